// CUDA-core helpers of the training path: on-device weight packing from the fp32 master
// parameters (PyTorch layouts) into the padded K-major operand layouts of the tensor-core kernels -- forward and
// transposed / rotated (dgrad) copies --, the inverse (packed fp32 weight gradients -> PyTorch layouts), PixelShuffle
// backward, bias gradients (column sums) and small elementwise pieces.  All memory-bound; none is on the
// inference path.
#include "ssr_device.cuh"

namespace ssr {

// output row n of a conv whose PixelShuffle(r) is folded into the store sits at source channel c*r*r + q (model.cu pack_conv)
__device__ __forceinline__ int ps_src_row(int n, int Cout, int ps_r) {
  if (ps_r <= 1) return n;
  const int rr = ps_r * ps_r, Cps = Cout / rr;
  const int q = n / Cps, c = n - q * Cps;
  return c * rr + q;
}

// W fp32 [Cout][Cin][taps] -> Wf bf16 [NP][taps*KP] (k = tap*KP + c), Wd bf16 [KP][taps*NP] (dgrad: rows = input
// channel, k' = tap'*NP + n with tap' = taps-1-tap, i.e. the 180-degree rotated kernel), bias -> bf [NP].
__global__ void pack_conv_dev_kernel(const float* __restrict__ W, const float* __restrict__ b, __nv_bfloat16* Wf, float* bf,
                                     __nv_bfloat16* Wd, int Cout, int Cin, int NP, int KP, int taps, int ps_r) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * Cin) return;
  const int n = idx / Cin, c = idx - n * Cin;
  const int sn = ps_src_row(n, Cout, ps_r);
  const float* src = W + ((size_t)sn * Cin + c) * taps;
  for (int t = 0; t < taps; ++t) {
    const __nv_bfloat16 v = __float2bfloat16_rn(src[t]);
    if (Wf) Wf[(size_t)n * taps * KP + (size_t)t * KP + c] = v;
    if (Wd) Wd[(size_t)c * taps * NP + (size_t)(taps - 1 - t) * NP + n] = v;
  }
  if (c == 0 && bf && b) bf[n] = b[sn];
}
int launch_pack_conv_dev(const float* W, const float* b, void* Wf, float* bf, void* Wd, int Cout, int Cin, int NP, int KP,
                         int taps, int ps_r, cudaStream_t s) {
  const int total = Cout * Cin;
  pack_conv_dev_kernel<<<(total + 255) / 256, 256, 0, s>>>(W, b, (__nv_bfloat16*)Wf, bf, (__nv_bfloat16*)Wd, Cout, Cin, NP, KP,
                                                           taps, ps_r);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// packed fp32 gradient dWp [NP][taps][KP] -> grad fp32 [Cout][Cin][taps] (PyTorch layout)
__global__ void unpack_wgrad_kernel(const float* __restrict__ dWp, float* grad, int Cout, int Cin, int KP, int taps, int ps_r) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * Cin) return;
  const int n = idx / Cin, c = idx - n * Cin;
  const int sn = ps_src_row(n, Cout, ps_r);
  float* dst = grad + ((size_t)sn * Cin + c) * taps;
  for (int t = 0; t < taps; ++t) dst[t] = dWp[((size_t)n * taps + t) * KP + c];
}
int launch_unpack_wgrad(const float* dWp, float* grad, int Cout, int Cin, int KP, int taps, int ps_r, cudaStream_t s) {
  const int total = Cout * Cin;
  unpack_wgrad_kernel<<<(total + 255) / 256, 256, 0, s>>>(dWp, grad, Cout, Cin, KP, taps, ps_r);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// bias gradient: out[sn(n)] = alpha * sum_m dY[m][n]; out must be zeroed (atomics over row strips)
__global__ void colsum_kernel(const void* __restrict__ dY, int elem, int ld, int M, int N, int Cout, int ps_r, float alpha,
                              float* out, int rows_per_block) {
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc = 0.0f;
    for (int r = r0; r < r1; ++r) acc += load_elem(dY, (size_t)r * ld + n, elem);
    atomicAdd(out + ps_src_row(n, Cout, ps_r), acc * alpha);
  }
}
int launch_colsum(const void* dY, int elem, int ld, int M, int N, int ps_r, float alpha, float* out, cudaStream_t s) {
  const int blocks = min((M + 63) / 64, 4 * 148);
  const int rpb = (M + blocks - 1) / blocks;
  SSR_CUDA(cudaMemsetAsync(out, 0, (size_t)N * 4, s));
  colsum_kernel<<<(M + rpb - 1) / rpb, 256, 0, s>>>(dY, elem, ld, M, N, N, ps_r, alpha, out, rpb);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// PixelShuffle backward: in bf16 [B][H*r][W*r][ld_in] (C channels used) -> out bf16 [B][H][W][r*r*C], column (i*r+j)*C + c
__global__ void unshuffle_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* out, int B, int H, int W, int C, int r,
                                 int ld_in) {
  const int vec_per_px = r * r * C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * vec_per_px) return;
  const int v = (int)(idx % vec_per_px);
  const long long p = idx / vec_per_px;
  const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
  const int col = v * 8, q = col / C, c = col - q * C;
  const int i = q / r, j = q - i * r;
  const size_t src = (((size_t)b * (H * r) + (y * r + i)) * (W * r) + (x * r + j)) * ld_in + c;
  *reinterpret_cast<uint4*>(out + (size_t)idx * 8) = *reinterpret_cast<const uint4*>(in + src);
}
int launch_unshuffle(const void* in, void* out, int B, int H, int W, int C, int r, int ld_in, cudaStream_t s) {
  SSR_CHECK(C % 8 == 0 && ld_in % 8 == 0, SSR_E_INVALID, "unshuffle: C=%d ld=%d", C, ld_in);
  const long long total = (long long)B * H * W * (r * r * C / 8);
  unshuffle_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, B, H, W, C, r, ld_in);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// fp32 NCHW [B,3,H,W] -> bf16 NHWC [B,H,W,64]: (x * scale + shift[c]) in lanes 0..2, zeros elsewhere
__global__ void nchw3_to_nhwc64_kernel(const float* __restrict__ in, __nv_bfloat16* out, int B, int H, int W, float scale,
                                       float s0, float s1, float s2) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)B * H * W) return;
  const size_t plane = (size_t)H * W;
  const size_t b = p / plane, rem = p - b * plane;
  const float* src = in + b * 3 * plane + rem;
  uint4 z = make_uint4(0, 0, 0, 0);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)p * 64);
  uint4 first = z;
  first.x = pack_bf16x2(src[0] * scale + s0, src[plane] * scale + s1);
  first.y = pack_bf16x2(src[2 * plane] * scale + s2, 0.0f);
  dst[0] = first;
#pragma unroll
  for (int i = 1; i < 8; ++i) dst[i] = z;
}
int launch_nchw3_to_nhwc64(const float* in, void* out, int B, int H, int W, float scale, const float* shift3, cudaStream_t s) {
  const long long total = (long long)B * H * W;
  nchw3_to_nhwc64_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(in, (__nv_bfloat16*)out, B, H, W, scale, shift3[0], shift3[1],
                                                                    shift3[2]);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// a += b (fp32), and an optional bf16 copy of the sum
__global__ void add_inplace_kernel(float* a, const float* __restrict__ b, __nv_bfloat16* out_bf, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 x = reinterpret_cast<float4*>(a)[i];
  const float4 y = reinterpret_cast<const float4*>(b)[i];
  x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
  reinterpret_cast<float4*>(a)[i] = x;
  if (out_bf) reinterpret_cast<uint2*>(out_bf)[i] = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
}
int launch_add_inplace(float* a, const float* b, void* out_bf, size_t n, cudaStream_t s) {
  SSR_CHECK(n % 4 == 0, SSR_E_INVALID, "add_inplace: n %% 4");
  add_inplace_kernel<<<(int)((n / 4 + 255) / 256), 256, 0, s>>>(a, b, (__nv_bfloat16*)out_bf, n / 4);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
