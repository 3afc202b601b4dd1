// CUDA-core helpers of the training path: on-device weight packing from the fp32 master
// parameters (PyTorch layouts) into the padded K-major operand layouts of the tensor-core kernels -- forward and
// transposed / rotated (dgrad) copies --, the inverse (packed fp32 weight gradients -> PyTorch layouts), PixelShuffle
// backward, bias gradients (column sums) and small elementwise pieces.  All memory-bound; none is on the
// inference path.
#include <string.h>

#include <algorithm>

#include "ssr_device.cuh"

namespace ssr {

// output row n of a conv whose PixelShuffle(r) is folded into the store sits at source channel c*r*r + q (model.cu pack_conv)
__device__ __forceinline__ int ps_src_row(int n, int Cout, int ps_r) {
  if (ps_r <= 1) return n;
  const int rr = ps_r * ps_r, Cps = Cout / rr;
  const int q = n / Cps, c = n - q * Cps;
  return c * rr + q;
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

// Column sums run in two stages -- per-strip partial sums, then one thread per column adds the strips -- so that no
// address sees hundreds of same-address atomics (measured: the single-stage atomic version cost 5x the data movement) and
// the result is deterministic.  `partial` is scratch of at least kColsumStrips * N floats.
constexpr int kColsumStrips = 592;
// stage 1: bf16 [M][ld] -> partial[strip][NP]; a thread owns 8 consecutive columns (one 16-byte load per row), RY row
// lanes of a CTA walk the strip's rows in parallel and meet in shared memory
__global__ void __launch_bounds__(256) colsum_partial_kernel(const __nv_bfloat16* __restrict__ dY, int ld, int M, int NP, float* partial,
                                                             int rows_per_block, int VX, int RY) {
  extern __shared__ float cs_red[];  // [RY][NP]
  const int V = NP / 8;
  const int vx = threadIdx.x % VX, ry = threadIdx.x / VX;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  if (ry < RY) {
    for (int v = vx; v < V; v += VX) {
      float acc[8], acc2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = acc2[e] = 0.0f;
      const __nv_bfloat16* base = dY + v * 8;
      int r = r0 + ry;
      for (; r + RY < r1; r += 2 * RY) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r * ld));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(r + RY) * ld));
        const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += __uint_as_float(wa[e] << 16);
          acc[2 * e + 1] += __uint_as_float(wa[e] & 0xffff0000u);
          acc2[2 * e] += __uint_as_float(wb[e] << 16);
          acc2[2 * e + 1] += __uint_as_float(wb[e] & 0xffff0000u);
        }
      }
      if (r < r1) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r * ld));
        const uint32_t wa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += __uint_as_float(wa[e] << 16);
          acc[2 * e + 1] += __uint_as_float(wa[e] & 0xffff0000u);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) cs_red[(size_t)ry * NP + v * 8 + e] = acc[e] + acc2[e];
    }
  }
  __syncthreads();
  for (int n = threadIdx.x; n < NP; n += blockDim.x) {
    float t = 0.0f;
    for (int y = 0; y < RY; ++y) t += cs_red[(size_t)y * NP + n];
    partial[(size_t)blockIdx.x * NP + n] = t;
  }
}
// second stage of the two-stage reductions: 32 columns x 8 strip lanes per CTA, 4 independent accumulators per thread so
// that 32 strip loads are in flight per column (one thread walking 592 strips serially took 42 us per launch)
__device__ __forceinline__ float strip_sum(const float* __restrict__ col, int strips, size_t stride, float (*red)[33]) {
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
  int b = threadIdx.y;
  for (; b + 24 < strips; b += 32) {
    a0 += col[(size_t)b * stride];
    a1 += col[(size_t)(b + 8) * stride];
    a2 += col[(size_t)(b + 16) * stride];
    a3 += col[(size_t)(b + 24) * stride];
  }
  for (; b < strips; b += 8) a0 += col[(size_t)b * stride];
  red[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.y == 0) {
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
  }
  return t;
}
// bias gradient of a conv: out[sn(n)] = alpha * sum_m dY[m][n]
__global__ void __launch_bounds__(256) colsum_final_conv_kernel(const float* __restrict__ partial, int strips, int NP, int Cout,
                                                                int ps_r, float alpha, float* out) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const float t = strip_sum(partial + min(n, NP - 1), strips, (size_t)NP, red);
  if (threadIdx.y == 0 && n < Cout) out[ps_src_row(n, Cout, ps_r)] = t * alpha;
}
static int colsum_stage1(const void* dY, int elem, int ld, int M, int NP, float* partial, int* strips, cudaStream_t s) {
  SSR_CHECK(elem == 2 && NP % 8 == 0 && ld % 8 == 0 && NP <= 2304, SSR_E_INVALID, "colsum: bf16 rows with NP %% 8 == 0 (NP=%d ld=%d)", NP, ld);
  const int blocks = min((M + 63) / 64, kColsumStrips);
  const int rpb = (M + blocks - 1) / blocks;
  *strips = (M + rpb - 1) / rpb;
  const int V = NP / 8, VX = V < 256 ? V : 256, RY = 256 / VX;
  ProfScope prof("bias_grad_colsum", 0.0, (double)M * NP * elem, s);
  colsum_partial_kernel<<<*strips, 256, (size_t)RY * NP * 4, s>>>((const __nv_bfloat16*)dY, ld, M, NP, partial, rpb, VX, RY);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}
// dY: bf16 [M][ld] of packed width NP (multiple of 8); out[sn(n)] for n < Cout
static float* red_take(DeferredRed* dr, size_t floats) {
  if (!dr || dr->used + floats > dr->pool_floats || dr->n >= dr->cap) return nullptr;
  float* p = dr->pool + dr->used;
  dr->used += (floats + 63) & ~(size_t)63;
  return p;
}
int launch_colsum(const void* dY, int elem, int ld, int M, int NP, int Cout, int ps_r, float alpha, float* out, float* partial,
                  cudaStream_t s, DeferredRed* dr) {
  int strips;
  float* mine = red_take(dr, (size_t)kColsumStrips * NP);
  SSR_TRY(colsum_stage1(dY, elem, ld, M, NP, mine ? mine : partial, &strips, s));
  if (mine) {
    RedEntry& e = dr->host[dr->n++];
    memset(&e, 0, sizeof(e));
    e.partial = mine; e.out = out; e.strips = strips; e.stride = NP; e.N = Cout; e.mode = 0; e.Cout = Cout; e.ps_r = ps_r; e.alpha = alpha;
    return SSR_OK;
  }
  colsum_final_conv_kernel<<<(Cout + 31) / 32, dim3(32, 8), 0, s>>>(partial, strips, NP, Cout, ps_r, alpha, out);
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

// ---------------------------------------------------------------------------------------------
// nn.Linear maps (model.cu finalize_swinir): source (n, k) -> packed (np, kp) and the row scale
__device__ __forceinline__ void lin_map(const LinMap& m, int n, int k, int* np, int* kp, float* scale) {
  *np = n;
  *kp = k;
  *scale = 1.0f;
  if (m.mode == 1) {
    const int part = n / m.C, r = n - part * m.C, h = r / m.d, j = r - h * m.d;
    *np = part * m.QP + h * m.DP + j;
    if (part == 0) *scale = m.qscale;
  } else if (m.mode == 2) {
    const int h = k / m.d, j = k - h * m.d;
    *kp = h * m.DP + j;
  }
}

// bias gradient of a linear layer: out[n] = scale(n) * sum_m dY[m][np(n)]  (NP = packed width of dY)
__global__ void __launch_bounds__(256) colsum_final_map_kernel(const float* __restrict__ partial, int strips, int NP, int N,
                                                               const LinMap map, float* out) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  int np = 0, kp;
  float sc = 0.0f;
  if (n < N) lin_map(map, n, 0, &np, &kp, &sc);
  const float t = strip_sum(partial + np, strips, (size_t)NP, red);
  if (threadIdx.y == 0 && n < N) out[n] = t * sc;
}
int launch_colsum_map(const void* dY, int elem, int ld, int M, int NP, int N, const LinMap& map, float* out, float* partial,
                      cudaStream_t s, DeferredRed* dr) {
  int strips;
  float* mine = red_take(dr, (size_t)kColsumStrips * NP);
  SSR_TRY(colsum_stage1(dY, elem, ld, M, NP, mine ? mine : partial, &strips, s));
  if (mine) {
    RedEntry& e = dr->host[dr->n++];
    memset(&e, 0, sizeof(e));
    e.partial = mine; e.out = out; e.strips = strips; e.stride = NP; e.N = N; e.mode = 1; e.map = map; e.alpha = 1.0f;
    return SSR_OK;
  }
  colsum_final_map_kernel<<<(N + 31) / 32, dim3(32, 8), 0, s>>>(partial, strips, NP, N, map, out);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// all deferred second stages in one launch: blockIdx.y = entry, blockIdx.x = block of 32 output columns
__global__ void __launch_bounds__(256) deferred_reductions_kernel(const RedEntry* __restrict__ entries) {
  __shared__ float red[8][33];
  const RedEntry e = entries[blockIdx.y];
  if ((int)blockIdx.x * 32 >= e.N) return;
  const int n = blockIdx.x * 32 + threadIdx.x;
  int col = 0, dst = 0;
  float sc = 0.0f;
  if (n < e.N) {
    if (e.mode == 1) {
      int kp;
      lin_map(e.map, n, 0, &col, &kp, &sc);
      dst = n;
    } else {
      col = n;
      dst = e.mode == 0 ? ps_src_row(n, e.Cout, e.ps_r) : n;
      sc = e.mode == 0 ? e.alpha : 1.0f;
    }
  }
  const float t = strip_sum(e.partial + col, e.strips, (size_t)e.stride, red);
  if (threadIdx.y == 0 && n < e.N) e.out[dst] = t * sc;
}
int launch_deferred_reductions(DeferredRed* dr, cudaStream_t s) {
  if (!dr || dr->n == 0) return SSR_OK;
  int maxN = 0;
  for (int i = 0; i < dr->n; ++i) maxN = dr->host[i].N > maxN ? dr->host[i].N : maxN;
  SSR_CUDA(cudaMemcpyAsync(dr->dev, dr->host, (size_t)dr->n * sizeof(RedEntry), cudaMemcpyHostToDevice, s));
  ProfScope prof("deferred_reductions", 0.0, 0.0, s);
  deferred_reductions_kernel<<<dim3((maxN + 31) / 32, dr->n), dim3(32, 8), 0, s>>>(dr->dev);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  dr->n = 0;
  dr->used = 0;
  return SSR_OK;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row (rows strided over all warps so dgamma / dbeta partials live in registers):
//   xhat = (x - mean) rstd;  g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
__global__ void __launch_bounds__(256, 4) ln_bwd_kernel(const LnBwdArgs a) {
  __shared__ float red[3][8][256];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * 8 + wib, nwarps = gridDim.x * 8;
  // a lane owns the float4 slots `lane` and `lane + 32` of a row: columns col(j) = 4 * (lane + 32 * (j / 4)) + j % 4
  const int Q = a.CP / 4;
  const bool has1 = lane + 32 < Q;
  float dg[8], db[8], ds[8], gam[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dg[j] = db[j] = ds[j] = 0.0f;
    const int n = 4 * (lane + 32 * (j >> 2)) + (j & 3);
    gam[j] = n < a.C ? __ldg(a.gamma + n) : 0.0f;
  }
  const float invC = 1.0f / (float)a.C;
  for (int row = warp; row < a.M; row += nwarps) {
    float v[8], dy[8], gi[8];
    {
      const float4* x4 = reinterpret_cast<const float4*>(a.x + (size_t)row * a.ldx);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 x0 = lane < Q ? x4[lane] : z, x1 = has1 ? x4[lane + 32] : z;
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
      if (a.elem_dy == 2) {
        const uint2* d2 = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(a.dy) + (size_t)row * a.ld_dy);
        const uint2 zz = make_uint2(0, 0);
        const uint2 d0 = lane < Q ? d2[lane] : zz, d1 = has1 ? d2[lane + 32] : zz;
        dy[0] = __uint_as_float(d0.x << 16); dy[1] = __uint_as_float(d0.x & 0xffff0000u);
        dy[2] = __uint_as_float(d0.y << 16); dy[3] = __uint_as_float(d0.y & 0xffff0000u);
        dy[4] = __uint_as_float(d1.x << 16); dy[5] = __uint_as_float(d1.x & 0xffff0000u);
        dy[6] = __uint_as_float(d1.y << 16); dy[7] = __uint_as_float(d1.y & 0xffff0000u);
      } else {
        const float4* d4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.dy) + (size_t)row * a.ld_dy);
        const float4 d0 = lane < Q ? d4[lane] : z, d1 = has1 ? d4[lane + 32] : z;
        dy[0] = d0.x; dy[1] = d0.y; dy[2] = d0.z; dy[3] = d0.w; dy[4] = d1.x; dy[5] = d1.y; dy[6] = d1.z; dy[7] = d1.w;
      }
      if (a.Gin) {
        const float4* g4 = reinterpret_cast<const float4*>(a.Gin + (size_t)row * a.ldg);
        const float4 g0 = lane < Q ? g4[lane] : z, g1 = has1 ? g4[lane + 32] : z;
        gi[0] = g0.x; gi[1] = g0.y; gi[2] = g0.z; gi[3] = g0.w; gi[4] = g1.x; gi[5] = g1.y; gi[6] = g1.z; gi[7] = g1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) gi[j] = 0.0f;
      }
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = 4 * (lane + 32 * (j >> 2)) + (j & 3);
      if (n >= a.C) v[j] = dy[j] = gi[j] = 0.0f;
      s += v[j];
    }
    const float mean = warp_sum(s) * invC;
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = 4 * (lane + 32 * (j >> 2)) + (j & 3);
      if (n < a.C) q += (v[j] - mean) * (v[j] - mean);
    }
    const float rstd = rsqrtf(warp_sum(q) * invC + a.eps);
    float sg = 0.0f, sgx = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = 4 * (lane + 32 * (j >> 2)) + (j & 3);
      v[j] = n < a.C ? (v[j] - mean) * rstd : 0.0f;  // xhat
      dg[j] = fmaf(dy[j], v[j], dg[j]);
      db[j] += dy[j];
      dy[j] *= gam[j];
      sg += dy[j];
      sgx = fmaf(dy[j], v[j], sgx);
    }
    sg = warp_sum(sg) * invC;
    sgx = warp_sum(sgx) * invC;
    float dx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = 4 * (lane + 32 * (j >> 2)) + (j & 3);
      dx[j] = n < a.C ? rstd * (dy[j] - sg - v[j] * sgx) + gi[j] : 0.0f;
    }
    float4* o4 = reinterpret_cast<float4*>(a.Gout + (size_t)row * a.ldg);
    if (lane < Q) o4[lane] = make_float4(dx[0], dx[1], dx[2], dx[3]);
    if (has1) o4[lane + 32] = make_float4(dx[4], dx[5], dx[6], dx[7]);
    if (a.Gb) {
      if (a.gb_scale) {
        const float sc = __ldg(a.gb_scale + row / a.rows_per_scale);
#pragma unroll
        for (int j = 0; j < 8; ++j) dx[j] *= sc;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ds[j] += dx[j];  // column sums of Gb: the consumer Linear layer's bias gradient
      uint2* b2 = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.Gb) + (size_t)row * a.ldg);
      if (lane < Q) b2[lane] = make_uint2(pack_bf16x2(dx[0], dx[1]), pack_bf16x2(dx[2], dx[3]));
      if (has1) b2[lane + 32] = make_uint2(pack_bf16x2(dx[4], dx[5]), pack_bf16x2(dx[6], dx[7]));
    }
  }
  if (!a.dgamma) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = 4 * (lane + 32 * (j >> 2)) + (j & 3);
    red[0][wib][n] = dg[j];
    red[1][wib][n] = db[j];
    red[2][wib][n] = ds[j];
  }
  __syncthreads();
  for (int n = threadIdx.x; n < a.C; n += 256) {
    float g = 0.0f, b = 0.0f, c = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      g += red[0][w][n];
      b += red[1][w][n];
      c += red[2][w][n];
    }
    const int nacc = a.gb_colsum ? 3 : 2;  // strips: [block][nacc][C]
    a.partial[((size_t)blockIdx.x * nacc) * a.C + n] = g;
    a.partial[((size_t)blockIdx.x * nacc + 1) * a.C + n] = b;
    if (a.gb_colsum) a.partial[((size_t)blockIdx.x * nacc + 2) * a.C + n] = c;
  }
}
__global__ void __launch_bounds__(256) ln_bwd_final_kernel(const float* __restrict__ partial, int blocks, int C, float* dgamma,
                                                           float* dbeta) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;  // n in [0, 2C): dgamma columns, then dbeta columns
  const int which = n >= C ? 1 : 0, c = min(n - which * C, C - 1);
  const float t = strip_sum(partial + (size_t)which * C + c, blocks, (size_t)2 * C, red);
  if (threadIdx.y == 0 && n < 2 * C) (which ? dbeta : dgamma)[c] = t;
}
int launch_ln_bwd(const LnBwdArgs& a0, cudaStream_t s, DeferredRed* dr) {
  LnBwdArgs a = a0;
  SSR_CHECK(a.C <= 256 && a.CP <= 256 && a.CP % 4 == 0 && a.ldx % 4 == 0 && a.ldg % 4 == 0 && a.ld_dy % 4 == 0, SSR_E_INVALID,
            "ln_bwd: C=%d CP=%d", a.C, a.CP);
  SSR_CHECK(!a.dgamma || a.partial, SSR_E_INVALID, "ln_bwd: partial scratch missing");
  const int blocks = min((a.M + 7) / 8, kColsumStrips);
  const int nacc = a.gb_colsum ? 3 : 2;
  float* mine = a.dgamma && dr && dr->n + nacc <= dr->cap ? red_take(dr, (size_t)blocks * nacc * a.C) : nullptr;
  if (mine) a.partial = mine;
  SSR_CHECK(!a.gb_colsum || (mine && a.Gb), SSR_E_INVALID, "ln_bwd: gb_colsum needs Gb, dgamma and a deferred-reduction pool");
  ProfScope prof("ln_bwd", 0.0, (double)a.M * a.C * (4 + a.elem_dy + (a.Gin ? 4 : 0) + 4 + (a.Gb ? 2 : 0)), s);
  ln_bwd_kernel<<<blocks, 256, 0, s>>>(a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  if (mine) {
    for (int which = 0; which < nacc; ++which) {
      RedEntry& e = dr->host[dr->n++];
      memset(&e, 0, sizeof(e));
      e.partial = mine + (size_t)which * a.C; e.out = which == 0 ? a.dgamma : which == 1 ? a.dbeta : a.gb_colsum; e.strips = blocks;
      e.stride = nacc * a.C; e.N = a.C;
      e.mode = 2; e.alpha = 1.0f;
    }
  } else if (a.dgamma) {
    ln_bwd_final_kernel<<<(2 * a.C + 31) / 32, dim3(32, 8), 0, s>>>(a.partial, blocks, a.C, a.dgamma, a.dbeta);
    count_launch();
    SSR_CUDA(cudaGetLastError());
  }
  return SSR_OK;
}

// network input for the first conv's weight gradient: pad (training flavour: reflect, common.py:277-282) + normalise
// fp32 NCHW [B,3,h,w] -> bf16 NHWC [B,Hp,Wp,64], lanes 0..2 = x * scale + shift[c], rest zero
__global__ void input_nhwc64_kernel(const float* __restrict__ in, __nv_bfloat16* out, int B, int h, int w, int Hp, int Wp,
                                    float scale, float s0, float s1, float s2) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)B * Hp * Wp) return;
  const int x = (int)(p % Wp), y = (int)((p / Wp) % Hp), b = (int)(p / ((long long)Wp * Hp));
  const int sy = y < h ? y : 2 * (h - 1) - y, sx = x < w ? x : 2 * (w - 1) - x;
  const size_t plane = (size_t)h * w;
  const float* src = in + (size_t)b * 3 * plane + (size_t)sy * w + sx;
  uint4 z = make_uint4(0, 0, 0, 0);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)p * 64);
  uint4 first = z;
  first.x = pack_bf16x2(src[0] * scale + s0, src[plane] * scale + s1);
  first.y = pack_bf16x2(src[2 * plane] * scale + s2, 0.0f);
  dst[0] = first;
#pragma unroll
  for (int i = 1; i < 8; ++i) dst[i] = z;
}
int launch_input_nhwc64(const float* x, void* out, int B, int h, int w, int Hp, int Wp, float scale, const float* shift3,
                        cudaStream_t s) {
  const long long total = (long long)B * Hp * Wp;
  input_nhwc64_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(x, (__nv_bfloat16*)out, B, h, w, Hp, Wp, scale, shift3[0],
                                                                 shift3[1], shift3[2]);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// dL/dx of the first conv (3 -> C, zero padding) composed with the input normalisation and the training-mode reflect pad
// (the adjoints of conv_first_kernel / input_nhwc64_kernel):
//   dxp[b][ci][y][x] = sum_{ky,kx,n} G[b][y-ky+1][x-kx+1][n] * W[n][ci][ky][kx]      on the padded Hp x Wp grid
//   dx[b][ci][sy][sx] += scale * dxp[b][ci][y][x]   for every padded (y, x) that reads source pixel (sy, sx)
// One warp per padded pixel, lanes over the channels; the <= 4 padded pixels of a reflected source pixel meet in an atomicAdd.
__global__ void __launch_bounds__(256) conv_first_dgrad_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ Wc, int C,
                                                               int B, int h, int w, int Hp, int Wp, float scale, float* __restrict__ dx) {
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= (long long)B * Hp * Wp) return;
  const int x = (int)(p % Wp), y = (int)((p / Wp) % Hp), b = (int)(p / ((long long)Wp * Hp));
  float acc[3] = {0.0f, 0.0f, 0.0f};
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y - ky + 1;
    if (yy < 0 || yy >= Hp) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x - kx + 1;
      if (xx < 0 || xx >= Wp) continue;
      const float* g = G + ((size_t)((size_t)b * Hp + yy) * Wp + xx) * ldg;
      for (int n = lane; n < C; n += 32) {
        const float gv = g[n];
        const float* wn = Wc + (size_t)n * 27 + ky * 3 + kx;
        acc[0] = fmaf(gv, wn[0], acc[0]);
        acc[1] = fmaf(gv, wn[9], acc[1]);
        acc[2] = fmaf(gv, wn[18], acc[2]);
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[ci] += __shfl_xor_sync(0xffffffffu, acc[ci], o);
  if (lane < 3) {
    const int sy = y < h ? y : 2 * (h - 1) - y, sx = x < w ? x : 2 * (w - 1) - x;
    atomicAdd(dx + ((size_t)b * 3 + lane) * h * w + (size_t)sy * w + sx, scale * acc[lane]);
  }
}
int launch_conv_first_dgrad(const float* G, int ldg, const float* Wc, int C, int B, int h, int w, int Hp, int Wp, float scale,
                            float* dx, cudaStream_t s) {
  SSR_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * 3 * h * w * 4, s));
  const long long warps = (long long)B * Hp * Wp;
  conv_first_dgrad_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(G, ldg, Wc, C, B, h, w, Hp, Wp, scale, dx);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// output gradient fp32 NCHW [B,3,ch,cw] (the cropped image) -> bf16 NHWC [B,Hs,Ws,64] of the un-cropped reconstruction
// conv output: lanes 0..2 = dy * scale inside the crop, zero outside and in the pad lanes
__global__ void grad_nhwc64_kernel(const float* __restrict__ dy, __nv_bfloat16* out, int B, int ch, int cw, int Hs, int Ws,
                                   float scale) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)B * Hs * Ws) return;
  const int x = (int)(p % Ws), y = (int)((p / Ws) % Hs), b = (int)(p / ((long long)Ws * Hs));
  uint4 z = make_uint4(0, 0, 0, 0);
  uint4 first = z;
  if (y < ch && x < cw) {
    const size_t plane = (size_t)ch * cw;
    const float* src = dy + (size_t)b * 3 * plane + (size_t)y * cw + x;
    first.x = pack_bf16x2(src[0] * scale, src[plane] * scale);
    first.y = pack_bf16x2(src[2 * plane] * scale, 0.0f);
  }
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)p * 64);
  dst[0] = first;
#pragma unroll
  for (int i = 1; i < 8; ++i) dst[i] = z;
}
int launch_grad_nhwc64(const float* dy, void* out, int B, int ch, int cw, int Hs, int Ws, float scale, cudaStream_t s) {
  const long long total = (long long)B * Hs * Ws;
  grad_nhwc64_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(dy, (__nv_bfloat16*)out, B, ch, cw, Hs, Ws, scale);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

__global__ void scale_to_bf16_kernel(const float* __restrict__ in, const float* __restrict__ scale, size_t elems_per_scale,
                                     __nv_bfloat16* out, size_t n4) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float sc = __ldg(scale + (i * 4) / elems_per_scale);
  const float4 x = reinterpret_cast<const float4*>(in)[i];
  reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(x.x * sc, x.y * sc), pack_bf16x2(x.z * sc, x.w * sc));
}
int launch_scale_to_bf16(const float* in, const float* scale, size_t elems_per_scale, void* out, size_t n, cudaStream_t s) {
  SSR_CHECK(n % 4 == 0 && elems_per_scale % 4 == 0, SSR_E_INVALID, "scale_to_bf16: n %% 4");
  scale_to_bf16_kernel<<<(int)((n / 4 + 255) / 256), 256, 0, s>>>(in, scale, elems_per_scale, (__nv_bfloat16*)out, n / 4);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// ---------------------------------------------------------------------------------------------
// batched re-pack: blockIdx.y = table entry, blockIdx.z = pass (0: forward pack, source-order threads; 1: dgrad pack,
// transposed thread order so that its stores are coalesced too), blockIdx.x = block of 256 elements
__global__ void __launch_bounds__(256) pack_batched_kernel(const PackEntry* __restrict__ entries) {
  const PackEntry e = entries[blockIdx.y];
  const int pass = blockIdx.z;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (e.kind == 2) {  // plain copy
    if (pass == 0 && idx < e.N) reinterpret_cast<float*>(e.Wf)[idx] = e.W[idx];
    return;
  }
  if (e.kind == 3) {  // [N][K] -> [K][N]
    if (pass == 0 && idx < e.N * e.K) {
      const int k = idx / e.N, n = idx - k * e.N;
      reinterpret_cast<float*>(e.Wf)[idx] = e.W[(size_t)n * e.K + k];
    }
    return;
  }
  if (idx >= e.N * e.K) return;
  if (pass == 1 && !e.Wd) return;
  if (e.kind == 0) {
    if (pass == 0) {
      const int n = idx / e.K, c = idx - n * e.K;
      const int sn = ps_src_row(n, e.N, e.ps_r);
      const float* src = e.W + ((size_t)sn * e.K + c) * e.taps;
      __nv_bfloat16* Wf = reinterpret_cast<__nv_bfloat16*>(e.Wf);
      for (int t = 0; t < e.taps; ++t) Wf[(size_t)n * e.taps * e.KP + (size_t)t * e.KP + c] = __float2bfloat16_rn(src[t]);
      if (c == 0 && e.bf && e.b) e.bf[n] = e.b[sn];
    } else {
      const int c = idx / e.N, n = idx - c * e.N;
      const int sn = ps_src_row(n, e.N, e.ps_r);
      const float* src = e.W + ((size_t)sn * e.K + c) * e.taps;
      __nv_bfloat16* Wd = reinterpret_cast<__nv_bfloat16*>(e.Wd);
      for (int t = 0; t < e.taps; ++t) Wd[(size_t)c * e.taps * e.NP + (size_t)(e.taps - 1 - t) * e.NP + n] = __float2bfloat16_rn(src[t]);
    }
    return;
  }
  // linear
  const int n = pass ? idx % e.N : idx / e.K, k = pass ? idx / e.N : idx % e.K;
  int np, kp;
  float sc;
  lin_map(e.map, n, k, &np, &kp, &sc);
  const __nv_bfloat16 v = __float2bfloat16_rn(e.W[(size_t)n * e.K + k] * sc);
  if (!pass) {
    reinterpret_cast<__nv_bfloat16*>(e.Wf)[(size_t)np * e.KP + kp] = v;
    if (k == 0 && e.bf && e.b) e.bf[np] = e.b[n] * sc;
  } else {
    reinterpret_cast<__nv_bfloat16*>(e.Wd)[(size_t)kp * e.NP + np] = v;
  }
}
int launch_pack_batched(const PackEntry* host, PackEntry* dev, int n, cudaStream_t s) {
  if (n == 0) return SSR_OK;
  long long max_elems = 1;
  for (int i = 0; i < n; ++i) {
    const long long el = host[i].kind == 2 ? host[i].N : (long long)host[i].N * host[i].K;
    max_elems = el > max_elems ? el : max_elems;
  }
  SSR_CUDA(cudaMemcpyAsync(dev, host, (size_t)n * sizeof(PackEntry), cudaMemcpyHostToDevice, s));
  ProfScope prof("weight_repack", 0.0, 0.0, s);
  pack_batched_kernel<<<dim3((unsigned)((max_elems + 255) / 256), n, 2), 256, 0, s>>>(dev);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

__global__ void __launch_bounds__(256) unpack_batched_kernel(const PackEntry* __restrict__ entries) {
  const PackEntry e = entries[blockIdx.y];
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= e.N * e.K) return;
  const int n = idx / e.K, k = idx - n * e.K;
  float* grad = reinterpret_cast<float*>(e.Wf);
  float* gbias = const_cast<float*>(e.b);  // destination of the bias gradient (its packed source is e.bf), or null
  if (e.kind == 0) {
    const int sn = ps_src_row(n, e.N, e.ps_r);
    float* dst = grad + ((size_t)sn * e.K + k) * e.taps;
    for (int t = 0; t < e.taps; ++t) dst[t] = e.W[((size_t)n * e.taps + t) * e.KP + k];
    if (k == 0 && gbias && e.bf) gbias[sn] = e.bf[n];
  } else {
    int np, kp;
    float sc;
    lin_map(e.map, n, k, &np, &kp, &sc);
    grad[idx] = e.W[(size_t)np * e.KP + kp] * sc;
    if (k == 0 && gbias && e.bf) gbias[n] = e.bf[np] * sc;
  }
}
int launch_unpack_batched(const PackEntry* host, PackEntry* dev, int n, cudaStream_t s) {
  if (n == 0) return SSR_OK;
  long long max_elems = 1;
  for (int i = 0; i < n; ++i) max_elems = std::max(max_elems, (long long)host[i].N * host[i].K);
  SSR_CUDA(cudaMemcpyAsync(dev, host, (size_t)n * sizeof(PackEntry), cudaMemcpyHostToDevice, s));
  ProfScope prof("wgrad_unpack", 0.0, 0.0, s);
  unpack_batched_kernel<<<dim3((unsigned)((max_elems + 255) / 256), n), 256, 0, s>>>(dev);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
