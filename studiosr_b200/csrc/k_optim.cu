// The rest of the Trainer step (trainer.py:102-109) on flat fp32 buffers: the L1 loss with its backward seed, and Adam.
//
//   loss = mean |out - y| ;  d loss / d out = sign(out - y) / N          (nn.L1Loss, trainer.py:45,102-104)
//   Adam (torch.optim.Adam, trainer.py:133-139), with the roundings of torch's single-tensor CUDA path (foreach=False) so that the
//   update is bit-identical to it: lerp_ as one fma, mul_ then addcmul_ (fma over a rounded product), sqrt, the division by the
//   scalar sqrt(1 - b2^t) as a multiplication by its fp32 reciprocal (ATen's cpu-scalar divisor shortcut), + eps, addcdiv_ as an
//   IEEE division followed by one fma:
//     m <- fma(1 - b1, g - m, m) ;  v <- fma((1 - b2) g, g, b2 v) ;  p <- fma(-lr / (1 - b1^t), m / (sqrt(v) * rbc2 + eps), p)
//
// Both are pure HBM streams: L1 reads 8 B and writes 4 B per output element, Adam moves 28 B per parameter (read p, g, m, v;
// write p, m, v) in ONE launch over the flat buffers the native backward already fills (models/common.py) -- against ~530
// per-tensor launches (or 3-4 multi-tensor launches plus their pointer tables) of the stock optimizer.
#include "ssr_device.cuh"

namespace ssr {

constexpr int L1_THREADS = 256;
constexpr int L1_MAX_BLOCKS = 148 * 8;

// stage 1: per-block partial sums of |out - y| (deterministic: fixed grid, fixed order) + the gradient seed
__global__ void __launch_bounds__(L1_THREADS) l1_partial_kernel(const float* __restrict__ out, const float* __restrict__ y,
                                                                float* __restrict__ dout, double* __restrict__ partial, long long n,
                                                                float inv_n) {
  __shared__ double red[L1_THREADS / 32];
  double acc = 0.0;
  const long long n4 = n >> 2;
  const float4* o4 = reinterpret_cast<const float4*>(out);
  const float4* y4 = reinterpret_cast<const float4*>(y);
  float4* d4 = reinterpret_cast<float4*>(dout);
  auto sgn = [inv_n](float d) { return d > 0.0f ? inv_n : (d < 0.0f ? -inv_n : 0.0f); };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(o4 + i), b = __ldg(y4 + i);
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
    acc += (double)(fabsf(dx) + fabsf(dy)) + (double)(fabsf(dz) + fabsf(dw));
    if (dout) d4[i] = make_float4(sgn(dx), sgn(dy), sgn(dz), sgn(dw));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // ragged tail
    const long long i = (n4 << 2) + threadIdx.x;
    const float d = out[i] - y[i];
    acc += (double)fabsf(d);
    if (dout) dout[i] = sgn(d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < L1_THREADS / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}
// stage 2: one warp adds the partials in a fixed order
__global__ void l1_final_kernel(const double* __restrict__ partial, int nblocks, float* __restrict__ loss, double inv_n) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *loss = (float)(s * inv_n);
}

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, float w1, float beta2, float w2,
                                                        float eps, float weight_decay, float neg_step_size, float inv_bc2_sqrt,
                                                        float grad_scale) {
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto one = [&](float& pp, float gg, float& mm, float& vv) {
    gg = __fmul_rn(gg, grad_scale);
    if (weight_decay != 0.0f) gg = __fmaf_rn(weight_decay, pp, gg);  // L2 form of torch.optim.Adam: grad.add(param, alpha=wd)
    mm = __fmaf_rn(w1, __fsub_rn(gg, mm), mm);                        // exp_avg.lerp_(grad, 1 - b1), |weight| < 0.5 branch
    vv = __fmaf_rn(__fmul_rn(w2, gg), gg, __fmul_rn(vv, beta2));      // exp_avg_sq.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(vv), inv_bc2_sqrt), eps);
    pp = __fmaf_rn(neg_step_size, __fdiv_rn(mm, denom), pp);          // param.addcdiv_(exp_avg, denom, value=-step_size)
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = p4[i], mv = m4[i], vv = v4[i];
    const float4 gv = __ldg(g4 + i);
    one(pv.x, gv.x, mv.x, vv.x);
    one(pv.y, gv.y, mv.y, vv.y);
    one(pv.z, gv.z, mv.z, vv.z);
    one(pv.w, gv.w, mv.w, vv.w);
    p4[i] = pv;
    m4[i] = mv;
    v4[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float pv = p[i], mv = m[i], vv = v[i];
    one(pv, g[i], mv, vv);
    p[i] = pv;
    m[i] = mv;
    v[i] = vv;
  }
}

// Checksum of a set of fp32 tensors (the packed-weight cache key of the Python boundary): one block per tensor sums x and
// x * (1 + (i & 1023)) in double in a fixed order; a second launch folds the per-tensor pairs with tensor-dependent weights.
__global__ void __launch_bounds__(256) checksum_tensors_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ numels,
                                                               double* __restrict__ partial) {
  __shared__ double r0[8], r1[8];
  const float* x = reinterpret_cast<const float*>(ptrs[blockIdx.x]);
  const long long n = numels[blockIdx.x];
  double a = 0.0, b = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = (double)__ldg(x + i);
    a += v;
    b += v * (double)(1 + (int)(i & 1023));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    r0[threadIdx.x >> 5] = a;
    r1[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < 8; ++w) {
      sa += r0[w];
      sb += r1[w];
    }
    partial[2 * blockIdx.x] = sa;
    partial[2 * blockIdx.x + 1] = sb;
  }
}
__global__ void checksum_fold_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < n; ++i) {
      a += partial[2 * i] * (double)(1 + (i % 61));
      b += partial[2 * i + 1] * (double)(1 + (i % 67));
    }
    out[0] = a;
    out[1] = b;
  }
}

}  // namespace ssr

using namespace ssr;

extern "C" {

size_t ssr_l1_loss_workspace_bytes(void) { return (size_t)L1_MAX_BLOCKS * sizeof(double); }

int ssr_l1_loss(const float* out, const float* y, int64_t n, float* loss, float* dout, void* workspace, size_t workspace_bytes,
                void* stream) {
  SSR_CHECK(out && y && loss && n > 0 && workspace, SSR_E_INVALID, "ssr_l1_loss: bad argument");
  SSR_CHECK(workspace_bytes >= ssr_l1_loss_workspace_bytes(), SSR_E_WORKSPACE, "ssr_l1_loss: workspace %zu B too small", workspace_bytes);
  SSR_CHECK((((uintptr_t)out | (uintptr_t)y | (uintptr_t)dout) & 15u) == 0, SSR_E_INVALID, "ssr_l1_loss: buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  long long blocks = (n / 4 + L1_THREADS - 1) / L1_THREADS;
  if (blocks > L1_MAX_BLOCKS) blocks = L1_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  {
    ProfScope prof("l1_loss", 0.0, (double)n * (dout ? 12 : 8), s);
    l1_partial_kernel<<<(int)blocks, L1_THREADS, 0, s>>>(out, y, dout, reinterpret_cast<double*>(workspace), n, 1.0f / (float)n);
    l1_final_kernel<<<1, 32, 0, s>>>(reinterpret_cast<const double*>(workspace), (int)blocks, loss, 1.0 / (double)n);
    count_launch(2);
  }
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

int ssr_tensors_checksum(const int64_t* dev_ptrs, const int64_t* dev_numels, int n, double* dev_partial, double* dev_out2, void* stream) {
  SSR_CHECK(dev_ptrs && dev_numels && dev_partial && dev_out2 && n > 0, SSR_E_INVALID, "ssr_tensors_checksum: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  checksum_tensors_kernel<<<n, 256, 0, s>>>(reinterpret_cast<const long long*>(dev_ptrs), reinterpret_cast<const long long*>(dev_numels),
                                           dev_partial);
  checksum_fold_kernel<<<1, 32, 0, s>>>(dev_partial, n, dev_out2);
  count_launch(2);
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

int ssr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                  double weight_decay, int64_t step, float grad_scale, void* stream) {
  SSR_CHECK(p && g && m && v && n > 0 && step >= 1, SSR_E_INVALID, "ssr_adam_step: bad argument");
  SSR_CHECK((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15u) == 0, SSR_E_INVALID,
            "ssr_adam_step: buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  // bias corrections on the host in double, as torch.optim.Adam's capturable=False path computes them from the python float
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float neg_step_size = (float)(-(lr / bc1)), inv_bc2_sqrt = 1.0f / (float)sqrt(bc2);
  const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = 148LL * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  ProfScope prof("adam_flat", 0.0, (double)n * 28, s);
  adam_flat_kernel<<<(int)blocks, 256, 0, s>>>(p, g, m, v, n, w1, (float)beta2, w2, (float)eps, (float)weight_decay, neg_step_size, inv_bc2_sqrt, grad_scale);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // extern "C"
