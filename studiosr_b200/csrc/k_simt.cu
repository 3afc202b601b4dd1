// CUDA-core kernels: the exact-fp32 implicit GEMM (SSR_PREC_FP32), fp32 window attention,
// LayerNorm, and the memory-bound ends of the network (first conv with fused input handling,
// last conv with fused output handling, tile blend) plus layout utilities.
#include <math.h>

#include "ssr_device.cuh"

namespace ssr {

// =============================================================================================
// fp32 implicit GEMM, fused epilogue.  Tile: 32 rows x BN columns, K step 16, 256 threads; thread
// (ty = warp, tx = lane) owns rows ty*4..+3 and columns tx + 32*j.
// =============================================================================================
constexpr int SG_BM = 32;
constexpr int SG_BK = 16;

template <int BN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs g) {
  extern __shared__ float smem[];
  float* As = smem;                          // [BM][BK+1]
  float* Ws = smem + SG_BM * (SG_BK + 1);    // [BK][BN]
  float* Vs = smem;                          // [BM][BN+1], aliases As/Ws after the main loop
  constexpr int NJ = BN / 32;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * BN;
  const float* __restrict__ A = reinterpret_cast<const float*>(g.A);
  const float* __restrict__ Wt = reinterpret_cast<const float*>(g.Wt);
  const int Ktot = g.taps * g.KP;

  float acc[4][NJ];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.0f;

  const int lr = tid >> 3, lk = (tid & 7) * 2;
  const int lm = m0 + lr;
  const bool lvalid = lm < g.M;
  int px = 0, py = 0, pb = 0;
  if (g.taps == 9 && lvalid) {
    px = lm % g.W;
    py = (lm / g.W) % g.H;
    pb = lm / (g.W * g.H);
  }
  for (int tap = 0; tap < g.taps; ++tap) {
    const float* arow = nullptr;
    if (lvalid) {
      if (g.taps == 9) {
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) arow = A + ((size_t)(pb * g.H + yy) * g.W + xx) * g.lda;
      } else {
        arow = A + (size_t)lm * g.lda;
      }
    }
    for (int k0 = 0; k0 < g.KP; k0 += SG_BK) {
      float2 av = arow ? *reinterpret_cast<const float2*>(arow + k0 + lk) : make_float2(0.f, 0.f);
      As[lr * (SG_BK + 1) + lk] = av.x;
      As[lr * (SG_BK + 1) + lk + 1] = av.y;
      for (int n = tid; n < BN; n += 256) {
        const float4* wp = reinterpret_cast<const float4*>(Wt + (size_t)(n0 + n) * Ktot + (size_t)tap * g.KP + k0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = __ldg(wp + q);
          Ws[(q * 4 + 0) * BN + n] = w.x;
          Ws[(q * 4 + 1) * BN + n] = w.y;
          Ws[(q * 4 + 2) * BN + n] = w.z;
          Ws[(q * 4 + 3) * BN + n] = w.w;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SG_BK; ++kk) {
        float a[4], w[NJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[(ty * 4 + i) * (SG_BK + 1) + kk];
#pragma unroll
        for (int j = 0; j < NJ; ++j) w[j] = Ws[kk * BN + tx + 32 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue ----
  const int Cps = g.ps_r > 1 ? g.N / (g.ps_r * g.ps_r) : 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    const bool valid = m < g.M;
    int b = 0, y = 0, x = 0;
    if (valid && g.ps_r > 1) {
      x = m % g.W;
      y = (m / g.W) % g.H;
      b = m / (g.W * g.H);
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int n = n0 + tx + 32 * j;
      float v = 0.0f;
      if (valid) {
        v = apply_act(acc[i][j] + __ldg(g.bias + n), g.act, g.slope) * g.alpha;
        if (g.res) v += g.res[(size_t)m * g.ldres + n];
        if (n >= g.N) v = 0.0f;
        if (g.out_f32) g.out_f32[(size_t)m * g.ld_f32 + n] = v;
        if (g.out_T) {
          float* o = reinterpret_cast<float*>(g.out_T);
          const float vs = g.round_tf32 ? round_tf32(v) : v;
          if (g.ps_r > 1) {
            if (n < g.N) o[ps_offset(b, y, x, n, g.H, g.W, g.ps_r, Cps, g.ld_T)] = vs;
          } else {
            o[(size_t)m * g.ld_T + n] = vs;
          }
        }
      }
      if (g.out_ln) acc[i][j] = v;
    }
  }
  if (g.out_ln) {  // BN == NP here (host guarantees gridDim.y == 1)
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) Vs[(ty * 4 + i) * (BN + 1) + tx + 32 * j] = acc[i][j];
    __syncthreads();
    float* o = reinterpret_cast<float*>(g.out_ln);
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i, m = m0 + r;
      if (m >= g.M) continue;
      float s = 0.0f;
      for (int n = tx; n < g.N; n += 32) s += Vs[r * (BN + 1) + n];
      const float mean = warp_sum(s) / (float)g.N;
      float q = 0.0f;
      for (int n = tx; n < g.N; n += 32) {
        const float d = Vs[r * (BN + 1) + n] - mean;
        q += d * d;
      }
      const float rstd = rsqrtf(warp_sum(q) / (float)g.N + g.eps);
      for (int n = tx; n < BN; n += 32) {
        float yv = 0.0f;
        if (n < g.N) yv = (Vs[r * (BN + 1) + n] - mean) * rstd * __ldg(g.gamma + n) + __ldg(g.beta + n);
        o[(size_t)m * g.ld_ln + n] = g.round_tf32 ? round_tf32(yv) : yv;
      }
    }
  }
}

template <int BN>
static int launch_gemm_simt_bn(const GemmArgs& g, cudaStream_t s) {
  const size_t main_bytes = (size_t)(SG_BM * (SG_BK + 1) + SG_BK * BN) * sizeof(float);
  const size_t ln_bytes = (size_t)SG_BM * (BN + 1) * sizeof(float);
  const size_t smem = main_bytes > ln_bytes ? main_bytes : ln_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CUDA(cudaFuncSetAttribute(gemm_simt_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  dim3 grid((g.M + SG_BM - 1) / SG_BM, g.NP / BN);
  ProfScope prof(g.taps == 9 ? "gemm_fp32_conv3x3" : "gemm_fp32_linear", gemm_alg_flops(g), gemm_alg_bytes(g, 4), s);
  gemm_simt_kernel<BN><<<grid, 256, smem, s>>>(g);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

int launch_gemm_simt(const GemmArgs& g, cudaStream_t s) {
  SSR_CHECK(g.NP % 64 == 0 && g.KP % 16 == 0, SSR_E_INVALID, "gemm_simt: NP=%d KP=%d not padded", g.NP, g.KP);
  if (g.out_ln) {
    SSR_CHECK(g.NP <= 256, SSR_E_INVALID, "gemm_simt: fused LayerNorm needs NP<=256 (NP=%d)", g.NP);
    switch (g.NP) {
      case 64: return launch_gemm_simt_bn<64>(g, s);
      case 128: return launch_gemm_simt_bn<128>(g, s);
      case 192: return launch_gemm_simt_bn<192>(g, s);
      case 256: return launch_gemm_simt_bn<256>(g, s);
    }
  }
  if (g.NP % 256 == 0) return launch_gemm_simt_bn<256>(g, s);
  if (g.NP % 192 == 0) return launch_gemm_simt_bn<192>(g, s);
  if (g.NP % 128 == 0) return launch_gemm_simt_bn<128>(g, s);
  return launch_gemm_simt_bn<64>(g, s);
}

// =============================================================================================
// fp32 (shifted-)window attention: one CTA per (window, head) -- at batch 1 a grid of windows alone leaves half the SMs idle
// and serialises the heads -- looping over 64-query row blocks.
// =============================================================================================
__global__ void __launch_bounds__(256) attn_simt_kernel(const AttnArgs a) {
  extern __shared__ float smem[];
  const int N = a.ws * a.ws;
  const int DPp = a.DP + 1;
  float* ks = smem;                // [N][DP+1]
  float* vs = ks + N * DPp;        // [N][DP+1]
  float* qs = vs + N * DPp;        // [64][DP+1]
  float* S = qs + 64 * DPp;        // [64][N+1]
  int* pix = reinterpret_cast<int*>(S + 64 * (N + 1));  // [N] source row index of each window token
  int* rid = pix + N;                                   // [N] shift-mask region id

  const int nwx = a.W / a.ws, nwy = a.H / a.ws;
  int w = blockIdx.x;
  const int wx = w % nwx;
  w /= nwx;
  const int wy = w % nwy;
  const int b = w / nwy;
  const int tid = threadIdx.x;
  const int elem = a.elem ? a.elem : 4;  // 4: fp32 activations (fp32 / tf32 models), 2: bf16
  const int nb = 2 * a.ws - 1;

  for (int t = tid; t < N; t += blockDim.x) {
    const int sy = wy * a.ws + t / a.ws, sx = wx * a.ws + t % a.ws;
    const int yy = (sy + a.shift) % a.H, xx = (sx + a.shift) % a.W;
    pix[t] = (b * a.H + yy) * a.W + xx;
    rid[t] = a.shift > 0 ? 3 * shift_region(sy, a.H, a.ws, a.shift) + shift_region(sx, a.W, a.ws, a.shift) : 0;
  }
  __syncthreads();

  {
    const int h = blockIdx.y;
    for (int e = tid; e < N * a.DP; e += blockDim.x) {
      const int t = e / a.DP, j = e % a.DP;
      const size_t row = (size_t)pix[t] * a.ld_qkv + h * a.DP + j;
      ks[t * DPp + j] = load_elem(a.qkv, row + a.QP, elem);
      vs[t * DPp + j] = load_elem(a.qkv, row + 2 * a.QP, elem);
    }
    const float* btab = a.bias + (size_t)h * nb * nb;
    for (int q0 = 0; q0 < N; q0 += 64) {
      __syncthreads();
      for (int e = tid; e < 64 * a.DP; e += blockDim.x) {
        const int t = e / a.DP, j = e % a.DP;
        qs[t * DPp + j] = load_elem(a.qkv, (size_t)pix[q0 + t] * a.ld_qkv + h * a.DP + j, elem);
      }
      __syncthreads();
      for (int e = tid; e < 64 * N; e += blockDim.x) {
        const int i = e / N, j = e % N;
        const int ti = q0 + i;
        float s = 0.0f;
        for (int c = 0; c < a.d; ++c) s = fmaf(qs[i * DPp + c], ks[j * DPp + c], s);
        const int dy = ti / a.ws - j / a.ws + a.ws - 1, dx = ti % a.ws - j % a.ws + a.ws - 1;
        s += __ldg(btab + dy * nb + dx);
        if (rid[ti] != rid[j]) s += -100.0f;
        S[i * (N + 1) + j] = s;
      }
      __syncthreads();
      const int warp = tid >> 5, lane = tid & 31;
      for (int i = warp; i < 64; i += 8) {
        float mx = -INFINITY;
        for (int j = lane; j < N; j += 32) mx = fmaxf(mx, S[i * (N + 1) + j]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < N; j += 32) {
          const float p = expf(S[i * (N + 1) + j] - mx);
          S[i * (N + 1) + j] = p;
          sum += p;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < N; j += 32) S[i * (N + 1) + j] *= inv;
      }
      __syncthreads();
      for (int e = tid; e < 64 * a.DP; e += blockDim.x) {
        const int i = e / a.DP, c = e % a.DP;
        float acc = 0.0f;
        if (c < a.d)
          for (int j = 0; j < N; ++j) acc = fmaf(S[i * (N + 1) + j], vs[j * DPp + c], acc);
        store_elem(a.o, (size_t)pix[q0 + i] * a.ld_o + h * a.DP + c, elem, acc, 0);
      }
    }
    __syncthreads();
  }
}

// =============================================================================================
// HAT overlapping cross-attention core (hat.py:257-284): queries = one ws x ws window, keys / values =
// the kws x kws window around it (nn.Unfold: stride ws, zero padding (kws - ws) / 2 -- out-of-image keys
// are ZERO vectors that still enter the softmax with their bias).  One CTA per window; per head the key
// tile is staged once per 32-query block for the scores and replaced by the value tile for P.V.
// Bias index = (ky - qy + ws - kws + 1) * nb + (kx - qx + ws - kws + 1), nb = ws + kws - 1, with the
// reference's negative-index wrap-around (hat.py:508-512 + Python indexing).
// =============================================================================================
__global__ void __launch_bounds__(256) attn_oca_kernel(const AttnArgs a) {
  extern __shared__ float smem[];
  const int Nq = a.ws * a.ws, Nk = a.kws * a.kws;
  const int DPp = a.DP + 1;
  float* kv = smem;              // [Nk][DP+1]
  float* qs = kv + Nk * DPp;     // [32][DP+1]
  float* S = qs + 32 * DPp;      // [32][Nk+1]
  int* kpix = reinterpret_cast<int*>(S + 32 * (Nk + 1));  // [Nk] source row or -1 (zero padding)
  int* qpix = kpix + Nk;                                  // [Nq]
  const int elem = a.elem ? a.elem : 4;
  const int nwx = a.W / a.ws, nwy = a.H / a.ws;
  int w = blockIdx.x;
  const int wx = w % nwx;
  w /= nwx;
  const int wy = w % nwy;
  const int b = w / nwy;
  const int tid = threadIdx.x;
  const int pad = (a.kws - a.ws) / 2;
  const int nb = a.ws + a.kws - 1, off = a.ws - a.kws + 1;
  for (int t = tid; t < Nk; t += blockDim.x) {
    const int y = wy * a.ws - pad + t / a.kws, x = wx * a.ws - pad + t % a.kws;
    kpix[t] = (y >= 0 && y < a.H && x >= 0 && x < a.W) ? (b * a.H + y) * a.W + x : -1;
  }
  for (int t = tid; t < Nq; t += blockDim.x) qpix[t] = (b * a.H + wy * a.ws + t / a.ws) * a.W + wx * a.ws + t % a.ws;
  __syncthreads();
  for (int h = 0; h < a.heads; ++h) {
    const float* btab = a.bias + (size_t)h * nb * nb;
    for (int q0 = 0; q0 < Nq; q0 += 32) {
      __syncthreads();
      for (int e = tid; e < Nk * a.DP; e += blockDim.x) {  // keys
        const int t = e / a.DP, j = e % a.DP;
        kv[t * DPp + j] = kpix[t] >= 0 ? load_elem(a.qkv, (size_t)kpix[t] * a.ld_qkv + a.QP + h * a.DP + j, elem) : 0.0f;
      }
      for (int e = tid; e < 32 * a.DP; e += blockDim.x) {
        const int t = e / a.DP, j = e % a.DP;
        qs[t * DPp + j] = load_elem(a.qkv, (size_t)qpix[q0 + t] * a.ld_qkv + h * a.DP + j, elem);
      }
      __syncthreads();
      for (int e = tid; e < 32 * Nk; e += blockDim.x) {
        const int i = e / Nk, j = e % Nk;
        const int ti = q0 + i;
        float sc = 0.0f;
        for (int c = 0; c < a.d; ++c) sc = fmaf(qs[i * DPp + c], kv[j * DPp + c], sc);
        int idx = (j / a.kws - ti / a.ws + off) * nb + (j % a.kws - ti % a.ws + off);
        if (idx < 0) idx += nb * nb;
        S[i * (Nk + 1) + j] = sc + __ldg(btab + idx);
      }
      __syncthreads();
      const int warp = tid >> 5, lane = tid & 31;
      for (int i = warp; i < 32; i += 8) {
        float mx = -INFINITY;
        for (int j = lane; j < Nk; j += 32) mx = fmaxf(mx, S[i * (Nk + 1) + j]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < Nk; j += 32) {
          const float p = expf(S[i * (Nk + 1) + j] - mx);
          S[i * (Nk + 1) + j] = p;
          sum += p;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < Nk; j += 32) S[i * (Nk + 1) + j] *= inv;
      }
      for (int e = tid; e < Nk * a.DP; e += blockDim.x) {  // values replace the keys (all score reads are done: the
        const int t = e / a.DP, j = e % a.DP;              // softmax above touches S only, the barrier below orders the rest)
        kv[t * DPp + j] = kpix[t] >= 0 ? load_elem(a.qkv, (size_t)kpix[t] * a.ld_qkv + 2 * a.QP + h * a.DP + j, elem) : 0.0f;
      }
      __syncthreads();
      for (int e = tid; e < 32 * a.DP; e += blockDim.x) {
        const int i = e / a.DP, c = e % a.DP;
        float acc = 0.0f;
        if (c < a.d)
          for (int j = 0; j < Nk; ++j) acc = fmaf(S[i * (Nk + 1) + j], kv[j * DPp + c], acc);
        store_elem(a.o, (size_t)qpix[q0 + i] * a.ld_o + h * a.DP + c, elem, acc, 0);
      }
    }
    __syncthreads();
  }
}

int launch_attn_oca(const AttnArgs& a, cudaStream_t s) {
  SSR_CHECK(a.H % a.ws == 0 && a.W % a.ws == 0, SSR_E_INVALID, "attn_oca: %dx%d not a multiple of ws=%d", a.H, a.W, a.ws);
  SSR_CHECK(a.kws > a.ws && (a.kws - a.ws) % 2 == 0 && (a.ws * a.ws) % 32 == 0, SSR_E_INVALID, "attn_oca: windows %d / %d", a.ws, a.kws);
  const int Nq = a.ws * a.ws, Nk = a.kws * a.kws;
  const size_t smem = (size_t)(Nk * (a.DP + 1) + 32 * (a.DP + 1) + 32 * (Nk + 1)) * sizeof(float) + (size_t)(Nk + Nq) * sizeof(int);
  SSR_CHECK(smem <= 227 * 1024, SSR_E_INVALID, "attn_oca: %zu B of shared memory", smem);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    SSR_CUDA(cudaFuncSetAttribute(attn_oca_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int nwin = a.B * (a.H / a.ws) * (a.W / a.ws);
  ProfScope prof("attn_oca", 4.0 * nwin * Nq * Nk * a.d * a.heads, (double)nwin * (Nq + 2.25 * Nq) * a.heads * a.d * (a.elem ? a.elem : 4), s);
  attn_oca_kernel<<<nwin, 256, smem, s>>>(a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

int launch_attn_simt(const AttnArgs& a, cudaStream_t s) {
  const int N = a.ws * a.ws;
  SSR_CHECK(N % 64 == 0, SSR_E_INVALID, "attn_simt: window_size^2 must be a multiple of 64 (ws=%d)", a.ws);
  SSR_CHECK(a.H % a.ws == 0 && a.W % a.ws == 0, SSR_E_INVALID, "attn: %dx%d not a multiple of ws=%d", a.H, a.W, a.ws);
  const size_t smem = (size_t)(2 * N * (a.DP + 1) + 64 * (a.DP + 1) + 64 * (N + 1)) * sizeof(float) + 2 * N * sizeof(int);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    SSR_CUDA(cudaFuncSetAttribute(attn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int nwin = a.B * (a.H / a.ws) * (a.W / a.ws);
  ProfScope prof("attn_fp32", 4.0 * nwin * N * N * a.d * a.heads, 4.0 * nwin * N * a.heads * a.d * 4, s);
  attn_simt_kernel<<<dim3(nwin, a.heads), 256, smem, s>>>(a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// =============================================================================================
// LayerNorm (one warp per row), optionally two chained LNs (patch_embed.norm -> blocks.0.norm1).
// =============================================================================================
__global__ void __launch_bounds__(256) layernorm_kernel(const LnArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.M) return;
  const float* x = a.in + (size_t)warp * a.ld_in;
  float v[8];
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane + 32 * j;
    v[j] = n < a.C ? x[n] : 0.0f;
    s += v[j];
  }
  float mean = warp_sum(s) / (float)a.C;
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane + 32 * j;
    if (n < a.C) q += (v[j] - mean) * (v[j] - mean);
  }
  float rstd = rsqrtf(warp_sum(q) / (float)a.C + a.eps);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane + 32 * j;
    v[j] = n < a.C ? (v[j] - mean) * rstd * __ldg(a.g1 + n) + __ldg(a.b1 + n) : 0.0f;
    if (a.out_f32 && n < a.CP) a.out_f32[(size_t)warp * a.ld_f32 + n] = v[j];
  }
  if (!a.out_T) return;
  if (a.g2) {
    s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    mean = warp_sum(s) / (float)a.C;
    q = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (lane + 32 * j < a.C) q += (v[j] - mean) * (v[j] - mean);
    rstd = rsqrtf(warp_sum(q) / (float)a.C + a.eps);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = lane + 32 * j;
      v[j] = n < a.C ? (v[j] - mean) * rstd * __ldg(a.g2 + n) + __ldg(a.b2 + n) : 0.0f;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane + 32 * j;
    if (n < a.CP) store_elem(a.out_T, (size_t)warp * a.ld_T + n, a.elem, v[j], a.round_tf32);
  }
}

int launch_layernorm(const LnArgs& a, cudaStream_t s) {
  SSR_CHECK(a.CP <= 256, SSR_E_INVALID, "layernorm: C padded %d > 256", a.CP);
  const int blocks = (a.M + 7) / 8;
  ProfScope prof("layernorm", 0.0, (double)a.M * a.C * (4 + (a.out_f32 ? 4 : 0) + (a.out_T ? a.elem : 0)), s);
  layernorm_kernel<<<blocks, 256, 0, s>>>(a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// =============================================================================================
// First conv: n_colors(3) -> Cout, fused uint8/float input, tile addressing, mirror pad, affine.
// CTA = 16 consecutive padded pixels x Cout channels (thread = channel).
// =============================================================================================
constexpr int CF_PIX = 16;

__device__ __forceinline__ int mirror_index(int i, int n, int pad_mode) {
  if (i < n) return i;
  return pad_mode == 0 ? 2 * n - 1 - i : 2 * (n - 1) - i;
}

__global__ void __launch_bounds__(256) conv_first_kernel(const ConvFirstArgs a) {
  __shared__ __align__(16) float patch[CF_PIX][28];  // 27 taps + 1 pad: a pixel's patch is seven 16-byte broadcast loads
  const int tid = threadIdx.x;
  const int M = a.B * a.Hp * a.Wp;
  const int m0 = blockIdx.x * CF_PIX;
  for (int e = tid; e < CF_PIX * 27; e += blockDim.x) {
    const int p = e / 27, k = e % 27;
    const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
    const int m = m0 + p;
    float v = 0.0f;
    if (m < M) {
      const int x = m % a.Wp, y = (m / a.Wp) % a.Hp, b = m / (a.Wp * a.Hp);
      const int yy = y + ky - 1, xx = x + kx - 1;
      if (yy >= 0 && yy < a.Hp && xx >= 0 && xx < a.Wp) {
        int sy = mirror_index(yy, a.h, a.pad_mode), sx = mirror_index(xx, a.w, a.pad_mode);
        int img = b;
        if (a.tile_mode) {
          const int t = a.tile_begin + b;
          const int ty = t / a.tiles_x, tx = t % a.tiles_x;
          int y0 = ty * a.stride, x0 = tx * a.stride;
          if (y0 > a.fh - a.h) y0 = a.fh - a.h;
          if (x0 > a.fw - a.w) x0 = a.fw - a.w;
          sy += y0;
          sx += x0;
          img = 0;
        }
        float raw;
        if (a.in_u8)
          raw = (float)reinterpret_cast<const uint8_t*>(a.in)[((size_t)(img * a.fh + sy) * a.fw + sx) * 3 + ci];
        else
          raw = reinterpret_cast<const float*>(a.in)[((size_t)(img * 3 + ci) * a.fh + sy) * a.fw + sx];
        v = raw * a.in_scale + a.in_shift[ci];
      }
    }
    patch[p][k] = v;
  }
  __syncthreads();
  const int c = tid;
  const int ldmax = a.out_f32 ? a.ld_f32 : a.ld_T;
  if (c >= ldmax && !(a.out_T && c < a.ld_T)) return;
  float w[27];
  float bias = 0.0f;
  if (c < a.Cout) {
#pragma unroll
    for (int k = 0; k < 27; ++k) w[k] = __ldg(a.Wc + c * 27 + k);
    bias = __ldg(a.bias + c);
  } else {
#pragma unroll
    for (int k = 0; k < 27; ++k) w[k] = 0.0f;
  }
  for (int p = 0; p < CF_PIX; ++p) {
    const int m = m0 + p;
    if (m >= M) break;
    float acc = bias;
    const float4* pp = reinterpret_cast<const float4*>(patch[p]);  // LDS.128 x 7 instead of LDS.32 x 27: the kernel was LDS-bound
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const float4 v = pp[q];
      acc = fmaf(v.x, w[4 * q], acc);
      if (4 * q + 1 < 27) acc = fmaf(v.y, w[4 * q + 1], acc);
      if (4 * q + 2 < 27) acc = fmaf(v.z, w[4 * q + 2], acc);
      if (4 * q + 3 < 27) acc = fmaf(v.w, w[4 * q + 3], acc);
    }
    if (a.out_f32 && c < a.ld_f32) a.out_f32[(size_t)m * a.ld_f32 + c] = acc;
    if (a.out_T && c < a.ld_T) store_elem(a.out_T, (size_t)m * a.ld_T + c, a.elem, acc, a.round_tf32);
  }
}

int launch_conv_first(const ConvFirstArgs& a, cudaStream_t s) {
  SSR_CHECK(a.Cout <= 256 && a.ld_f32 <= 256 && a.ld_T <= 256, SSR_E_INVALID, "conv_first: Cout=%d > 256", a.Cout);
  const int M = a.B * a.Hp * a.Wp;
  ProfScope prof("conv_first", 2.0 * M * 27 * a.Cout,
                 (double)M * (3 * (a.in_u8 ? 1 : 4) + a.Cout * ((a.out_f32 ? 4 : 0) + (a.out_T ? a.elem : 0))), s);
  conv_first_kernel<<<(M + CF_PIX - 1) / CF_PIX, 256, 0, s>>>(a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// =============================================================================================
// First conv + patch-embed LayerNorm + first block's norm1 in ONE pass (SwinIR, swinir.py:358-361,344-346,151):
//   x0 = conv_first(pad(normalise(x)))  ->  g = LN(x0; patch_embed.norm)  ->  xn = LN(g; layers.0.blocks.0.norm1)
// The three outputs are the whole cost (12 B + sizeof(T) per channel and pixel written, 3 bytes read): an HBM stream.  The
// two-kernel version wrote x0, read it back and ran its conv with one thread per channel (15 % / 57 % of the HBM roofline).
// A warp owns 8 pixels at a time; lane l holds channels (2l, 2l+1) + 64 i of every pixel, so weights are LDS.64 from a
// [tap][channel] table, the input patch is a broadcast LDS and every store is a full 128- / 256-byte line per warp.
// =============================================================================================
constexpr int CFL_WARPS = 8, CFL_PX = 8;  // 64 pixels per CTA pass

template <int NG>  // channel groups of 64: CP = 64 * NG
__global__ void __launch_bounds__(32 * CFL_WARPS) conv_first_ln_kernel(const ConvFirstArgs a, int n_groups_px) {
  extern __shared__ __align__(16) float cfl_smem[];
  float* s_w = cfl_smem;                       // [27][64 NG]
  float* s_bias = s_w + 27 * 64 * NG;          // [64 NG]
  float* s_g1 = s_bias + 64 * NG;
  float* s_b1 = s_g1 + 64 * NG;
  float* s_g2 = s_b1 + 64 * NG;
  float* s_b2 = s_g2 + 64 * NG;
  float* s_patch = s_b2 + 64 * NG;             // [CFL_WARPS][CFL_PX][28]
  constexpr int CP = 64 * NG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 27 * CP; e += blockDim.x) {
    const int k = e / CP, c = e - k * CP;
    s_w[e] = c < a.Cout ? __ldg(a.Wc + c * 27 + k) : 0.0f;
  }
  for (int c = tid; c < CP; c += blockDim.x) {
    s_bias[c] = c < a.Cout ? __ldg(a.bias + c) : 0.0f;
    s_g1[c] = __ldg(a.g1 + c);
    s_b1[c] = __ldg(a.b1 + c);
    s_g2[c] = a.g2 ? __ldg(a.g2 + c) : 0.0f;
    s_b2[c] = a.g2 ? __ldg(a.b2 + c) : 0.0f;
  }
  __syncthreads();
  const int M = a.B * a.Hp * a.Wp;
  const float invC = 1.0f / (float)a.Cout;
  float* patch = s_patch + warp * (CFL_PX * 28);
  for (int grp = blockIdx.x * CFL_WARPS + warp; grp < n_groups_px; grp += gridDim.x * CFL_WARPS) {
    const int m0 = grp * CFL_PX;
    // ---- the 8 pixels' 3x3x3 input patches (pad + normalise + tile addressing as conv_first_kernel) ----
    __syncwarp();
    for (int e = lane; e < CFL_PX * 27; e += 32) {
      const int p = e / 27, k = e - p * 27;
      const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
      const int m = m0 + p;
      float v = 0.0f;
      if (m < M) {
        const int x = m % a.Wp, y = (m / a.Wp) % a.Hp, b = m / (a.Wp * a.Hp);
        const int yy = y + ky - 1, xx = x + kx - 1;
        if (yy >= 0 && yy < a.Hp && xx >= 0 && xx < a.Wp) {
          int sy = mirror_index(yy, a.h, a.pad_mode), sx = mirror_index(xx, a.w, a.pad_mode);
          int img = b;
          if (a.tile_mode) {
            const int t = a.tile_begin + b;
            const int ty = t / a.tiles_x, tx = t % a.tiles_x;
            int y0 = ty * a.stride, x0 = tx * a.stride;
            if (y0 > a.fh - a.h) y0 = a.fh - a.h;
            if (x0 > a.fw - a.w) x0 = a.fw - a.w;
            sy += y0;
            sx += x0;
            img = 0;
          }
          float raw;
          if (a.in_u8)
            raw = (float)reinterpret_cast<const uint8_t*>(a.in)[((size_t)(img * a.fh + sy) * a.fw + sx) * 3 + ci];
          else
            raw = reinterpret_cast<const float*>(a.in)[((size_t)(img * 3 + ci) * a.fh + sy) * a.fw + sx];
          v = raw * a.in_scale + a.in_shift[ci];
        }
      }
      patch[p * 28 + k] = v;
    }
    __syncwarp();
    // ---- conv: acc[p][g] = (channels 64 g + 2 lane, + 1) of pixel p ----
    float2 acc[CFL_PX][NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const float2 bv = *reinterpret_cast<const float2*>(s_bias + 64 * g + 2 * lane);
#pragma unroll
      for (int p = 0; p < CFL_PX; ++p) acc[p][g] = bv;
    }
#pragma unroll 3
    for (int k = 0; k < 27; ++k) {
      float2 w[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) w[g] = *reinterpret_cast<const float2*>(s_w + k * CP + 64 * g + 2 * lane);
#pragma unroll
      for (int p = 0; p < CFL_PX; ++p) {
        const float xv = patch[p * 28 + k];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          acc[p][g].x = fmaf(xv, w[g].x, acc[p][g].x);
          acc[p][g].y = fmaf(xv, w[g].y, acc[p][g].y);
        }
      }
    }
    // ---- per pixel: store x0, LN -> g, LN -> xn ----
#pragma unroll
    for (int p = 0; p < CFL_PX; ++p) {
      const int m = m0 + p;
      if (m >= M) break;  // warp-uniform
      if (a.out_f32) {
#pragma unroll
        for (int g = 0; g < NG; ++g) *reinterpret_cast<float2*>(a.out_f32 + (size_t)m * a.ld_f32 + 64 * g + 2 * lane) = acc[p][g];
      }
      float2 v[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) v[g] = acc[p][g];
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {  // pass 0: patch_embed.norm (g1, b1); pass 1: norm1 of the first block (g2, b2)
        if (pass == 1 && !a.g2) break;
        float sum = 0.0f;
#pragma unroll
        for (int g = 0; g < NG; ++g) sum += v[g].x + v[g].y;  // padded channels are exact zeros
        const float mean = warp_sum(sum) * invC;
        float sq = 0.0f;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int c = 64 * g + 2 * lane;
          const float dx = c < a.Cout ? v[g].x - mean : 0.0f, dy = c + 1 < a.Cout ? v[g].y - mean : 0.0f;
          sq += dx * dx + dy * dy;
        }
        const float rstd = rsqrtf(warp_sum(sq) * invC + a.eps);
        const float* gam = pass == 0 ? s_g1 : s_g2;
        const float* bet = pass == 0 ? s_b1 : s_b2;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int c = 64 * g + 2 * lane;
          const float2 gv = *reinterpret_cast<const float2*>(gam + c), bv = *reinterpret_cast<const float2*>(bet + c);
          // pad channels take the (padded) beta: 0, or the 1.0 that carries the qkv bias of the fused attention kernel
          v[g].x = c < a.Cout ? (v[g].x - mean) * rstd * gv.x + bv.x : bv.x;
          v[g].y = c + 1 < a.Cout ? (v[g].y - mean) * rstd * gv.y + bv.y : bv.y;
        }
        if (pass == 0 && a.out_g) {
#pragma unroll
          for (int g = 0; g < NG; ++g) *reinterpret_cast<float2*>(a.out_g + (size_t)m * a.ld_g + 64 * g + 2 * lane) = v[g];
        }
      }
      if (a.out_T) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const size_t idx = (size_t)m * a.ld_T + 64 * g + 2 * lane;
          if (a.elem == 2) {
            reinterpret_cast<uint32_t*>(a.out_T)[idx >> 1] = pack_bf16x2(v[g].x, v[g].y);
          } else {
            float2 o = v[g];
            if (a.round_tf32) {
              o.x = round_tf32(o.x);
              o.y = round_tf32(o.y);
            }
            *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.out_T) + idx) = o;
          }
        }
      }
    }
  }
}

int launch_conv_first_ln(const ConvFirstArgs& a, int CP, cudaStream_t s) {
  SSR_CHECK(CP % 64 == 0 && CP >= 64 && CP <= 192 && a.Cout <= CP, SSR_E_INVALID, "conv_first_ln: padded channels %d not in {64, 128, 192}", CP);
  SSR_CHECK(a.g1 && a.b1, SSR_E_INVALID, "conv_first_ln: missing LayerNorm parameters");
  const int M = a.B * a.Hp * a.Wp;
  const int n_groups = (M + CFL_PX - 1) / CFL_PX;
  const size_t smem = (size_t)(27 * CP + 5 * CP + CFL_WARPS * CFL_PX * 28) * sizeof(float);
  int blocks = (n_groups + CFL_WARPS - 1) / CFL_WARPS;
  const int cap = 148 * 4;  // B200: 148 SMs (sm_100a only)
  if (blocks > cap) blocks = cap;
  ProfScope prof("conv_first_ln", 2.0 * M * 27 * a.Cout,
                 (double)M * (3 * (a.in_u8 ? 1 : 4) + a.Cout * ((a.out_f32 ? 4 : 0) + (a.out_g ? 4 : 0) + (a.out_T ? a.elem : 0))), s);
  const int ng = CP / 64;
  if (ng == 1)
    conv_first_ln_kernel<1><<<blocks, 32 * CFL_WARPS, smem, s>>>(a, n_groups);
  else if (ng == 2)
    conv_first_ln_kernel<2><<<blocks, 32 * CFL_WARPS, smem, s>>>(a, n_groups);
  else
    conv_first_ln_kernel<3><<<blocks, 32 * CFL_WARPS, smem, s>>>(a, n_groups);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// =============================================================================================
// Last conv: Cin -> 3 with fused bias / affine / crop / (fp32 NCHW | uint8 HWC) store.
// One thread per output pixel; weights staged in shared memory as [tap][ci][4].
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(128) conv_last_kernel(const ConvLastArgs a) {
  extern __shared__ float wsm[];  // [9*Cin][4]
  for (int e = threadIdx.x; e < 9 * a.Cin * 4; e += blockDim.x) wsm[e] = __ldg(a.Wc + e);
  __syncthreads();
  const long long total = (long long)a.B * a.ch * a.cw;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % a.cw), y = (int)((idx / a.cw) % a.ch), b = (int)(idx / ((long long)a.cw * a.ch));
  const T* in = reinterpret_cast<const T*>(a.in);
  float acc0 = a.bias[0], acc1 = a.bias[1], acc2 = a.bias[2];
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= a.Hs || xx < 0 || xx >= a.Ws) continue;
    const T* row = in + ((size_t)(b * a.Hs + yy) * a.Ws + xx) * a.ldi;
    const float4* w4 = reinterpret_cast<const float4*>(wsm) + tap * a.Cin;
    if constexpr (sizeof(T) == 2) {
      for (int c = 0; c < a.Cin; c += 8) {
        const uint4 raw = *reinterpret_cast<const uint4*>(row + c);
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(hp[q]);
          const float4 wa = w4[c + 2 * q], wb = w4[c + 2 * q + 1];
          acc0 = fmaf(f.x, wa.x, acc0); acc1 = fmaf(f.x, wa.y, acc1); acc2 = fmaf(f.x, wa.z, acc2);
          acc0 = fmaf(f.y, wb.x, acc0); acc1 = fmaf(f.y, wb.y, acc1); acc2 = fmaf(f.y, wb.z, acc2);
        }
      }
    } else {
      for (int c = 0; c < a.Cin; c += 4) {
        const float4 f = *reinterpret_cast<const float4*>(row + c);
        const float4 w0 = w4[c], w1 = w4[c + 1], w2 = w4[c + 2], w3 = w4[c + 3];
        acc0 = fmaf(f.x, w0.x, acc0); acc1 = fmaf(f.x, w0.y, acc1); acc2 = fmaf(f.x, w0.z, acc2);
        acc0 = fmaf(f.y, w1.x, acc0); acc1 = fmaf(f.y, w1.y, acc1); acc2 = fmaf(f.y, w1.z, acc2);
        acc0 = fmaf(f.z, w2.x, acc0); acc1 = fmaf(f.z, w2.y, acc1); acc2 = fmaf(f.z, w2.z, acc2);
        acc0 = fmaf(f.w, w3.x, acc0); acc1 = fmaf(f.w, w3.y, acc1); acc2 = fmaf(f.w, w3.z, acc2);
      }
    }
  }
  const float r0 = (acc0 + a.out_shift[0]) * a.out_scale;
  const float r1 = (acc1 + a.out_shift[1]) * a.out_scale;
  const float r2 = (acc2 + a.out_shift[2]) * a.out_scale;
  if (a.out_f32) {
    const size_t plane = (size_t)a.ch * a.cw;
    float* o = a.out_f32 + (size_t)b * 3 * plane + (size_t)y * a.cw + x;
    o[0] = r0;
    o[plane] = r1;
    o[2 * plane] = r2;
  }
  if (a.out_u8) {
    uint8_t* o = a.out_u8 + ((size_t)(b * a.ch + y) * a.cw + x) * 3;
    o[0] = (uint8_t)fminf(fmaxf(rintf(r0 * a.u8_scale), 0.0f), 255.0f);
    o[1] = (uint8_t)fminf(fmaxf(rintf(r1 * a.u8_scale), 0.0f), 255.0f);
    o[2] = (uint8_t)fminf(fmaxf(rintf(r2 * a.u8_scale), 0.0f), 255.0f);
  }
}

int launch_conv_last(const ConvLastArgs& a, cudaStream_t s) {
  SSR_CHECK(a.Cin % 8 == 0 && a.Cin <= 256, SSR_E_INVALID, "conv_last: Cin=%d", a.Cin);
  const long long total = (long long)a.B * a.ch * a.cw;
  const int blocks = (int)((total + 127) / 128);
  const size_t smem = (size_t)9 * a.Cin * 4 * sizeof(float);
  ProfScope prof("conv_last", 2.0 * total * 9 * a.Cin * 3,
                 (double)a.B * a.Hs * a.Ws * a.Cin * a.elem + (double)total * 3 * (a.out_f32 ? 4 : 1), s);
  if (a.elem == 2)
    conv_last_kernel<__nv_bfloat16><<<blocks, 128, smem, s>>>(a);
  else
    conv_last_kernel<float><<<blocks, 128, smem, s>>>(a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// =============================================================================================
// Tile blend (gather form): out(Y,X) = sum_t w_t * tile_t / sum_t w_t over the tiles covering it.
// =============================================================================================
__device__ __forceinline__ int tile_origin(int i, int stride, int L, int tile) {
  const int o = i * stride;
  return o > L - tile ? L - tile : o;
}
__device__ __forceinline__ float ramp_w(int p, int n, int o, bool first, bool last) {
  // p in [0,n): position inside the tile (output samples); o: ramp length (output samples)
  float w = 1.0f;
  if (!first && p < o) w = ((float)p + 0.5f) / (float)o;
  if (!last && p >= n - o) w *= ((float)(n - 1 - p) + 0.5f) / (float)o;
  return w;
}

__global__ void __launch_bounds__(256) blend_kernel(const BlendArgs a) {
  const int OW = a.W * a.scale;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)OW * (a.row_end - a.row_begin)) return;
  const int X = (int)(idx % OW), Y = a.row_begin + (int)(idx / OW);
  const int th = a.tile < a.H ? a.tile : a.H, tw = a.tile < a.W ? a.tile : a.W;
  const int stride = a.tile - a.overlap;
  const int ts_h = th * a.scale, ts_w = tw * a.scale;
  const int oy = (a.overlap < th ? a.overlap : th) * a.scale, ox = (a.overlap < tw ? a.overlap : tw) * a.scale;
  float acc[3] = {0.f, 0.f, 0.f};
  float wsum = 0.0f;
  for (int iy = 0; iy < a.tiles_y; ++iy) {
    const int y0 = tile_origin(iy, stride, a.H, th) * a.scale;
    if (y0 > Y) break;
    if (Y >= y0 + ts_h) continue;
    const float wy = ramp_w(Y - y0, ts_h, oy, iy == 0, iy == a.tiles_y - 1);
    for (int ix = 0; ix < a.tiles_x; ++ix) {
      const int x0 = tile_origin(ix, stride, a.W, tw) * a.scale;
      if (x0 > X) break;
      if (X >= x0 + ts_w) continue;
      const float w = wy * ramp_w(X - x0, ts_w, ox, ix == 0, ix == a.tiles_x - 1);
      const float* t = a.tiles + (size_t)(iy * a.tiles_x + ix) * 3 * ts_h * ts_w + (size_t)(Y - y0) * ts_w + (X - x0);
      acc[0] += w * t[0];
      acc[1] += w * t[(size_t)ts_h * ts_w];
      acc[2] += w * t[2 * (size_t)ts_h * ts_w];
      wsum += w;
    }
  }
  uint8_t* o = a.out + ((size_t)Y * OW + X) * 3;
  const float inv = 1.0f / wsum;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = (uint8_t)fminf(fmaxf(rintf(acc[c] * inv * a.u8_scale), 0.0f), 255.0f);
}

int launch_blend(const BlendArgs& a, cudaStream_t s) {
  SSR_CHECK(a.row_begin >= 0 && a.row_end <= a.H * a.scale && a.row_begin <= a.row_end, SSR_E_INVALID, "blend: bad row band [%d, %d)",
            a.row_begin, a.row_end);
  const long long total = (long long)a.W * a.scale * (a.row_end - a.row_begin);
  if (total == 0) return SSR_OK;
  const double frac = (double)(a.row_end - a.row_begin) / (a.H * a.scale);
  ProfScope prof("blend", 0.0, (double)total * 3 + frac * a.tiles_x * a.tiles_y * 3 * 4 * a.tile * a.tile * a.scale * a.scale, s);
  blend_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// =============================================================================================
// layout utilities (op-level tests and model-independent glue)
// =============================================================================================
__global__ void pack_rows_kernel(const float* in, int rows, int cols, void* out, int ld, int elem, int rtf32) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * ld) return;
  const int c = (int)(idx % ld);
  const long long r = idx / ld;
  store_elem(out, (size_t)idx, elem, c < cols ? in[r * cols + c] : 0.0f, rtf32);
}
// =============================================================================================
// RCAN channel attention (common.py:156-170) around the second conv of an RCAB (rcan.py:21-24):
//   y = sigmoid(W2 relu(W1 mean_hw(t) + b1) + b2);  out = res + t * y
// Stage 1 reduces t over pixel slabs into partial[b][split][c] (deterministic, no atomics); stage 2
// finishes the mean, evaluates the tiny MLP once per block and applies gate + residual.
// =============================================================================================
__global__ void __launch_bounds__(256) ca_pool_kernel(const void* t, int elem_t, int ld, int HW, int C, int nsplit, float* partial) {
  __shared__ float red[256];
  const int b = blockIdx.y, split = blockIdx.x;
  const int chunk = (HW + nsplit - 1) / nsplit;
  const int p0 = split * chunk, p1 = min(HW, p0 + chunk);
  const int lanes = 256 / 64;  // pixel lanes per 64-channel slab
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int c = c0 + (threadIdx.x & 63), pl = threadIdx.x >> 6;
    float acc = 0.0f;
    if (c < C)
      for (int p = p0 + pl; p < p1; p += lanes) acc += load_elem(t, ((size_t)b * HW + p) * ld + c, elem_t == 2 ? 2 : 4);
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < 64 && c < C)
      partial[((size_t)b * nsplit + split) * C + c] = red[threadIdx.x] + red[threadIdx.x + 64] + red[threadIdx.x + 128] + red[threadIdx.x + 192];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) ca_apply_kernel(const CaArgs a) {
  __shared__ float pooled[256], hid[64], gate[256];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float s = 0.0f;
    for (int k = 0; k < a.nsplit; ++k) s += a.partial[((size_t)b * a.nsplit + k) * a.C + c];
    pooled[c] = s / (float)a.HW;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < a.R; r += blockDim.x) {
    float s = a.b1[r];
    for (int c = 0; c < a.C; ++c) s = fmaf(a.W1[r * a.C + c], pooled[c], s);
    hid[r] = fmaxf(s, 0.0f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.CP; c += blockDim.x) {
    float g = 0.0f;
    if (c < a.C) {
      float s = a.b2[c];
      for (int r = 0; r < a.R; ++r) s = fmaf(a.W2[c * a.R + r], hid[r], s);
      g = a.scale / (1.0f + expf(-s));
    }
    gate[c] = g;  // padded channels: t is zero there anyway
  }
  __syncthreads();
  if (a.save_gate && blockIdx.x == 0) {  // training forward: keep the MLP's intermediates for the backward
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      a.save_pool[(size_t)b * a.C + c] = pooled[c];
      a.save_gate[(size_t)b * a.C + c] = gate[c];
    }
    for (int r = threadIdx.x; r < a.R; r += blockDim.x) a.save_hid[(size_t)b * a.R + r] = hid[r];
  }
  const int chunk = (a.HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * chunk, p1 = min(a.HW, p0 + chunk);
  const int c4n = a.CP / 4;
  for (int e = threadIdx.x; e < (p1 - p0) * c4n; e += blockDim.x) {
    const int p = p0 + e / c4n, c = (e % c4n) * 4;
    const size_t m = (size_t)b * a.HW + p;
    float4 tv;
    if (a.elem_t == 2) {
      const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(a.t) + m * a.ld + c);
      tv = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
    } else {
      tv = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.t) + m * a.ld + c);
    }
    const float4 rv = *reinterpret_cast<const float4*>(a.res + m * a.ld + c);
    float4 o;
    o.x = fmaf(tv.x, gate[c], rv.x); o.y = fmaf(tv.y, gate[c + 1], rv.y);
    o.z = fmaf(tv.z, gate[c + 2], rv.z); o.w = fmaf(tv.w, gate[c + 3], rv.w);  // gate already carries a.scale
    *reinterpret_cast<float4*>(a.out_f32 + m * a.ld + c) = o;
    if (!a.out_T) continue;
    if (a.elem == 2) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.out_T) + m * a.ld_T + c) = pk;
    } else {
      float4 q = o;
      if (a.round_tf32) { q.x = round_tf32(q.x); q.y = round_tf32(q.y); q.z = round_tf32(q.z); q.w = round_tf32(q.w); }
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out_T) + m * a.ld_T + c) = q;
    }
  }
}

int launch_channel_attention(const CaArgs& a, cudaStream_t s) {
  SSR_CHECK(a.C <= 256 && a.R <= 64 && a.CP % 4 == 0 && a.nsplit >= 1, SSR_E_INVALID, "channel attention: C=%d R=%d", a.C, a.R);
  const double px = (double)a.B * a.HW;
  {
    ProfScope prof("ca_pool", 0.0, px * a.C * (a.elem_t == 2 ? 2 : 4), s);
    ca_pool_kernel<<<dim3(a.nsplit, a.B), 256, 0, s>>>(a.t, a.elem_t, a.ld, a.HW, a.C, a.nsplit, a.partial);
    count_launch();
  }
  {
    ProfScope prof("ca_apply", 0.0, px * a.C * ((a.elem_t == 2 ? 2 : 4) + 4 + 4 + (a.out_T ? a.elem : 0)), s);
    const int blocks = a.HW >= 4096 ? 64 : (a.HW + 63) / 64;
    ca_apply_kernel<<<dim3(blocks, a.B), 256, 0, s>>>(a);
    count_launch();
  }
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// =============================================================================================
// HAN (han.py): layer attention (LAM_Module :12-33) and channel-spatial attention (CSAM_Module :36-52)
// The N = 11 stacked feature maps live as N fp32 planes [B*HW][ld]; the flattened C*H*W axis of the reference is any fixed
// order of one image's (pixel, channel) pairs, so the pixel-major planes serve as they are.
// =============================================================================================
constexpr int HAN_N = 11, HAN_NP = HAN_N * (HAN_N + 1) / 2;
// stage 1: energy[b][n][m] = sum over one image of X_n X_m (upper triangle, fp64 accumulation across blocks)
__global__ void __launch_bounds__(256) han_gram_kernel(const float* stack, size_t plane, int ld, int HW, int C, double* energy) {
  __shared__ float red[8][HAN_NP];
  const int b = blockIdx.y;
  const size_t n_el = (size_t)HW * C;
  float acc[HAN_NP];
#pragma unroll
  for (int i = 0; i < HAN_NP; ++i) acc[i] = 0.0f;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (size_t)gridDim.x * blockDim.x) {
    const size_t p = e / C;
    const int c = (int)(e - p * C);
    const float* src = stack + ((size_t)b * HW + p) * ld + c;
    float v[HAN_N];
#pragma unroll
    for (int n = 0; n < HAN_N; ++n) v[n] = src[(size_t)n * plane];
    int k = 0;
#pragma unroll
    for (int n = 0; n < HAN_N; ++n)
#pragma unroll
      for (int m2 = n; m2 < HAN_N; ++m2) acc[k] = fmaf(v[n], v[m2], acc[k]), ++k;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < HAN_NP; ++i) {
    float a = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[warp][i] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HAN_NP; i += blockDim.x) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)red[w][i];
    atomicAdd(energy + (size_t)b * HAN_NP + i, t);
  }
}
// stage 2: attention = softmax(row max - energy) (han.py:24-26); out[n] = gamma * sum_m attention[n][m] X_m + X_n, written as the
// [pixel][n * C + c] operand of last_conv
__global__ void __launch_bounds__(256) han_lam_apply_kernel(const float* stack, size_t plane, int ld, int HW, int C, const double* energy,
                                                            const float* gamma, void* out, int ld_out, int elem, int rtf32) {
  __shared__ float att[HAN_N][HAN_N];
  const int b = blockIdx.y;
  if (threadIdx.x < HAN_N) {
    const int n = threadIdx.x;
    double e[HAN_N], mx = -1e300;
    for (int m2 = 0; m2 < HAN_N; ++m2) {
      const int lo = n < m2 ? n : m2, hi = n < m2 ? m2 : n;
      e[m2] = energy[(size_t)b * HAN_NP + lo * HAN_N - lo * (lo - 1) / 2 + (hi - lo)];
      mx = e[m2] > mx ? e[m2] : mx;
    }
    // energy_new = max - e >= 0; its softmax, shifted by its own maximum (= max - min e)
    double mn = 1e300, sum = 0.0;
    for (int m2 = 0; m2 < HAN_N; ++m2) mn = e[m2] < mn ? e[m2] : mn;
    for (int m2 = 0; m2 < HAN_N; ++m2) {
      e[m2] = exp((mx - e[m2]) - (mx - mn));
      sum += e[m2];
    }
    for (int m2 = 0; m2 < HAN_N; ++m2) att[n][m2] = (float)(e[m2] / sum);
  }
  __syncthreads();
  const float g = gamma[0];
  const size_t n_el = (size_t)HW * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (size_t)gridDim.x * blockDim.x) {
    const size_t p = e / C;
    const int c = (int)(e - p * C);
    const size_t row = (size_t)b * HW + p;
    const float* src = stack + row * ld + c;
    float v[HAN_N];
#pragma unroll
    for (int n = 0; n < HAN_N; ++n) v[n] = src[(size_t)n * plane];
#pragma unroll
    for (int n = 0; n < HAN_N; ++n) {
      float a = 0.0f;
#pragma unroll
      for (int m2 = 0; m2 < HAN_N; ++m2) a = fmaf(att[n][m2], v[m2], a);
      store_elem(out, row * ld_out + (size_t)n * C + c, elem, fmaf(g, a, v[n]), rtf32);
    }
  }
}
int launch_han_lam(const float* stack, size_t plane, int ld, int B, int HW, int C, double* energy, const float* gamma, void* out,
                   int ld_out, int elem, int rtf32, cudaStream_t s) {
  SSR_CUDA(cudaMemsetAsync(energy, 0, (size_t)B * HAN_NP * sizeof(double), s));
  const int blocks = (int)std::min<size_t>(((size_t)HW * C + 255) / 256, 256);
  ProfScope prof("han_lam", 2.0 * B * HW * C * (HAN_NP + HAN_N * HAN_N), (double)B * HW * C * (2.0 * HAN_N * 4 + HAN_N * elem), s);
  han_gram_kernel<<<dim3(blocks, B), 256, 0, s>>>(stack, plane, ld, HW, C, energy);
  count_launch();
  han_lam_apply_kernel<<<dim3(blocks, B), 256, 0, s>>>(stack, plane, ld, HW, C, energy, gamma, out, ld_out, elem, rtf32);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}
// CSAM: s = sigmoid(Conv3d(1, 1, 3, padding 1) over the (C, H, W) volume); out = x * (gamma * s) + x
__global__ void __launch_bounds__(256) han_csam_kernel(const float* x, int ld, int B, int H, int W, int C, const float* w27, const float* bias,
                                                       const float* gamma, void* out, int ld_out, int elem, int rtf32) {
  const size_t n_el = (size_t)B * H * W * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const size_t p = e / C;
    const int xx = (int)(p % W), yy = (int)((p / W) % H), b = (int)(p / ((size_t)W * H));
    float a = bias[0];
    for (int dc = 0; dc < 3; ++dc) {
      const int cc = c + dc - 1;
      if (cc < 0 || cc >= C) continue;
      for (int dy = 0; dy < 3; ++dy) {
        const int y2 = yy + dy - 1;
        if (y2 < 0 || y2 >= H) continue;
        for (int dx = 0; dx < 3; ++dx) {
          const int x2 = xx + dx - 1;
          if (x2 < 0 || x2 >= W) continue;
          a = fmaf(w27[(dc * 3 + dy) * 3 + dx], x[(((size_t)b * H + y2) * W + x2) * ld + cc], a);
        }
      }
    }
    const float v = x[p * ld + c];
    const float sgm = 1.0f / (1.0f + expf(-a));
    store_elem(out, p * ld_out + c, elem, fmaf(v, gamma[0] * sgm, v), rtf32);
  }
}
int launch_han_csam(const float* x, int ld, int B, int H, int W, int C, const float* w27, const float* bias, const float* gamma,
                    void* out, int ld_out, int elem, int rtf32, cudaStream_t s) {
  const size_t n = (size_t)B * H * W * C;
  ProfScope prof("han_csam", 2.0 * 27 * n, (double)n * (4 + elem), s);
  han_csam_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 32), 256, 0, s>>>(x, ld, B, H, W, C, w27, bias, gamma, out, ld_out, elem,
                                                                                      rtf32);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// ---- HAN backward (training) ----
// LAM: out_n = gamma * sum_m A[n][m] X_m + X_n, A = softmax(max - E), E = X X^T.  With D[n][m] = <dOut_n, X_m>:
//   dgamma = sum A (.) D;  dA = gamma D;  dEnew = A (.) (dA - rowsum(A (.) dA));  dE = -dEnew (the row-max term cancels: softmax
//   gradients sum to zero over a row);  dX_n = dOut_n + gamma sum_m A[m][n] dOut_m + sum_m (dE[n][m] + dE[m][n]) X_m.
// stage 1: D[b][n][m] (blockIdx.z = n), fp64 accumulation across blocks.  dOut is [pixel][n * C + c].
__global__ void __launch_bounds__(256) han_gram2_kernel(const float* dout, int ld_d, const float* stack, size_t plane, int ld, int HW, int C,
                                                        double* D) {
  __shared__ float red[8][HAN_N];
  const int b = blockIdx.y, n = blockIdx.z;
  const size_t n_el = (size_t)HW * C;
  float acc[HAN_N];
#pragma unroll
  for (int i = 0; i < HAN_N; ++i) acc[i] = 0.0f;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (size_t)gridDim.x * blockDim.x) {
    const size_t p = e / C;
    const int c = (int)(e - p * C);
    const size_t row = (size_t)b * HW + p;
    const float g = dout[row * ld_d + (size_t)n * C + c];
    const float* src = stack + row * ld + c;
#pragma unroll
    for (int m2 = 0; m2 < HAN_N; ++m2) acc[m2] = fmaf(g, src[(size_t)m2 * plane], acc[m2]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < HAN_N; ++i) {
    float a = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[warp][i] = a;
  }
  __syncthreads();
  if (threadIdx.x < HAN_N) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)red[w][threadIdx.x];
    atomicAdd(D + ((size_t)b * HAN_N + n) * HAN_N + threadIdx.x, t);
  }
}
// stage 2 (one block per image): the 11 x 11 algebra.  coef[b][0][n][m] = gamma A[m][n], coef[b][1][n][m] = dE[n][m] + dE[m][n]
__global__ void han_lam_bwd_small_kernel(const double* energy, const double* D, const float* gamma, float* coef, float* dgamma) {
  __shared__ double A[HAN_N][HAN_N], dE[HAN_N][HAN_N];
  const int b = blockIdx.x, n = threadIdx.x;
  const double g = (double)gamma[0];
  if (n < HAN_N) {
    double e[HAN_N], mx = -1e300, mn = 1e300, sum = 0.0;
    for (int m2 = 0; m2 < HAN_N; ++m2) {
      const int lo = n < m2 ? n : m2, hi = n < m2 ? m2 : n;
      e[m2] = energy[(size_t)b * HAN_NP + lo * HAN_N - lo * (lo - 1) / 2 + (hi - lo)];
      mx = e[m2] > mx ? e[m2] : mx;
      mn = e[m2] < mn ? e[m2] : mn;
    }
    for (int m2 = 0; m2 < HAN_N; ++m2) {
      e[m2] = exp((mx - e[m2]) - (mx - mn));
      sum += e[m2];
    }
    double dot = 0.0, dg = 0.0;
    for (int m2 = 0; m2 < HAN_N; ++m2) {
      A[n][m2] = e[m2] / sum;
      const double d = D[((size_t)b * HAN_N + n) * HAN_N + m2];
      dg += A[n][m2] * d;
      dot += A[n][m2] * g * d;
    }
    for (int m2 = 0; m2 < HAN_N; ++m2) dE[n][m2] = -A[n][m2] * (g * D[((size_t)b * HAN_N + n) * HAN_N + m2] - dot);
    atomicAdd(dgamma, (float)dg);
  }
  __syncthreads();
  if (n < HAN_N)
    for (int m2 = 0; m2 < HAN_N; ++m2) {
      coef[((size_t)b * 2 + 0) * HAN_N * HAN_N + n * HAN_N + m2] = (float)(g * A[m2][n]);
      coef[((size_t)b * 2 + 1) * HAN_N * HAN_N + n * HAN_N + m2] = (float)(dE[n][m2] + dE[m2][n]);
    }
}
// stage 3: dX planes (fp32 [11][B*HW][ld])
__global__ void __launch_bounds__(256) han_lam_bwd_apply_kernel(const float* dout, int ld_d, const float* stack, size_t plane, int ld, int HW,
                                                                int C, const float* coef, float* dstack) {
  __shared__ float cf[2][HAN_N][HAN_N];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * HAN_N * HAN_N; i += blockDim.x) (&cf[0][0][0])[i] = coef[(size_t)b * 2 * HAN_N * HAN_N + i];
  __syncthreads();
  const size_t n_el = (size_t)HW * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (size_t)gridDim.x * blockDim.x) {
    const size_t p = e / C;
    const int c = (int)(e - p * C);
    const size_t row = (size_t)b * HW + p;
    float g[HAN_N], v[HAN_N];
#pragma unroll
    for (int m2 = 0; m2 < HAN_N; ++m2) {
      g[m2] = dout[row * ld_d + (size_t)m2 * C + c];
      v[m2] = stack[row * ld + c + (size_t)m2 * plane];
    }
#pragma unroll
    for (int n = 0; n < HAN_N; ++n) {
      float a = g[n];
#pragma unroll
      for (int m2 = 0; m2 < HAN_N; ++m2) a = fmaf(cf[0][n][m2], g[m2], fmaf(cf[1][n][m2], v[m2], a));
      dstack[(size_t)n * plane + row * ld + c] = a;
    }
  }
}
int launch_han_lam_bwd(const float* dout, int ld_d, const float* stack, size_t plane, int ld, int B, int HW, int C, const double* energy,
                       const float* gamma, double* D, float* coef, float* dgamma, float* dstack, cudaStream_t s) {
  SSR_CUDA(cudaMemsetAsync(D, 0, (size_t)B * HAN_N * HAN_N * sizeof(double), s));
  SSR_CHECK(dgamma != nullptr, SSR_E_INVALID, "han_lam_bwd: dgamma scratch missing");
  SSR_CUDA(cudaMemsetAsync(dgamma, 0, 4, s));
  const int blocks = (int)std::min<size_t>(((size_t)HW * C + 255) / 256, 128);
  ProfScope prof("han_lam_bwd", 0.0, (double)B * HW * C * 4.0 * 4 * HAN_N, s);
  han_gram2_kernel<<<dim3(blocks, B, HAN_N), 256, 0, s>>>(dout, ld_d, stack, plane, ld, HW, C, D);
  count_launch();
  han_lam_bwd_small_kernel<<<B, 32, 0, s>>>(energy, D, gamma, coef, dgamma);
  count_launch();
  han_lam_bwd_apply_kernel<<<dim3(blocks, B), 256, 0, s>>>(dout, ld_d, stack, plane, ld, HW, C, coef, dstack);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}
// CSAM backward.  out = x (1 + gamma s), s = sigmoid(pre), pre = conv3d(x) + b.  With g = dL/dout:
//   dpre = g x gamma s (1 - s);  dx = g (1 + gamma s) + conv3d^T(dpre);  dgamma = sum g x s;  db = sum dpre;  dW[tap] = sum dpre x[. + off(tap)]
// pass 1: dpre, the direct part of dx, and the 29 scalar sums (scal: [0] dgamma, [1] db, [2..29) dW)
__global__ void __launch_bounds__(256) han_csam_bwd1_kernel(const float* x, int ld, const float* g, int ld_g, int B, int H, int W, int C,
                                                            const float* w27, const float* bias, const float* gamma, float* dpre, float* dxd,
                                                            float* scal) {
  __shared__ float red[8][29];
  float acc[29];
#pragma unroll
  for (int i = 0; i < 29; ++i) acc[i] = 0.0f;
  const float gm = gamma[0];
  const size_t n_el = (size_t)B * H * W * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const size_t p = e / C;
    const int xx = (int)(p % W), yy = (int)((p / W) % H), b = (int)(p / ((size_t)W * H));
    float nb[27];
    float pre = bias[0];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int cc = c + t / 9 - 1, y2 = yy + (t / 3) % 3 - 1, x2 = xx + t % 3 - 1;
      nb[t] = (cc >= 0 && cc < C && y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) ? x[(((size_t)b * H + y2) * W + x2) * ld + cc] : 0.0f;
      pre = fmaf(w27[t], nb[t], pre);
    }
    const float v = nb[13], gg = g[p * ld_g + c];
    const float sg = 1.0f / (1.0f + expf(-pre));
    const float dp = gg * v * gm * sg * (1.0f - sg);
    dpre[p * ld + c] = dp;
    dxd[p * ld + c] = gg * fmaf(gm, sg, 1.0f);
    acc[0] = fmaf(gg * v, sg, acc[0]);
    acc[1] += dp;
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[2 + t] = fmaf(dp, nb[t], acc[2 + t]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 29; ++i) {
    float a = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[warp][i] = a;
  }
  __syncthreads();
  if (threadIdx.x < 29) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(scal + threadIdx.x, t);
  }
}
// pass 2: G0 = acc_in (the LAM's gradient of the same plane) + dxd + conv3d^T(dpre); fp32 + bf16 copy
__global__ void __launch_bounds__(256) han_csam_bwd2_kernel(const float* dpre, const float* dxd, const float* acc_in, int ld, int B, int H, int W,
                                                            int C, const float* w27, float* out, __nv_bfloat16* out_bf) {
  const size_t n_el = (size_t)B * H * W * C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const size_t p = e / C;
    const int xx = (int)(p % W), yy = (int)((p / W) % H), b = (int)(p / ((size_t)W * H));
    float a = dxd[p * ld + c] + acc_in[p * ld + c];
#pragma unroll
    for (int t = 0; t < 27; ++t) {  // x[q] fed pre[q - off(t)] with weight w[t]
      const int cc = c - (t / 9 - 1), y2 = yy - ((t / 3) % 3 - 1), x2 = xx - (t % 3 - 1);
      if (cc >= 0 && cc < C && y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) a = fmaf(w27[t], dpre[(((size_t)b * H + y2) * W + x2) * ld + cc], a);
    }
    out[p * ld + c] = a;
    out_bf[p * ld + c] = __float2bfloat16_rn(a);
  }
}
int launch_han_csam_bwd(const float* x, int ld, const float* g, int ld_g, const float* acc_in, int B, int H, int W, int C, const float* w27,
                        const float* bias, const float* gamma, float* dpre, float* dxd, float* scal, float* out, void* out_bf,
                        cudaStream_t s) {
  SSR_CUDA(cudaMemsetAsync(scal, 0, 29 * 4, s));
  const size_t n = (size_t)B * H * W * C;
  const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
  ProfScope prof("han_csam_bwd", 4.0 * 27 * n, (double)n * 4 * 6, s);
  han_csam_bwd1_kernel<<<grid, 256, 0, s>>>(x, ld, g, ld_g, B, H, W, C, w27, bias, gamma, dpre, dxd, scal);
  count_launch();
  han_csam_bwd2_kernel<<<grid, 256, 0, s>>>(dpre, dxd, acc_in, ld, B, H, W, C, w27, out, reinterpret_cast<__nv_bfloat16*>(out_bf));
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// ---- channel attention backward (training; see CaBwdArgs) ----
// stage 1: partial[b][split][c] = sum over a pixel slab of G * t
__global__ void __launch_bounds__(256) ca_bwd_reduce_kernel(const float* G, const __nv_bfloat16* t, int ld, int HW, int C, int nsplit,
                                                            float* partial) {
  __shared__ float red[256];
  const int b = blockIdx.y, split = blockIdx.x;
  const int chunk = (HW + nsplit - 1) / nsplit;
  const int p0 = split * chunk, p1 = min(HW, p0 + chunk);
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int c = c0 + (threadIdx.x & 63), pl = threadIdx.x >> 6;
    float acc = 0.0f;
    if (c < C)
      for (int p = p0 + pl; p < p1; p += 4) {
        const size_t i = ((size_t)b * HW + p) * ld + c;
        acc = fmaf(G[i], __bfloat162float(t[i]), acc);
      }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < 64 && c < C)
      partial[((size_t)b * nsplit + split) * C + c] = red[threadIdx.x] + red[threadIdx.x + 64] + red[threadIdx.x + 128] + red[threadIdx.x + 192];
    __syncthreads();
  }
}
// stage 2 (one block; the samples are walked in order, so the parameter gradients are deterministic): the gate MLP's backward
__global__ void __launch_bounds__(256) ca_bwd_gate_kernel(const CaBwdArgs a) {
  __shared__ float dz2[256], dz1[64];
  for (int b = 0; b < a.B; ++b) {
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      float ds = 0.0f;
      for (int k = 0; k < a.nsplit; ++k) ds += a.partial[((size_t)b * a.nsplit + k) * a.C + c];
      const float g = a.gate[(size_t)b * a.C + c];
      dz2[c] = ds * g * (1.0f - g);
    }
    __syncthreads();
    for (int r = threadIdx.x; r < a.R; r += blockDim.x) {
      float dh = 0.0f;
      for (int c = 0; c < a.C; ++c) dh = fmaf(a.W2[c * a.R + r], dz2[c], dh);
      dz1[r] = a.hid[(size_t)b * a.R + r] > 0.0f ? dh : 0.0f;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      float dp = 0.0f;
      for (int r = 0; r < a.R; ++r) dp = fmaf(a.W1[r * a.C + c], dz1[r], dp);
      a.dpool[(size_t)b * a.C + c] = dp / (float)a.HW;
      if (a.db2) a.db2[c] = (b ? a.db2[c] : 0.0f) + dz2[c];
    }
    if (a.db1)
      for (int r = threadIdx.x; r < a.R; r += blockDim.x) a.db1[r] = (b ? a.db1[r] : 0.0f) + dz1[r];
    for (int e = threadIdx.x; e < a.C * a.R; e += blockDim.x) {  // every entry is owned by one thread across all samples
      if (a.dW2) {
        const int c = e / a.R, r = e - c * a.R;
        a.dW2[e] = (b ? a.dW2[e] : 0.0f) + dz2[c] * a.hid[(size_t)b * a.R + r];
      }
      if (a.dW1) {
        const int r = e / a.C, c = e - r * a.C;
        a.dW1[e] = (b ? a.dW1[e] : 0.0f) + dz1[r] * a.pool[(size_t)b * a.C + c];
      }
    }
    __syncthreads();
  }
}
// stage 3: dt = G * gate + dpool (bf16; padded channels zero)
__global__ void __launch_bounds__(256) ca_bwd_apply_kernel(const CaBwdArgs a) {
  const int c4n = a.CP / 4;
  const size_t n = (size_t)a.B * a.HW * c4n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const size_t m = e / c4n;
    const int c = (int)(e - m * c4n) * 4, b = (int)(m / a.HW);
    const float4 g = *reinterpret_cast<const float4*>(a.G + m * a.ld + c);
    float o[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = c + i < a.C ? fmaf(o[i], a.gate[(size_t)b * a.C + c + i], a.dpool[(size_t)b * a.C + c + i]) : 0.0f;
    uint2 pk;
    pk.x = pack_bf16x2(o[0], o[1]);
    pk.y = pack_bf16x2(o[2], o[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.dt) + m * a.ld_dt + c) = pk;
  }
}
int launch_channel_attention_bwd(const CaBwdArgs& a, cudaStream_t s) {
  SSR_CHECK(a.C <= 256 && a.R <= 64 && a.CP % 4 == 0 && a.nsplit >= 1 && a.ld % 4 == 0 && a.ld_dt % 4 == 0, SSR_E_INVALID,
            "channel attention backward: C=%d R=%d", a.C, a.R);
  const double px = (double)a.B * a.HW;
  {
    ProfScope prof("ca_bwd", 0.0, px * a.C * (4 + 2 + 4 + 2), s);
    ca_bwd_reduce_kernel<<<dim3(a.nsplit, a.B), 256, 0, s>>>(a.G, reinterpret_cast<const __nv_bfloat16*>(a.t), a.ld, a.HW, a.C, a.nsplit,
                                                            a.partial);
    count_launch();
    ca_bwd_gate_kernel<<<1, 256, 0, s>>>(a);
    count_launch();
    const size_t n = (size_t)a.B * a.HW * (a.CP / 4);
    ca_bwd_apply_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, s>>>(a);
    count_launch();
  }
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// "pixelshuffledirect" tail of the lightweight SwinIR (swinir.py:327-329, 367-372): the single conv C -> 3 s^2 leaves
// fp32 rows [pixel][3 s^2]; this kernel is nn.PixelShuffle(s) + un-normalise + crop + fp32 NCHW / uint8 HWC store.
__global__ void shuffle_finish_kernel(const float* in, int ld, int B, int Hp, int Wp, int scale, int ch, int cw, float s0, float s1,
                                      float s2, float out_scale, float u8_scale, float* out_f32, uint8_t* out_u8) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * ch * cw) return;
  const int X = (int)(idx % cw), Y = (int)((idx / cw) % ch), b = (int)(idx / ((long long)cw * ch));
  const int y = Y / scale, x = X / scale, q = (Y % scale) * scale + X % scale, ss = scale * scale;
  const float* row = in + ((size_t)(b * Hp + y) * Wp + x) * ld;
  const float sh[3] = {s0, s1, s2};
  float r[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) r[c] = (row[c * ss + q] + sh[c]) * out_scale;
  if (out_f32) {
    const size_t plane = (size_t)ch * cw;
    float* o = out_f32 + (size_t)b * 3 * plane + (size_t)Y * cw + X;
    o[0] = r[0];
    o[plane] = r[1];
    o[2 * plane] = r[2];
  }
  if (out_u8) {
    uint8_t* o = out_u8 + ((size_t)(b * ch + Y) * cw + X) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (uint8_t)fminf(fmaxf(rintf(r[c] * u8_scale), 0.0f), 255.0f);
  }
}
int launch_shuffle_finish(const float* in, int ld, int B, int Hp, int Wp, int scale, int ch, int cw, const float* shift3,
                          float out_scale, float u8_scale, float* out_f32, uint8_t* out_u8, cudaStream_t s) {
  const long long total = (long long)B * ch * cw;
  ProfScope prof("shuffle_finish", 0.0, (double)total * 3 * (4 + (out_f32 ? 4 : 1)), s);
  shuffle_finish_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(in, ld, B, Hp, Wp, scale, ch, cw, shift3[0], shift3[1], shift3[2],
                                                                  out_scale, u8_scale, out_f32, out_u8);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// Fold a LayerNorm affine into the Linear that consumes it (op-level tests; the model does this on the host at
// pack time): Wf[n][k] = W[n][k] * gamma[k],  bf[n] = b[n] + sum_k W[n][k] * beta[k].
__global__ void fold_ln_linear_kernel(const float* W, const float* b, const float* gamma, const float* beta, float* Wf, float* bf,
                                      int N, int K) {
  const int n = blockIdx.x;
  float acc = 0.0f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float w = W[(size_t)n * K + k];
    Wf[(size_t)n * K + k] = w * gamma[k];
    acc += w * beta[k];
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = b[n];
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    bf[n] = t;
  }
}
int launch_fold_ln_linear(const float* W, const float* b, const float* gamma, const float* beta, float* Wf, float* bf, int N, int K,
                          cudaStream_t s) {
  fold_ln_linear_kernel<<<N, 128, 0, s>>>(W, b, gamma, beta, Wf, bf, N, K);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

int launch_pack_rows(const float* in, int rows, int cols, void* out, int ld, int elem, int rtf32, cudaStream_t s) {
  const long long total = (long long)rows * ld;
  pack_rows_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(in, rows, cols, out, ld, elem, rtf32);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

__global__ void unpack_rows_kernel(const void* in, int ld, int elem, float* out, int rows, int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int c = (int)(idx % cols);
  const long long r = idx / cols;
  out[idx] = load_elem(in, (size_t)(r * ld + c), elem);
}
int launch_unpack_rows(const void* in, int ld, int elem, float* out, int rows, int cols, cudaStream_t s) {
  const long long total = (long long)rows * cols;
  unpack_rows_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(in, ld, elem, out, rows, cols);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

int launch_fill_zero(void* p, size_t bytes, cudaStream_t s) {
  SSR_CUDA(cudaMemsetAsync(p, 0, bytes, s));
  return SSR_OK;
}

// reference qkv layout [M][3][heads][d] (fp32, un-scaled) -> padded [M][3][heads][DP], q scaled
__global__ void repack_qkv_kernel(const float* qkv, void* out, int M, int C, int heads, int DP, int QP, float qscale,
                                  int elem, int rtf32) {
  const int d = C / heads;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * 3 * QP) return;
  const int col = (int)(idx % (3 * QP));
  const long long m = idx / (3 * QP);
  const int part = col / QP, hc = col % QP, h = hc / DP, j = hc % DP;
  float v = 0.0f;
  if (h < heads && j < d) {
    v = qkv[m * 3 * C + part * C + h * d + j];
    if (part == 0) v *= qscale;
  }
  store_elem(out, (size_t)idx, elem, v, rtf32);
}
int launch_repack_qkv(const float* qkv, void* out, int M, int C, int heads, int DP, int QP, float qscale, int elem,
                      int rtf32, cudaStream_t s) {
  const long long total = (long long)M * 3 * QP;
  repack_qkv_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(qkv, out, M, C, heads, DP, QP, qscale, elem, rtf32);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

__global__ void unpack_heads_kernel(const void* o, int ld, int elem, float* out, int M, int heads, int d, int DP) {
  const int C = heads * d;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * C) return;
  const int c = (int)(idx % C);
  const long long m = idx / C;
  out[idx] = load_elem(o, (size_t)(m * ld + (c / d) * DP + c % d), elem);
}
int launch_unpack_heads(const void* o, int ld, int elem, float* out, int M, int heads, int d, int DP, cudaStream_t s) {
  const long long total = (long long)M * heads * d;
  unpack_heads_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(o, ld, elem, out, M, heads, d, DP);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

// reference channel order [M][heads*d] -> padded head layout [M][ld] (column = head*DP + j), zero pads
__global__ void pack_heads_kernel(const float* in, void* out, int M, int heads, int d, int DP, int ld, int elem, int rtf32) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * ld) return;
  const int c = (int)(idx % ld);
  const long long m = idx / ld;
  const int h = c / DP, j = c % DP;
  store_elem(out, (size_t)idx, elem, (h < heads && j < d) ? in[m * heads * d + h * d + j] : 0.0f, rtf32);
}
int launch_pack_heads(const float* in, void* out, int M, int heads, int d, int DP, int ld, int elem, int rtf32,
                      cudaStream_t s) {
  const long long total = (long long)M * ld;
  pack_heads_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(in, out, M, heads, d, DP, ld, elem, rtf32);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

__global__ void nchw_to_nhwc_kernel(const float* in, void* out, int B, int C, int H, int W, int ld, int elem, int rtf32) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * ld) return;
  const int c = (int)(idx % ld);
  const long long p = idx / ld;
  const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
  store_elem(out, (size_t)idx, elem, c < C ? in[((size_t)(b * C + c) * H + y) * W + x] : 0.0f, rtf32);
}
int launch_nchw_to_nhwc(const float* in, void* out, int B, int C, int H, int W, int ld, int elem, int rtf32,
                        cudaStream_t s) {
  const long long total = (long long)B * H * W * ld;
  nchw_to_nhwc_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(in, out, B, C, H, W, ld, elem, rtf32);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

__global__ void nhwc_to_nchw_kernel(const void* in, int ld, int elem, float* out, int B, int C, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * C * H * W) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H), c = (int)((idx / ((long long)W * H)) % C);
  const int b = (int)(idx / ((long long)W * H * C));
  out[idx] = load_elem(in, ((size_t)(b * H + y) * W + x) * ld + c, elem);
}
int launch_nhwc_to_nchw(const void* in, int ld, int elem, float* out, int B, int C, int H, int W, cudaStream_t s) {
  const long long total = (long long)B * C * H * W;
  nhwc_to_nchw_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(in, ld, elem, out, B, C, H, W);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
