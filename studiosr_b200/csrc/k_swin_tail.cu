// Fused tail of a Swin block (bf16, C=180 -> padded 192, hidden 360 -> 384), one persistent kernel:
//
//   t'  = o @ Wproj^T + bproj + res                (swinir.py:103,171)    [tcgen05 SS, acc in TMEM]
//   xn2 = LayerNorm2(t')                           (swinir.py:172)        [epilogue -> smem A operand]
//   h   = GELU(xn2 @ W1^T + b1)                    (common.py:185-186)    [3 chunks of 128 hidden units]
//   t'' = t' + h @ W2^T + b2                       (common.py:188, swinir.py:172)
//   out: t'' (fp32 residual stream), LayerNorm_next(t'') or a bf16 copy of t''
//
// Design notes (all measured on B200, profiles/r01_micro_tc.txt):
//  * h never touches shared memory: the GELU epilogue packs it to bf16 and writes it back over its own fc1
//    accumulator columns with tcgen05.st; fc2 consumes it as the TMEM A operand of a TS-mode tcgen05.mma.
//  * Every bulk transfer is TMA.  The fp32 residual arrives in per-warp [32 x 32] SWIZZLE_128B boxes (thread <-> row
//    reads are bank-conflict free) prefetched two chunks ahead; the outputs leave as per-warp TMA stores from a
//    swizzled staging box.  No thread ever waits on a global load, and the ragged last tile is clipped by TMA.
//  * A single producer thread sustains only one wait->issue round per ~500 cycles however deep the ring is, so the
//    weight stream is split over two producer warps (even / odd ring entries) and the o tile has its own.
//  * LayerNorm statistics are one-pass (sum, sum of squares in fp32) on values held in registers.
//
// TMEM (512 columns): Y = [0,192) fc2 accumulator (pre-loaded with t' + b2); [192,384) projection accumulator,
// re-used as the fc1 chunk accumulators X0 = [192,320), X1 = [320,448).
// Warps: 0,1 = weight TMA producers, 2 = o-tile producer + L2 prefetch of the next residual tile, 3 = MMA issuer,
// 4..11 = epilogue; epilogue warp w owns TMEM lane quadrant (w % 4) and column half (w - 4) / 4.
#include "ssr_tc.cuh"

namespace ssr {

constexpr int ST_THREADS = 384;
constexpr int ST_EPI_WARP0 = 4;
constexpr int ST_WSLOTS = 3;
constexpr int ST_WENTRIES = 18;               // weight ring entries per tile
constexpr uint32_t ST_TILE = 16384;           // one [128 rows][128 B] k-block tile
constexpr uint32_t ST_WSLOT = 192 * 128;      // weight ring slot (fc1 entries use 128 rows of it)
constexpr uint32_t ST_OFF_OX = 0;                                  // 3 tiles: o, later xn2
constexpr uint32_t ST_OFF_W = ST_OFF_OX + 3 * ST_TILE;             // weight ring
constexpr uint32_t ST_OFF_IO = ST_OFF_W + ST_WSLOTS * ST_WSLOT;    // per epilogue warp: R0, R1 (residual), S (staging)
constexpr uint32_t ST_IO_WARP = 3 * 4096;
constexpr uint32_t ST_OFF_PAR = ST_OFF_IO + 8 * ST_IO_WARP;        // fp32 parameters
constexpr int ST_NPAR = 192 * 6 + 384;                             // bp b2 g2 be2 g3 be3 | b1
constexpr uint32_t ST_OFF_RED = ST_OFF_PAR + ST_NPAR * 4;          // [128][2] float2 cross-half reductions
constexpr uint32_t ST_OFF_BAR = ST_OFF_RED + 128 * 2 * 8;
constexpr uint32_t ST_SMEM = ST_OFF_BAR + 512 + 1024;
static_assert(ST_SMEM <= 232448, "fused tail kernel exceeds the 227 KB shared-memory limit");

enum {  // mbarrier indices
  SB_WFULL = 0,                       // [ST_WSLOTS]
  SB_WEMPTY = SB_WFULL + ST_WSLOTS,   // [ST_WSLOTS]
  SB_OFULL = SB_WEMPTY + ST_WSLOTS,
  SB_OEMPTY,
  SB_PFULL,    // projection accumulator complete
  SB_XNREADY,  // xn2 in smem + Y initialised (8 arrivals)
  SB_XFULL,    // [2] fc1 chunk accumulator complete
  SB_HREADY = SB_XFULL + 2,  // [2] h chunk packed into TMEM (8 arrivals)
  SB_YFULL = SB_HREADY + 2,
  SB_RFULL,    // [8 warps][2] residual chunk landed
  SB_COUNT = SB_RFULL + 16
};

struct TailArgs {
  int M, C, n_tiles;
  const float *bp, *b1, *b2, *g2, *be2, *g3, *be3;
  int has_f32;  // store t'' as fp32
  int has_bf;   // store a bf16 tensor: LayerNorm_next(t'') if do_ln else t''
  int do_ln;
  float eps;
  long long* dbg;  // optional phase timestamps (developer diagnostics): [CTA][tile < 32][16]
};

// byte offset of (row r, 16-byte chunk j) inside a [rows][128 B] SWIZZLE_128B box / [rows][64 B] SWIZZLE_64B box
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t sw64_off(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// GELU(x) = x * Phi(x), Phi(x) ~ 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4))): max |err| 1.3e-4 against the erf form
// (relative 7.5e-4 for x > -3), one MUFU + 7 FMA-pipe instructions.  x^2 is clamped where the fit's x^4 term would
// turn the polynomial around (|x| > 6: Phi is 0 / 1 to nine digits anyway).
__device__ __forceinline__ float gelu_tanh3(float x) {
  const float x2 = fminf(x * x, 36.0f);
  float p = fmaf(x2, -3.99928615e-04f, 3.74196843e-02f);
  p = fmaf(p, x2, 7.96738290e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

__global__ void __launch_bounds__(ST_THREADS, 1)
swin_tail_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWp,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmResPf,
                 const __grid_constant__ CUtensorMap tmOutF, const __grid_constant__ CUtensorMap tmOutB, const TailArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* par = reinterpret_cast<float*>(smem + ST_OFF_PAR);
  float *s_bp = par, *s_b2 = par + 192, *s_g2 = par + 384, *s_be2 = par + 576, *s_g3 = par + 768, *s_be3 = par + 960,
        *s_b1 = par + 1152;
  float2* red = reinterpret_cast<float2*>(smem + ST_OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + SB_COUNT);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 192; i += ST_THREADS) {
    s_bp[i] = __ldg(a.bp + i);
    s_b2[i] = __ldg(a.b2 + i);
    s_g2[i] = __ldg(a.g2 + i);
    s_be2[i] = __ldg(a.be2 + i);
    s_g3[i] = a.g3 ? __ldg(a.g3 + i) : 0.f;
    s_be3[i] = a.be3 ? __ldg(a.be3 + i) : 0.f;
  }
  for (int i = threadIdx.x; i < 384; i += ST_THREADS) s_b1[i] = __ldg(a.b1 + i);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmO);
    prefetch_tmap(&tmWp);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    prefetch_tmap(&tmRes);
    prefetch_tmap(&tmResPf);
    prefetch_tmap(&tmOutF);
    prefetch_tmap(&tmOutB);
    for (int i = 0; i < SB_COUNT; ++i) {
      const bool epi8 = (i == SB_XNREADY) || (i >= SB_HREADY && i < SB_HREADY + 2);
      mbar_init(bar(i), epi8 ? 8 : 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 3) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = a.n_tiles;
  constexpr uint32_t IDESC_192 = umma_idesc(1, 128, 192), IDESC_128 = umma_idesc(1, 128, 128);

  if (warp < 2) {
    // =========================== weight producers (even / odd ring entries) ===========================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        for (int e = warp; e < ST_WENTRIES; e += 2) {
          // ring entries in exactly the order the MMA warp consumes them:
          //  0-2 proj | 3-5 fc1 c0 | 6-8 fc1 c1 | 9-10 fc2 c0 | 11-13 fc1 c2 | 14-15 fc2 c1 | 16-17 fc2 c2
          const CUtensorMap* map;
          uint32_t bytes;
          int c0, c1;
          if (e < 3) {
            map = &tmWp; bytes = 192 * 128; c0 = e * 64; c1 = 0;
          } else if (e < 9) {
            map = &tmW1; bytes = 128 * 128; c0 = ((e - 3) % 3) * 64; c1 = ((e - 3) / 3) * 128;
          } else if (e < 11) {
            map = &tmW2; bytes = 192 * 128; c0 = (e - 9) * 64; c1 = 0;
          } else if (e < 14) {
            map = &tmW1; bytes = 128 * 128; c0 = (e - 11) * 64; c1 = 256;
          } else {
            map = &tmW2; bytes = 192 * 128; c0 = 128 + (e - 14) * 64; c1 = 0;
          }
          const uint32_t k = (uint32_t)it * ST_WENTRIES + e;
          const int s = k % ST_WSLOTS;
          mbar_wait(bar(SB_WEMPTY + s), ((k / ST_WSLOTS) & 1u) ^ 1u);
          mbar_expect_tx(bar(SB_WFULL + s), bytes);
          tma_load_2d(sbase + ST_OFF_W + s * ST_WSLOT, map, bar(SB_WFULL + s), c0, c1);
        }
      }
    }
  } else if (warp == 2) {
    // =========================== o-tile producer ===========================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        // the buffer is free once the previous tile's last fc1 chunk has consumed xn2
        mbar_wait(bar(SB_OEMPTY), ((uint32_t)it & 1u) ^ 1u);
        mbar_expect_tx(bar(SB_OFULL), 3 * ST_TILE);
        for (int kb = 0; kb < 3; ++kb) tma_load_2d(sbase + ST_OFF_OX + kb * ST_TILE, &tmO, bar(SB_OFULL), kb * 64, tile * 128);
        // pull the residual rows of the tile after this one into L2 (they are fetched to smem ~1 tile from now)
        const int nt = tile + gridDim.x;
        if (nt < n_tiles)
          for (int c = 0; c < 6; ++c) tma_prefetch_2d(&tmResPf, c * 32, nt * 128);
      }
    }
  } else if (warp == 3) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      uint32_t wk = 0;  // weight ring position
      uint32_t n_hready[2] = {0, 0};
      const uint32_t tY = tmem_base, tP = tmem_base + 192, tX[2] = {tmem_base + 192, tmem_base + 320};
      auto w_wait = [&]() -> uint32_t {
        const int s = wk % ST_WSLOTS;
        mbar_wait(bar(SB_WFULL + s), (wk / ST_WSLOTS) & 1u);
        tc_fence_after();
        return sbase + ST_OFF_W + s * ST_WSLOT;
      };
      auto w_release = [&]() {
        umma_commit(bar(SB_WEMPTY + (wk % ST_WSLOTS)));
        ++wk;
      };
      // one k-block, both operands in smem: A tile at `a_addr`, W from the ring, 4 UMMAs of K=16
      auto kblock_ss = [&](uint32_t a_addr, uint32_t tmem_d, uint32_t idesc, bool first_clears) {
        const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(w_wait());
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<false>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (first_clears && k == 0) ? 0u : 1u);
        w_release();
      };
      auto fc1_chunk = [&](int c) {
        const int b = c & 1;
        for (int kb = 0; kb < 3; ++kb) kblock_ss(sbase + ST_OFF_OX + kb * ST_TILE, tX[b], IDESC_128, kb == 0);
        umma_commit(bar(SB_XFULL + b));
      };
      auto fc2_chunk = [&](int c) {
        const int b = c & 1;
        mbar_wait(bar(SB_HREADY + b), n_hready[b] & 1u);
        ++n_hready[b];
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) {
          // h of hidden units [kb*64, kb*64+64) of this chunk: packed bf16 pairs in X_b columns [kb*64, kb*64+32)
          const uint64_t bdesc = umma_desc_sw128(w_wait());
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ts(tY, tX[b] + (uint32_t)(kb * 64 + k * 8), bdesc + 2 * k, IDESC_192, 1u);
          w_release();
        }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t ph = (uint32_t)it & 1u;
        // projection into [192,384): everything that read X0/X1 of the previous tile was either waited for (GELU
        // epilogues, via HREADY) or is an earlier tcgen05.mma of this thread (fc2's TS reads; MMAs execute in order)
        mbar_wait(bar(SB_OFULL), ph);
        tc_fence_after();
        for (int kb = 0; kb < 3; ++kb) kblock_ss(sbase + ST_OFF_OX + kb * ST_TILE, tP, IDESC_192, kb == 0);
        umma_commit(bar(SB_PFULL));
        mbar_wait(bar(SB_XNREADY), ph);  // xn2 written over the o tile, Y = t' + b2
        tc_fence_after();
        fc1_chunk(0);
        fc1_chunk(1);
        fc2_chunk(0);
        fc1_chunk(2);
        umma_commit(bar(SB_OEMPTY));  // all reads of xn2 are done once this commit fires
        fc2_chunk(1);
        fc2_chunk(2);
        umma_commit(bar(SB_YFULL));
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue (8 warps) ===========================
    const int ew = warp - ST_EPI_WARP0;
    const int hf = ew >> 2;     // column half of every accumulator
    const int quad = warp & 3;  // TMEM lane quadrant
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint8_t* io = smem + ST_OFF_IO + ew * ST_IO_WARP;  // R0 | R1 | S
    const uint32_t io_s = smem_u32(io);
    const int rbar = SB_RFULL + ew * 2;
    uint32_t n_xfull[2] = {0, 0};
    const float invC = 1.0f / (float)a.C;
    const bool mask_tail = (hf == 1);  // this warp's last chunk holds the padded channels [C, 192)

    // residual chunk q (running index over this CTA's tiles, 3 per tile) lives in R[q % 2]
    uint32_t q = 0;
    auto res_issue = [&](uint32_t qq) {  // lane 0 only
      const int itl = (int)(qq / 3), c = (int)(qq % 3);
      const long long tile = (long long)blockIdx.x + (long long)itl * gridDim.x;
      if (tile < n_tiles) {
        const uint32_t b = bar(rbar + (qq & 1));
        mbar_expect_tx(b, 4096);
        tma_load_2d(io_s + (qq & 1) * 4096, &tmRes, b, hf * 96 + c * 32, (int)tile * 128 + quad * 32);
      }
    };
    if (lane == 0) {
      res_issue(0);
      res_issue(1);
    }

    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const int row0 = tile * 128 + quad * 32;  // first row of this warp's boxes

      long long* dbg = (a.dbg && ew == 0 && lane == 0 && it < 32) ? a.dbg + 16 * ((size_t)blockIdx.x * 32 + it) : nullptr;
      // ---------------- projection epilogue: t' = acc + bp + res ; Y <- t' + b2 ; xn2 -> smem ----------------
      float tv[3][32];
      if (dbg) dbg[0] = clock64();
      mbar_wait(bar(SB_PFULL), ph);
      if (dbg) dbg[1] = clock64();
      tc_fence_after();
      {
        uint32_t raw[3][32];
#pragma unroll
        for (int c = 0; c < 3; ++c) tmem_ld32_nowait(tlane + 192 + hf * 96 + c * 32, raw[c]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) tv[c][i] = __uint_as_float(raw[c][i]);
      }
      float sum = 0.0f, sq = 0.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int nb = hf * 96 + c * 32;
        mbar_wait(bar(rbar + (q & 1)), (q >> 1) & 1u);
        const uint8_t* rb = io + (q & 1) * 4096;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r = *reinterpret_cast<const float4*>(rb + sw128_off(lane, j));
          const float4 b = *reinterpret_cast<const float4*>(s_bp + nb + 4 * j);
          tv[c][4 * j + 0] += r.x + b.x;
          tv[c][4 * j + 1] += r.y + b.y;
          tv[c][4 * j + 2] += r.z + b.z;
          tv[c][4 * j + 3] += r.w + b.w;
        }
        __syncwarp();
        if (lane == 0) res_issue(q + 2);  // refill this buffer two chunks ahead
        ++q;
        if (c == 2 && mask_tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (nb + i >= a.C) tv[c][i] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          sum += tv[c][i];
          sq = fmaf(tv[c][i], tv[c][i], sq);
        }
      }
      if (dbg) dbg[2] = clock64();
      red[row * 2 + hf] = make_float2(sum, sq);
      // Y = t' + b2: fc2 accumulates on top of the residual (issued before the exchange so it overlaps the barrier)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int nb = hf * 96 + c * 32;
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = tv[c][i] + s_b2[nb + i];
        tmem_st32(tlane + nb, y);
      }
      named_bar_sync(1 + quad, 64);
      {
        const float2 r0 = red[row * 2], r1 = red[row * 2 + 1];
        const float mean = (r0.x + r1.x) * invC;
        const float var = fmaxf((r0.y + r1.y) * invC - mean * mean, 0.0f);
        const float rstd = rsqrtf(var + a.eps);
        // the projection MMAs completed before SB_PFULL fired, so the o tile is dead: overwrite it with xn2 in the
        // SWIZZLE_128B K-major layout TMA would have produced
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int nb = hf * 96 + c * 32;
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            float n[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int col = nb + 8 * qd + i;
              n[i] = fmaf((tv[c][8 * qd + i] - mean) * rstd, s_g2[col], s_be2[col]);  // pads: gamma = beta = 0
            }
            const int col = nb + 8 * qd;
            *reinterpret_cast<uint4*>(smem + ST_OFF_OX + (col >> 6) * ST_TILE + sw128_off(row, (col & 63) >> 3)) =
                make_uint4(pack_bf16x2(n[0], n[1]), pack_bf16x2(n[2], n[3]), pack_bf16x2(n[4], n[5]), pack_bf16x2(n[6], n[7]));
          }
        }
      }
      tc_fence_before();    // orders the tcgen05.st of Y (tmem_st32 waits for completion) before the arrive
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(SB_XNREADY));
      if (dbg) dbg[3] = clock64();

      // ---------------- fc1 chunk epilogues: h = GELU(acc + b1) -> packed bf16 over its own accumulator ----------------
#pragma unroll 1
      for (int ch = 0; ch < 3; ++ch) {
        const int b = ch & 1;
        mbar_wait(bar(SB_XFULL + b), n_xfull[b] & 1u);
        ++n_xfull[b];
        tc_fence_after();
        if (dbg) dbg[4 + 2 * ch] = clock64();
        const uint32_t tx = tlane + 192 + b * 128 + hf * 64;  // this warp's 64 hidden columns of the chunk
        uint32_t raw[2][32];
        tmem_ld32_nowait(tx, raw[0]);
        tmem_ld32_nowait(tx + 32, raw[1]);
        tmem_wait_ld();
        uint32_t pk[32];
        const float* bb = s_b1 + ch * 128 + hf * 64;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x0 = __uint_as_float(raw[i >> 4][(2 * i) & 31]) + bb[2 * i];
          const float x1 = __uint_as_float(raw[i >> 4][(2 * i + 1) & 31]) + bb[2 * i + 1];
          pk[i] = pack_bf16x2(gelu_tanh3(x0), gelu_tanh3(x1));
        }
        tmem_st32_u32(tx, pk);  // K index 2i, 2i+1 of this half -> column i (low half = even k)
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(SB_HREADY + b));
        if (dbg) dbg[5 + 2 * ch] = clock64();
      }

      // ---------------- final epilogue: t'' = Y ; TMA stores + LayerNorm_next ----------------
      mbar_wait(bar(SB_YFULL), ph);
      tc_fence_after();
      if (dbg) dbg[10] = clock64();
      {
        uint32_t raw[3][32];
#pragma unroll
        for (int c = 0; c < 3; ++c) tmem_ld32_nowait(tlane + hf * 96 + c * 32, raw[c]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) tv[c][i] = __uint_as_float(raw[c][i]);
      }
      if (mask_tail) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (hf * 96 + 64 + i >= a.C) tv[2][i] = 0.0f;
      }
      if (a.do_ln) {
        sum = 0.0f;
        sq = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            sum += tv[c][i];
            sq = fmaf(tv[c][i], tv[c][i], sq);
          }
        red[row * 2 + hf] = make_float2(sum, sq);
      }
      uint8_t* stg = io + 2 * 4096;
      const uint32_t stg_s = io_s + 2 * 4096;
      if (a.has_f32) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (lane == 0) bulk_wait_read<0>();  // the previous store has drained the staging box
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + sw128_off(lane, j)) =
                make_float4(tv[c][4 * j], tv[c][4 * j + 1], tv[c][4 * j + 2], tv[c][4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOutF, stg_s, hf * 96 + c * 32, row0);
            bulk_commit();
          }
        }
      }
      if (a.has_bf) {
        float mean = 0.0f, rstd = 1.0f;
        if (a.do_ln) {
          named_bar_sync(1 + quad, 64);
          const float2 r0 = red[row * 2], r1 = red[row * 2 + 1];
          mean = (r0.x + r1.x) * invC;
          const float var = fmaxf((r0.y + r1.y) * invC - mean * mean, 0.0f);
          rstd = rsqrtf(var + a.eps);
          named_bar_sync(1 + quad, 64);  // the partner has read `red` before the next tile's projection epilogue rewrites it
        }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int nb = hf * 96 + c * 32;
          if (c == 2) {  // the first half-box is re-used
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
          uint8_t* sb = stg + (c & 1) * 2048;
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            float n[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int col = nb + 8 * qd + i;
              const float v = tv[c][8 * qd + i];
              n[i] = a.do_ln ? fmaf((v - mean) * rstd, s_g3[col], s_be3[col]) : v;
            }
            *reinterpret_cast<uint4*>(sb + sw64_off(lane, qd)) =
                make_uint4(pack_bf16x2(n[0], n[1]), pack_bf16x2(n[2], n[3]), pack_bf16x2(n[4], n[5]), pack_bf16x2(n[6], n[7]));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOutB, stg_s + (c & 1) * 2048, nb, row0);
            bulk_commit();
          }
        }
      }
      tc_fence_before();  // Y and the projection columns are re-written by this warp in the next tile (program order)
      if (dbg) dbg[11] = clock64();
    }
    if (lane == 0) bulk_wait_all();  // every store of this warp has landed before the CTA may exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

long long* g_tail_dbg = nullptr;

int launch_mlp_fused(const MlpFusedArgs& f, cudaStream_t s) {
  SSR_CHECK(f.CP == 192 && f.HP == 384 && f.QP == 192, SSR_E_INVALID, "swin_tail: unsupported padded dims %d/%d/%d", f.CP, f.HP,
            f.QP);
  SSR_CHECK(!(f.out_T && f.out_ln), SSR_E_INVALID, "swin_tail: at most one bf16 output");
  SSR_CHECK(f.ldres % 4 == 0 && (!f.out_f32 || f.ld_f32 % 4 == 0), SSR_E_INVALID, "swin_tail: fp32 leading dims must be 16-byte multiples");
  CUtensorMap tmO, tmWp, tmW1, tmW2, tmRes, tmResPf, tmOutF, tmOutB;
  auto map2d = [&](CUtensorMap* m, const void* base, int elem, int cols, int rows, int ld, int box_c, int box_r, int sw) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)ld * elem};
    cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
    return make_tmap(m, base, elem, 2, dims, str, box, sw);
  };
  SSR_TRY(map2d(&tmO, f.o, 2, f.ld_o, f.M, f.ld_o, 64, 128, 128));
  SSR_TRY(map2d(&tmWp, f.Wp, 2, 192, 192, 192, 64, 192, 128));
  SSR_TRY(map2d(&tmW1, f.W1, 2, 192, 384, 192, 64, 128, 128));
  SSR_TRY(map2d(&tmW2, f.W2, 2, 384, 192, 384, 64, 192, 128));
  SSR_TRY(map2d(&tmRes, f.res, 4, 192, f.M, f.ldres, 32, 32, 128));
  SSR_TRY(map2d(&tmResPf, f.res, 4, 192, f.M, f.ldres, 32, 128, 128));
  const void* outf = f.out_f32 ? (const void*)f.out_f32 : (const void*)f.res;  // placeholder map when unused
  SSR_TRY(map2d(&tmOutF, outf, 4, 192, f.M, f.out_f32 ? f.ld_f32 : f.ldres, 32, 32, 128));
  const void* outb = f.out_ln ? f.out_ln : f.out_T;
  const int ldb = f.out_ln ? f.ld_ln : f.ld_T;
  if (outb) {
    SSR_CHECK(ldb % 8 == 0, SSR_E_INVALID, "swin_tail: bf16 leading dim must be a 16-byte multiple");
    SSR_TRY(map2d(&tmOutB, outb, 2, 192, f.M, ldb, 32, 32, 64));
  } else {
    tmOutB = tmO;
  }
  TailArgs a;
  a.M = f.M; a.C = f.C; a.n_tiles = (f.M + 127) / 128;
  a.bp = f.bp; a.b1 = f.b1; a.b2 = f.b2; a.g2 = f.g2; a.be2 = f.be2; a.g3 = f.g3; a.be3 = f.be3;
  a.has_f32 = f.out_f32 != nullptr;
  a.has_bf = outb != nullptr;
  a.do_ln = f.out_ln != nullptr;
  a.eps = f.eps;
  a.dbg = g_tail_dbg;
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CUDA(cudaFuncSetAttribute(swin_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
    attr_set = true;
  }
  const int sms = num_sms_cached();
  const double flops = 2.0 * f.M * ((double)f.C * f.C + 2.0 * f.C * f.Hid);
  const double bytes = (double)f.M * f.C * (2 + 4 + (f.out_f32 ? 4 : 0) + (f.out_T ? 2 : 0) + (f.out_ln ? 2 : 0));
  ProfScope prof("swin_tail", flops, bytes, s);
  swin_tail_kernel<<<a.n_tiles < sms ? a.n_tiles : sms, ST_THREADS, ST_SMEM, s>>>(tmO, tmWp, tmW1, tmW2, tmRes, tmResPf, tmOutF,
                                                                                 tmOutB, a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
