// Fused tail of a Swin block (bf16, C=180 -> padded 192, hidden 360 -> 384), one persistent kernel:
//
//   t'  = o @ Wproj^T + bproj + res                (swinir.py:103,171)    [tcgen05, acc in TMEM]
//   xn2 = LayerNorm2(t')                           (swinir.py:172)        [epilogue -> smem A operand]
//   h   = GELU(xn2 @ W1^T + b1)                    (common.py:185-186)    [3 chunks of 128 hidden units]
//   t'' = t' + h @ W2^T + b2                       (common.py:188, swinir.py:172)
//   out: t'' (fp32 residual stream), LayerNorm_next(t'') or a bf16 copy of t''
//
// Only `o` (384 B), `res` (768 B) are read and t'' (768 B), xn (384 B) written per token: 2.3 kB instead of the
// 6.1 kB of three separate GEMM launches; xn2 and h never leave the SM.  The residual add is free: the epilogue
// of the projection writes (t' + b2) into the TMEM columns that then serve as the fc2 accumulator.
//
// TMEM (512 columns): Y = [0,192) fc2 accumulator; [192,384) projection accumulator, re-used as the two
// fc1 chunk accumulators X0 = [192,320), X1 = [320,448).
// Warps: 0 = TMA producer (o tile + weight k-blocks, L2-resident), 1 = MMA issuer, 2..9 = epilogue; epilogue
// warp w owns TMEM lane quadrant (w % 4) and column half (w - 2) / 4 of every accumulator.
#include "ssr_tc.cuh"

namespace ssr {

constexpr int MF_THREADS = 320;
constexpr int MF_C = 192, MF_H = 384, MF_NC = 128, MF_CHUNKS = 3;
constexpr int MF_WSLOTS = 4;
constexpr uint32_t MF_TILE = 16384;          // one [128 rows][128 B] k-block tile
constexpr uint32_t MF_WSLOT = 192 * 128;     // weight ring slot (fc1 uses 128 rows of it)
constexpr uint32_t MF_OFF_OX = 0;                              // 3 tiles: o, later xn2
constexpr uint32_t MF_OFF_H = MF_OFF_OX + 3 * MF_TILE;         // 2 buffers x 2 tiles (aliased as epilogue staging)
constexpr uint32_t MF_OFF_W = MF_OFF_H + 4 * MF_TILE;          // weight ring
constexpr uint32_t MF_OFF_PAR = MF_OFF_W + MF_WSLOTS * MF_WSLOT;  // fp32 parameters
constexpr int MF_NPAR = 192 * 6 + 384;                         // bp b2 g2 be2 g3 be3 | b1
constexpr uint32_t MF_OFF_RED = MF_OFF_PAR + MF_NPAR * 4;      // [2][128][2] floats cross-half reductions
constexpr uint32_t MF_OFF_BAR = MF_OFF_RED + 2 * 128 * 2 * 4;
constexpr uint32_t MF_SMEM = MF_OFF_BAR + 256 + 1024;
static_assert(MF_SMEM <= 232448, "fused MLP kernel exceeds the 227 KB shared-memory limit");

enum {  // mbarrier indices
  MB_WFULL = 0,                      // [MF_WSLOTS]
  MB_WEMPTY = MB_WFULL + MF_WSLOTS,  // [MF_WSLOTS]
  MB_OFULL = MB_WEMPTY + MF_WSLOTS,
  MB_OEMPTY,
  MB_PFULL,    // projection accumulator complete
  MB_XNREADY,  // xn2 in smem + Y initialised
  MB_XFULL,    // [2] fc1 chunk accumulator complete
  MB_XEMPTY = MB_XFULL + 2,  // [2] drained by the epilogue
  MB_HREADY = MB_XEMPTY + 2, // [2] h chunk in smem
  MB_HEMPTY = MB_HREADY + 2, // [2] consumed by fc2 MMAs
  MB_YFULL = MB_HEMPTY + 2,
  MB_COUNT
};

struct MlpArgs {
  const float* res;  // fp32 [M][ldres]
  int ldres;
  int M, C;
  const float *bp, *b1, *b2, *g2, *be2, *g3, *be3;
  float* out_f32;  // t'' or null
  int ld_f32;
  __nv_bfloat16* out_T;  // bf16 copy of t'' or null
  int ld_T;
  __nv_bfloat16* out_ln;  // LayerNorm_next(t'') or null
  int ld_ln;
  float eps;
};

// byte offset of (row r, 16-byte chunk j) inside a [128][128 B] SWIZZLE_128B K-major tile
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(MF_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWp,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const MlpArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* par = reinterpret_cast<float*>(smem + MF_OFF_PAR);
  float *s_bp = par, *s_b2 = par + 192, *s_g2 = par + 384, *s_be2 = par + 576, *s_g3 = par + 768, *s_be3 = par + 960,
        *s_b1 = par + 1152;
  float* red = reinterpret_cast<float*>(smem + MF_OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MF_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + MB_COUNT);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 192; i += MF_THREADS) {
    s_bp[i] = __ldg(a.bp + i);
    s_b2[i] = __ldg(a.b2 + i);
    s_g2[i] = __ldg(a.g2 + i);
    s_be2[i] = __ldg(a.be2 + i);
    s_g3[i] = a.g3 ? __ldg(a.g3 + i) : 0.f;
    s_be3[i] = a.be3 ? __ldg(a.be3 + i) : 0.f;
  }
  for (int i = threadIdx.x; i < 384; i += MF_THREADS) s_b1[i] = __ldg(a.b1 + i);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmO);
    prefetch_tmap(&tmWp);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    for (int i = 0; i < MB_COUNT; ++i) {
      const bool epi_arrives = (i == MB_XNREADY) || (i >= MB_XEMPTY && i < MB_HEMPTY);
      mbar_init(bar(i), epi_arrives ? 8 : 1);  // epilogue-signalled barriers: one arrival per epilogue warp
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = (a.M + 127) / 128;
  constexpr uint32_t IDESC_192 = umma_idesc(1, 128, 192), IDESC_128 = umma_idesc(1, 128, 128);

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t wk = 0;  // weight ring position
      auto load_w = [&](const CUtensorMap* map, uint32_t bytes, int c0, int c1) {
        const int s = wk % MF_WSLOTS;
        mbar_wait(bar(MB_WEMPTY + s), ((wk / MF_WSLOTS) & 1u) ^ 1u);
        mbar_expect_tx(bar(MB_WFULL + s), bytes);
        tma_load_2d(sbase + MF_OFF_W + s * MF_WSLOT, map, bar(MB_WFULL + s), c0, c1);
        ++wk;
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        // o tile (the buffer is free once the previous tile's last fc1 chunk has consumed xn2)
        mbar_wait(bar(MB_OEMPTY), ((uint32_t)it & 1u) ^ 1u);
        mbar_expect_tx(bar(MB_OFULL), 3 * MF_TILE);
        for (int kb = 0; kb < 3; ++kb) tma_load_2d(sbase + MF_OFF_OX + kb * MF_TILE, &tmO, bar(MB_OFULL), kb * 64, tile * 128);
        // weights, in exactly the order the MMA warp consumes them
        for (int kb = 0; kb < 3; ++kb) load_w(&tmWp, 192 * 128, kb * 64, 0);            // proj
        for (int kb = 0; kb < 3; ++kb) load_w(&tmW1, 128 * 128, kb * 64, 0);            // fc1 chunk 0
        for (int kb = 0; kb < 3; ++kb) load_w(&tmW1, 128 * 128, kb * 64, 128);          // fc1 chunk 1
        for (int kb = 0; kb < 2; ++kb) load_w(&tmW2, 192 * 128, 0 * 128 + kb * 64, 0);  // fc2 chunk 0
        for (int kb = 0; kb < 3; ++kb) load_w(&tmW1, 128 * 128, kb * 64, 256);          // fc1 chunk 2
        for (int kb = 0; kb < 2; ++kb) load_w(&tmW2, 192 * 128, 1 * 128 + kb * 64, 0);  // fc2 chunk 1
        for (int kb = 0; kb < 2; ++kb) load_w(&tmW2, 192 * 128, 2 * 128 + kb * 64, 0);  // fc2 chunk 2
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      uint32_t wk = 0;
      uint32_t n_xempty[2] = {0, 0}, n_hready[2] = {0, 0};  // completed-phase counters of the barriers we wait on
      const uint32_t tY = tmem_base, tP = tmem_base + 192, tX[2] = {tmem_base + 192, tmem_base + 320};
      // one k-block: A tile at `a_addr`, W from the ring, 4 UMMAs of K=16
      auto kblock = [&](uint32_t a_addr, uint32_t tmem_d, uint32_t idesc, bool first_clears) {
        const int s = wk % MF_WSLOTS;
        mbar_wait(bar(MB_WFULL + s), (wk / MF_WSLOTS) & 1u);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(sbase + MF_OFF_W + s * MF_WSLOT);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<false>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (first_clears && k == 0) ? 0u : 1u);
        umma_commit(bar(MB_WEMPTY + s));
        ++wk;
      };
      auto wait_xempty = [&](int b) {  // wait until the epilogue has drained X[b] as many times as we filled it
        if (n_xempty[b] > 0) mbar_wait(bar(MB_XEMPTY + b), (n_xempty[b] - 1) & 1u);
      };
      auto fc1_chunk = [&](int c) {
        const int b = c & 1;
        wait_xempty(b);
        tc_fence_after();
        for (int kb = 0; kb < 3; ++kb) kblock(sbase + MF_OFF_OX + kb * MF_TILE, tX[b], IDESC_128, kb == 0);
        umma_commit(bar(MB_XFULL + b));
        ++n_xempty[b];
      };
      auto fc2_chunk = [&](int c) {
        const int b = c & 1;
        mbar_wait(bar(MB_HREADY + b), n_hready[b] & 1u);
        ++n_hready[b];
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) kblock(sbase + MF_OFF_H + (b * 2 + kb) * MF_TILE, tY, IDESC_192, false);
        umma_commit(bar(MB_HEMPTY + b));
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t ph = (uint32_t)it & 1u;
        // projection into [192,384): both fc1 accumulators of the previous tile must have been drained
        wait_xempty(0);
        wait_xempty(1);
        mbar_wait(bar(MB_OFULL), ph);
        tc_fence_after();
        for (int kb = 0; kb < 3; ++kb) kblock(sbase + MF_OFF_OX + kb * MF_TILE, tP, IDESC_192, kb == 0);
        umma_commit(bar(MB_PFULL));
        mbar_wait(bar(MB_XNREADY), ph);  // xn2 written over the o tile, Y = t' + b2
        tc_fence_after();
        fc1_chunk(0);
        fc1_chunk(1);
        fc2_chunk(0);
        fc1_chunk(2);
        umma_commit(bar(MB_OEMPTY));  // all reads of xn2 are done once this commit fires
        fc2_chunk(1);
        fc2_chunk(2);
        umma_commit(bar(MB_YFULL));
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue (8 warps) ===========================
    const int ew = warp - 2;
    const int hf = ew >> 2;     // column half of every accumulator
    const int quad = warp & 3;  // TMEM lane quadrant
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    float* st = reinterpret_cast<float*>(smem + MF_OFF_H) + ew * (TC_STAGE_BYTES / 4);  // staging aliases the h buffers
    uint32_t n_xfull[2] = {0, 0}, n_hempty[2] = {0, 0};
    const float invC = 1.0f / (float)a.C;

    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const int mm = tile * 128 + row;
      const int m = mm < a.M ? mm : -1;
      const RowMap rm = make_rowmap(m, lane);

      // ---------------- projection epilogue: t' = acc + bp + res ; Y <- t' + b2 ; xn2 -> smem ----------------
      float4 resv[8];
      res_prefetch(resv, rm, lane, a.res, a.ldres, hf * 96);
      mbar_wait(bar(MB_PFULL), ph);
      tc_fence_after();
      float sum = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const int nb = hf * 96 + c * 32;
        float v[32];
        tmem_ld32(tlane + 192 + nb, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += s_bp[nb + i];
        stage_put_f32(st, lane, v);
        __syncwarp();
        stage_add_store_f32(st, lane, rm, resv, nullptr, 0, nb);
        if (c < 2) res_prefetch(resv, rm, lane, a.res, a.ldres, nb + 32);
        __syncwarp();
        stage_get_f32(st, lane, v);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (nb + i >= a.C) v[i] = 0.0f;
          sum += v[i];
        }
        tmem_st32(tlane + 192 + nb, v);  // t' stays in the projection columns for the LayerNorm passes
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += s_b2[nb + i];
        tmem_st32(tlane + nb, v);  // Y = t' + b2: fc2 accumulates on top of the residual
      }
      red[row * 2 + hf] = sum;
      named_bar_sync(1, 256);
      const float mean = (red[row * 2] + red[row * 2 + 1]) * invC;
      float sq = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const int nb = hf * 96 + c * 32;
        float v[32];
        tmem_ld32(tlane + 192 + nb, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = v[i] - mean;
          if (nb + i < a.C) sq += d * d;
        }
      }
      red[256 + row * 2 + hf] = sq;
      named_bar_sync(1, 256);
      const float rstd = rsqrtf((red[256 + row * 2] + red[256 + row * 2 + 1]) * invC + a.eps);
      // all threads have read the o tile?  The projection MMAs completed before MB_PFULL fired, so the o tile is
      // dead and can be overwritten with xn2 (same SWIZZLE_128B K-major layout TMA would have produced).
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const int nb = hf * 96 + c * 32;
        float v[32];
        tmem_ld32(tlane + 192 + nb, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * s_g2[nb + i] + s_be2[nb + i];  // pads: gamma = beta = 0
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = nb + 8 * q;
          const uint4 pk = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                      pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
          *reinterpret_cast<uint4*>(smem + MF_OFF_OX + (col >> 6) * MF_TILE + sw128_off(row, (col & 63) >> 3)) = pk;
        }
      }
      tc_fence_before();
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(MB_XNREADY));

      // ---------------- fc1 chunk epilogues: h = GELU(acc + b1) -> smem ----------------
#pragma unroll 1
      for (int ch = 0; ch < MF_CHUNKS; ++ch) {
        const int b = ch & 1;
        mbar_wait(bar(MB_XFULL + b), n_xfull[b] & 1u);
        ++n_xfull[b];
        tc_fence_after();
        if (n_hempty[b] > 0) mbar_wait(bar(MB_HEMPTY + b), (n_hempty[b] - 1) & 1u);  // fc2 finished reading this h buffer
        ++n_hempty[b];
        if (ch == 0) named_bar_sync(1, 256);  // the staging tiles alias the h buffers: everyone is past the staging phase
        uint8_t* htile = smem + MF_OFF_H + (b * 2 + hf) * MF_TILE;  // this half's 64 hidden columns = one k-block tile
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int nl = hf * 64 + c * 32;  // column inside the chunk
          float v[32];
          tmem_ld32(tlane + 192 + b * 128 + nl, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = v[i] + s_b1[ch * MF_NC + nl + i];
            v[i] = 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752440f));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 pk = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                        pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
            *reinterpret_cast<uint4*>(htile + sw128_off(row, c * 4 + q)) = pk;
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(MB_XEMPTY + b));
          mbar_arrive(bar(MB_HREADY + b));
        }
      }

      // ---------------- final epilogue: t'' = Y ; stores + LayerNorm_next ----------------
      mbar_wait(bar(MB_YFULL), ph);
      tc_fence_after();
      named_bar_sync(1, 256);  // h buffers are dead for every warp -> safe to use them as staging again
      sum = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const int nb = hf * 96 + c * 32;
        float v[32];
        tmem_ld32(tlane + nb, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (nb + i >= a.C) v[i] = 0.0f;
          sum += v[i];
        }
        if (a.out_f32) {
          stage_put_f32(st, lane, v);
          __syncwarp();
          stage_add_store_f32(st, lane, rm, nullptr, a.out_f32, a.ld_f32, nb);
          __syncwarp();
        }
        if (a.out_T)
          stage_store_bf16(st, lane, v, [&](int i) { return rm.mh[i] >= 0 ? a.out_T + (size_t)rm.mh[i] * a.ld_T + nb : nullptr; });
      }
      if (a.out_ln) {
        red[row * 2 + hf] = sum;
        named_bar_sync(1, 256);
        const float mean3 = (red[row * 2] + red[row * 2 + 1]) * invC;
        sq = 0.0f;
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          const int nb = hf * 96 + c * 32;
          float v[32];
          tmem_ld32(tlane + nb, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float d = v[i] - mean3;
            if (nb + i < a.C) sq += d * d;
          }
        }
        red[256 + row * 2 + hf] = sq;
        named_bar_sync(1, 256);
        const float rstd3 = rsqrtf((red[256 + row * 2] + red[256 + row * 2 + 1]) * invC + a.eps);
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          const int nb = hf * 96 + c * 32;
          float v[32];
          tmem_ld32(tlane + nb, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean3) * rstd3 * s_g3[nb + i] + s_be3[nb + i];
          stage_store_bf16(st, lane, v, [&](int i) { return rm.mh[i] >= 0 ? a.out_ln + (size_t)rm.mh[i] * a.ld_ln + nb : nullptr; });
        }
      }
      tc_fence_before();  // Y and the projection columns are re-written by this warp in the next tile (program order)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_mlp_fused(const MlpFusedArgs& f, cudaStream_t s) {
  SSR_CHECK(f.CP == MF_C && f.HP == MF_H && f.QP == MF_C, SSR_E_INVALID, "mlp_fused: unsupported padded dims %d/%d/%d",
            f.CP, f.HP, f.QP);
  CUtensorMap tmO, tmWp, tmW1, tmW2;
  {
    cuuint64_t dims[2] = {(cuuint64_t)f.ld_o, (cuuint64_t)f.M};
    cuuint64_t str[1] = {(cuuint64_t)f.ld_o * 2};
    cuuint32_t box[2] = {64, 128};
    SSR_TRY(make_tmap(&tmO, f.o, 2, 2, dims, str, box));
  }
  {
    cuuint64_t dims[2] = {192, 192};
    cuuint64_t str[1] = {192 * 2};
    cuuint32_t box[2] = {64, 192};
    SSR_TRY(make_tmap(&tmWp, f.Wp, 2, 2, dims, str, box));
  }
  {
    cuuint64_t dims[2] = {192, 384};
    cuuint64_t str[1] = {192 * 2};
    cuuint32_t box[2] = {64, 128};
    SSR_TRY(make_tmap(&tmW1, f.W1, 2, 2, dims, str, box));
  }
  {
    cuuint64_t dims[2] = {384, 192};
    cuuint64_t str[1] = {384 * 2};
    cuuint32_t box[2] = {64, 192};
    SSR_TRY(make_tmap(&tmW2, f.W2, 2, 2, dims, str, box));
  }
  MlpArgs a;
  a.res = f.res; a.ldres = f.ldres; a.M = f.M; a.C = f.C;
  a.bp = f.bp; a.b1 = f.b1; a.b2 = f.b2; a.g2 = f.g2; a.be2 = f.be2; a.g3 = f.g3; a.be3 = f.be3;
  a.out_f32 = f.out_f32; a.ld_f32 = f.ld_f32;
  a.out_T = reinterpret_cast<__nv_bfloat16*>(f.out_T); a.ld_T = f.ld_T;
  a.out_ln = reinterpret_cast<__nv_bfloat16*>(f.out_ln); a.ld_ln = f.ld_ln;
  a.eps = f.eps;
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CUDA(cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MF_SMEM));
    attr_set = true;
  }
  const int n_tiles = (f.M + 127) / 128;
  const int sms = num_sms_cached();
  const double flops = 2.0 * f.M * ((double)f.C * f.C + 2.0 * f.C * f.Hid);
  const double bytes = (double)f.M * f.C * (2 + 4 + (f.out_f32 ? 4 : 0) + (f.out_T ? 2 : 0) + (f.out_ln ? 2 : 0));
  ProfScope prof("mlp_fused", flops, bytes, s);
  mlp_fused_kernel<<<n_tiles < sms ? n_tiles : sms, MF_THREADS, MF_SMEM, s>>>(tmO, tmWp, tmW1, tmW2, a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
