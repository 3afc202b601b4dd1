// Fused tail of a Swin block (bf16, C=180 -> padded 192, hidden 360 -> 384), one persistent kernel:
//
//   t'  = o @ Wproj^T + bproj + res                (swinir.py:103,171)    [tcgen05 SS, acc in TMEM]
//   xn2 = (t' - mean) * rstd                       (swinir.py:172; norm2's gamma/beta are folded into W1/b1 at pack time)
//   h   = GELU(xn2 @ W1^T + b1)                    (common.py:185-186)    [6 chunks of 64 hidden units; b1 rides in the GEMM:
//                                                   xn2's pad channels C, C+1 are 1.0 and W1's columns there hold b1 as hi + lo]
//   t'' = t' + h @ W2^T + b2                       (common.py:188, swinir.py:172)
//   out: t'' (fp32 residual stream), LayerNorm_next(t'') or a bf16 copy of t''
//
// Round 2: TWO TILES IN FLIGHT.  Round 1 ran a tile's phases back to back on one set of 8 epilogue warps (21k cycles per
// 128-token tile: projection epilogue 4.3k, MLP 9.3k + 1.8k, final epilogue 7k -- the tensor pipe idle during 11k of them).
// Now the residual-stream accumulator Y is double-buffered in TMEM and the epilogue is split into two roles that work on
// different tiles at the same time:
//   IO warps   (8): E1(j) = projection epilogue of tile j (residual add, LayerNorm statistics, xn2 -> smem), then
//                   E3(j-1) = final epilogue of tile j-1 (fp32 stream + LayerNorm_next / bf16 copy -> TMA stores);
//   GELU warps (8): the six fc1-chunk epilogues of tile j;
//   MMA issuer    : fc1 / fc2 chunks of tile j with the projection of tile j+1 slotted in after the first two fc1 chunks.
// So the memory-bound phases of one tile run under the tensor-bound MLP of the other.
//
// Design notes (all measured on B200, profiles/r01_micro_tc.txt):
//  * h never touches shared memory: the GELU epilogue packs it to bf16 and writes it back over its own fc1
//    accumulator columns with tcgen05.st; fc2 consumes it as the TMEM A operand of a TS-mode tcgen05.mma.
//  * Every bulk transfer is TMA.  The fp32 residual arrives in per-warp [32 x 32] SWIZZLE_128B boxes (thread <-> row
//    reads are bank-conflict free); the outputs leave as per-warp TMA stores from the same boxes.  No thread ever waits on
//    a global load, and the ragged last tile is clipped by TMA.
//  * The o tile streams through the weight ring (it is only needed by the three projection k-blocks), so the only
//    tile-sized operand buffer is xn2; two producer threads fill the ring (a single thread sustains one wait -> issue round
//    per ~500 cycles).
//  * LayerNorm statistics are one-pass (sum, sum of squares in fp32) on values held in registers.
//
// TMEM (512 columns): Y0 = [0,192), Y1 = [192,384): projection accumulator -> t' + b2 -> fc2 accumulator of tiles with even /
// odd local index; X0 = [384,448), X1 = [448,512): fc1 chunk accumulators, overwritten in place by the packed bf16 h.
// Warps (20, registers re-balanced with setmaxnreg): 0,1,2 = ring producers (+ L2 prefetch of the residual rows), 3 = MMA
// issuer, 4..11 = IO warps, 12..19 = GELU warps; warp w owns TMEM lane quadrant w % 4 and column half ((w - 4) % 8) / 4.
#include "ssr_tc.cuh"

namespace ssr {

constexpr int ST_THREADS = 640;
constexpr int ST_IO_WARP0 = 4, ST_GELU_WARP0 = 12;
// register budgets after setmaxnreg (the CTA's pool is 640 x 96 = 61440 registers)
#ifndef ST_REGS_WG0
#define ST_REGS_WG0 40
#define ST_REGS_IO 152
#define ST_REGS_GELU 64
#endif
static_assert(128 * ST_REGS_WG0 + 256 * ST_REGS_IO + 256 * ST_REGS_GELU <= 640 * 96, "setmaxnreg budgets exceed the CTA's register pool");
constexpr int ST_WSLOTS = 4;
constexpr int ST_NCHUNK = 6;                  // fc1 / fc2 chunks of 64 hidden units
constexpr uint32_t ST_TILE = 16384;           // one [128 rows][128 B] k-block tile
constexpr uint32_t ST_WSLOT = 192 * 128;      // ring slot: Wproj / W2 k-block [192 x 64], W1 chunk 3 x [64 x 64], o k-block [128 x 64]
constexpr uint32_t ST_OFF_XN = 0;                                  // 3 k-block tiles of xn2 (A operand of fc1)
constexpr uint32_t ST_OFF_W = ST_OFF_XN + 3 * ST_TILE;             // ring
constexpr uint32_t ST_OFF_IO = ST_OFF_W + ST_WSLOTS * ST_WSLOT;    // per IO warp: B0, B1 (residual in, outputs out)
constexpr uint32_t ST_IO_WARP = 2 * 4096;
constexpr uint32_t ST_OFF_PAR = ST_OFF_IO + 8 * ST_IO_WARP;        // fp32 parameters
constexpr int ST_NPAR = 192 * 4;                                   // bp b2 g3 be3
constexpr uint32_t ST_OFF_RED = ST_OFF_PAR + ST_NPAR * 4;          // [128][2] float2 cross-half reductions
constexpr uint32_t ST_OFF_BAR = ST_OFF_RED + 128 * 2 * 8;
constexpr uint32_t ST_SMEM = ST_OFF_BAR + 512 + 1024;
static_assert(ST_SMEM <= 232448, "fused tail kernel exceeds the 227 KB shared-memory limit");

enum {  // mbarrier indices
  SB_WFULL = 0,                       // [ST_WSLOTS]
  SB_WEMPTY = SB_WFULL + ST_WSLOTS,   // [ST_WSLOTS]
  SB_PFULL = SB_WEMPTY + ST_WSLOTS,   // [2] projection accumulator in Y0 / Y1 complete
  SB_XNREADY = SB_PFULL + 2,          // xn2 in smem + Y = t' + b2 (8 arrivals: IO warps)
  SB_XNFREE,                          // every fc1 MMA of the tile has read xn2
  SB_XFULL,                           // [2] fc1 chunk accumulator X0 / X1 complete
  SB_HREADY = SB_XFULL + 2,           // [2] h chunk packed into TMEM (8 arrivals: GELU warps)
  SB_YFULL = SB_HREADY + 2,           // [2] fc2 of the tile complete
  SB_YFREE = SB_YFULL + 2,            // [2] the final epilogue holds Y in registers (8 arrivals: IO warps)
  SB_RFULL = SB_YFREE + 2,            // [8 warps][2] residual chunk landed in B0 / B1
  SB_COUNT = SB_RFULL + 16
};
static_assert(SB_COUNT * 8 + 8 <= 512, "barrier area");

struct TailArgs {
  int M, C, n_tiles;
  const float *bp, *b2, *g3, *be3;  // fc1's bias (with norm2's beta folded in) rides in W1's columns C, C+1
  int has_f32;  // store t'' as fp32
  int has_bf;   // store a bf16 tensor: LayerNorm_next(t'') if do_ln else t''
  int do_ln;
  float eps;
  long long* dbg;  // optional phase timestamps (developer diagnostics): 3 regions of [CTA][tile < 32][16]
};

// byte offset of (row r, 16-byte chunk j) inside a [rows][128 B] SWIZZLE_128B box / [rows][64 B] SWIZZLE_64B box
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t sw64_off(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// GELU(x) = x * Phi(x), Phi(x) ~ 0.5 (1 + tanh(x (c0 + c1 x^2))) with (c0, c1) least-squares fitted to the erf form
// on [-8, 8]: max |err| 3.1e-4, far inside the bf16 rounding of h.  Two elements per call on the packed fp32x2 pipe:
// 6 FFMA2-class instructions + 2 MUFU.TANH + 1 pack for a pair.
__device__ __forceinline__ uint32_t gelu2_bf16(f32x2 x, f32x2 kC0, f32x2 kC1, f32x2 kHalf) {
  const f32x2 x2 = f2_mul(x, x);
  const f32x2 u = f2_mul(x, f2_fma(x2, kC1, kC0));
  float ulo, uhi, tlo, thi;
  f2_unpack(u, ulo, uhi);
  asm("tanh.approx.f32 %0, %1;" : "=f"(tlo) : "f"(ulo));
  asm("tanh.approx.f32 %0, %1;" : "=f"(thi) : "f"(uhi));
  const f32x2 hx = f2_mul(x, kHalf);
  return f2_to_bf16x2(f2_fma(hx, f2_pack(tlo, thi), hx));
}

// The ring entries of one CTA in the order the MMA issuer consumes them.  `f(kind, tile_it, idx)`: kind 0 = o k-block idx of
// local tile tile_it, 1 = Wproj k-block idx, 2 = W1 chunk idx, 3 = W2 chunk idx.
template <typename F>
__device__ __forceinline__ void for_each_ring_entry(int my_tiles, F f) {
  if (my_tiles <= 0) return;
  for (int kb = 0; kb < 3; ++kb) {
    f(0, 0, kb);
    f(1, 0, kb);
  }
  for (int it = 0; it < my_tiles; ++it) {
    f(2, it, 0);
    f(2, it, 1);
    if (it + 1 < my_tiles)
      for (int kb = 0; kb < 3; ++kb) {
        f(0, it + 1, kb);
        f(1, it + 1, kb);
      }
    for (int c = 0; c < ST_NCHUNK; ++c) {
      f(3, it, c);
      if (c + 2 < ST_NCHUNK) f(2, it, c + 2);
    }
  }
}

// CL = CTAs per cluster.  With CL = 2 the two CTAs walk the same ring-entry sequence; each fetches 1/CL of every weight
// entry and TMA-multicasts it into the slot of both, so Wproj / W1 / W2 (366 KB per 128-token tile, 56 % of the kernel's L2
// reads -- the kernel runs at the chip's L2 throughput cap) leave L2 once per cluster.  A slot is reusable once the MMA
// issuers of BOTH CTAs have released it: WEMPTY counts CL multicast commits.
template <int CL>
__global__ void __launch_bounds__(ST_THREADS, 1)
swin_tail_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWp,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmResPf,
                 const __grid_constant__ CUtensorMap tmOutF, const __grid_constant__ CUtensorMap tmOutB,
                 const __grid_constant__ CUtensorMap tmOutB2, const TailArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // pointer arithmetic (not an integer round trip) keeps the shared address space visible to the compiler: LDS/STS
  // instead of generic LD/ST for every staging access
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* par = reinterpret_cast<float*>(smem + ST_OFF_PAR);
  float *s_bp = par, *s_b2 = par + 192, *s_g3 = par + 384, *s_be3 = par + 576;
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
    s_g3[i] = a.g3 ? __ldg(a.g3 + i) : 0.f;
    s_be3[i] = a.be3 ? __ldg(a.be3 + i) : 0.f;
  }
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmO);
    prefetch_tmap(&tmWp);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    prefetch_tmap(&tmRes);
    prefetch_tmap(&tmResPf);
    prefetch_tmap(&tmOutF);
    prefetch_tmap(&tmOutB);
    prefetch_tmap(&tmOutB2);
    for (int i = 0; i < SB_COUNT; ++i) {
      const bool cnt8 = i == SB_XNREADY || (i >= SB_HREADY && i < SB_HREADY + 2) || (i >= SB_YFREE && i < SB_YFREE + 2);
      const bool wempty = i >= SB_WEMPTY && i < SB_WEMPTY + ST_WSLOTS;
      mbar_init(bar(i), cnt8 ? 8 : wempty ? CL : 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 3) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = a.n_tiles;
  // tiles of this CTA.  In a cluster every CTA takes the same number (the ring sequences must match): tile indices >= n_tiles
  // are dummies whose rows lie beyond M -- TMA zero-fills their loads and drops their stores.
  const int my_tiles = CL > 1 ? (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x
                              : (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kClMask = (uint16_t)((1u << CL) - 1u);
  constexpr uint32_t IDESC_192 = umma_idesc(1, 128, 192), IDESC_64 = umma_idesc(1, 128, 64);
  constexpr uint32_t TY = 0, TX = 384;  // Y_s = TY + 192 s, X_b = TX + 64 b

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ST_REGS_WG0));
    if (warp < 3) {
      // =========================== ring producers: entry e is loaded by warp e % 3 ===========================
      if (lane == 0) {
        uint32_t k = 0;
        for_each_ring_entry(my_tiles, [&](int kind, int it, int idx) {
          const uint32_t e = k++;
          if ((int)(e % 3u) != warp) return;
          const int s = e % ST_WSLOTS;
          mbar_wait(bar(SB_WEMPTY + s), ((e / ST_WSLOTS) & 1u) ^ 1u);
          const uint32_t dst = sbase + ST_OFF_W + s * ST_WSLOT, fb = bar(SB_WFULL + s);
          if (kind == 0) {
            const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * 128;
            mbar_expect_tx(fb, ST_TILE);
            tma_load_2d(dst, &tmO, fb, idx * 64, row0);
            // pull the tile's residual rows into L2 (the IO warps fetch them into smem about one tile from now)
            if (it >= 2) {
              tma_prefetch_2d(&tmResPf, idx * 64, row0);
              tma_prefetch_2d(&tmResPf, idx * 64 + 32, row0);
            }
          } else {
            // weight entry: this CTA fetches rows [rank, rank + 1) / CL of every box; the whole slot (own part + the peers'
            // multicast parts) completes on the own barrier
            mbar_expect_tx(fb, ST_WSLOT);
            auto wload = [&](uint32_t d, const CUtensorMap* map, int c0, int r0, int rows) {
              const int part = rows / CL;
              if (CL > 1)
                tma_load_2d_mc(d + crank * (uint32_t)part * 128u, map, fb, c0, r0 + (int)crank * part, kClMask);
              else
                tma_load_2d(d, map, fb, c0, r0);
            };
            if (kind == 1) {
              wload(dst, &tmWp, idx * 64, 0, 192);
            } else if (kind == 2) {
              for (int kb = 0; kb < 3; ++kb) wload(dst + kb * 8192, &tmW1, kb * 64, idx * 64, 64);
            } else {
              wload(dst, &tmW2, idx * 64, 0, 192);
            }
          }
        });
      }
    } else {
      // =========================== MMA issuer ===========================
      if (lane == 0 && my_tiles > 0) {
        uint32_t wk = 0;  // ring position
        uint32_t n_hready[2] = {0, 0};
        // `opaque`: keeps ptxas from hoisting descriptors / TMEM addresses out of the loops (they would be spilled on this
        // warpgroup's small register budget)
        auto opaque = [](uint32_t v) {
          asm volatile("" : "+r"(v));
          return v;
        };
        auto w_wait = [&]() -> uint32_t {
          const int s = wk % ST_WSLOTS;
          mbar_wait(bar(SB_WFULL + s), (wk / ST_WSLOTS) & 1u);
          tc_fence_after();
          return opaque(sbase) + ST_OFF_W + s * ST_WSLOT;
        };
        auto slot_free = [&](int s) {  // the slot is free once the MMAs issued so far are done -- in every CTA of the cluster
          if (CL > 1)
            umma_commit_mc(bar(SB_WEMPTY + s), kClMask);
          else
            umma_commit(bar(SB_WEMPTY + s));
        };
        auto w_release = [&]() {
          slot_free(wk % ST_WSLOTS);
          ++wk;
        };
        auto proj = [&](int it) {  // Y_s = o(it) Wproj^T ; s = it & 1
          const uint32_t tY = opaque(tmem_base) + TY + 192u * (uint32_t)(it & 1);
          for (int kb = 0; kb < 3; ++kb) {
            const uint32_t o_addr = w_wait();
            const int so = wk % ST_WSLOTS;
            ++wk;  // the o slot is released together with the weight slot below
            const uint32_t w_addr = w_wait();
            const uint64_t adesc = umma_desc_sw128(o_addr), bdesc = umma_desc_sw128(w_addr);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma<false>(tY, adesc + 2 * k, bdesc + 2 * k, IDESC_192, (kb | k) ? 1u : 0u);
            slot_free(so);
            w_release();
          }
          umma_commit(bar(SB_PFULL + (it & 1)));
        };
        auto fc1_chunk = [&](int c) {  // X_b = xn2 W1[64c : 64c+64]^T
          const int b = c & 1;
          const uint32_t w_addr = w_wait();
          const uint32_t tX = opaque(tmem_base) + TX + 64u * (uint32_t)b;
          for (int kb = 0; kb < 3; ++kb) {
            const uint64_t adesc = umma_desc_sw128(opaque(sbase) + ST_OFF_XN + kb * ST_TILE), bdesc = umma_desc_sw128(w_addr + kb * 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma<false>(tX, adesc + 2 * k, bdesc + 2 * k, IDESC_64, (kb | k) ? 1u : 0u);
          }
          w_release();
          umma_commit(bar(SB_XFULL + b));
        };
        auto fc2_chunk = [&](int it, int c) {  // Y_s += h_c W2[:, 64c : 64c+64]^T, h_c = packed bf16 in X_b columns [0,16) and [32,48)
          const int b = c & 1;
          mbar_wait(bar(SB_HREADY + b), n_hready[b] & 1u);
          ++n_hready[b];
          tc_fence_after();
          const uint64_t bdesc = umma_desc_sw128(w_wait());
          const uint32_t tb = opaque(tmem_base);
          const uint32_t tY = tb + TY + 192u * (uint32_t)(it & 1), tX = tb + TX + 64u * (uint32_t)b;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ts(tY, tX + (uint32_t)((k >> 1) * 32 + (k & 1) * 8), bdesc + 2 * k, IDESC_192, 1u);
          w_release();
        };
        proj(0);
        for (int it = 0; it < my_tiles; ++it) {
          const uint32_t ph = (uint32_t)it & 1u;
          long long* dbg = (a.dbg && it < 32) ? a.dbg + 16 * ((size_t)(2 * gridDim.x + blockIdx.x) * 32 + it) : nullptr;
          if (dbg) dbg[0] = clock64();
          mbar_wait(bar(SB_XNREADY), ph);  // xn2 of tile `it` in smem, Y_s = t' + b2
          tc_fence_after();
          if (dbg) dbg[1] = clock64();
          fc1_chunk(0);
          fc1_chunk(1);
          if (dbg) dbg[2] = clock64();
          if (it + 1 < my_tiles) {  // projection of the next tile into the other Y: its final-epilogue reader is tile it-1
            if (it >= 1) {
              mbar_wait(bar(SB_YFREE + (ph ^ 1u)), ((uint32_t)(it - 1) >> 1) & 1u);
              tc_fence_after();
            }
            proj(it + 1);
          }
          if (dbg) dbg[3] = clock64();
          for (int c = 0; c < ST_NCHUNK; ++c) {
            fc2_chunk(it, c);
            if (c + 2 < ST_NCHUNK) fc1_chunk(c + 2);
            if (c + 2 == ST_NCHUNK - 1) umma_commit(bar(SB_XNFREE));  // all reads of xn2 are done once this commit fires
            if (dbg && c < 6) dbg[4 + c] = clock64();
          }
          umma_commit(bar(SB_YFULL + ph));
        }
      }
      __syncwarp();
    }
  } else if (warp < ST_GELU_WARP0) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ST_REGS_IO));
    // =========================== IO warps (8): E1(j), E3(j-1), E1(j+1), ... ===========================
    const int ew = warp - ST_IO_WARP0;
    const int hf = ew >> 2;     // column half of Y
    const int quad = warp & 3;  // TMEM lane quadrant
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint8_t* io = smem + ST_OFF_IO + ew * ST_IO_WARP;  // B0 | B1
    const uint32_t io_s = smem_u32(io);
    const int rbar = SB_RFULL + ew * 2;
    const float invC = 1.0f / (float)a.C;
    const bool mask_tail = (hf == 1);  // this warp's last chunk holds the padded channels [C, 192)
    uint32_t n_res[2] = {0, 0};        // completed uses of this warp's two residual barriers

    // The warp's two 4 KB boxes carry residual chunks [32 rows][32 cols] fp32 (SWIZZLE_128B): chunks 0 and 1 of a tile sit in
    // B0 / B1 from the moment the previous output stores have drained them until that tile's projection epilogue, chunk 2
    // follows into B0 as soon as chunk 0 has been consumed.
    auto res_issue = [&](int c, int it) {  // lane 0 only; chunk c -> box c & 1
      if (it < my_tiles) {
        const uint32_t b = bar(rbar + (c & 1));
        mbar_expect_tx(b, 4096);
        tma_load_2d(io_s + (c & 1) * 4096, &tmRes, b, hf * 96 + c * 32, ((int)blockIdx.x + it * (int)gridDim.x) * 128 + quad * 32);
      }
    };
    auto drain = [&]() {  // every store issued so far has read its staging data
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
    };

    // ---------------- E1: projection epilogue of local tile `it`: t' = acc + bp + res ; Y <- t' + b2 ; xn2 -> smem ----------------
    auto E1 = [&](int it) {
      const uint32_t tY = tlane + TY + 192u * (uint32_t)(it & 1) + (uint32_t)hf * 96u;
      long long* dbg = (a.dbg && ew == 0 && lane == 0 && it < 32) ? a.dbg + 16 * ((size_t)blockIdx.x * 32 + it) : nullptr;
      f32x2 tv[3][16];  // this thread's 96 values of the row, as (even, odd) column pairs
      if (dbg) dbg[0] = clock64();
      mbar_wait_warp(bar(SB_PFULL + (it & 1)), ((uint32_t)it >> 1) & 1u, lane);
      if (dbg) dbg[1] = clock64();
      tc_fence_after();
      {
        uint32_t raw[3][32];
#pragma unroll
        for (int c = 0; c < 3; ++c) tmem_ld32_nowait(tY + c * 32, raw[c]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < 16; ++i) tv[c][i] = f2_pack_u(raw[c][2 * i], raw[c][2 * i + 1]);
      }
      f32x2 sm[4] = {0ull, 0ull, 0ull, 0ull}, sqv[4] = {0ull, 0ull, 0ull, 0ull};  // independent chains: no 96-deep FADD dependency
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int nb = hf * 96 + c * 32;
        mbar_wait_warp(bar(rbar + (c & 1)), n_res[c & 1] & 1u, lane);
        ++n_res[c & 1];
        const uint8_t* rb = io + (c & 1) * 4096;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r = *reinterpret_cast<const float4*>(rb + sw128_off(lane, j));
          const float4 b = *reinterpret_cast<const float4*>(s_bp + nb + 4 * j);
          tv[c][2 * j] = f2_add(tv[c][2 * j], f2_add(f2_pack(r.x, r.y), f2_pack(b.x, b.y)));
          tv[c][2 * j + 1] = f2_add(tv[c][2 * j + 1], f2_add(f2_pack(r.z, r.w), f2_pack(b.z, b.w)));
        }
        if (c == 0) {  // B0 is free again: fetch the third chunk while the second is processed
          __syncwarp();
          if (lane == 0) res_issue(2, it);
        }
        if (c == 2 && mask_tail) {  // padded channels stay exact zeros
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float lo, hi;
            f2_unpack(tv[c][i], lo, hi);
            tv[c][i] = f2_pack(nb + 2 * i < a.C ? lo : 0.0f, nb + 2 * i + 1 < a.C ? hi : 0.0f);
          }
        }
        // Y = t' + b2: fc2 accumulates on top of the residual
        {
          uint32_t y[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = *reinterpret_cast<const float4*>(s_b2 + nb + 4 * j);
            f2_unpack_u(f2_add(tv[c][2 * j], f2_pack(b.x, b.y)), y[4 * j], y[4 * j + 1]);
            f2_unpack_u(f2_add(tv[c][2 * j + 1], f2_pack(b.z, b.w)), y[4 * j + 2], y[4 * j + 3]);
          }
          tmem_st32_u32(tY + c * 32, y);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          sm[i & 3] = f2_add(sm[i & 3], tv[c][i]);
          sqv[i & 3] = f2_fma(tv[c][i], tv[c][i], sqv[i & 3]);
        }
      }
      if (dbg) dbg[2] = clock64();
      red[row * 2 + hf] = make_float2(f2_hsum(f2_add(f2_add(sm[0], sm[1]), f2_add(sm[2], sm[3]))),
                                      f2_hsum(f2_add(f2_add(sqv[0], sqv[1]), f2_add(sqv[2], sqv[3]))));
      named_bar_sync(1 + quad, 64);
      const float2 r0 = red[row * 2], r1 = red[row * 2 + 1];
      const float mean = (r0.x + r1.x) * invC;
      const float var = fmaxf((r0.y + r1.y) * invC - mean * mean, 0.0f);
      const float rstd = rsqrtf(var + a.eps);
      const f32x2 sc2 = f2_splat(rstd), sh2 = f2_splat(-mean * rstd);
      // the xn2 tile is shared by consecutive tiles: the fc1 MMAs of the previous tile must have read it
      if (it > 0) mbar_wait_warp(bar(SB_XNFREE), (uint32_t)(it - 1) & 1u, lane);
      if (dbg) dbg[3] = clock64();
      // xn2 in the SWIZZLE_128B K-major layout TMA would have produced
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int nb = hf * 96 + c * 32;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            w[i] = f2_to_bf16x2(f2_fma(tv[c][4 * qd + i], sc2, sh2));
            if (c == 2 && mask_tail) {  // padded channels: 1.0 in C, C+1 (they carry fc1's bias: W1 holds it there), zeros after
              const int col = nb + 8 * qd + 2 * i;
              w[i] = col == a.C ? 0x3F803F80u : (col > a.C ? 0u : w[i]);
            }
          }
          const int col = nb + 8 * qd;
          *reinterpret_cast<uint4*>(smem + ST_OFF_XN + (col >> 6) * ST_TILE + sw128_off(row, (col & 63) >> 3)) =
              make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      tmem_wait_st();
      tc_fence_before();    // orders the tcgen05.st of Y before the arrive
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(SB_XNREADY));
      if (dbg) dbg[4] = clock64();
    };

    // ---------------- E3: final epilogue of local tile `it`: t'' = Y ; TMA stores + LayerNorm_next; refill the boxes ----------------
    auto E3 = [&](int it, int it_refill) {
      const uint32_t tY = tlane + TY + 192u * (uint32_t)(it & 1) + (uint32_t)hf * 96u;
      const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * 128 + quad * 32;  // first row of this warp's boxes
      long long* dbg = (a.dbg && ew == 0 && lane == 0 && it < 32) ? a.dbg + 16 * ((size_t)blockIdx.x * 32 + it) : nullptr;
      f32x2 tv[3][16];
      if (dbg) dbg[8] = clock64();
      mbar_wait_warp(bar(SB_YFULL + (it & 1)), ((uint32_t)it >> 1) & 1u, lane);
      tc_fence_after();
      if (dbg) dbg[9] = clock64();
      {
        uint32_t raw[3][32];
#pragma unroll
        for (int c = 0; c < 3; ++c) tmem_ld32_nowait(tY + c * 32, raw[c]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(SB_YFREE + (it & 1)));  // Y_s may take the projection of tile it + 2
        if (mask_tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (hf * 96 + 64 + i >= a.C) raw[2][i] = 0u;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < 16; ++i) tv[c][i] = f2_pack_u(raw[c][2 * i], raw[c][2 * i + 1]);
      }
      auto stage_f32 = [&](int c, int box) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 w;
          f2_unpack_u(tv[c][2 * j], w.x, w.y);
          f2_unpack_u(tv[c][2 * j + 1], w.z, w.w);
          *reinterpret_cast<uint4*>(io + box * 4096 + sw128_off(lane, j)) = w;
        }
      };
      if (a.has_f32) {  // round 1: chunks 0, 1 -> B0, B1
        stage_f32(0, 0);
        stage_f32(1, 1);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOutF, io_s, hf * 96, row0);
          tma_store_2d(&tmOutF, io_s + 4096, hf * 96 + 32, row0);
          bulk_commit();
        }
      }
      if (dbg) dbg[10] = clock64();
      float mean = 0.0f, rstd = 1.0f;
      if (a.do_ln) {  // statistics overlap the stores' drain
        f32x2 sm[4] = {0ull, 0ull, 0ull, 0ull}, sqv[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            sm[i & 3] = f2_add(sm[i & 3], tv[c][i]);
            sqv[i & 3] = f2_fma(tv[c][i], tv[c][i], sqv[i & 3]);
          }
        named_bar_sync(1 + quad, 64);  // the partner has read E1's entries of `red`
        red[row * 2 + hf] = make_float2(f2_hsum(f2_add(f2_add(sm[0], sm[1]), f2_add(sm[2], sm[3]))),
                                        f2_hsum(f2_add(f2_add(sqv[0], sqv[1]), f2_add(sqv[2], sqv[3]))));
        named_bar_sync(1 + quad, 64);
        const float2 r0 = red[row * 2], r1 = red[row * 2 + 1];
        mean = (r0.x + r1.x) * invC;
        const float var = fmaxf((r0.y + r1.y) * invC - mean * mean, 0.0f);
        rstd = rsqrtf(var + a.eps);
        named_bar_sync(1 + quad, 64);  // the partner has read `red` before the next E1 rewrites it
      }
      if (dbg) dbg[11] = clock64();
      // bf16 word (two columns) i of chunk c: LayerNorm_next or a plain copy
      const f32x2 sc2 = f2_splat(a.do_ln ? rstd : 1.0f), sh2 = f2_splat(a.do_ln ? -mean * rstd : 0.0f);
      auto bf_quad = [&](int c, int j) -> uint4 {  // columns [4j*2 .. 4j*2+8) of chunk c -> 16 bytes
        const int nb = hf * 96 + c * 32;
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          f32x2 n0 = f2_fma(tv[c][4 * j + 2 * i], sc2, sh2), n1 = f2_fma(tv[c][4 * j + 2 * i + 1], sc2, sh2);
          if (a.do_ln) {  // pads: gamma = beta = 0
            const float4 gv = *reinterpret_cast<const float4*>(s_g3 + nb + 8 * j + 4 * i);
            const float4 bv = *reinterpret_cast<const float4*>(s_be3 + nb + 8 * j + 4 * i);
            n0 = f2_fma(n0, f2_pack(gv.x, gv.y), f2_pack(bv.x, bv.y));
            n1 = f2_fma(n1, f2_pack(gv.z, gv.w), f2_pack(bv.z, bv.w));
          }
          w[2 * i] = f2_to_bf16x2(n0);
          w[2 * i + 1] = f2_to_bf16x2(n1);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
      };
      drain();  // round 1 has been read
      if (dbg) dbg[12] = clock64();
      // round 2: fp32 chunk 2 -> B0 ; bf16 columns [0,64) of this half (one [32 x 64] SWIZZLE_128B box) -> B1
      if (a.has_f32) stage_f32(2, 0);
      if (a.has_bf) {
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(io + 4096 + sw128_off(lane, j)) = bf_quad(j >> 2, j & 3);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (a.has_f32) tma_store_2d(&tmOutF, io_s, hf * 96 + 64, row0);
        if (a.has_bf) tma_store_2d(&tmOutB2, io_s + 4096, hf * 96, row0);
        bulk_commit();
      }
      drain();
      // round 3: bf16 columns [64,96): a [32 x 32] SWIZZLE_64B box -> B0 ; B1 already takes the refill tile's residual
      if (lane == 0) res_issue(1, it_refill);
      if (a.has_bf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(io + sw64_off(lane, j)) = bf_quad(2, j);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOutB, io_s, hf * 96 + 64, row0);
          bulk_commit();
        }
        drain();
      }
      if (lane == 0) res_issue(0, it_refill);
      if (dbg) dbg[13] = clock64();
    };

    if (my_tiles > 0) {
      if (lane == 0)
        for (int c = 0; c < 2; ++c) res_issue(c, 0);
      E1(0);
      __syncwarp();
      if (lane == 0)  // no final epilogue runs between the first two projection epilogues: refill the boxes here
        for (int c = 0; c < 2; ++c) res_issue(c, 1);
      for (int it = 1; it < my_tiles; ++it) {
        E1(it);
        E3(it - 1, it + 1);
      }
      E3(my_tiles - 1, my_tiles);
      if (lane == 0) bulk_wait_all();  // every store of this warp has landed before the CTA may exit
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ST_REGS_GELU));  // below the launch value of 96: a release
    // =========================== GELU warps (8): h = GELU(acc + b1) -> packed bf16 over its own accumulator ===========================
    const int ew = warp - ST_GELU_WARP0;
    const int hf = ew >> 2;     // 32-column half of the 64-unit chunk
    const int quad = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const f32x2 kC0 = f2_splat(0.79973199f), kC1 = f2_splat(0.03489978f), kHalf = f2_splat(0.5f);
    uint32_t n_xfull[2] = {0, 0};
    for (int it = 0; it < my_tiles; ++it) {
      long long* dbg = (a.dbg && ew == 0 && lane == 0 && it < 32) ? a.dbg + 16 * ((size_t)(gridDim.x + blockIdx.x) * 32 + it) : nullptr;
#pragma unroll 1
      for (int ch = 0; ch < ST_NCHUNK; ++ch) {
        const int b = ch & 1;
        if (dbg) dbg[2 * ch] = clock64();
        mbar_wait_warp(bar(SB_XFULL + b), n_xfull[b] & 1u, lane);
        ++n_xfull[b];
        tc_fence_after();
        if (dbg) dbg[2 * ch + 1] = clock64();
        const uint32_t tx = tlane + TX + 64u * (uint32_t)b + 32u * (uint32_t)hf;  // this warp's 32 hidden columns of the chunk
        uint32_t raw[32];
        tmem_ld32_nowait(tx, raw);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // b1 is already in the accumulator (xn2's channels C, C+1 are 1.0, W1's columns there hold it)
          pk[2 * j + 0] = gelu2_bf16(f2_pack_u(raw[4 * j], raw[4 * j + 1]), kC0, kC1, kHalf);
          pk[2 * j + 1] = gelu2_bf16(f2_pack_u(raw[4 * j + 2], raw[4 * j + 3]), kC0, kC1, kHalf);
        }
        tmem_st16_u32(tx, pk);  // K index 2i, 2i+1 of this half -> column i (low half = even k)
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(SB_HREADY + b));
      }
      if (dbg) dbg[12] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer's commit may still arrive on its barriers
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
  SSR_CHECK(f.C % 2 == 0 && f.C + 2 <= f.CP, SSR_E_INVALID, "swin_tail: C=%d leaves no pad channel pair for the fc1 bias", f.C);
  SSR_CHECK(f.ldres % 4 == 0 && (!f.out_f32 || f.ld_f32 % 4 == 0), SSR_E_INVALID, "swin_tail: fp32 leading dims must be 16-byte multiples");
  CUtensorMap tmO, tmWp, tmW1, tmW2, tmRes, tmResPf, tmOutF, tmOutB, tmOutB2;
  auto map2d = [&](CUtensorMap* m, const void* base, int elem, int cols, int rows, int ld, int box_c, int box_r, int sw) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)ld * elem};
    cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
    return make_tmap(m, base, elem, 2, dims, str, box, sw);
  };
  SSR_TRY(map2d(&tmO, f.o, 2, f.ld_o, f.M, f.ld_o, 64, 128, 128));
  // STUDIOSR_B200_TAIL_CLUSTER=2 switches the weight multicast on.  Measured (profiles/r02_cluster_multicast.txt): it takes
  // 28 % of the kernel's L2 reads away and the tile period does not move (17.7k -> 18.1k cycles: the two CTAs now stall
  // together), i.e. the kernel is not L2-bound; kept as an experiment switch, off by default.
  static int cl = 0;
  if (!cl) {
    const char* e = getenv("STUDIOSR_B200_TAIL_CLUSTER");
    cl = (e && e[0] == '2') ? 2 : 1;
  }
  SSR_TRY(map2d(&tmWp, f.Wp, 2, 192, 192, 192, 64, 192 / cl, 128));
  SSR_TRY(map2d(&tmW1, f.W1, 2, 192, 384, 192, 64, 64 / cl, 128));
  SSR_TRY(map2d(&tmW2, f.W2, 2, 384, 192, 384, 64, 192 / cl, 128));
  SSR_TRY(map2d(&tmRes, f.res, 4, 192, f.M, f.ldres, 32, 32, 128));
  SSR_TRY(map2d(&tmResPf, f.res, 4, 192, f.M, f.ldres, 32, 128, 128));
  const void* outf = f.out_f32 ? (const void*)f.out_f32 : (const void*)f.res;  // placeholder map when unused
  SSR_TRY(map2d(&tmOutF, outf, 4, 192, f.M, f.out_f32 ? f.ld_f32 : f.ldres, 32, 32, 128));
  const void* outb = f.out_ln ? f.out_ln : f.out_T;
  const int ldb = f.out_ln ? f.ld_ln : f.ld_T;
  if (outb) {
    SSR_CHECK(ldb % 8 == 0, SSR_E_INVALID, "swin_tail: bf16 leading dim must be a 16-byte multiple");
    SSR_TRY(map2d(&tmOutB, outb, 2, 192, f.M, ldb, 32, 32, 64));
    SSR_TRY(map2d(&tmOutB2, outb, 2, 192, f.M, ldb, 64, 32, 128));
  } else {
    tmOutB = tmO;
    tmOutB2 = tmO;
  }
  TailArgs a;
  a.M = f.M; a.C = f.C; a.n_tiles = (f.M + 127) / 128;
  a.bp = f.bp; a.b2 = f.b2; a.g3 = f.g3; a.be3 = f.be3;
  a.has_f32 = f.out_f32 != nullptr;
  a.has_bf = outb != nullptr;
  a.do_ln = f.out_ln != nullptr;
  a.eps = f.eps;
  a.dbg = g_tail_dbg;
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CUDA(cudaFuncSetAttribute(swin_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
    SSR_CUDA(cudaFuncSetAttribute(swin_tail_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
    attr_set = true;
  }
  const int sms = num_sms_cached();
  const double flops = 2.0 * f.M * ((double)f.C * f.C + 2.0 * f.C * f.Hid);
  const double bytes = (double)f.M * f.C * (2 + 4 + (f.out_f32 ? 4 : 0) + (f.out_T ? 2 : 0) + (f.out_ln ? 2 : 0));
  ProfScope prof("swin_tail", flops, bytes, s);
  int grid = a.n_tiles < sms ? a.n_tiles : sms;
  if (cl == 1) {
    swin_tail_kernel<1><<<grid, ST_THREADS, ST_SMEM, s>>>(tmO, tmWp, tmW1, tmW2, tmRes, tmResPf, tmOutF, tmOutB, tmOutB2, a);
  } else {
    grid = (grid + 1) / 2 * 2;
    if (grid > sms) grid -= 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(ST_THREADS);
    cfg.dynamicSmemBytes = ST_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SSR_CUDA(cudaLaunchKernelEx(&cfg, swin_tail_kernel<2>, tmO, tmWp, tmW1, tmW2, tmRes, tmResPf, tmOutF, tmOutB, tmOutB2, a));
  }
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
