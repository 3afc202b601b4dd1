// tcgen05 / TMEM / TMA implicit-GEMM kernel with the fused epilogue (sm_100a only).
//
//   D[128 x BLOCK_N] (fp32, TMEM) = sum_kb  A_kb[128 x 128B] (smem, SW128, K-major) * W_kb[BLOCK_N x 128B]^T
//
// A tile is 128 NHWC pixels: 128 consecutive rows (linear) or a BW x BH x BB pixel patch (3x3 conv,
// one TMA box per tap with the tap's (dy,dx) added to the box origin; out-of-image pixels are
// zero-filled by TMA, which *is* the conv's zero padding).  W is the packed [NP][taps*KP] weight.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quadrant each; thread <-> accumulator row).
// Two CTAs are co-resident per SM (<= 113 KB smem, <= 256 TMEM columns each) so one CTA's
// epilogue overlaps the other's TMA/MMA main loop.
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "ssr_tc.cuh"

namespace ssr {

#ifdef SSR_TC_POLL
#define TC_ROLE_WAIT mbar_wait_poll
#else
#define TC_ROLE_WAIT mbar_wait
#endif

// ---------------------------------------------------------------------------------------------
struct TcGeom {
  int conv;              // 0: 128 consecutive rows per tile; 1: BW x BH x BB pixel patch per tile
  int BW, BH, BB;        // patch shape (BW*BH*BB == 128)
  int tiles_x, tiles_y;  // patch grid per image (conv)
  int kc_per_tap;        // KP / BK
  int nkb;               // taps * kc_per_tap
  int m_tiles, n_tiles;  // work items = m_tiles * n_tiles, n fastest (A tile stays hot in L2)
};

constexpr int TC_BM = 128;
constexpr int TC_EPI_WARPS = 8;                      // two groups of 4 (one per TMEM accumulator buffer)
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS + 32;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue, warp 10 second TMA producer
constexpr int TC_MAX_NP = 2304;
// "Halo" mode of the 3x3 conv (bf16): the 128-pixel tile is an 8 x 16 patch and ONE TMA box of (8+2) x (16+2) pixels x 64
// channels serves all nine taps -- the tap's A operand is the same shared-memory tile read from a start address shifted by
// ((dy+1) * 10 + (dx+1)) rows with an 8-row-group stride of 10 rows.  That is legal because the tensor core derives the
// 128B-swizzle phase from absolute shared-memory address bits, exactly like TMA (measured for every shift / pitch:
// scripts/micro_halo.cu, profiles/r01_micro_halo_descriptor.txt).  The conv kernels were bound by the SM's shared-memory
// fill rate (~58-60 B/clk/SM at 40-48 KB per k-block); the halo tile cuts the A-operand fill 6x.
constexpr int TC_HALO_BW = 8, TC_HALO_BH = 16;
constexpr int TC_HALO_PITCH = TC_HALO_BW + 2;                                   // pixels per halo row
constexpr uint32_t TC_HALO_BYTES = (TC_HALO_BH + 2) * TC_HALO_PITCH * 128;      // 23040
constexpr uint32_t TC_HALO_SLOT = 24576;                                        // 1 KB aligned slot per halo tile

template <int BLOCK_N>
constexpr int tc_stages() {
  return BLOCK_N >= 256 ? 3 : BLOCK_N <= 64 ? 6 : 4;  // stage = 16 KB (A) + BLOCK_N*128 B (W); small stages need a deeper ring
}
template <int BLOCK_N>
constexpr size_t tc_smem_bytes() {
  return (size_t)tc_stages<BLOCK_N>() * (16384 + BLOCK_N * 128) + TC_EPI_WARPS * TC_STAGE_BYTES + TC_MAX_NP * 4 +
         2 * 256 * 4 + 256 /*barriers*/ + 1024 /*align slack*/;
}


// 3xTF32 ("tf32x3" precision, fp32 activations and weights): every operand is split into a tf32 head and a tf32 tail,
//   a = a_hi + a_lo (+ O(2^-22 a)),  w = w_hi + w_lo;   D += a_hi w_hi + a_lo w_hi + a_hi w_lo    (three tcgen05.mma per k-step)
// which carries the fp32 claim of the north star (max-abs <= 1e-3; measured ~1e-6) on the tensor cores -- a single tf32 pass sits
// AT that tolerance (1.05e-3).  The weight tails are packed next to the heads (rows [NP, 2 NP) of the packed matrix); the
// activation tile is split in shared memory by four extra warps between TMA arrival and the MMAs (in place: head; second tile:
// tail), so the activations stay un-rounded fp32 in HBM and no producer kernel has to know about the mode.
template <int BLOCK_N>
constexpr size_t tc_smem_bytes_split() {
  return (size_t)2 * (2 * 16384 + 2 * BLOCK_N * 128) + TC_EPI_WARPS * TC_STAGE_BYTES + TC_MAX_NP * 4 + 2 * 256 * 4 + 256 + 1024;
}

template <typename T, int BLOCK_N, bool kHalo, bool kSplit = false>
__global__ void __launch_bounds__(TC_THREADS + (kHalo ? 32 : 0) + (kSplit ? 128 : 0), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmP, const GemmArgs g, const TcGeom geo) {
  constexpr bool kTf32 = sizeof(T) == 4;
  constexpr int BK = 128 / (int)sizeof(T);  // elements per 128-byte swizzle row
  static_assert(!kSplit || (sizeof(T) == 4 && !kHalo), "the 3xTF32 split exists for fp32 operands only");
  constexpr int kStages = kSplit ? 2 : tc_stages<BLOCK_N>();
  constexpr int kTmemCols = 2 * BLOCK_N <= 64 ? 64 : 2 * BLOCK_N <= 128 ? 128 : 2 * BLOCK_N <= 256 ? 256 : 512;
  constexpr uint32_t A_BYTES = TC_BM * 128 * (kSplit ? 2 : 1), W_BYTES = BLOCK_N * 128 * (kSplit ? 2 : 1);  // split: head | tail
  constexpr uint32_t A_HALF = TC_BM * 128, W_HALF = BLOCK_N * 128;
  constexpr uint32_t idesc = umma_idesc(kTf32 ? 2 : 1, TC_BM, BLOCK_N);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // pointer arithmetic (not an integer round trip) keeps the shared address space visible to the compiler: LDS/STS
  // instead of generic LD/ST for every staging access
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smA = smem;
  uint8_t* smW = smem + kStages * A_BYTES;
  float* stage_all = reinterpret_cast<float*>(smem + kStages * (A_BYTES + W_BYTES));
  float* s_bias = stage_all + TC_EPI_WARPS * (TC_STAGE_BYTES / 4);
  float* s_gamma = s_bias + TC_MAX_NP;
  float* s_beta = s_gamma + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + 256);
  // bars: [0,kStages) full, [kStages,2kStages) empty, then tmem_full[2], tmem_empty[2]; then the TMEM base address
  // halo mode: + full[2], empty[2] of the two halo-tile slots (they live in the A region of the stage ring)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * kStages + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * kStages + 2 + b); };
  // halo tiles live in the A region of the stage ring: as many slots as fit (narrow tiles have 6 stages -> 4 slots, which
  // they need: nine N=64 k-blocks last only ~1.7k cycles, less than one TMA round trip)
  constexpr int kHaloSlots = (kStages * A_BYTES) / TC_HALO_SLOT >= 4 ? 4 : 2;
  auto afull_bar = [&](int b) { return bar0 + 8u * (2 * kStages + 4 + b); };
  auto aempty_bar = [&](int b) { return bar0 + 8u * (2 * kStages + 8 + b); };
  static_assert(!kHalo || (kHaloSlots * TC_HALO_SLOT <= kStages * A_BYTES), "the halo tiles must fit the A region of the ring");

  for (int i = threadIdx.x; i < g.NP; i += blockDim.x) s_bias[i] = __ldg(g.bias + i);
  if (g.out_ln)
    for (int i = threadIdx.x; i < g.NP; i += blockDim.x) {
      s_gamma[i] = __ldg(g.gamma + i);
      s_beta[i] = __ldg(g.beta + i);
    }
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);  // one arrival per epilogue warp of the group
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(afull_bar(b), kSplit ? 4 : 1);  // split mode: "stage b has been split" (one arrival per splitter warp)
      mbar_init(aempty_bar(b), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = geo.m_tiles * geo.n_tiles;

  if (warp == 0 || warp == 2 + TC_EPI_WARPS) {
    // =========================== TMA producers ===========================
    // One thread completes only one wait->issue round per ~500 cycles whatever the ring depth (measured,
    // profiles/r01_micro_tc.txt), less than a k-block's MMAs take, so even and odd k-blocks get their own thread.
    const uint32_t pid = warp == 0 ? 0u : 1u;
    if (lane == 0) {
      uint32_t kbg = 0;  // running k-block counter across items (ring position)
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mt = item / geo.n_tiles, nt = item - mt * geo.n_tiles;
        const int n0 = nt * BLOCK_N;
        int m0 = 0, tx0 = 0, ty0 = 0, tb0 = 0;
        if (geo.conv) {
          int t = mt;
          tx0 = (t % geo.tiles_x) * geo.BW;
          t /= geo.tiles_x;
          ty0 = (t % geo.tiles_y) * geo.BH;
          tb0 = (t / geo.tiles_y) * geo.BB;
        } else {
          m0 = mt * TC_BM;
        }
        for (int kb = 0; kb < geo.nkb; ++kb, ++kbg) {
          if ((kbg & 1u) != pid) continue;
          const int s = kbg % kStages;
          const uint32_t ph = (kbg / kStages) & 1u;
          TC_ROLE_WAIT(empty_bar(s), ph ^ 1u);
          const uint32_t dstA = smem_u32(smA + s * A_BYTES), dstW = smem_u32(smW + s * W_BYTES);
          if constexpr (kHalo) {
            // k-blocks run channel-block major (cb, tap): the halo tile of a channel block lasts nine k-blocks, only W streams
            const int cb = kb / 9, tap = kb - cb * 9;
            mbar_expect_tx(full_bar(s), W_BYTES);
            tma_load_2d(dstW, &tmW, full_bar(s), (tap * geo.kc_per_tap + cb) * BK, n0);
          } else {
            mbar_expect_tx(full_bar(s), A_HALF + W_BYTES);
            const int tap = kb / geo.kc_per_tap, kc = kb - tap * geo.kc_per_tap;
            if (geo.conv) {
              const int dy = (g.taps == 9) ? tap / 3 - 1 : 0, dx = (g.taps == 9) ? tap % 3 - 1 : 0;
              tma_load_4d(dstA, &tmA, full_bar(s), kc * BK, tx0 + dx, ty0 + dy, tb0);
            } else {
              tma_load_2d(dstA, &tmA, full_bar(s), kc * BK, m0);
            }
            tma_load_2d(dstW, &tmW, full_bar(s), kb * BK, n0);
            if constexpr (kSplit) tma_load_2d(dstW + W_HALF, &tmW, full_bar(s), kb * BK, g.NP + n0);  // the tf32 tails of the weights
          }
        }
      }
    }
  } else if (kSplit && warp >= 3 + TC_EPI_WARPS) {
    // =========================== activation splitter (3xTF32 only): a -> (tf32 head in place, tf32 tail) ===========================
    const int t = threadIdx.x - 32 * (3 + TC_EPI_WARPS);  // 0..127 = tile row
    uint32_t kbg = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      for (int kb = 0; kb < geo.nkb; ++kb, ++kbg) {
        const int s = kbg % kStages;
        mbar_wait_warp(full_bar(s), (kbg / kStages) & 1u, lane);
        uint8_t* hi = smA + s * A_BYTES + t * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = (j ^ (t & 7)) * 16;  // rows are 128 bytes apart: rotate the chunk order so a quarter-warp hits 8 bank groups
          float4 v = *reinterpret_cast<float4*>(hi + c);
          float4 h4 = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
          float4 l4 = make_float4(round_tf32(v.x - h4.x), round_tf32(v.y - h4.y), round_tf32(v.z - h4.z), round_tf32(v.w - h4.w));
          *reinterpret_cast<float4*>(hi + c) = h4;
          *reinterpret_cast<float4*>(hi + A_HALF + c) = l4;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(afull_bar(s));
      }
    }
  } else if (kHalo && warp == 3 + TC_EPI_WARPS) {
    // =========================== halo-tile producer (halo mode only) ===========================
    if (lane == 0) {
      uint32_t ag = 0;  // running halo-tile counter (slot = ag & 1)
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int t = item / geo.n_tiles;
        const int tx0 = (t % geo.tiles_x) * geo.BW;
        t /= geo.tiles_x;
        const int ty0 = (t % geo.tiles_y) * geo.BH, tb0 = t / geo.tiles_y;
        for (int cb = 0; cb < geo.kc_per_tap; ++cb, ++ag) {
          const int sa = ag % kHaloSlots;
          TC_ROLE_WAIT(aempty_bar(sa), ((ag / kHaloSlots) & 1u) ^ 1u);
          mbar_expect_tx(afull_bar(sa), TC_HALO_BYTES);
          tma_load_4d(smem_u32(smA + sa * TC_HALO_SLOT), &tmA, afull_bar(sa), cb * BK, tx0 - 1, ty0 - 1, tb0);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      uint32_t kbg = 0;
      uint32_t ag0 = 0;  // halo mode: running halo-tile counter at the start of the item
      int il = 0;  // local item index
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++il) {
        const int b = il & 1;
        TC_ROLE_WAIT(tempty_bar(b), ((uint32_t)(il >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * BLOCK_N);
        for (int kb = 0; kb < geo.nkb; ++kb, ++kbg) {
          const int s = kbg % kStages;
          const uint32_t ph = (kbg / kStages) & 1u;
          uint64_t adesc;
          if constexpr (kHalo) {
            const int cb = kb / 9, tap = kb - cb * 9;
            const uint32_t ag = ag0 + (uint32_t)cb;
            const int sa = ag % kHaloSlots;
            if (tap == 0) TC_ROLE_WAIT(afull_bar(sa), (ag / kHaloSlots) & 1u);
            // same tile, start shifted by (dy+1) halo rows and (dx+1) pixels; 8-row groups (= patch rows) 10 pixels apart
            const uint32_t a_addr = smem_u32(smA + sa * TC_HALO_SLOT) + (uint32_t)((tap / 3) * TC_HALO_PITCH + tap % 3) * 128u;
            adesc = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((TC_HALO_PITCH * 128) >> 4) << 32) | (1ull << 46) |
                    (2ull << 61);
          } else {
            adesc = umma_desc_sw128(smem_u32(smA + s * A_BYTES));
          }
          TC_ROLE_WAIT(full_bar(s), ph);
          if constexpr (kSplit) TC_ROLE_WAIT(afull_bar(s), ph);  // the splitter warps have written head and tail
          tc_fence_after();
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smW + s * W_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of K per 128-byte row: +2 in the (addr >> 4) field
            umma<kTf32>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            if constexpr (kSplit) {  // small terms: a_lo w_hi and a_hi w_lo
              const uint64_t adesc_lo = umma_desc_sw128(smem_u32(smA + s * A_BYTES + A_HALF));
              const uint64_t bdesc_lo = umma_desc_sw128(smem_u32(smW + s * W_BYTES + W_HALF));
              umma<kTf32>(tmem_d, adesc_lo + 2 * k, bdesc + 2 * k, idesc, 1u);
              umma<kTf32>(tmem_d, adesc + 2 * k, bdesc_lo + 2 * k, idesc, 1u);
            }
          }
          umma_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
          if constexpr (kHalo) {
            if (kb % 9 == 8) umma_commit(aempty_bar((ag0 + (uint32_t)(kb / 9)) % kHaloSlots));  // the halo tile has served its nine taps
          }
        }
        if constexpr (kHalo) ag0 += (uint32_t)geo.kc_per_tap;
        umma_commit(tfull_bar(b));  // accumulator complete
      }
    }
    __syncwarp();
  } else if (warp < 2 + TC_EPI_WARPS) {
    // =========================== epilogue ===========================
    const int ew = warp - 2;
    const int grp = ew >> 2;    // accumulator buffer / item parity served by this warp
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    float* st = stage_all + ew * (TC_STAGE_BYTES / 4);
    const int Cps = g.ps_r > 1 ? g.N / (g.ps_r * g.ps_r) : 0;
    const bool do_ln = g.out_ln != nullptr;
    T* outT = reinterpret_cast<T*>(g.out_T);
    T* oln = reinterpret_cast<T*>(g.out_ln);

    int il = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++il) {
      if ((il & 1) != grp) continue;
      const int mt = item / geo.n_tiles, nt = item - mt * geo.n_tiles;
      const int n0 = nt * BLOCK_N;
      int m = -1;  // this thread's row as a pixel / token index, -1 if outside the tensor
      int tx0 = 0, ty0 = 0, tb0 = 0;  // patch origin (conv)
      if (geo.conv) {
        int t = mt;
        tx0 = (t % geo.tiles_x) * geo.BW;
        t /= geo.tiles_x;
        ty0 = (t % geo.tiles_y) * geo.BH;
        tb0 = (t / geo.tiles_y) * geo.BB;
        const int ww = row % geo.BW, hh = (row / geo.BW) % geo.BH, bb = row / (geo.BW * geo.BH);
        const int pb = tb0 + bb, py = ty0 + hh, px = tx0 + ww;
        if (pb < g.B && py < g.H && px < g.W) m = (pb * g.H + py) * g.W + px;
      } else {
        const int mm = mt * TC_BM + row;
        if (mm < g.M) m = mm;
      }
      const RowMap rm = make_rowmap(m, lane);
      // destination of transposed-view row i for the T-typed outputs (plain rows or pixel-shuffled)
      auto dstT = [&](T* base, int ld, int mi, int nb) -> T* {
        if (mi < 0) return nullptr;
        if (g.ps_r > 1) {
          const int px = mi % g.W, py = (mi / g.W) % g.H, pb = mi / (g.W * g.H);
          return base + ps_offset(pb, py, px, nb, g.H, g.W, g.ps_r, Cps, ld);
        }
        return base + (size_t)mi * ld + nb;
      };
      auto store_T = [&](float (&v)[32], T* base, int ld, int nb) {
        if constexpr (sizeof(T) == 2) {
          stage_store_bf16(st, lane, v, [&](int i) { return dstT(base, ld, rm.mh[i], nb); });
        } else {
          stage_store_tf32(st, lane, v, [&](int i) { return dstT(base, ld, rm.mf[i], nb); }, !kSplit);
        }
      };

      // plain-rows bf16 outputs (linear layers): pack the chunk into a SWIZZLE_64B box [32 rows][32 cols] of this warp's staging
      // area and let TMA write it (rows beyond M are clipped); no read-back, no per-lane global stores.  Two boxes per warp.
      auto tma_store_chunk = [&](const float (&v)[32], const CUtensorMap* map, int buf, int nb_) {
        uint8_t* box = reinterpret_cast<uint8_t*>(st) + buf * 2048;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(box + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                         pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (g.tma_store == 2) {
            // nn.PixelShuffle folded into the store: the warp's 32 pixels are a (bw x bh) sub-patch, the 32 columns one (i, j)
            // sub-position and half of its channels; the output is addressed as (c, j, x, i, b*H + y) -- a 5-D tensor map
            const int Cps_ = g.N / (g.ps_r * g.ps_r), q_ = nb_ / Cps_, r0_ = quad * 32;
            tma_store_5d(map, smem_u32(box), nb_ - q_ * Cps_, q_ % g.ps_r, tx0 + r0_ % geo.BW, q_ / g.ps_r,
                         tb0 * g.H + ty0 + r0_ / geo.BW);
          } else {
            tma_store_2d(map, smem_u32(box), nb_, mt * TC_BM + quad * 32);
          }
          bulk_commit();
        }
      };
      float4 resv[8];
      if (g.res) res_prefetch(resv, rm, lane, g.res, g.ldres, n0);  // overlaps the wait for the accumulator
      // activation-backward mask (training dgrad): this thread's row, 32 columns = 4 x 16 B (bf16), fetched one chunk ahead
      // so that the HBM latency hides behind the previous chunk (un-prefetched it cost ~0.8 us per chunk: 154 -> 60 us)
      uint4 mq[4] = {};
      auto mask_fetch = [&](int nb_) {
        if constexpr (sizeof(T) == 2) {
          if (g.mask && m >= 0) {
            const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(g.mask) + (size_t)m * g.ld_mask + nb_);
#pragma unroll
            for (int q = 0; q < 4; ++q) mq[q] = __ldg(mp + q);
          }
        }
      };
      mask_fetch(n0);

      long long* dbg = (g.dbg && (ew & 3) == 0 && lane == 0 && il < 64) ? g.dbg + 16 * ((size_t)blockIdx.x * 64 + il) : nullptr;
      if (dbg) dbg[0] = clock64();
      mbar_wait(tfull_bar(grp), (uint32_t)(il >> 1) & 1u);
      tc_fence_after();
      if (dbg) dbg[1] = clock64();
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * BLOCK_N);
      float sum = 0.0f, sumsq = 0.0f;

      if (g.out3_f32 || g.out3_u8) {
        // reconstruction conv (Cout = 3): bias + output affine + crop + fp32 NCHW / uint8 HWC store straight from the
        // first three accumulator columns; consecutive lanes are consecutive pixels of a row, so the stores coalesce
        float v[32];
        tmem_ld32(trow, v);
        if (m >= 0) {
          const int px = m % g.W, py = (m / g.W) % g.H, pb = m / (g.W * g.H);
          if (py < g.crop_h && px < g.crop_w) {
            float r[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) r[c] = (v[c] + s_bias[c] + g.out_shift[c]) * g.out_scale;
            if (g.out3_f32) {
              const size_t plane = (size_t)g.crop_h * g.crop_w;
              float* o = g.out3_f32 + (size_t)pb * 3 * plane + (size_t)py * g.crop_w + px;
              o[0] = r[0];
              o[plane] = r[1];
              o[2 * plane] = r[2];
            }
            if (g.out3_u8) {
              uint8_t* o = g.out3_u8 + ((size_t)(pb * g.crop_h + py) * g.crop_w + px) * 3;
#pragma unroll
              for (int c = 0; c < 3; ++c) o[c] = (uint8_t)fminf(fmaxf(rintf(r[c] * g.u8_scale), 0.0f), 255.0f);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(grp));
        continue;
      }

#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int nb = n0 + c * 32;
        float v[32];
        if (dbg && c == 1) dbg[4] = clock64();
        tmem_ld32(trow + c * 32, v);  // (requesting chunk c + 1 here, ahead of its use, was measured 30 % SLOWER: DESIGN.md 5.3)
        if (dbg && c == 1) dbg[5] = clock64();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bv = *reinterpret_cast<const float4*>(s_bias + nb + 4 * q);  // warp-wide broadcast
          v[4 * q + 0] += bv.x;
          v[4 * q + 1] += bv.y;
          v[4 * q + 2] += bv.z;
          v[4 * q + 3] += bv.w;
        }
        if (g.out_pre && g.pre_mode == 1) {  // GELU and its derivative in one pass (training forward of fc1)
          float gp[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = v[i], ax = fabsf(x) * 0.70710678118654752440f;
            float t, e;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));  // exp(-x^2 / 2)
            float p = fmaf(1.061405429f, t, -1.453152027f);
            p = fmaf(p, t, 1.421413741f);
            p = fmaf(p, t, -0.284496736f);
            p = fmaf(p, t, 0.254829592f);
            const float cdf = 0.5f * (1.0f + copysignf(fmaf(-p * t, e, 1.0f), x));
            gp[i] = fmaf(x * 0.3989422804014327f, e, cdf);
            v[i] = x * cdf * g.alpha;
          }
          if (g.tma_store) {
            if (lane == 0) bulk_wait_read<0>();  // both boxes of the previous chunk have been read by TMA
            __syncwarp();
            tma_store_chunk(gp, &tmP, 0, nb);
          } else {
            store_T(gp, reinterpret_cast<T*>(g.out_pre), g.ld_pre, nb);
          }
        } else {
          if (g.out_pre) {
            if (g.tma_store) {
              if (lane == 0) bulk_wait_read<0>();
              __syncwarp();
              tma_store_chunk(v, &tmP, 0, nb);
            } else {
              store_T(v, reinterpret_cast<T*>(g.out_pre), g.ld_pre, nb);
            }
          }
          epilogue_act(v, g.act, g.slope, g.alpha);
        }
        if (g.mask) {  // activation backward (training dgrad): gate by the sign of the saved forward output
          float mk[32];
          if (m >= 0) {
            const T* mp = reinterpret_cast<const T*>(g.mask) + (size_t)m * g.ld_mask + nb;
            if constexpr (sizeof(T) == 2) {
              (void)mp;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u = mq[q];
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  mk[8 * q + 2 * j] = __uint_as_float(w4[j] << 16);
                  mk[8 * q + 2 * j + 1] = __uint_as_float(w4[j] & 0xffff0000u);
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(mp) + q);
                mk[4 * q] = u.x; mk[4 * q + 1] = u.y; mk[4 * q + 2] = u.z; mk[4 * q + 3] = u.w;
              }
            }
            if (c + 1 < BLOCK_N / 32) mask_fetch(nb + 32);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (g.mask_mode == 2) {
                v[i] *= mk[i];
              } else if (g.mask_mode == 1) {  // d/du [u Phi(u)] = Phi(u) + u phi(u)
                const float u = mk[i];
                const float cdf = 0.5f * (1.0f + fast_erf(u * 0.70710678118654752440f));
                v[i] *= fmaf(u * 0.3989422804014327f, __expf(-0.5f * u * u), cdf);
              } else {
                v[i] = mk[i] > 0.0f ? v[i] : v[i] * g.mask_slope;
              }
            }
          }
        }
        if (g.row_scale && m >= 0) {
          const float rs = __ldg(g.row_scale + m / g.rows_per_scale);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= rs;
        }
        if (dbg && c == 1) dbg[6] = clock64();
        if (g.res || g.out_f32) {
          // transposed domain: add the (prefetched, coalesced) residual and store the fp32 stream
          stage_put_f32(st, lane, v);
          __syncwarp();
          stage_add_store_f32(st, lane, rm, g.res ? resv : nullptr, g.out_f32, g.ld_f32, nb);
          if (g.res && c + 1 < BLOCK_N / 32) res_prefetch(resv, rm, lane, g.res, g.ldres, nb + 32);
          __syncwarp();
          if (g.res) stage_get_f32(st, lane, v);
          __syncwarp();
        }
        if (nb + 32 > g.N) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (nb + i >= g.N) v[i] = 0.0f;
        }
        if (do_ln) {  // one-pass statistics (pad columns are exact zeros and add nothing)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            sum += v[i];
            sumsq = fmaf(v[i], v[i], sumsq);
          }
          tmem_st32(trow + c * 32, v);  // keep v for the normalisation pass
        }
        if (outT) {
          if (g.tma_store) {
            if (g.out_pre) {
              tma_store_chunk(v, &tmO, 1, nb);  // box 0 holds out_pre of this chunk (waited for above)
            } else {
              if (lane == 0) bulk_wait_read<1>();  // the box written two chunks ago is free again
              __syncwarp();
              tma_store_chunk(v, &tmO, c & 1, nb);
            }
          } else {
            store_T(v, outT, g.ld_T, nb);
          }
        }
        if (dbg && c == 1) dbg[7] = clock64();
      }
      if (dbg) dbg[2] = clock64();

      if (do_ln) {  // host guarantees n_tiles == 1 and BLOCK_N == NP: the thread owns the whole row
        const float mean = sum / (float)g.N;
        const float sq = fmaxf(sumsq - sum * mean, 0.0f);  // sum (v - mean)^2
        const float rstd = rsqrtf(sq / (float)g.N + g.eps);
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          float v[32];
          tmem_ld32(trow + c * 32, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 gv = *reinterpret_cast<const float4*>(s_gamma + c * 32 + 4 * q);
            const float4 bv = *reinterpret_cast<const float4*>(s_beta + c * 32 + 4 * q);
            v[4 * q + 0] = (v[4 * q + 0] - mean) * rstd * gv.x + bv.x;
            v[4 * q + 1] = (v[4 * q + 1] - mean) * rstd * gv.y + bv.y;
            v[4 * q + 2] = (v[4 * q + 2] - mean) * rstd * gv.z + bv.z;
            v[4 * q + 3] = (v[4 * q + 3] - mean) * rstd * gv.w + bv.w;
          }
          // LN output is always plain rows (never pixel-shuffled)
          if constexpr (sizeof(T) == 2) {
            stage_store_bf16(st, lane, v, [&](int i) { return rm.mh[i] >= 0 ? oln + (size_t)rm.mh[i] * g.ld_ln + c * 32 : nullptr; });
          } else {
            stage_store_tf32(st, lane, v, [&](int i) { return rm.mf[i] >= 0 ? oln + (size_t)rm.mf[i] * g.ld_ln + c * 32 : nullptr; }, !kSplit);
          }
        }
      }
      if (dbg) dbg[3] = clock64();
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld inside tmem_ld32): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(grp));
    }
    if (g.tma_store && lane == 0) bulk_wait_all();  // the staging boxes must outlive their TMA stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* map, const void* base, int elem, int rank, const cuuint64_t* dims,
              const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  SSR_CHECK(fn != nullptr, SSR_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapDataType dt = elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSR_CHECK(r == CUDA_SUCCESS, SSR_E_CUDA, "cuTensorMapEncodeTiled failed (%d), rank %d base %p dims %llu,%llu", (int)r,
            rank, base, (unsigned long long)dims[0], (unsigned long long)dims[1]);
  return SSR_OK;
}

int num_sms_cached() {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      num_sms = 148;
  }
  return num_sms;
}

// choose a BW x BH x BB = 128 pixel patch that wastes the fewest out-of-image pixels
static void choose_patch(int B, int H, int W, int* bw, int* bh, int* bb) {
  long long best = -1;
  for (int w = 1; w <= 128; w *= 2)
    for (int h = 1; w * h <= 128; h *= 2) {
      const int b = 128 / (w * h);
      if (w > 2 * W || h > 2 * H) continue;
      const long long cost = (long long)((W + w - 1) / w) * ((H + h - 1) / h) * ((B + b - 1) / b);
      // prefer wide patches (longer contiguous runs) on ties
      if (best < 0 || cost < best || (cost == best && w > *bw)) {
        best = cost;
        *bw = w;
        *bh = h;
        *bb = b;
      }
    }
}

template <typename T, int BLOCK_N, bool kSplit = false>
static int launch_tc_bn(const GemmArgs& g, cudaStream_t s) {
  constexpr int elem = (int)sizeof(T);
  constexpr int BK = 128 / elem;
  SSR_CHECK(g.KP % BK == 0 && g.lda % (16 / elem) == 0, SSR_E_INVALID, "gemm_tc: KP=%d lda=%d", g.KP, g.lda);
  TcGeom geo{};
  geo.kc_per_tap = g.KP / BK;
  geo.nkb = g.taps * geo.kc_per_tap;
  CUtensorMap tmA, tmW;
  int grid_x;
  SSR_CHECK(g.NP <= TC_MAX_NP, SSR_E_INVALID, "gemm_tc: NP=%d > %d", g.NP, TC_MAX_NP);
  bool halo = false;
  if (g.taps == 9) {
    geo.conv = 1;
    choose_patch(g.B, g.H, g.W, &geo.BW, &geo.BH, &geo.BB);
    geo.tiles_x = (g.W + geo.BW - 1) / geo.BW;
    geo.tiles_y = (g.H + geo.BH - 1) / geo.BH;
    grid_x = geo.tiles_x * geo.tiles_y * ((g.B + geo.BB - 1) / geo.BB);
    // Halo mode pays where the kernel is shared-memory-fill bound: wide tiles (BLOCK_N = 256: EDSR body / upsampler convs,
    // +8 % measured) whose images tile exactly into 8 x 16 patches.  Narrower tiles are epilogue- or latency-bound and got
    // slower; STUDIOSR_B200_HALO=1 / =0 forces it on / off for experiments.
    const char* henv = getenv("STUDIOSR_B200_HALO");
    const bool halo_auto = BLOCK_N == 256 && g.W % TC_HALO_BW == 0 && g.H % TC_HALO_BH == 0;
    if (elem == 2 && (henv ? henv[0] == '1' : halo_auto)) {
      // halo mode needs the 8 x 16 patch; take it unless it wastes > 25 % more out-of-image rows than the best free patch
      const long long hx = (g.W + TC_HALO_BW - 1) / TC_HALO_BW, hy = (g.H + TC_HALO_BH - 1) / TC_HALO_BH;
      if (hx * hy * g.B * 4 <= (long long)grid_x * 5) {
        halo = true;
        geo.BW = TC_HALO_BW;
        geo.BH = TC_HALO_BH;
        geo.BB = 1;
        geo.tiles_x = (int)hx;
        geo.tiles_y = (int)hy;
        grid_x = (int)(hx * hy * g.B);
      }
    }
    cuuint64_t dims[4] = {(cuuint64_t)g.lda, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
    cuuint64_t str[3] = {(cuuint64_t)g.lda * elem, (cuuint64_t)g.W * g.lda * elem, (cuuint64_t)g.H * g.W * g.lda * elem};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)geo.BW, (cuuint32_t)geo.BH, (cuuint32_t)geo.BB};
    if (halo) {  // one box = the patch plus its one-pixel halo; out-of-image pixels are zero-filled (= the conv padding)
      box[1] = TC_HALO_BW + 2;
      box[2] = TC_HALO_BH + 2;
    }
    SSR_TRY(make_tmap(&tmA, g.A, elem, 4, dims, str, box));
  } else {
    geo.conv = 0;
    geo.BW = 128;
    geo.BH = geo.BB = 1;
    grid_x = (g.M + TC_BM - 1) / TC_BM;
    cuuint64_t dims[2] = {(cuuint64_t)g.lda, (cuuint64_t)g.M};
    cuuint64_t str[1] = {(cuuint64_t)g.lda * elem};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)TC_BM};
    SSR_TRY(make_tmap(&tmA, g.A, elem, 2, dims, str, box));
  }
  {
    const cuuint64_t Ktot = (cuuint64_t)g.taps * g.KP;
    cuuint64_t dims[2] = {Ktot, (cuuint64_t)g.NP * (kSplit ? 2 : 1)};  // split: rows [NP, 2 NP) hold the tf32 tails
    cuuint64_t str[1] = {Ktot * elem};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BLOCK_N};
    SSR_TRY(make_tmap(&tmW, g.Wt, elem, 2, dims, str, box));
  }
  // TMA-store epilogue: bf16 plain-row outputs of linear layers without residual / LayerNorm / fp32 outputs
  GemmArgs ga = g;
  CUtensorMap tmO, tmP;
  memset(&tmO, 0, sizeof(tmO));
  memset(&tmP, 0, sizeof(tmP));
  ga.tma_store = 0;
  if (elem == 2 && g.taps == 1 && g.out_T && !g.out_f32 && !g.res && !g.out_ln && g.ps_r <= 1 && !g.out3_f32 && !g.out3_u8 &&
      g.ld_T % 8 == 0 && (!g.out_pre || g.ld_pre % 8 == 0) && !getenv("STUDIOSR_B200_NO_TMA_STORE")) {
    cuuint32_t box[2] = {32, 32};
    {
      cuuint64_t dims[2] = {(cuuint64_t)g.NP, (cuuint64_t)g.M};
      cuuint64_t str[1] = {(cuuint64_t)g.ld_T * 2};
      SSR_TRY(make_tmap(&tmO, g.out_T, 2, 2, dims, str, box, 64));
    }
    if (g.out_pre) {
      cuuint64_t dims[2] = {(cuuint64_t)g.NP, (cuuint64_t)g.M};
      cuuint64_t str[1] = {(cuuint64_t)g.ld_pre * 2};
      SSR_TRY(make_tmap(&tmP, g.out_pre, 2, 2, dims, str, box, 64));
    }
    ga.tma_store = 1;
  } else if (elem == 2 && g.taps == 9 && g.ps_r > 1 && g.out_T && !g.out_f32 && !g.res && !g.out_ln && !g.out_pre && geo.BB == 1 &&
             g.H % geo.BH == 0 && g.ld_T % 8 == 0 && !getenv("STUDIOSR_B200_NO_TMA_STORE")) {
    // pixel-shuffled output [B][H r][W r][ld] seen as (c, j, x, i, t = b*H + y); a warp stores a (32 ch, 1, bw, 1, bh) box
    const int r = g.ps_r, Cps = g.N / (r * r);
    const int bwp = geo.BW < 32 ? geo.BW : 32, bhp = 32 / bwp;
    const cuuint64_t ldb = (cuuint64_t)g.ld_T * 2;
    cuuint64_t dims[5] = {(cuuint64_t)Cps, (cuuint64_t)r, (cuuint64_t)g.W, (cuuint64_t)r, (cuuint64_t)g.B * g.H};
    cuuint64_t str[4] = {ldb, ldb * r, ldb * r * g.W, ldb * r * g.W * r};
    cuuint32_t box[5] = {32, 1, (cuuint32_t)bwp, 1, (cuuint32_t)bhp};
    SSR_TRY(make_tmap(&tmO, g.out_T, 2, 5, dims, str, box, 64));
    ga.tma_store = 2;
  }
  constexpr size_t smem = kSplit ? tc_smem_bytes_split<BLOCK_N>() : tc_smem_bytes<BLOCK_N>();
  static_assert(smem <= 232448, "gemm_tc: shared memory budget");
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<T, BLOCK_N, false, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if constexpr (elem == 2)
      SSR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<T, BLOCK_N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  geo.m_tiles = grid_x;
  geo.n_tiles = g.NP / BLOCK_N;
  const int items = geo.m_tiles * geo.n_tiles;
  const int num_sms = num_sms_cached();
  dim3 grid(items < num_sms ? items : num_sms);
  ProfScope prof((g.out3_f32 || g.out3_u8) ? "gemm_tc_conv_last" : g.taps == 9 ? "gemm_tc_conv3x3" : "gemm_tc_linear", gemm_alg_flops(g),
                 gemm_alg_bytes(g, elem), s);
  if constexpr (elem == 2) {
    if (halo)
      gemm_tc_kernel<T, BLOCK_N, true><<<grid, TC_THREADS + 32, smem, s>>>(tmA, tmW, tmO, tmP, ga, geo);
    else
      gemm_tc_kernel<T, BLOCK_N, false><<<grid, TC_THREADS, smem, s>>>(tmA, tmW, tmO, tmP, ga, geo);
  } else {
    gemm_tc_kernel<T, BLOCK_N, false, kSplit><<<grid, TC_THREADS + (kSplit ? 128 : 0), smem, s>>>(tmA, tmW, tmO, tmP, ga, geo);
  }
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

template <typename T>
static int launch_tc_t(const GemmArgs& g, cudaStream_t s) {
  SSR_CHECK(g.NP % 64 == 0, SSR_E_INVALID, "gemm_tc: NP=%d must be a multiple of 64", g.NP);
  if (g.ps_r > 1) {
    const int Cps = g.N / (g.ps_r * g.ps_r);
    SSR_CHECK(Cps % 32 == 0 && g.N == g.NP, SSR_E_INVALID, "gemm_tc: pixel-shuffle store needs N/r^2 %% 32 == 0 (N=%d r=%d)",
              g.N, g.ps_r);
  }
  if (g.out_ln) {
    SSR_CHECK(g.NP <= 256, SSR_E_INVALID, "gemm_tc: fused LayerNorm needs NP<=256 (NP=%d)", g.NP);
    switch (g.NP) {
      case 64: return launch_tc_bn<T, 64>(g, s);
      case 128: return launch_tc_bn<T, 128>(g, s);
      case 192: return launch_tc_bn<T, 192>(g, s);
      case 256: return launch_tc_bn<T, 256>(g, s);
    }
  }
  if (g.NP % 256 == 0) return launch_tc_bn<T, 256>(g, s);
  if (g.NP % 192 == 0) return launch_tc_bn<T, 192>(g, s);
  if (g.NP % 128 == 0) return launch_tc_bn<T, 128>(g, s);
  return launch_tc_bn<T, 64>(g, s);
}

// 3xTF32: tiles of at most 192 columns (two stages of head + tail operands must fit the shared memory)
static int launch_tc_split(const GemmArgs& g, cudaStream_t s) {
  SSR_CHECK(g.NP % 64 == 0, SSR_E_INVALID, "gemm_tc: NP=%d must be a multiple of 64", g.NP);
  if (g.ps_r > 1) {
    const int Cps = g.N / (g.ps_r * g.ps_r);
    SSR_CHECK(Cps % 32 == 0 && g.N == g.NP, SSR_E_INVALID, "gemm_tc: pixel-shuffle store needs N/r^2 %% 32 == 0 (N=%d r=%d)", g.N, g.ps_r);
  }
  if (g.out_ln) {
    SSR_CHECK(g.NP <= 192, SSR_E_INVALID, "gemm_tc (tf32x3): fused LayerNorm needs NP<=192 (NP=%d)", g.NP);
    switch (g.NP) {
      case 64: return launch_tc_bn<float, 64, true>(g, s);
      case 128: return launch_tc_bn<float, 128, true>(g, s);
      case 192: return launch_tc_bn<float, 192, true>(g, s);
    }
  }
  if (g.NP % 192 == 0) return launch_tc_bn<float, 192, true>(g, s);
  if (g.NP % 128 == 0) return launch_tc_bn<float, 128, true>(g, s);
  return launch_tc_bn<float, 64, true>(g, s);
}

int launch_gemm_tc(const GemmArgs& g, int elem, cudaStream_t s) {
  if (elem == 2) return launch_tc_t<__nv_bfloat16>(g, s);
  if (g.split_tf32) return launch_tc_split(g, s);
  return launch_tc_t<float>(g, s);
}

}  // namespace ssr
