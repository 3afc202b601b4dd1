// Fused (shifted-)window attention of a Swin block (bf16, 8x8 windows, head_dim <= 32 padded to 32, C padded to 192):
//
//   qkv = xn @ Wqkv^T + b                  (swinir.py:80)            [tcgen05 SS, N = 192 per head pair]
//   s   = q k^T * d^-1/2 + rel_pos_bias (+ shift mask)  (swinir.py:83-95)  [tcgen05 SS, scores in TMEM]
//   p   = softmax(s)                       (swinir.py:97)            [registers; P written back over S as bf16]
//   o   = p v                              (swinir.py:100)           [tcgen05 TS: A = P from TMEM, B = V MN-major]
//
// torch.roll, window_partition and window_reverse (swinir.py:154-168, common.py:236-247) are TMA tile addressing:
// one work item is two windows = 128 tokens; every window is fetched as two [8 rows x 4 px x 64 ch] boxes (four when it
// wraps around the bottom edge of the cyclically shifted image) straight out of the pixel-ordered NHWC activation, and
// the result leaves through the same boxes.  That makes the token order inside a window
//   r = (tx / 4) * 32 + ty * 4 + tx % 4
// which only the (host-permuted) bias table and the mask have to know about.  The shift mask is -100 on whole 16-column
// groups of that order (the 4x4 quadrants of the window), so it costs one predicated add per group.
//
// Work per item is software-pipelined over the three head pairs g: while the epilogue warps run the softmax of pair g
// the tensor core already computes the QKV projection of pair g+1.  q is pre-scaled by d^-1/2 * log2(e) and the bias
// table by log2(e) at pack time, so the softmax is a bare ex2.
//
// TMEM (512 columns): [0,192) QKV accumulator of the current pair (q h0 h1 | k h0 h1 | v h0 h1, 32 each);
// [192,320) / [320,448) scores of head 0 / 1 of the pair over all 128 keys of the item (only the own window's 64 are
// read), overwritten in place by P (packed bf16: 64 columns, the other window's half zeroed = block-diagonal P);
// [448,512) O of the pair.
// Warps: 0 = output stores, 1 = weight producer, 2 = activation loads, 3 = MMA issuer, 4..11 = epilogue
// (lane quadrant = warp % 4; group = (warp - 4) / 4: column half in the QKV epilogue, head of the pair afterwards).
#include "ssr_tc.cuh"

namespace ssr {

constexpr int SA_THREADS = 384;
constexpr int SA_POLY_DEFAULT = 4;  // every 4th pair of exponentials runs on the FMA pipe (0 = all on MUFU.EX2)
constexpr uint32_t SA_TILE = 16384;
constexpr uint32_t SA_WSLOT = 192 * 128;
constexpr int SA_WSLOTS = 3;  // one per k-block of a pair's projection: the next pair's weights are complete before its MMAs start
constexpr uint32_t SA_OFF_XN = 0;                                  // 3 k-block tiles of the LayerNorm-ed input
constexpr uint32_t SA_OFF_W = SA_OFF_XN + 3 * SA_TILE;             // weight ring
constexpr uint32_t SA_OFF_Q = SA_OFF_W + SA_WSLOTS * SA_WSLOT;     // [128][64] bf16 SW128: q of the pair
constexpr uint32_t SA_OFF_K = SA_OFF_Q + SA_TILE;                  // [128][64] bf16 SW128: k of the pair
constexpr uint32_t SA_OFF_V = SA_OFF_K + SA_TILE;                  // 2 buffers x 2 heads x [128 tokens][32] bf16 SW64 (MN-major B operand)
constexpr uint32_t SA_OFF_OST = SA_OFF_V + 2 * SA_TILE;            // 2 x [128][64] bf16 SW128 output staging (pair g -> buffer g & 1)
// relative-position bias, compact: per head 2 copies (one per 4-byte alignment of a window row's start) of the reversed
// 15 x 15 table, row pitch 16 bf16: the 64 values of a token row are 16 aligned runs of 4 (one per key quad)
constexpr uint32_t SA_BT_PITCH = 16, SA_BT_COPY = 15 * SA_BT_PITCH * 2, SA_BT_HEAD = 2 * SA_BT_COPY;
constexpr uint32_t SA_BIAS_BYTES = 6 * SA_BT_HEAD;
constexpr uint32_t SA_OFF_BIAS = SA_OFF_OST + 2 * SA_TILE;
constexpr uint32_t SA_OFF_BAR = SA_OFF_BIAS + SA_BIAS_BYTES;
constexpr uint32_t SA_SMEM = SA_OFF_BAR + 256;  // the kernel has no static shared memory: the dynamic base is 1024-aligned (checked)
static_assert(SA_SMEM <= 232448, "fused attention kernel exceeds the 227 KB shared-memory limit");

enum {
  AB_WFULL = 0,                      // [2]
  AB_WEMPTY = AB_WFULL + SA_WSLOTS,  // [2]
  AB_XNFULL = AB_WEMPTY + SA_WSLOTS,
  AB_XNEMPTY,
  AB_QKVFULL,   // accumulator of pair g complete
  AB_OPREADY,   // q/k/v operands of pair g in smem, accumulator drained (8 arrivals)
  AB_SFULL,     // [2] scores of head 0/1
  AB_PREADY = AB_SFULL + 2,  // [2] P of head 0/1 in TMEM (4 arrivals)
  AB_OFULL = AB_PREADY + 2,  // [2] O of head 0/1
  AB_OSTAGED = AB_OFULL + 2,    // [2] output of pair g staged in buffer g & 1 (8 arrivals)
  AB_OSTFREE = AB_OSTAGED + 2,  // [2] staging buffer drained by its TMA stores
  AB_END = AB_OSTFREE + 2,
  AB_COUNT = AB_END
};

struct AttnKArgs {
  const uint4* bias_tab;  // [6 heads][2 copies][15][16] bf16 (pack_attn_fused_host), times log2(e)
  int B, H, W, shift;
  int nwx, nwy, n_windows, n_tiles;
  long long* dbg;  // optional phase timestamps (developer diagnostics): [CTA][g < 96][16]: 0..5 epilogue warp 4, 8..13 MMA issuer, 14..15 store warp
};

__device__ __forceinline__ uint32_t sa_sw128(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t sa_sw64(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

// MN-major B operand, SWIZZLE_64B: rows of 64 bytes (32 bf16 of N) per K index, 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_mn_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}

struct WinPos {
  int valid, b, x0, y0, ywrap;
};
__device__ __forceinline__ WinPos win_pos(const AttnKArgs& a, int widx) {
  WinPos p;
  p.valid = widx < a.n_windows;
  const int wx = widx % a.nwx, t = widx / a.nwx;
  const int wy = t % a.nwy;
  p.b = p.valid ? t / a.nwy : a.B;  // an out-of-range image index makes TMA zero-fill (loads) / drop (stores) the box
  p.x0 = wx * 8 + a.shift;          // < W: shift < 8 and the last window starts at W - 8 ... the halves wrap separately
  p.y0 = wy * 8 + a.shift;
  p.ywrap = p.y0 + 8 > a.H;
  return p;
}

// CL = CTAs per cluster: with CL = 2 each CTA fetches half of every Wqkv k-block and TMA-multicasts it into the ring slot of
// both (the 221 KB of weights per 128-token item are 2/3 of the kernel's L2 reads); both CTAs walk the same item count.
template <int SA_POLY_EVERY, int CL>
__global__ void __launch_bounds__(SA_THREADS, 1)
swin_attn_kernel(const __grid_constant__ CUtensorMap tmX8, const __grid_constant__ CUtensorMap tmX4,
                 const __grid_constant__ CUtensorMap tmO8, const __grid_constant__ CUtensorMap tmO4,
                 const __grid_constant__ CUtensorMap tmW, const AttnKArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();  // SWIZZLE_128B operand tiles need the 1 KB alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SA_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AB_COUNT);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (int)(SA_BIAS_BYTES / 16); i += SA_THREADS)
    reinterpret_cast<uint4*>(smem + SA_OFF_BIAS)[i] = __ldg(a.bias_tab + i);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX8);
    prefetch_tmap(&tmX4);
    prefetch_tmap(&tmO8);
    prefetch_tmap(&tmO4);
    prefetch_tmap(&tmW);
    for (int i = 0; i < AB_COUNT; ++i) {
      int cnt = 1;
      if (i == AB_OPREADY || i == AB_OSTAGED || i == AB_OSTAGED + 1) cnt = 8;
      if (i == AB_PREADY || i == AB_PREADY + 1) cnt = 4;
      if (i >= AB_WEMPTY && i < AB_WEMPTY + SA_WSLOTS) cnt = CL;  // released by the MMA issuer of every CTA of the cluster
      mbar_init(bar(i), cnt);
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
  // items of this CTA; in a cluster every CTA takes the same number (item indices >= n_tiles are dummies: zero-filled loads,
  // no stores)
  const int my_tiles = CL > 1 ? (a.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x
                              : (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kClMask = (uint16_t)((1u << CL) - 1u);
  const int G = 3 * my_tiles;                                                               // head-pair steps
  constexpr uint32_t IDESC_QKV = umma_idesc(1, 128, 192), IDESC_S = umma_idesc(1, 128, 128);
  constexpr uint32_t IDESC_PV = umma_idesc(1, 128, 32) | (1u << 16);  // B operand MN-major
  const uint32_t tACC = tmem_base, tS[2] = {tmem_base + 192, tmem_base + 320}, tO = tmem_base + 448;

  if (warp == 0) {
    // =========================== output stores: lane = (window, x half) of the pair's staging tile ===========================
    // (one thread needs ~500 cycles to issue a TMA store: four lanes issue their boxes with the same instruction)
    for (int g = 0; g < G; ++g) {
      const int it = g / 3, hp = g - 3 * it;
      mbar_wait_warp(bar(AB_OSTAGED + (g & 1)), ((uint32_t)g >> 1) & 1u, lane);
      long long* dbg = (a.dbg && lane == 0 && g < 96) ? a.dbg + 16 * ((size_t)blockIdx.x * 96 + g) : nullptr;
      if (dbg) dbg[14] = clock64();
      if (lane < 4) {
        const int w = lane >> 1, hh = lane & 1;
        const int tile = blockIdx.x + it * gridDim.x;
        const WinPos p = win_pos(a, 2 * tile + w);
        if (p.valid) {
          const int x = (p.x0 + 4 * hh) % a.W;
          const uint32_t src = sbase + SA_OFF_OST + (uint32_t)(g & 1) * SA_TILE + (uint32_t)(64 * w + 32 * hh) * 128u;
          if (!p.ywrap) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmO8), "r"(src),
                         "r"(hp * 64), "r"(x), "r"(p.y0), "r"(p.b)
                         : "memory");
          } else {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmO4), "r"(src),
                         "r"(hp * 64), "r"(x), "r"(p.y0), "r"(p.b)
                         : "memory");
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmO4),
                         "r"(src + 16 * 128), "r"(hp * 64), "r"(x), "r"(0), "r"(p.b)
                         : "memory");
          }
        }
        bulk_commit();
        bulk_wait_read<0>();  // ~3.8k cycles after the staging was published; the other buffer takes the next pair meanwhile
      }
      __syncwarp();
      if (dbg) dbg[15] = clock64();
      if (lane == 0) mbar_arrive(bar(AB_OSTFREE + (g & 1)));
    }
    if (lane < 4) bulk_wait_all();
  } else if (warp == 1) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      for (int e = 0; e < 3 * G; ++e) {
        const int g = e / 3, kb = e - 3 * g, hp = g % 3;
        const int s = e % SA_WSLOTS;
        mbar_wait(bar(AB_WEMPTY + s), (((uint32_t)e / SA_WSLOTS) & 1u) ^ 1u);
        mbar_expect_tx(bar(AB_WFULL + s), SA_WSLOT);
        if (CL > 1)  // own rows of the k-block -> both CTAs; the slot completes with the peer's part
          tma_load_2d_mc(sbase + SA_OFF_W + s * SA_WSLOT + crank * (192u / CL) * 128u, &tmW, bar(AB_WFULL + s), kb * 64,
                         hp * 192 + (int)crank * (192 / CL), kClMask);
        else
          tma_load_2d(sbase + SA_OFF_W + s * SA_WSLOT, &tmW, bar(AB_WFULL + s), kb * 64, hp * 192);
      }
    }
  } else if (warp == 2) {
    // =========================== activation loads: lane = (window, x half) of the item ===========================
    // The next item's input is requested the moment the last projection of this item has read the tile (XNEMPTY); issued
    // from the store thread it arrived ~1.8k cycles after the first projection of the item wanted it.
    for (int it = 0; it < my_tiles; ++it) {
      if (it > 0) mbar_wait_warp(bar(AB_XNEMPTY), (uint32_t)(it - 1) & 1u, lane);
      if (lane == 0) mbar_expect_tx(bar(AB_XNFULL), 3 * SA_TILE);
      __syncwarp();
      if (lane < 4) {
        const int w = lane >> 1, hh = lane & 1;
        const int tile = blockIdx.x + it * gridDim.x;
        const WinPos p = win_pos(a, 2 * tile + w);
        const int x = (p.x0 + 4 * hh) % a.W;
        const uint32_t row_off = (uint32_t)(64 * w + 32 * hh) * 128u;
        for (int kb = 0; kb < 3; ++kb) {
          const uint32_t dst = sbase + SA_OFF_XN + kb * SA_TILE + row_off;
          if (!p.ywrap) {
            tma_load_4d(dst, &tmX8, bar(AB_XNFULL), kb * 64, x, p.y0, p.b);
          } else {  // rows ty 0..3 at the bottom edge, ty 4..7 wrapped to the top
            tma_load_4d(dst, &tmX4, bar(AB_XNFULL), kb * 64, x, p.y0, p.b);
            tma_load_4d(dst + 16 * 128, &tmX4, bar(AB_XNFULL), kb * 64, x, 0, p.b);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // =========================== MMA issuer ===========================
    if (lane == 0 && G > 0) {
      uint32_t wk = 0;
      auto proj = [&](int g) {  // QKV accumulator of pair g: [128 x 192] = xn [128 x 192] * Whp[g % 3]^T
        for (int kb = 0; kb < 3; ++kb) {
          const int s = wk % SA_WSLOTS;
          mbar_wait(bar(AB_WFULL + s), (wk / SA_WSLOTS) & 1u);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(sbase + SA_OFF_XN + kb * SA_TILE);
          const uint64_t bdesc = umma_desc_sw128(sbase + SA_OFF_W + s * SA_WSLOT);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<false>(tACC, adesc + 2 * k, bdesc + 2 * k, IDESC_QKV, (kb | k) ? 1u : 0u);
          if (CL > 1)
            umma_commit_mc(bar(AB_WEMPTY + s), kClMask);
          else
            umma_commit(bar(AB_WEMPTY + s));
          ++wk;
        }
        umma_commit(bar(AB_QKVFULL));
        if (g % 3 == 2) umma_commit(bar(AB_XNEMPTY));  // last reader of this item's input tile
      };
      auto scores = [&](int h) {  // S_h = q_h k_h^T over all 128 keys of the item (K = 32: bytes [64h, 64h+64) of a row)
        const uint64_t qd = umma_desc_sw128(sbase + SA_OFF_Q), kd = umma_desc_sw128(sbase + SA_OFF_K);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma<false>(tS[h], qd + 4 * h + 2 * k, kd + 4 * h + 2 * k, IDESC_S, k ? 1u : 0u);
        umma_commit(bar(AB_SFULL + h));
      };
      auto proj_next = [&](int g) {
        if (g >= G) return;
        if (g % 3 == 0) {
          mbar_wait(bar(AB_XNFULL), (uint32_t)(g / 3) & 1u);
          tc_fence_after();
        }
        proj(g);
      };
      // Issue order per pair: scores(g), projection(g+1), P.V(g).  (Tried: P.V_h(g) -> scores_h(g+1) interleaved with the
      // projection last, so that the next softmax starts earlier: pair period 4.9k -> 5.2k cycles, the projection then
      // completes after the epilogue warps come back for it.)
      proj_next(0);
      for (int g = 0; g < G; ++g) {
        const uint32_t ph = (uint32_t)g & 1u;
        long long* dbg = (a.dbg && g < 96) ? a.dbg + 16 * ((size_t)blockIdx.x * 96 + g) : nullptr;
        if (dbg) dbg[8] = clock64();
        mbar_wait(bar(AB_OPREADY), ph);  // q, k, v of pair g staged; accumulator drained
        tc_fence_after();
        if (dbg) dbg[9] = clock64();
        scores(0);
        scores(1);
        if (dbg) dbg[10] = clock64();
        proj_next(g + 1);  // runs under this pair's softmax
        if (dbg) dbg[11] = clock64();
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // O_h = P_h v_h: A = P (TMEM, block-diagonal over the two windows), B = V_h MN-major
          mbar_wait(bar(AB_PREADY + h), ph);
          tc_fence_after();
          if (dbg) dbg[12 + h] = clock64();
          const uint64_t vd = umma_desc_mn_sw64(sbase + SA_OFF_V + (uint32_t)(g & 1) * SA_TILE + h * 8192);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 keys per step: 8 packed columns of P, 16 rows (1 KB) of V
            umma_ts(tO + 32 * h, tS[h] + 8 * k, vd + 64 * k, IDESC_PV, k ? 1u : 0u);
          umma_commit(bar(AB_OFULL + h));
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue (8 warps) ===========================
    const int ew = warp - 4;
    const int grp = ew >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // token row of the item
    const int w = row >> 6;            // window of the item (warp-uniform)
    const int ri = row & 63;           // token index inside the window (pi order)
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint8_t* sQ = smem + SA_OFF_Q;
    uint8_t* sK = smem + SA_OFF_K;
    uint8_t* sV = smem + SA_OFF_V;
    uint8_t* sO = smem + SA_OFF_OST;
    // this token's window coordinates (order r = (tx / 4) * 32 + ty * 4 + tx % 4) and its copy of the bias table
    const int yi = (ri & 31) >> 2, xi = (ri >> 5) * 4 + (ri & 3);
    const int bt_s = (7 - xi) & 1;
    const uint8_t* sBiasRow = smem + SA_OFF_BIAS + bt_s * SA_BT_COPY + (7 - yi) * (SA_BT_PITCH * 2) + (7 - xi + bt_s) * 2;
    constexpr float kMask = -100.0f * 1.4426950408889634f;
    bool yflag = false, xflag = false;

    // output of head `grp` of pair gg: O / l -> bf16 -> staging (runs one pair late, under the next pair's MMAs)
    auto out_epilogue = [&](int gg, float inv_l, long long* dbg) {
      mbar_wait_warp(bar(AB_OFULL + grp), (uint32_t)gg & 1u, lane);
      tc_fence_after();
      if (dbg) dbg[6] = clock64();
      uint32_t raw[32];
      tmem_ld32_nowait(tlane + (tO - tmem_base) + 32 * grp, raw);
      if (gg > 1) mbar_wait_warp(bar(AB_OSTFREE + (gg & 1)), (((uint32_t)gg >> 1) & 1u) ^ 1u, lane);  // stores of pair gg-2 have read this buffer
      tmem_wait_ld();
      if (dbg) dbg[7] = clock64();
      const f32x2 il = f2_splat(inv_l);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        v.x = f2_to_bf16x2(f2_mul(f2_pack_u(raw[8 * j + 0], raw[8 * j + 1]), il));
        v.y = f2_to_bf16x2(f2_mul(f2_pack_u(raw[8 * j + 2], raw[8 * j + 3]), il));
        v.z = f2_to_bf16x2(f2_mul(f2_pack_u(raw[8 * j + 4], raw[8 * j + 5]), il));
        v.w = f2_to_bf16x2(f2_mul(f2_pack_u(raw[8 * j + 6], raw[8 * j + 7]), il));
        *reinterpret_cast<uint4*>(sO + (gg & 1) * SA_TILE + sa_sw128(row, 4 * grp + j)) = v;
      }
    };  // the caller publishes the staging (fence.proxy.async + arrive on AB_OSTAGED) together with its own smem writes
    // QKV epilogue of one 32-column chunk: bf16 into the operand tile `dst_of(chunk j)` selects.  The qkv bias is already in
    // the accumulator: norm1's output carries 1.0 in the two pad channels C, C+1 and the packed weight holds the bias there
    // (hi + lo bf16 parts), so there is no shared-memory bias fetch in front of every store.
    auto qkv_chunk = [&](const uint32_t (&raw)[32], auto dst_of) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // 8 columns -> one 16-byte chunk
        const uint32_t* r8 = &raw[8 * j];
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(r8[0]), __uint_as_float(r8[1]));
        v.y = pack_bf16x2(__uint_as_float(r8[2]), __uint_as_float(r8[3]));
        v.z = pack_bf16x2(__uint_as_float(r8[4]), __uint_as_float(r8[5]));
        v.w = pack_bf16x2(__uint_as_float(r8[6]), __uint_as_float(r8[7]));
        *reinterpret_cast<uint4*>(dst_of(j)) = v;
      }
    };

    float inv_l_prev = 1.0f;
    for (int g = 0; g < G; ++g) {
      const uint32_t ph = (uint32_t)g & 1u;
      const int it = g / 3, hp = g - 3 * it;
      if (hp == 0 && a.shift > 0) {  // which edges of the shifted image does this thread's window touch?
        const int widx = 2 * ((int)blockIdx.x + it * (int)gridDim.x) + w;
        const int wx = widx % a.nwx, wy = (widx / a.nwx) % a.nwy;
        yflag = wy == a.nwy - 1;
        xflag = wx == a.nwx - 1;
      }
      long long* dbg = (a.dbg && ew == 0 && lane == 0 && g < 96) ? a.dbg + 16 * ((size_t)blockIdx.x * 96 + g) : nullptr;
      if (dbg) dbg[0] = clock64();

      // ---------------- QKV epilogue: bf16 operand tiles (V into buffer g & 1: P.V of pair g-1 may still run) ----------------
      mbar_wait_warp(bar(AB_QKVFULL), ph, lane);
      if (dbg) dbg[1] = clock64();
      tc_fence_after();
      {
        uint32_t raw[3][32];
#pragma unroll
        for (int c = 0; c < 3; ++c) tmem_ld32_nowait(tlane + grp * 96 + c * 32, raw[c]);
        tmem_wait_ld();
        uint8_t* vbuf = sV + (g & 1) * SA_TILE;
        if (grp == 0) {  // columns [0,96): q h0 | q h1 | k h0
          qkv_chunk(raw[0], [&](int j) { return sQ + sa_sw128(row, j); });
          qkv_chunk(raw[1], [&](int j) { return sQ + sa_sw128(row, 4 + j); });
          qkv_chunk(raw[2], [&](int j) { return sK + sa_sw128(row, j); });
        } else {  // columns [96,192): k h1 | v h0 | v h1
          qkv_chunk(raw[0], [&](int j) { return sK + sa_sw128(row, 4 + j); });
          qkv_chunk(raw[1], [&](int j) { return vbuf + sa_sw64(row, j); });
          qkv_chunk(raw[2], [&](int j) { return vbuf + 8192 + sa_sw64(row, j); });
        }
      }
      // publish the operand tiles first: the score MMAs (and the next pair's projection) then run under the output epilogue
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(AB_OPREADY));
      if (dbg) dbg[2] = clock64();
      // ---------------- output of the previous pair (its P.V ran under this pair's QKV epilogue) ----------------
      if (g > 0) {
        out_epilogue(g - 1, inv_l_prev, dbg);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(AB_OSTAGED + ((g - 1) & 1)));
      }
      if (dbg) dbg[3] = clock64();

      // ---------------- softmax of head `grp` of the pair ----------------
      const int head = 2 * hp + grp;
      // bias row of this token: 64 bf16 = 8 chunks (fetched while the score MMAs finish)
      uint4 bq[8];
      {
        // keys 8c .. 8c+7 = key rows yj = 2c % 8, 2c % 8 + 1 at key columns 4 (c / 4) .. +3: two aligned runs of the table
        const uint8_t* brow = sBiasRow + head * SA_BT_HEAD;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint8_t* p0 = brow + ((2 * c) & 7) * (SA_BT_PITCH * 2) + (c >> 2) * 8;
          const uint32_t* q0 = reinterpret_cast<const uint32_t*>(p0);
          const uint32_t* q1 = reinterpret_cast<const uint32_t*>(p0 + SA_BT_PITCH * 2);
          bq[c] = make_uint4(q0[0], q0[1], q1[0], q1[1]);
        }
      }
      mbar_wait_warp(bar(AB_SFULL + grp), ph, lane);
      if (dbg) dbg[4] = clock64();
      tc_fence_after();
      {
        uint32_t raw[2][32];
        tmem_ld32_nowait(tlane + (tS[grp] - tmem_base) + 64 * w, raw[0]);
        tmem_ld32_nowait(tlane + (tS[grp] - tmem_base) + 64 * w + 32, raw[1]);
        tmem_wait_ld();
        f32x2 s[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t wds[4] = {bq[c].x, bq[c].y, bq[c].z, bq[c].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int p = 4 * c + i;  // pair index: key columns 2p, 2p+1
            const f32x2 bias2 = f2_pack_u(wds[i] << 16, wds[i] & 0xffff0000u);
            s[p] = f2_add(f2_pack_u(raw[p >> 4][(2 * p) & 31], raw[p >> 4][(2 * p + 1) & 31]), bias2);
          }
        }
        if (yflag || xflag) {  // shift mask: -100 on the 16-key groups whose quadrant lies across the image seam
          const int gi = ri >> 4;
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const bool m = (xflag && ((gq >> 1) != (gi >> 1))) || (yflag && ((gq & 1) != (gi & 1)));
            const f32x2 mv = f2_splat(m ? kMask : 0.0f);
#pragma unroll
            for (int i = 0; i < 8; ++i) s[8 * gq + i] = f2_add(s[8 * gq + i], mv);
          }
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int p = 0; p < 32; ++p) {
          float lo, hi;
          f2_unpack(s[p], lo, hi);
          mx[p & 3] = fmaxf(mx[p & 3], fmaxf(lo, hi));
        }
        const f32x2 nm = f2_splat(-fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
        f32x2 acc2[4] = {0ull, 0ull, 0ull, 0ull};
        uint32_t pk[32];
#pragma unroll
        for (int p = 0; p < 32; ++p) {
          f32x2 e2;
          if (SA_POLY_EVERY > 0 && p % (SA_POLY_EVERY > 0 ? SA_POLY_EVERY : 1) == SA_POLY_EVERY - 1) {
            e2 = f2_exp2_poly(f2_add(s[p], nm));  // every SA_POLY_EVERY-th pair: FMA pipe instead of MUFU (the phase is MUFU-bound)
          } else {
            float lo, hi;
            f2_unpack(f2_add(s[p], nm), lo, hi);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(lo) : "f"(lo));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(hi) : "f"(hi));
            e2 = f2_pack(lo, hi);
          }
          acc2[p & 3] = f2_add(acc2[p & 3], e2);
          pk[p] = f2_to_bf16x2(e2);
        }
        inv_l_prev = 1.0f / f2_hsum(f2_add(f2_add(acc2[0], acc2[1]), f2_add(acc2[2], acc2[3])));
        // P over the scores: own window's 64 keys -> 32 packed columns, the other window's half zeroed
        uint32_t zero[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) zero[i] = 0u;
        tmem_st32_u32(tlane + (tS[grp] - tmem_base) + 32 * w, pk);
        tmem_st32_u32(tlane + (tS[grp] - tmem_base) + 32 * (1 - w), zero);
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(AB_PREADY + grp));
      if (dbg) dbg[5] = clock64();
    }
    if (G > 0) {
      out_epilogue(G - 1, inv_l_prev, nullptr);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(AB_OSTAGED + ((G - 1) & 1)));
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

extern long long* g_tail_dbg;

int launch_swin_attn_fused(const AttnFusedArgs& f, cudaStream_t s) {
  SSR_CHECK(f.H % 8 == 0 && f.W % 8 == 0, SSR_E_INVALID, "swin_attn: %dx%d not a multiple of the 8x8 window", f.H, f.W);
  SSR_CHECK(f.shift == 0 || f.shift == 4, SSR_E_INVALID, "swin_attn: shift %d not in {0, 4}", f.shift);
  SSR_CHECK(f.ld_x == 192 && f.ld_o == 192, SSR_E_INVALID, "swin_attn: leading dims must be 192 (got %d / %d)", f.ld_x, f.ld_o);
  // developer A/B switches: STUDIOSR_B200_ATTN_POLY=0 keeps every softmax exponential on MUFU, STUDIOSR_B200_ATTN_CLUSTER=2
  // switches the cluster-of-2 weight multicast on (measured neutral: profiles/r02_cluster_multicast.txt)
  static int poly = -1, cl = 1;
  if (poly < 0) {
    const char* e = getenv("STUDIOSR_B200_ATTN_POLY");
    poly = (e && e[0] == '0') ? 0 : SA_POLY_DEFAULT;
    const char* c = getenv("STUDIOSR_B200_ATTN_CLUSTER");
    cl = (c && c[0] == '2') ? 2 : 1;
    SSR_CUDA(cudaFuncSetAttribute(swin_attn_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_SMEM));
    SSR_CUDA(cudaFuncSetAttribute(swin_attn_kernel<SA_POLY_DEFAULT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_SMEM));
    SSR_CUDA(cudaFuncSetAttribute(swin_attn_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_SMEM));
    SSR_CUDA(cudaFuncSetAttribute(swin_attn_kernel<SA_POLY_DEFAULT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_SMEM));
  }
  CUtensorMap tmX8, tmX4, tmO8, tmO4, tmW;
  auto map4d = [&](CUtensorMap* m, const void* base, int ld, int bh) {
    cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)f.W, (cuuint64_t)f.H, (cuuint64_t)f.B};
    cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)f.W * ld * 2, (cuuint64_t)f.H * f.W * ld * 2};
    cuuint32_t box[4] = {64, 4, (cuuint32_t)bh, 1};
    return make_tmap(m, base, 2, 4, dims, str, box, 128);
  };
  SSR_TRY(map4d(&tmX8, f.xn, f.ld_x, 8));
  SSR_TRY(map4d(&tmX4, f.xn, f.ld_x, 4));
  SSR_TRY(map4d(&tmO8, f.o, f.ld_o, 8));
  SSR_TRY(map4d(&tmO4, f.o, f.ld_o, 4));
  {
    cuuint64_t dims[2] = {192, 576};
    cuuint64_t str[1] = {192 * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(192 / cl)};
    SSR_TRY(make_tmap(&tmW, f.Whp, 2, 2, dims, str, box, 128));
  }
  AttnKArgs a;
  a.bias_tab = reinterpret_cast<const uint4*>(f.bias_tab);
  a.B = f.B; a.H = f.H; a.W = f.W; a.shift = f.shift;
  a.nwx = f.W / 8; a.nwy = f.H / 8;
  a.n_windows = f.B * a.nwx * a.nwy;
  a.n_tiles = (a.n_windows + 1) / 2;
  a.dbg = g_tail_dbg;
  const int sms = num_sms_cached();
  const double T = (double)f.B * f.H * f.W;
  const double flops = T * (2.0 * 3 * f.C * f.C + 4.0 * 64 * f.C);  // qkv projection + (q k^T, p v) over 64 keys
  const double bytes = T * f.C * (2 + 2);
  ProfScope prof("swin_attn", flops, bytes, s);
  int grid = a.n_tiles < sms ? a.n_tiles : sms;
  if (cl == 1) {
    if (poly == 0)
      swin_attn_kernel<0, 1><<<grid, SA_THREADS, SA_SMEM, s>>>(tmX8, tmX4, tmO8, tmO4, tmW, a);
    else
      swin_attn_kernel<SA_POLY_DEFAULT, 1><<<grid, SA_THREADS, SA_SMEM, s>>>(tmX8, tmX4, tmO8, tmO4, tmW, a);
  } else {
    grid = (grid + 1) / 2 * 2;
    if (grid > sms) grid -= 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(SA_THREADS);
    cfg.dynamicSmemBytes = SA_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (poly == 0)
      SSR_CUDA(cudaLaunchKernelEx(&cfg, swin_attn_kernel<0, 2>, tmX8, tmX4, tmO8, tmO4, tmW, a));
    else
      SSR_CUDA(cudaLaunchKernelEx(&cfg, swin_attn_kernel<SA_POLY_DEFAULT, 2>, tmX8, tmX4, tmO8, tmO4, tmW, a));
  }
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
