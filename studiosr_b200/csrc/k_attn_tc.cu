// HAT attention cores on tcgen05 / TMEM (bf16, head_dim <= 32 padded to 32):
//   * (shifted-)window self-attention over 16x16 windows = 256 tokens (hat.py:84-111 inside HAB :154-195), and
//   * overlapping cross-attention: queries = a 16x16 window, keys / values = the 24x24 window around it cut out by
//     nn.Unfold(kernel 24, stride 16, padding 4) -- out-of-image keys are ZERO vectors that still take part in the softmax
//     with their bias -- with the reference's negative-index wrap of the bias lookup (hat.py:257-284, 490-513).
// Round 1 ran both on mma.sync tiles (k_attn_flash.cu: 126 / 100 TFLOP/s, issue-bound); this kernel is their tcgen05 form.
//
// Work item = (window, head).  torch.roll / window_partition / nn.Unfold / window_reverse are TMA addressing: q, k, v of the
// item arrive as boxes of the pixel-ordered qkv activation [B, H, W, 3 * 192] (SWIZZLE_64B, 64-byte rows = one head), the
// 24x24 key window is ONE box whose out-of-image part is zero-filled by TMA, and the output leaves through the same boxes.
// A shifted 16x16 window is fetched as two 8-pixel-wide halves (each wraps around the image edge on its own; a window that
// wraps in y takes two 8-row boxes per half), which makes the token order inside a window
//   r = (tx / 8) * 128 + ty * 8 + tx % 8
// known only to the bias / mask arithmetic of the epilogue.
//
// Per item the 256 queries are two halves of 128 rows (= TMEM lanes), each with its own accumulators and its own four epilogue
// warps.  The softmax subtracts an UPPER BOUND of the row maximum, m' = max_j(q k_j) log2e + max(bias table of the head)
// (any constant >= the true maximum leaves softmax unchanged; it must only not be so loose that exp2 underflows, and a
// relative-position table spans a few units), so the first pass over the scores is a bare max with no bias lookups:
//   window (256 keys): S = q k^T for all keys stays in TMEM (256 columns); pass 0: row max; pass 1: p = exp2(s log2e + bias
//     (+ mask) - m'), row sum, P packed to bf16 over S; O = P v (tcgen05 TS: A = P from TMEM, B = V MN-major) lands in the
//     columns of S that P no longer needs;
//   24x24 keys (576): three chunks of 192 keys, walked twice -- pass 0: S_c -> row max; pass 1: S_c again (two K = 16 MMAs)
//     -> P_c over S_c -> O += P_c v_c -- so S never needs more than one chunk of columns and nothing is rescaled.
// The bias of 8 consecutive keys is 8 consecutive floats of a per-head table in shared memory (reversed / offset at load time,
// pitch chosen so that a warp's 32 rows hit 32 banks).
// TMEM (512 columns), half x at 256 x: window: S [0,256) -> P [0,128), O [128,160); 24x24: S / P chunk [0,192), O [192,224).
// Warps: 0 = TMA loads + output stores, 1 = MMA issuer, 2..5 = epilogue of query half 0, 6..9 = half 1 (lane quadrant = warp % 4).
#include "ssr_tc.cuh"

namespace ssr {

constexpr int AT_THREADS = 320;
constexpr uint32_t AT_QBYTES = 256 * 64;  // q of one item: 256 tokens x 32 bf16
constexpr uint32_t AT_TAB_BYTES = 8192;

template <bool kOca>
struct AtCfg {
  static constexpr int NK = kOca ? 576 : 256;   // keys per item
  static constexpr int CK = kOca ? 192 : 128;   // keys per chunk
  static constexpr int NC = NK / CK;
  static constexpr uint32_t KV = NK * 64;       // bytes of k (or v) of one item
  static constexpr uint32_t BUF = AT_QBYTES + 2 * KV;
  static constexpr uint32_t OFF_ST = 2 * BUF;                    // output staging [256][64 B] SW64
  static constexpr uint32_t OFF_TAB = OFF_ST + AT_QBYTES;        // bias table of the current head (fp32, x log2e)
  static constexpr uint32_t OFF_BAR = OFF_TAB + AT_TAB_BYTES;
  static constexpr uint32_t SMEM = OFF_BAR + 256 + 1024;
};

enum {
  TB_OPFULL = 0,   // [2] q, k, v of the item in buffer b
  TB_OPEMPTY = 2,  // [2] every MMA of the item has read buffer b
  TB_SFULL = 4,    // [2 halves] S chunk complete
  TB_SREADY = 6,   // [2] pass 0: chunk consumed / pass 1: P chunk written (4 arrivals: the half's epilogue warps)
  TB_OFULL = 8,    // [2] O of the item complete
  TB_OFREE = 10,   // [2] O copied out of TMEM (4 arrivals)
  TB_STFULL = 12,  // output of the item staged (8 arrivals)
  TB_STFREE = 13,  // staging read by the TMA stores
  TB_COUNT = 14
};

struct AttnTcArgs {
  const float* bias;  // [heads][nb * nb] relative-position bias table (fp32, reference layout)
  int B, H, W, heads, shift;
  int nwx, nwy, nwin, n_items;
  int QP;  // channel offset between q, k and v inside a qkv row
};

// K-major operand, SWIZZLE_64B: rows of 64 bytes (32 bf16 of K), 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t at_desc_k_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// MN-major B operand, SWIZZLE_64B: rows of 64 bytes (32 bf16 of N) per K index (as k_swin_attn.cu's V operand)
__device__ __forceinline__ uint64_t at_desc_mn_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t at_sw64(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

template <bool kOca>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmIn2,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2, const AttnTcArgs a) {
  using C = AtCfg<kOca>;
  constexpr int CK = C::CK, NC = C::NC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* s_tab = reinterpret_cast<float*>(smem + C::OFF_TAB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TB_COUNT);
  const uint32_t sbase = smem_u32(smem), bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmIn);
    prefetch_tmap(&tmIn2);
    prefetch_tmap(&tmOut);
    prefetch_tmap(&tmOut2);
    for (int i = 0; i < TB_COUNT; ++i) {
      int cnt = 1;
      if ((i >= TB_SREADY && i < TB_SREADY + 2) || (i >= TB_OFREE && i < TB_OFREE + 2)) cnt = 4;
      if (i == TB_STFULL) cnt = 8;
      mbar_init(bar(i), cnt);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // items in head-major order, so that a CTA re-loads its bias table at most `heads` times
  const int my_items = (a.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto item_of = [&](int it, int& h, int& b, int& wy, int& wx) {
    const int item = (int)blockIdx.x + it * (int)gridDim.x;
    h = item / a.nwin;
    int w = item - h * a.nwin;
    wx = w % a.nwx;
    w /= a.nwx;
    wy = w % a.nwy;
    b = w / a.nwy;
  };

  if (warp == 0) {
    // =========================== TMA loads of q, k, v and stores of the output ===========================
    if (lane == 0) {
      auto load_item = [&](int it) {
        int h, b, wy, wx;
        item_of(it, h, b, wy, wx);
        const int buf = it & 1;
        mbar_wait(bar(TB_OPEMPTY + buf), (((uint32_t)it >> 1) & 1u) ^ 1u);
        const uint32_t fb = bar(TB_OPFULL + buf), dst = sbase + buf * C::BUF;
        mbar_expect_tx(fb, C::BUF);
        if constexpr (kOca) {
          tma_load_4d(dst, &tmIn, fb, h * 32, wx * 16, wy * 16, b);
          tma_load_4d(dst + AT_QBYTES, &tmIn2, fb, a.QP + h * 32, wx * 16 - 4, wy * 16 - 4, b);  // outside the image: zeros
          tma_load_4d(dst + AT_QBYTES + C::KV, &tmIn2, fb, 2 * a.QP + h * 32, wx * 16 - 4, wy * 16 - 4, b);
        } else {
          const int y0 = wy * 16 + a.shift, x0 = wx * 16 + a.shift;
          const bool ywrap = y0 + 16 > a.H;
          for (int op = 0; op < 3; ++op) {
            const int c0 = op * a.QP + h * 32;
            for (int xh = 0; xh < 2; ++xh) {
              const int x = (x0 + 8 * xh) % a.W;
              const uint32_t d = dst + op * AT_QBYTES + xh * 8192;
              if (!ywrap) {
                tma_load_4d(d, &tmIn, fb, c0, x, y0, b);
              } else {  // rows 0..7 at the bottom edge, rows 8..15 wrapped to the top
                tma_load_4d(d, &tmIn2, fb, c0, x, y0, b);
                tma_load_4d(d + 4096, &tmIn2, fb, c0, x, 0, b);
              }
            }
          }
        }
      };
      auto store_item = [&](int it) {
        int h, b, wy, wx;
        item_of(it, h, b, wy, wx);
        const uint32_t src = sbase + C::OFF_ST;
        if constexpr (kOca) {
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmOut), "r"(src),
                       "r"(h * 32), "r"(wx * 16), "r"(wy * 16), "r"(b)
                       : "memory");
        } else {
          const int y0 = wy * 16 + a.shift, x0 = wx * 16 + a.shift;
          const bool ywrap = y0 + 16 > a.H;
          for (int xh = 0; xh < 2; ++xh) {
            const int x = (x0 + 8 * xh) % a.W;
            const uint32_t s = src + xh * 8192;
            if (!ywrap) {
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmOut), "r"(s),
                           "r"(h * 32), "r"(x), "r"(y0), "r"(b)
                           : "memory");
            } else {
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmOut2), "r"(s),
                           "r"(h * 32), "r"(x), "r"(y0), "r"(b)
                           : "memory");
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmOut2),
                           "r"(s + 4096), "r"(h * 32), "r"(x), "r"(0), "r"(b)
                           : "memory");
            }
          }
        }
        bulk_commit();
      };
      if (my_items > 0) load_item(0);
      for (int it = 0; it < my_items; ++it) {
        if (it + 1 < my_items) load_item(it + 1);
        mbar_wait(bar(TB_STFULL), (uint32_t)it & 1u);
        store_item(it);
        bulk_wait_read<0>();
        mbar_arrive(bar(TB_STFREE));
      }
      bulk_wait_all();
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t IDESC_S = umma_idesc(1, 128, CK);
      constexpr uint32_t IDESC_PV = umma_idesc(1, 128, 32) | (1u << 16);  // B operand MN-major
      uint32_t n_ready[2] = {0, 0};
      for (int it = 0; it < my_items; ++it) {
        const int buf = it & 1;
        const uint32_t qa = sbase + buf * C::BUF, ka = qa + AT_QBYTES, va = ka + C::KV;
        mbar_wait(bar(TB_OPFULL + buf), ((uint32_t)it >> 1) & 1u);
        tc_fence_after();
        auto issue_s = [&](int x, int c) {  // S_x = q[x * 128 ...] k_c^T
          const uint64_t ad = at_desc_k_sw64(qa + x * 8192), bd = at_desc_k_sw64(ka + (uint32_t)c * CK * 64);
          const uint32_t t = tmem_base + 256u * (uint32_t)x;
          umma<false>(t, ad, bd, IDESC_S, 0u);
          umma<false>(t, ad + 2, bd + 2, IDESC_S, 1u);
          umma_commit(bar(TB_SFULL + x));
        };
        auto wait_ready = [&](int x) {
          mbar_wait(bar(TB_SREADY + x), n_ready[x] & 1u);
          ++n_ready[x];
          tc_fence_after();
        };
        auto issue_pv = [&](int x, int c) {  // O_x += P_c v_c: A = P (TMEM), B = V rows [c * CK, +CK) MN-major
          const uint32_t t = tmem_base + 256u * (uint32_t)x;
          const uint64_t vd = at_desc_mn_sw64(va + (uint32_t)c * CK * 64);
#pragma unroll
          for (int k = 0; k < CK / 16; ++k) umma_ts(t + 192, t + 8 * k, vd + 64 * k, IDESC_PV, (c | k) ? 1u : 0u);
        };
        if constexpr (!kOca) {
          // window: S for all 256 keys at once, then P.V into the columns of S that the packed P leaves free
          for (int x = 0; x < 2; ++x) {
            if (it > 0) {  // O of the previous item (aliased by this S) has been copied out
              mbar_wait(bar(TB_OFREE + x), (uint32_t)(it - 1) & 1u);
              tc_fence_after();
            }
            const uint64_t ad = at_desc_k_sw64(qa + x * 8192), bd = at_desc_k_sw64(ka);
            const uint32_t t = tmem_base + 256u * (uint32_t)x;
            umma<false>(t, ad, bd, umma_idesc(1, 128, 256), 0u);
            umma<false>(t, ad + 2, bd + 2, umma_idesc(1, 128, 256), 1u);
            umma_commit(bar(TB_SFULL + x));
          }
          for (int x = 0; x < 2; ++x) {
            wait_ready(x);  // P of half x is in TMEM
            const uint32_t t = tmem_base + 256u * (uint32_t)x;
            const uint64_t vd = at_desc_mn_sw64(va);
#pragma unroll
            for (int k = 0; k < 16; ++k) umma_ts(t + 128, t + 8 * k, vd + 64 * k, IDESC_PV, k ? 1u : 0u);
            umma_commit(bar(TB_OFULL + x));
          }
        } else {
        // pass 0: row maxima
        for (int c = 0; c < NC; ++c)
          for (int x = 0; x < 2; ++x) {
            if (c > 0) wait_ready(x);
            issue_s(x, c);
          }
        // pass 1: probabilities and P.V
        for (int x = 0; x < 2; ++x) {
          wait_ready(x);
          issue_s(x, 0);
        }
        for (int c = 0; c < NC; ++c)
          for (int x = 0; x < 2; ++x) {
            wait_ready(x);  // P_c of half x is in TMEM
            if (c == 0 && it > 0) {  // the previous item's O has been copied out
              mbar_wait(bar(TB_OFREE + x), (uint32_t)(it - 1) & 1u);
              tc_fence_after();
            }
            issue_pv(x, c);
            if (c + 1 < NC)
              issue_s(x, c + 1);
            else
              umma_commit(bar(TB_OFULL + x));
          }
        }
        umma_commit(bar(TB_OPEMPTY + buf));
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue: query half x, lane quadrant quad ===========================
    const int x = (warp - 2) >> 2, quad = warp & 3;
    const int row = quad * 32 + lane, R = x * 128 + row;  // query token (smem / TMEM order)
    const uint32_t tS = tmem_base + 256u * (uint32_t)x + ((uint32_t)(quad * 32) << 16);
    int qy, qx;
    if constexpr (kOca) {
      qy = R >> 4;
      qx = R & 15;
    } else {
      qy = (R >> 3) & 15;
      qx = (R >> 7) * 8 + (R & 7);
    }
    constexpr float LOG2E = 1.4426950408889634f, kMask = -100.0f * LOG2E;
    uint32_t n_sfull = 0;
    int cur_head = -1;
    uint8_t* stg = smem + C::OFF_ST;
    for (int it = 0; it < my_items; ++it) {
      int h, b, wy, wx;
      item_of(it, h, b, wy, wx);
      if (h != cur_head) {  // (re)load this head's bias table: every epilogue thread is between items here
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int et = threadIdx.x - 64;
        if constexpr (kOca) {  // T2[i] = T[(i - 880) mod 1521]: index = flat (k - q - 7) offset with the reference's negative wrap
          for (int i = et; i < 1521; i += 256) {
            int src = i - 880;
            if (src < 0) src += 1521;
            s_tab[i] = __ldg(a.bias + (size_t)h * 1521 + src) * LOG2E;
          }
        } else {  // tab[(qy - ky + 15) * 40 + (15 - qx + kx)] = T[(qy - ky + 15) * 31 + (qx - kx + 15)]
          for (int i = et; i < 31 * 31; i += 256) {
            const int ar = i / 31, bc = i - ar * 31;
            s_tab[ar * 40 + (30 - bc)] = __ldg(a.bias + (size_t)h * 961 + i) * LOG2E;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2) {  // max of the (scaled) table -> last float of the table area: the softmax's shift is max(q k) log2e + this
          constexpr int n_tab = kOca ? 1521 : 31 * 31;
          float tm = -INFINITY;
          for (int i = lane; i < n_tab; i += 32) tm = fmaxf(tm, __ldg(a.bias + (size_t)h * n_tab + i) * LOG2E);
          tm = warp_max(tm);
          if (lane == 0) s_tab[AT_TAB_BYTES / 4 - 1] = tm;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_head = h;
      }
      bool xflag = false, yflag = false;
      if constexpr (!kOca) {
        xflag = a.shift > 0 && wx == a.nwx - 1;
        yflag = a.shift > 0 && wy == a.nwy - 1;
      }
      // per 8-key group g of chunk c: first table index and whether the group is masked
      auto group_info = [&](int c, int g, int& tidx, bool& masked) {
        if constexpr (kOca) {
          const int ky = c * 8 + g / 3, kx0 = (g % 3) * 8;
          tidx = (ky - qy - 7) * 39 + (kx0 - qx - 7) + 880;
          masked = false;
        } else {
          const int ky = g, xh = c;  // chunk = x-half of the window, group = key row
          tidx = (qy - ky + 15) * 40 + 15 - qx + 8 * xh;
          masked = (xflag && xh != (qx >> 3)) || (yflag && (ky >> 3) != (qy >> 3));
        }
      };
      auto wait_s = [&]() {
        mbar_wait_warp(bar(TB_SFULL + x), n_sfull & 1u, lane);
        ++n_sfull;
        tc_fence_after();
      };
      constexpr int NP = kOca ? CK / 32 : 8;      // 32-column pieces per pass-1 step
      constexpr int NSTEP = kOca ? NC : 1;        // hand-offs with the MMA issuer per pass
      constexpr int OCOL = kOca ? 192 : 128;      // first TMEM column of O
      // ---------------- pass 0: upper bound of the row maximum (bare scores, four independent chains) ----------------
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      for (int c = 0; c < NSTEP; ++c) {
        wait_s();
#pragma unroll 1
        for (int p = 0; p < NP; p += 2) {
          uint32_t ra[32], rb[32];
          tmem_ld32_nowait(tS + 32 * p, ra);
          tmem_ld32_nowait(tS + 32 * p + 32, rb);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], fmaxf(__uint_as_float(ra[j]), __uint_as_float(rb[j])));
        }
        if constexpr (kOca) {  // the chunk may be overwritten by the next one
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(TB_SREADY + x));
        }
      }
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * LOG2E + s_tab[AT_TAB_BYTES / 4 - 1];  // + max of the bias table
      // ---------------- pass 1: p = exp2(s log2e + bias (+ mask) - m'), packed over S ----------------
      float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int c = 0; c < NSTEP; ++c) {
        if constexpr (kOca) wait_s();
#pragma unroll 1
        for (int p = 0; p < NP; p += 2) {
          // two pieces per iteration: both TMEM loads are issued before the arithmetic of the first
          uint32_t raw2[2][32];
          tmem_ld32_nowait(tS + 32 * p, raw2[0]);
          tmem_ld32_nowait(tS + 32 * p + 32, raw2[1]);
          tmem_wait_ld();
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
          const uint32_t (&raw)[32] = raw2[hp];
          const int pp = p + hp;
          uint32_t pk[16];
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            int tidx;
            bool masked;
            group_info(kOca ? c : (pp >> 2), kOca ? 4 * pp + gq : 4 * (pp & 3) + gq, tidx, masked);
            const float off = (masked ? kMask : 0.0f) - mx;
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              float e0 = fmaf(__uint_as_float(raw[8 * gq + j]), LOG2E, s_tab[tidx + j] + off);
              float e1 = fmaf(__uint_as_float(raw[8 * gq + j + 1]), LOG2E, s_tab[tidx + j + 1] + off);
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(e0));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(e1));
              l4[(j >> 1) & 3] += e0 + e1;
              pk[4 * gq + (j >> 1)] = pack_bf16x2(e0, e1);
            }
          }
          // columns [16pp, 16pp+16) of this lane were read in this or an earlier iteration (16pp + 16 <= 32p + 64): in-place packing is safe
          tmem_st16_u32(tS + 16 * pp, pk);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(TB_SREADY + x));
      }
      const float lsum = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      // ---------------- output: O / l -> bf16 -> staging ----------------
      mbar_wait_warp(bar(TB_OFULL + x), (uint32_t)it & 1u, lane);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld32_nowait(tS + OCOL, raw);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(TB_OFREE + x));
      if (it > 0) mbar_wait_warp(bar(TB_STFREE), (uint32_t)(it - 1) & 1u, lane);  // the previous item's stores have read the staging
      const float inv = 1.0f / lsum;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(raw[8 * j + 0]) * inv, __uint_as_float(raw[8 * j + 1]) * inv);
        v.y = pack_bf16x2(__uint_as_float(raw[8 * j + 2]) * inv, __uint_as_float(raw[8 * j + 3]) * inv);
        v.z = pack_bf16x2(__uint_as_float(raw[8 * j + 4]) * inv, __uint_as_float(raw[8 * j + 5]) * inv);
        v.w = pack_bf16x2(__uint_as_float(raw[8 * j + 6]) * inv, __uint_as_float(raw[8 * j + 7]) * inv);
        *reinterpret_cast<uint4*>(stg + at_sw64(R, j)) = v;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(TB_STFULL));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// a: the AttnArgs of the mma.sync kernel (qkv bf16 [B*H*W][ld_qkv], q pre-scaled by d^-1/2; o bf16 [B*H*W][ld_o]); oca = 0 / 1
int launch_attn_tc(const AttnArgs& a, int oca, cudaStream_t s) {
  SSR_CHECK(a.ws == 16 && a.DP == 32 && (!oca || a.kws == 24), SSR_E_INVALID, "attn_tc: only 16x16 windows (24x24 keys), head dim <= 32");
  SSR_CHECK(a.H % 16 == 0 && a.W % 16 == 0, SSR_E_INVALID, "attn_tc: %dx%d not a multiple of the 16x16 window", a.H, a.W);
  SSR_CHECK(oca || a.shift == 0 || a.shift == 8, SSR_E_INVALID, "attn_tc: shift %d not in {0, 8}", a.shift);
  SSR_CHECK(a.ld_qkv % 8 == 0 && a.ld_o % 8 == 0 && a.QP % 8 == 0, SSR_E_INVALID, "attn_tc: unaligned leading dims");
  CUtensorMap tmIn, tmIn2, tmOut, tmOut2;
  auto map4d = [&](CUtensorMap* m, const void* base, int ld, int bw, int bh) {
    cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)a.W * ld * 2, (cuuint64_t)a.H * a.W * ld * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    return make_tmap(m, base, 2, 4, dims, str, box, 64);
  };
  if (oca) {
    SSR_TRY(map4d(&tmIn, a.qkv, a.ld_qkv, 16, 16));
    SSR_TRY(map4d(&tmIn2, a.qkv, a.ld_qkv, 24, 24));
    SSR_TRY(map4d(&tmOut, a.o, a.ld_o, 16, 16));
    tmOut2 = tmOut;
  } else {
    SSR_TRY(map4d(&tmIn, a.qkv, a.ld_qkv, 8, 16));
    SSR_TRY(map4d(&tmIn2, a.qkv, a.ld_qkv, 8, 8));
    SSR_TRY(map4d(&tmOut, a.o, a.ld_o, 8, 16));
    SSR_TRY(map4d(&tmOut2, a.o, a.ld_o, 8, 8));
  }
  AttnTcArgs k;
  k.bias = a.bias;
  k.B = a.B; k.H = a.H; k.W = a.W; k.heads = a.heads; k.shift = oca ? 0 : a.shift;
  k.nwx = a.W / 16; k.nwy = a.H / 16;
  k.nwin = a.B * k.nwx * k.nwy;
  k.n_items = k.nwin * a.heads;
  k.QP = a.QP;
  const int sms = num_sms_cached();
  const int grid = k.n_items < sms ? k.n_items : sms;
  const double Nq = 256, Nk = oca ? 576 : 256;
  ProfScope prof(oca ? "attn_oca_tc" : "attn_win_tc", 4.0 * k.nwin * Nq * Nk * a.d * a.heads, (double)k.nwin * (2.0 * Nq + 2.0 * Nk) * a.heads * a.d * 2, s);
  if (oca) {
    static bool attr = false;
    if (!attr) {
      SSR_CUDA(cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AtCfg<true>::SMEM));
      attr = true;
    }
    attn_tc_kernel<true><<<grid, AT_THREADS, AtCfg<true>::SMEM, s>>>(tmIn, tmIn2, tmOut, tmOut2, k);
  } else {
    static bool attr = false;
    if (!attr) {
      SSR_CUDA(cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AtCfg<false>::SMEM));
      attr = true;
    }
    attn_tc_kernel<false><<<grid, AT_THREADS, AtCfg<false>::SMEM, s>>>(tmIn, tmIn2, tmOut, tmOut2, k);
  }
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
