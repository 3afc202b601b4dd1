// Weight-gradient kernel of the training path on tcgen05 / TMEM / TMA, bf16 operands, fp32 accumulate.
//
//   dW[n][tap][c] += alpha * sum_p dY[p][n] * X[p + off(tap)][c]          (conv3x3, zero pad 1; taps = 1: nn.Linear)
//
// which is what autograd computes for nn.Conv2d / nn.Linear weights (the reference relies on torch autograd:
// trainer.py:104 `loss.backward()`).  The contraction runs over PIXELS, so both operands are "MN-major" for the
// tensor core: a TMA box of 64 pixels x 64 channels lands in shared memory as 64 rows (K = pixel) of 128 bytes
// (64 channels of M or N), 128B-swizzled -- exactly the canonical MN-major SWIZZLE_128B atom
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: LBO = distance between 64-channel blocks, SBO = 1024 B (8 pixels).
// The tap shift of the 3x3 conv is the box origin of the X load; TMA's out-of-bounds zero fill is the conv padding
// and also masks ragged image edges (dY reads 0 there).
//
// Work item = (128 output channels) x (one tap) x (BN <= 256 input channels) x (one split of the pixel range);
// splits accumulate into the packed fp32 gradient with red.global.add.v4.f32.
// Warp roles (224 threads): warp 0 / warp 6 = TMA producers (even / odd ring entries), warp 1 = TMEM + MMA issuer,
// warps 2..5 = epilogue (thread <-> output-channel row).
#include "ssr_tc.cuh"

namespace ssr {

constexpr int WG_THREADS = 224;
constexpr int WG_STAGES = 4;
constexpr uint32_t WG_A_BYTES = 2 * 8192, WG_B_BYTES = 4 * 8192, WG_STAGE = WG_A_BYTES + WG_B_BYTES;
constexpr size_t WG_SMEM = (size_t)WG_STAGES * WG_STAGE + 256 + 1024;

struct WgGeom {
  int conv;                // 1: 4-D (C,W,H,B) boxes of bw x bh x bb = 64 pixels; 0: 64 consecutive rows
  int bw, bh, bb;
  int cx, cy;              // chunk grid per image (conv)
  int n_chunks;            // pixel chunks in total
  int m_blocks, n_blocks;  // 128-row blocks of dW, BN-column blocks
  int BN;
  int splits;
};

// MN-major operand, SWIZZLE_128B: rows of 128 bytes (64 bf16 of M / N) per K index; LBO = byte distance between
// 64-element blocks along M / N, SBO = 1024 B between 8-row K groups
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const WgradArgs a,
                const WgGeom geo) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem), bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (WG_STAGES + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * WG_STAGES);

  // bias gradient: the CTAs of the centre tap (tap 0 of a Linear) and the first column block also sum the columns of the dY
  // tiles; their four epilogue warps, idle during the main loop, read each stage and co-sign its release
  const bool do_bias = a.dBp != nullptr && (blockIdx.x / geo.splits) % geo.n_blocks == 0 &&
                       ((blockIdx.x / geo.splits) / geo.n_blocks) % a.taps == (a.taps == 9 ? 4 : 0);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), do_bias ? 5 : 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // decode the work item: split fastest (CTAs of one tile run together and share X / dY through L2)
  int item = blockIdx.x;
  const int split = item % geo.splits;
  item /= geo.splits;
  const int nb = item % geo.n_blocks;
  item /= geo.n_blocks;
  const int tap = item % a.taps;
  const int mb = item / a.taps;
  const int c_begin = (int)((long long)geo.n_chunks * split / geo.splits);
  const int c_end = (int)((long long)geo.n_chunks * (split + 1) / geo.splits);
  const int n_iter = c_end - c_begin;
  const int m0 = mb * 128, n0 = nb * geo.BN;
  const int m1 = min(m0 + 64, a.NoutP - 64);  // second 64-row block (re-reads the first when only 64 rows are left)
  const int dy = a.taps == 9 ? tap / 3 - 1 : 0, dx = a.taps == 9 ? tap % 3 - 1 : 0;
  const int nbox = geo.BN / 64;
  const uint32_t stage_tx = WG_A_BYTES + (uint32_t)nbox * 8192u;

  if (warp == 0 || warp == 6) {
    if (lane == 0) {
      const int pid = warp == 0 ? 0 : 1;
      for (int it = pid; it < n_iter; it += 2) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (uint32_t)(it / WG_STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), stage_tx);
        const uint32_t dA = sbase + s * WG_STAGE, dB = dA + WG_A_BYTES;
        const int ch = c_begin + it;
        if (geo.conv) {
          int t = ch;
          const int x0 = (t % geo.cx) * geo.bw;
          t /= geo.cx;
          const int y0 = (t % geo.cy) * geo.bh, b0 = (t / geo.cy) * geo.bb;
          tma_load_4d(dA, &tmY, full_bar(s), m0, x0, y0, b0);
          tma_load_4d(dA + 8192, &tmY, full_bar(s), m1, x0, y0, b0);
          for (int j = 0; j < nbox; ++j) tma_load_4d(dB + j * 8192, &tmX, full_bar(s), n0 + 64 * j, x0 + dx, y0 + dy, b0);
        } else {
          const int r0 = ch * 64;
          tma_load_2d(dA, &tmY, full_bar(s), m0, r0);
          tma_load_2d(dA + 8192, &tmY, full_bar(s), m1, r0);
          for (int j = 0; j < nbox; ++j) tma_load_2d(dB + j * 8192, &tmX, full_bar(s), n0 + 64 * j, r0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: bf16 x bf16 -> fp32, A and B both MN-major (bits 15 / 16)
      const uint32_t idesc = umma_idesc(1, 128, geo.BN) | (1u << 15) | (1u << 16);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % WG_STAGES;
        mbar_wait(full_bar(s), (uint32_t)(it / WG_STAGES) & 1u);
        tc_fence_after();
        const uint64_t ad = umma_desc_mn_sw128(sbase + s * WG_STAGE, 8192);
        const uint64_t bd = umma_desc_mn_sw128(sbase + s * WG_STAGE + WG_A_BYTES, 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 16 pixels (= 16 rows of 128 B) per MMA
          umma<false>(tmem_base, ad + 128 * k, bd + 128 * k, idesc, (it | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(s));
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else if (warp < 6) {
    const int quad = warp & 3;
    const int row = m0 + quad * 32 + lane;
    if (do_bias) {
      // 128 threads = 16 column groups (8 channels = one 16-byte chunk) x 8 row lanes; a thread adds rows rl, rl + 8, ... of its
      // chunk with LDS.128 (the 8 row lanes of a quarter-warp hit 8 different swizzled chunks: conflict-free) into 8 fp32
      // accumulators; the row lanes meet by shuffles after the last stage.  te = thread index among the epilogue warps.
      const int te = (warp - 2) * 32 + lane;
      const int cg = te >> 3, rl = te & 7;  // column group 0..15 (block = cg >> 3), row lane
      const uint8_t* gbase = smem + (cg >> 3) * 8192;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % WG_STAGES;
        mbar_wait_warp(full_bar(s), (uint32_t)(it / WG_STAGES) & 1u, lane);
        const uint8_t* tile = gbase + s * WG_STAGE;
#pragma unroll
        for (int r8 = 0; r8 < 8; ++r8) {
          const int r = r8 * 8 + rl;  // r & 7 == rl
          const uint4 u = *reinterpret_cast<const uint4*>(tile + r * 128 + ((((cg & 7)) ^ rl) << 4));
          acc[0] += __uint_as_float(u.x << 16); acc[1] += __uint_as_float(u.x & 0xffff0000u);
          acc[2] += __uint_as_float(u.y << 16); acc[3] += __uint_as_float(u.y & 0xffff0000u);
          acc[4] += __uint_as_float(u.z << 16); acc[5] += __uint_as_float(u.z & 0xffff0000u);
          acc[6] += __uint_as_float(u.w << 16); acc[7] += __uint_as_float(u.w & 0xffff0000u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 2);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 4);
      }
      if (rl == 0 && n_iter > 0) {
        const int colb = cg * 8;  // first of this thread's 8 channels within the 128-row block
        const bool blk_valid = colb < 64 || m1 == m0 + 64;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rrow = m0 + colb + i;
          if (blk_valid && rrow < a.NoutP) atomicAdd(a.dBp + rrow, acc[i] * a.alpha);
        }
      }
    }
    mbar_wait_warp(done_bar, 0, lane);
    tc_fence_after();
    if (n_iter > 0) {
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
      // rows of the duplicated second block (m1 != m0 + 64) are not written
      const bool valid = row < a.NoutP && (quad < 2 || m1 == m0 + 64);
      float* dst = a.dWp + ((size_t)row * a.taps + tap) * a.CinP + n0;
#pragma unroll 1
      for (int c = 0; c < geo.BN / 32; ++c) {
        float v[32];
        tmem_ld32(trow + c * 32, v);
        if (valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            red_add_v4(dst + c * 32 + 4 * q, v[4 * q] * a.alpha, v[4 * q + 1] * a.alpha, v[4 * q + 2] * a.alpha,
                       v[4 * q + 3] * a.alpha);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

static void choose_chunk(int B, int H, int W, int* bw, int* bh, int* bb) {
  long long best = -1;
  for (int w = 1; w <= 64; w *= 2)
    for (int h = 1; w * h <= 64; h *= 2) {
      const int b = 64 / (w * h);
      if (w > 2 * W || h > 2 * H) continue;
      const long long cost = (long long)((W + w - 1) / w) * ((H + h - 1) / h) * ((B + b - 1) / b);
      if (best < 0 || cost < best || (cost == best && w > *bw)) {
        best = cost;
        *bw = w;
        *bh = h;
        *bb = b;
      }
    }
}

int launch_wgrad_tc(const WgradArgs& a, cudaStream_t s) {
  SSR_CHECK(a.NoutP % 64 == 0 && a.CinP % 64 == 0 && a.NoutP >= 64 && a.CinP >= 64, SSR_E_INVALID,
            "wgrad_tc: NoutP=%d CinP=%d must be multiples of 64", a.NoutP, a.CinP);
  SSR_CHECK(a.ldy % 8 == 0 && a.ldx % 8 == 0 && a.ldy >= a.NoutP && a.ldx >= a.CinP, SSR_E_INVALID, "wgrad_tc: ldy=%d ldx=%d",
            a.ldy, a.ldx);
  SSR_CHECK(a.taps == 1 || a.taps == 9, SSR_E_INVALID, "wgrad_tc: taps=%d", a.taps);
  WgGeom geo{};
  geo.BN = a.CinP <= 256 ? a.CinP : (a.CinP % 256 == 0 ? 256 : a.CinP % 192 == 0 ? 192 : a.CinP % 128 == 0 ? 128 : 64);
  geo.n_blocks = a.CinP / geo.BN;
  geo.m_blocks = (a.NoutP + 127) / 128;
  CUtensorMap tmY, tmX;
  if (a.taps == 9) {
    geo.conv = 1;
    choose_chunk(a.B, a.H, a.W, &geo.bw, &geo.bh, &geo.bb);
    geo.cx = (a.W + geo.bw - 1) / geo.bw;
    geo.cy = (a.H + geo.bh - 1) / geo.bh;
    geo.n_chunks = geo.cx * geo.cy * ((a.B + geo.bb - 1) / geo.bb);
    cuuint32_t box[4] = {64, (cuuint32_t)geo.bw, (cuuint32_t)geo.bh, (cuuint32_t)geo.bb};
    {
      cuuint64_t dims[4] = {(cuuint64_t)a.NoutP, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
      cuuint64_t str[3] = {(cuuint64_t)a.ldy * 2, (cuuint64_t)a.W * a.ldy * 2, (cuuint64_t)a.H * a.W * a.ldy * 2};
      SSR_TRY(make_tmap(&tmY, a.dY, 2, 4, dims, str, box));
    }
    {
      cuuint64_t dims[4] = {(cuuint64_t)a.CinP, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
      cuuint64_t str[3] = {(cuuint64_t)a.ldx * 2, (cuuint64_t)a.W * a.ldx * 2, (cuuint64_t)a.H * a.W * a.ldx * 2};
      SSR_TRY(make_tmap(&tmX, a.X, 2, 4, dims, str, box));
    }
  } else {
    geo.conv = 0;
    geo.n_chunks = (a.M + 63) / 64;
    cuuint32_t box[2] = {64, 64};
    {
      cuuint64_t dims[2] = {(cuuint64_t)a.NoutP, (cuuint64_t)a.M};
      cuuint64_t str[1] = {(cuuint64_t)a.ldy * 2};
      SSR_TRY(make_tmap(&tmY, a.dY, 2, 2, dims, str, box));
    }
    {
      cuuint64_t dims[2] = {(cuuint64_t)a.CinP, (cuuint64_t)a.M};
      cuuint64_t str[1] = {(cuuint64_t)a.ldx * 2};
      SSR_TRY(make_tmap(&tmX, a.X, 2, 2, dims, str, box));
    }
  }
  const int tiles = geo.m_blocks * a.taps * geo.n_blocks;
  // one CTA per SM is resident (192 KB of operand stages): aim at exactly two full waves, never a third partial one
  const int target = 2 * num_sms_cached();
  geo.splits = target / tiles;
  if (geo.splits > geo.n_chunks) geo.splits = geo.n_chunks;
  if (geo.splits < 1) geo.splits = 1;
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
    attr_set = true;
  }
  const double flops = 2.0 * a.M * (double)a.N_alg * a.K_alg * a.taps;
  const double bytes = (double)a.M * (a.N_alg + a.K_alg) * 2 + (double)a.N_alg * a.K_alg * a.taps * 4;
  ProfScope prof(a.taps == 9 ? "wgrad_tc_conv3x3" : "wgrad_tc_linear", flops, bytes, s);
  wgrad_tc_kernel<<<tiles * geo.splits, WG_THREADS, WG_SMEM, s>>>(tmY, tmX, a, geo);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
