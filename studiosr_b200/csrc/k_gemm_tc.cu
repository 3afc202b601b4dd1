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
#include <cuda.h>

#include <mutex>

#include "ssr_device.cuh"

namespace ssr {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap (-> CUDA error on the host), never hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr (kTf32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
}

// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (base_lane + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
      "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
      "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
      "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
      "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major operand, 128B swizzle, rows of 128 bytes packed densely:
// start>>4 | LBO(=1, unused for swizzled K-major) | SBO = 1024 B (8 rows) | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B format (1 = bf16, 2 = tf32), both K-major, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
struct TcGeom {
  int conv;            // 0: 128 consecutive rows per tile; 1: BW x BH x BB pixel patch per tile
  int BW, BH, BB;      // patch shape (BW*BH*BB == 128)
  int tiles_x, tiles_y;  // patch grid per image (conv)
  int kc_per_tap;      // KP / BK
  int nkb;             // taps * kc_per_tap
};

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 192;

template <int BLOCK_N>
constexpr int tc_tmem_cols() {
  return BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
}
template <int BLOCK_N>
constexpr int tc_stages() {
  // stage = 16 KB (A) + BLOCK_N*128 B (W); keep a CTA <= ~100 KB so two fit per SM
  return BLOCK_N >= 192 ? 2 : (BLOCK_N >= 128 ? 3 : 4);
}
template <int BLOCK_N>
constexpr size_t tc_smem_bytes() {
  return (size_t)tc_stages<BLOCK_N>() * (16384 + BLOCK_N * 128) + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <typename T, int BLOCK_N>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmArgs g,
               const TcGeom geo) {
  constexpr bool kTf32 = sizeof(T) == 4;
  constexpr int BK = 128 / (int)sizeof(T);  // elements per 128-byte swizzle row
  constexpr int kStages = tc_stages<BLOCK_N>();
  constexpr int kTmemCols = tc_tmem_cols<BLOCK_N>();
  constexpr uint32_t A_BYTES = TC_BM * 128, W_BYTES = BLOCK_N * 128;
  constexpr uint32_t idesc = umma_idesc(kTf32 ? 2 : 1, TC_BM, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smA = smem;
  uint8_t* smW = smem + kStages * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (A_BYTES + W_BYTES));
  // bars[0..kStages) full, [kStages..2kStages) empty, [2kStages] tmem_full; then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * kStages);

  // ---- tile coordinates ----
  const int n0 = blockIdx.y * BLOCK_N;
  int m0 = 0, tx0 = 0, ty0 = 0, tb0 = 0;
  if (geo.conv) {
    int t = blockIdx.x;
    tx0 = (t % geo.tiles_x) * geo.BW;
    t /= geo.tiles_x;
    ty0 = (t % geo.tiles_y) * geo.BH;
    tb0 = (t / geo.tiles_y) * geo.BB;
  } else {
    m0 = blockIdx.x * TC_BM;
  }

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      for (int kb = 0; kb < geo.nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), A_BYTES + W_BYTES);
        const int tap = kb / geo.kc_per_tap, kc = kb - tap * geo.kc_per_tap;
        const uint32_t dstA = smem_u32(smA + s * A_BYTES), dstW = smem_u32(smW + s * W_BYTES);
        if (geo.conv) {
          const int dy = (g.taps == 9) ? tap / 3 - 1 : 0, dx = (g.taps == 9) ? tap % 3 - 1 : 0;
          tma_load_4d(dstA, &tmA, full_bar(s), kc * BK, tx0 + dx, ty0 + dy, tb0);
        } else {
          tma_load_2d(dstA, &tmA, full_bar(s), kc * BK, m0);
        }
        tma_load_2d(dstW, &tmW, full_bar(s), kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      for (int kb = 0; kb < geo.nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(smA + s * A_BYTES));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smW + s * W_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 4 x 32 bytes of K per 128-byte row: +2 in the (addr >> 4) field
          umma<kTf32>(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    int m;
    bool valid;
    int pb = 0, py = 0, px = 0;
    if (geo.conv) {
      const int ww = row % geo.BW, hh = (row / geo.BW) % geo.BH, bb = row / (geo.BW * geo.BH);
      pb = tb0 + bb;
      py = ty0 + hh;
      px = tx0 + ww;
      valid = pb < g.B && py < g.H && px < g.W;
      m = (pb * g.H + py) * g.W + px;
    } else {
      m = m0 + row;
      valid = m < g.M;
      if (g.ps_r > 1 && valid) {
        px = m % g.W;
        py = (m / g.W) % g.H;
        pb = m / (g.W * g.H);
      }
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int Cps = g.ps_r > 1 ? g.N / (g.ps_r * g.ps_r) : 0;
    const bool do_ln = g.out_ln != nullptr;
    float sum = 0.0f;
    T* outT = reinterpret_cast<T*>(g.out_T);

#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      float v[32];
      tmem_ld32(trow + c * 32, v);
      const int nb = n0 + c * 32;
      if (valid) {
        const float4* b4 = reinterpret_cast<const float4*>(g.bias + nb);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bv = __ldg(b4 + q);
          v[4 * q + 0] = apply_act(v[4 * q + 0] + bv.x, g.act, g.slope) * g.alpha;
          v[4 * q + 1] = apply_act(v[4 * q + 1] + bv.y, g.act, g.slope) * g.alpha;
          v[4 * q + 2] = apply_act(v[4 * q + 2] + bv.z, g.act, g.slope) * g.alpha;
          v[4 * q + 3] = apply_act(v[4 * q + 3] + bv.w, g.act, g.slope) * g.alpha;
        }
        if (g.res) {
          const float4* r4 = reinterpret_cast<const float4*>(g.res + (size_t)m * g.ldres + nb);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 rv = r4[q];
            v[4 * q + 0] += rv.x;
            v[4 * q + 1] += rv.y;
            v[4 * q + 2] += rv.z;
            v[4 * q + 3] += rv.w;
          }
        }
        if (nb + 32 > g.N) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (nb + i >= g.N) v[i] = 0.0f;
        }
        if (g.out_f32) {
          float4* o4 = reinterpret_cast<float4*>(g.out_f32 + (size_t)m * g.ld_f32 + nb);
#pragma unroll
          for (int q = 0; q < 8; ++q) o4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        if (outT) {
          T* dst;
          bool ok = true;
          if (g.ps_r > 1) {
            // 32 consecutive GEMM columns stay inside one (i,j) sub-pixel because Cps % 32 == 0
            ok = nb < g.N;
            dst = outT + ps_offset(pb, py, px, nb, g.H, g.W, g.ps_r, Cps, g.ld_T);
          } else {
            dst = outT + (size_t)m * g.ld_T + nb;
          }
          if (ok) {
            if constexpr (kTf32) {
              float4* o4 = reinterpret_cast<float4*>(dst);
#pragma unroll
              for (int q = 0; q < 8; ++q)
                o4[q] = make_float4(round_tf32(v[4 * q]), round_tf32(v[4 * q + 1]), round_tf32(v[4 * q + 2]),
                                    round_tf32(v[4 * q + 3]));
            } else {
              uint4* o4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                o4[q] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                   pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
            }
          }
        }
        if (do_ln) {
#pragma unroll
          for (int i = 0; i < 32; ++i) sum += v[i];
        }
      }
      if (do_ln) tmem_st32(trow + c * 32, v);  // keep v for the two LayerNorm passes
    }

    if (do_ln) {  // host guarantees n0 == 0 and BLOCK_N == NP: the thread owns the whole row
      const float mean = sum / (float)g.N;
      float sq = 0.0f;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        float v[32];
        tmem_ld32(trow + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = v[i] - mean;
          if (c * 32 + i < g.N) sq += d * d;
        }
      }
      const float rstd = rsqrtf(sq / (float)g.N + g.eps);
      T* oln = reinterpret_cast<T*>(g.out_ln);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        float v[32];
        tmem_ld32(trow + c * 32, v);
        if (valid) {
          const float4* g4 = reinterpret_cast<const float4*>(g.gamma + c * 32);
          const float4* be4 = reinterpret_cast<const float4*>(g.beta + c * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 gv = __ldg(g4 + q), bv = __ldg(be4 + q);
            v[4 * q + 0] = (v[4 * q + 0] - mean) * rstd * gv.x + bv.x;
            v[4 * q + 1] = (v[4 * q + 1] - mean) * rstd * gv.y + bv.y;
            v[4 * q + 2] = (v[4 * q + 2] - mean) * rstd * gv.z + bv.z;
            v[4 * q + 3] = (v[4 * q + 3] - mean) * rstd * gv.w + bv.w;
          }
          T* dst = oln + (size_t)m * g.ld_ln + c * 32;
          if constexpr (kTf32) {
            float4* o4 = reinterpret_cast<float4*>(dst);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              o4[q] = make_float4(round_tf32(v[4 * q]), round_tf32(v[4 * q + 1]), round_tf32(v[4 * q + 2]),
                                  round_tf32(v[4 * q + 3]));
          } else {
            uint4* o4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              o4[q] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                 pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
          }
        }
      }
    }
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

static int make_tmap(CUtensorMap* map, const void* base, int elem, int rank, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  SSR_CHECK(fn != nullptr, SSR_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapDataType dt = elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSR_CHECK(r == CUDA_SUCCESS, SSR_E_CUDA, "cuTensorMapEncodeTiled failed (%d), rank %d base %p dims %llu,%llu", (int)r,
            rank, base, (unsigned long long)dims[0], (unsigned long long)dims[1]);
  return SSR_OK;
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

template <typename T, int BLOCK_N>
static int launch_tc_bn(const GemmArgs& g, cudaStream_t s) {
  constexpr int elem = (int)sizeof(T);
  constexpr int BK = 128 / elem;
  SSR_CHECK(g.KP % BK == 0 && g.lda % (16 / elem) == 0, SSR_E_INVALID, "gemm_tc: KP=%d lda=%d", g.KP, g.lda);
  TcGeom geo{};
  geo.kc_per_tap = g.KP / BK;
  geo.nkb = g.taps * geo.kc_per_tap;
  CUtensorMap tmA, tmW;
  int grid_x;
  if (g.taps == 9) {
    geo.conv = 1;
    choose_patch(g.B, g.H, g.W, &geo.BW, &geo.BH, &geo.BB);
    geo.tiles_x = (g.W + geo.BW - 1) / geo.BW;
    geo.tiles_y = (g.H + geo.BH - 1) / geo.BH;
    grid_x = geo.tiles_x * geo.tiles_y * ((g.B + geo.BB - 1) / geo.BB);
    cuuint64_t dims[4] = {(cuuint64_t)g.lda, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
    cuuint64_t str[3] = {(cuuint64_t)g.lda * elem, (cuuint64_t)g.W * g.lda * elem, (cuuint64_t)g.H * g.W * g.lda * elem};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)geo.BW, (cuuint32_t)geo.BH, (cuuint32_t)geo.BB};
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
    cuuint64_t dims[2] = {Ktot, (cuuint64_t)g.NP};
    cuuint64_t str[1] = {Ktot * elem};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BLOCK_N};
    SSR_TRY(make_tmap(&tmW, g.Wt, elem, 2, dims, str, box));
  }
  constexpr size_t smem = tc_smem_bytes<BLOCK_N>();
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<T, BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(grid_x, g.NP / BLOCK_N);
  ProfScope prof(g.taps == 9 ? "gemm_tc_conv3x3" : "gemm_tc_linear", gemm_alg_flops(g), gemm_alg_bytes(g, elem), s);
  gemm_tc_kernel<T, BLOCK_N><<<grid, TC_THREADS, smem, s>>>(tmA, tmW, g, geo);
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

int launch_gemm_tc(const GemmArgs& g, int elem, cudaStream_t s) {
  if (elem == 2) return launch_tc_t<__nv_bfloat16>(g, s);
  return launch_tc_t<float>(g, s);
}

}  // namespace ssr
