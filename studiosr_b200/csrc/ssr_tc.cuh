// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and the warp-transposing epilogue I/O helpers shared by
// the tensor-core kernels (k_gemm_tc.cu, k_mlp_fused.cu).  sm_100a only.
#pragma once
#include <cuda.h>

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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (prompt wake-up) or the
// hint expires, instead of spinning.  A spinning waiter is not free: ncu showed the old loop (try_wait + clock64 +
// branch) taking 45 % of all issued warp instructions of the fused kernels and, through CS2R, half of the XU pipe
// that MUFU.TANH needs.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap (-> CUDA error on the host), never hang the GPU box.  The clock is only
// consulted once every 256 failed (i.e. timed-out) attempts.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 8000000000LL) __trap();
    }
  }
}
// Polling wait (mbarrier.test_wait in a tight loop): for single-thread roles (TMA producer, MMA issuer) whose wake-up latency
// is on the critical path of a software pipeline; a suspended try_wait wakes up late.  Bounded like mbar_wait.
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 0xfffffu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 8000000000LL) __trap();
    }
  }
}
// whole-warp wait: one lane polls, the warp re-converges behind it (32x less barrier traffic, one wake-up)
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
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
// ---- thread-block clusters: weight tiles are fetched from L2 once per cluster and multicast into every CTA's ring slot ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the box lands at the same shared-memory offset, and completes bytes on the mbarrier at the same offset, in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// TMA store of a staged smem box (bulk async-group completion) and L2 prefetch of a box
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() {  // smem of all but the newest kPending store groups may be re-used
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

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
// arrives on the mbarrier at this offset in every CTA of `mask` once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
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

// A operand from TMEM (row = lane, K packed two bf16 per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// issue only (no wait): several loads can be in flight before one tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_st32_u32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}

// single-column variants: one 32-bit value per lane (cross-warp exchange of per-row scalars through TMEM)
__device__ __forceinline__ void tmem_ld1_nowait(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st1_u32(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
// 16-column variants (32 lanes x 16 columns)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16_u32(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
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

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2): two lanes per issued instruction.  The fused
// epilogues are issue-bound, not FMA-pipe bound, so this halves their cost. ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 f2_pack_u(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 a, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
}
__device__ __forceinline__ void f2_unpack_u(f32x2 a, uint32_t& lo, uint32_t& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(a));
}
__device__ __forceinline__ f32x2 f2_splat(float x) { return f2_pack(x, x); }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(f32x2 a) {
  float lo, hi;
  f2_unpack(a, lo, hi);
  return pack_bf16x2(lo, hi);
}
// 2^x for a pair of non-positive x on the FMA pipe instead of the 16-lanes-per-clock MUFU: round-to-nearest split x = j + f
// by the 1.5 * 2^23 magic add, degree-3 minimax of 2^f on [-0.5, 0.5] (relative error 7.5e-5, far below the bf16 rounding
// of the probabilities it feeds), j added into the exponent field.  x is clamped at -126 (result 2^-126 instead of 0).
__device__ __forceinline__ f32x2 f2_exp2_poly(f32x2 x) {
  float lo, hi;
  f2_unpack(x, lo, hi);
  x = f2_pack(fmaxf(lo, -126.0f), fmaxf(hi, -126.0f));
  const f32x2 magic = f2_splat(12582912.0f);
  const f32x2 t = f2_add(x, magic);
  const f32x2 f = f2_fma(f2_add(t, f2_splat(-12582912.0f)), f2_splat(-1.0f), x);
  f32x2 p = f2_fma(f2_splat(0.0551716648042202f), f, f2_splat(0.2426111251115799f));
  p = f2_fma(p, f, f2_splat(0.6932609677314758f));
  p = f2_fma(p, f, f2_splat(0.9999280571937561f));
  uint32_t pl, ph, tl, th;
  f2_unpack_u(p, pl, ph);
  f2_unpack_u(t, tl, th);
  return f2_pack_u(pl + (tl << 23), ph + (th << 23));
}
__device__ __forceinline__ float f2_hsum(f32x2 a) {
  float lo, hi;
  f2_unpack(a, lo, hi);
  return lo + hi;
}

constexpr int TC_STAGE_ROW = 36;                     // floats per staging row (32 + 4 pad: conflict-free 16 B access)
constexpr int TC_STAGE_BYTES = 32 * TC_STAGE_ROW * 4;  // per-warp transpose staging tile

// ---- per-warp transposing I/O: thread <-> row in registers, lanes <-> columns in global memory ----
// The TMEM accumulator layout gives every thread one full row; writing rows straight from registers
// makes each 16-byte store its own L2 request (measured: ~6 cycles/request, 36k cycles per tile).  These
// helpers bounce a 32x32 chunk through a padded shared tile so global accesses are 128 B (fp32) or
// 64 B (bf16) contiguous per row, 4 / 8 rows per instruction.  Row indices of the transposed view are
// gathered once per work item (RowMap) so the per-chunk code is shuffle- and branch-free.
struct RowMap {
  int mf[8];  // fp32 view: row 4*i + (lane>>3) -> pixel index or -1
  int mh[4];  // bf16 view: row 8*i + (lane>>2) -> pixel index or -1
};
__device__ __forceinline__ RowMap make_rowmap(int m_own, int lane) {
  RowMap r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.mf[i] = __shfl_sync(0xffffffffu, m_own, 4 * i + (lane >> 3));
#pragma unroll
  for (int i = 0; i < 4; ++i) r.mh[i] = __shfl_sync(0xffffffffu, m_own, 8 * i + (lane >> 2));
  return r;
}
__device__ __forceinline__ void stage_put_f32(float* st, int lane, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(st + lane * TC_STAGE_ROW + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void stage_get_f32(const float* st, int lane, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 x = *reinterpret_cast<const float4*>(st + lane * TC_STAGE_ROW + 4 * q);
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
}
// coalesced residual prefetch: 8 x float4 per lane covering the warp's 32 rows x 32 columns at column nb
__device__ __forceinline__ void res_prefetch(float4 (&x)[8], const RowMap& rm, int lane, const float* res, int ldres, int nb) {
  const int c4 = (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    x[i] = rm.mf[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(res + (size_t)rm.mf[i] * ldres + nb + c4))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
}
// transposed domain: st += residual (if any), optional coalesced fp32 store of the sum
__device__ __forceinline__ void stage_add_store_f32(float* st, int lane, const RowMap& rm, const float4* resv, float* out,
                                                    int ld_out, int nb) {
  const int c4 = (lane & 7) * 4;
  float4 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(st + (4 * i + (lane >> 3)) * TC_STAGE_ROW + c4);
  if (resv) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i].x += resv[i].x; x[i].y += resv[i].y; x[i].z += resv[i].z; x[i].w += resv[i].w;
      *reinterpret_cast<float4*>(st + (4 * i + (lane >> 3)) * TC_STAGE_ROW + c4) = x[i];
    }
  }
  if (out) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (rm.mf[i] >= 0) *reinterpret_cast<float4*>(out + (size_t)rm.mf[i] * ld_out + nb + c4) = x[i];
  }
}
// rows -> global bf16: staging rows hold 32 bf16 (64 B) at an 80-byte pitch.  `dst_of(i)` = destination of row-view i.
template <typename DstF>
__device__ __forceinline__ void stage_store_bf16(float* stf, int lane, const float (&v)[32], DstF dst_of) {
  uint8_t* st = reinterpret_cast<uint8_t*>(stf);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(st + lane * 80 + 16 * q) =
        make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                   pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
  __syncwarp();
  uint4 x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const uint4*>(st + (8 * i + (lane >> 2)) * 80 + (lane & 3) * 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat16* d = dst_of(i);
    if (d) *reinterpret_cast<uint4*>(d + (lane & 3) * 8) = x[i];
  }
  __syncwarp();
}
template <typename DstF>
__device__ __forceinline__ void stage_store_tf32(float* st, int lane, float (&v)[32], DstF dst_of, bool round = true) {
  if (round) {  // single-pass tf32 consumers read exact operands; the 3xTF32 mode keeps full fp32 activations
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = round_tf32(v[i]);
  }
  stage_put_f32(st, lane, v);
  __syncwarp();
  float4 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(st + (4 * i + (lane >> 3)) * TC_STAGE_ROW + (lane & 7) * 4);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* d = dst_of(i);
    if (d) *reinterpret_cast<float4*>(d + (lane & 7) * 4) = x[i];
  }
  __syncwarp();
}

// erf with |error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26) on the SFU: rcp.approx + ex2.approx + 7 FMA-class
// ops instead of erff's branchy ~40-instruction polynomial (measured: GELU epilogue 19k -> 6k cycles / tile).
// Used by the tensor-core epilogues only; the fp32 CUDA-core path keeps erff.
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  return copysignf(fmaf(-p * t, e, 1.0f), x);
}
__device__ __forceinline__ void epilogue_act(float (&v)[32], int act, float slope, float alpha) {
  // one uniform branch around each 32-element loop: a per-element switch gets if-converted by ptxas and
  // then evaluates the erf polynomial for every element whatever `act` is (measured 2.5k cycles / chunk)
  if (act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.5f * v[i] * (1.0f + fast_erf(v[i] * 0.70710678118654752440f));
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (act == ACT_LEAKY) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.0f ? v[i] : v[i] * slope;
  }
  if (alpha != 1.0f) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= alpha;
  }
}

// host helper (k_gemm_tc.cu): 128B-swizzled tiled tensor map over a row-major tensor
int make_tmap(CUtensorMap* map, const void* base, int elem, int rank, const cuuint64_t* dims,
              const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes = 128);
int num_sms_cached();

}  // namespace ssr
