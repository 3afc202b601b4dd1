// Developer microbenchmark (GPU box): tcgen05.mma issue rate per N with SMEM operands, with and
// without concurrent LDS/STS traffic from other warps, and tcgen05.ld throughput for 4/8/16 warps.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I studiosr_b200/csrc scripts/micro_tc.cu -o scripts/micro_tc -lcuda
#include <cstdio>
#include <cstdlib>

#include "ssr_tc.cuh"

using namespace ssr;

namespace ssr {
void set_error(const char*, ...) {}
}  // namespace ssr

constexpr int ROUNDS = 256;
__device__ __forceinline__ void mbar_poll(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

// mode bit0: background LDS traffic from warps 2..9; bit1: background tcgen05.ld from warps 2..9
template <int N>
__global__ void __launch_bounds__(320, 1) mma_rate(long long* out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 256 * 128 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
    fence_proxy_async();
    done = 0;
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t idesc = umma_idesc(1, 128, N);
  if (warp == 1) {
    if (lane == 0) {
      const uint64_t adesc = umma_desc_sw128(smem_u32(smem)), bdesc = umma_desc_sw128(smem_u32(smem + 16384));
      const long long t0 = clock64();
      for (int r = 0; r < ROUNDS; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<false>(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      const long long t1 = clock64();
      done = 1;
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncwarp();
  } else if (warp >= 2) {
    float acc = 0.f;
    const uint8_t* bg = smem + 16384 + 256 * 128 + (warp - 2) * 8192;
    long long n = 0;
    if (mode & 1) {
      while (!done) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(smem_u32(bg + i * 512 + lane * 16)));
          acc += x.x + x.w;
        }
        ++n;
      }
    } else if (mode & 2) {
      const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256;
      while (!done) {
        float v[32];
        tmem_ld32(tl, v);
        acc += v[0] + v[31];
        ++n;
      }
    }
    if (acc == 12345.f) out[7] = n;
    if (blockIdx.x == 0 && warp == 2 && lane == 0) out[1] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

__global__ void __launch_bounds__(512, 1) ld_rate(long long* out, int nwarps) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp < nwarps) {
    const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < ROUNDS; ++r) {
      float v[32];
      tmem_ld32(tl, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += v[i];
    }
    const long long t1 = clock64();
    if (acc == 12345.f) out[7] = 1;
    if (blockIdx.x == 0 && lane == 0 && warp == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// L2 -> SMEM streaming rate: every CTA walks the same `total` bytes (L2-resident weights) with 1-D bulk copies of
// `chunk` bytes, `depth` copies in flight.  mode 0: all CTAs start at offset 0; mode 1: staggered start.
__global__ void __launch_bounds__(64, 1) l2_rate(const uint8_t* src, int total, int chunk, int depth, int passes, int stagger,
                                                 long long* out, int copies = 1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nchunks = total / chunk;
    const int n = nchunks * passes;
    const int start = stagger == 1 ? (blockIdx.x * 7) % nchunks : 0;
    const long long t0 = clock64();
    for (int i = 0; i < n + depth; ++i) {
      if (i >= depth) {
        if (stagger == 2) mbar_poll(smem_u32(&bars[(i - depth) % depth]), ((i - depth) / depth) & 1);
        else mbar_wait(smem_u32(&bars[(i - depth) % depth]), ((i - depth) / depth) & 1);
      }
      if (i < n) {
        const int s = i % depth;
        const uint32_t bar = smem_u32(&bars[s]);
        mbar_expect_tx(bar, chunk);
        const uint8_t* g = src + (size_t)(blockIdx.x % copies) * total + (size_t)((start + i) % nchunks) * chunk;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + s * chunk)),
                     "l"(g), "r"(chunk), "r"(bar)
                     : "memory");
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}


// Same walk, but the CTAs of a cluster of CL share every chunk: CTA r fetches part r and multicasts it to all CL CTAs.
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __launch_bounds__(64, 1) l2_mc(const uint8_t* src, int total, int chunk, int depth, int passes, int CL,
                                               long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8];
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), CL);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) {
    const int nchunks = total / chunk;
    const int n = nchunks * passes;
    const int part = chunk / CL;
    const uint16_t mask = (uint16_t)((1u << CL) - 1);
    const long long t0 = clock64();
    for (int i = 0; i < n + depth; ++i) {
      if (i >= depth) {
        const int j = i - depth, s = j % depth;
        mbar_wait(smem_u32(&full[s]), (j / depth) & 1);
        for (int c = 0; c < CL; ++c) {
          uint32_t ra;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&empty[s])), "r"(c));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
        }
      }
      if (i < n) {
        const int s = i % depth;
        const uint32_t u = i / depth;
        mbar_wait(smem_u32(&empty[s]), (u & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&full[s]);
        mbar_expect_tx(bar, chunk);
        const uint8_t* g = src + (size_t)(i % nchunks) * chunk + rank * part;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                smem_u32(smem + s * chunk + rank * part)),
            "l"(g), "r"(part), "r"(bar), "h"(mask)
            : "memory");
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  cluster_sync_all();
}


// k issuing warps per CTA, each with its own ring: does the ~557 cyc/op cost serialise per thread or per SM?
__global__ void __launch_bounds__(512, 1) l2_rate_mw(const uint8_t* src, int total, int chunk, int depth, int passes, int nwarps,
                                                      long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[16][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 16; ++w)
      for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[w][i]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (lane == 0 && warp < nwarps) {
    const int nchunks = total / chunk;
    const int n = nchunks * passes / nwarps;
    uint8_t* my = smem + warp * depth * chunk;
    const long long t0 = clock64();
    for (int i = 0; i < n + depth; ++i) {
      if (i >= depth) mbar_wait(smem_u32(&bars[warp][(i - depth) % depth]), ((i - depth) / depth) & 1);
      if (i < n) {
        const int s = i % depth;
        const uint32_t bar = smem_u32(&bars[warp][s]);
        mbar_expect_tx(bar, chunk);
        const uint8_t* g = src + (size_t)((warp * 5 + i) % nchunks) * chunk;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(my + s * chunk)),
                     "l"(g), "r"(chunk), "r"(bar)
                     : "memory");
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) out[0] = t1 - t0;
  }
}


// issue cost: K back-to-back loads (no waits in between), then wait for all.  tensor = 1: 2-D tensor-map boxes [64 x rows] bf16.
__global__ void __launch_bounds__(64, 1) issue_cost(const uint8_t* src, const __grid_constant__ CUtensorMap tm, int chunk, int K, int tensor,
                                                    long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot_issue = 0, tot_all = 0;
    for (int rep = 0; rep < 8; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < K; ++i) {
        const uint32_t bar = smem_u32(&bars[i]);
        mbar_expect_tx(bar, chunk);
        if (tensor)
          tma_load_2d(smem_u32(smem + i * chunk), &tm, bar, 0, ((rep * K + i) * (chunk / 128)) % 4096);
        else
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(smem + i * chunk)),
                       "l"(src + (size_t)((rep * K + i) % 32) * chunk), "r"(chunk), "r"(bar)
                       : "memory");
      }
      const long long t1 = clock64();
      for (int i = 0; i < K; ++i) mbar_wait(smem_u32(&bars[i]), rep & 1);
      const long long t2 = clock64();
      if (rep >= 2) {
        tot_issue += t1 - t0;
        tot_all += t2 - t0;
      }
    }
    if (blockIdx.x == 0) {
      out[0] = tot_issue / 6;
      out[1] = tot_all / 6;
    }
  }
}


// ring with `per` bulk copies per stage (one mbarrier per stage): is the ~600 cyc cost per op or per wait->issue round?
__global__ void __launch_bounds__(64, 1) ring_multi(const uint8_t* src, int total, int chunk, int depth, int per, int passes, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int stage_bytes = chunk * per;
    const int nst = total / stage_bytes;
    const int n = nst * passes;
    const long long t0 = clock64();
    for (int i = 0; i < n + depth; ++i) {
      if (i >= depth) mbar_wait(smem_u32(&bars[(i - depth) % depth]), ((i - depth) / depth) & 1);
      if (blockIdx.x == 0 && i >= 200 && i < 232) out[8 + (i - 200)] = clock64() - t0;
      if (i < n) {
        const int s = i % depth;
        const uint32_t bar = smem_u32(&bars[s]);
        mbar_expect_tx(bar, stage_bytes);
        for (int j = 0; j < per; ++j) {
          const uint8_t* g = src + (size_t)(i % nst) * stage_bytes + j * chunk;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(smem + s * stage_bytes + j * chunk)),
                       "l"(g), "r"(chunk), "r"(bar)
                       : "memory");
        }
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}


// producer warp / consumer warp ring (the real kernels' pattern): warp 0 waits empty -> issues, warp 1 waits full -> arrives empty
__global__ void __launch_bounds__(64, 1) ring_pc(const uint8_t* src, int total, int chunk, int depth, int passes, int spin, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  const int nchunks = total / chunk;
  const int n = nchunks * passes;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const int s = i % depth;
      mbar_wait(smem_u32(&empty[s]), ((i / depth) & 1) ^ 1);
      mbar_expect_tx(smem_u32(&full[s]), chunk);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem + s * chunk)),
                   "l"(src + (size_t)(i % nchunks) * chunk), "r"(chunk), "r"(smem_u32(&full[s]))
                   : "memory");
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[1] = t1 - t0;
  } else if (threadIdx.x == 32) {
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const int s = i % depth;
      mbar_wait(smem_u32(&full[s]), (i / depth) & 1);
      if (spin) {
        const long long c0 = clock64();
        while (clock64() - c0 < spin) {}
      }
      mbar_arrive(smem_u32(&empty[s]));
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}


// np producer warps (warp w handles ops i with i % np == w), one consumer warp
__global__ void __launch_bounds__(192, 1) ring_np(const uint8_t* src, int total, int chunk, int depth, int passes, int np, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  const int nchunks = total / chunk;
  const int n = nchunks * passes;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) != 0) return;
  if (warp < np) {
    for (int i = warp; i < n; i += np) {
      const int s = i % depth;
      mbar_wait(smem_u32(&empty[s]), ((i / depth) & 1) ^ 1);
      mbar_expect_tx(smem_u32(&full[s]), chunk);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem + s * chunk)),
                   "l"(src + (size_t)(i % nchunks) * chunk), "r"(chunk), "r"(smem_u32(&full[s]))
                   : "memory");
    }
  } else if (warp == 5) {
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const int s = i % depth;
      mbar_wait(smem_u32(&full[s]), (i / depth) & 1);
      mbar_arrive(smem_u32(&empty[s]));
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}
// latencies of the individual producer-side instructions, single thread, steady state
__global__ void __launch_bounds__(32, 1) instr_lat(const uint8_t* src, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar2), 1);
    fence_barrier_init();
    fence_proxy_async();
    long long acc[6] = {0, 0, 0, 0, 0, 0};
    for (int it = 0; it < 20; ++it) {
      const long long a = clock64();
      mbar_expect_tx(smem_u32(&bar), 8192);
      const long long b = clock64();
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(src),
                   "r"(8192), "r"(smem_u32(&bar))
                   : "memory");
      const long long c = clock64();
      mbar_wait(smem_u32(&bar), it & 1);  // blocks until the data lands
      const long long d = clock64();
      mbar_arrive(smem_u32(&bar2));  // plain arrive completes bar2's phase
      const long long e = clock64();
      mbar_wait(smem_u32(&bar2), it & 1);  // already complete
      const long long f = clock64();
      const long long g0 = clock64();
      if (it >= 4) {
        acc[0] += b - a; acc[1] += c - b; acc[2] += d - c; acc[3] += e - d; acc[4] += f - e; acc[5] += g0 - f;
      }
    }
    for (int i = 0; i < 6; ++i) out[i] = acc[i] / 16;
  }
}


// mbarrier ping-pong between two warps: round-trip latency with try_wait (blocking) vs test_wait (polling)
__global__ void __launch_bounds__(64, 1) pingpong(int poll, long long* out) {
  __shared__ uint64_t b1, b2;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&b1), 1);
    mbar_init(smem_u32(&b2), 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int N = 200;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < N; ++i) {
      mbar_arrive(smem_u32(&b1));
      if (poll) mbar_poll(smem_u32(&b2), i & 1); else mbar_wait(smem_u32(&b2), i & 1);
    }
    out[0] = (clock64() - t0) / N;
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < N; ++i) {
      if (poll) mbar_poll(smem_u32(&b1), i & 1); else mbar_wait(smem_u32(&b1), i & 1);
      mbar_arrive(smem_u32(&b2));
    }
  }
}

template <int N>
static void run_mma(long long* d_out) {
  const int smem = 16384 + 256 * 128 + 65536 + 2048;
  cudaFuncSetAttribute(mma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 3; ++mode) {
    long long h[8] = {0};
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d_out, 0, 64);
      mma_rate<N><<<148, 320, smem>>>(d_out, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("N=%d mode=%d: %s\n", N, mode, cudaGetErrorString(e));
        exit(1);
      }
    }
    cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
    printf("mma M128 N%-3d K16 mode=%d (%s): %.1f cyc/mma (floor %d), bg iters %lld\n", N, mode,
           mode == 0 ? "quiet" : mode == 1 ? "bg LDS x8 warps" : "bg tcgen05.ld x8 warps", (double)h[0] / (ROUNDS * 4), N / 2, h[1]);
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 512);
  run_mma<32>(d_out);
  run_mma<64>(d_out);
  run_mma<128>(d_out);
  run_mma<192>(d_out);
  run_mma<256>(d_out);
  for (int nw : {1, 4, 8, 16}) {
    long long h[8] = {0};
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d_out, 0, 64);
      ld_rate<<<148, 512>>>(d_out, nw);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("ld nw=%d: %s\n", nw, cudaGetErrorString(e));
        exit(1);
      }
    }
    cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
    printf("tcgen05.ld 32x32b.x32 + 32 FADD, %2d warps: %.1f cyc per ld per warp (4 KB each)\n", nw, (double)h[0] / ROUNDS);
  }

  {
    const int total = 98304 * 6;  // 590 KB
    uint8_t* w;
    cudaMalloc(&w, total);
    cudaMemset(w, 1, total);
    cudaFuncSetAttribute(l2_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int stagger = 0; stagger < 3; stagger += 2)
      for (int depth : {1, 2, 4, 8})
        for (int chunk : {8192, 24576}) {
          if (depth * chunk > 196608) continue;
          long long h[8] = {0};
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0);
          cudaEventCreate(&e1);
          const int passes = 40;
          for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(d_out, 0, 64);
            cudaEventRecord(e0);
            l2_rate<<<148, 64, 200 * 1024>>>(w, total, chunk, depth, passes, stagger, d_out);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("l2_rate: %s\n", cudaGetErrorString(e));
              exit(1);
            }
          }
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
          const double bytes = (double)total * passes;
          printf("L2->SMEM bulk copy, 148 CTAs same 590 KB, chunk %5d depth %d stagger %d: %.1f B/cyc/SM, chip %.2f TB/s (%.3f ms)\n", chunk,
                 depth, stagger, bytes / (double)h[0], bytes * 148 / (ms * 1e-3) / 1e12, ms);
        }
  }

  {
    const int total = 24576 * 24;
    uint8_t* w;
    cudaMalloc(&w, total);
    cudaMemset(w, 1, total);
    cudaFuncSetAttribute(l2_mc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(l2_mc, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int CL : {1, 2, 4, 8})
      for (int depth : {3, 6}) {
        const int chunk = 24576;
        int nclusters = 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148 / CL * CL);
        cfg.blockDim = dim3(64);
        cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaOccupancyMaxActiveClusters(&nclusters, l2_mc, &cfg);
        cfg.gridDim = dim3(nclusters * CL);
        long long h[8] = {0};
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const int passes = 40;
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(d_out, 0, 64);
          cudaEventRecord(e0);
          cudaLaunchKernelEx(&cfg, l2_mc, (const uint8_t*)w, total, chunk, depth, passes, CL, d_out);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("l2_mc CL=%d: %s\n", CL, cudaGetErrorString(e));
            exit(1);
          }
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
        const double bytes = (double)total * passes;
        printf("multicast cluster %d (%d clusters = %d SMs) chunk %d depth %d: %.1f B/cyc/SM delivered, %.3f ms\n", CL, nclusters,
               nclusters * CL, chunk, depth, bytes / (double)h[0], ms);
      }
  }

  {
    const int total = 98304 * 6;
    uint8_t* w;
    cudaMalloc(&w, total);
    cudaMemset(w, 1, total);
    cudaFuncSetAttribute(l2_rate_mw, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int nw : {1, 2, 4, 8, 16})
      for (int chunk : {2048, 4096, 16384}) {
        const int depth = 2;
        if (nw * depth * chunk > 196608) continue;
        long long h[8] = {0};
        const int passes = 16;
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(d_out, 0, 64);
          l2_rate_mw<<<148, 512, 200 * 1024>>>(w, total, chunk, depth, passes, nw, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("l2_rate_mw: %s\n", cudaGetErrorString(e));
            exit(1);
          }
        }
        cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
        const double bytes = (double)total * passes;
        printf("bulk copy from %2d issuing warps, chunk %5d depth 2: %.1f B/cyc/SM, %.0f cyc per op per warp\n", nw, chunk,
               bytes / (double)h[0], (double)h[0] / (total / chunk * passes / nw));
      }
  }

  {
    typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncFn enc = (EncFn)fp;
    uint8_t* w;
    const size_t total = (size_t)8192 * 128;  // [8192 rows][64 bf16]
    cudaMalloc(&w, total);
    cudaMemset(w, 1, total);
    cudaFuncSetAttribute(issue_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int tensor = 0; tensor < 2; ++tensor)
      for (int rows : {32, 128, 192})
        for (int K : {1, 2, 4, 8}) {
          const int chunk = rows * 128;
          CUtensorMap tm;
          cuuint64_t dims[2] = {64, 8192};
          cuuint64_t str[1] = {128};
          cuuint32_t box[2] = {64, (cuuint32_t)rows};
          cuuint32_t es[2] = {1, 1};
          CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
          long long h[8] = {0};
          cudaMemset(d_out, 0, 64);
          issue_cost<<<148, 64, 200 * 1024>>>(w, tm, chunk, K, tensor, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("issue_cost: %s\n", cudaGetErrorString(e)); exit(1); }
          cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
          printf("issue_cost %s chunk %5d K=%d: issue %lld cyc total (%.0f per op), all done after %lld cyc\n", tensor ? "tensor2d" : "bulk1d  ",
                 chunk, K, h[0], (double)h[0] / K, h[1]);
        }
  }

  {
    const int total = 24576 * 24;
    uint8_t* w;
    cudaMalloc(&w, (size_t)total * 148);
    cudaMemset(w, 1, (size_t)total * 148);
    for (int copies : {1, 2, 4, 8, 16, 37, 148})
      for (int depth : {1, 2, 4}) {
        const int chunk = 24576, passes = 20;
        long long h[8] = {0};
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
          cudaMemset(d_out, 0, 64);
          cudaEventRecord(e0);
          l2_rate<<<148, 64, 200 * 1024>>>(w, total, chunk, depth, passes, 0, d_out, copies);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("l2_rate copies: %s\n", cudaGetErrorString(e)); exit(1); }
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
        const double bytes = (double)total * passes;
        printf("replicated weights: %3d copies, chunk %d depth %d: %.1f B/cyc/SM, chip %.2f TB/s\n", copies, chunk, depth,
               bytes / (double)h[0], bytes * 148 / (ms * 1e-3) / 1e12);
      }
  }

  {
    const int total = 24576 * 24;
    uint8_t* w;
    cudaMalloc(&w, (size_t)total);
    cudaMemset(w, 1, (size_t)total);
    cudaFuncSetAttribute(ring_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int per : {1, 2, 4, 8})
      for (int depth : {2, 4}) {
        const int chunk = 8192, passes = 20;
        if (per * chunk * depth > 196608) continue;
        long long h[8] = {0};
        for (int rep = 0; rep < 3; ++rep) {
          cudaMemset(d_out, 0, 512);
          ring_multi<<<148, 64, 200 * 1024>>>(w, total, chunk, depth, per, passes, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("ring_multi: %s\n", cudaGetErrorString(e)); exit(1); }
        }
        cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
        const double bytes = (double)total * passes;
        printf("ring: %d x 8 KB copies per stage, depth %d: %.1f B/cyc/SM, %.0f cyc per stage\n", per, depth, bytes / (double)h[0],
               (double)h[0] / (total / (chunk * per) * passes));
        {
          long long tl[40];
          cudaMemcpy(tl, d_out, 320, cudaMemcpyDeviceToHost);
          printf("   wait-return deltas:");
          for (int i = 1; i < 24; ++i) printf(" %lld", tl[8 + i] - tl[8 + i - 1]);
          printf("\n");
        }
      }
  }

  {
    const int total = 24576 * 24;
    uint8_t* w;
    cudaMalloc(&w, (size_t)total);
    cudaMemset(w, 1, (size_t)total);
    cudaFuncSetAttribute(ring_pc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int spin : {0, 200, 400})
      for (int depth : {1, 2, 4, 8})
        for (int chunk : {8192, 24576}) {
          const int passes = 20;
          if (chunk * depth > 196608) continue;
          long long h[8] = {0};
          for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(d_out, 0, 512);
            ring_pc<<<148, 64, 200 * 1024>>>(w, total, chunk, depth, passes, spin, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("ring_pc: %s\n", cudaGetErrorString(e)); exit(1); }
          }
          cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
          const double bytes = (double)total * passes;
          printf("ring_pc (2 warps) chunk %5d depth %d consumer-spin %3d: %.1f B/cyc/SM, %.0f cyc per op\n", chunk, depth, spin,
                 bytes / (double)h[0], (double)h[0] / (total / chunk * passes));
        }
  }

  {
    const int total = 24576 * 24;
    uint8_t* w;
    cudaMalloc(&w, (size_t)total);
    cudaMemset(w, 1, (size_t)total);
    cudaFuncSetAttribute(ring_np, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int np : {1, 2, 4})
      for (int depth : {4, 8}) {
        const int passes = 20, chunk = 24576;
        long long h[8] = {0};
        for (int rep = 0; rep < 3; ++rep) {
          cudaMemset(d_out, 0, 512);
          ring_np<<<148, 192, 200 * 1024>>>(w, total, chunk, depth, passes, np, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("ring_np: %s\n", cudaGetErrorString(e)); exit(1); }
        }
        cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
        const double bytes = (double)total * passes;
        printf("ring_np %d producer warps, chunk %5d depth %d: %.1f B/cyc/SM, %.0f cyc per op\n", np, chunk, depth, bytes / (double)h[0],
               (double)h[0] / (total / chunk * passes));
      }
    cudaFuncSetAttribute(instr_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    long long h[8] = {0};
    cudaMemset(d_out, 0, 512);
    instr_lat<<<1, 32, 64 * 1024>>>(w, d_out);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
    printf("instr_lat (cycles incl. ~clock64 pair): expect_tx %lld | bulk issue %lld | wait(blocking, 8 KB) %lld | arrive %lld | wait(complete) %lld | clock pair %lld\n",
           h[0], h[1], h[2], h[3], h[4], h[5]);
  }

  for (int poll = 0; poll < 2; ++poll) {
    long long h[8] = {0};
    cudaMemset(d_out, 0, 512);
    pingpong<<<1, 64>>>(poll, d_out);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
    printf("mbarrier ping-pong round trip (%s): %lld cycles\n", poll ? "test_wait polling" : "try_wait", h[0]);
  }
  return 0;
}
