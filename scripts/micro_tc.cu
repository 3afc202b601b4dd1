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
  cudaMalloc(&d_out, 64);
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
  return 0;
}
