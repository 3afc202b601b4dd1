// Developer experiment (GPU box): can ONE TMA-loaded halo tile in shared memory serve all nine taps of a 3x3 conv as the
// A operand of tcgen05.mma?  The tile is [R rows = halo pixels][64 ch] bf16, 128 B per row, SWIZZLE_128B as written by TMA.
// A tap's operand is the same tile read from a start address shifted by r0 rows with an 8-row-group stride (SBO) of
// `pitch` rows.  Whether that works depends on how the tensor core derives the swizzle phase (absolute address bits vs
// row index + the descriptor's base_offset field), which the guides do not state -- this program measures it:
// for several (r0, pitch, base_offset policy) it compares D = A_shifted * W^T against a CPU reference.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I studiosr_b200/csrc scripts/micro_halo.cu -o /tmp/micro_halo -lcuda
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ssr_tc.cuh"

using namespace ssr;
namespace ssr {
void set_error(const char*, ...) {}
}  // namespace ssr

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int R = 320;  // rows of the halo tile (two TMA boxes of 160 rows)

__global__ void __launch_bounds__(128, 1) halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                     float* out, int r0, int pitch, int bo_policy) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sX = smem;             // R x 128 B
  uint8_t* sW = smem + R * 128;   // 64 x 128 B (R * 128 is a multiple of 1024)
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc<64>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bars[0]), R * 128 + 64 * 128);
    tma_load_2d(smem_u32(sX), &tmX, smem_u32(&bars[0]), 0, 0);
    tma_load_2d(smem_u32(sX + 160 * 128), &tmX, smem_u32(&bars[0]), 0, 160);
    tma_load_2d(smem_u32(sW), &tmW, smem_u32(&bars[0]), 0, 0);
    mbar_wait(smem_u32(&bars[0]), 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sX) + (uint32_t)r0 * 128u;
    uint64_t adesc = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((pitch * 128) >> 4) << 32) | (1ull << 46) |
                     (2ull << 61);
    if (bo_policy == 1) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sW));
    constexpr uint32_t idesc = umma_idesc(1, 128, 64);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma<false>(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
    umma_commit(smem_u32(&bars[1]));
  }
  __syncthreads();
  mbar_wait_warp(smem_u32(&bars[1]), 0, lane);
  tc_fence_after();
  float v[32];
  for (int c = 0; c < 2; ++c) {
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c * 32 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

// issue rate of M128 x N x K16 MMAs whose A operand is the shifted / pitched view (B = the W tile repeated)
template <int N>
__global__ void __launch_bounds__(128, 1) halo_rate(long long* out, int r0, int pitch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (R * 128 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc<256>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem) + (uint32_t)r0 * 128u;
    const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((pitch * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + R * 128));
    constexpr uint32_t idesc = umma_idesc(1, 128, N);
    const long long t0 = clock64();
    for (int r = 0; r < 256; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma<false>(tmem_slot, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_slot);
  }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  std::vector<float> X(R * 64), W(64 * 64);
  srand(1);
  for (auto& x : X) x = bf((rand() % 2001 - 1000) / 1000.0f);
  for (auto& x : W) x = bf((rand() % 2001 - 1000) / 1000.0f);
  std::vector<__nv_bfloat16> Xh(X.size()), Wh(W.size());
  for (size_t i = 0; i < X.size(); ++i) Xh[i] = __float2bfloat16_rn(X[i]);
  for (size_t i = 0; i < W.size(); ++i) Wh[i] = __float2bfloat16_rn(W[i]);
  __nv_bfloat16 *dX, *dW;
  float* dO;
  cudaMalloc(&dX, Xh.size() * 2);
  cudaMalloc(&dW, Wh.size() * 2);
  cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dX, Xh.data(), Xh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, Wh.data(), Wh.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmX, tmW;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)R};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, 160};
    enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  {
    cuuint64_t dims[2] = {64, 64};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, 64};
    enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  const size_t smem = R * 128 + 64 * 128 + 2048;
  cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> out(128 * 64);
  for (int pitch : {8, 16, 10, 18})
    for (int r0 : {0, 1, 3, 8, 9, 17, 19, 37})
      for (int pol : {0, 1}) {
        if (r0 + 15 * pitch + 7 >= R) continue;
        halo_kernel<<<1, 128, smem>>>(tmX, tmW, dO, r0, pitch, pol);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("pitch %d r0 %d pol %d: %s\n", pitch, r0, pol, cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < 128; ++m) {
          const int row = r0 + (m / 8) * pitch + m % 8;
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)X[row * 64 + k] * W[n * 64 + k];
            worst = fmax(worst, fabs(ref - out[m * 64 + n]));
          }
        }
        printf("pitch %2d rows  r0 %2d  base_offset %s : max |err| = %.4f  %s\n", pitch, r0, pol ? "(addr>>7)&7" : "0          ", worst,
               worst < 1e-2 ? "OK" : "MISMATCH");
      }
  long long* dT;
  cudaMalloc(&dT, 64);
  const size_t smem2 = R * 128 + 256 * 128 + 2048;
  cudaFuncSetAttribute(halo_rate<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  cudaFuncSetAttribute(halo_rate<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  cudaFuncSetAttribute(halo_rate<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  for (int pitch : {8, 10, 16})
    for (int r0 : {0, 1, 4, 11, 21}) {
      long long h64, h192, h256;
      halo_rate<64><<<1, 128, smem2>>>(dT, r0, pitch);
      cudaDeviceSynchronize();
      cudaMemcpy(&h64, dT, 8, cudaMemcpyDeviceToHost);
      halo_rate<192><<<1, 128, smem2>>>(dT, r0, pitch);
      cudaDeviceSynchronize();
      cudaMemcpy(&h192, dT, 8, cudaMemcpyDeviceToHost);
      halo_rate<256><<<1, 128, smem2>>>(dT, r0, pitch);
      cudaDeviceSynchronize();
      cudaMemcpy(&h256, dT, 8, cudaMemcpyDeviceToHost);
      printf("MMA rate  pitch %2d r0 %2d : N=64 %.1f  N=192 %.1f  N=256 %.1f cycles per MMA (M128 K16)\n", pitch, r0, h64 / 1024.0, h192 / 1024.0,
             h256 / 1024.0);
    }
  return 0;
}
