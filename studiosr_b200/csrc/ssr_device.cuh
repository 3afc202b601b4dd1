// Device-side helpers shared by all kernels.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ssr_internal.cuh"

namespace ssr {

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_LEAKY: return v > 0.0f ? v : v * slope;
    case ACT_GELU: return gelu_erf(v);
    default: return v;
  }
}

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float load(const float* p) { return *p; }
  static __device__ __forceinline__ void store(float* p, float v, int rtf32) { *p = rtf32 ? round_tf32(v) : v; }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float v, int) { *p = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// generic element load/store by byte size (2 = bf16, 4 = fp32)
__device__ __forceinline__ float load_elem(const void* base, size_t idx, int elem) {
  return elem == 2 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx])
                   : reinterpret_cast<const float*>(base)[idx];
}
__device__ __forceinline__ void store_elem(void* base, size_t idx, int elem, float v, int rtf32) {
  if (elem == 2)
    reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<float*>(base)[idx] = rtf32 ? round_tf32(v) : v;
}

// Pixel-shuffle destination of GEMM column n (already permuted to (i, j, c) order at pack time)
// for source pixel m = (b, y, x): element offset into out_T [B][H*r][W*r][ld].
__device__ __forceinline__ size_t ps_offset(int b, int y, int x, int n, int H, int W, int r, int Cps, int ld) {
  int q = n / Cps, c = n - q * Cps;
  int i = q / r, j = q - i * r;
  return ((size_t)((size_t)b * (H * r) + (y * r + i)) * (W * r) + (x * r + j)) * ld + c;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// region id of a coordinate in the shifted frame (reference common.py:253-262)
__device__ __forceinline__ int shift_region(int p, int L, int ws, int shift) {
  return p >= L - shift ? 2 : (p >= L - ws ? 1 : 0);
}

}  // namespace ssr
