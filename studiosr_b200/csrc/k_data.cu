// Device-side evaluation metric and training augmentation: the reference computes Y-PSNR on the host per image
// (utils/metrics.py:11-49, evaluator.py:53-79 -- one D2H of the 4x image per evaluation image) and augments training pairs with
// numpy on the data-loader workers (data/transforms.py:8-61, dataset.py:50-58).  Both are byte streams: one pass each.
#include "ssr_device.cuh"

namespace ssr {

constexpr int PS_THREADS = 256;
constexpr int PS_MAX_BLOCKS = 148 * 8;

// to_y (metrics.py:11-17): image.astype(float32) / 255 -> float32, dot with [65.481, 128.553, 24.966] in float64, + 16
__device__ __forceinline__ float to_y_f32(const uint8_t* p) {
  const float r = (float)p[0] / 255.0f, g = (float)p[1] / 255.0f, b = (float)p[2] / 255.0f;
  return (float)((double)r * 65.481 + (double)g * 128.553 + (double)b * 24.966 + 16.0);  // compute_psnr casts back to float32
}

__global__ void __launch_bounds__(PS_THREADS) psnr_partial_kernel(const uint8_t* __restrict__ a, int wa, const uint8_t* __restrict__ b,
                                                                  int wb, int y0, int x0, int hh, int ww, int y_only,
                                                                  double* __restrict__ partial) {
  __shared__ double red[PS_THREADS / 32];
  double acc = 0.0;
  const long long n = (long long)hh * ww;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = y0 + (int)(i / ww), x = x0 + (int)(i % ww);
    const uint8_t* pa = a + ((size_t)y * wa + x) * 3;
    const uint8_t* pb = b + ((size_t)y * wb + x) * 3;
    if (y_only) {
      const float d = to_y_f32(pa) - to_y_f32(pb);
      acc += (double)(d * d);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = (float)pa[c] - (float)pb[c];
        acc += (double)(d * d);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < PS_THREADS / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}
__global__ void psnr_final_kernel(const double* __restrict__ partial, int nblocks, double inv_n, double* __restrict__ mse) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *mse = s * inv_n;
}

// One training pair per blockIdx.y: crop -> fliplr -> flipud -> rot90 (the order of dataset.py:50-58) -> CHW float / 255
// (transforms.py:64-68 array2tensor), for the LR patch (blockIdx.z = 0) and the scale-times larger HR patch (blockIdx.z = 1).
struct AugPair {
  const uint8_t* lq;
  const uint8_t* gt;
  int lq_w, gt_w;  // row pitch in pixels
  int xs, ys;      // crop origin in the LR image
  int flags;       // bit 0 fliplr, bit 1 flipud, bit 2 rot90 (np.rot90: counter-clockwise)
  int pad;
};
__global__ void __launch_bounds__(256) augment_pairs_kernel(const AugPair* __restrict__ pairs, int size, int scale, float* __restrict__ x_out,
                                                            float* __restrict__ y_out) {
  const AugPair p = pairs[blockIdx.y];
  const bool hr = blockIdx.z == 1;
  const int n = hr ? size * scale : size;
  const uint8_t* src = hr ? p.gt : p.lq;
  const int pitch = hr ? p.gt_w : p.lq_w;
  const int oy = hr ? p.ys * scale : p.ys, ox = hr ? p.xs * scale : p.xs;
  float* out = (hr ? y_out : x_out) + (size_t)blockIdx.y * 3 * n * n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
    int i = idx / n, j = idx - i * n;      // output position
    if (p.flags & 4) {                     // rot90: out[i][j] = in[j][n - 1 - i]
      const int t = i;
      i = j;
      j = n - 1 - t;
    }
    if (p.flags & 2) i = n - 1 - i;        // flipud
    if (p.flags & 1) j = n - 1 - j;        // fliplr
    const uint8_t* px = src + ((size_t)(oy + i) * pitch + (ox + j)) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(size_t)c * n * n + idx] = (float)px[c] / 255.0f;
  }
}

}  // namespace ssr

using namespace ssr;

extern "C" {

size_t ssr_psnr_workspace_bytes(void) { return (size_t)PS_MAX_BLOCKS * sizeof(double); }

int ssr_psnr_mse_u8(const uint8_t* im1, int h1, int w1, const uint8_t* im2, int h2, int w2, int crop_border, int y_only, double* dev_mse,
                    void* workspace, size_t workspace_bytes, void* stream) {
  SSR_CHECK(im1 && im2 && dev_mse && workspace, SSR_E_INVALID, "ssr_psnr_mse_u8: null argument");
  SSR_CHECK(workspace_bytes >= ssr_psnr_workspace_bytes(), SSR_E_WORKSPACE, "ssr_psnr_mse_u8: workspace %zu B too small", workspace_bytes);
  const int h = h1 < h2 ? h1 : h2, w = w1 < w2 ? w1 : w2;  // crop_img_to_equal (metrics.py:20-33): trailing rows / columns dropped
  const int hh = h - 2 * crop_border, ww = w - 2 * crop_border;
  SSR_CHECK(crop_border >= 0 && hh > 0 && ww > 0, SSR_E_INVALID, "ssr_psnr_mse_u8: nothing left after cropping %d from %dx%d", crop_border, h, w);
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)hh * ww;
  long long blocks = (n + PS_THREADS - 1) / PS_THREADS;
  if (blocks > PS_MAX_BLOCKS) blocks = PS_MAX_BLOCKS;
  ProfScope prof("psnr_u8", 0.0, (double)n * 6, s);
  psnr_partial_kernel<<<(int)blocks, PS_THREADS, 0, s>>>(im1, w1, im2, w2, crop_border, crop_border, hh, ww, y_only,
                                                        reinterpret_cast<double*>(workspace));
  psnr_final_kernel<<<1, 32, 0, s>>>(reinterpret_cast<const double*>(workspace), (int)blocks, 1.0 / ((double)n * (y_only ? 1 : 3)), dev_mse);
  count_launch(2);
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

int ssr_augment_pairs_u8(const void* dev_pairs, int n_pairs, int size, int scale, float* x_out, float* y_out, void* stream) {
  SSR_CHECK(dev_pairs && x_out && y_out && n_pairs > 0 && size > 0 && scale > 0, SSR_E_INVALID, "ssr_augment_pairs_u8: bad argument");
  static_assert(sizeof(AugPair) == 40, "AugPair layout is part of the ABI (see include/ssr_b200.h)");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = size * scale;
  int bx = (n * n + 255) / 256;
  if (bx > 64) bx = 64;
  ProfScope prof("augment_pairs", 0.0, (double)n_pairs * (size * size + (double)n * n) * 15, s);
  augment_pairs_kernel<<<dim3(bx, n_pairs, 2), 256, 0, s>>>(reinterpret_cast<const AugPair*>(dev_pairs), size, scale, x_out, y_out);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // extern "C"
