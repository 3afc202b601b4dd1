// Reconstruction conv (64 -> 3 channels, 3x3; swinir.py:326,366, hat.py conv_last, rcan.py tail) on tcgen05 with the TAPS in N:
//
//   D[p][tap * 3 + c] = sum_ch x[p][ch] * w[c][ch][tap]          one [halo pixels x 64] x [64 x 27] GEMM per tile (K = 64 only)
//   y[c](oy, ox)      = bias[c] + sum_tap D[(oy + ky) * (PW + 2) + (ox + kx)][tap * 3 + c]     gathered from shared memory
//
// As an implicit GEMM with N = 3 padded to 64 the conv needed nine K = 64 k-blocks per 128 pixels and an SS-mode tcgen05.mma
// never takes less than ~44 cycles (profiles/r01_micro_tc.txt): 36 MMAs = 1.6k cycles per 128 pixels, 2.5 ms per cfg5 frame at
// 16 % of the HBM roofline.  With the taps in N one tile of 8 x 32 output pixels is 12 MMAs over its (8+2) x (32+2) halo
// (ONE TMA box, out-of-image pixels zero-filled = the conv padding) and the kernel is what it should be: a single read of the
// 64-channel HR activation (128 B per pixel) and a 3-channel write.
// Warps: 0 = TMA producer (halo tiles, two slots), 1 = MMA issuer, 2..5 = epilogue (TMEM -> smem, 9-tap gather, affine, store).
#include "ssr_tc.cuh"

namespace ssr {

constexpr int CL_PH = 8, CL_PW = 32;                    // output pixels per tile
constexpr int CL_HW = CL_PW + 2, CL_HH = CL_PH + 2;     // halo tile
constexpr int CL_ROWS = CL_HW * CL_HH;                  // 340 halo pixels
constexpr int CL_MT = (CL_ROWS + 127) / 128;            // 3 M tiles
constexpr uint32_t CL_SLOT = CL_MT * 128 * 128;         // 48 KB: [384 rows][64 ch] bf16, SWIZZLE_128B
constexpr int CL_SPITCH = 33;                           // floats per row of the D staging (odd: conflict-free column walks)
constexpr uint32_t CL_OFF_W = 2 * CL_SLOT;              // [32][64] bf16 SW128: rows = tap * 3 + c
constexpr uint32_t CL_OFF_S = CL_OFF_W + 4096;
constexpr uint32_t CL_OFF_BAR = CL_OFF_S + CL_MT * 128 * CL_SPITCH * 4;
constexpr uint32_t CL_SMEM = CL_OFF_BAR + 128 + 1024;
constexpr int CL_THREADS = 192;

struct ConvLastTcArgs {
  int B, H, W;          // input (= padded output) size
  int crop_h, crop_w;   // stored output size
  int tiles_x, tiles_y, n_tiles;
  float bias[3], out_shift[3];
  float out_scale, u8_scale;
  float* out_f32;       // [B][3][crop_h][crop_w] or null
  uint8_t* out_u8;      // [B][crop_h][crop_w][3] or null
};

__global__ void __launch_bounds__(CL_THREADS, 1)
conv_last_tapn_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvLastTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* S = reinterpret_cast<float*>(smem + CL_OFF_S);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CL_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const uint32_t sbase = smem_u32(smem), bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 + s); };
  auto dfull_bar = [&](int b) { return bar0 + 8u * (4 + b); };
  auto dempty_bar = [&](int b) { return bar0 + 8u * (6 + b); };
  const uint32_t wfull_bar = bar0 + 64u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
      mbar_init(dfull_bar(i), 1);
      mbar_init(dempty_bar(i), 4);
    }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto tile_origin = [&](int tile, int& b, int& y0, int& x0) {
    x0 = (tile % a.tiles_x) * CL_PW;
    const int t = tile / a.tiles_x;
    y0 = (t % a.tiles_y) * CL_PH;
    b = t / a.tiles_y;
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, 4096);
      tma_load_2d(sbase + CL_OFF_W, &tmW, wfull_bar, 0, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(empty_bar(s), (((uint32_t)it >> 1) & 1u) ^ 1u);
        int b, y0, x0;
        tile_origin(tile, b, y0, x0);
        mbar_expect_tx(full_bar(s), CL_ROWS * 128);
        tma_load_4d(sbase + s * CL_SLOT, &tmX, full_bar(s), 0, x0 - 1, y0 - 1, b);  // out-of-image pixels arrive as zeros
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc(1, 128, 32);
      mbar_wait(wfull_bar, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(dempty_bar(s), (((uint32_t)it >> 1) & 1u) ^ 1u);  // the epilogue has copied D_s of tile it-2 to smem
        mbar_wait(full_bar(s), ((uint32_t)it >> 1) & 1u);
        tc_fence_after();
        const uint64_t bdesc = umma_desc_sw128(sbase + CL_OFF_W);
#pragma unroll
        for (int mt = 0; mt < CL_MT; ++mt) {
          const uint64_t adesc = umma_desc_sw128(sbase + s * CL_SLOT + mt * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<false>(tmem_base + (uint32_t)(s * 96 + mt * 32), adesc + 2 * k, bdesc + 2 * k, IDESC, k ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // the halo tile has been read
        umma_commit(dfull_bar(s));
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int et = threadIdx.x - 64;  // 0..127
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait_warp(dfull_bar(s), ((uint32_t)it >> 1) & 1u, lane);
      tc_fence_after();
      // phase 1: D (27 of 32 columns) -> S[halo pixel][tap * 3 + c]
#pragma unroll
      for (int mt = 0; mt < CL_MT; ++mt) {
        float v[32];
        tmem_ld32(tlane + (uint32_t)(s * 96 + mt * 32), v);
        float* dst = S + (size_t)(mt * 128 + quad * 32 + lane) * CL_SPITCH;
#pragma unroll
        for (int j = 0; j < 27; ++j) dst[j] = v[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dempty_bar(s));
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // phase 2: 9-tap gather, affine, crop, store.  Consecutive threads = consecutive pixels of an output row.
      int b, y0, x0;
      tile_origin(tile, b, y0, x0);
#pragma unroll
      for (int r = 0; r < (CL_PH * CL_PW) / 128; ++r) {
        const int idx = r * 128 + et, oy = idx / CL_PW, ox = idx - oy * CL_PW;
        float acc[3] = {a.bias[0], a.bias[1], a.bias[2]};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float* src = S + (size_t)((oy + ky) * CL_HW + ox + kx) * CL_SPITCH + (ky * 3 + kx) * 3;
            acc[0] += src[0];
            acc[1] += src[1];
            acc[2] += src[2];
          }
        const int py = y0 + oy, px = x0 + ox;
        if (b < a.B && py < a.crop_h && px < a.crop_w) {
          float o3[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) o3[c] = (acc[c] + a.out_shift[c]) * a.out_scale;
          if (a.out_f32) {
            const size_t plane = (size_t)a.crop_h * a.crop_w;
            float* o = a.out_f32 + (size_t)b * 3 * plane + (size_t)py * a.crop_w + px;
            o[0] = o3[0];
            o[plane] = o3[1];
            o[2 * plane] = o3[2];
          }
          if (a.out_u8) {
            uint8_t* o = a.out_u8 + ((size_t)(b * a.crop_h + py) * a.crop_w + px) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = (uint8_t)fminf(fmaxf(rintf(o3[c] * a.u8_scale), 0.0f), 255.0f);
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // S may be overwritten by the next tile's phase 1
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// x: bf16 [B][H][W][ld] (ld >= 64, channels [0,64) used); w27: bf16 [32][64] rows = tap * 3 + c (27 used, 5 zero rows)
int launch_conv_last_tapn(const void* x, int ld, const void* w27, const float* bias3, const float* out_shift, float out_scale, float u8_scale,
                          int B, int H, int W, int crop_h, int crop_w, float* out_f32, uint8_t* out_u8, cudaStream_t s) {
  CUtensorMap tmX, tmW;
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {64, CL_HW, CL_HH, 1};
    SSR_TRY(make_tmap(&tmX, x, 2, 4, dims, str, box, 128));
  }
  {
    cuuint64_t dims[2] = {64, 32};
    cuuint64_t str[1] = {64 * 2};
    cuuint32_t box[2] = {64, 32};
    SSR_TRY(make_tmap(&tmW, w27, 2, 2, dims, str, box, 128));
  }
  ConvLastTcArgs a;
  a.B = B; a.H = H; a.W = W; a.crop_h = crop_h; a.crop_w = crop_w;
  a.tiles_x = (W + CL_PW - 1) / CL_PW;
  a.tiles_y = (H + CL_PH - 1) / CL_PH;
  a.n_tiles = a.tiles_x * a.tiles_y * B;
  for (int i = 0; i < 3; ++i) {
    a.bias[i] = bias3[i];
    a.out_shift[i] = out_shift[i];
  }
  a.out_scale = out_scale; a.u8_scale = u8_scale; a.out_f32 = out_f32; a.out_u8 = out_u8;
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CUDA(cudaFuncSetAttribute(conv_last_tapn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CL_SMEM));
    attr_set = true;
  }
  const int sms = num_sms_cached();
  const double px = (double)B * H * W;
  ProfScope prof("conv_last_tapn", 2.0 * px * 9 * 64 * 3, px * 64 * 2 + (double)B * crop_h * crop_w * 3 * (out_f32 ? 4 : 1), s);
  conv_last_tapn_kernel<<<a.n_tiles < sms ? a.n_tiles : sms, CL_THREADS, CL_SMEM, s>>>(tmX, tmW, a);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
