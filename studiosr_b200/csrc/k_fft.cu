// 2-D real FFT pair of SwinFIR's FourierUnit (swinfir.py:9-34: torch.fft.rfftn / irfftn over (H, W), norm "ortho") on
// pixel-major fp32 activations.  The images of this path are small (tiles / patches of 16..128 pixels per side, any even or
// odd length), so each 1-D transform is a direct DFT with an exact twiddle table in shared memory: thread = one output
// (position, channel), consecutive threads = consecutive channels (coalesced), O(L) work per output.  A complex array is
// stored as rows of `ld` floats with the real parts in columns [0, c) and the imaginary parts in [c, 2c) -- the
// `cat(real, imag)` channel layout the 1x1 conv of the FourierUnit consumes (swinfir.py:22).
#include <math.h>

#include "ssr_device.cuh"

namespace ssr {

constexpr int FFT_MAX_LEN = 1024;

// twiddle table tw[j] = (cos, sin)(2 pi j / L), j < L
__device__ __forceinline__ void fft_twiddles(float2* tw, int L) {
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float sn, cs;
    sincospif(2.0f * (float)j / (float)L, &sn, &cs);
    tw[j] = make_float2(cs, sn);
  }
  __syncthreads();
}

// pass 1: real -> complex along W.  in [B*H*W][ld_in] (c real channels); out [B*H*Wf][ld_out], Wf = W/2 + 1
__global__ void __launch_bounds__(256) rfft_w_kernel(const float* in, int ld_in, float* out, int ld_out, int BH, int W, int c, float scale) {
  __shared__ float2 tw[FFT_MAX_LEN];
  fft_twiddles(tw, W);
  const int Wf = W / 2 + 1;
  const size_t n = (size_t)BH * Wf * c;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c);
    const size_t r = e / c;
    const int k = (int)(r % Wf);
    const size_t bh = r / Wf;
    const float* src = in + bh * W * ld_in + ch;
    float re = 0.0f, im = 0.0f;
    int j = 0;
    for (int x = 0; x < W; ++x) {
      const float v = src[(size_t)x * ld_in];
      re = fmaf(v, tw[j].x, re);
      im = fmaf(-v, tw[j].y, im);
      j += k;
      if (j >= W) j -= W;
    }
    out[r * ld_out + ch] = re * scale;
    out[r * ld_out + c + ch] = im * scale;
  }
}
// pass 2 / 3: complex -> complex along H (sign = -1 forward, +1 inverse).  in / out [B][H][Wf][ld]
__global__ void __launch_bounds__(256) cfft_h_kernel(const float* in, int ld_in, float* out, int ld_out, int B, int H, int Wf, int c, float sign,
                                                     float scale, int rtf32) {
  __shared__ float2 tw[FFT_MAX_LEN];
  fft_twiddles(tw, H);
  const size_t n = (size_t)B * H * Wf * c;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c);
    const size_t r = e / c;
    const int kw = (int)(r % Wf);
    const int k = (int)((r / Wf) % H);
    const size_t b = r / ((size_t)Wf * H);
    const float* src = in + ((b * H) * Wf + kw) * ld_in + ch;
    const size_t step = (size_t)Wf * ld_in;
    float re = 0.0f, im = 0.0f;
    int j = 0;
    for (int y = 0; y < H; ++y) {
      const float a = src[y * step], bb = src[y * step + c];
      const float cs = tw[j].x, sn = sign * tw[j].y;  // e^{sign * i * theta}
      re += a * cs - bb * sn;
      im += a * sn + bb * cs;
      j += k;
      if (j >= H) j -= H;
    }
    re *= scale;
    im *= scale;
    if (rtf32) {  // the spectrum is the A operand of a tf32 GEMM: round-to-nearest here instead of the MMA's truncation
      re = round_tf32(re);
      im = round_tf32(im);
    }
    out[r * ld_out + ch] = re;
    out[r * ld_out + c + ch] = im;
  }
}
// pass 4: complex (Hermitian half spectrum) -> real along W, + `add` (the FourierUnit's caller adds its input, swinfir.py:49).
// The imaginary parts of the DC and (even W) Nyquist bins are ignored, as by every c2r transform.
__global__ void __launch_bounds__(256) irfft_w_kernel(const float* in, int ld_in, const float* add, int ld_add, void* out, int ld_out,
                                                      int BH, int W, int c, float scale, int elem, int rtf32) {
  __shared__ float2 tw[FFT_MAX_LEN];
  fft_twiddles(tw, W);
  const int Wf = W / 2 + 1;
  const size_t n = (size_t)BH * W * c;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c);
    const size_t r = e / c;
    const int x = (int)(r % W);
    const size_t bh = r / W;
    const float* src = in + bh * Wf * ld_in + ch;
    float acc = src[0];  // k = 0
    int j = 0;
    for (int k = 1; k < Wf; ++k) {
      j += x;
      if (j >= W) j -= W;
      const float a = src[(size_t)k * ld_in], bb = src[(size_t)k * ld_in + c];
      if (2 * k == W)
        acc = fmaf(a, tw[j].x, acc);  // Nyquist: cos(pi x) = +-1, counted once
      else
        acc += 2.0f * (a * tw[j].x - bb * tw[j].y);
    }
    store_elem(out, r * ld_out + ch, elem, fmaf(acc, scale, add ? add[r * ld_add + ch] : 0.0f), rtf32);
  }
}

static unsigned fft_grid(size_t n) { return (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16); }

int launch_rfft2(const float* in, int ld_in, float* tmp, float* spec, int ld_spec, int B, int H, int W, int c, int rtf32, cudaStream_t s) {
  SSR_CHECK(H <= FFT_MAX_LEN && W <= FFT_MAX_LEN && 2 * c <= ld_spec, SSR_E_INVALID, "rfft2: %dx%d, c=%d", H, W, c);
  const int Wf = W / 2 + 1;
  const float sc = 1.0f / sqrtf((float)H * (float)W);  // norm = "ortho", split evenly over the two passes' product
  ProfScope prof("fft", 8.0 * B * H * Wf * c * (W / 2 + H), (double)B * H * (W + 4.0 * Wf) * c * 4, s);
  rfft_w_kernel<<<fft_grid((size_t)B * H * Wf * c), 256, 0, s>>>(in, ld_in, tmp, ld_spec, B * H, W, c, sc);
  count_launch();
  cfft_h_kernel<<<fft_grid((size_t)B * H * Wf * c), 256, 0, s>>>(tmp, ld_spec, spec, ld_spec, B, H, Wf, c, -1.0f, 1.0f, rtf32);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}
int launch_irfft2_add(const float* spec, int ld_spec, float* tmp, const float* add, int ld_add, void* out, int ld_out, int B, int H, int W,
                      int c, int elem, int rtf32, cudaStream_t s) {
  SSR_CHECK(H <= FFT_MAX_LEN && W <= FFT_MAX_LEN && 2 * c <= ld_spec, SSR_E_INVALID, "irfft2: %dx%d, c=%d", H, W, c);
  const int Wf = W / 2 + 1;
  const float sc = 1.0f / sqrtf((float)H * (float)W);
  ProfScope prof("fft", 8.0 * B * H * Wf * c * (W / 2 + H), (double)B * H * (W + 4.0 * Wf) * c * 4, s);
  cfft_h_kernel<<<fft_grid((size_t)B * H * Wf * c), 256, 0, s>>>(spec, ld_spec, tmp, ld_spec, B, H, Wf, c, 1.0f, 1.0f, 0);
  count_launch();
  irfft_w_kernel<<<fft_grid((size_t)B * H * W * c), 256, 0, s>>>(tmp, ld_spec, add, ld_add, out, ld_out, B * H, W, c, sc, elem, rtf32);
  count_launch();
  SSR_CUDA(cudaGetLastError());
  return SSR_OK;
}

}  // namespace ssr
