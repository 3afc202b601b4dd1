// Internal declarations shared by the kernels and the host-side model executor.
// Not part of the public ABI (see include/ssr_b200.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssr_b200.h"

namespace ssr {

// ---------------------------------------------------------------------------------------------
// errors
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SSR_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ssr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SSR_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)

#define SSR_CHECK(cond, code, ...)   \
  do {                               \
    if (!(cond)) {                   \
      ssr::set_error(__VA_ARGS__);   \
      return (code);                 \
    }                                \
  } while (0)

#define SSR_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != SSR_OK) return _r; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// optional per-launch profiling (ssr_profile_begin / ssr_profile_end): CUDA events around every
// launch on its own stream, aggregated per kernel class with the launch's algorithmic FLOPs/bytes.
struct ProfScope {
  void* rec = nullptr;
  cudaStream_t stream;
  ProfScope(const char* cls, double flops, double bytes, cudaStream_t s);
  ~ProfScope();
};

// ---------------------------------------------------------------------------------------------
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_GELU = 3 };

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// One fused "implicit GEMM + epilogue" launch.  Rows are NHWC pixels / tokens.
//   acc[m][n] = sum_{tap,c} A[pix(m)+tap][c] * Wt[n][tap*KP + c]
//   v = act(acc + bias[n]) * alpha + res[m][n]
//   out_f32[m][n] = v ; out_T[m][n] (or pixel-shuffled) = T(v) ; out_ln[m][n] = T(LN(v)*gamma+beta)
struct GemmArgs {
  const void* A;  // T [B*H*W][lda]
  int lda;
  int M;  // B*H*W
  int B, H, W;
  int taps;        // 1 (linear) or 9 (3x3 conv, zero pad 1)
  int KP;          // padded input channels (multiple of 64)
  const void* Wt;  // T [NP][taps*KP]
  int N, NP;       // real / padded output channels (NP multiple of 64; pads are zero)
  const float* bias;  // [NP]
  int act;
  float slope;
  float alpha;
  const float* res;  // fp32 [M][ldres] or null
  int ldres;
  float* out_f32;  // or null
  int ld_f32;
  void* out_T;  // or null
  int ld_T;
  int ps_r;      // PixelShuffle factor folded into the out_T store (0/1 = none); N = r*r*Cps
  void* out_ln;  // or null
  int ld_ln;
  const float* gamma;  // [NP] (pads zero)
  const float* beta;
  float eps;
  int round_tf32;  // round T=float stores to tf32 (rna) so the next tf32 MMA sees exact operands
  int split_tf32;  // 3xTF32: Wt holds [2 NP] rows (tf32 heads, then tails); activations are full fp32 and split in the kernel
  // reconstruction-conv mode (tensor-core path only): v[0..2] -> (v + out_shift) * out_scale, cropped to crop_h x crop_w,
  // stored as fp32 NCHW [B,3,crop_h,crop_w] and / or uint8 HWC (round-half-even, clip); all other outputs unused
  float* out3_f32;
  uint8_t* out3_u8;
  int crop_h, crop_w;
  float out_shift[3];
  float out_scale, u8_scale;
  int K_alg, N_alg;  // un-padded contraction / output widths, for FLOP and byte accounting only
  int tma_store;     // set by the launcher: out_T (and out_pre) leave through TMA stores of swizzled 32 x 32 boxes
  // backward of an activation folded into a dgrad epilogue (tensor-core path only): after alpha,
  //   v = mask[m][n] > 0 ? v : v * mask_slope     (ReLU: slope 0 with mask = the forward OUTPUT; LeakyReLU likewise)
  // mask_mode 1 (GELU backward): v *= gelu'(mask[m][n]) with mask = the saved PRE-activation
  // mask_mode 2: v *= mask[m][n] (mask = the activation's derivative, saved by the forward with pre_mode 1)
  const void* mask;  // T [M][ld_mask] or null
  int ld_mask;
  float mask_slope;
  int mask_mode;
  void* out_pre;  // T [M][ld_pre]: copy of (acc + bias) BEFORE the activation (kept for the backward), or null
  int ld_pre;
  int pre_mode;   // 0: out_pre = pre-activation u; 1 (act == GELU only): out_pre = gelu'(u) = Phi(u) + u phi(u), which shares the
                  // exponential of the erf evaluation, so the backward's epilogue is one multiply
  // stochastic depth (timm DropPath, swinir.py:137,171-172): v *= row_scale[m / rows_per_scale] before the residual add
  const float* row_scale;  // [ceil(M / rows_per_scale)] or null (tensor-core path only)
  int rows_per_scale;
  long long* dbg;    // optional per-CTA phase timestamps (developer diagnostics), 8 slots per CTA
};
double gemm_alg_flops(const GemmArgs& g);
double gemm_alg_bytes(const GemmArgs& g, int elem);

struct AttnArgs {
  const void* qkv;  // T [B*H*W][ld_qkv]; q | k | v, each [heads][DP] (q pre-scaled)
  int ld_qkv;
  int QP;   // heads*DP
  void* o;  // T [B*H*W][ld_o], channel = head*DP + j
  int ld_o;
  const float* bias;  // [heads][(2ws-1)^2]
  int B, H, W, ws, shift, heads, d, DP;
  int elem;  // bytes per activation element for the CUDA-core kernels (0 / 4: fp32, 2: bf16)
  int kws;   // key / value window of the HAT overlapping cross-attention (launch_attn_oca)
};

struct LnArgs {
  const float* in;  // [M][ld_in]
  int ld_in;
  int M, C, CP;
  const float *g1, *b1;  // first LN
  float* out_f32;        // LN1 result (or null)
  int ld_f32;
  const float *g2, *b2;  // optional second LN applied to the first one's output (or null)
  void* out_T;           // T(LN2(LN1(x))) or T(LN1(x)) if g2 null (or null)
  int ld_T;
  int elem;  // sizeof(T): 2 or 4
  int round_tf32;
  float eps;
};

// first conv (n_colors -> Cout) fused with: uint8/float input, tile addressing into a frame,
// mirror padding (eval / train flavour) and the input affine (normalisation / MeanShift).
struct ConvFirstArgs {
  const void* in;  // float NCHW [Bimg,3,fh,fw]  or uint8 HWC [Bimg,fh,fw,3]
  int in_u8;
  int fh, fw;  // source image / frame size
  // tile mode: batch item b is the th x tw crop at (ty0[b], tx0[b]) of image 0; otherwise the whole image b
  int tile_mode, tile, stride, tiles_x, tile_begin;
  int h, w;    // logical (un-padded) input size per batch item
  int Hp, Wp;  // padded size
  int pad_mode;  // 0 eval (edge-inclusive mirror), 1 train (reflect), 2 none
  int B;
  float in_scale;
  float in_shift[3];
  const float* Wc;    // [Cout][27] (ci,ky,kx)
  const float* bias;  // [Cout]
  int Cout;
  float* out_f32;  // [B*Hp*Wp][ld_f32] or null
  int ld_f32;
  void* out_T;
  int ld_T;
  int elem;
  int round_tf32;
  // fused patch-embed LayerNorm chain (launch_conv_first_ln, SwinIR): g = LN(conv; g1, b1) -> out_g (fp32), LN(g; g2, b2) -> out_T
  const float *g1, *b1, *g2, *b2;  // padded fp32 vectors [CP]
  float* out_g;
  int ld_g;
  float eps;
};
int launch_conv_first_ln(const ConvFirstArgs& a, int CP, cudaStream_t s);
// reconstruction conv 64 -> 3 with the taps in N (k_conv_last.cu); x bf16 NHWC, w27 bf16 [32][64] (row = tap * 3 + c)
int launch_conv_last_tapn(const void* x, int ld, const void* w27, const float* bias3, const float* out_shift, float out_scale, float u8_scale,
                          int B, int H, int W, int crop_h, int crop_w, float* out_f32, uint8_t* out_u8, cudaStream_t s);

// last conv (Cin -> 3) fused with bias, output affine (un-normalise / MeanShift), crop and either
// fp32 NCHW store or uint8 HWC quantisation (round-half-even, clip).
struct ConvLastArgs {
  const void* in;  // T [B][Hs][Ws][ldi]
  int ldi, Cin, elem;
  int B, Hs, Ws;  // input (padded, upscaled) size
  int ch, cw;     // cropped output size
  const float* Wc;  // [9][Cin][4] (tap, ci, co padded to 4)
  float bias[3];
  float out_shift[3];
  float out_scale;
  float* out_f32;  // NCHW [B,3,ch,cw] or null
  uint8_t* out_u8;  // HWC [B,ch,cw,3] or null
  float u8_scale;
};

struct BlendArgs {
  const float* tiles;  // [nt][3][ts][ts] fp32 (ts = tile*scale)
  uint8_t* out;        // [H*s][W*s][3]
  int H, W, scale, tile, overlap, tiles_x, tiles_y;
  float u8_scale;
  int row_begin, row_end;  // output rows [row_begin, row_end) are written (a rank's band of a sharded frame)
};

// fused proj + residual + LN2 + fc1 + GELU + fc2 + residual (+ LN_next) of one Swin block (bf16, padded 192/384).
// LN2's affine is NOT applied by the kernel: W1 / b1 must carry it (W1[n][k] * gamma2[k], b1 + W1 beta2).
struct MlpFusedArgs {
  const void* o;  // bf16 [M][ld_o] attention output in the padded head layout
  int ld_o;
  int M, C, Hid, CP, HP, QP;
  const void *Wp, *W1, *W2;  // packed bf16 [CP][QP], [HP][CP], [CP][HP]
  const float *bp, *b2, *g3, *be3;  // padded fp32 vectors (g3/be3 may be null).  fc1's bias is not an argument: W1's
                                     // columns C, C+1 hold it as hi + lo bf16 parts (the kernel feeds 1.0 there)
  const float* res;
  int ldres;
  float* out_f32;
  int ld_f32;
  void* out_T;
  int ld_T;
  void* out_ln;
  int ld_ln;
  float eps;
};
int launch_mlp_fused(const MlpFusedArgs& a, cudaStream_t s);

// fused QKV projection + (shifted-)window attention of one Swin block (bf16, 8x8 windows, 6 heads x 32 padded dims)
struct AttnFusedArgs {
  const void* xn;  // bf16 [B*H*W][ld_x]: LayerNorm1 output, pixel order; channels C and C+1 must hold 1.0 (they carry the qkv bias)
  int ld_x;
  void* o;  // bf16 [B*H*W][ld_o], channel = head*32 + j
  int ld_o;
  const void* Whp;       // bf16 [3 pairs][192][192]: rows q h0 h1 | k h0 h1 | v h0 h1 (32 each); q rows x d^-1/2 log2(e);
                         // columns C, C+1 = the qkv bias of the row as hi + lo bf16 parts
  const void* bias_tab;  // bf16 [6 heads][2 shifted copies][15][16]: reversed relative-position bias table x log2(e)
  int B, H, W, shift;
  int C, d;  // un-padded sizes (accounting)
};
int launch_swin_attn_fused(const AttnFusedArgs& a, cudaStream_t s);
constexpr size_t kAttnWhpBytes = 3 * 192 * 192 * 2, kAttnBiasBytes = 6 * 2 * 15 * 16 * 2;
// host-side packing of the two buffers above from the reference parameters (qkv.weight [3C][C], qkv.bias [3C],
// relative_position_bias_table [225][heads]); heads must be 6, C / heads <= 32 and C <= 190 (two pad channels for the bias)
int pack_attn_fused_host(const float* Wqkv, const float* bqkv, const float* table, int C, int heads, void* Whp, void* bias_tab);
constexpr int kAttnOnesChannels = 2;  // norm1's output channels C .. C+1 are 1.0 for the fused attention kernel
int launch_fold_ln_linear(const float* W, const float* b, const float* gamma, const float* beta, float* Wf, float* bf, int N, int K,
                          cudaStream_t s);
int launch_pack_heads(const float* in, void* out, int M, int heads, int d, int DP, int ld, int elem, int round_tf32,
                      cudaStream_t s);

// RCAN channel-attention gate + RCAB residual: out = res + t * sigmoid(W2 relu(W1 mean_hw(t) + b1) + b2)
struct CaArgs {
  const void* t;     // [B][HW][ld]: output of the second conv, fp32 (elem_t 0 / 4) or bf16 (elem_t 2: HAT's bf16 path)
  int elem_t;
  const float* res;  // fp32 [B][HW][ld]: RCAB input
  int ld, B, HW, C, CP, R;
  const float *W1, *b1, *W2, *b2;  // [R][C], [R], [C][R], [C] fp32
  float* partial;                  // scratch [B][nsplit][C]
  int nsplit;
  float* out_f32;  // [B][HW][ld]
  void* out_T;     // T-typed copy for the next conv
  int ld_T, elem, round_tf32;
  float scale;  // out = res + t * gate * scale (1 for RCAN, conv_scale for HAT's CAB)
  // training forward only (null otherwise): the gate MLP's intermediates, kept for launch_channel_attention_bwd
  float *save_pool, *save_hid, *save_gate;  // fp32 [B][C], [B][R], [B][C]
};
int launch_channel_attention(const CaArgs& a, cudaStream_t s);

// SwinFIR's FourierUnit (swinfir.py:9-34), k_fft.cu: torch.fft.rfftn / irfftn over (H, W) with norm "ortho" as direct DFTs.
// Complex rows hold the real parts in columns [0, c) and the imaginary parts in [c, 2c).  tmp: scratch like `spec`.
int launch_rfft2(const float* in, int ld_in, float* tmp, float* spec, int ld_spec, int B, int H, int W, int c, int rtf32, cudaStream_t s);
// out = irfft2(spec) + add (T-typed, [B*H*W][ld_out])
int launch_irfft2_add(const float* spec, int ld_spec, float* tmp, const float* add, int ld_add, void* out, int ld_out, int B, int H, int W,
                      int c, int elem, int rtf32, cudaStream_t s);

// HAN (han.py:12-52).  stack: 11 fp32 planes [B*HW][ld], `plane` elements apart (plane 0 = newest body output).
// launch_han_lam: layer attention, out = T-typed [B*HW][ld_out] with column n * C + c; energy = scratch [B][66] doubles.
int launch_han_lam(const float* stack, size_t plane, int ld, int B, int HW, int C, double* energy, const float* gamma, void* out,
                   int ld_out, int elem, int rtf32, cudaStream_t s);
// launch_han_csam: x fp32 [B*H*W][ld] -> out = x * (gamma * sigmoid(conv3d(x))) + x, T-typed [B*H*W][ld_out]
int launch_han_csam(const float* x, int ld, int B, int H, int W, int C, const float* w27, const float* bias, const float* gamma,
                    void* out, int ld_out, int elem, int rtf32, cudaStream_t s);

// HAN training.  launch_han_lam_bwd: dout = fp32 [B*HW][ld_d] gradient of the layer attention's output (column n * C + c),
// energy = what launch_han_lam left behind; D [B][121] doubles and coef [B][2][121] floats are scratch; dgamma (1 float) and dstack (11 fp32 planes like `stack`) are outputs.
int launch_han_lam_bwd(const float* dout, int ld_d, const float* stack, size_t plane, int ld, int B, int HW, int C, const double* energy,
                       const float* gamma, double* D, float* coef, float* dgamma, float* dstack, cudaStream_t s);
// launch_han_csam_bwd: g = dL/d(csa output) (fp32, ld_g), acc_in = gradient already at x from elsewhere; out / out_bf = total
// gradient at x; scal [29] = dgamma, dbias, dW[27]; dpre / dxd = scratch like x.
int launch_han_csam_bwd(const float* x, int ld, const float* g, int ld_g, const float* acc_in, int B, int H, int W, int C, const float* w27,
                        const float* bias, const float* gamma, float* dpre, float* dxd, float* scal, float* out, void* out_bf,
                        cudaStream_t s);

// backward of out = res + t * sigmoid(W2 relu(W1 mean_hw(t) + b1) + b2) (common.py:156-170 inside rcan.py:21-24):
//   dt = G * gate + W1^T dz1 / HW,  dz1 = relu'(.) * W2^T dz2,  dz2 = gate (1 - gate) * sum_hw G t   (+ the four parameter gradients)
struct CaBwdArgs {
  const float* G;   // fp32 [B][HW][ld]: dL/d(out) (also dL/d(res): the caller keeps it)
  const void* t;    // bf16 [B][HW][ld]: saved output of the second conv
  int ld, B, HW, C, CP, R;
  const float *W1, *W2;                 // fp32 [R][C], [C][R] (current values)
  const float *pool, *hid, *gate;       // saved by the training forward
  float* partial;                       // scratch [B][nsplit][C]
  int nsplit;
  float* dpool;                         // scratch [B][C]
  float *dW1, *db1, *dW2, *db2;         // parameter gradients (PyTorch layouts, overwritten; any may be null)
  void* dt;                             // bf16 [B][HW][ld_dt]
  int ld_dt;
};
int launch_channel_attention_bwd(const CaBwdArgs& a, cudaStream_t s);

// weight gradient of a conv3x3 / linear layer (k_wgrad_tc.cu): dWp[n][tap][c] += alpha * sum_p dY[p][n] X[p+off(tap)][c]
struct WgradArgs {
  const void* dY;  // bf16 [B*H*W][ldy]
  int ldy;
  const void* X;  // bf16 [B*H*W][ldx]
  int ldx;
  int B, H, W, M;
  int taps;
  int NoutP, CinP;  // padded widths (multiples of 64) of dY / X that take part
  float* dWp;       // fp32 [NoutP][taps][CinP], accumulated into (zero it first)
  float* dBp;       // fp32 [NoutP] or null: packed bias gradient alpha * sum_p dY[p][n], accumulated into (zero it first) --
                    // the column sums of the dY tiles that stream through shared memory anyway, by the otherwise idle epilogue warps
  float alpha;
  int N_alg, K_alg;  // un-padded widths (accounting)
};
int launch_wgrad_tc(const WgradArgs& a, cudaStream_t s);

// backward of the window-attention core (k_attn_bwd.cu), bf16, 8x8 windows
struct AttnBwdArgs {
  const void* qkv;  // bf16 [B*H*W][ld_qkv]: q | k | v, each [heads][DP], q pre-scaled (as the forward reads it)
  int ld_qkv, QP;
  const void* d_o;  // bf16 [B*H*W][ld_do]: gradient of the attention output, channel = head*DP + j
  int ld_do;
  void* dqkv;         // bf16 [B*H*W][ld_qkv]: gradient of qkv (w.r.t. the pre-scaled q); pad lanes must be pre-zeroed
  const float* bias;  // [heads][225]
  float* dB;          // scratch fp32 [heads][64][64]
  float* dtable;      // fp32 [225][heads] = d(relative_position_bias_table), overwritten (or null)
  int B, H, W, shift, heads, d, DP;
};
int launch_attn_bwd(const AttnBwdArgs& a, cudaStream_t s);

// row / column maps between a PyTorch nn.Linear [N][K] and its padded pack (model.cu finalize_swinir)
struct LinMap {
  int mode;  // 0: identity; 1: qkv rows n = part*C + h*d + j -> part*QP + h*DP + j (q rows scaled); 2: proj columns k = h*d + j -> h*DP + j
  int C, d, DP, QP;
  float qscale;
};
// Second stages of the two-stage reductions of a backward pass can be DEFERRED: every first stage writes its strips into
// its own slice of a pool, and one batched kernel at the end of the backward finishes all of them (hundreds of 5-15 us
// latency-bound launches become one).  dr == nullptr keeps the immediate behaviour.
struct RedEntry {  // out[dst(n)] = scale(n) * sum_b partial[b * stride + col(n)],  n < N
  const float* partial;
  float* out;
  int strips, stride, N;
  int mode;  // 0: conv bias (dst = pixel-shuffle source row, scale = alpha); 1: linear bias (LinMap); 2: plain
  int Cout, ps_r;
  float alpha;
  LinMap map;
};
struct DeferredRed {
  float* pool = nullptr;  // device scratch for the strips
  size_t pool_floats = 0, used = 0;
  RedEntry* dev = nullptr;  // device copy of the entries
  int cap = 0;
  RedEntry* host = nullptr;  // host entries (owned by the caller, >= cap)
  int n = 0;
};
int launch_deferred_reductions(DeferredRed* dr, cudaStream_t s);
// One launch re-packs every parameter of a model (train_forward): conv / linear packs (forward + dgrad operands), plain
// vector copies (LayerNorm affine, first-conv weights) and the bias-table transposition, driven by a device table.
struct PackEntry {
  const float* W;   // source (PyTorch layout)
  const float* b;   // bias source or null
  void* Wf;         // forward pack (bf16) / copy destination (fp32) / transposed table (fp32)
  float* bf;        // packed bias or null
  void* Wd;         // dgrad pack (bf16) or null
  int kind;         // 0 conv [N][K][taps], 1 linear [N][K] with LinMap, 2 plain copy of N floats, 3 table [N][K] -> [K][N]
  int N, K, NP, KP, taps, ps_r;
  LinMap map;
};
int launch_pack_batched(const PackEntry* host, PackEntry* dev, int n, cudaStream_t s);
// ... and one launch un-packs every packed fp32 weight gradient into the PyTorch layouts at the end of the backward
// (PackEntry re-used: W = packed gradient, Wf = destination, kind 0 conv / 1 linear)
int launch_unpack_batched(const PackEntry* host, PackEntry* dev, int n, cudaStream_t s);
// LayerNorm backward (one warp per row): G_out = G_in + dLN(x; dy, gamma), dgamma += sum dy*xhat, dbeta += sum dy
struct LnBwdArgs {
  const float* x;  // fp32 [M][ldx]: LayerNorm input
  int ldx;
  const void* dy;  // gradient w.r.t. the LayerNorm output, [M][ld_dy], elem_dy = 2 (bf16) or 4 (fp32)
  int ld_dy, elem_dy;
  const float* gamma;  // [C]
  const float* Gin;    // fp32 [M][ldg] residual-stream gradient to add, or null
  float* Gout;         // fp32 [M][ldg]
  void* Gb;            // bf16 copy of Gout [M][ldg], or null
  const float* gb_scale;  // optional per-sample factor of the bf16 copy only (stochastic depth of the next branch): [M / rows_per_scale]
  int rows_per_scale;
  int ldg;
  int M, C, CP;
  float eps;
  float *dgamma, *dbeta;  // [C], overwritten, or null
  float* gb_colsum;       // [C] or null: column sums of the values written to Gb = the bias gradient of the Linear layer that
                          // consumes Gb as its output gradient (needs dgamma and a DeferredRed)
  float* partial;         // scratch, >= kTrainPartialFloats floats (per-CTA partial sums)
};
int launch_ln_bwd(const LnBwdArgs& a, cudaStream_t s, DeferredRed* dr = nullptr);
int launch_colsum_map(const void* dY, int elem, int ld, int M, int NP, int N, const LinMap& map, float* out, float* partial,
                      cudaStream_t s, DeferredRed* dr = nullptr);
constexpr size_t kTrainPartialFloats = (size_t)592 * 2304;  // column-sum strips x widest packed row
int launch_input_nhwc64(const float* x, void* out, int B, int h, int w, int Hp, int Wp, float scale, const float* shift3,
                        cudaStream_t s);
int launch_grad_nhwc64(const float* dy, void* out, int B, int ch, int cw, int Hs, int Ws, float scale, cudaStream_t s);
// out[i] = bf16(in[i] * scale[i / elems_per_scale])
int launch_scale_to_bf16(const float* in, const float* scale, size_t elems_per_scale, void* out, size_t n, cudaStream_t s);

// k_train.cu: on-device (re)packing of the fp32 master parameters, gradient unpacking, small backward pieces
int launch_unpack_wgrad(const float* dWp, float* grad, int Cout, int Cin, int KP, int taps, int ps_r, cudaStream_t s);
int launch_colsum(const void* dY, int elem, int ld, int M, int NP, int Cout, int ps_r, float alpha, float* out, float* partial,
                  cudaStream_t s, DeferredRed* dr = nullptr);
int launch_unshuffle(const void* in, void* out, int B, int H, int W, int C, int r, int ld_in, cudaStream_t s);
int launch_nchw3_to_nhwc64(const float* in, void* out, int B, int H, int W, float scale, const float* shift3, cudaStream_t s);
int launch_add_inplace(float* a, const float* b, void* out_bf, size_t n, cudaStream_t s);
// dL/dx from the gradient at the first conv's output (fp32 [B*Hp*Wp][ldg]); dx fp32 NCHW [B][3][h][w], overwritten
int launch_conv_first_dgrad(const float* G, int ldg, const float* Wc, int C, int B, int h, int w, int Hp, int Wp, float scale,
                            float* dx, cudaStream_t s);

int launch_gemm_simt(const GemmArgs& g, cudaStream_t s);
int launch_gemm_tc(const GemmArgs& g, int elem, cudaStream_t s);
int launch_attn_simt(const AttnArgs& a, cudaStream_t s);
int launch_attn_oca(const AttnArgs& a, cudaStream_t s);
int launch_attn_mma(const AttnArgs& a, cudaStream_t s);
int launch_attn_flash(const AttnArgs& a, int oca, cudaStream_t s);  // bf16, large windows (HAT): k_attn_flash.cu
// tcgen05 / TMEM form of the two HAT attention cores (k_attn_tc.cu): 16x16 windows (shift 0 / 8), 24x24 key windows, DP = 32, bf16
int launch_attn_tc(const AttnArgs& a, int oca, cudaStream_t s);
int launch_layernorm(const LnArgs& a, cudaStream_t s);
int launch_conv_first(const ConvFirstArgs& a, cudaStream_t s);
int launch_conv_last(const ConvLastArgs& a, cudaStream_t s);
int launch_blend(const BlendArgs& a, cudaStream_t s);
int launch_shuffle_finish(const float* in, int ld, int B, int Hp, int Wp, int scale, int ch, int cw, const float* shift3,
                          float out_scale, float u8_scale, float* out_f32, uint8_t* out_u8, cudaStream_t s);
// fp32 [rows][cols] -> T [rows][ld] (zero pad), optional per-row-block scale; used by op-level tests
int launch_pack_rows(const float* in, int rows, int cols, void* out, int ld, int elem, int round_tf32, cudaStream_t s);
int launch_unpack_rows(const void* in, int ld, int elem, float* out, int rows, int cols, cudaStream_t s);
int launch_fill_zero(void* p, size_t bytes, cudaStream_t s);
int launch_repack_qkv(const float* qkv, void* out, int M, int C, int heads, int DP, int QP, float qscale, int elem,
                      int round_tf32, cudaStream_t s);
int launch_unpack_heads(const void* o, int ld, int elem, float* out, int M, int heads, int d, int DP, cudaStream_t s);
int launch_nchw_to_nhwc(const float* in, void* out, int B, int C, int H, int W, int ld, int elem, int round_tf32,
                        cudaStream_t s);
int launch_nhwc_to_nchw(const void* in, int ld, int elem, float* out, int B, int C, int H, int W, cudaStream_t s);

}  // namespace ssr
