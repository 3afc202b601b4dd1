/*
 * libssr_b200 -- C ABI of the B200-native (sm_100a) super-resolution hot path.
 *
 * The reference (veritross/studiosr) is pure PyTorch and has no FFI of its own
 * (SURVEY.md §8b); the boundary it fixes is the Python class API of
 * `studiosr.models`.  This header is the C-ABI that sits directly beneath the
 * drop-in classes in studiosr_b200/models/ and is what any other host (C++,
 * ctypes, cgo, JNI ...) would bind.  Each entry point names the reference
 * interface it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - plain C types only; every pointer is either a HOST pointer or a CUDA DEVICE
 *     pointer as documented per argument; no torch types, no hidden allocations of
 *     activation memory (the caller owns inputs, outputs and the workspace).
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as
 *     void*); nothing synchronises the host unless documented (the *_host entry).
 *   - return value: 0 = success, negative = SSR_E_* ; the message is available from
 *     ssr_last_error() (thread-local).
 *   - there is no CPU fallback: on a device that is not sm_100 every compute entry
 *     returns SSR_E_ARCH.
 */
#ifndef SSR_B200_H
#define SSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSR_VERSION 100

enum {
  SSR_OK = 0,
  SSR_E_INVALID = -1,   /* bad argument / shape / unknown parameter name */
  SSR_E_ARCH = -2,      /* device is not sm_100 (B200) */
  SSR_E_CUDA = -3,      /* CUDA runtime / driver error (message has the detail) */
  SSR_E_STATE = -4,     /* model not finalised / missing parameter */
  SSR_E_WORKSPACE = -5  /* workspace too small */
};

enum { SSR_ARCH_SWINIR = 0, SSR_ARCH_EDSR = 1, SSR_ARCH_RCAN = 2, SSR_ARCH_HAT = 3, SSR_ARCH_HAN = 4,
       SSR_ARCH_SWINFIR = 5 /* SwinIR's config; every RSTB conv and conv_after_body is an SFB (swinfir.py:68-114) */ };

/* arithmetic of the contractions */
enum {
  SSR_PREC_FP32 = 0, /* CUDA-core fp32 FMA (bit-faithful to the reference's fp32 semantics up to summation order) */
  SSR_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 activations, fp32 accumulate */
  SSR_PREC_BF16 = 2, /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate, fp32 residual stream / LN / softmax */
  SSR_PREC_TF32X3 = 3 /* tcgen05 kind::tf32, every operand split into a tf32 head + tail and three MMAs per k-step (a_hi w_hi +
                         a_lo w_hi + a_hi w_lo): fp32-level accuracy (~1e-6) on the tensor cores; fp32 activations, fp32
                         attention / LN / softmax.  Model entry points only (the ssr_op_* calls take 0..2). */
};

/* padding applied before the network, i.e. which reference branch of SwinIR.forward is mirrored */
enum {
  SSR_PAD_EVAL = 0, /* swinir.py:249-255 check_image_size_for_eval: always pads, edge-inclusive mirror */
  SSR_PAD_TRAIN = 1 /* common.py:277-282 check_image_size: reflect pad to the next multiple */
};

#define SSR_MAX_LAYERS 16

/* Mirrors the constructor arguments of studiosr.models.SwinIR (swinir.py:259-274) and
 * studiosr.models.EDSR (edsr.py:13-21) and studiosr.models.RCAN (rcan.py:40-49). */
typedef struct ssr_model_config {
  int arch;      /* SSR_ARCH_* */
  int precision; /* SSR_PREC_* */
  int scale;     /* 2, 3, 4 or 8 */
  int n_colors;  /* 3 */
  float img_range;
  /* SwinIR */
  int embed_dim;
  int n_layers;
  int depths[SSR_MAX_LAYERS];
  int num_heads[SSR_MAX_LAYERS];
  int window_size;
  float mlp_ratio;
  int upsampler; /* 0 = "pixelshuffle", 1 = "pixelshuffledirect" */
  /* EDSR */
  int n_feats;
  int n_resblocks;
  float res_scale;
  /* RCAN (n_feats, n_resblocks as above) */
  int n_resgroups;
  int reduction;
  /* HAT (hat.py:389-406; embed_dim, depths, num_heads, window_size, mlp_ratio as SwinIR) */
  int compress_ratio;
  int squeeze_factor;
  float conv_scale;
  float overlap_ratio;
} ssr_model_config;

typedef struct ssr_model ssr_model_t;

int ssr_version(void);
const char* ssr_last_error(void);
/* 0 if `device` (CUDA ordinal) is an sm_100 part, SSR_E_ARCH otherwise. */
int ssr_device_check(int device);

/* ---- model life cycle: replaces nn.Module construction + load_state_dict ------------------
 * ssr_model_set_param takes the reference's state_dict entry `name` (e.g.
 * "layers.0.residual_group.blocks.1.attn.qkv.weight") as a HOST fp32 array in PyTorch layout
 * ([out,in] linear, [out,in,kh,kw] conv).  Integer buffers (relative_position_index) are
 * derived internally and need not be passed.  ssr_model_finalize re-packs everything into the
 * padded K-major device layouts the kernels consume and uploads it; it may be
 * called again after parameters change. */
int ssr_model_create(const ssr_model_config* cfg, int device, ssr_model_t** out);
int ssr_model_set_param(ssr_model_t* m, const char* name, const float* host_data, int64_t numel);
int ssr_model_finalize(ssr_model_t* m);
void ssr_model_destroy(ssr_model_t* m);

/* ---- forward: replaces SwinIR.forward (swinir.py:353-372) / EDSR.forward (edsr.py:39-48) ----
 * x: DEVICE fp32 NCHW [B, n_colors, H, W];  y: DEVICE fp32 NCHW [B, n_colors, scale*H, scale*W].
 * pad_mode selects the eval / train padding branch (ignored by EDSR). */
size_t ssr_model_workspace_bytes(const ssr_model_t* m, int B, int H, int W, int pad_mode);
int ssr_model_forward(ssr_model_t* m, const float* x, float* y, int B, int H, int W, int pad_mode,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- Model.inference (common.py:36-48) on device buffers ----------------------------------
 * img: DEVICE uint8 HWC [H, W, 3] -> out: DEVICE uint8 HWC [scale*H, scale*W, 3];
 * /255 (img_range == 1), eval-mode forward, *255, round-half-even, clip, all fused into the
 * first / last kernels.  Batch of `B` independent images of the same size. */
int ssr_model_upscale_u8(ssr_model_t* m, const uint8_t* img, uint8_t* out, int B, int H, int W,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- tiled full-frame inference (BASELINE.json config 5; new capability, the reference has no
 * tiler): frame DEVICE uint8 HWC [H, W, 3] -> DEVICE uint8 HWC [scale*H, scale*W, 3].  Tiles of
 * tile x tile LR pixels at stride tile-overlap (last clamped to the border) run as one batch
 * through the eval forward; outputs are blended with separable linear ramps (oracle:
 * oracle/sr_oracle.py:tiled_upscale). */
int ssr_tiled_num_tiles(int H, int W, int tile, int overlap);
size_t ssr_model_tiled_workspace_bytes(const ssr_model_t* m, int H, int W, int tile, int overlap, int chunk_tiles);
int ssr_model_upscale_tiled_u8(ssr_model_t* m, const uint8_t* frame, uint8_t* out, int H, int W, int tile,
                               int overlap, int chunk_tiles, void* workspace, size_t workspace_bytes, void* stream);
/* The two halves of the call above, for sharding ONE frame over several GPUs (tiles are independent, the
 * only exchange is the gather of their outputs).  A rank runs
 *   ssr_model_tiles_u8        tiles [tile_begin, tile_end) of the row-major tile list (tile_end < 0: to the end) ->
 *                             tiles_out, DEVICE fp32 [tile_end - tile_begin][3][th*scale][tw*scale] (th = min(tile, H), tw
 *                             likewise; ssr_tiled_tile_elems floats per tile), chunk_tiles tiles per network pass (0 = all);
 *   (the caller all-gathers the tile outputs, e.g. ncclAllGather into the full [n_tiles] list)
 *   ssr_model_blend_tiles_u8  tiles = the FULL row-major list -> rows [row_begin, row_end) of the output frame (`out` is the
 *                             frame base; row_end < 0: to the last row).  Blending is a gather per output pixel, so a band
 *                             is bit-identical to the same rows of the whole-frame call.
 * studiosr_b200/sharding.py drives this protocol over torch.distributed. */
size_t ssr_tiled_tile_elems(const ssr_model_t* m, int H, int W, int tile);
size_t ssr_model_tiles_workspace_bytes(const ssr_model_t* m, int H, int W, int tile, int max_tiles_per_pass);
int ssr_model_tiles_u8(ssr_model_t* m, const uint8_t* frame, float* tiles_out, int H, int W, int tile, int overlap,
                       int tile_begin, int tile_end, int chunk_tiles, void* workspace, size_t workspace_bytes, void* stream);
int ssr_model_blend_tiles_u8(ssr_model_t* m, const float* tiles, uint8_t* out, int H, int W, int tile, int overlap,
                             int row_begin, int row_end, void* stream);
/* Same, HOST buffers (pinned or pageable): H2D copy, compute, D2H copy, stream synchronise. */
int ssr_model_upscale_tiled_u8_host(ssr_model_t* m, const uint8_t* frame_host, uint8_t* out_host, int H, int W,
                                    int tile, int overlap, int chunk_tiles, void* workspace, size_t workspace_bytes,
                                    void* stream);

/* ---- training step (trainer.py:97-105: `out = model(x)` ... `loss.backward()`), bf16 tensor-core path ----------
 * The reference differentiates the model with torch autograd; these two calls are that forward/backward pair for the
 * whole model.  The fp32 master parameters and their gradients stay in caller-owned DEVICE memory:
 *   ssr_model_train_bind     names the state_dict entries (weights AND biases, in any fixed order) once; the model must
 *                            have been finalised (ssr_model_set_param + ssr_model_finalize) so the layouts exist.
 *   ssr_model_train_forward  params[i] = DEVICE fp32 pointer of bound entry i (PyTorch layout).  Re-packs the weights on
 *                            the device (forward and transposed/rotated dgrad operands), runs the training-mode forward
 *                            x [B,3,H,W] -> y [B,3,sH,sW] (fp32 NCHW) and keeps every GEMM operand in `workspace`.
 *   ssr_model_train_backward dy = dL/dy (DEVICE fp32 NCHW).  grads[i] = DEVICE fp32 buffer for dL/d(param i), overwritten
 *                            (NULL: not wanted, e.g. frozen entries).  `workspace` must be the buffer, untouched, that
 *                            the matching train_forward used.
 *   ssr_model_train_input_grad  dL/dx (DEVICE fp32 NCHW [B,3,H,W], overwritten) of the step whose train_backward ran last on
 *                            this handle (the Trainer never asks for it; autograd users with x.requires_grad do): the first
 *                            conv's data gradient through the input normalisation and the training-mode reflect pad.  The
 *                            workspace of that backward must still be untouched.
 *   drop_scale               stochastic depth of SwinIR (timm DropPath, swinir.py:137,171-172): DEVICE fp32
 *                            [2 * n_blocks][B], entry [2k][b] / [2k+1][b] = the factor (0 or 1/keep_prob) applied to sample b's
 *                            attention / MLP branch of block k (blocks in forward order); NULL = no stochastic depth.  The host
 *                            draws it (the Python module consumes torch's RNG exactly as the reference's DropPath calls do);
 *                            the same array must be passed to the matching backward.
 * Forward/backward pairs on one handle must not interleave with another forward on the same handle. */
int ssr_model_train_bind(ssr_model_t* m, int n, const char* const* names, const int64_t* numels);
size_t ssr_model_train_workspace_bytes(const ssr_model_t* m, int B, int H, int W);
int ssr_model_train_forward(ssr_model_t* m, const float* const* params, const float* drop_scale, const float* x, float* y, int B,
                            int H, int W, void* workspace, size_t workspace_bytes, void* stream);
int ssr_model_train_backward(ssr_model_t* m, const float* dy, const float* drop_scale, float* const* grads, int B, int H, int W,
                             void* workspace, size_t workspace_bytes, void* stream);
int ssr_model_train_input_grad(ssr_model_t* m, float* dx, int B, int H, int W, void* stream);

/* ---- the rest of the Trainer step on flat fp32 DEVICE buffers (trainer.py:102-109,133-139) -----------------------------
 * ssr_l1_loss    nn.L1Loss (trainer.py:45): *loss = mean |out - y| over n elements and, when dout != NULL, the backward seed
 *                dout = sign(out - y) / n in the same pass (deterministic two-stage sum).  workspace: ssr_l1_loss_workspace_bytes().
 * ssr_adam_step  torch.optim.Adam's update (trainer.py:133-139; L2 weight decay, no amsgrad) for ONE flat parameter vector of n
 *                elements in one launch: p, m (exp_avg), v (exp_avg_sq) updated in place from g * grad_scale; `step` is the 1-based
 *                step count, `lr` the current learning rate (MultiStepLR stays on the host: it only changes this scalar).  The
 *                hyper-parameters are doubles (python floats) because torch derives its fp32 constants from them in double; the
 *                update is bit-identical to torch.optim.Adam(foreach=False).
 * All pointers must be 16-byte aligned. */
size_t ssr_l1_loss_workspace_bytes(void);
int ssr_l1_loss(const float* out, const float* y, int64_t n, float* loss, float* dout, void* workspace, size_t workspace_bytes,
                void* stream);
int ssr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                  double weight_decay, int64_t step, float grad_scale, void* stream);

/* ---- device-side evaluation metric and training augmentation -------------------------------------------------
 * ssr_psnr_mse_u8       compute_psnr (utils/metrics.py:36-49) up to its last line: im1 [h1,w1,3], im2 [h2,w2,3] DEVICE uint8 HWC are
 *                       cropped to the common size (trailing rows / columns), `crop_border` pixels are cut from every side, the
 *                       squared error of the BT.601 luma (y_only != 0, metrics.py:11-17) or of the three channels is averaged
 *                       into *dev_mse (DEVICE double); PSNR = 20 log10(255 / sqrt(mse)), inf for mse == 0.
 * ssr_augment_pairs_u8  the training transform of dataset.py:50-58 + array2tensor (transforms.py:8-68) for n_pairs (LR, HR) pairs
 *                       in one launch: x_out DEVICE fp32 [n_pairs,3,size,size], y_out [n_pairs,3,size*scale,size*scale], values / 255.
 *                       dev_pairs: DEVICE array of ssr_aug_pair; the HOST draws xs, ys and the flags (the reference uses python's
 *                       `random`, so its seeded sequence can be reproduced draw for draw). */
typedef struct ssr_aug_pair {
  const uint8_t* lq; /* DEVICE uint8 HWC LR image */
  const uint8_t* gt; /* DEVICE uint8 HWC HR image (scale x larger) */
  int lq_w, gt_w;    /* image widths in pixels */
  int xs, ys;        /* crop origin in the LR image (paired_random_crop, transforms.py:8-22) */
  int flags;         /* bit 0: fliplr, bit 1: flipud, bit 2: rot90 (np.rot90, counter-clockwise) -- applied in that order */
  int pad;
} ssr_aug_pair;
size_t ssr_psnr_workspace_bytes(void);
int ssr_psnr_mse_u8(const uint8_t* im1, int h1, int w1, const uint8_t* im2, int h2, int w2, int crop_border, int y_only,
                    double* dev_mse, void* workspace, size_t workspace_bytes, void* stream);
int ssr_augment_pairs_u8(const void* dev_pairs, int n_pairs, int size, int scale, float* x_out, float* y_out, void* stream);

/* Checksum of n fp32 DEVICE tensors in two launches (the Python boundary keys its packed-weight cache on it: writes through
 * `tensor.data` do not bump torch's version counters).  dev_ptrs / dev_numels: DEVICE int64 [n] (addresses, element counts);
 * dev_partial: DEVICE double [2n] scratch; dev_out2: DEVICE double [2]. */
int ssr_tensors_checksum(const int64_t* dev_ptrs, const int64_t* dev_numels, int n, double* dev_partial, double* dev_out2, void* stream);

/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
int64_t ssr_launch_count(void);
/* A host that captured calls of this library into a CUDA graph reports each replay here (kernels = the count's increase during
 * the capture), so that ssr_launch_count keeps meaning "kernels of this library that ran". */
void ssr_note_graph_replay(int64_t kernels);
/* Per-launch device timing for the roofline report: between begin and end every kernel launch is
 * bracketed by CUDA events on its own stream.  ssr_profile_end synchronises the device and writes a
 * JSON object {"<kernel class>": {"launches", "ms", "flops", "bytes"}} (algorithmic FLOPs / bytes
 * with un-padded dimensions) into `json`.  Not for use inside timed regions. */
int ssr_profile_begin(void);
int ssr_profile_end(char* json, size_t capacity);

/* ---- op-level entry points (used by the parity tests; same kernels the model path runs) -----
 * All tensors DEVICE fp32 in the reference's own layouts; the library packs / pads internally
 * into scratch carved from `workspace`. */

/* nn.Linear with the fused epilogue of the model path:
 *   v = act(x[M,K] @ W[N,K]^T + b) (+ res[M,N]);  y = v;  y_ln = LayerNorm(v) * ln_w + ln_b (eps 1e-5)
 * act: 0 none, 1 relu, 2 leaky-relu(0.01), 3 exact-erf GELU.  res, ln_w/ln_b, y, y_ln may be NULL.
 * Replaces nn.Linear / Mlp / nn.LayerNorm call sites swinir.py:80,103,151,172 and common.py:184-194. */
int ssr_op_linear(int precision, const float* x, const float* W, const float* b, const float* res, int act,
                  const float* ln_w, const float* ln_b, float* y, float* y_ln, int M, int K, int N, void* workspace,
                  size_t workspace_bytes, void* stream);
/* nn.Conv2d(Cin, Cout, 3, 1, 1) on NCHW fp32 (swinir.py:241,316,321-326; common.py:104,124-153):
 *   v = act(conv(x) + b) * alpha (+ res);  optional nn.PixelShuffle(ps_r) folded into the store, in which
 * case y is [B, Cout/ps_r^2, H*ps_r, W*ps_r] (ps_r = 0/1: none; exclusive with res). */
int ssr_op_conv3x3(int precision, const float* x, const float* W, const float* b, const float* res, float* y, int B,
                   int Cin, int Cout, int H, int Wd, int act, float alpha, int ps_r, void* workspace,
                   size_t workspace_bytes, void* stream);
/* Weight / bias gradient of nn.Conv2d(Cin, Cout, 3, 1, 1) (taps = 9) or of a 1x1 conv / nn.Linear over pixels (taps = 1) as
 * torch autograd computes it for the reference (trainer.py:104), bf16 tensor-core path (k_wgrad_tc.cu):
 *   dW[co][ci][ky][kx] = alpha * sum_{b,y,x} dy[b,co,y,x] * x[b,ci,y+ky-1,x+kx-1];   db[co] = alpha * sum dy[b,co,y,x]
 * dy [B,Cout,H,W], x [B,Cin,H,W] fp32 NCHW; dW [Cout,Cin,3,3] (or [Cout,Cin]); db [Cout] or NULL. */
int ssr_op_conv3x3_wgrad(const float* dy, const float* x, float* dW, float* db, int B, int Cin, int Cout, int H, int Wd, int taps,
                         float alpha, void* workspace, size_t workspace_bytes, void* stream);
/* Fused tail of a SwinTransformerBlock (swinir.py:103,171-172; common.py:184-194), bf16 tensor-core path:
 *   t1 = o @ Wp^T + bp + res;  h = GELU(LN(t1; g2,be2) @ W1^T + b1);  y = t1 + h @ W2^T + b2;
 *   y_ln = LN(y; g3,be3) (or, when g3 is NULL, y rounded to bf16).  o, res, y: [M, C] fp32; weights in PyTorch
 *   layouts ([C,C], [hidden,C], [C,hidden]).  Only the C=180 / 6 heads / hidden=360 class of shapes. */
int ssr_op_swin_mlp(const float* o, const float* res, const float* Wp, const float* bp, const float* g2, const float* be2,
                    const float* W1, const float* b1, const float* W2, const float* b2, const float* g3, const float* be3,
                    float* y, float* y_ln, int M, int C, int heads, int hidden, void* workspace, size_t workspace_bytes,
                    void* stream);
/* SwinIR (shifted-)window attention core, swinir.py:78-105 minus the two Linear layers, together with
 * the roll / window_partition / window_reverse / calculate_mask around it (swinir.py:154-168,
 * common.py:236-274): qkv [B,H,W,3*C] (q|k|v, each head-major, un-scaled) -> o [B,H,W,C]; q scaling,
 * relative-position bias (table [(2ws-1)^2, heads]) and the shift mask are applied inside the kernel. */
int ssr_op_window_attention(int precision, const float* qkv, const float* bias_table, float* o, int B, int H, int W,
                            int C, int heads, int ws, int shift, void* workspace, size_t workspace_bytes,
                            void* stream);
/* Fused qkv projection + (shifted-)window attention of a SwinTransformerBlock (swinir.py:78-102 and 154-168 without
 * the output projection), bf16 tensor-core path: xn [B,H,W,C] (the LayerNorm1 output) -> o [B,H,W,C]; Wqkv [3C,C],
 * bqkv [3C], bias_table [225, heads] in the reference's layouts; 8x8 windows, shift 0 or 4, C=180 / 6 heads class. */
int ssr_op_swin_attn(const float* xn, const float* Wqkv, const float* bqkv, const float* bias_table, float* o, int B, int H,
                     int W, int C, int heads, int shift, void* workspace, size_t workspace_bytes, void* stream);
size_t ssr_op_workspace_bytes(int64_t max_elems);

#ifdef __cplusplus
}
#endif
#endif /* SSR_B200_H */
