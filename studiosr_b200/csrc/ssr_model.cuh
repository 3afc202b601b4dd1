// Model handle and packed-layer descriptors shared by the inference executor (model.cu) and the training
// executor (train.cu).  Internal; the public ABI is include/ssr_b200.h.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "ssr_internal.cuh"

namespace ssr {
// ---------------------------------------------------------------------------------------------
struct Lin {           // one packed (implicit-)GEMM layer
  size_t w_off = 0;    // arena offset of T [NP][taps*KP]
  size_t b_off = 0;    // arena offset of float [NP]
  int K = 0, KP = 0, N = 0, NP = 0, taps = 1, ps_r = 0;
  int N_alg = 0;       // un-padded output width (accounting)
};
struct LNp {
  size_t g_off = 0, b_off = 0;
};
struct Block {
  LNp norm1, norm2;
  Lin qkv, proj, fc1, fc2;
  size_t bias_off = 0;  // float [heads][(2ws-1)^2]
  size_t whp_off = 0, btab_off = 0;  // fused attention kernel operands (k_swin_attn.cu)
  // HAT: channel attention block of a HAB (hat.py:41-52)
  Lin cab0, cab2;
  size_t ca_w1 = 0, ca_b1 = 0, ca_w2 = 0, ca_b2 = 0;
};
// SwinFIR's SFB (swinfir.py:68-80): SpatialB convs, SpectralTransform 1x1 convs (as Linears), fusion
struct Sfb {
  Lin s0, s2, before, fu, after, fusion;
};
struct Layer {
  std::vector<Block> blocks;
  Lin conv;
  Sfb sfb;  // replaces `conv` when the model is a SwinFIR
  int heads = 0, d = 0, DP = 0, QP = 0;
  Block ocab;  // HAT: overlapping cross-attention block closing the group (hat.py:198-293)
};

}  // namespace ssr

struct ssr_train_state;
using namespace ssr;  // internal header: every includer is library code inside / around namespace ssr

struct ssr_model {
  ssr_model_config cfg;
  int device = 0;
  int elem = 4;  // bytes per activation / weight element
  bool finalized = false;
  std::map<std::string, std::vector<float>> params;
  // packed
  std::vector<uint8_t> host_arena;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  // SwinIR
  int C = 0, CP = 0, HID = 0, HP = 0, QPmax = 0;
  std::vector<Layer> layers;
  LNp pe_norm, final_norm;
  size_t conv_first_w = 0, conv_first_b = 0;
  Lin conv_after_body, conv_before_up;
  std::vector<Lin> up;  // upsample convs
  size_t conv_last_w = 0;
  float conv_last_bias[3] = {0, 0, 0};
  int last_cin = 64;
  Lin last_lin;  // the same conv packed for the tensor-core implicit GEMM (bf16 / tf32 models)
  size_t last_w27 = 0;  // bf16 models with 64 input channels: [32][64] bf16, row = tap * 3 + c (k_conv_last.cu); 0 = not packed
  // EDSR
  int F = 0, FP = 0;
  std::vector<Lin> res_a, res_b;
  Lin body_tail;
  float sub_bias[3] = {0, 0, 0}, add_bias[3] = {0, 0, 0};
  // RCAN: res_a / res_b hold the two convs of every RCAB (group-major), grp_tail the conv closing each group
  std::vector<Lin> grp_tail;
  struct CaP {
    size_t w1 = 0, b1 = 0, w2 = 0, b2 = 0;  // fp32 [R][C], [R], [C][R], [C]
  };
  std::vector<CaP> ca;
  // HAN: the RCAN trunk plus last_conv (11 F -> F), last (2 F -> F), the CSAM Conv3d (27 weights + bias) and the two gammas
  Sfb sfb_after_body;  // SwinFIR: conv_after_body
  bool sfb = false;
  Lin han_last_conv, han_last;
  size_t csa_w = 0, csa_b = 0, csa_gamma = 0, la_gamma = 0;

  ssr_train_state* train = nullptr;  // training executor state (train.cu), owned

  template <typename T>
  T* dev(size_t off) const { return reinterpret_cast<T*>(arena + off); }
};

namespace ssr {
struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<uint8_t*>(p)) {}
  void* take(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};
struct InputSpec {
  const void* in;
  int in_u8;
  int fh, fw;  // frame / image size
  int tile_mode, tile, stride, tiles_x, tile_begin;
};
struct OutputSpec {
  float* out_f32;
  uint8_t* out_u8;
};
// model.cu
size_t arena_alloc(ssr_model* m, size_t bytes);
void upsampler_plan(int scale, std::vector<int>* rs);
int run_gemm(const ssr_model* m, GemmArgs& g, cudaStream_t s);
GemmArgs gemm_base(const ssr_model* m, const Lin& L, const void* A, int lda, int B, int H, int W);
int check_ready(ssr_model* m);
// train.cu
void train_state_destroy(ssr_model* m);
}  // namespace ssr
