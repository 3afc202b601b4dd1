// Training executor: forward with saved activations + backward (data and weight gradients) behind
// the C ABI, bf16 tensor-core path.  Replaces what `loss.backward()` (trainer.py:104) runs through torch autograd for
// the model: the reference has no hand-written backward, so every formula below is the adjoint of the cited forward.
// The fp32 master parameters stay in PyTorch-owned device memory; each train_forward re-packs them on the device.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ssr_model.cuh"

using namespace ssr;

namespace ssr {

struct ConvT {       // a trainable conv3x3 executed by the implicit-GEMM kernels
  const Lin* fwd;    // forward pack (m->arena)
  int wi = -1, bi = -1;  // bound-parameter indices
  int Cout = 0, Cin = 0;
  size_t dg_off = 0;     // dgrad pack in train->arena2: bf16 [KP][9*NP]
  size_t dwp_off = 0;    // packed fp32 weight gradient [NP][9][KP] (float offset into the dwp workspace block)
  size_t dbp_off = 0;    // packed fp32 bias gradient [NP] (same block)
};

struct LinT {  // a trainable nn.Linear executed by the GEMM kernels
  const Lin* fwd = nullptr;
  int wi = -1, bi = -1;
  int N = 0, K = 0;  // PyTorch [N][K]
  LinMap map{};
  size_t dg_off = 0;   // dgrad pack in arena2: bf16 [KP][NP]
  size_t dwp_off = 0;  // packed fp32 weight gradient [NP][KP]
  size_t dbp_off = 0;  // packed fp32 bias gradient [NP]
};
struct LnT {
  const LNp* p = nullptr;
  int gi = -1, bi = -1;
};
struct CaT {  // channel-attention gate of an RCAB: bound-parameter indices
  int w1 = -1, b1 = -1, w2 = -1, b2 = -1;
};
struct BlockT {
  LnT n1, n2;
  LinT qkv, proj, fc1, fc2;
  int table_i = -1;
  size_t bias_off = 0;  // head-major table in m->arena
};

}  // namespace ssr

struct ssr_train_state {
  std::vector<std::string> names;
  std::vector<int64_t> numels;
  std::map<std::string, int> index;
  uint8_t* arena2 = nullptr;
  size_t arena2_bytes = 0;
  size_t zero_bias_off = 0;
  std::vector<ConvT> convs;
  size_t dwp_floats = 0;
  size_t red_floats = 0;  // pool for the strips of every deferred reduction of one backward
  int red_entries = 0;
  std::vector<RedEntry> red_host;
  DeferredRed red;
  std::vector<PackEntry> pack_host, unpack_host;
  // EDSR
  int head_w = -1, head_b = -1;
  size_t head_dwp = 0, head_dbp = 0;
  std::vector<int> e_res_a, e_res_b, e_up;  // indices into convs
  int e_body_tail = -1, e_last = -1;
  // RCAN (head / body tail / upsampler / last conv share the EDSR fields)
  std::vector<int> r_a, r_b, r_gt;  // indices into convs: the two convs of every RCAB (group-major), the conv closing each group
  std::vector<ssr::CaT> r_ca;
  int pack_cap = 0;                 // entries the batched (un)pack launches may carry
  // HAN (the RCAN fields plus): convs last_conv / last, bound indices of csa.conv.weight / .bias, csa.gamma, la.gamma
  int h_last_conv = -1, h_last = -1, h_csa_w = -1, h_csa_b = -1, h_csa_g = -1, h_la_g = -1;
  // left behind by the last train_backward for ssr_model_train_input_grad: dL/d(first conv's output), fp32
  const float* dx_G = nullptr;
  int dx_ld = 0, dx_C = 0, dx_B = 0, dx_h = 0, dx_w = 0, dx_Hp = 0, dx_Wp = 0;
  float dx_scale = 1.0f;
  // SwinIR
  std::vector<std::vector<BlockT>> s_blocks;
  std::vector<int> s_conv;  // RSTB convs
  LnT s_pe, s_fin;
  int s_cab = -1, s_cbu = -1;
};

namespace ssr {

void train_state_destroy(ssr_model* m) {
  if (!m->train) return;
  if (m->train->arena2) cudaFree(m->train->arena2);
  delete m->train;
  m->train = nullptr;
}

static int find_idx(const ssr_train_state* t, const std::string& name, int64_t numel) {
  auto it = t->index.find(name);
  if (it == t->index.end()) {
    set_error("train: parameter '%s' was not bound", name.c_str());
    return -1;
  }
  if (t->numels[it->second] != numel) {
    set_error("train: parameter '%s' has %lld elements, expected %lld", name.c_str(), (long long)t->numels[it->second],
              (long long)numel);
    return -1;
  }
  return it->second;
}

static int add_conv(ssr_train_state* t, const std::string& name, const Lin* fwd, int Cout, int Cin, size_t* a2) {
  ConvT c;
  c.fwd = fwd;
  c.Cout = Cout;
  c.Cin = Cin;
  c.wi = find_idx(t, name + ".weight", (int64_t)Cout * Cin * 9);
  c.bi = find_idx(t, name + ".bias", Cout);
  if (c.wi < 0 || c.bi < 0) return -1;
  *a2 = (*a2 + 255) & ~(size_t)255;
  c.dg_off = *a2;
  *a2 += (size_t)fwd->KP * 9 * fwd->NP * 2;
  c.dwp_off = t->dwp_floats;
  t->dwp_floats += (size_t)fwd->NP * 9 * fwd->KP;
  c.dbp_off = t->dwp_floats;
  t->dwp_floats += (size_t)fwd->NP;
  t->red_floats += (size_t)592 * fwd->NP + 64;
  t->red_entries += 1;
  t->convs.push_back(c);
  return (int)t->convs.size() - 1;
}

static int bind_edsr(ssr_model* m) {
  ssr_train_state* t = m->train;
  const ssr_model_config& c = m->cfg;
  size_t a2 = 0;
  t->zero_bias_off = a2;
  a2 += 4096 * 4;
  t->head_w = find_idx(t, "head.0.weight", (int64_t)m->F * 27);
  t->head_b = find_idx(t, "head.0.bias", m->F);
  if (t->head_w < 0 || t->head_b < 0) return SSR_E_STATE;
  t->head_dwp = t->dwp_floats;
  t->dwp_floats += (size_t)m->FP * 9 * 64;
  t->head_dbp = t->dwp_floats;
  t->dwp_floats += (size_t)m->FP;
  t->red_floats += (size_t)592 * m->FP + 64;
  t->red_entries += 1;
  char nm[64];
  for (int i = 0; i < c.n_resblocks; ++i) {
    snprintf(nm, sizeof(nm), "body.%d.body.0", i);
    int ia = add_conv(t, nm, &m->res_a[i], m->F, m->F, &a2);
    snprintf(nm, sizeof(nm), "body.%d.body.2", i);
    int ib = add_conv(t, nm, &m->res_b[i], m->F, m->F, &a2);
    if (ia < 0 || ib < 0) return SSR_E_STATE;
    t->e_res_a.push_back(ia);
    t->e_res_b.push_back(ib);
  }
  snprintf(nm, sizeof(nm), "body.%d", c.n_resblocks);
  t->e_body_tail = add_conv(t, nm, &m->body_tail, m->F, m->F, &a2);
  if (t->e_body_tail < 0) return SSR_E_STATE;
  for (size_t i = 0; i < m->up.size(); ++i) {
    snprintf(nm, sizeof(nm), "tail.0.%d", (int)(2 * i));
    int iu = add_conv(t, nm, &m->up[i], m->up[i].N, m->F, &a2);
    if (iu < 0) return SSR_E_STATE;
    t->e_up.push_back(iu);
  }
  t->e_last = add_conv(t, "tail.1", &m->last_lin, 3, m->F, &a2);
  if (t->e_last < 0) return SSR_E_STATE;
  t->arena2_bytes = a2 + 1024;
  SSR_CUDA(cudaMalloc(&t->arena2, t->arena2_bytes));
  SSR_CUDA(cudaMemset(t->arena2, 0, t->arena2_bytes));
  return SSR_OK;
}

// ---------------------------------------------------------------------------------------------
struct EdsrTrainWs {
  void* xin64;
  float *x0, *r;
  std::vector<void*> rb, tmp;
  void* bt;
  std::vector<void*> hr, ghr;
  void *dy64, *gU;
  float *G, *Gt;
  void *Gb, *Dh;
  float* dwp;
  float* partial;
  float* red_pool;
  RedEntry* red_dev;
  PackEntry* pack_dev;
};

static size_t plan_edsr_train(const ssr_model* m, void* base, int B, int H, int W, EdsrTrainWs* w) {
  Carver c(base);
  const size_t T = (size_t)B * H * W;
  const int FP = m->FP, nb = m->cfg.n_resblocks;
  w->xin64 = c.take(T * 64 * 2);
  w->x0 = (float*)c.take(T * FP * 4);
  w->r = (float*)c.take(T * FP * 4);
  w->rb.resize(nb + 1);
  w->tmp.resize(nb);
  for (int i = 0; i <= nb; ++i) w->rb[i] = c.take(T * FP * 2);
  for (int i = 0; i < nb; ++i) w->tmp[i] = c.take(T * FP * 2);
  w->bt = c.take(T * FP * 2);
  size_t px = T, gu_max = 0;
  w->hr.resize(m->up.size());
  w->ghr.resize(m->up.size());
  for (size_t i = 0; i < m->up.size(); ++i) {
    gu_max = std::max(gu_max, px * (size_t)m->up[i].NP);
    px *= (size_t)m->up[i].ps_r * m->up[i].ps_r;
    w->hr[i] = c.take(px * FP * 2);
    w->ghr[i] = c.take(px * FP * 2);
  }
  w->dy64 = c.take(px * 64 * 2);
  w->gU = c.take(gu_max * 2);
  w->G = (float*)c.take(T * FP * 4);
  w->Gt = (float*)c.take(T * FP * 4);
  w->Gb = c.take(T * FP * 2);
  w->Dh = c.take(T * FP * 2);
  w->dwp = (float*)c.take(m->train->dwp_floats * 4);
  w->partial = (float*)c.take(kTrainPartialFloats * 4);
  w->red_pool = (float*)c.take(m->train->red_floats * 4);
  w->red_dev = (RedEntry*)c.take((size_t)(m->train->red_entries + 8) * sizeof(RedEntry));
  w->pack_dev = (PackEntry*)c.take((size_t)(2 * m->train->red_entries + 16) * sizeof(PackEntry));
  return c.off + 1024;
}

static void push_copy(ssr_train_state* t, const float* src, float* dst, int n) {
  PackEntry e;
  memset(&e, 0, sizeof(e));
  e.W = src;
  e.Wf = dst;
  e.kind = 2;
  e.N = n;
  e.K = 1;
  t->pack_host.push_back(e);
}
static int repack_conv(ssr_model* m, const ConvT& c, const float* const* params, cudaStream_t) {
  const Lin& L = *c.fwd;
  PackEntry e;
  memset(&e, 0, sizeof(e));
  e.W = params[c.wi];
  e.b = params[c.bi];
  e.Wf = m->arena + L.w_off;
  e.bf = m->dev<float>(L.b_off);
  e.Wd = m->train->arena2 + c.dg_off;
  e.kind = 0;
  e.N = c.Cout;
  e.K = c.Cin;
  e.NP = L.NP;
  e.KP = L.KP;
  e.taps = 9;
  e.ps_r = L.ps_r;
  m->train->pack_host.push_back(e);
  return SSR_OK;
}

static int train_forward_edsr(ssr_model* m, const float* const* params, const float* x, float* y, int B, int h, int w, void* ws,
                              size_t ws_bytes, cudaStream_t s) {
  ssr_train_state* t = m->train;
  const ssr_model_config& c = m->cfg;
  EdsrTrainWs W;
  const size_t need = plan_edsr_train(m, ws, B, h, w, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "train workspace %zu B < required %zu B", ws_bytes, need);
  const int FP = m->FP;
  // ---- re-pack the (updated) fp32 master weights: forward + dgrad operand layouts ----
  t->pack_host.clear();
  push_copy(t, params[t->head_w], m->dev<float>(m->conv_first_w), m->F * 27);
  push_copy(t, params[t->head_b], m->dev<float>(m->conv_first_b), m->F);
  for (const ConvT& cv : t->convs) SSR_TRY(repack_conv(m, cv, params, s));
  SSR_TRY(launch_pack_batched(t->pack_host.data(), W.pack_dev, (int)t->pack_host.size(), s));
  // ---- forward (edsr.py:39-48), every GEMM operand kept for the backward ----
  SSR_TRY(launch_nchw3_to_nhwc64(x, W.xin64, B, h, w, 1.0f, m->sub_bias, s));
  {
    ConvFirstArgs a;
    memset(&a, 0, sizeof(a));
    a.in = x;
    a.fh = h;
    a.fw = w;
    a.h = h;
    a.w = w;
    a.Hp = h;
    a.Wp = w;
    a.pad_mode = 2;
    a.B = B;
    a.in_scale = 1.0f;
    for (int i = 0; i < 3; ++i) a.in_shift[i] = m->sub_bias[i];
    a.Wc = m->dev<float>(m->conv_first_w);
    a.bias = m->dev<float>(m->conv_first_b);
    a.Cout = m->F;
    a.out_f32 = W.x0;
    a.ld_f32 = FP;
    a.out_T = W.rb[0];
    a.ld_T = FP;
    a.elem = 2;
    SSR_TRY(launch_conv_first(a, s));
  }
  for (int i = 0; i < c.n_resblocks; ++i) {  // ResBlock (common.py:150-153)
    GemmArgs ga = gemm_base(m, m->res_a[i], W.rb[i], FP, B, h, w);
    ga.act = ACT_RELU;
    ga.out_T = W.tmp[i];
    ga.ld_T = FP;
    SSR_TRY(run_gemm(m, ga, s));
    GemmArgs gb = gemm_base(m, m->res_b[i], W.tmp[i], FP, B, h, w);
    gb.alpha = c.res_scale;
    gb.res = i == 0 ? W.x0 : W.r;
    gb.ldres = FP;
    gb.out_f32 = W.r;
    gb.ld_f32 = FP;
    gb.out_T = W.rb[i + 1];
    gb.ld_T = FP;
    SSR_TRY(run_gemm(m, gb, s));
  }
  {
    GemmArgs g = gemm_base(m, m->body_tail, W.rb[c.n_resblocks], FP, B, h, w);
    g.res = W.x0;
    g.ldres = FP;
    g.out_T = W.bt;
    g.ld_T = FP;
    SSR_TRY(run_gemm(m, g, s));
  }
  const void* cur = W.bt;
  int H = h, Wd = w;
  for (size_t i = 0; i < m->up.size(); ++i) {
    GemmArgs g = gemm_base(m, m->up[i], cur, FP, B, H, Wd);
    g.out_T = W.hr[i];
    g.ld_T = FP;
    SSR_TRY(run_gemm(m, g, s));
    cur = W.hr[i];
    H *= m->up[i].ps_r;
    Wd *= m->up[i].ps_r;
  }
  GemmArgs g = gemm_base(m, m->last_lin, cur, FP, B, H, Wd);
  g.out3_f32 = y;
  g.crop_h = H;
  g.crop_w = Wd;
  for (int i = 0; i < 3; ++i) g.out_shift[i] = m->add_bias[i];
  g.out_scale = 1.0f;
  g.u8_scale = 1.0f;
  return run_gemm(m, g, s);
}

static DeferredRed* red_begin(ssr_train_state* t, float* pool, RedEntry* dev) {
  t->red_host.resize((size_t)t->red_entries + 8);
  t->red.pool = pool;
  t->red.pool_floats = t->red_floats;
  t->red.used = 0;
  t->red.dev = dev;
  t->red.cap = t->red_entries + 8;
  t->red.host = t->red_host.data();
  t->red.n = 0;
  t->unpack_host.clear();
  return &t->red;
}

// dgrad of a conv: dX[p][c] = sum_{tap',n} dY[p + off(tap')][n] * Wd[c][tap'*NP + n]  (the forward kernel on the rotated pack)
static GemmArgs dgrad_base(const ssr_model* m, const ConvT& c, const void* dY, int B, int H, int W) {
  const Lin& L = *c.fwd;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = dY;
  g.lda = L.NP;
  g.B = B;
  g.H = H;
  g.W = W;
  g.M = B * H * W;
  g.taps = 9;
  g.KP = L.NP;
  g.Wt = m->train->arena2 + c.dg_off;
  g.N = c.Cin;
  g.NP = L.KP;
  g.bias = reinterpret_cast<const float*>(m->train->arena2 + m->train->zero_bias_off);
  g.act = ACT_NONE;
  g.slope = 0.0f;
  g.alpha = 1.0f;
  g.eps = 1e-5f;
  g.K_alg = c.Cout;
  g.N_alg = c.Cin;
  return g;
}

static int wgrad_conv(const ssr_model* m, const ConvT& c, const void* dY, const void* X, int ldx, int B, int H, int W, float alpha,
                      float* dwp, float* partial, float* const* grads, cudaStream_t s, int ldy = 0) {
  const Lin& L = *c.fwd;
  if (ldy <= 0) ldy = L.NP;
  if (grads[c.wi]) {
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.dY = dY;
    a.ldy = ldy;
    a.X = X;
    a.ldx = ldx;
    a.B = B;
    a.H = H;
    a.W = W;
    a.M = B * H * W;
    a.taps = 9;
    a.NoutP = L.NP;
    a.CinP = L.KP;
    a.dWp = dwp + c.dwp_off;
    // (the in-kernel bias gradient is a Linear-layer feature: with nine taps only the centre-tap CTAs would carry the extra
    // column sums and become the critical path of the launch -- measured +0.9 ms on cfg2 against 0.64 ms of column sums saved)
    a.dBp = nullptr;
    a.alpha = alpha;
    a.N_alg = c.Cout;
    a.K_alg = c.Cin;
    SSR_TRY(launch_wgrad_tc(a, s));
    PackEntry e;  // un-packed with all the others by one launch at the end of the backward
    memset(&e, 0, sizeof(e));
    e.W = dwp + c.dwp_off;
    e.Wf = grads[c.wi];
    e.bf = a.dBp;
    e.b = grads[c.bi];
    e.kind = 0;
    e.N = c.Cout;
    e.K = c.Cin;
    e.KP = L.KP;
    e.taps = 9;
    e.ps_r = L.ps_r;
    m->train->unpack_host.push_back(e);
  }
  if (grads[c.bi])
    SSR_TRY(launch_colsum(dY, 2, ldy, B * H * W, L.NP, c.Cout, L.ps_r, alpha, grads[c.bi], partial, s, &m->train->red));
  return SSR_OK;
}

static int train_backward_edsr(ssr_model* m, const float* dy, float* const* grads, int B, int h, int w, void* ws, size_t ws_bytes,
                               cudaStream_t s) {
  ssr_train_state* t = m->train;
  const ssr_model_config& c = m->cfg;
  EdsrTrainWs W;
  const size_t need = plan_edsr_train(m, ws, B, h, w, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "train workspace %zu B < required %zu B", ws_bytes, need);
  const int FP = m->FP, nb = c.n_resblocks;
  const size_t T = (size_t)B * h * w;
  red_begin(t, W.red_pool, W.red_dev);
  SSR_CUDA(cudaMemsetAsync(W.dwp, 0, t->dwp_floats * 4, s));
  int H = h * c.scale, Wd = w * c.scale;
  // add_mean is a frozen identity-weight 1x1 conv (common.py:108-121): dL/d(tail.1 output) = dy
  SSR_TRY(launch_nchw_to_nhwc(dy, W.dy64, B, 3, H, Wd, 64, 2, 0, s));
  const int nup = (int)m->up.size();
  {  // tail.1 (edsr.py:37,46)
    const ConvT& cv = t->convs[t->e_last];
    const void* X = nup ? W.hr[nup - 1] : W.bt;
    SSR_TRY(wgrad_conv(m, cv, W.dy64, X, FP, B, H, Wd, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, W.dy64, B, H, Wd);
    if (nup) {
      g.out_T = W.ghr[nup - 1];
      g.ld_T = FP;
    } else {
      g.out_f32 = W.Gt;
      g.ld_f32 = FP;
      g.out_T = W.Gb;
      g.ld_T = FP;
    }
    SSR_TRY(launch_gemm_tc(g, 2, s));
  }
  for (int k = nup - 1; k >= 0; --k) {  // Upsampler (common.py:124-137): PixelShuffle backward, then the conv
    const ConvT& cv = t->convs[t->e_up[k]];
    const int r = m->up[k].ps_r;
    H /= r;
    Wd /= r;
    SSR_TRY(launch_unshuffle(W.ghr[k], W.gU, B, H, Wd, m->F, r, FP, s));
    const void* X = k ? W.hr[k - 1] : W.bt;
    SSR_TRY(wgrad_conv(m, cv, W.gU, X, FP, B, H, Wd, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, W.gU, B, H, Wd);
    if (k) {
      g.out_T = W.ghr[k - 1];
      g.ld_T = FP;
    } else {
      g.out_f32 = W.Gt;
      g.ld_f32 = FP;
      g.out_T = W.Gb;
      g.ld_T = FP;
    }
    SSR_TRY(launch_gemm_tc(g, 2, s));
  }
  {  // res = body(x) + x (edsr.py:43-44): Gt = dL/d(res) feeds the body's last conv and, through the skip, the head output
    const ConvT& cv = t->convs[t->e_body_tail];
    SSR_TRY(wgrad_conv(m, cv, W.Gb, W.rb[nb], FP, B, h, w, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, W.Gb, B, h, w);
    g.out_f32 = W.G;
    g.ld_f32 = FP;
    g.out_T = W.Dh;  // W.Gb is this launch's A operand, so the bf16 copy of G goes to the other buffer and the roles swap
    g.ld_T = FP;
    SSR_TRY(launch_gemm_tc(g, 2, s));
  }
  void* Gb = W.Dh;  // bf16 copy of G
  void* Dh = W.Gb;  // scratch for the gradient at the ReLU
  for (int i = nb - 1; i >= 0; --i) {  // ResBlock: r' = r + res_scale * conv_b(relu(conv_a(r)))  (common.py:150-153)
    const ConvT& ca = t->convs[t->e_res_a[i]];
    const ConvT& cb = t->convs[t->e_res_b[i]];
    SSR_TRY(wgrad_conv(m, cb, Gb, W.tmp[i], FP, B, h, w, c.res_scale, W.dwp, W.partial, grads, s));
    GemmArgs gb = dgrad_base(m, cb, Gb, B, h, w);
    gb.alpha = c.res_scale;
    gb.mask = W.tmp[i];  // ReLU backward: gate by the saved ReLU output
    gb.ld_mask = FP;
    gb.mask_slope = 0.0f;
    gb.out_T = Dh;
    gb.ld_T = FP;
    SSR_TRY(launch_gemm_tc(gb, 2, s));
    SSR_TRY(wgrad_conv(m, ca, Dh, W.rb[i], FP, B, h, w, 1.0f, W.dwp, W.partial, grads, s));
    if (i == 0) SSR_TRY(launch_add_inplace(W.G, W.Gt, nullptr, T * FP, s));  // long skip joins at the head output
    GemmArgs ga = dgrad_base(m, ca, Dh, B, h, w);
    ga.res = W.G;
    ga.ldres = FP;
    ga.out_f32 = W.G;
    ga.ld_f32 = FP;
    ga.out_T = Gb;
    ga.ld_T = FP;
    SSR_TRY(launch_gemm_tc(ga, 2, s));
  }
  if (nb == 0) SSR_TRY(launch_add_inplace(W.G, W.Gt, Gb, T * FP, s));
  // head conv (edsr.py:27,41): dW[n][ci][tap] = sum_p G[p][n] * (x - mean)[p + off(tap)][ci]
  if (grads[t->head_w]) {
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.dY = Gb;
    a.ldy = FP;
    a.X = W.xin64;
    a.ldx = 64;
    a.B = B;
    a.H = h;
    a.W = w;
    a.M = (int)T;
    a.taps = 9;
    a.NoutP = FP;
    a.CinP = 64;
    a.dWp = W.dwp + t->head_dwp;
    a.dBp = nullptr;
    a.alpha = 1.0f;
    a.N_alg = m->F;
    a.K_alg = 3;
    SSR_TRY(launch_wgrad_tc(a, s));
    PackEntry e;
    memset(&e, 0, sizeof(e));
    e.W = W.dwp + t->head_dwp;
    e.Wf = grads[t->head_w];
    e.bf = a.dBp;
    e.b = grads[t->head_b];
    e.kind = 0;
    e.N = m->F;
    e.K = 3;
    e.KP = 64;
    e.taps = 9;
    t->unpack_host.push_back(e);
  }
  if (grads[t->head_b]) SSR_TRY(launch_colsum(Gb, 2, FP, (int)T, FP, m->F, 0, 1.0f, grads[t->head_b], W.partial, s, &t->red));
  t->dx_G = W.G; t->dx_ld = FP; t->dx_C = m->F; t->dx_B = B; t->dx_h = h; t->dx_w = w; t->dx_Hp = h; t->dx_Wp = w;
  t->dx_scale = 1.0f;  // sub_mean is an identity-weight 1x1 conv (common.py:108-121)
  SSR_TRY(launch_unpack_batched(t->unpack_host.data(), W.pack_dev, (int)t->unpack_host.size(), s));
  return launch_deferred_reductions(&t->red, s);  // every bias gradient's second stage, one launch
}


// =============================================================================================
// RCAN (rcan.py:39-77) training executor: EDSR's head / upsampler / reconstruction conv around residual groups of RCABs
// (conv - ReLU - conv - channel attention, rcan.py:11-24)
// =============================================================================================
static int bind_rcan(ssr_model* m) {
  ssr_train_state* t = m->train;
  const ssr_model_config& c = m->cfg;
  const int F = m->F, R = F / c.reduction;
  size_t a2 = 0;
  t->zero_bias_off = a2;
  a2 += 4096 * 4;
  t->head_w = find_idx(t, "head.0.weight", (int64_t)F * 27);
  t->head_b = find_idx(t, "head.0.bias", F);
  if (t->head_w < 0 || t->head_b < 0) return SSR_E_STATE;
  t->head_dwp = t->dwp_floats;
  t->dwp_floats += (size_t)m->FP * 9 * 64;
  t->head_dbp = t->dwp_floats;
  t->dwp_floats += (size_t)m->FP;
  t->red_floats += (size_t)592 * m->FP + 64;
  t->red_entries += 1;
  char nm[96];
  for (int g = 0; g < c.n_resgroups; ++g) {
    for (int b = 0; b < c.n_resblocks; ++b) {
      const int i = g * c.n_resblocks + b;
      snprintf(nm, sizeof(nm), "body.%d.body.%d.body", g, b);
      const std::string p(nm);
      const int ia = add_conv(t, p + ".0", &m->res_a[i], F, F, &a2);
      const int ib = add_conv(t, p + ".2", &m->res_b[i], F, F, &a2);
      if (ia < 0 || ib < 0) return SSR_E_STATE;
      t->r_a.push_back(ia);
      t->r_b.push_back(ib);
      CaT ca;
      ca.w1 = find_idx(t, p + ".3.conv_du.0.weight", (int64_t)R * F);
      ca.b1 = find_idx(t, p + ".3.conv_du.0.bias", R);
      ca.w2 = find_idx(t, p + ".3.conv_du.2.weight", (int64_t)F * R);
      ca.b2 = find_idx(t, p + ".3.conv_du.2.bias", F);
      if (ca.w1 < 0 || ca.b1 < 0 || ca.w2 < 0 || ca.b2 < 0) return SSR_E_STATE;
      t->r_ca.push_back(ca);
    }
    snprintf(nm, sizeof(nm), "body.%d.body.%d", g, c.n_resblocks);
    const int ig = add_conv(t, nm, &m->grp_tail[g], F, F, &a2);
    if (ig < 0) return SSR_E_STATE;
    t->r_gt.push_back(ig);
  }
  snprintf(nm, sizeof(nm), "body.%d", c.n_resgroups);
  t->e_body_tail = add_conv(t, nm, &m->body_tail, F, F, &a2);
  if (t->e_body_tail < 0) return SSR_E_STATE;
  for (size_t i = 0; i < m->up.size(); ++i) {
    snprintf(nm, sizeof(nm), "tail.0.%d", (int)(2 * i));
    const int iu = add_conv(t, nm, &m->up[i], m->up[i].N, F, &a2);
    if (iu < 0) return SSR_E_STATE;
    t->e_up.push_back(iu);
  }
  t->e_last = add_conv(t, "tail.1", &m->last_lin, 3, F, &a2);
  if (t->e_last < 0) return SSR_E_STATE;
  if (m->cfg.arch == SSR_ARCH_HAN) {  // han.py:84-88
    t->h_last_conv = add_conv(t, "last_conv", &m->han_last_conv, F, 11 * F, &a2);
    t->h_last = add_conv(t, "last", &m->han_last, F, 2 * F, &a2);
    t->h_csa_w = find_idx(t, "csa.conv.weight", 27);
    t->h_csa_b = find_idx(t, "csa.conv.bias", 1);
    t->h_csa_g = find_idx(t, "csa.gamma", 1);
    t->h_la_g = find_idx(t, "la.gamma", 1);
    if (t->h_last_conv < 0 || t->h_last < 0 || t->h_csa_w < 0 || t->h_csa_b < 0 || t->h_csa_g < 0 || t->h_la_g < 0) return SSR_E_STATE;
  }
  t->pack_cap = (int)t->convs.size() + 4 * (int)t->r_ca.size() + 24;
  t->arena2_bytes = a2 + 1024;
  SSR_CUDA(cudaMalloc(&t->arena2, t->arena2_bytes));
  SSR_CUDA(cudaMemset(t->arena2, 0, t->arena2_bytes));
  return SSR_OK;
}

struct RcanTrainWs {
  void* xin64;
  float *x0, *r, *gin;
  std::vector<void*> S;        // bf16 copies of the stream: head output, every RCAB output, every group output
  std::vector<void*> tmp, t2;  // per RCAB: ReLU output, second conv's output (bf16)
  float* gates;                // per RCAB: pool [B][F] | gate [B][F] | hid [B][R]
  size_t gate_stride;
  float *ca_partial, *dpool;
  int nsplit;
  void* bt;
  std::vector<void*> hr, ghr;
  void *dy64, *gU;
  float *G, *G2, *Gt;
  void *Gb, *Dh, *Dt;
  float* dwp;
  float* partial;
  float* red_pool;
  RedEntry* red_dev;
  PackEntry* pack_dev;
  // HAN: the 11 fp32 planes of the stack and of its gradient, the layer attention's scratch, last_conv's operand / gradient,
  // the concatenation and its gradient, the CSAM backward's scratch
  float *h_stack, *h_dstack, *h_dlam, *h_dcat, *h_dpre, *h_dxd, *h_coef, *h_scal;
  double *h_energy, *h_D;
  void *h_lam, *h_cat, *h_dcatb;
};

static size_t plan_rcan_train(const ssr_model* m, void* base, int B, int H, int W, RcanTrainWs* w) {
  Carver c(base);
  const size_t T = (size_t)B * H * W;
  const int FP = m->FP, nb = m->cfg.n_resblocks, ng = m->cfg.n_resgroups, nblk = nb * ng, R = m->F / m->cfg.reduction;
  w->xin64 = c.take(T * 64 * 2);
  w->x0 = (float*)c.take(T * FP * 4);
  w->r = (float*)c.take(T * FP * 4);
  w->gin = (float*)c.take(T * FP * 4);
  w->S.resize(1 + (size_t)ng * (nb + 1));
  for (void*& p : w->S) p = c.take(T * FP * 2);
  w->tmp.resize(nblk);
  w->t2.resize(nblk);
  for (int i = 0; i < nblk; ++i) {
    w->tmp[i] = c.take(T * FP * 2);
    w->t2[i] = c.take(T * FP * 2);
  }
  w->gate_stride = (size_t)B * (2 * m->F + R);
  w->gates = (float*)c.take(w->gate_stride * nblk * 4);
  w->nsplit = std::max(1, std::min(64, H * W / 256));
  w->ca_partial = (float*)c.take((size_t)B * w->nsplit * m->F * 4);
  w->dpool = (float*)c.take((size_t)B * m->F * 4);
  w->bt = c.take(T * FP * 2);
  size_t px = T, gu_max = 0;
  w->hr.resize(m->up.size());
  w->ghr.resize(m->up.size());
  for (size_t i = 0; i < m->up.size(); ++i) {
    gu_max = std::max(gu_max, px * (size_t)m->up[i].NP);
    px *= (size_t)m->up[i].ps_r * m->up[i].ps_r;
    w->hr[i] = c.take(px * FP * 2);
    w->ghr[i] = c.take(px * FP * 2);
  }
  w->dy64 = c.take(px * 64 * 2);
  w->gU = c.take(gu_max * 2);
  w->G = (float*)c.take(T * FP * 4);
  w->G2 = (float*)c.take(T * FP * 4);
  w->Gt = (float*)c.take(T * FP * 4);
  w->Gb = c.take(T * FP * 2);
  w->Dh = c.take(T * FP * 2);
  w->Dt = c.take(T * FP * 2);
  w->dwp = (float*)c.take(m->train->dwp_floats * 4);
  w->partial = (float*)c.take(kTrainPartialFloats * 4);
  w->red_pool = (float*)c.take(m->train->red_floats * 4);
  w->red_dev = (RedEntry*)c.take((size_t)(m->train->red_entries + 8) * sizeof(RedEntry));
  w->pack_dev = (PackEntry*)c.take((size_t)m->train->pack_cap * sizeof(PackEntry));
  if (m->cfg.arch == SSR_ARCH_HAN) {
    w->h_stack = (float*)c.take(11 * T * FP * 4);
    w->h_dstack = (float*)c.take(11 * T * FP * 4);
    w->h_dlam = (float*)c.take(T * 11 * FP * 4);
    w->h_dcat = (float*)c.take(T * 2 * FP * 4);
    w->h_dpre = (float*)c.take(T * FP * 4);
    w->h_dxd = (float*)c.take(T * FP * 4);
    w->h_coef = (float*)c.take((size_t)B * 2 * 121 * 4);
    w->h_scal = (float*)c.take(32 * 4);
    w->h_energy = (double*)c.take((size_t)B * 66 * 8);
    w->h_D = (double*)c.take((size_t)B * 121 * 8);
    w->h_lam = c.take(T * 11 * FP * 2);
    w->h_cat = c.take(T * 2 * FP * 2);
    w->h_dcatb = c.take(T * 2 * FP * 2);
  }
  return c.off + 1024;
}

static int train_forward_rcan(ssr_model* m, const float* const* params, const float* x, float* y, int B, int h, int w, void* ws,
                              size_t ws_bytes, cudaStream_t s) {
  ssr_train_state* t = m->train;
  const ssr_model_config& c = m->cfg;
  RcanTrainWs W;
  const size_t need = plan_rcan_train(m, ws, B, h, w, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "train workspace %zu B < required %zu B", ws_bytes, need);
  const int FP = m->FP, F = m->F, R = F / c.reduction, nb = c.n_resblocks;
  // ---- re-pack the (updated) fp32 master weights: forward + dgrad operand layouts, the gates' small matrices as they are ----
  t->pack_host.clear();
  push_copy(t, params[t->head_w], m->dev<float>(m->conv_first_w), F * 27);
  push_copy(t, params[t->head_b], m->dev<float>(m->conv_first_b), F);
  for (const ConvT& cv : t->convs) SSR_TRY(repack_conv(m, cv, params, s));
  for (size_t i = 0; i < t->r_ca.size(); ++i) {
    push_copy(t, params[t->r_ca[i].w1], m->dev<float>(m->ca[i].w1), R * F);
    push_copy(t, params[t->r_ca[i].b1], m->dev<float>(m->ca[i].b1), R);
    push_copy(t, params[t->r_ca[i].w2], m->dev<float>(m->ca[i].w2), F * R);
    push_copy(t, params[t->r_ca[i].b2], m->dev<float>(m->ca[i].b2), F);
  }
  const bool han = c.arch == SSR_ARCH_HAN;
  if (han) {
    push_copy(t, params[t->h_csa_w], m->dev<float>(m->csa_w), 27);
    push_copy(t, params[t->h_csa_b], m->dev<float>(m->csa_b), 1);
    push_copy(t, params[t->h_csa_g], m->dev<float>(m->csa_gamma), 1);
    push_copy(t, params[t->h_la_g], m->dev<float>(m->la_gamma), 1);
  }
  SSR_CHECK((int)t->pack_host.size() <= t->pack_cap, SSR_E_STATE, "train: pack list overflow");
  SSR_TRY(launch_pack_batched(t->pack_host.data(), W.pack_dev, (int)t->pack_host.size(), s));
  const size_t plane = (size_t)B * h * w * FP;
  // ---- forward (rcan.py:68-77 / han.py:90-113), every GEMM operand kept for the backward ----
  SSR_TRY(launch_nchw3_to_nhwc64(x, W.xin64, B, h, w, 1.0f, m->sub_bias, s));
  {
    ConvFirstArgs a;
    memset(&a, 0, sizeof(a));
    a.in = x; a.fh = h; a.fw = w; a.h = h; a.w = w; a.Hp = h; a.Wp = w; a.pad_mode = 2; a.B = B;
    a.in_scale = 1.0f;
    for (int i = 0; i < 3; ++i) a.in_shift[i] = m->sub_bias[i];
    a.Wc = m->dev<float>(m->conv_first_w);
    a.bias = m->dev<float>(m->conv_first_b);
    a.Cout = F;
    a.out_f32 = W.x0; a.ld_f32 = FP; a.out_T = W.S[0]; a.ld_T = FP; a.elem = 2;
    SSR_TRY(launch_conv_first(a, s));
  }
  int k = 0;  // index of the stream state that feeds the next layer
  const float* gcur = W.x0;
  for (int g = 0; g < c.n_resgroups; ++g) {
    const float* rcur = gcur;
    for (int b = 0; b < nb; ++b, ++k) {  // RCAB (rcan.py:21-24)
      const int i = g * nb + b;
      GemmArgs ga = gemm_base(m, m->res_a[i], W.S[k], FP, B, h, w);
      ga.act = ACT_RELU;
      ga.out_T = W.tmp[i];
      ga.ld_T = FP;
      SSR_TRY(run_gemm(m, ga, s));
      GemmArgs gb = gemm_base(m, m->res_b[i], W.tmp[i], FP, B, h, w);
      gb.out_T = W.t2[i];
      gb.ld_T = FP;
      SSR_TRY(run_gemm(m, gb, s));
      float* gs = W.gates + W.gate_stride * i;
      CaArgs ca;
      memset(&ca, 0, sizeof(ca));
      ca.t = W.t2[i]; ca.elem_t = 2; ca.res = rcur; ca.ld = FP; ca.B = B; ca.HW = h * w; ca.C = F; ca.CP = FP; ca.R = R;
      ca.W1 = m->dev<float>(m->ca[i].w1); ca.b1 = m->dev<float>(m->ca[i].b1);
      ca.W2 = m->dev<float>(m->ca[i].w2); ca.b2 = m->dev<float>(m->ca[i].b2);
      ca.partial = W.ca_partial; ca.nsplit = W.nsplit;
      ca.out_f32 = W.r; ca.out_T = W.S[k + 1]; ca.ld_T = FP; ca.elem = 2;
      ca.scale = 1.0f;
      ca.save_pool = gs; ca.save_gate = gs + (size_t)B * F; ca.save_hid = gs + (size_t)2 * B * F;
      SSR_TRY(launch_channel_attention(ca, s));
      rcur = W.r;
    }
    // group tail conv + group skip (rcan.py:33-36)
    float* gout = han ? W.h_stack + (size_t)(c.n_resgroups - g) * plane : W.gin;  // HAN: the group's plane of the stack
    GemmArgs gt = gemm_base(m, m->grp_tail[g], W.S[k], FP, B, h, w);
    gt.res = gcur;
    gt.ldres = FP;
    gt.out_f32 = gout;
    gt.ld_f32 = FP;
    gt.out_T = W.S[k + 1];
    gt.ld_T = FP;
    SSR_TRY(run_gemm(m, gt, s));
    gcur = gout;
    ++k;
  }
  if (!han) {  // body tail conv + long skip (rcan.py:72-73)
    GemmArgs g = gemm_base(m, m->body_tail, W.S[k], FP, B, h, w);
    g.res = W.x0;
    g.ldres = FP;
    g.out_T = W.bt;
    g.ld_T = FP;
    SSR_TRY(run_gemm(m, g, s));
  } else {  // han.py:93-108: plane 0, layer attention -> last_conv, channel-spatial attention, last + long skip
    GemmArgs g0 = gemm_base(m, m->body_tail, W.S[k], FP, B, h, w);
    g0.out_f32 = W.h_stack;
    g0.ld_f32 = FP;
    SSR_TRY(run_gemm(m, g0, s));
    SSR_TRY(launch_han_lam(W.h_stack, plane, FP, B, h * w, F, W.h_energy, m->dev<float>(m->la_gamma), W.h_lam, 11 * FP, 2, 0, s));
    GemmArgs g1 = gemm_base(m, m->han_last_conv, W.h_lam, 11 * FP, B, h, w);
    g1.out_T = reinterpret_cast<uint8_t*>(W.h_cat) + (size_t)FP * 2;
    g1.ld_T = 2 * FP;
    SSR_TRY(run_gemm(m, g1, s));
    SSR_TRY(launch_han_csam(W.h_stack, FP, B, h, w, F, m->dev<float>(m->csa_w), m->dev<float>(m->csa_b), m->dev<float>(m->csa_gamma),
                            W.h_cat, 2 * FP, 2, 0, s));
    GemmArgs g2 = gemm_base(m, m->han_last, W.h_cat, 2 * FP, B, h, w);
    g2.res = W.x0;
    g2.ldres = FP;
    g2.out_T = W.bt;
    g2.ld_T = FP;
    SSR_TRY(run_gemm(m, g2, s));
  }
  const void* cur = W.bt;
  int H = h, Wd = w;
  for (size_t i = 0; i < m->up.size(); ++i) {
    GemmArgs g = gemm_base(m, m->up[i], cur, FP, B, H, Wd);
    g.out_T = W.hr[i];
    g.ld_T = FP;
    SSR_TRY(run_gemm(m, g, s));
    cur = W.hr[i];
    H *= m->up[i].ps_r;
    Wd *= m->up[i].ps_r;
  }
  GemmArgs g = gemm_base(m, m->last_lin, cur, FP, B, H, Wd);
  g.out3_f32 = y;
  g.crop_h = H;
  g.crop_w = Wd;
  for (int i = 0; i < 3; ++i) g.out_shift[i] = m->add_bias[i];
  g.out_scale = 1.0f;
  g.u8_scale = 1.0f;
  return run_gemm(m, g, s);
}

static int train_backward_rcan(ssr_model* m, const float* const* params, const float* dy, float* const* grads, int B, int h, int w,
                               void* ws, size_t ws_bytes, cudaStream_t s) {
  ssr_train_state* t = m->train;
  const ssr_model_config& c = m->cfg;
  RcanTrainWs W;
  const size_t need = plan_rcan_train(m, ws, B, h, w, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "train workspace %zu B < required %zu B", ws_bytes, need);
  (void)params;
  const int FP = m->FP, F = m->F, R = F / c.reduction, nb = c.n_resblocks, ng = c.n_resgroups;
  const size_t T = (size_t)B * h * w;
  red_begin(t, W.red_pool, W.red_dev);
  SSR_CUDA(cudaMemsetAsync(W.dwp, 0, t->dwp_floats * 4, s));
  int H = h * c.scale, Wd = w * c.scale;
  SSR_TRY(launch_nchw_to_nhwc(dy, W.dy64, B, 3, H, Wd, 64, 2, 0, s));
  const int nup = (int)m->up.size();
  {  // tail.1 (rcan.py:66,75)
    const ConvT& cv = t->convs[t->e_last];
    const void* X = nup ? W.hr[nup - 1] : W.bt;
    SSR_TRY(wgrad_conv(m, cv, W.dy64, X, FP, B, H, Wd, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, W.dy64, B, H, Wd);
    if (nup) {
      g.out_T = W.ghr[nup - 1];
      g.ld_T = FP;
    } else {
      g.out_f32 = W.Gt;
      g.ld_f32 = FP;
      g.out_T = W.Gb;
      g.ld_T = FP;
    }
    SSR_TRY(launch_gemm_tc(g, 2, s));
  }
  for (int k = nup - 1; k >= 0; --k) {  // Upsampler (common.py:124-137): PixelShuffle backward, then the conv
    const ConvT& cv = t->convs[t->e_up[k]];
    const int r = m->up[k].ps_r;
    H /= r;
    Wd /= r;
    SSR_TRY(launch_unshuffle(W.ghr[k], W.gU, B, H, Wd, F, r, FP, s));
    const void* X = k ? W.hr[k - 1] : W.bt;
    SSR_TRY(wgrad_conv(m, cv, W.gU, X, FP, B, H, Wd, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, W.gU, B, H, Wd);
    if (k) {
      g.out_T = W.ghr[k - 1];
      g.ld_T = FP;
    } else {
      g.out_f32 = W.Gt;
      g.ld_f32 = FP;
      g.out_T = W.Gb;
      g.ld_T = FP;
    }
    SSR_TRY(launch_gemm_tc(g, 2, s));
  }
  // res = body(x) + x (rcan.py:72-73): Gt = dL/d(res) feeds the body's last conv and, through the long skip, the head output
  int k = ng * (nb + 1);  // stream state that fed the layer being differentiated
  float* Ga = W.G;        // dL/d(current stream state), fp32 ...
  float* Gn = W.G2;
  void* Gb = W.Dh;        // ... and its bf16 copy
  void* Dh = W.Gb;        // scratch for the gradient at the ReLU
  const bool han = c.arch == SSR_ARCH_HAN;
  const size_t plane = T * FP;
  const void* body_dy = W.Gb;  // dL/d(output of the body's last conv), bf16
  if (han) {  // han.py:101-108 in reverse: last, last_conv, layer attention, channel-spatial attention
    const ConvT& cl = t->convs[t->h_last];
    SSR_TRY(wgrad_conv(m, cl, W.Gb, W.h_cat, 2 * FP, B, h, w, 1.0f, W.dwp, W.partial, grads, s));
    {
      GemmArgs g = dgrad_base(m, cl, W.Gb, B, h, w);
      g.out_f32 = W.h_dcat;
      g.ld_f32 = 2 * FP;
      g.out_T = W.h_dcatb;
      g.ld_T = 2 * FP;
      SSR_TRY(launch_gemm_tc(g, 2, s));
    }
    const ConvT& clc = t->convs[t->h_last_conv];
    const void* dy2 = reinterpret_cast<const uint8_t*>(W.h_dcatb) + (size_t)FP * 2;  // channels [F, 2F): last_conv's output
    SSR_TRY(wgrad_conv(m, clc, dy2, W.h_lam, 11 * FP, B, h, w, 1.0f, W.dwp, W.partial, grads, s, 2 * FP));
    for (int j = 0; j < 11; ++j) {  // data gradient plane by plane (64 of the 704 input channels per GEMM)
      GemmArgs g = dgrad_base(m, clc, dy2, B, h, w);
      g.lda = 2 * FP;
      g.Wt = m->train->arena2 + clc.dg_off + (size_t)j * FP * 9 * clc.fwd->NP * 2;
      g.N = FP;
      g.NP = FP;
      g.N_alg = FP;
      g.out_f32 = W.h_dlam + (size_t)j * FP;
      g.ld_f32 = 11 * FP;
      SSR_TRY(launch_gemm_tc(g, 2, s));
    }
    float* dla = grads[t->h_la_g] ? grads[t->h_la_g] : W.h_scal + 30;
    SSR_TRY(launch_han_lam_bwd(W.h_dlam, 11 * FP, W.h_stack, plane, FP, B, h * w, F, W.h_energy, m->dev<float>(m->la_gamma), W.h_D, W.h_coef,
                               dla, W.h_dstack, s));
    SSR_TRY(launch_han_csam_bwd(W.h_stack, FP, W.h_dcat, 2 * FP, W.h_dstack, B, h, w, F, m->dev<float>(m->csa_w), m->dev<float>(m->csa_b),
                                m->dev<float>(m->csa_gamma), W.h_dpre, W.h_dxd, W.h_scal, Gn, W.Dt, s));
    if (grads[t->h_csa_g]) SSR_CUDA(cudaMemcpyAsync(grads[t->h_csa_g], W.h_scal, 4, cudaMemcpyDeviceToDevice, s));
    if (grads[t->h_csa_b]) SSR_CUDA(cudaMemcpyAsync(grads[t->h_csa_b], W.h_scal + 1, 4, cudaMemcpyDeviceToDevice, s));
    if (grads[t->h_csa_w]) SSR_CUDA(cudaMemcpyAsync(grads[t->h_csa_w], W.h_scal + 2, 27 * 4, cudaMemcpyDeviceToDevice, s));
    body_dy = W.Dt;
  }
  {
    const ConvT& cv = t->convs[t->e_body_tail];
    SSR_TRY(wgrad_conv(m, cv, body_dy, W.S[k], FP, B, h, w, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, body_dy, B, h, w);
    g.out_f32 = Ga;
    g.ld_f32 = FP;
    g.out_T = Gb;
    g.ld_T = FP;
    SSR_TRY(launch_gemm_tc(g, 2, s));
    if (han) SSR_TRY(launch_add_inplace(Ga, W.h_dstack + plane, Gb, T * FP, s));  // the last group's own plane of the stack
  }
  for (int g = ng - 1; g >= 0; --g) {  // ResidualGroup: out = conv(rcabs(x)) + x  (rcan.py:27-36)
    {  // Ga = dL/d(out) stays for the group skip; the conv's input gradient goes to Gn
      const ConvT& cv = t->convs[t->r_gt[g]];
      --k;
      SSR_TRY(wgrad_conv(m, cv, Gb, W.S[k], FP, B, h, w, 1.0f, W.dwp, W.partial, grads, s));
      GemmArgs gg = dgrad_base(m, cv, Gb, B, h, w);
      gg.out_f32 = Gn;
      gg.ld_f32 = FP;
      SSR_TRY(launch_gemm_tc(gg, 2, s));
    }
    for (int b = nb - 1; b >= 0; --b) {  // RCAB: out = x + t * gate(t), t = conv_b(relu(conv_a(x)))  (rcan.py:11-24)
      const int i = g * nb + b;
      --k;
      const ConvT& ca = t->convs[t->r_a[i]];
      const ConvT& cb = t->convs[t->r_b[i]];
      const float* gs = W.gates + W.gate_stride * i;
      CaBwdArgs a;
      memset(&a, 0, sizeof(a));
      a.G = Gn; a.t = W.t2[i]; a.ld = FP; a.B = B; a.HW = h * w; a.C = F; a.CP = FP; a.R = R;
      a.W1 = m->dev<float>(m->ca[i].w1); a.W2 = m->dev<float>(m->ca[i].w2);
      a.pool = gs; a.gate = gs + (size_t)B * F; a.hid = gs + (size_t)2 * B * F;
      a.partial = W.ca_partial; a.nsplit = W.nsplit; a.dpool = W.dpool;
      a.dW1 = grads[t->r_ca[i].w1]; a.db1 = grads[t->r_ca[i].b1]; a.dW2 = grads[t->r_ca[i].w2]; a.db2 = grads[t->r_ca[i].b2];
      a.dt = W.Dt; a.ld_dt = FP;
      SSR_TRY(launch_channel_attention_bwd(a, s));
      SSR_TRY(wgrad_conv(m, cb, W.Dt, W.tmp[i], FP, B, h, w, 1.0f, W.dwp, W.partial, grads, s));
      GemmArgs gb = dgrad_base(m, cb, W.Dt, B, h, w);
      gb.mask = W.tmp[i];  // ReLU backward: gate by the saved ReLU output
      gb.ld_mask = FP;
      gb.mask_slope = 0.0f;
      gb.out_T = Dh;
      gb.ld_T = FP;
      SSR_TRY(launch_gemm_tc(gb, 2, s));
      SSR_TRY(wgrad_conv(m, ca, Dh, W.S[k], FP, B, h, w, 1.0f, W.dwp, W.partial, grads, s));
      GemmArgs ga = dgrad_base(m, ca, Dh, B, h, w);
      ga.res = Gn;  // the RCAB's skip
      ga.ldres = FP;
      ga.out_f32 = Gn;
      ga.ld_f32 = FP;
      SSR_TRY(launch_gemm_tc(ga, 2, s));
    }
    // group skip: dL/d(group input) = Gn + Ga; with the long skip joining at the head output (g == 0)
    if (g == 0) SSR_TRY(launch_add_inplace(Gn, W.Gt, nullptr, T * FP, s));
    SSR_TRY(launch_add_inplace(Gn, Ga, Gb, T * FP, s));
    std::swap(Ga, Gn);
    if (han && g > 0)  // the previous group's output is also plane ng - (g - 1) of the stack
      SSR_TRY(launch_add_inplace(Ga, W.h_dstack + (size_t)(ng - g + 1) * plane, Gb, T * FP, s));
  }
  // head conv (rcan.py:63,70): dW[n][ci][tap] = sum_p G[p][n] * (x - mean)[p + off(tap)][ci]
  if (grads[t->head_w]) {
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.dY = Gb; a.ldy = FP; a.X = W.xin64; a.ldx = 64; a.B = B; a.H = h; a.W = w; a.M = (int)T; a.taps = 9;
    a.NoutP = FP; a.CinP = 64;
    a.dWp = W.dwp + t->head_dwp;
    a.dBp = nullptr;
    a.alpha = 1.0f;
    a.N_alg = F;
    a.K_alg = 3;
    SSR_TRY(launch_wgrad_tc(a, s));
    PackEntry e;
    memset(&e, 0, sizeof(e));
    e.W = W.dwp + t->head_dwp;
    e.Wf = grads[t->head_w];
    e.bf = a.dBp;
    e.b = grads[t->head_b];
    e.kind = 0;
    e.N = F;
    e.K = 3;
    e.KP = 64;
    e.taps = 9;
    t->unpack_host.push_back(e);
  }
  if (grads[t->head_b]) SSR_TRY(launch_colsum(Gb, 2, FP, (int)T, FP, F, 0, 1.0f, grads[t->head_b], W.partial, s, &t->red));
  t->dx_G = Ga; t->dx_ld = FP; t->dx_C = F; t->dx_B = B; t->dx_h = h; t->dx_w = w; t->dx_Hp = h; t->dx_Wp = w;
  t->dx_scale = 1.0f;
  SSR_CHECK((int)t->unpack_host.size() <= t->pack_cap, SSR_E_STATE, "train: unpack list overflow");
  SSR_TRY(launch_unpack_batched(t->unpack_host.data(), W.pack_dev, (int)t->unpack_host.size(), s));
  return launch_deferred_reductions(&t->red, s);  // every bias gradient's second stage, one launch
}

// =============================================================================================
// SwinIR (swinir.py:353-372) training executor
// =============================================================================================
static int add_linear(ssr_train_state* t, const std::string& name, const Lin* fwd, int N, int K, const LinMap& map, size_t* a2,
                      LinT* out) {
  out->fwd = fwd;
  out->N = N;
  out->K = K;
  out->map = map;
  out->wi = find_idx(t, name + ".weight", (int64_t)N * K);
  out->bi = find_idx(t, name + ".bias", N);
  if (out->wi < 0 || out->bi < 0) return SSR_E_STATE;
  *a2 = (*a2 + 255) & ~(size_t)255;
  out->dg_off = *a2;
  *a2 += (size_t)fwd->KP * fwd->NP * 2;
  out->dwp_off = t->dwp_floats;
  t->dwp_floats += (size_t)fwd->NP * fwd->KP;
  out->dbp_off = t->dwp_floats;
  t->dwp_floats += (size_t)fwd->NP;
  t->red_floats += (size_t)592 * fwd->NP + 64;
  t->red_entries += 1;
  return SSR_OK;
}
static int add_ln(ssr_train_state* t, const std::string& name, const LNp* p, int C, LnT* out) {
  out->p = p;
  out->gi = find_idx(t, name + ".weight", C);
  out->bi = find_idx(t, name + ".bias", C);
  t->red_floats += (size_t)592 * 3 * C + 64;
  t->red_entries += 3;
  return (out->gi < 0 || out->bi < 0) ? SSR_E_STATE : SSR_OK;
}

static int bind_swinir(ssr_model* m) {
  ssr_train_state* t = m->train;
  const ssr_model_config& c = m->cfg;
  SSR_CHECK(c.upsampler == 0, SSR_E_INVALID, "training path: only the 'pixelshuffle' upsampler is built");
  SSR_CHECK(c.window_size == 8, SSR_E_INVALID, "training path: window_size 8 only");
  const int C = m->C;
  size_t a2 = 0;
  t->zero_bias_off = a2;
  a2 += 4096 * 4;
  t->head_w = find_idx(t, "conv_first.weight", (int64_t)C * 27);
  t->head_b = find_idx(t, "conv_first.bias", C);
  if (t->head_w < 0 || t->head_b < 0) return SSR_E_STATE;
  t->head_dwp = t->dwp_floats;
  t->dwp_floats += (size_t)m->CP * 9 * 64;
  t->head_dbp = t->dwp_floats;
  t->dwp_floats += (size_t)m->CP;
  t->red_floats += (size_t)592 * m->CP + 64;
  t->red_entries += 1;
  SSR_TRY(add_ln(t, "patch_embed.norm", &m->pe_norm, C, &t->s_pe));
  t->s_blocks.resize(m->layers.size());
  for (size_t li = 0; li < m->layers.size(); ++li) {
    const Layer& L = m->layers[li];
    LinMap ident{0, C, L.d, L.DP, L.QP, 1.0f};
    LinMap mq{1, C, L.d, L.DP, L.QP, 1.0f / sqrtf((float)L.d)};
    LinMap mp{2, C, L.d, L.DP, L.QP, 1.0f};
    for (size_t bi = 0; bi < L.blocks.size(); ++bi) {
      const Block& B = L.blocks[bi];
      BlockT bt;
      char pre[128];
      snprintf(pre, sizeof(pre), "layers.%d.residual_group.blocks.%d", (int)li, (int)bi);
      const std::string p(pre);
      SSR_TRY(add_ln(t, p + ".norm1", &B.norm1, C, &bt.n1));
      SSR_TRY(add_ln(t, p + ".norm2", &B.norm2, C, &bt.n2));
      SSR_TRY(add_linear(t, p + ".attn.qkv", &B.qkv, 3 * C, C, mq, &a2, &bt.qkv));
      SSR_TRY(add_linear(t, p + ".attn.proj", &B.proj, C, C, mp, &a2, &bt.proj));
      SSR_TRY(add_linear(t, p + ".mlp.fc1", &B.fc1, m->HID, C, ident, &a2, &bt.fc1));
      SSR_TRY(add_linear(t, p + ".mlp.fc2", &B.fc2, C, m->HID, ident, &a2, &bt.fc2));
      bt.table_i = find_idx(t, p + ".attn.relative_position_bias_table", (int64_t)225 * L.heads);
      if (bt.table_i < 0) return SSR_E_STATE;
      bt.bias_off = B.bias_off;
      t->s_blocks[li].push_back(bt);
    }
    char nm[64];
    snprintf(nm, sizeof(nm), "layers.%d.conv", (int)li);
    const int ic = add_conv(t, nm, &L.conv, C, C, &a2);
    if (ic < 0) return SSR_E_STATE;
    t->s_conv.push_back(ic);
  }
  SSR_TRY(add_ln(t, "norm", &m->final_norm, C, &t->s_fin));
  t->s_cab = add_conv(t, "conv_after_body", &m->conv_after_body, C, C, &a2);
  t->s_cbu = add_conv(t, "conv_before_upsample.0", &m->conv_before_up, 64, C, &a2);
  if (t->s_cab < 0 || t->s_cbu < 0) return SSR_E_STATE;
  for (size_t i = 0; i < m->up.size(); ++i) {
    char nm[64];
    snprintf(nm, sizeof(nm), "upsample.%d", (int)(2 * i));
    const int iu = add_conv(t, nm, &m->up[i], m->up[i].N, 64, &a2);
    if (iu < 0) return SSR_E_STATE;
    t->e_up.push_back(iu);
  }
  t->e_last = add_conv(t, "conv_last", &m->last_lin, 3, 64, &a2);
  if (t->e_last < 0) return SSR_E_STATE;
  t->arena2_bytes = a2 + 1024;
  SSR_CUDA(cudaMalloc(&t->arena2, t->arena2_bytes));
  SSR_CUDA(cudaMemset(t->arena2, 0, t->arena2_bytes));
  return SSR_OK;
}

struct SwinBlockWs {
  const float* tin;  // fp32 LN1 input (layer input or the previous block's output)
  float* tmid;       // fp32 after the attention residual
  float* tout;       // fp32 block output (null for the last block of a layer: only the bf16 copy is needed)
  void *xn1, *qkv, *o, *xn2, *u, *h;
};
struct SwinTrainWs {
  void* xin64;
  float* x0;
  std::vector<float*> g;  // n_layers + 1
  std::vector<std::vector<SwinBlockWs>> blk;
  std::vector<void*> tb;
  void *xnf, *tbf, *cbu;
  std::vector<void*> hr, ghr;
  void *dy64, *gU, *dcbu;
  float *G, *Gt, *Gres, *dB;
  void *Gb, *Gtb, *dH, *dXn, *dO, *dQKV;
  float* dwp;
  float* partial;
  float* red_pool;
  RedEntry* red_dev;
  PackEntry* pack_dev;
};

static size_t plan_swin_train(const ssr_model* m, void* base, int B, int Hp, int Wp, SwinTrainWs* w) {
  Carver c(base);
  const size_t T = (size_t)B * Hp * Wp;
  const int CP = m->CP, HP = m->HP;
  const size_t nL = m->layers.size();
  w->xin64 = c.take(T * 64 * 2);
  w->x0 = (float*)c.take(T * CP * 4);
  w->g.resize(nL + 1);
  for (size_t i = 0; i <= nL; ++i) w->g[i] = (float*)c.take(T * CP * 4);
  w->blk.resize(nL);
  w->tb.resize(nL);
  int QPmax = 0, heads_max = 0;
  for (size_t li = 0; li < nL; ++li) {
    const Layer& L = m->layers[li];
    QPmax = std::max(QPmax, L.QP);
    heads_max = std::max(heads_max, L.heads);
    w->blk[li].resize(L.blocks.size());
    for (size_t bi = 0; bi < L.blocks.size(); ++bi) {
      SwinBlockWs& b = w->blk[li][bi];
      b.tin = bi == 0 ? w->g[li] : w->blk[li][bi - 1].tout;
      b.xn1 = c.take(T * CP * 2);
      b.qkv = c.take(T * 3 * L.QP * 2);
      b.o = c.take(T * L.QP * 2);
      b.tmid = (float*)c.take(T * CP * 4);
      b.xn2 = c.take(T * CP * 2);
      b.u = c.take(T * HP * 2);
      b.h = c.take(T * HP * 2);
      b.tout = bi + 1 < L.blocks.size() ? (float*)c.take(T * CP * 4) : nullptr;
    }
    w->tb[li] = c.take(T * CP * 2);
  }
  w->xnf = c.take(T * CP * 2);
  w->tbf = c.take(T * CP * 2);
  w->cbu = c.take(T * 64 * 2);
  w->dcbu = c.take(T * 64 * 2);
  size_t px = T, gu_max = 0;
  w->hr.resize(m->up.size());
  w->ghr.resize(m->up.size());
  for (size_t i = 0; i < m->up.size(); ++i) {
    gu_max = std::max(gu_max, px * (size_t)m->up[i].NP);
    px *= (size_t)m->up[i].ps_r * m->up[i].ps_r;
    w->hr[i] = c.take(px * 64 * 2);
    w->ghr[i] = c.take(px * 64 * 2);
  }
  w->dy64 = c.take(px * 64 * 2);
  w->gU = c.take(gu_max * 2);
  w->G = (float*)c.take(T * CP * 4);
  w->Gt = (float*)c.take(T * CP * 4);
  w->Gres = (float*)c.take(T * CP * 4);
  w->Gb = c.take(T * CP * 2);
  w->Gtb = c.take(T * CP * 2);
  w->dH = c.take(T * HP * 2);
  w->dXn = c.take(T * CP * 2);
  w->dO = c.take(T * QPmax * 2);
  w->dQKV = c.take(T * 3 * QPmax * 2);
  w->dB = (float*)c.take((size_t)heads_max * 64 * 64 * 4);
  w->dwp = (float*)c.take(m->train->dwp_floats * 4);
  w->partial = (float*)c.take(kTrainPartialFloats * 4);
  w->red_pool = (float*)c.take(m->train->red_floats * 4);
  w->red_dev = (RedEntry*)c.take((size_t)(m->train->red_entries + 8) * sizeof(RedEntry));
  w->pack_dev = (PackEntry*)c.take((size_t)(2 * m->train->red_entries + 16) * sizeof(PackEntry));
  return c.off + 1024;
}

static int repack_linear(ssr_model* m, const LinT& l, const float* const* params, cudaStream_t) {
  const Lin& L = *l.fwd;
  PackEntry e;
  memset(&e, 0, sizeof(e));
  e.W = params[l.wi];
  e.b = params[l.bi];
  e.Wf = m->arena + L.w_off;
  e.bf = m->dev<float>(L.b_off);
  e.Wd = m->train->arena2 + l.dg_off;
  e.kind = 1;
  e.N = l.N;
  e.K = l.K;
  e.NP = L.NP;
  e.KP = L.KP;
  e.taps = 1;
  e.map = l.map;
  m->train->pack_host.push_back(e);
  return SSR_OK;
}
static int repack_ln(ssr_model* m, const LnT& l, const float* const* params, int C, cudaStream_t) {
  push_copy(m->train, params[l.gi], m->dev<float>(l.p->g_off), C);
  push_copy(m->train, params[l.bi], m->dev<float>(l.p->b_off), C);
  return SSR_OK;
}
static void repack_table(ssr_model* m, const float* table, float* dst, int nb, int heads) {
  PackEntry e;
  memset(&e, 0, sizeof(e));
  e.W = table;  // [nb][heads]
  e.Wf = dst;   // [heads][nb]
  e.kind = 3;  // out[k * N + n] = W[n * K + k]: N = nb rows of the source, K = heads
  e.N = nb;
  e.K = heads;
  m->train->pack_host.push_back(e);
}

static void train_padded(const ssr_model* m, int h, int w, int* Hp, int* Wp) {
  const int ws = m->cfg.window_size;
  *Hp = (h + ws - 1) / ws * ws;
  *Wp = (w + ws - 1) / ws * ws;
}

static const float kMean3[3] = {0.4488f, 0.4371f, 0.4040f};  // common.py:223

static int train_forward_swinir(ssr_model* m, const float* const* params, const float* drop, const float* x, float* y, int B, int h,
                                int w, void* ws, size_t ws_bytes, cudaStream_t s) {
  ssr_train_state* t = m->train;
  const ssr_model_config& c = m->cfg;
  int Hp, Wp;
  train_padded(m, h, w, &Hp, &Wp);
  SSR_CHECK(Hp - h < h && Wp - w < w, SSR_E_INVALID, "reflect padding needs pad < size (%dx%d)", h, w);
  SwinTrainWs W;
  const size_t need = plan_swin_train(m, ws, B, Hp, Wp, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "train workspace %zu B < required %zu B", ws_bytes, need);
  const int C = m->C, CP = m->CP, HP = m->HP;
  const int T = B * Hp * Wp;
  // ---- re-pack the fp32 master parameters ----
  t->pack_host.clear();
  push_copy(t, params[t->head_w], m->dev<float>(m->conv_first_w), C * 27);
  push_copy(t, params[t->head_b], m->dev<float>(m->conv_first_b), C);
  SSR_TRY(repack_ln(m, t->s_pe, params, C, s));
  SSR_TRY(repack_ln(m, t->s_fin, params, C, s));
  for (size_t li = 0; li < t->s_blocks.size(); ++li)
    for (const BlockT& b : t->s_blocks[li]) {
      SSR_TRY(repack_ln(m, b.n1, params, C, s));
      SSR_TRY(repack_ln(m, b.n2, params, C, s));
      SSR_TRY(repack_linear(m, b.qkv, params, s));
      SSR_TRY(repack_linear(m, b.proj, params, s));
      SSR_TRY(repack_linear(m, b.fc1, params, s));
      SSR_TRY(repack_linear(m, b.fc2, params, s));
      repack_table(m, params[b.table_i], m->dev<float>(b.bias_off), 225, m->layers[li].heads);
    }
  for (const ConvT& cv : t->convs) SSR_TRY(repack_conv(m, cv, params, s));
  SSR_TRY(launch_pack_batched(t->pack_host.data(), W.pack_dev, (int)t->pack_host.size(), s));
  // ---- forward (swinir.py:353-372, training branch: reflect pad), un-fused so that every GEMM operand is kept ----
  const float in_scale = 1.0f / c.img_range;
  float in_shift[3] = {-kMean3[0], -kMean3[1], -kMean3[2]};
  SSR_TRY(launch_input_nhwc64(x, W.xin64, B, h, w, Hp, Wp, in_scale, in_shift, s));
  {
    ConvFirstArgs a;
    memset(&a, 0, sizeof(a));
    a.in = x;
    a.fh = h;
    a.fw = w;
    a.h = h;
    a.w = w;
    a.Hp = Hp;
    a.Wp = Wp;
    a.pad_mode = SSR_PAD_TRAIN;
    a.B = B;
    a.in_scale = in_scale;
    for (int i = 0; i < 3; ++i) a.in_shift[i] = in_shift[i];
    a.Wc = m->dev<float>(m->conv_first_w);
    a.bias = m->dev<float>(m->conv_first_b);
    a.Cout = C;
    a.out_f32 = W.x0;
    a.ld_f32 = CP;
    a.elem = 2;
    SSR_TRY(launch_conv_first(a, s));
  }
  {
    LnArgs a;
    memset(&a, 0, sizeof(a));
    a.in = W.x0;
    a.ld_in = CP;
    a.M = T;
    a.C = C;
    a.CP = CP;
    a.g1 = m->dev<float>(m->pe_norm.g_off);
    a.b1 = m->dev<float>(m->pe_norm.b_off);
    a.out_f32 = W.g[0];
    a.ld_f32 = CP;
    const Block& b0 = m->layers[0].blocks[0];
    a.g2 = m->dev<float>(b0.norm1.g_off);
    a.b2 = m->dev<float>(b0.norm1.b_off);
    a.out_T = W.blk[0][0].xn1;
    a.ld_T = CP;
    a.elem = 2;
    a.eps = 1e-5f;
    SSR_TRY(launch_layernorm(a, s));
  }
  const int nL = (int)m->layers.size();
  int kblk = 0;  // running block index: drop holds [2 * kblk] (attention branch) and [2 * kblk + 1] (MLP branch), B floats each
  for (int li = 0; li < nL; ++li) {
    const Layer& L = m->layers[li];
    const int depth = (int)L.blocks.size();
    for (int bi = 0; bi < depth; ++bi, ++kblk) {
      const Block& blk = L.blocks[bi];
      SwinBlockWs& bw = W.blk[li][bi];
      {
        GemmArgs g = gemm_base(m, blk.qkv, bw.xn1, CP, B, Hp, Wp);
        g.out_T = bw.qkv;
        g.ld_T = 3 * L.QP;
        SSR_TRY(run_gemm(m, g, s));
      }
      {
        AttnArgs a;
        memset(&a, 0, sizeof(a));
        a.qkv = bw.qkv;
        a.ld_qkv = 3 * L.QP;
        a.QP = L.QP;
        a.o = bw.o;
        a.ld_o = L.QP;
        a.bias = m->dev<float>(blk.bias_off);
        a.B = B;
        a.H = Hp;
        a.W = Wp;
        a.ws = c.window_size;
        a.shift = (bi % 2 == 0) ? 0 : c.window_size / 2;
        a.heads = L.heads;
        a.d = L.d;
        a.DP = L.DP;
        SSR_CUDA(cudaMemsetAsync(bw.o, 0, (size_t)T * L.QP * 2, s));
        SSR_TRY(launch_attn_mma(a, s));  // (the per-(window, head) attn_flash kernel measured 11 % slower on 8x8 windows)
      }
      {
        GemmArgs g = gemm_base(m, blk.proj, bw.o, L.QP, B, Hp, Wp);
        g.res = bw.tin;
        g.ldres = CP;
        if (drop) {  // x = shortcut + drop_path(attn)  (swinir.py:171)
          g.row_scale = drop + (size_t)(2 * kblk) * B;
          g.rows_per_scale = Hp * Wp;
        }
        g.out_f32 = bw.tmid;
        g.ld_f32 = CP;
        g.out_ln = bw.xn2;
        g.ld_ln = CP;
        g.gamma = m->dev<float>(blk.norm2.g_off);
        g.beta = m->dev<float>(blk.norm2.b_off);
        SSR_TRY(run_gemm(m, g, s));
      }
      {
        GemmArgs g = gemm_base(m, blk.fc1, bw.xn2, CP, B, Hp, Wp);
        g.act = ACT_GELU;
        g.out_pre = bw.u;  // holds gelu'(u), not u: the dgrad of fc2 only multiplies by it
        g.ld_pre = HP;
        g.pre_mode = 1;
        g.out_T = bw.h;
        g.ld_T = HP;
        SSR_TRY(run_gemm(m, g, s));
      }
      {
        GemmArgs g = gemm_base(m, blk.fc2, bw.h, HP, B, Hp, Wp);
        g.res = bw.tmid;
        g.ldres = CP;
        if (drop) {  // x = x + drop_path(mlp(norm2(x)))  (swinir.py:172)
          g.row_scale = drop + (size_t)(2 * kblk + 1) * B;
          g.rows_per_scale = Hp * Wp;
        }
        if (bi + 1 < depth) {
          g.out_f32 = bw.tout;
          g.ld_f32 = CP;
          g.out_ln = W.blk[li][bi + 1].xn1;
          g.ld_ln = CP;
          g.gamma = m->dev<float>(L.blocks[bi + 1].norm1.g_off);
          g.beta = m->dev<float>(L.blocks[bi + 1].norm1.b_off);
        } else {
          g.out_T = W.tb[li];
          g.ld_T = CP;
        }
        SSR_TRY(run_gemm(m, g, s));
      }
    }
    {
      GemmArgs g = gemm_base(m, L.conv, W.tb[li], CP, B, Hp, Wp);
      g.res = W.g[li];
      g.ldres = CP;
      g.out_f32 = W.g[li + 1];
      g.ld_f32 = CP;
      const LNp& nx = li + 1 < nL ? m->layers[li + 1].blocks[0].norm1 : m->final_norm;
      g.out_ln = li + 1 < nL ? W.blk[li + 1][0].xn1 : W.xnf;
      g.ld_ln = CP;
      g.gamma = m->dev<float>(nx.g_off);
      g.beta = m->dev<float>(nx.b_off);
      SSR_TRY(run_gemm(m, g, s));
    }
  }
  {
    GemmArgs g = gemm_base(m, m->conv_after_body, W.xnf, CP, B, Hp, Wp);
    g.res = W.x0;
    g.ldres = CP;
    g.out_T = W.tbf;
    g.ld_T = CP;
    SSR_TRY(run_gemm(m, g, s));
  }
  {
    GemmArgs g = gemm_base(m, m->conv_before_up, W.tbf, CP, B, Hp, Wp);
    g.act = ACT_LEAKY;
    g.slope = 0.01f;
    g.out_T = W.cbu;
    g.ld_T = 64;
    SSR_TRY(run_gemm(m, g, s));
  }
  const void* cur = W.cbu;
  int H = Hp, Wd = Wp;
  for (size_t i = 0; i < m->up.size(); ++i) {
    GemmArgs g = gemm_base(m, m->up[i], cur, 64, B, H, Wd);
    g.out_T = W.hr[i];
    g.ld_T = 64;
    SSR_TRY(run_gemm(m, g, s));
    cur = W.hr[i];
    H *= m->up[i].ps_r;
    Wd *= m->up[i].ps_r;
  }
  GemmArgs g = gemm_base(m, m->last_lin, cur, 64, B, H, Wd);
  g.out3_f32 = y;
  g.crop_h = h * c.scale;
  g.crop_w = w * c.scale;
  for (int i = 0; i < 3; ++i) g.out_shift[i] = kMean3[i];
  g.out_scale = c.img_range;
  g.u8_scale = 1.0f;
  return run_gemm(m, g, s);
}

// dgrad of a linear layer: dX[M][K] = dY[M][N] W[N][K] -> the forward GEMM kernel on the transposed pack
static GemmArgs dgrad_lin(const ssr_model* m, const LinT& l, const void* dY, int M) {
  const Lin& L = *l.fwd;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = dY;
  g.lda = L.NP;
  g.B = 1;
  g.H = 1;
  g.W = M;
  g.M = M;
  g.taps = 1;
  g.KP = L.NP;
  g.Wt = m->train->arena2 + l.dg_off;
  g.N = L.KP;  // every padded column is produced (pads are exact zeros): downstream kernels read whole rows
  g.NP = L.KP;
  g.bias = reinterpret_cast<const float*>(m->train->arena2 + m->train->zero_bias_off);
  g.alpha = 1.0f;
  g.eps = 1e-5f;
  g.K_alg = l.N;
  g.N_alg = l.K;
  return g;
}
static int wgrad_lin(const ssr_model* m, const LinT& l, const void* dY, const void* X, int M, float* dwp, float* partial,
                     float* const* grads, cudaStream_t s, bool bias_done = false) {
  const Lin& L = *l.fwd;
  if (grads[l.wi]) {
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.dY = dY;
    a.ldy = L.NP;
    a.X = X;
    a.ldx = L.KP;
    a.B = 1;
    a.H = 1;
    a.W = M;
    a.M = M;
    a.taps = 1;
    a.NoutP = L.NP;
    a.CinP = L.KP;
    a.dWp = dwp + l.dwp_off;
    a.dBp = (grads[l.bi] && !bias_done) ? dwp + l.dbp_off : nullptr;
    a.alpha = 1.0f;
    a.N_alg = l.N;
    a.K_alg = l.K;
    SSR_TRY(launch_wgrad_tc(a, s));
    PackEntry e;
    memset(&e, 0, sizeof(e));
    e.W = dwp + l.dwp_off;
    e.Wf = grads[l.wi];
    e.bf = a.dBp;
    e.b = a.dBp ? grads[l.bi] : nullptr;
    e.kind = 1;
    e.N = l.N;
    e.K = l.K;
    e.KP = L.KP;
    e.taps = 1;
    e.map = l.map;
    m->train->unpack_host.push_back(e);
  }
  if (grads[l.bi] && !bias_done && !grads[l.wi])  // bias_done: the LayerNorm backward that produced dY already summed its columns
    SSR_TRY(launch_colsum_map(dY, 2, L.NP, M, L.NP, l.N, l.map, grads[l.bi], partial, s, &m->train->red));
  return SSR_OK;
}
static int ln_backward(ssr_train_state* ts, const LnT& l, const float* gamma_dev, const float* x, const void* dy, int elem_dy, const float* Gin, float* Gout,
                       void* Gb, int M, int C, int CP, float* partial, float* const* grads, cudaStream_t s,
                       const float* gb_scale = nullptr, int rows_per_scale = 1, float* gb_colsum = nullptr) {
  LnBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x;
  a.ldx = CP;
  a.dy = dy;
  a.ld_dy = CP;
  a.elem_dy = elem_dy;
  a.gamma = gamma_dev;
  a.Gin = Gin;
  a.Gout = Gout;
  a.Gb = Gb;
  a.ldg = CP;
  a.M = M;
  a.C = C;
  a.CP = CP;
  a.eps = 1e-5f;
  a.partial = partial;
  a.gb_scale = gb_scale;
  a.rows_per_scale = rows_per_scale;
  a.gb_colsum = (grads[l.gi] && grads[l.bi]) ? gb_colsum : nullptr;
  if (grads[l.gi] && grads[l.bi]) {
    a.dgamma = grads[l.gi];
    a.dbeta = grads[l.bi];
  }
  return launch_ln_bwd(a, s, &ts->red);
}

static int train_backward_swinir(ssr_model* m, const float* dy, const float* drop, float* const* grads, int B, int h, int w, void* ws,
                                 size_t ws_bytes, cudaStream_t s) {
  ssr_train_state* t = m->train;
  const ssr_model_config& c = m->cfg;
  int Hp, Wp;
  train_padded(m, h, w, &Hp, &Wp);
  SwinTrainWs W;
  const size_t need = plan_swin_train(m, ws, B, Hp, Wp, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "train workspace %zu B < required %zu B", ws_bytes, need);
  const int C = m->C, CP = m->CP;
  const int T = B * Hp * Wp;
  red_begin(t, W.red_pool, W.red_dev);
  SSR_CUDA(cudaMemsetAsync(W.dwp, 0, t->dwp_floats * 4, s));
  int H = Hp * c.scale, Wd = Wp * c.scale;
  // y = (conv_last(.) + mean) * img_range, cropped (swinir.py:366-372): dL/d(conv_last output) = dy * img_range inside the crop
  SSR_TRY(launch_grad_nhwc64(dy, W.dy64, B, h * c.scale, w * c.scale, H, Wd, c.img_range, s));
  const int nup = (int)m->up.size();
  {
    const ConvT& cv = t->convs[t->e_last];
    SSR_TRY(wgrad_conv(m, cv, W.dy64, nup ? W.hr[nup - 1] : W.cbu, 64, B, H, Wd, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, W.dy64, B, H, Wd);
    g.out_T = nup ? W.ghr[nup - 1] : W.dcbu;
    g.ld_T = 64;
    if (!nup) {
      g.mask = W.cbu;
      g.ld_mask = 64;
      g.mask_slope = 0.01f;
    }
    SSR_TRY(launch_gemm_tc(g, 2, s));
  }
  for (int k = nup - 1; k >= 0; --k) {
    const ConvT& cv = t->convs[t->e_up[k]];
    const int r = m->up[k].ps_r;
    H /= r;
    Wd /= r;
    SSR_TRY(launch_unshuffle(W.ghr[k], W.gU, B, H, Wd, 64, r, 64, s));
    SSR_TRY(wgrad_conv(m, cv, W.gU, k ? W.hr[k - 1] : W.cbu, 64, B, H, Wd, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, W.gU, B, H, Wd);
    g.out_T = k ? W.ghr[k - 1] : W.dcbu;
    g.ld_T = 64;
    if (k == 0) {  // LeakyReLU(0.01) of conv_before_upsample (swinir.py:321-324): gate by the sign of its output
      g.mask = W.cbu;
      g.ld_mask = 64;
      g.mask_slope = 0.01f;
    }
    SSR_TRY(launch_gemm_tc(g, 2, s));
  }
  {  // conv_before_upsample.0
    const ConvT& cv = t->convs[t->s_cbu];
    SSR_TRY(wgrad_conv(m, cv, W.dcbu, W.tbf, CP, B, Hp, Wp, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, W.dcbu, B, Hp, Wp);
    g.out_f32 = W.Gres;
    g.ld_f32 = CP;
    g.out_T = W.Gb;
    g.ld_T = CP;
    SSR_TRY(launch_gemm_tc(g, 2, s));
  }
  {  // conv_after_body(norm(features)) + x0 (swinir.py:362): Gres also flows to x0 through the long skip
    const ConvT& cv = t->convs[t->s_cab];
    SSR_TRY(wgrad_conv(m, cv, W.Gb, W.xnf, CP, B, Hp, Wp, 1.0f, W.dwp, W.partial, grads, s));
    GemmArgs g = dgrad_base(m, cv, W.Gb, B, Hp, Wp);
    g.out_T = W.dXn;
    g.ld_T = CP;
    SSR_TRY(launch_gemm_tc(g, 2, s));
  }
  const int nL = (int)m->layers.size();
  int kblk = 0;
  for (int li = 0; li < nL; ++li) kblk += (int)m->layers[li].blocks.size();
  const size_t per_sample = (size_t)Hp * Wp * CP;
  SSR_TRY(ln_backward(t, t->s_fin, m->dev<float>(m->final_norm.g_off), W.g[nL], W.dXn, 2, nullptr, W.G, W.Gb, T, C, CP, W.partial, grads, s));
  for (int li = nL - 1; li >= 0; --li) {
    const Layer& L = m->layers[li];
    const int depth = (int)L.blocks.size();
    {  // RSTB: g' = g + conv(t_last) (swinir.py:245-246); G = dL/dg' keeps flowing through the skip
      const ConvT& cv = t->convs[t->s_conv[li]];
      SSR_TRY(wgrad_conv(m, cv, W.Gb, W.tb[li], CP, B, Hp, Wp, 1.0f, W.dwp, W.partial, grads, s));
      GemmArgs g = dgrad_base(m, cv, W.Gb, B, Hp, Wp);
      g.out_f32 = W.Gt;
      g.ld_f32 = CP;
      g.out_T = W.Gtb;
      g.ld_T = CP;
      SSR_TRY(launch_gemm_tc(g, 2, s));
    }
    for (int bi = depth - 1; bi >= 0; --bi) {
      --kblk;
      const Block& blk = L.blocks[bi];
      const BlockT& bt = t->s_blocks[li][bi];
      // proj / fc2 bias gradients ride in the LayerNorm backward kernels (column sums of the bf16 gradient copy they write):
      // LN2 backward of block b -> proj bias of b; LN1 backward of block b (b > 0) -> fc2 bias of block b - 1
      auto sums_ok = [&](int b_) {
        const BlockT& q = t->s_blocks[li][b_];
        return grads[q.n1.gi] && grads[q.n1.bi] && grads[q.n2.gi] && grads[q.n2.bi];
      };
      const bool ln_sums = sums_ok(bi);
      const bool fc2_bias_done = bi + 1 < depth && sums_ok(bi + 1);
      const SwinBlockWs& bw = W.blk[li][bi];
      // ---- MLP: t_out = t_mid + fc2(GELU(fc1(LN2(t_mid))))  (swinir.py:172, common.py:184-194) ----
      // with stochastic depth the branch sees dL/dt_out scaled per sample: Gtb = bf16(Gt * drop[2k+1][b])
      // (for the other blocks the LN1 backward of the block after this one already wrote the scaled copy)
      if (drop && bi == depth - 1)
        SSR_TRY(launch_scale_to_bf16(W.Gt, drop + (size_t)(2 * kblk + 1) * B, per_sample, W.Gtb, (size_t)T * CP, s));
      // (the bias gradient of fc2 = column sums of Gtb came with the LN1 backward of the block after this one)
      SSR_TRY(wgrad_lin(m, bt.fc2, W.Gtb, bw.h, T, W.dwp, W.partial, grads, s, fc2_bias_done));
      {
        GemmArgs g = dgrad_lin(m, bt.fc2, W.Gtb, T);
        g.mask = bw.u;  // GELU backward: the forward saved gelu'(u)
        g.ld_mask = m->HP;
        g.mask_mode = 2;
        g.out_T = W.dH;
        g.ld_T = m->HP;
        SSR_TRY(launch_gemm_tc(g, 2, s));
      }
      SSR_TRY(wgrad_lin(m, bt.fc1, W.dH, bw.xn2, T, W.dwp, W.partial, grads, s));
      {
        GemmArgs g = dgrad_lin(m, bt.fc1, W.dH, T);
        g.out_T = W.dXn;
        g.ld_T = CP;
        SSR_TRY(launch_gemm_tc(g, 2, s));
      }
      // the bf16 copy feeds the attention branch: scaled by its stochastic-depth factor
      SSR_TRY(ln_backward(t, bt.n2, m->dev<float>(blk.norm2.g_off), bw.tmid, W.dXn, 2, W.Gt, W.Gt, W.Gtb, T, C, CP, W.partial, grads, s,
                          drop ? drop + (size_t)(2 * kblk) * B : nullptr, Hp * Wp, ln_sums ? grads[bt.proj.bi] : nullptr));
      // ---- attention: t_mid = t_in + proj(W-MSA(LN1(t_in)))  (swinir.py:149-171) ----

      SSR_TRY(wgrad_lin(m, bt.proj, W.Gtb, bw.o, T, W.dwp, W.partial, grads, s, ln_sums));
      {
        GemmArgs g = dgrad_lin(m, bt.proj, W.Gtb, T);
        g.out_T = W.dO;
        g.ld_T = L.QP;
        SSR_TRY(launch_gemm_tc(g, 2, s));
      }
      {
        AttnBwdArgs a;
        memset(&a, 0, sizeof(a));
        a.qkv = bw.qkv;
        a.ld_qkv = 3 * L.QP;
        a.QP = L.QP;
        a.d_o = W.dO;
        a.ld_do = L.QP;
        a.dqkv = W.dQKV;
        a.bias = m->dev<float>(blk.bias_off);
        a.dB = W.dB;
        a.dtable = grads[bt.table_i];
        a.B = B;
        a.H = Hp;
        a.W = Wp;
        a.shift = (bi % 2 == 0) ? 0 : c.window_size / 2;
        a.heads = L.heads;
        a.d = L.d;
        a.DP = L.DP;
        if (L.heads * L.DP < L.QP) SSR_CUDA(cudaMemsetAsync(W.dQKV, 0, (size_t)T * 3 * L.QP * 2, s));
        SSR_TRY(launch_attn_bwd(a, s));
      }
      SSR_TRY(wgrad_lin(m, bt.qkv, W.dQKV, bw.xn1, T, W.dwp, W.partial, grads, s));
      {
        GemmArgs g = dgrad_lin(m, bt.qkv, W.dQKV, T);
        g.out_T = W.dXn;
        g.ld_T = CP;
        SSR_TRY(launch_gemm_tc(g, 2, s));
      }
      // the bf16 copy feeds the MLP branch of the previous block (unused for the first block of a layer)
      SSR_TRY(ln_backward(t, bt.n1, m->dev<float>(blk.norm1.g_off), bw.tin, W.dXn, 2, W.Gt, W.Gt, W.Gtb, T, C, CP, W.partial, grads, s,
                          (drop && bi > 0) ? drop + (size_t)(2 * (kblk - 1) + 1) * B : nullptr, Hp * Wp,
                          (ln_sums && bi > 0) ? grads[t->s_blocks[li][bi - 1].fc2.bi] : nullptr));
    }
    SSR_TRY(launch_add_inplace(W.G, W.Gt, W.Gb, (size_t)T * CP, s));  // blocks' path joins the group skip
  }
  // g0 = patch_embed.norm(x0) (swinir.py:22-32, 343-344); x0 also receives the long skip Gres
  SSR_TRY(ln_backward(t, t->s_pe, m->dev<float>(m->pe_norm.g_off), W.x0, W.G, 4, W.Gres, W.Gt, W.Gtb, T, C, CP, W.partial, grads, s));
  if (grads[t->head_w]) {
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.dY = W.Gtb;
    a.ldy = CP;
    a.X = W.xin64;
    a.ldx = 64;
    a.B = B;
    a.H = Hp;
    a.W = Wp;
    a.M = T;
    a.taps = 9;
    a.NoutP = CP;
    a.CinP = 64;
    a.dWp = W.dwp + t->head_dwp;
    a.dBp = nullptr;
    a.alpha = 1.0f;
    a.N_alg = C;
    a.K_alg = 3;
    SSR_TRY(launch_wgrad_tc(a, s));
    PackEntry e;
    memset(&e, 0, sizeof(e));
    e.W = W.dwp + t->head_dwp;
    e.Wf = grads[t->head_w];
    e.bf = a.dBp;
    e.b = grads[t->head_b];
    e.kind = 0;
    e.N = C;
    e.K = 3;
    e.KP = 64;
    e.taps = 9;
    t->unpack_host.push_back(e);
  }
  if (grads[t->head_b]) SSR_TRY(launch_colsum(W.Gtb, 2, CP, T, CP, C, 0, 1.0f, grads[t->head_b], W.partial, s, &t->red));
  t->dx_G = W.Gt; t->dx_ld = CP; t->dx_C = C; t->dx_B = B; t->dx_h = h; t->dx_w = w; t->dx_Hp = Hp; t->dx_Wp = Wp;
  t->dx_scale = 1.0f / m->cfg.img_range;  // Normalizer (common.py:222-233): x_in = (x - mean * range) / range
  SSR_TRY(launch_unpack_batched(t->unpack_host.data(), W.pack_dev, (int)t->unpack_host.size(), s));
  return launch_deferred_reductions(&t->red, s);  // bias / LayerNorm-parameter gradients: all second stages, one launch
}

}  // namespace ssr

// =============================================================================================
extern "C" {

int ssr_model_train_bind(ssr_model_t* m, int n, const char* const* names, const int64_t* numels) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(m->cfg.precision == SSR_PREC_BF16, SSR_E_INVALID, "the training path is built for the bf16 tensor-core precision only");
  SSR_CHECK(m->cfg.arch == SSR_ARCH_EDSR || m->cfg.arch == SSR_ARCH_SWINIR || m->cfg.arch == SSR_ARCH_RCAN || m->cfg.arch == SSR_ARCH_HAN,
            SSR_E_INVALID, "the training path is built for EDSR, RCAN, HAN and SwinIR (HAT and SwinFIR are inference-only)");
  train_state_destroy(m);
  m->train = new ssr_train_state();
  for (int i = 0; i < n; ++i) {
    m->train->names.push_back(names[i]);
    m->train->numels.push_back(numels[i]);
    m->train->index[names[i]] = i;
  }
  const bool rcan_like = m->cfg.arch == SSR_ARCH_RCAN || m->cfg.arch == SSR_ARCH_HAN;
  int r = m->cfg.arch == SSR_ARCH_EDSR ? bind_edsr(m) : rcan_like ? bind_rcan(m) : bind_swinir(m);
  if (r != SSR_OK) train_state_destroy(m);
  return r;
}

size_t ssr_model_train_workspace_bytes(const ssr_model_t* m, int B, int H, int W) {
  if (!m || !m->train) return 0;
  if (m->cfg.arch == SSR_ARCH_SWINIR) {
    int Hp, Wp;
    train_padded(m, H, W, &Hp, &Wp);
    SwinTrainWs w;
    return plan_swin_train(m, nullptr, B, Hp, Wp, &w);
  }
  if (m->cfg.arch == SSR_ARCH_RCAN || m->cfg.arch == SSR_ARCH_HAN) {
    RcanTrainWs w;
    return plan_rcan_train(m, nullptr, B, H, W, &w);
  }
  EdsrTrainWs w;
  return plan_edsr_train(m, nullptr, B, H, W, &w);
}

int ssr_model_train_forward(ssr_model_t* m, const float* const* params, const float* drop_scale, const float* x, float* y, int B,
                            int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(m->train != nullptr, SSR_E_STATE, "ssr_model_train_bind has not been called");
  SSR_CHECK(params && x && y && B > 0 && H > 0 && W > 0, SSR_E_INVALID, "train_forward: bad argument");
  if (m->cfg.arch == SSR_ARCH_SWINIR)
    return train_forward_swinir(m, params, drop_scale, x, y, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
  SSR_CHECK(drop_scale == nullptr, SSR_E_INVALID, "train_forward: drop_scale is a SwinIR option");
  if (m->cfg.arch == SSR_ARCH_RCAN || m->cfg.arch == SSR_ARCH_HAN) return train_forward_rcan(m, params, x, y, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
  return train_forward_edsr(m, params, x, y, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ssr_model_train_backward(ssr_model_t* m, const float* dy, const float* drop_scale, float* const* grads, int B, int H, int W,
                             void* workspace, size_t workspace_bytes, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(m->train != nullptr, SSR_E_STATE, "ssr_model_train_bind has not been called");
  SSR_CHECK(dy && grads && B > 0 && H > 0 && W > 0, SSR_E_INVALID, "train_backward: bad argument");
  if (m->cfg.arch == SSR_ARCH_SWINIR)
    return train_backward_swinir(m, dy, drop_scale, grads, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
  if (m->cfg.arch == SSR_ARCH_RCAN || m->cfg.arch == SSR_ARCH_HAN)
    return train_backward_rcan(m, nullptr, dy, grads, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
  return train_backward_edsr(m, dy, grads, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ssr_model_train_input_grad(ssr_model_t* m, float* dx, int B, int H, int W, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(m->train != nullptr && m->train->dx_G != nullptr, SSR_E_STATE, "train_input_grad: no ssr_model_train_backward has run");
  const ssr_train_state* t = m->train;
  SSR_CHECK(dx && B == t->dx_B && H == t->dx_h && W == t->dx_w, SSR_E_INVALID, "train_input_grad: shape [%d,3,%d,%d] is not the last backward's [%d,3,%d,%d]",
            B, H, W, t->dx_B, t->dx_h, t->dx_w);
  return launch_conv_first_dgrad(t->dx_G, t->dx_ld, m->dev<float>(m->conv_first_w), t->dx_C, B, H, W, t->dx_Hp, t->dx_Wp, t->dx_scale, dx,
                                 (cudaStream_t)stream);
}

}  // extern "C"
