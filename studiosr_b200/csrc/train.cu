// Training executor: forward with saved activations + backward (data and weight gradients) behind
// the C ABI, bf16 tensor-core path.  Replaces what `loss.backward()` (trainer.py:104) runs through torch autograd for
// the model: the reference has no hand-written backward, so every formula below is the adjoint of the cited forward.
// The fp32 master parameters stay in PyTorch-owned device memory; each train_forward re-packs them on the device.
#include <string.h>

#include "ssr_model.cuh"

using namespace ssr;

namespace ssr {

struct ConvT {       // a trainable conv3x3 executed by the implicit-GEMM kernels
  const Lin* fwd;    // forward pack (m->arena)
  int wi = -1, bi = -1;  // bound-parameter indices
  int Cout = 0, Cin = 0;
  size_t dg_off = 0;     // dgrad pack in train->arena2: bf16 [KP][9*NP]
  size_t dwp_off = 0;    // packed fp32 weight gradient [NP][9][KP] (float offset into the dwp workspace block)
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
  // EDSR
  int head_w = -1, head_b = -1;
  size_t head_dwp = 0;
  std::vector<int> e_res_a, e_res_b, e_up;  // indices into convs
  int e_body_tail = -1, e_last = -1;
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
  return c.off + 1024;
}

static int repack_conv(ssr_model* m, const ConvT& c, const float* const* params, cudaStream_t s) {
  const Lin& L = *c.fwd;
  return launch_pack_conv_dev(params[c.wi], params[c.bi], m->arena + L.w_off, m->dev<float>(L.b_off), m->train->arena2 + c.dg_off,
                              c.Cout, c.Cin, L.NP, L.KP, 9, L.ps_r, s);
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
  SSR_CUDA(cudaMemcpyAsync(m->dev<float>(m->conv_first_w), params[t->head_w], (size_t)m->F * 27 * 4, cudaMemcpyDeviceToDevice, s));
  SSR_CUDA(cudaMemcpyAsync(m->dev<float>(m->conv_first_b), params[t->head_b], (size_t)m->F * 4, cudaMemcpyDeviceToDevice, s));
  for (const ConvT& cv : t->convs) SSR_TRY(repack_conv(m, cv, params, s));
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
                      float* dwp, float* const* grads, cudaStream_t s) {
  const Lin& L = *c.fwd;
  if (grads[c.wi]) {
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.dY = dY;
    a.ldy = L.NP;
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
    a.alpha = alpha;
    a.N_alg = c.Cout;
    a.K_alg = c.Cin;
    SSR_TRY(launch_wgrad_tc(a, s));
    SSR_TRY(launch_unpack_wgrad(dwp + c.dwp_off, grads[c.wi], c.Cout, c.Cin, L.KP, 9, L.ps_r, s));
  }
  if (grads[c.bi]) SSR_TRY(launch_colsum(dY, 2, L.NP, B * H * W, c.Cout, L.ps_r, alpha, grads[c.bi], s));
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
  SSR_CUDA(cudaMemsetAsync(W.dwp, 0, t->dwp_floats * 4, s));
  int H = h * c.scale, Wd = w * c.scale;
  // add_mean is a frozen identity-weight 1x1 conv (common.py:108-121): dL/d(tail.1 output) = dy
  SSR_TRY(launch_nchw_to_nhwc(dy, W.dy64, B, 3, H, Wd, 64, 2, 0, s));
  const int nup = (int)m->up.size();
  {  // tail.1 (edsr.py:37,46)
    const ConvT& cv = t->convs[t->e_last];
    const void* X = nup ? W.hr[nup - 1] : W.bt;
    SSR_TRY(wgrad_conv(m, cv, W.dy64, X, FP, B, H, Wd, 1.0f, W.dwp, grads, s));
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
    SSR_TRY(wgrad_conv(m, cv, W.gU, X, FP, B, H, Wd, 1.0f, W.dwp, grads, s));
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
    SSR_TRY(wgrad_conv(m, cv, W.Gb, W.rb[nb], FP, B, h, w, 1.0f, W.dwp, grads, s));
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
    SSR_TRY(wgrad_conv(m, cb, Gb, W.tmp[i], FP, B, h, w, c.res_scale, W.dwp, grads, s));
    GemmArgs gb = dgrad_base(m, cb, Gb, B, h, w);
    gb.alpha = c.res_scale;
    gb.mask = W.tmp[i];  // ReLU backward: gate by the saved ReLU output
    gb.ld_mask = FP;
    gb.mask_slope = 0.0f;
    gb.out_T = Dh;
    gb.ld_T = FP;
    SSR_TRY(launch_gemm_tc(gb, 2, s));
    SSR_TRY(wgrad_conv(m, ca, Dh, W.rb[i], FP, B, h, w, 1.0f, W.dwp, grads, s));
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
    a.alpha = 1.0f;
    a.N_alg = m->F;
    a.K_alg = 3;
    SSR_TRY(launch_wgrad_tc(a, s));
    SSR_TRY(launch_unpack_wgrad(W.dwp + t->head_dwp, grads[t->head_w], m->F, 3, 64, 9, 0, s));
  }
  if (grads[t->head_b]) SSR_TRY(launch_colsum(Gb, 2, FP, (int)T, m->F, 0, 1.0f, grads[t->head_b], s));
  return SSR_OK;
}

}  // namespace ssr

// =============================================================================================
extern "C" {

int ssr_model_train_bind(ssr_model_t* m, int n, const char* const* names, const int64_t* numels) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(m->cfg.precision == SSR_PREC_BF16, SSR_E_INVALID, "the training path is built for the bf16 tensor-core precision only");
  SSR_CHECK(m->cfg.arch == SSR_ARCH_EDSR, SSR_E_INVALID, "the training path is built for EDSR so far");
  train_state_destroy(m);
  m->train = new ssr_train_state();
  for (int i = 0; i < n; ++i) {
    m->train->names.push_back(names[i]);
    m->train->numels.push_back(numels[i]);
    m->train->index[names[i]] = i;
  }
  int r = bind_edsr(m);
  if (r != SSR_OK) train_state_destroy(m);
  return r;
}

size_t ssr_model_train_workspace_bytes(const ssr_model_t* m, int B, int H, int W) {
  if (!m || !m->train) return 0;
  EdsrTrainWs w;
  return plan_edsr_train(m, nullptr, B, H, W, &w);
}

int ssr_model_train_forward(ssr_model_t* m, const float* const* params, const float* x, float* y, int B, int H, int W,
                            void* workspace, size_t workspace_bytes, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(m->train != nullptr, SSR_E_STATE, "ssr_model_train_bind has not been called");
  SSR_CHECK(params && x && y && B > 0 && H > 0 && W > 0, SSR_E_INVALID, "train_forward: bad argument");
  return train_forward_edsr(m, params, x, y, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ssr_model_train_backward(ssr_model_t* m, const float* dy, float* const* grads, int B, int H, int W, void* workspace,
                             size_t workspace_bytes, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(m->train != nullptr, SSR_E_STATE, "ssr_model_train_bind has not been called");
  SSR_CHECK(dy && grads && B > 0 && H > 0 && W > 0, SSR_E_INVALID, "train_backward: bad argument");
  return train_backward_edsr(m, dy, grads, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
