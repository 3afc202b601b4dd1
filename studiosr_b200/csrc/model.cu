// Host side of libssr_b200: model handle, weight packing, workspace planning and the launch
// sequences of SwinIR (swinir.py:353-372) and EDSR (edsr.py:39-48), plus the C ABI.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "ssr_model.cuh"

namespace ssr {

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- profiling -------------------------------------------------------------------------------
struct ProfRec {
  const char* cls;
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::vector<ProfRec*> g_prof;

ProfScope::ProfScope(const char* cls, double flops, double bytes, cudaStream_t s) : stream(s) {
  if (!g_prof_on) return;
  ProfRec* r = new ProfRec{cls, flops, bytes, nullptr, nullptr};
  cudaEventCreate(&r->e0);
  cudaEventCreate(&r->e1);
  cudaEventRecord(r->e0, s);
  rec = r;
}
ProfScope::~ProfScope() {
  if (!rec) return;
  ProfRec* r = reinterpret_cast<ProfRec*>(rec);
  cudaEventRecord(r->e1, stream);
  g_prof.push_back(r);
}

double gemm_alg_flops(const GemmArgs& g) { return 2.0 * g.M * (double)g.N_alg * g.K_alg * g.taps; }
double gemm_alg_bytes(const GemmArgs& g, int elem) {
  double b = (double)g.M * g.K_alg * elem + (double)g.N_alg * g.K_alg * g.taps * elem;
  if (g.res) b += (double)g.M * g.N_alg * 4;
  if (g.out_f32) b += (double)g.M * g.N_alg * 4;
  if (g.out_T) b += (double)g.M * g.N_alg * elem;
  if (g.out_ln) b += (double)g.M * g.N_alg * elem;
  if (g.out3_f32) b += (double)g.B * g.crop_h * g.crop_w * 3 * 4;
  if (g.out3_u8) b += (double)g.B * g.crop_h * g.crop_w * 3;
  return b;
}

static const float kRgbMean[3] = {0.4488f, 0.4371f, 0.4040f};  // common.py:223 / :111

static inline float tf32_round_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) {
    u += 0x00001000u;
    u &= 0xffffe000u;
  }
  memcpy(&x, &u, 4);
  return x;
}

}  // namespace ssr

using namespace ssr;

namespace ssr {

// ---------------------------------------------------------------------------------------------
// packing helpers (host)
size_t arena_alloc(ssr_model* m, size_t bytes) {
  size_t off = (m->host_arena.size() + 255) & ~(size_t)255;
  m->host_arena.resize(off + bytes, 0);
  return off;
}

// element `idx` of a packed weight matrix of `total` elements; the 3xTF32 mode stores the tf32 head there and the tf32 tail
// `total` elements further on (the second half of a double-size allocation, see w_bytes)
static void store_w(ssr_model* m, size_t off, size_t idx, float v, size_t total) {
  if (m->elem == 2) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    memcpy(m->host_arena.data() + off + idx * 2, &h, 2);
  } else if (m->cfg.precision == SSR_PREC_TF32X3) {
    const float hi = tf32_round_host(v), lo = tf32_round_host(v - hi);
    memcpy(m->host_arena.data() + off + idx * 4, &hi, 4);
    memcpy(m->host_arena.data() + off + (total + idx) * 4, &lo, 4);
  } else {
    if (m->cfg.precision == SSR_PREC_TF32) v = tf32_round_host(v);
    memcpy(m->host_arena.data() + off + idx * 4, &v, 4);
  }
}
static size_t w_bytes(const ssr_model* m, size_t elems) { return elems * m->elem * (m->cfg.precision == SSR_PREC_TF32X3 ? 2 : 1); }

static const std::vector<float>* find_param(ssr_model* m, const std::string& name, size_t numel) {
  auto it = m->params.find(name);
  if (it == m->params.end()) {
    set_error("missing parameter '%s'", name.c_str());
    return nullptr;
  }
  if (it->second.size() != numel) {
    set_error("parameter '%s' has %zu elements, expected %zu", name.c_str(), it->second.size(), numel);
    return nullptr;
  }
  return &it->second;
}

static size_t pack_vec(ssr_model* m, const std::vector<float>& v, int NP, float scale = 1.0f) {
  size_t off = arena_alloc(m, (size_t)NP * 4);
  float* dst = reinterpret_cast<float*>(m->host_arena.data() + off);
  for (size_t i = 0; i < v.size(); ++i) dst[i] = v[i] * scale;
  return off;
}

// nn.Linear [N][K] -> [NP][KP]; row_src(n') / col_src(k') give the source index or -1 (zero)
template <typename RowF, typename ColF, typename ScaleF>
static int pack_linear(ssr_model* m, const std::string& name, int N, int K, int NP, int KP, RowF row_src, ColF col_src,
                       ScaleF row_scale, Lin* out) {
  const std::vector<float>* W = find_param(m, name + ".weight", (size_t)N * K);
  const std::vector<float>* B = find_param(m, name + ".bias", (size_t)N);
  if (!W || !B) return SSR_E_STATE;
  out->K = K; out->KP = KP; out->N = N; out->NP = NP; out->taps = 1; out->ps_r = 0; out->N_alg = N;
  out->w_off = arena_alloc(m, w_bytes(m, (size_t)NP * KP));
  out->b_off = arena_alloc(m, (size_t)NP * 4);
  float* bd = reinterpret_cast<float*>(m->host_arena.data() + out->b_off);
  for (int n = 0; n < NP; ++n) {
    const int sn = row_src(n);
    if (sn < 0) continue;
    const float sc = row_scale(n);
    bd = reinterpret_cast<float*>(m->host_arena.data() + out->b_off);
    bd[n] = (*B)[sn] * sc;
    for (int k = 0; k < KP; ++k) {
      const int sk = col_src(k);
      if (sk >= 0) store_w(m, out->w_off, (size_t)n * KP + k, (*W)[(size_t)sn * K + sk] * sc, (size_t)NP * KP);
    }
  }
  return SSR_OK;
}

// nn.Conv2d [Cout][Cin][3][3] -> [NP][9*KP] (k = tap*KP + c); ps_r > 1 permutes the output rows to
// (i, j, c) order so the pixel-shuffled store is channel-contiguous.
static int pack_conv(ssr_model* m, const std::string& name, int Cout, int Cin, int ps_r, Lin* out) {
  const std::vector<float>* W = find_param(m, name + ".weight", (size_t)Cout * Cin * 9);
  const std::vector<float>* B = find_param(m, name + ".bias", (size_t)Cout);
  if (!W || !B) return SSR_E_STATE;
  const int NP = round_up(Cout, 64), KP = round_up(Cin, 64);
  out->K = Cin; out->KP = KP; out->N = Cout; out->NP = NP; out->taps = 9; out->ps_r = ps_r; out->N_alg = Cout;
  out->w_off = arena_alloc(m, w_bytes(m, (size_t)NP * 9 * KP));
  out->b_off = arena_alloc(m, (size_t)NP * 4);
  const int rr = ps_r > 1 ? ps_r * ps_r : 1, Cps = Cout / rr;
  for (int n = 0; n < Cout; ++n) {
    int sn = n;
    if (ps_r > 1) {
      const int q = n / Cps, c = n % Cps;
      sn = c * rr + q;  // q = i*r + j
    }
    reinterpret_cast<float*>(m->host_arena.data() + out->b_off)[n] = (*B)[sn];
    for (int tap = 0; tap < 9; ++tap)
      for (int c = 0; c < Cin; ++c)
        store_w(m, out->w_off, (size_t)n * 9 * KP + (size_t)tap * KP + c, (*W)[((size_t)sn * Cin + c) * 9 + tap], (size_t)NP * 9 * KP);
  }
  return SSR_OK;
}

static int pack_ln(ssr_model* m, const std::string& name, int C, int CP, LNp* out) {
  const std::vector<float>* g = find_param(m, name + ".weight", (size_t)C);
  const std::vector<float>* b = find_param(m, name + ".bias", (size_t)C);
  if (!g || !b) return SSR_E_STATE;
  out->g_off = pack_vec(m, *g, CP);
  out->b_off = pack_vec(m, *b, CP);
  return SSR_OK;
}

void upsampler_plan(int scale, std::vector<int>* rs) {  // common.py:124-137
  rs->clear();
  if ((scale & (scale - 1)) == 0) {
    for (int s = scale; s > 1; s >>= 1) rs->push_back(2);
  } else {
    rs->push_back(scale);
  }
}

static int pack_conv_last(ssr_model* m, const std::string& name, int Cin, size_t* w_off, float* bias3) {
  const std::vector<float>* W = find_param(m, name + ".weight", (size_t)3 * Cin * 9);
  const std::vector<float>* B = find_param(m, name + ".bias", 3);
  if (!W || !B) return SSR_E_STATE;
  *w_off = arena_alloc(m, (size_t)9 * Cin * 4 * 4);
  float* dst = reinterpret_cast<float*>(m->host_arena.data() + *w_off);
  for (int tap = 0; tap < 9; ++tap)
    for (int c = 0; c < Cin; ++c)
      for (int co = 0; co < 3; ++co) dst[((size_t)tap * Cin + c) * 4 + co] = (*W)[((size_t)co * Cin + c) * 9 + tap];
  for (int i = 0; i < 3; ++i) bias3[i] = (*B)[i];
  m->last_w27 = 0;
  if (Cin == 64 && m->cfg.precision == SSR_PREC_BF16) {  // taps-in-N operand of k_conv_last.cu
    m->last_w27 = arena_alloc(m, 32 * 64 * 2);
    for (int n = 0; n < 32; ++n)
      for (int c = 0; c < 64; ++c) {
        const float v = n < 27 ? (*W)[((size_t)(n % 3) * Cin + c) * 9 + n / 3] : 0.0f;
        const __nv_bfloat16 h = __float2bfloat16(v);
        memcpy(m->host_arena.data() + m->last_w27 + ((size_t)n * 64 + c) * 2, &h, 2);
      }
  }
  return SSR_OK;
}

static int pack_conv_first(ssr_model* m, const std::string& name, int Cout, size_t* w_off, size_t* b_off) {
  const std::vector<float>* W = find_param(m, name + ".weight", (size_t)Cout * 27);
  const std::vector<float>* B = find_param(m, name + ".bias", (size_t)Cout);
  if (!W || !B) return SSR_E_STATE;
  *w_off = pack_vec(m, *W, Cout * 27);
  *b_off = pack_vec(m, *B, Cout);
  return SSR_OK;
}

// The fused Swin tail kernel (k_swin_tail.cu) is used for bf16 models of the 180/6/360 class.
static bool fused_tail(const ssr_model* m, const Layer& L) {
  return m->cfg.precision == SSR_PREC_BF16 && m->CP == 192 && m->HP == 384 && L.QP == 192;
}

// The fused QKV + window-attention kernel (k_swin_attn.cu): same class, 6 heads, 8x8 windows.
static bool fused_attn(const ssr_model* m, const Layer& L) {
  return fused_tail(m, L) && L.heads == 6 && L.DP == 32 && m->cfg.window_size == 8 && getenv("STUDIOSR_B200_UNFUSED_ATTN") == nullptr;
}

// Re-pack `lin` (already packed un-folded by pack_linear) with the preceding LayerNorm's affine folded in:
// W'[n][k] = W[n][k] * gamma[k], b'[n] = b[n] + sum_k W[n][k] * beta[k]; the kernel then only normalises.
static int fold_norm_into_linear(ssr_model* m, const std::string& norm, const std::string& lin, int N, int K, int KP, const Lin& out) {
  const std::vector<float>* W = find_param(m, lin + ".weight", (size_t)N * K);
  const std::vector<float>* B = find_param(m, lin + ".bias", (size_t)N);
  const std::vector<float>* g = find_param(m, norm + ".weight", (size_t)K);
  const std::vector<float>* be = find_param(m, norm + ".bias", (size_t)K);
  if (!W || !B || !g || !be) return SSR_E_STATE;
  for (int n = 0; n < N; ++n) {
    double acc = (*B)[n];
    for (int k = 0; k < K; ++k) {
      const float w = (*W)[(size_t)n * K + k];
      store_w(m, out.w_off, (size_t)n * KP + k, w * (*g)[k], (size_t)out.NP * KP);  // bf16 fused path only
      acc += (double)w * (*be)[k];
    }
    reinterpret_cast<float*>(m->host_arena.data() + out.b_off)[n] = (float)acc;
    // the fused tail kernel takes the bias from the GEMM itself: its A operand is 1.0 in channels K, K+1
    SSR_CHECK(K + 2 <= KP, SSR_E_INVALID, "fold_norm_into_linear: K=%d leaves no pad channel pair for the bias", K);
    const float hi = __bfloat162float(__float2bfloat16_rn((float)acc));
    store_w(m, out.w_off, (size_t)n * KP + K, hi, (size_t)out.NP * KP);
    store_w(m, out.w_off, (size_t)n * KP + K + 1, (float)acc - hi, (size_t)out.NP * KP);
  }
  return SSR_OK;
}

int pack_attn_fused_host(const float* Wqkv, const float* bqkv, const float* table, int C, int heads, void* Whp, void* bias_tab) {
  SSR_CHECK(heads == 6 && C % heads == 0 && C / heads <= 32 && C + kAttnOnesChannels <= 192, SSR_E_INVALID,
            "fused attention packing: C=%d heads=%d unsupported", C, heads);
  const int d = C / heads;
  const float log2e = 1.4426950408889634f, qs = log2e / sqrtf((float)d);
  __nv_bfloat16* W = reinterpret_cast<__nv_bfloat16*>(Whp);
  for (int hp = 0; hp < 3; ++hp)
    for (int n = 0; n < 192; ++n) {
      const int sec = n / 64, hh = (n % 64) / 32, j = n % 32, head = 2 * hp + hh;
      const bool real = j < d;
      const int src = sec * C + head * d + j;
      const float sc = sec == 0 ? qs : 1.0f;
      __nv_bfloat16* row = W + ((size_t)hp * 192 + n) * 192;
      for (int k = 0; k < 192; ++k) row[k] = __float2bfloat16_rn(real && k < C ? Wqkv[(size_t)src * C + k] * sc : 0.0f);
      // the bias rides in the GEMM: the A operand (norm1's output) is 1.0 in channels C, C+1; hi + lo keeps 16 mantissa bits
      const float b = real ? bqkv[src] * sc : 0.0f;
      row[C] = __float2bfloat16_rn(b);
      row[C + 1] = __float2bfloat16_rn(b - __bfloat162float(row[C]));
    }
  // relative-position bias (swinir.py:57-67, 92-95), compact: B[i][j] = T[yi - yj + 7][xi - xj + 7] = R[7 - yi + yj][7 - xi + xj]
  // with R the table reversed in both axes, so the 8 x 8 keys of a window seen from token (yi, xi) are an 8 x 8 sub-block of R
  // and every key quad (4 consecutive xj) is a contiguous run.  Two copies, shifted by 0 / 1 element, make every run start
  // 4-byte aligned: copy s serves tokens with (7 - xi) % 2 == s and stores R[a][b] at [a][b + s], row pitch 16.
  __nv_bfloat16* bt = reinterpret_cast<__nv_bfloat16*>(bias_tab);
  for (size_t i = 0; i < kAttnBiasBytes / 2; ++i) bt[i] = __float2bfloat16_rn(0.0f);
  for (int hd = 0; hd < heads; ++hd)
    for (int sft = 0; sft < 2; ++sft)
      for (int a = 0; a < 15; ++a)
        for (int b = 0; b < 15; ++b)
          bt[((size_t)(hd * 2 + sft) * 15 + a) * 16 + b + sft] =
              __float2bfloat16_rn(table[(size_t)((14 - a) * 15 + (14 - b)) * heads + hd] * log2e);
  return SSR_OK;
}

static bool is_swin(const ssr_model* m) { return m->cfg.arch == SSR_ARCH_SWINIR || m->cfg.arch == SSR_ARCH_SWINFIR; }

// SFB (swinfir.py:68-80).  The 1x1 convs are Linears over pixels / spectrum bins; the fusion conv reads the concatenation
// [S(x) | F(x)] whose halves are CP columns apart in the workspace, hence its column map.
static int pack_sfb(ssr_model* m, const std::string& pre, Sfb* S) {
  const int C = m->C, CP = m->CP, c2 = C / 2, c2P = round_up(c2, 64);
  SSR_CHECK(C % 2 == 0, SSR_E_INVALID, "SwinFIR: embed_dim %d is odd", C);
  auto one = [](int) { return 1.0f; };
  SSR_TRY(pack_conv(m, pre + ".S.body.0", C, C, 0, &S->s0));
  SSR_TRY(pack_conv(m, pre + ".S.body.2", C, C, 0, &S->s2));
  SSR_TRY(pack_linear(m, pre + ".F.conv_before_fft.0", c2, C, c2P, CP, [=](int n) { return n < c2 ? n : -1; },
                      [=](int k) { return k < C ? k : -1; }, one, &S->before));
  SSR_TRY(pack_linear(m, pre + ".F.fu.conv_layer", C, C, CP, CP, [=](int n) { return n < C ? n : -1; }, [=](int k) { return k < C ? k : -1; },
                      one, &S->fu));
  SSR_TRY(pack_linear(m, pre + ".F.conv_after_fft", C, c2, CP, c2P, [=](int n) { return n < C ? n : -1; },
                      [=](int k) { return k < c2 ? k : -1; }, one, &S->after));
  SSR_TRY(pack_linear(m, pre + ".fusion", C, 2 * C, CP, 2 * CP, [=](int n) { return n < C ? n : -1; },
                      [=](int k) { return k < C ? k : (k >= CP && k - CP < C ? C + k - CP : -1); }, one, &S->fusion));
  return SSR_OK;
}

static int finalize_swinir(ssr_model* m) {
  const ssr_model_config& c = m->cfg;
  SSR_CHECK(c.n_colors == 3, SSR_E_INVALID, "n_colors must be 3");
  SSR_CHECK(c.window_size == 8 || c.precision == SSR_PREC_FP32, SSR_E_INVALID,
            "tensor-core window attention supports window_size 8 (got %d)", c.window_size);
  const int C = c.embed_dim;
  m->sfb = c.arch == SSR_ARCH_SWINFIR;
  SSR_CHECK(!m->sfb || c.precision != SSR_PREC_BF16, SSR_E_INVALID,
            "SwinFIR runs in the fp32-class precisions (fp32 / tf32 / tf32x3): the reference trains and ships it in fp32 (swinfir.py:126)");
  m->C = C;
  m->CP = round_up(C, 64);
  m->HID = (int)(C * c.mlp_ratio);
  m->HP = round_up(m->HID, 64);
  SSR_CHECK(m->CP <= 256, SSR_E_INVALID, "embed_dim %d > 256 not supported", C);
  const int nb = (2 * c.window_size - 1) * (2 * c.window_size - 1);
  SSR_TRY(pack_conv_first(m, "conv_first", C, &m->conv_first_w, &m->conv_first_b));
  SSR_TRY(pack_ln(m, "patch_embed.norm", C, m->CP, &m->pe_norm));
  m->layers.clear();
  m->QPmax = 0;
  for (int li = 0; li < c.n_layers; ++li) {
    Layer L;
    L.heads = c.num_heads[li];
    SSR_CHECK(L.heads > 0 && C % L.heads == 0, SSR_E_INVALID, "embed_dim %d not divisible by heads %d", C, L.heads);
    L.d = C / L.heads;
    SSR_CHECK(L.d <= 32, SSR_E_INVALID, "head_dim %d > 32 not supported", L.d);
    L.DP = L.d <= 16 ? 16 : 32;
    L.QP = round_up(L.heads * L.DP, 64);
    if (L.QP > m->QPmax) m->QPmax = L.QP;
    const int d = L.d, DP = L.DP, QP = L.QP, heads = L.heads;
    const float qscale = 1.0f / sqrtf((float)d);
    for (int bi = 0; bi < c.depths[li]; ++bi) {
      Block B;
      char pre[128];
      snprintf(pre, sizeof(pre), "layers.%d.residual_group.blocks.%d", li, bi);
      const std::string p(pre);
      SSR_TRY(pack_ln(m, p + ".norm1", C, m->CP, &B.norm1));
      SSR_TRY(pack_ln(m, p + ".norm2", C, m->CP, &B.norm2));
      // qkv rows: n' = part*QP + h*DP + j  <-  part*C + h*d + j ; q rows scaled by d^-0.5 (swinir.py:83)
      auto qkv_row = [=](int n) {
        const int part = n / QP, hc = n % QP, h = hc / DP, j = hc % DP;
        return (h < heads && j < d) ? part * C + h * d + j : -1;
      };
      auto qkv_scale = [=](int n) { return n < QP ? qscale : 1.0f; };
      auto ident_c = [=](int k) { return k < C ? k : -1; };
      auto one = [](int) { return 1.0f; };
      SSR_TRY(pack_linear(m, p + ".attn.qkv", 3 * C, C, 3 * QP, m->CP, qkv_row, ident_c, qkv_scale, &B.qkv));
      B.qkv.N = 3 * QP;  // all padded columns are produced (zeros) so the attention kernel can read them
      auto proj_col = [=](int k) {
        const int h = k / DP, j = k % DP;
        return (h < heads && j < d) ? h * d + j : -1;
      };
      SSR_TRY(pack_linear(m, p + ".attn.proj", C, C, m->CP, QP, ident_c, proj_col, one, &B.proj));
      const int HID = m->HID;
      auto ident_h = [=](int k) { return k < HID ? k : -1; };
      SSR_TRY(pack_linear(m, p + ".mlp.fc1", HID, C, m->HP, m->CP, ident_h, ident_c, one, &B.fc1));
      if (fused_tail(m, L)) SSR_TRY(fold_norm_into_linear(m, p + ".norm2", p + ".mlp.fc1", HID, C, m->CP, B.fc1));
      SSR_TRY(pack_linear(m, p + ".mlp.fc2", C, HID, m->CP, m->HP, ident_c, ident_h, one, &B.fc2));
      // relative position bias table [(2ws-1)^2][heads] -> [heads][(2ws-1)^2]
      const std::vector<float>* T = find_param(m, p + ".attn.relative_position_bias_table", (size_t)nb * heads);
      if (!T) return SSR_E_STATE;
      B.bias_off = arena_alloc(m, (size_t)heads * nb * 4);
      float* bt = reinterpret_cast<float*>(m->host_arena.data() + B.bias_off);
      for (int h = 0; h < heads; ++h)
        for (int i = 0; i < nb; ++i) bt[h * nb + i] = (*T)[(size_t)i * heads + h];
      if (fused_attn(m, L)) {  // operands of k_swin_attn.cu: head-pair ordered qkv weights, pi-ordered bias table
        const std::vector<float>* Wq = find_param(m, p + ".attn.qkv.weight", (size_t)3 * C * C);
        const std::vector<float>* Bq = find_param(m, p + ".attn.qkv.bias", (size_t)3 * C);
        if (!Wq || !Bq) return SSR_E_STATE;
        B.whp_off = arena_alloc(m, kAttnWhpBytes);
        B.btab_off = arena_alloc(m, kAttnBiasBytes);
        uint8_t* base = m->host_arena.data();
        SSR_TRY(pack_attn_fused_host(Wq->data(), Bq->data(), T->data(), C, heads, base + B.whp_off, base + B.btab_off));
        // norm1 (gamma pad = 0) writes its beta into the pad channels: 1.0 in C, C+1 switches the bias columns of Whp on
        float* beta = reinterpret_cast<float*>(base + B.norm1.b_off);
        for (int k = 0; k < kAttnOnesChannels; ++k) beta[C + k] = 1.0f;
      }
      L.blocks.push_back(B);
    }
    char nm[64];
    snprintf(nm, sizeof(nm), "layers.%d.conv", li);
    if (m->sfb)
      SSR_TRY(pack_sfb(m, nm, &L.sfb));
    else
      SSR_TRY(pack_conv(m, nm, C, C, 0, &L.conv));
    m->layers.push_back(L);
  }
  SSR_TRY(pack_ln(m, "norm", C, m->CP, &m->final_norm));
  if (m->sfb)
    SSR_TRY(pack_sfb(m, "conv_after_body", &m->sfb_after_body));
  else
    SSR_TRY(pack_conv(m, "conv_after_body", C, C, 0, &m->conv_after_body));
  std::vector<int> rs;
  upsampler_plan(c.scale, &rs);
  m->up.clear();
  if (c.upsampler == 0) {
    SSR_TRY(pack_conv(m, "conv_before_upsample.0", 64, C, 0, &m->conv_before_up));
    for (size_t i = 0; i < rs.size(); ++i) {
      Lin L;
      char nm[64];
      snprintf(nm, sizeof(nm), "upsample.%d", (int)(2 * i));
      SSR_TRY(pack_conv(m, nm, rs[i] * rs[i] * 64, 64, rs[i], &L));
      m->up.push_back(L);
    }
    m->last_cin = 64;
    SSR_TRY(pack_conv_last(m, "conv_last", 64, &m->conv_last_w, m->conv_last_bias));
    if (c.precision != SSR_PREC_FP32) SSR_TRY(pack_conv(m, "conv_last", 3, 64, 0, &m->last_lin));
  } else {
    // pixelshuffledirect: one conv C -> scale^2 * 3, stored un-shuffled; the shuffle happens in the finish kernel
    Lin L;
    SSR_TRY(pack_conv(m, "upsample.0", c.scale * c.scale * 3, C, 0, &L));
    m->up.push_back(L);
  }
  return SSR_OK;
}

static size_t pack_raw(ssr_model* m, const std::string& name, size_t numel);

// qkv / proj / fc1 / fc2 of one attention block in the padded-head layout of the un-fused GEMM path, plus the
// relative-position bias table transposed to [heads][nbias] (shared by SwinIR's un-fused path and HAT)
static int pack_attn_mlp(ssr_model* m, const std::string& pa, const std::string& pm, const std::string& ptab, const std::string& pn2,
                         const Layer& L, int nbias, Block* B) {
  const int C = m->C, HID = m->HID, heads = L.heads, d = L.d, DP = L.DP, QP = L.QP;
  const float qscale = 1.0f / sqrtf((float)d);
  auto qkv_row = [=](int n) {
    const int part = n / QP, hc = n % QP, h = hc / DP, j = hc % DP;
    return (h < heads && j < d) ? part * C + h * d + j : -1;
  };
  auto qkv_scale = [=](int n) { return n < QP ? qscale : 1.0f; };
  auto ident_c = [=](int k) { return k < C ? k : -1; };
  auto one = [](int) { return 1.0f; };
  SSR_TRY(pack_linear(m, pa + ".qkv", 3 * C, C, 3 * QP, m->CP, qkv_row, ident_c, qkv_scale, &B->qkv));
  B->qkv.N = 3 * QP;
  auto proj_col = [=](int k) {
    const int h = k / DP, j = k % DP;
    return (h < heads && j < d) ? h * d + j : -1;
  };
  SSR_TRY(pack_linear(m, pa + ".proj", C, C, m->CP, QP, ident_c, proj_col, one, &B->proj));
  auto ident_h = [=](int k) { return k < HID ? k : -1; };
  SSR_TRY(pack_linear(m, pm + ".fc1", HID, C, m->HP, m->CP, ident_h, ident_c, one, &B->fc1));
  if (fused_tail(m, L)) SSR_TRY(fold_norm_into_linear(m, pn2, pm + ".fc1", HID, C, m->CP, B->fc1));  // k_swin_tail.cu applies no LN2 affine
  SSR_TRY(pack_linear(m, pm + ".fc2", C, HID, m->CP, m->HP, ident_c, ident_h, one, &B->fc2));
  const std::vector<float>* T = find_param(m, ptab, (size_t)nbias * heads);
  if (!T) return SSR_E_STATE;
  B->bias_off = arena_alloc(m, (size_t)heads * nbias * 4);
  float* bt = reinterpret_cast<float*>(m->host_arena.data() + B->bias_off);
  for (int h = 0; h < heads; ++h)
    for (int i = 0; i < nbias; ++i) bt[h * nbias + i] = (*T)[(size_t)i * heads + h];
  return SSR_OK;
}

static int finalize_hat(ssr_model* m) {  // hat.py:388-470
  const ssr_model_config& c = m->cfg;
  SSR_CHECK(c.n_colors == 3, SSR_E_INVALID, "n_colors must be 3");
  const int C = c.embed_dim, ws = c.window_size;
  m->C = C;
  m->CP = round_up(C, 64);
  m->HID = (int)(C * c.mlp_ratio);
  m->HP = round_up(m->HID, 64);
  SSR_CHECK(m->CP <= 256, SSR_E_INVALID, "embed_dim %d > 256 not supported", C);
  SSR_CHECK(c.compress_ratio > 0 && C % c.compress_ratio == 0 && c.squeeze_factor > 0 && C % c.squeeze_factor == 0, SSR_E_INVALID,
            "embed_dim %d not divisible by compress_ratio %d / squeeze_factor %d", C, c.compress_ratio, c.squeeze_factor);
  const int wse = ws + (int)(c.overlap_ratio * ws);
  SSR_CHECK((ws * ws) % 64 == 0 && wse > ws && (wse - ws) % 2 == 0, SSR_E_INVALID, "window %d / overlap window %d unsupported", ws, wse);
  const int R = C / c.squeeze_factor, Cc = C / c.compress_ratio;
  SSR_TRY(pack_conv_first(m, "conv_first", C, &m->conv_first_w, &m->conv_first_b));
  SSR_TRY(pack_ln(m, "patch_embed.norm", C, m->CP, &m->pe_norm));
  m->layers.clear();
  m->QPmax = 0;
  for (int li = 0; li < c.n_layers; ++li) {
    Layer L;
    L.heads = c.num_heads[li];
    SSR_CHECK(L.heads > 0 && C % L.heads == 0, SSR_E_INVALID, "embed_dim %d not divisible by heads %d", C, L.heads);
    L.d = C / L.heads;
    SSR_CHECK(L.d <= 32, SSR_E_INVALID, "head_dim %d > 32 not supported", L.d);
    L.DP = L.d <= 16 ? 16 : 32;
    L.QP = round_up(L.heads * L.DP, 64);
    if (L.QP > m->QPmax) m->QPmax = L.QP;
    char pre[128];
    for (int bi = 0; bi < c.depths[li]; ++bi) {
      Block B;
      snprintf(pre, sizeof(pre), "layers.%d.residual_group.blocks.%d", li, bi);
      const std::string p(pre);
      SSR_TRY(pack_ln(m, p + ".norm1", C, m->CP, &B.norm1));
      SSR_TRY(pack_ln(m, p + ".norm2", C, m->CP, &B.norm2));
      SSR_TRY(pack_attn_mlp(m, p + ".attn", p + ".mlp", p + ".attn.relative_position_bias_table", p + ".norm2", L,
                            (2 * ws - 1) * (2 * ws - 1), &B));
      SSR_TRY(pack_conv(m, p + ".conv_block.cab.0", Cc, C, 0, &B.cab0));
      SSR_TRY(pack_conv(m, p + ".conv_block.cab.2", C, Cc, 0, &B.cab2));
      B.ca_w1 = pack_raw(m, p + ".conv_block.cab.3.attention.1.weight", (size_t)R * C);
      B.ca_b1 = pack_raw(m, p + ".conv_block.cab.3.attention.1.bias", (size_t)R);
      B.ca_w2 = pack_raw(m, p + ".conv_block.cab.3.attention.3.weight", (size_t)C * R);
      B.ca_b2 = pack_raw(m, p + ".conv_block.cab.3.attention.3.bias", (size_t)C);
      if (B.ca_w1 == (size_t)-1 || B.ca_b1 == (size_t)-1 || B.ca_w2 == (size_t)-1 || B.ca_b2 == (size_t)-1) return SSR_E_STATE;
      L.blocks.push_back(B);
    }
    snprintf(pre, sizeof(pre), "layers.%d.residual_group.overlap_attn", li);
    const std::string po(pre);
    SSR_TRY(pack_ln(m, po + ".norm1", C, m->CP, &L.ocab.norm1));
    SSR_TRY(pack_ln(m, po + ".norm2", C, m->CP, &L.ocab.norm2));
    SSR_TRY(pack_attn_mlp(m, po, po + ".mlp", po + ".relative_position_bias_table", po + ".norm2", L, (ws + wse - 1) * (ws + wse - 1),
                          &L.ocab));
    snprintf(pre, sizeof(pre), "layers.%d.conv", li);
    SSR_TRY(pack_conv(m, pre, C, C, 0, &L.conv));
    m->layers.push_back(L);
  }
  SSR_TRY(pack_ln(m, "norm", C, m->CP, &m->final_norm));
  SSR_TRY(pack_conv(m, "conv_after_body", C, C, 0, &m->conv_after_body));
  std::vector<int> rs;
  upsampler_plan(c.scale, &rs);
  m->up.clear();
  SSR_TRY(pack_conv(m, "conv_before_upsample.0", 64, C, 0, &m->conv_before_up));
  for (size_t i = 0; i < rs.size(); ++i) {
    Lin L;
    char nm[64];
    snprintf(nm, sizeof(nm), "upsample.%d", (int)(2 * i));
    SSR_TRY(pack_conv(m, nm, rs[i] * rs[i] * 64, 64, rs[i], &L));
    m->up.push_back(L);
  }
  m->last_cin = 64;
  SSR_TRY(pack_conv_last(m, "conv_last", 64, &m->conv_last_w, m->conv_last_bias));
  if (c.precision != SSR_PREC_FP32) SSR_TRY(pack_conv(m, "conv_last", 3, 64, 0, &m->last_lin));
  return SSR_OK;
}

static int finalize_edsr(ssr_model* m) {
  const ssr_model_config& c = m->cfg;
  SSR_CHECK(c.n_colors == 3, SSR_E_INVALID, "n_colors must be 3");
  m->F = c.n_feats;
  m->FP = round_up(m->F, 64);
  SSR_CHECK(m->FP <= 256, SSR_E_INVALID, "n_feats %d > 256 not supported", m->F);
  SSR_TRY(pack_conv_first(m, "head.0", m->F, &m->conv_first_w, &m->conv_first_b));
  const std::vector<float>* sb = find_param(m, "sub_mean.bias", 3);
  const std::vector<float>* ab = find_param(m, "add_mean.bias", 3);
  if (!sb || !ab) return SSR_E_STATE;
  for (int i = 0; i < 3; ++i) {
    m->sub_bias[i] = (*sb)[i];
    m->add_bias[i] = (*ab)[i];
  }
  m->res_a.assign(c.n_resblocks, Lin());
  m->res_b.assign(c.n_resblocks, Lin());
  for (int i = 0; i < c.n_resblocks; ++i) {
    char nm[64];
    snprintf(nm, sizeof(nm), "body.%d.body.0", i);
    SSR_TRY(pack_conv(m, nm, m->F, m->F, 0, &m->res_a[i]));
    snprintf(nm, sizeof(nm), "body.%d.body.2", i);
    SSR_TRY(pack_conv(m, nm, m->F, m->F, 0, &m->res_b[i]));
  }
  char nm[64];
  snprintf(nm, sizeof(nm), "body.%d", c.n_resblocks);
  SSR_TRY(pack_conv(m, nm, m->F, m->F, 0, &m->body_tail));
  std::vector<int> rs;
  upsampler_plan(c.scale, &rs);
  m->up.clear();
  for (size_t i = 0; i < rs.size(); ++i) {
    Lin L;
    snprintf(nm, sizeof(nm), "tail.0.%d", (int)(2 * i));
    SSR_TRY(pack_conv(m, nm, rs[i] * rs[i] * m->F, m->F, rs[i], &L));
    m->up.push_back(L);
  }
  m->last_cin = m->F;
  SSR_TRY(pack_conv_last(m, "tail.1", m->F, &m->conv_last_w, m->conv_last_bias));
  if (c.precision != SSR_PREC_FP32) SSR_TRY(pack_conv(m, "tail.1", 3, m->F, 0, &m->last_lin));
  return SSR_OK;
}

static size_t pack_raw(ssr_model* m, const std::string& name, size_t numel) {
  const std::vector<float>* v = find_param(m, name, numel);
  if (!v) return (size_t)-1;
  return pack_vec(m, *v, (int)numel);
}

static int finalize_rcan(ssr_model* m) {  // rcan.py:39-66
  const ssr_model_config& c = m->cfg;
  SSR_CHECK(c.n_colors == 3, SSR_E_INVALID, "n_colors must be 3");
  m->F = c.n_feats;
  m->FP = round_up(m->F, 64);
  SSR_CHECK(m->FP <= 256, SSR_E_INVALID, "n_feats %d > 256 not supported", m->F);
  SSR_CHECK(c.reduction > 0 && m->F % c.reduction == 0 && m->F / c.reduction <= 64, SSR_E_INVALID, "bad reduction %d", c.reduction);
  SSR_CHECK(c.n_resgroups > 0 && c.n_resblocks > 0, SSR_E_INVALID, "bad RCAN depth %dx%d", c.n_resgroups, c.n_resblocks);
  const int F = m->F, R = F / c.reduction;
  SSR_TRY(pack_conv_first(m, "head.0", F, &m->conv_first_w, &m->conv_first_b));
  const std::vector<float>* sb = find_param(m, "sub_mean.bias", 3);
  const std::vector<float>* ab = find_param(m, "add_mean.bias", 3);
  if (!sb || !ab) return SSR_E_STATE;
  for (int i = 0; i < 3; ++i) {
    m->sub_bias[i] = (*sb)[i];
    m->add_bias[i] = (*ab)[i];
  }
  const int nblk = c.n_resgroups * c.n_resblocks;
  m->res_a.assign(nblk, Lin());
  m->res_b.assign(nblk, Lin());
  m->ca.assign(nblk, ssr_model::CaP());
  m->grp_tail.assign(c.n_resgroups, Lin());
  char nm[96];
  for (int g = 0; g < c.n_resgroups; ++g) {
    for (int b = 0; b < c.n_resblocks; ++b) {
      const int i = g * c.n_resblocks + b;
      snprintf(nm, sizeof(nm), "body.%d.body.%d.body", g, b);
      const std::string p(nm);
      SSR_TRY(pack_conv(m, p + ".0", F, F, 0, &m->res_a[i]));
      SSR_TRY(pack_conv(m, p + ".2", F, F, 0, &m->res_b[i]));
      ssr_model::CaP& ca = m->ca[i];
      ca.w1 = pack_raw(m, p + ".3.conv_du.0.weight", (size_t)R * F);
      ca.b1 = pack_raw(m, p + ".3.conv_du.0.bias", (size_t)R);
      ca.w2 = pack_raw(m, p + ".3.conv_du.2.weight", (size_t)F * R);
      ca.b2 = pack_raw(m, p + ".3.conv_du.2.bias", (size_t)F);
      if (ca.w1 == (size_t)-1 || ca.b1 == (size_t)-1 || ca.w2 == (size_t)-1 || ca.b2 == (size_t)-1) return SSR_E_STATE;
    }
    snprintf(nm, sizeof(nm), "body.%d.body.%d", g, c.n_resblocks);
    SSR_TRY(pack_conv(m, nm, F, F, 0, &m->grp_tail[g]));
  }
  snprintf(nm, sizeof(nm), "body.%d", c.n_resgroups);
  SSR_TRY(pack_conv(m, nm, F, F, 0, &m->body_tail));
  std::vector<int> rs;
  upsampler_plan(c.scale, &rs);
  m->up.clear();
  for (size_t i = 0; i < rs.size(); ++i) {
    Lin L;
    snprintf(nm, sizeof(nm), "tail.0.%d", (int)(2 * i));
    SSR_TRY(pack_conv(m, nm, rs[i] * rs[i] * F, F, rs[i], &L));
    m->up.push_back(L);
  }
  m->last_cin = F;
  SSR_TRY(pack_conv_last(m, "tail.1", F, &m->conv_last_w, m->conv_last_bias));
  if (c.precision != SSR_PREC_FP32) SSR_TRY(pack_conv(m, "tail.1", 3, F, 0, &m->last_lin));
  return SSR_OK;
}

static int finalize_han(ssr_model* m) {  // han.py:55-88 = the RCAN trunk (finalize_rcan) + csa, la, last_conv, last
  const ssr_model_config& c = m->cfg;
  SSR_CHECK(c.n_resgroups == 10, SSR_E_INVALID, "HAN: last_conv takes n_feats * 11 channels (han.py:87), i.e. n_resgroups must be 10 (got %d)",
            c.n_resgroups);
  SSR_CHECK(c.n_feats % 64 == 0, SSR_E_INVALID, "HAN: n_feats %d is not a multiple of 64", c.n_feats);
  SSR_TRY(finalize_rcan(m));
  const int F = m->F;
  SSR_TRY(pack_conv(m, "last_conv", F, 11 * F, 0, &m->han_last_conv));
  SSR_TRY(pack_conv(m, "last", F, 2 * F, 0, &m->han_last));
  m->csa_w = pack_raw(m, "csa.conv.weight", 27);
  m->csa_b = pack_raw(m, "csa.conv.bias", 1);
  m->csa_gamma = pack_raw(m, "csa.gamma", 1);
  m->la_gamma = pack_raw(m, "la.gamma", 1);
  if (m->csa_w == (size_t)-1 || m->csa_b == (size_t)-1 || m->csa_gamma == (size_t)-1 || m->la_gamma == (size_t)-1) return SSR_E_STATE;
  return SSR_OK;
}

// ---------------------------------------------------------------------------------------------
// workspace planning
static void padded_size(const ssr_model* m, int H, int W, int pad_mode, int* Hp, int* Wp) {
  if (!is_swin(m) && m->cfg.arch != SSR_ARCH_HAT) {
    *Hp = H;
    *Wp = W;
    return;
  }
  const int ws = m->cfg.window_size;
  if (pad_mode == SSR_PAD_EVAL) {
    *Hp = (H / ws + 1) * ws;
    *Wp = (W / ws + 1) * ws;
  } else {
    *Hp = (H + ws - 1) / ws * ws;
    *Wp = (W + ws - 1) / ws * ws;
  }
}

struct SwinWs {
  float *x0, *g, *t;
  void *xn, *qkv, *o, *hbuf, *tb, *cbu, *hr[2];
  size_t hr_elems[2];
  // SwinFIR (fp32-class precisions only, so every buffer is fp32): SpatialB's hidden map, [S(x) | F(x)], the SpectralTransform's
  // half-width maps before / after the Fourier unit, three half-spectrum arrays [B*Hp*(Wp/2+1)][CP]
  float *sfb_s1, *sfb_cat, *sfb_y, *sfb_z, *sfb_fa, *sfb_fb, *sfb_fc;
};

static size_t plan_swinir(const ssr_model* m, void* base, int B, int Hp, int Wp, SwinWs* w) {
  Carver c(base);
  const size_t T = (size_t)B * Hp * Wp, e = m->elem;
  w->x0 = (float*)c.take(T * m->CP * 4);
  w->g = (float*)c.take(T * m->CP * 4);
  w->t = (float*)c.take(T * m->CP * 4);
  w->xn = c.take(T * m->CP * e);
  w->qkv = c.take(T * 3 * m->QPmax * e);
  w->o = c.take(T * m->QPmax * e);
  w->hbuf = c.take(T * m->HP * e);
  w->tb = c.take(T * m->CP * e);
  w->cbu = c.take(T * 64 * e);
  if (m->sfb) {
    const size_t c2P = (size_t)round_up(m->C / 2, 64), Mf = (size_t)B * Hp * (Wp / 2 + 1);
    w->sfb_s1 = (float*)c.take(T * m->CP * 4);
    w->sfb_cat = (float*)c.take(T * 2 * m->CP * 4);
    w->sfb_y = (float*)c.take(T * c2P * 4);
    w->sfb_z = (float*)c.take(T * c2P * 4);
    w->sfb_fa = (float*)c.take(Mf * m->CP * 4);
    w->sfb_fb = (float*)c.take(Mf * m->CP * 4);
    w->sfb_fc = (float*)c.take(Mf * m->CP * 4);
  }
  // high-resolution ping-pong buffers of the pixel-shuffle tail (64 channels)
  size_t need[2] = {0, 0};
  size_t px = T;
  for (size_t i = 0; i < m->up.size() && m->cfg.upsampler == 0; ++i) {
    px *= (size_t)m->up[i].ps_r * m->up[i].ps_r;
    if (px * 64 > need[i & 1]) need[i & 1] = px * 64;
  }
  if (m->cfg.upsampler == 1) need[0] = T * m->up[0].NP;  // fp32 rows of the pixelshuffledirect conv
  for (int i = 0; i < 2; ++i) {
    w->hr_elems[i] = need[i];
    w->hr[i] = need[i] ? c.take(need[i] * (m->cfg.upsampler == 1 ? 4 : e)) : nullptr;
  }
  return c.off + 1024;
}

struct EdsrWs {
  float *x, *r;
  void *rb, *tmp, *hr[2];
};
static size_t plan_edsr(const ssr_model* m, void* base, int B, int H, int W, EdsrWs* w) {
  Carver c(base);
  const size_t T = (size_t)B * H * W, e = m->elem;
  w->x = (float*)c.take(T * m->FP * 4);
  w->r = (float*)c.take(T * m->FP * 4);
  w->rb = c.take(T * m->FP * e);
  w->tmp = c.take(T * m->FP * e);
  size_t need[2] = {0, 0};
  size_t px = T;
  for (size_t i = 0; i < m->up.size(); ++i) {
    px *= (size_t)m->up[i].ps_r * m->up[i].ps_r;
    if (px * m->FP > need[i & 1]) need[i & 1] = px * m->FP;
  }
  for (int i = 0; i < 2; ++i) w->hr[i] = need[i] ? c.take(need[i] * e) : nullptr;
  return c.off + 1024;
}

// ---------------------------------------------------------------------------------------------
// launch helpers
int run_gemm(const ssr_model* m, GemmArgs& g, cudaStream_t s) {
  g.round_tf32 = m->cfg.precision == SSR_PREC_TF32;
  g.split_tf32 = m->cfg.precision == SSR_PREC_TF32X3;
  if (m->cfg.precision == SSR_PREC_FP32) return launch_gemm_simt(g, s);
  return launch_gemm_tc(g, m->elem, s);
}

GemmArgs gemm_base(const ssr_model* m, const Lin& L, const void* A, int lda, int B, int H, int W) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A;
  g.lda = lda;
  g.B = B;
  g.H = H;
  g.W = W;
  g.M = B * H * W;
  g.taps = L.taps;
  g.KP = L.KP;
  g.Wt = m->arena + L.w_off;
  g.N = L.N;
  g.NP = L.NP;
  g.bias = m->dev<float>(L.b_off);
  g.act = ACT_NONE;
  g.slope = 0.01f;
  g.alpha = 1.0f;
  g.eps = 1e-5f;
  g.ps_r = L.ps_r;
  g.K_alg = L.K;
  g.N_alg = L.N_alg;
  return g;
}

static void set_ln(const ssr_model* m, GemmArgs& g, const LNp& ln, void* out, int ld) {
  g.out_ln = out;
  g.ld_ln = ld;
  g.gamma = m->dev<float>(ln.g_off);
  g.beta = m->dev<float>(ln.b_off);
}

// pixel-shuffle tail shared by SwinIR ("pixelshuffle") and EDSR: cur [B,H,W,ch] -> conv_last
static int run_tail(ssr_model* m, const void* cur, int ch_ld, int B, int Hp, int Wp, void* hr0, void* hr1, int h, int w,
                    const float* out_shift, float out_scale, const OutputSpec& out, cudaStream_t s) {
  int H = Hp, W = Wp;
  void* bufs[2] = {hr0, hr1};
  for (size_t i = 0; i < m->up.size(); ++i) {
    const Lin& L = m->up[i];
    GemmArgs g = gemm_base(m, L, cur, ch_ld, B, H, W);
    g.out_T = bufs[i & 1];
    g.ld_T = ch_ld;
    SSR_TRY(run_gemm(m, g, s));
    cur = bufs[i & 1];
    H *= L.ps_r;
    W *= L.ps_r;
  }
  if (m->last_w27 && ch_ld >= 64) {  // bf16, 64 input channels: one K = 64 GEMM with the nine taps in N + a shared-memory gather
    return launch_conv_last_tapn(cur, ch_ld, m->arena + m->last_w27, m->conv_last_bias, out_shift, out_scale,
                                 m->cfg.img_range == 1.0f ? 255.0f : 1.0f, B, H, W, h * m->cfg.scale, w * m->cfg.scale, out.out_f32,
                                 out.out_u8, s);
  }
  if (m->cfg.precision != SSR_PREC_FP32) {  // tensor-core implicit GEMM (N padded 3 -> 64) with the reconstruction epilogue
    GemmArgs g = gemm_base(m, m->last_lin, cur, ch_ld, B, H, W);
    g.out3_f32 = out.out_f32;
    g.out3_u8 = out.out_u8;
    g.crop_h = h * m->cfg.scale;
    g.crop_w = w * m->cfg.scale;
    for (int i = 0; i < 3; ++i) g.out_shift[i] = out_shift[i];
    g.out_scale = out_scale;
    g.u8_scale = m->cfg.img_range == 1.0f ? 255.0f : 1.0f;
    return run_gemm(m, g, s);
  }
  ConvLastArgs a;
  memset(&a, 0, sizeof(a));
  a.in = cur;
  a.ldi = ch_ld;
  a.Cin = m->last_cin;
  a.elem = m->elem;
  a.B = B;
  a.Hs = H;
  a.Ws = W;
  a.ch = h * m->cfg.scale;
  a.cw = w * m->cfg.scale;
  a.Wc = m->dev<float>(m->conv_last_w);
  for (int i = 0; i < 3; ++i) {
    a.bias[i] = m->conv_last_bias[i];
    a.out_shift[i] = out_shift[i];
  }
  a.out_scale = out_scale;
  a.out_f32 = out.out_f32;
  a.out_u8 = out.out_u8;
  a.u8_scale = m->cfg.img_range == 1.0f ? 255.0f : 1.0f;
  return launch_conv_last(a, s);
}

// developer diagnostics (STUDIOSR_B200_DEBUG_NAN=1): count non-finite values of an activation buffer after a launch
__global__ void count_nonfinite_kernel(const void* p, size_t n, int elem, unsigned long long* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  unsigned long long c = 0;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = elem == 2 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
    if (!isfinite(v)) ++c;
  }
  if (c) atomicAdd(out, c);
}
static void debug_nonfinite(const char* what, int li, int bi, const void* p, size_t n, int elem, cudaStream_t s) {
  static const bool on = getenv("STUDIOSR_B200_DEBUG_NAN") != nullptr;
  if (!on) return;
  unsigned long long* d;
  cudaMalloc(&d, 8);
  cudaMemsetAsync(d, 0, 8, s);
  count_nonfinite_kernel<<<256, 256, 0, s>>>(p, n, elem, d);
  unsigned long long h = 0;
  cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  fprintf(stderr, "[debug] layer %d block %d %-10s non-finite %llu of %zu (%s)\n", li, bi, what, h, n, cudaGetErrorString(e));
  cudaFree(d);
  if (h && elem == 2 && li == 0 && bi == 0) {  // pattern of the first bad buffer: per 32-column group, and the first bad rows
    std::vector<__nv_bfloat16> hb(n);
    cudaMemcpy(hb.data(), p, n * 2, cudaMemcpyDeviceToHost);
    size_t rows = n / 192, colg[6] = {0, 0, 0, 0, 0, 0};
    int printed = 0;
    for (size_t r = 0; r < rows; ++r) {
      int bad = 0;
      for (int c = 0; c < 192; ++c)
        if (!isfinite(__bfloat162float(hb[r * 192 + c]))) { ++bad; ++colg[c / 32]; }
      if (bad && printed < 24) { fprintf(stderr, "   row %zu (y %zu x %zu): %d bad\n", r, r / 72, r % 72, bad); ++printed; }
    }
    fprintf(stderr, "   bad per head: %zu %zu %zu %zu %zu %zu\n", colg[0], colg[1], colg[2], colg[3], colg[4], colg[5]);
  }
}

// SFB(x) = fusion(cat[S(x), F(x)]) (swinfir.py:68-80) on pixel-major fp32 maps.  xa = x as the GEMM operand (tf32-rounded in
// the tf32 mode), xr = x for the SpatialB skip.  `epi` puts the caller's epilogue (skip, fp32 / LayerNorm outputs) on the
// fusion GEMM, which takes the place of the plain conv's.
template <typename Epi>
static int run_sfb(ssr_model* m, const Sfb& S, const void* xa, const float* xr, int B, int Hp, int Wp, const SwinWs& W, Epi epi,
                   cudaStream_t s) {
  const int CP = m->CP, c2 = m->C / 2, c2P = round_up(c2, 64), Wf = Wp / 2 + 1;
  const int rtf = m->cfg.precision == SSR_PREC_TF32;
  {  // S: conv - LeakyReLU(0.2) - conv + x (swinfir.py:53-65) -> columns [0, CP) of the concatenation
    GemmArgs a = gemm_base(m, S.s0, xa, CP, B, Hp, Wp);
    a.act = ACT_LEAKY;
    a.slope = 0.2f;
    a.out_T = W.sfb_s1;
    a.ld_T = CP;
    SSR_TRY(run_gemm(m, a, s));
    GemmArgs b = gemm_base(m, S.s2, W.sfb_s1, CP, B, Hp, Wp);
    b.res = xr;
    b.ldres = CP;
    b.out_T = W.sfb_cat;
    b.ld_T = 2 * CP;
    SSR_TRY(run_gemm(m, b, s));
  }
  {  // F: y = LeakyReLU(conv1x1(x)) (swinfir.py:40-43,47)
    GemmArgs a = gemm_base(m, S.before, xa, CP, B, Hp, Wp);
    a.act = ACT_LEAKY;
    a.slope = 0.2f;
    a.out_f32 = W.sfb_y;  // un-rounded copy: input of the transform and of the `output + x` skip
    a.ld_f32 = c2P;
    SSR_TRY(run_gemm(m, a, s));
  }
  // FourierUnit (swinfir.py:18-34): rfft2 -> 1x1 conv + LeakyReLU on [real | imag] -> irfft2, then + y (swinfir.py:49)
  SSR_TRY(launch_rfft2(W.sfb_y, c2P, W.sfb_fa, W.sfb_fb, CP, B, Hp, Wp, c2, rtf, s));
  {
    GemmArgs a = gemm_base(m, S.fu, W.sfb_fb, CP, B, Hp, Wf);
    a.act = ACT_LEAKY;
    a.slope = 0.2f;
    a.out_f32 = W.sfb_fc;
    a.ld_f32 = CP;
    SSR_TRY(run_gemm(m, a, s));
  }
  SSR_TRY(launch_irfft2_add(W.sfb_fc, CP, W.sfb_fa, W.sfb_y, c2P, W.sfb_z, c2P, B, Hp, Wp, c2, 4, rtf, s));
  {  // conv_after_fft (swinfir.py:49) -> columns [CP, 2 CP)
    GemmArgs a = gemm_base(m, S.after, W.sfb_z, c2P, B, Hp, Wp);
    a.out_T = W.sfb_cat + CP;
    a.ld_T = 2 * CP;
    SSR_TRY(run_gemm(m, a, s));
  }
  GemmArgs f = gemm_base(m, S.fusion, W.sfb_cat, 2 * CP, B, Hp, Wp);
  epi(f);
  return run_gemm(m, f, s);
}

static int forward_swinir(ssr_model* m, const InputSpec& in, const OutputSpec& out, int B, int h, int w, int pad_mode,
                          void* ws, size_t ws_bytes, cudaStream_t s) {
  const ssr_model_config& c = m->cfg;
  int Hp, Wp;
  padded_size(m, h, w, pad_mode, &Hp, &Wp);
  if (pad_mode == SSR_PAD_TRAIN)
    SSR_CHECK(Hp - h < h && Wp - w < w, SSR_E_INVALID, "reflect padding needs pad < size (%dx%d)", h, w);
  SwinWs W;
  const size_t need = plan_swinir(m, ws, B, Hp, Wp, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, need);
  const int CP = m->CP, e = m->elem;
  const int T = B * Hp * Wp;
  const int rtf = c.precision == SSR_PREC_TF32;
  if (m->sfb) {  // the FFT kernels write channels [0, 2 c2) / [0, c2) only: the padded K columns of the GEMMs that follow must be 0
    SSR_CUDA(cudaMemsetAsync(W.sfb_fb, 0, (size_t)B * Hp * (Wp / 2 + 1) * CP * 4, s));
    SSR_CUDA(cudaMemsetAsync(W.sfb_z, 0, (size_t)T * round_up(m->C / 2, 64) * 4, s));
  }

  {  // pad + normalise + conv_first (swinir.py:356-361)
    ConvFirstArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in.in;
    a.in_u8 = in.in_u8;
    a.fh = in.fh;
    a.fw = in.fw;
    a.tile_mode = in.tile_mode;
    a.tile = in.tile;
    a.stride = in.stride;
    a.tiles_x = in.tiles_x;
    a.tile_begin = in.tile_begin;
    a.h = h;
    a.w = w;
    a.Hp = Hp;
    a.Wp = Wp;
    a.pad_mode = pad_mode;
    a.B = B;
    const float u8s = (in.in_u8 && c.img_range == 1.0f) ? 1.0f / 255.0f : 1.0f;
    a.in_scale = u8s / c.img_range;
    for (int i = 0; i < 3; ++i) a.in_shift[i] = -kRgbMean[i];
    a.Wc = m->dev<float>(m->conv_first_w);
    a.bias = m->dev<float>(m->conv_first_b);
    a.Cout = m->C;
    a.out_f32 = W.x0;
    a.ld_f32 = CP;
    a.elem = e;
    if (CP <= 192) {  // + patch_embed.norm -> g (residual stream) and layers.0.blocks.0.norm1 -> xn in the same pass
      const Block& b0 = m->layers[0].blocks[0];
      a.g1 = m->dev<float>(m->pe_norm.g_off);
      a.b1 = m->dev<float>(m->pe_norm.b_off);
      a.g2 = m->dev<float>(b0.norm1.g_off);
      a.b2 = m->dev<float>(b0.norm1.b_off);
      a.out_g = W.g;
      a.ld_g = CP;
      a.out_T = W.xn;
      a.ld_T = CP;
      a.round_tf32 = rtf;
      a.eps = 1e-5f;
      SSR_TRY(launch_conv_first_ln(a, CP, s));
    } else {
      SSR_TRY(launch_conv_first(a, s));
    }
  }
  if (CP > 192) {  // patch_embed.norm -> g (residual stream), chained with layers.0.blocks.0.norm1 -> xn
    LnArgs a;
    memset(&a, 0, sizeof(a));
    a.in = W.x0;
    a.ld_in = CP;
    a.M = T;
    a.C = m->C;
    a.CP = CP;
    a.g1 = m->dev<float>(m->pe_norm.g_off);
    a.b1 = m->dev<float>(m->pe_norm.b_off);
    a.out_f32 = W.g;
    a.ld_f32 = CP;
    const Block& b0 = m->layers[0].blocks[0];
    a.g2 = m->dev<float>(b0.norm1.g_off);
    a.b2 = m->dev<float>(b0.norm1.b_off);
    a.out_T = W.xn;
    a.ld_T = CP;
    a.elem = e;
    a.round_tf32 = rtf;
    a.eps = 1e-5f;
    SSR_TRY(launch_layernorm(a, s));
  }
  SSR_CUDA(cudaMemsetAsync(W.o, 0, (size_t)T * m->QPmax * e, s));
  const int nL = (int)m->layers.size();
  for (int li = 0; li < nL; ++li) {
    const Layer& L = m->layers[li];
    const int depth = (int)L.blocks.size();
    const bool fused_mlp = fused_tail(m, L);
    for (int bi = 0; bi < depth; ++bi) {
      const Block& blk = L.blocks[bi];
      if (fused_attn(m, L)) {  // qkv projection + roll + partition + attention + reverse + roll in ONE kernel
        AttnFusedArgs f;
        memset(&f, 0, sizeof(f));
        f.xn = W.xn; f.ld_x = CP; f.o = W.o; f.ld_o = L.QP;
        f.Whp = m->arena + blk.whp_off; f.bias_tab = m->arena + blk.btab_off;
        f.B = B; f.H = Hp; f.W = Wp; f.shift = (bi % 2 == 0) ? 0 : c.window_size / 2;
        f.C = m->C; f.d = L.d;
        debug_nonfinite("xn", li, bi, W.xn, (size_t)T * CP, e, s);
        if (li == 0 && bi == 0) {
          debug_nonfinite("Whp", li, bi + 100, f.Whp, kAttnWhpBytes / 2, 2, s);
          debug_nonfinite("btab", li, bi + 100, f.bias_tab, kAttnBiasBytes / 2, 2, s);
        }
        SSR_TRY(launch_swin_attn_fused(f, s));
        debug_nonfinite("o", li, bi, W.o, (size_t)T * L.QP, e, s);
      } else {
      {  // qkv projection (swinir.py:80); q scale folded into the packed weights
        GemmArgs g = gemm_base(m, blk.qkv, W.xn, CP, B, Hp, Wp);
        g.out_T = W.qkv;
        g.ld_T = 3 * L.QP;
        SSR_TRY(run_gemm(m, g, s));
      }
      {  // roll + partition + attention + reverse + roll (swinir.py:154-168, 83-102)
        AttnArgs a;
        memset(&a, 0, sizeof(a));
        a.qkv = W.qkv;
        a.ld_qkv = 3 * L.QP;
        a.QP = L.QP;
        a.o = W.o;
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
        if (e == 2)
          SSR_TRY(launch_attn_mma(a, s));
        else
          SSR_TRY(launch_attn_simt(a, s));
      }
      }
      if (fused_mlp) {  // proj + res + LN2 + fc1 + GELU + fc2 + res (+ next norm1 | bf16 copy) in ONE kernel
        MlpFusedArgs f;
        memset(&f, 0, sizeof(f));
        f.o = W.o; f.ld_o = L.QP; f.M = T; f.C = m->C; f.Hid = m->HID; f.CP = CP; f.HP = m->HP; f.QP = L.QP;
        f.Wp = m->arena + blk.proj.w_off; f.W1 = m->arena + blk.fc1.w_off; f.W2 = m->arena + blk.fc2.w_off;
        f.bp = m->dev<float>(blk.proj.b_off); f.b2 = m->dev<float>(blk.fc2.b_off);
        f.res = bi == 0 ? W.g : W.t; f.ldres = CP;
        f.eps = 1e-5f;
        if (bi + 1 < depth) {
          f.out_f32 = W.t; f.ld_f32 = CP;
          f.out_ln = W.xn; f.ld_ln = CP;
          f.g3 = m->dev<float>(L.blocks[bi + 1].norm1.g_off); f.be3 = m->dev<float>(L.blocks[bi + 1].norm1.b_off);
        } else {
          f.out_T = W.tb; f.ld_T = CP;
        }
        SSR_TRY(launch_mlp_fused(f, s));
        debug_nonfinite("t", li, bi, W.t, (size_t)T * CP, 4, s);
        continue;
      }
      {  // proj + residual (swinir.py:103,171) with norm2 fused into the epilogue
        GemmArgs g = gemm_base(m, blk.proj, W.o, L.QP, B, Hp, Wp);
        g.res = bi == 0 ? W.g : W.t;
        g.ldres = CP;
        g.out_f32 = W.t;
        g.ld_f32 = CP;
        set_ln(m, g, blk.norm2, W.xn, CP);
        SSR_TRY(run_gemm(m, g, s));
      }
      {  // fc1 + GELU (common.py:185-186)
        GemmArgs g = gemm_base(m, blk.fc1, W.xn, CP, B, Hp, Wp);
        g.act = ACT_GELU;
        g.out_T = W.hbuf;
        g.ld_T = m->HP;
        SSR_TRY(run_gemm(m, g, s));
      }
      {  // fc2 + residual (swinir.py:172); epilogue = next block's norm1, or a T copy for the RSTB conv
        GemmArgs g = gemm_base(m, blk.fc2, W.hbuf, m->HP, B, Hp, Wp);
        g.res = W.t;
        g.ldres = CP;
        if (bi + 1 < depth) {
          g.out_f32 = W.t;
          g.ld_f32 = CP;
          set_ln(m, g, L.blocks[bi + 1].norm1, W.xn, CP);
        } else {
          g.out_T = W.tb;
          g.ld_T = CP;
          if (m->sfb) {  // SwinFIR: the SFB's SpatialB skip reads the un-rounded fp32 stream
            g.out_f32 = W.t;
            g.ld_f32 = CP;
          }
        }
        SSR_TRY(run_gemm(m, g, s));
      }
    }
    {  // RSTB conv + group residual (swinir.py:245-246); epilogue = next layer's first norm1 or the final norm
      auto epilogue = [&](GemmArgs& g) {
        g.res = W.g;
        g.ldres = CP;
        if (li + 1 < nL) {
          g.out_f32 = W.g;
          g.ld_f32 = CP;
          set_ln(m, g, m->layers[li + 1].blocks[0].norm1, W.xn, CP);
        } else {
          set_ln(m, g, m->final_norm, W.xn, CP);
        }
      };
      if (m->sfb) {  // SwinFIR: resi_connection = SFB (swinfir.py:112)
        SSR_TRY(run_sfb(m, L.sfb, W.tb, W.t, B, Hp, Wp, W, epilogue, s));
      } else {
        GemmArgs g = gemm_base(m, L.conv, W.tb, CP, B, Hp, Wp);
        epilogue(g);
        SSR_TRY(run_gemm(m, g, s));
      }
    }
  }
  {  // conv_after_body + long skip (swinir.py:362)
    auto epilogue = [&](GemmArgs& g) {
      g.res = W.x0;
      g.ldres = CP;
      g.out_T = W.tb;
      g.ld_T = CP;
    };
    if (m->sfb) {  // SwinFIR: conv_after_body = SFB (swinfir.py:114); its input is the final norm's (T-typed = fp32) output
      SSR_TRY(run_sfb(m, m->sfb_after_body, W.xn, reinterpret_cast<const float*>(W.xn), B, Hp, Wp, W, epilogue, s));
    } else {
      GemmArgs g = gemm_base(m, m->conv_after_body, W.xn, CP, B, Hp, Wp);
      epilogue(g);
      SSR_TRY(run_gemm(m, g, s));
    }
  }
  float shift[3] = {kRgbMean[0], kRgbMean[1], kRgbMean[2]};
  if (c.upsampler == 0) {
    GemmArgs g = gemm_base(m, m->conv_before_up, W.tb, CP, B, Hp, Wp);
    g.act = ACT_LEAKY;
    g.slope = 0.01f;
    g.out_T = W.cbu;
    g.ld_T = 64;
    SSR_TRY(run_gemm(m, g, s));
    return run_tail(m, W.cbu, 64, B, Hp, Wp, W.hr[0], W.hr[1], h, w, shift, c.img_range, out, s);
  }
  // lightweight SR ("pixelshuffledirect", swinir.py:327-329,367-369): one conv C -> 3 s^2, then shuffle + un-normalise + crop
  GemmArgs g = gemm_base(m, m->up[0], W.tb, CP, B, Hp, Wp);
  g.out_f32 = reinterpret_cast<float*>(W.hr[0]);
  g.ld_f32 = m->up[0].NP;
  SSR_TRY(run_gemm(m, g, s));
  return launch_shuffle_finish(reinterpret_cast<float*>(W.hr[0]), m->up[0].NP, B, Hp, Wp, c.scale, h * c.scale, w * c.scale, shift,
                               c.img_range, c.img_range == 1.0f ? 255.0f : 1.0f, out.out_f32, out.out_u8, s);
}

static int forward_edsr(ssr_model* m, const InputSpec& in, const OutputSpec& out, int B, int h, int w, void* ws,
                        size_t ws_bytes, cudaStream_t s) {
  const ssr_model_config& c = m->cfg;
  EdsrWs W;
  const size_t need = plan_edsr(m, ws, B, h, w, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, need);
  const int FP = m->FP, e = m->elem;
  {  // sub_mean + head (edsr.py:40-41)
    ConvFirstArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in.in;
    a.in_u8 = in.in_u8;
    a.fh = in.fh;
    a.fw = in.fw;
    a.tile_mode = in.tile_mode;
    a.tile = in.tile;
    a.stride = in.stride;
    a.tiles_x = in.tiles_x;
    a.tile_begin = in.tile_begin;
    a.h = h;
    a.w = w;
    a.Hp = h;
    a.Wp = w;
    a.pad_mode = 2;
    a.B = B;
    a.in_scale = (in.in_u8 && c.img_range == 1.0f) ? 1.0f / 255.0f : 1.0f;
    for (int i = 0; i < 3; ++i) a.in_shift[i] = m->sub_bias[i];
    a.Wc = m->dev<float>(m->conv_first_w);
    a.bias = m->dev<float>(m->conv_first_b);
    a.Cout = m->F;
    a.out_f32 = W.x;
    a.ld_f32 = FP;
    a.out_T = W.rb;
    a.ld_T = FP;
    a.elem = e;
    a.round_tf32 = c.precision == SSR_PREC_TF32;
    SSR_TRY(launch_conv_first(a, s));
  }
  for (int i = 0; i < c.n_resblocks; ++i) {  // ResBlock (common.py:150-153)
    GemmArgs ga = gemm_base(m, m->res_a[i], W.rb, FP, B, h, w);
    ga.act = ACT_RELU;
    ga.out_T = W.tmp;
    ga.ld_T = FP;
    SSR_TRY(run_gemm(m, ga, s));
    GemmArgs gb = gemm_base(m, m->res_b[i], W.tmp, FP, B, h, w);
    gb.alpha = c.res_scale;
    gb.res = i == 0 ? W.x : W.r;
    gb.ldres = FP;
    gb.out_f32 = W.r;
    gb.ld_f32 = FP;
    gb.out_T = W.rb;
    gb.ld_T = FP;
    SSR_TRY(run_gemm(m, gb, s));
  }
  {  // body tail conv + long skip (edsr.py:43-44)
    GemmArgs g = gemm_base(m, m->body_tail, W.rb, FP, B, h, w);
    g.res = W.x;
    g.ldres = FP;
    g.out_T = W.tmp;
    g.ld_T = FP;
    SSR_TRY(run_gemm(m, g, s));
  }
  return run_tail(m, W.tmp, FP, B, h, w, W.hr[0], W.hr[1], h, w, m->add_bias, 1.0f, out, s);
}

struct HatWs {
  float *x0, *g, *t, *t2, *partial;
  void *xn, *qkv, *o, *hbuf, *tb, *cbu, *c1, *hr[2];
  int nsplit;
};
static size_t plan_hat(const ssr_model* m, void* base, int B, int Hp, int Wp, HatWs* w) {
  Carver c(base);
  const size_t T = (size_t)B * Hp * Wp, e = m->elem;
  const int HW = Hp * Wp;
  w->nsplit = HW >= 16384 ? 64 : (HW + 255) / 256;
  w->x0 = (float*)c.take(T * m->CP * 4);
  w->g = (float*)c.take(T * m->CP * 4);
  w->t = (float*)c.take(T * m->CP * 4);
  w->t2 = (float*)c.take(T * m->CP * 4);
  w->partial = (float*)c.take((size_t)B * w->nsplit * m->C * 4);
  w->xn = c.take(T * m->CP * e);
  w->qkv = c.take(T * 3 * m->QPmax * e);
  w->o = c.take(T * m->QPmax * e);
  w->hbuf = c.take(T * m->HP * e);
  w->tb = c.take(T * m->CP * e);
  w->cbu = c.take(T * 64 * e);
  w->c1 = c.take(T * round_up(m->C / m->cfg.compress_ratio, 64) * e);
  size_t need[2] = {0, 0};
  size_t px = T;
  for (size_t i = 0; i < m->up.size(); ++i) {
    px *= (size_t)m->up[i].ps_r * m->up[i].ps_r;
    if (px * 64 > need[i & 1]) need[i & 1] = px * 64;
  }
  for (int i = 0; i < 2; ++i) w->hr[i] = need[i] ? c.take(need[i] * e) : nullptr;
  return c.off + 1024;
}

// HAT forward (hat.py:542-554): the un-fused GEMM path of SwinIR with 16x16 windows, plus the channel-attention
// block inside every HAB and the overlapping cross-attention block closing every group.
static int forward_hat(ssr_model* m, const InputSpec& in, const OutputSpec& out, int B, int h, int w, void* ws, size_t ws_bytes,
                       cudaStream_t s) {
  const ssr_model_config& c = m->cfg;
  int Hp, Wp;
  padded_size(m, h, w, SSR_PAD_TRAIN, &Hp, &Wp);  // hat.py:544: reflect pad in both modes
  SSR_CHECK(Hp - h < h && Wp - w < w, SSR_E_INVALID, "reflect padding needs pad < size (%dx%d)", h, w);
  HatWs W;
  const size_t need = plan_hat(m, ws, B, Hp, Wp, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, need);
  const int CP = m->CP, e = m->elem;
  const int T = B * Hp * Wp;
  const int rtf = c.precision == SSR_PREC_TF32;
  const int wse = c.window_size + (int)(c.overlap_ratio * c.window_size);
  const int R = m->C / c.squeeze_factor, CcP = round_up(m->C / c.compress_ratio, 64);

  {  // reflect pad + normalise + conv_first (hat.py:544-548)
    ConvFirstArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in.in; a.in_u8 = in.in_u8; a.fh = in.fh; a.fw = in.fw;
    a.tile_mode = in.tile_mode; a.tile = in.tile; a.stride = in.stride; a.tiles_x = in.tiles_x; a.tile_begin = in.tile_begin;
    a.h = h; a.w = w; a.Hp = Hp; a.Wp = Wp; a.pad_mode = SSR_PAD_TRAIN; a.B = B;
    const float u8s = (in.in_u8 && c.img_range == 1.0f) ? 1.0f / 255.0f : 1.0f;
    a.in_scale = u8s / c.img_range;
    for (int i = 0; i < 3; ++i) a.in_shift[i] = -kRgbMean[i];
    a.Wc = m->dev<float>(m->conv_first_w);
    a.bias = m->dev<float>(m->conv_first_b);
    a.Cout = m->C;
    a.out_f32 = W.x0; a.ld_f32 = CP; a.elem = e;
    SSR_TRY(launch_conv_first(a, s));
  }
  auto layer_norm = [&](const float* src, const LNp& ln, float* dst_f32, const LNp* ln2, void* dst_T) {
    LnArgs a;
    memset(&a, 0, sizeof(a));
    a.in = src; a.ld_in = CP; a.M = T; a.C = m->C; a.CP = CP;
    a.g1 = m->dev<float>(ln.g_off); a.b1 = m->dev<float>(ln.b_off);
    a.out_f32 = dst_f32; a.ld_f32 = CP;
    if (ln2) {
      a.g2 = m->dev<float>(ln2->g_off);
      a.b2 = m->dev<float>(ln2->b_off);
    }
    a.out_T = dst_T; a.ld_T = CP; a.elem = e; a.round_tf32 = rtf; a.eps = 1e-5f;
    return launch_layernorm(a, s);
  };
  // patch_embed.norm -> g (residual stream), chained with layers.0.blocks.0.norm1 -> xn
  SSR_TRY(layer_norm(W.x0, m->pe_norm, W.g, &m->layers[0].blocks[0].norm1, W.xn));
  SSR_CUDA(cudaMemsetAsync(W.o, 0, (size_t)T * m->QPmax * e, s));
  SSR_CUDA(cudaMemsetAsync(W.c1, 0, (size_t)T * CcP * e, s));
  const int nL = (int)m->layers.size();

  // shared tail of HAB / OCAB: norm2 -> fc1 + GELU -> fc2 + residual, epilogue = `next` LayerNorm into xn or a T copy into tb
  auto mlp = [&](const Block& blk, bool ln2_fused, const LNp* next) -> int {
    if (!ln2_fused) SSR_TRY(layer_norm(W.t, blk.norm2, nullptr, nullptr, W.xn));
    {
      GemmArgs g = gemm_base(m, blk.fc1, W.xn, CP, B, Hp, Wp);
      g.act = ACT_GELU;
      g.out_T = W.hbuf;
      g.ld_T = m->HP;
      SSR_TRY(run_gemm(m, g, s));
    }
    GemmArgs g = gemm_base(m, blk.fc2, W.hbuf, m->HP, B, Hp, Wp);
    g.res = W.t;
    g.ldres = CP;
    if (next) {
      g.out_f32 = W.t;
      g.ld_f32 = CP;
      set_ln(m, g, *next, W.xn, CP);
    } else {
      g.out_T = W.tb;
      g.ld_T = CP;
    }
    return run_gemm(m, g, s);
  };

  for (int li = 0; li < nL; ++li) {
    const Layer& L = m->layers[li];
    const int depth = (int)L.blocks.size();
    const bool fused = fused_tail(m, L);  // same predicate finalize_hat used to fold norm2 into fc1
    for (int bi = 0; bi < depth; ++bi) {  // HAB (hat.py:154-195); xn = norm1(x) on entry
      const Block& blk = L.blocks[bi];
      const float* shortcut = bi == 0 ? W.g : W.t;
      {  // channel attention block on norm1(x): conv 3x3 -> GELU -> conv 3x3 (hat.py:44-49); the gate is applied below
        GemmArgs g = gemm_base(m, blk.cab0, W.xn, CP, B, Hp, Wp);
        g.act = ACT_GELU;
        g.out_T = W.c1;
        g.ld_T = CcP;
        SSR_TRY(run_gemm(m, g, s));
        GemmArgs g2 = gemm_base(m, blk.cab2, W.c1, CcP, B, Hp, Wp);
        if (e == 2) {  // bf16 path: the channel-attention input stays bf16 (half the bytes for the conv store, pool and gate)
          g2.out_T = W.t2;
          g2.ld_T = CP;
        } else {
          g2.out_f32 = W.t2;
          g2.ld_f32 = CP;
        }
        SSR_TRY(run_gemm(m, g2, s));
      }
      {  // (S)W-MSA over 16x16 windows (hat.py:167-183)
        GemmArgs g = gemm_base(m, blk.qkv, W.xn, CP, B, Hp, Wp);
        g.out_T = W.qkv;
        g.ld_T = 3 * L.QP;
        SSR_TRY(run_gemm(m, g, s));
        AttnArgs a;
        memset(&a, 0, sizeof(a));
        a.qkv = W.qkv; a.ld_qkv = 3 * L.QP; a.QP = L.QP; a.o = W.o; a.ld_o = L.QP;
        a.bias = m->dev<float>(blk.bias_off);
        a.B = B; a.H = Hp; a.W = Wp; a.ws = c.window_size; a.shift = (bi % 2 == 0) ? 0 : c.window_size / 2;
        a.heads = L.heads; a.d = L.d; a.DP = L.DP; a.elem = e;
        static const bool no_tc = getenv("STUDIOSR_B200_HAT_ATTN_MMA") != nullptr;  // experiments: force the mma.sync kernel
        if (e == 2 && a.DP == 32 && a.ws == 16 && !no_tc)
          SSR_TRY(launch_attn_tc(a, 0, s));  // bf16, 16x16 windows: tcgen05 / TMEM two-pass softmax (k_attn_tc.cu)
        else if (e == 2 && (a.DP == 16 || a.DP == 32) && (a.ws * a.ws) % 64 == 0)
          SSR_TRY(launch_attn_flash(a, 0, s));  // other bf16 shapes: mma.sync tiles with an online softmax over key chunks
        else
          SSR_TRY(launch_attn_simt(a, s));
        if (!fused) {
          GemmArgs gp = gemm_base(m, blk.proj, W.o, L.QP, B, Hp, Wp);
          gp.res = shortcut;
          gp.ldres = CP;
          gp.out_f32 = W.t;
          gp.ld_f32 = CP;
          SSR_TRY(run_gemm(m, gp, s));
        }
      }
      {  // x = shortcut + attn + conv_scale * (t2 * sigmoid-gate(mean t2))  (hat.py:188, :25-38); in the fused path the
         // channel-attention term joins the shortcut first and the projection is part of the tail kernel
        CaArgs ca;
        memset(&ca, 0, sizeof(ca));
        ca.t = W.t2; ca.elem_t = e == 2 ? 2 : 4; ca.res = fused ? shortcut : W.t; ca.ld = CP; ca.B = B; ca.HW = Hp * Wp; ca.C = m->C; ca.CP = CP; ca.R = R;
        ca.W1 = m->dev<float>(blk.ca_w1); ca.b1 = m->dev<float>(blk.ca_b1);
        ca.W2 = m->dev<float>(blk.ca_w2); ca.b2 = m->dev<float>(blk.ca_b2);
        ca.partial = W.partial; ca.nsplit = W.nsplit;
        ca.out_f32 = W.t; ca.out_T = nullptr; ca.elem = e; ca.round_tf32 = rtf; ca.scale = c.conv_scale;
        SSR_TRY(launch_channel_attention(ca, s));
      }
      const LNp* next = bi + 1 < depth ? &L.blocks[bi + 1].norm1 : &L.ocab.norm1;
      if (fused) {  // proj + residual + LN2 + fc1 + GELU + fc2 + residual + next norm1 in ONE kernel (k_swin_tail.cu)
        MlpFusedArgs f;
        memset(&f, 0, sizeof(f));
        f.o = W.o; f.ld_o = L.QP; f.M = T; f.C = m->C; f.Hid = m->HID; f.CP = CP; f.HP = m->HP; f.QP = L.QP;
        f.Wp = m->arena + blk.proj.w_off; f.W1 = m->arena + blk.fc1.w_off; f.W2 = m->arena + blk.fc2.w_off;
        f.bp = m->dev<float>(blk.proj.b_off); f.b2 = m->dev<float>(blk.fc2.b_off);
        f.res = W.t; f.ldres = CP; f.eps = 1e-5f;
        f.out_f32 = W.t; f.ld_f32 = CP; f.out_ln = W.xn; f.ld_ln = CP;
        f.g3 = m->dev<float>(next->g_off); f.be3 = m->dev<float>(next->b_off);
        SSR_TRY(launch_mlp_fused(f, s));
      } else {
        SSR_TRY(mlp(blk, false, next));
      }
    }
    {  // OCAB (hat.py:240-293); xn = ocab.norm1(x) on entry
      const Block& blk = L.ocab;
      GemmArgs g = gemm_base(m, blk.qkv, W.xn, CP, B, Hp, Wp);
      g.out_T = W.qkv;
      g.ld_T = 3 * L.QP;
      SSR_TRY(run_gemm(m, g, s));
      AttnArgs a;
      memset(&a, 0, sizeof(a));
      a.qkv = W.qkv; a.ld_qkv = 3 * L.QP; a.QP = L.QP; a.o = W.o; a.ld_o = L.QP;
      a.bias = m->dev<float>(blk.bias_off);
      a.B = B; a.H = Hp; a.W = Wp; a.ws = c.window_size; a.kws = wse; a.heads = L.heads; a.d = L.d; a.DP = L.DP; a.elem = e;
      static const bool no_tc = getenv("STUDIOSR_B200_HAT_ATTN_MMA") != nullptr;
      if (e == 2 && a.DP == 32 && a.ws == 16 && a.kws == 24 && !no_tc)
        SSR_TRY(launch_attn_tc(a, 1, s));  // bf16: tcgen05 / TMEM, the 24x24 key window is one zero-filled TMA box
      else if (e == 2 && (a.DP == 16 || a.DP == 32) && (a.ws * a.ws) % 64 == 0 && (a.kws * a.kws) % 64 == 0)
        SSR_TRY(launch_attn_flash(a, 1, s));  // other bf16 shapes: mma.sync tiles with an online softmax over key chunks
      else
        SSR_TRY(launch_attn_oca(a, s));
      if (fused) {
        MlpFusedArgs f;
        memset(&f, 0, sizeof(f));
        f.o = W.o; f.ld_o = L.QP; f.M = T; f.C = m->C; f.Hid = m->HID; f.CP = CP; f.HP = m->HP; f.QP = L.QP;
        f.Wp = m->arena + blk.proj.w_off; f.W1 = m->arena + blk.fc1.w_off; f.W2 = m->arena + blk.fc2.w_off;
        f.bp = m->dev<float>(blk.proj.b_off); f.b2 = m->dev<float>(blk.fc2.b_off);
        f.res = W.t; f.ldres = CP; f.eps = 1e-5f;
        f.out_T = W.tb; f.ld_T = CP;
        SSR_TRY(launch_mlp_fused(f, s));
      } else {
        GemmArgs gp = gemm_base(m, blk.proj, W.o, L.QP, B, Hp, Wp);
        gp.res = W.t;
        gp.ldres = CP;
        gp.out_f32 = W.t;
        gp.ld_f32 = CP;
        set_ln(m, gp, blk.norm2, W.xn, CP);
        SSR_TRY(run_gemm(m, gp, s));
        SSR_TRY(mlp(blk, true, nullptr));
      }
    }
    {  // group conv + group residual (hat.py:385); epilogue = next group's first norm1 or the final norm
      GemmArgs g = gemm_base(m, L.conv, W.tb, CP, B, Hp, Wp);
      g.res = W.g;
      g.ldres = CP;
      if (li + 1 < nL) {
        g.out_f32 = W.g;
        g.ld_f32 = CP;
        set_ln(m, g, m->layers[li + 1].blocks[0].norm1, W.xn, CP);
      } else {
        set_ln(m, g, m->final_norm, W.xn, CP);
      }
      SSR_TRY(run_gemm(m, g, s));
    }
  }
  {  // conv_after_body + long skip (hat.py:549)
    GemmArgs g = gemm_base(m, m->conv_after_body, W.xn, CP, B, Hp, Wp);
    g.res = W.x0;
    g.ldres = CP;
    g.out_T = W.tb;
    g.ld_T = CP;
    SSR_TRY(run_gemm(m, g, s));
  }
  float shift[3] = {kRgbMean[0], kRgbMean[1], kRgbMean[2]};
  GemmArgs g = gemm_base(m, m->conv_before_up, W.tb, CP, B, Hp, Wp);
  g.act = ACT_LEAKY;
  g.slope = 0.01f;
  g.out_T = W.cbu;
  g.ld_T = 64;
  SSR_TRY(run_gemm(m, g, s));
  return run_tail(m, W.cbu, 64, B, Hp, Wp, W.hr[0], W.hr[1], h, w, shift, c.img_range, out, s);
}

struct RcanWs {
  float *x, *gin, *r, *t2, *partial;
  void *rb, *tmp, *hr[2];
  int nsplit;
};
static size_t plan_rcan(const ssr_model* m, void* base, int B, int H, int W, RcanWs* w) {
  Carver c(base);
  const size_t T = (size_t)B * H * W, e = m->elem;
  const int HW = H * W;
  w->nsplit = HW >= 16384 ? 64 : (HW + 255) / 256;
  w->x = (float*)c.take(T * m->FP * 4);
  w->gin = (float*)c.take(T * m->FP * 4);
  w->r = (float*)c.take(T * m->FP * 4);
  w->t2 = (float*)c.take(T * m->FP * 4);
  w->partial = (float*)c.take((size_t)B * w->nsplit * m->F * 4);
  w->rb = c.take(T * m->FP * e);
  w->tmp = c.take(T * m->FP * e);
  size_t need[2] = {0, 0};
  size_t px = T;
  for (size_t i = 0; i < m->up.size(); ++i) {
    px *= (size_t)m->up[i].ps_r * m->up[i].ps_r;
    if (px * m->FP > need[i & 1]) need[i & 1] = px * m->FP;
  }
  for (int i = 0; i < 2; ++i) w->hr[i] = need[i] ? c.take(need[i] * e) : nullptr;
  return c.off + 1024;
}

static int forward_rcan(ssr_model* m, const InputSpec& in, const OutputSpec& out, int B, int h, int w, void* ws,
                        size_t ws_bytes, cudaStream_t s) {
  const ssr_model_config& c = m->cfg;
  RcanWs W;
  const size_t need = plan_rcan(m, ws, B, h, w, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, need);
  const int FP = m->FP, e = m->elem, R = m->F / c.reduction;
  {  // sub_mean + head (rcan.py:69-70)
    ConvFirstArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in.in; a.in_u8 = in.in_u8; a.fh = in.fh; a.fw = in.fw;
    a.tile_mode = in.tile_mode; a.tile = in.tile; a.stride = in.stride; a.tiles_x = in.tiles_x; a.tile_begin = in.tile_begin;
    a.h = h; a.w = w; a.Hp = h; a.Wp = w; a.pad_mode = 2; a.B = B;
    a.in_scale = (in.in_u8 && c.img_range == 1.0f) ? 1.0f / 255.0f : 1.0f;
    for (int i = 0; i < 3; ++i) a.in_shift[i] = m->sub_bias[i];
    a.Wc = m->dev<float>(m->conv_first_w);
    a.bias = m->dev<float>(m->conv_first_b);
    a.Cout = m->F;
    a.out_f32 = W.x; a.ld_f32 = FP; a.out_T = W.rb; a.ld_T = FP; a.elem = e;
    a.round_tf32 = c.precision == SSR_PREC_TF32;
    SSR_TRY(launch_conv_first(a, s));
  }
  const float* gcur = W.x;  // input of the current residual group (fp32 stream)
  for (int g = 0; g < c.n_resgroups; ++g) {
    const float* rcur = gcur;  // input of the current RCAB
    for (int b = 0; b < c.n_resblocks; ++b) {  // RCAB (rcan.py:21-24)
      const int i = g * c.n_resblocks + b;
      GemmArgs ga = gemm_base(m, m->res_a[i], W.rb, FP, B, h, w);
      ga.act = ACT_RELU;
      ga.out_T = W.tmp;
      ga.ld_T = FP;
      SSR_TRY(run_gemm(m, ga, s));
      GemmArgs gb = gemm_base(m, m->res_b[i], W.tmp, FP, B, h, w);
      gb.out_f32 = W.t2;
      gb.ld_f32 = FP;
      SSR_TRY(run_gemm(m, gb, s));
      CaArgs ca;
      memset(&ca, 0, sizeof(ca));
      ca.t = W.t2; ca.res = rcur; ca.ld = FP; ca.B = B; ca.HW = h * w; ca.C = m->F; ca.CP = FP; ca.R = R;
      ca.W1 = m->dev<float>(m->ca[i].w1); ca.b1 = m->dev<float>(m->ca[i].b1);
      ca.W2 = m->dev<float>(m->ca[i].w2); ca.b2 = m->dev<float>(m->ca[i].b2);
      ca.partial = W.partial; ca.nsplit = W.nsplit;
      ca.out_f32 = W.r; ca.out_T = W.rb; ca.ld_T = FP; ca.elem = e; ca.round_tf32 = c.precision == SSR_PREC_TF32;
      ca.scale = 1.0f;
      SSR_TRY(launch_channel_attention(ca, s));
      rcur = W.r;
    }
    // group tail conv + group skip (rcan.py:33-36)
    GemmArgs gt = gemm_base(m, m->grp_tail[g], W.rb, FP, B, h, w);
    gt.res = gcur;
    gt.ldres = FP;
    gt.out_f32 = W.gin;
    gt.ld_f32 = FP;
    gt.out_T = W.rb;
    gt.ld_T = FP;
    SSR_TRY(run_gemm(m, gt, s));
    gcur = W.gin;
  }
  {  // body tail conv + long skip (rcan.py:72-73)
    GemmArgs g = gemm_base(m, m->body_tail, W.rb, FP, B, h, w);
    g.res = W.x;
    g.ldres = FP;
    g.out_T = W.tmp;
    g.ld_T = FP;
    SSR_TRY(run_gemm(m, g, s));
  }
  return run_tail(m, W.tmp, FP, B, h, w, W.hr[0], W.hr[1], h, w, m->add_bias, 1.0f, out, s);
}

// ---- HAN (han.py:90-113) ----
struct HanWs {
  float *x, *r, *t2, *partial, *stack;
  double* energy;
  void *rb, *tmp, *lam, *cat, *hr[2];
  int nsplit;
};
static size_t plan_han(const ssr_model* m, void* base, int B, int H, int W, HanWs* w) {
  Carver c(base);
  const size_t T = (size_t)B * H * W, e = m->elem;
  const int HW = H * W;
  w->nsplit = HW >= 16384 ? 64 : (HW + 255) / 256;
  w->x = (float*)c.take(T * m->FP * 4);
  w->r = (float*)c.take(T * m->FP * 4);
  w->t2 = (float*)c.take(T * m->FP * 4);
  w->stack = (float*)c.take(11 * T * m->FP * 4);  // plane n: 0 = body's last conv, 1 + k = group 9 - k (newest first, han.py:94-99)
  w->partial = (float*)c.take((size_t)B * w->nsplit * m->F * 4);
  w->energy = (double*)c.take((size_t)B * 66 * 8);
  w->rb = c.take(T * m->FP * e);
  w->tmp = c.take(T * m->FP * e);
  w->lam = c.take(T * 11 * m->FP * e);
  w->cat = c.take(T * 2 * m->FP * e);
  size_t need[2] = {0, 0};
  size_t px = T;
  for (size_t i = 0; i < m->up.size(); ++i) {
    px *= (size_t)m->up[i].ps_r * m->up[i].ps_r;
    if (px * m->FP > need[i & 1]) need[i & 1] = px * m->FP;
  }
  for (int i = 0; i < 2; ++i) w->hr[i] = need[i] ? c.take(need[i] * e) : nullptr;
  return c.off + 1024;
}

static int forward_han(ssr_model* m, const InputSpec& in, const OutputSpec& out, int B, int h, int w, void* ws, size_t ws_bytes,
                       cudaStream_t s) {
  const ssr_model_config& c = m->cfg;
  HanWs W;
  const size_t need = plan_han(m, ws, B, h, w, &W);
  SSR_CHECK(ws && need <= ws_bytes, SSR_E_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, need);
  const int FP = m->FP, F = m->F, e = m->elem, R = F / c.reduction, ng = c.n_resgroups;
  const int rtf = c.precision == SSR_PREC_TF32;
  const size_t T = (size_t)B * h * w, plane = T * FP;
  {  // sub_mean + head (han.py:91-92)
    ConvFirstArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in.in; a.in_u8 = in.in_u8; a.fh = in.fh; a.fw = in.fw;
    a.tile_mode = in.tile_mode; a.tile = in.tile; a.stride = in.stride; a.tiles_x = in.tiles_x; a.tile_begin = in.tile_begin;
    a.h = h; a.w = w; a.Hp = h; a.Wp = w; a.pad_mode = 2; a.B = B;
    a.in_scale = (in.in_u8 && c.img_range == 1.0f) ? 1.0f / 255.0f : 1.0f;
    for (int i = 0; i < 3; ++i) a.in_shift[i] = m->sub_bias[i];
    a.Wc = m->dev<float>(m->conv_first_w);
    a.bias = m->dev<float>(m->conv_first_b);
    a.Cout = F;
    a.out_f32 = W.x; a.ld_f32 = FP; a.out_T = W.rb; a.ld_T = FP; a.elem = e;
    a.round_tf32 = rtf;
    SSR_TRY(launch_conv_first(a, s));
  }
  const float* gcur = W.x;  // input of the current residual group (fp32 stream)
  for (int g = 0; g < ng; ++g) {
    const float* rcur = gcur;
    for (int b = 0; b < c.n_resblocks; ++b) {  // RCAB (rcan.py:21-24)
      const int i = g * c.n_resblocks + b;
      GemmArgs ga = gemm_base(m, m->res_a[i], W.rb, FP, B, h, w);
      ga.act = ACT_RELU;
      ga.out_T = W.tmp;
      ga.ld_T = FP;
      SSR_TRY(run_gemm(m, ga, s));
      GemmArgs gb = gemm_base(m, m->res_b[i], W.tmp, FP, B, h, w);
      gb.out_f32 = W.t2;
      gb.ld_f32 = FP;
      SSR_TRY(run_gemm(m, gb, s));
      CaArgs ca;
      memset(&ca, 0, sizeof(ca));
      ca.t = W.t2; ca.res = rcur; ca.ld = FP; ca.B = B; ca.HW = h * w; ca.C = F; ca.CP = FP; ca.R = R;
      ca.W1 = m->dev<float>(m->ca[i].w1); ca.b1 = m->dev<float>(m->ca[i].b1);
      ca.W2 = m->dev<float>(m->ca[i].w2); ca.b2 = m->dev<float>(m->ca[i].b2);
      ca.partial = W.partial; ca.nsplit = W.nsplit;
      ca.out_f32 = W.r; ca.out_T = W.rb; ca.ld_T = FP; ca.elem = e; ca.round_tf32 = rtf;
      ca.scale = 1.0f;
      SSR_TRY(launch_channel_attention(ca, s));
      rcur = W.r;
    }
    // group tail conv + group skip (rcan.py:33-36); the group's output is plane ng - g of the stack
    float* slot = W.stack + (size_t)(ng - g) * plane;
    GemmArgs gt = gemm_base(m, m->grp_tail[g], W.rb, FP, B, h, w);
    gt.res = gcur;
    gt.ldres = FP;
    gt.out_f32 = slot;
    gt.ld_f32 = FP;
    gt.out_T = W.rb;
    gt.ld_T = FP;
    SSR_TRY(run_gemm(m, gt, s));
    gcur = slot;
  }
  {  // body's last conv (han.py:93-99, no skip here): plane 0
    GemmArgs g = gemm_base(m, m->body_tail, W.rb, FP, B, h, w);
    g.out_f32 = W.stack;
    g.ld_f32 = FP;
    SSR_TRY(run_gemm(m, g, s));
  }
  // layer attention over the 11 planes (han.py:101) -> last_conv's operand [pixel][11 F]
  SSR_TRY(launch_han_lam(W.stack, plane, FP, B, h * w, F, W.energy, m->dev<float>(m->la_gamma), W.lam, 11 * FP, e, rtf, s));
  {  // out2 = last_conv(.) (han.py:102) -> channels [F, 2F) of the concatenation
    GemmArgs g = gemm_base(m, m->han_last_conv, W.lam, 11 * FP, B, h, w);
    g.out_T = reinterpret_cast<uint8_t*>(W.cat) + (size_t)FP * e;
    g.ld_T = 2 * FP;
    SSR_TRY(run_gemm(m, g, s));
  }
  // out1 = csa(out1) (han.py:104) -> channels [0, F)
  SSR_TRY(launch_han_csam(W.stack, FP, B, h, w, F, m->dev<float>(m->csa_w), m->dev<float>(m->csa_b), m->dev<float>(m->csa_gamma), W.cat,
                          2 * FP, e, rtf, s));
  {  // res = last(cat) + x (han.py:105-108)
    GemmArgs g = gemm_base(m, m->han_last, W.cat, 2 * FP, B, h, w);
    g.res = W.x;
    g.ldres = FP;
    g.out_T = W.tmp;
    g.ld_T = FP;
    SSR_TRY(run_gemm(m, g, s));
  }
  return run_tail(m, W.tmp, FP, B, h, w, W.hr[0], W.hr[1], h, w, m->add_bias, 1.0f, out, s);
}

int check_ready(ssr_model* m) {
  SSR_CHECK(m != nullptr, SSR_E_INVALID, "null model");
  SSR_CHECK(m->finalized, SSR_E_STATE, "model not finalised (call ssr_model_finalize)");
  int dev = -1;
  SSR_CUDA(cudaGetDevice(&dev));
  if (dev != m->device) SSR_CUDA(cudaSetDevice(m->device));
  return SSR_OK;
}

static int forward_any(ssr_model* m, const InputSpec& in, const OutputSpec& out, int B, int h, int w, int pad_mode,
                       void* ws, size_t ws_bytes, cudaStream_t s) {
  SSR_CHECK(B > 0 && h > 0 && w > 0, SSR_E_INVALID, "bad shape B=%d H=%d W=%d", B, h, w);
  if (is_swin(m)) return forward_swinir(m, in, out, B, h, w, pad_mode, ws, ws_bytes, s);
  if (m->cfg.arch == SSR_ARCH_RCAN) return forward_rcan(m, in, out, B, h, w, ws, ws_bytes, s);
  if (m->cfg.arch == SSR_ARCH_HAN) return forward_han(m, in, out, B, h, w, ws, ws_bytes, s);
  if (m->cfg.arch == SSR_ARCH_HAT) return forward_hat(m, in, out, B, h, w, ws, ws_bytes, s);
  return forward_edsr(m, in, out, B, h, w, ws, ws_bytes, s);
}

static size_t workspace_any(const ssr_model* m, int B, int H, int W, int pad_mode) {
  if (is_swin(m)) {
    int Hp, Wp;
    padded_size(m, H, W, pad_mode, &Hp, &Wp);
    SwinWs w;
    return plan_swinir(m, nullptr, B, Hp, Wp, &w);
  }
  if (m->cfg.arch == SSR_ARCH_RCAN) {
    RcanWs w;
    return plan_rcan(m, nullptr, B, H, W, &w);
  }
  if (m->cfg.arch == SSR_ARCH_HAN) {
    HanWs w;
    return plan_han(m, nullptr, B, H, W, &w);
  }
  if (m->cfg.arch == SSR_ARCH_HAT) {
    int Hp, Wp;
    padded_size(m, H, W, SSR_PAD_TRAIN, &Hp, &Wp);
    HatWs w;
    return plan_hat(m, nullptr, B, Hp, Wp, &w);
  }
  EdsrWs w;
  return plan_edsr(m, nullptr, B, H, W, &w);
}

static int num_tiles_1d(int L, int tile, int stride) {
  if (L <= tile) return 1;
  return (L - tile + stride - 1) / stride + 1;
}

}  // namespace ssr

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int ssr_version(void) { return SSR_VERSION; }
const char* ssr_last_error(void) { return g_err; }
int64_t ssr_launch_count(void) { return (int64_t)g_launches.load(); }
void ssr_note_graph_replay(int64_t kernels) { g_launches.fetch_add((long long)kernels, std::memory_order_relaxed); }

int ssr_profile_begin(void) {
  for (ProfRec* r : g_prof) {
    cudaEventDestroy(r->e0);
    cudaEventDestroy(r->e1);
    delete r;
  }
  g_prof.clear();
  g_prof_on = true;
  return SSR_OK;
}

int ssr_profile_end(char* json, size_t cap) {
  g_prof_on = false;
  SSR_CUDA(cudaDeviceSynchronize());
  struct Agg {
    double ms = 0, flops = 0, bytes = 0;
    long long n = 0;
  };
  std::map<std::string, Agg> agg;
  for (ProfRec* r : g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r->e0, r->e1);
    Agg& a = agg[r->cls];
    a.ms += ms;
    a.flops += r->flops;
    a.bytes += r->bytes;
    a.n += 1;
    cudaEventDestroy(r->e0);
    cudaEventDestroy(r->e1);
    delete r;
  }
  g_prof.clear();
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
    out += buf;
    first = false;
  }
  out += "}";
  SSR_CHECK(json && out.size() + 1 <= cap, SSR_E_INVALID, "ssr_profile_end: buffer too small (%zu needed)", out.size() + 1);
  memcpy(json, out.c_str(), out.size() + 1);
  return SSR_OK;
}

int ssr_device_check(int device) {
  cudaDeviceProp p;
  SSR_CUDA(cudaGetDeviceProperties(&p, device));
  SSR_CHECK(p.major == 10, SSR_E_ARCH, "device %d is sm_%d%d; libssr_b200 needs sm_100 (B200) and has no fallback", device,
            p.major, p.minor);
  return SSR_OK;
}

int ssr_model_create(const ssr_model_config* cfg, int device, ssr_model_t** out) {
  SSR_CHECK(cfg && out, SSR_E_INVALID, "null argument");
  SSR_CHECK(cfg->arch >= SSR_ARCH_SWINIR && cfg->arch <= SSR_ARCH_SWINFIR, SSR_E_INVALID, "unknown arch %d", cfg->arch);
  SSR_CHECK(cfg->precision >= 0 && cfg->precision <= 3, SSR_E_INVALID, "unknown precision %d", cfg->precision);
  SSR_CHECK(cfg->scale >= 1 && cfg->scale <= 8, SSR_E_INVALID, "bad scale %d", cfg->scale);
  if (cfg->arch == SSR_ARCH_SWINIR || cfg->arch == SSR_ARCH_SWINFIR || cfg->arch == SSR_ARCH_HAT)
    SSR_CHECK(cfg->n_layers > 0 && cfg->n_layers <= SSR_MAX_LAYERS, SSR_E_INVALID, "bad n_layers %d", cfg->n_layers);
  SSR_TRY(ssr_device_check(device));
  ssr_model* m = new ssr_model();
  m->cfg = *cfg;
  m->device = device;
  m->elem = cfg->precision == SSR_PREC_BF16 ? 2 : 4;
  *out = m;
  return SSR_OK;
}

int ssr_model_set_param(ssr_model_t* m, const char* name, const float* host_data, int64_t numel) {
  SSR_CHECK(m && name && host_data && numel > 0, SSR_E_INVALID, "bad argument to ssr_model_set_param");
  m->params[name].assign(host_data, host_data + numel);
  m->finalized = false;
  return SSR_OK;
}

int ssr_model_finalize(ssr_model_t* m) {
  SSR_CHECK(m != nullptr, SSR_E_INVALID, "null model");
  SSR_CUDA(cudaSetDevice(m->device));
  m->host_arena.clear();
  int r = is_swin(m) ? finalize_swinir(m)
          : m->cfg.arch == SSR_ARCH_RCAN ? finalize_rcan(m)
          : m->cfg.arch == SSR_ARCH_HAN  ? finalize_han(m)
          : m->cfg.arch == SSR_ARCH_HAT  ? finalize_hat(m)
                                         : finalize_edsr(m);
  if (r != SSR_OK) return r;
  if (m->arena && m->arena_bytes < m->host_arena.size()) {
    cudaFree(m->arena);
    m->arena = nullptr;
  }
  if (!m->arena) {
    m->arena_bytes = m->host_arena.size();
    SSR_CUDA(cudaMalloc(&m->arena, m->arena_bytes));
  }
  if (getenv("STUDIOSR_B200_DEBUG_NAN") && m->cfg.arch == SSR_ARCH_SWINIR && !m->layers.empty() && m->layers[0].blocks[0].whp_off) {
    const Block& b0 = m->layers[0].blocks[0];
    auto scan = [&](const char* nm, size_t off, size_t n) {
      size_t bad = 0, first = 0;
      for (size_t i = 0; i < n; ++i) {
        __nv_bfloat16 h;
        memcpy(&h, m->host_arena.data() + off + 2 * i, 2);
        if (!isfinite(__bfloat162float(h))) { if (!bad) first = i; ++bad; }
      }
      fprintf(stderr, "[debug] finalize: %s at arena+%zu: %zu non-finite of %zu (first at %zu)\n", nm, off, bad, n, first);
    };
    scan("Whp", b0.whp_off, kAttnWhpBytes / 2);
    scan("btab", b0.btab_off, kAttnBiasBytes / 2);
  }
  SSR_CUDA(cudaMemcpy(m->arena, m->host_arena.data(), m->host_arena.size(), cudaMemcpyHostToDevice));
  // a pageable-memory cudaMemcpy may return before its DMA has landed and only the legacy stream is ordered behind it: callers
  // launch on non-blocking streams (CUDA-graph capture streams, torch side streams)
  SSR_CUDA(cudaDeviceSynchronize());
  m->host_arena.clear();
  m->host_arena.shrink_to_fit();
  m->finalized = true;
  return SSR_OK;
}

void ssr_model_destroy(ssr_model_t* m) {
  if (!m) return;
  train_state_destroy(m);
  if (m->arena) cudaFree(m->arena);
  delete m;
}

size_t ssr_model_workspace_bytes(const ssr_model_t* m, int B, int H, int W, int pad_mode) {
  if (!m || !m->finalized) return 0;
  return workspace_any(m, B, H, W, pad_mode);
}

int ssr_model_forward(ssr_model_t* m, const float* x, float* y, int B, int H, int W, int pad_mode, void* workspace,
                      size_t workspace_bytes, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(x && y, SSR_E_INVALID, "null tensor");
  InputSpec in{x, 0, H, W, 0, 0, 0, 0, 0};
  OutputSpec out{y, nullptr};
  return forward_any(m, in, out, B, H, W, pad_mode, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ssr_model_upscale_u8(ssr_model_t* m, const uint8_t* img, uint8_t* outp, int B, int H, int W, void* workspace,
                         size_t workspace_bytes, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(img && outp, SSR_E_INVALID, "null tensor");
  InputSpec in{img, 1, H, W, 0, 0, 0, 0, 0};
  OutputSpec out{nullptr, outp};
  return forward_any(m, in, out, B, H, W, SSR_PAD_EVAL, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ssr_tiled_num_tiles(int H, int W, int tile, int overlap) {
  if (tile <= overlap || H <= 0 || W <= 0) return 0;
  return num_tiles_1d(H, tile, tile - overlap) * num_tiles_1d(W, tile, tile - overlap);
}

static size_t tiled_layout(const ssr_model_t* m, int H, int W, int tile, int overlap, int chunk, size_t* tiles_off,
                           size_t* dev_in_off, size_t* dev_out_off) {
  const int th = tile < H ? tile : H, tw = tile < W ? tile : W;
  const int nt = ssr_tiled_num_tiles(H, W, tile, overlap);
  if (chunk <= 0 || chunk > nt) chunk = nt;
  const int s = m->cfg.scale;
  size_t off = (workspace_any(m, chunk, th, tw, SSR_PAD_EVAL) + 1023) & ~(size_t)1023;
  *tiles_off = off;
  off += ((size_t)nt * 3 * th * s * tw * s * 4 + 1023) & ~(size_t)1023;
  *dev_in_off = off;
  off += ((size_t)H * W * 3 + 1023) & ~(size_t)1023;
  *dev_out_off = off;
  off += ((size_t)H * s * W * s * 3 + 1023) & ~(size_t)1023;
  return off;
}

size_t ssr_model_tiled_workspace_bytes(const ssr_model_t* m, int H, int W, int tile, int overlap, int chunk_tiles) {
  if (!m || !m->finalized || tile <= overlap) return 0;
  size_t a, b, c;
  return tiled_layout(m, H, W, tile, overlap, chunk_tiles, &a, &b, &c);
}

size_t ssr_tiled_tile_elems(const ssr_model_t* m, int H, int W, int tile) {
  if (!m || tile <= 0) return 0;
  const int th = tile < H ? tile : H, tw = tile < W ? tile : W, s = m->cfg.scale;
  return (size_t)3 * th * s * tw * s;
}

size_t ssr_model_tiles_workspace_bytes(const ssr_model_t* m, int H, int W, int tile, int max_tiles_per_pass) {
  if (!m || !m->finalized || tile <= 0 || max_tiles_per_pass <= 0) return 0;
  const int th = tile < H ? tile : H, tw = tile < W ? tile : W;
  return workspace_any(m, max_tiles_per_pass, th, tw, SSR_PAD_EVAL);
}

int ssr_model_tiles_u8(ssr_model_t* m, const uint8_t* frame, float* tiles_out, int H, int W, int tile, int overlap, int tile_begin,
                       int tile_end, int chunk_tiles, void* workspace, size_t workspace_bytes, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(frame && tiles_out && workspace, SSR_E_INVALID, "null buffer");
  SSR_CHECK(tile > overlap && overlap >= 0, SSR_E_INVALID, "tile %d must exceed overlap %d", tile, overlap);
  const int th = tile < H ? tile : H, tw = tile < W ? tile : W, stride = tile - overlap;
  const int tiles_x = num_tiles_1d(W, tile, stride), tiles_y = num_tiles_1d(H, tile, stride);
  const int nt = tiles_x * tiles_y;
  if (tile_end < 0) tile_end = nt;
  SSR_CHECK(tile_begin >= 0 && tile_begin <= tile_end && tile_end <= nt, SSR_E_INVALID, "tile range [%d, %d) outside [0, %d)", tile_begin,
            tile_end, nt);
  const int n = tile_end - tile_begin;
  if (n == 0) return SSR_OK;
  int chunk = chunk_tiles;
  if (chunk <= 0 || chunk > n) chunk = n;
  const size_t need = workspace_any(m, chunk, th, tw, SSR_PAD_EVAL);
  SSR_CHECK(need <= workspace_bytes, SSR_E_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  const size_t elems = ssr_tiled_tile_elems(m, H, W, tile);
  for (int t0 = 0; t0 < n; t0 += chunk) {
    const int nb = n - t0 < chunk ? n - t0 : chunk;
    InputSpec in{frame, 1, H, W, 1, tile, stride, tiles_x, tile_begin + t0};
    OutputSpec out{tiles_out + (size_t)t0 * elems, nullptr};
    SSR_TRY(forward_any(m, in, out, nb, th, tw, SSR_PAD_EVAL, workspace, workspace_bytes, (cudaStream_t)stream));
  }
  return SSR_OK;
}

int ssr_model_blend_tiles_u8(ssr_model_t* m, const float* tiles, uint8_t* outp, int H, int W, int tile, int overlap, int row_begin,
                             int row_end, void* stream) {
  SSR_CHECK(m && tiles && outp, SSR_E_INVALID, "null argument");
  SSR_CHECK(tile > overlap && overlap >= 0, SSR_E_INVALID, "tile %d must exceed overlap %d", tile, overlap);
  const int stride = tile - overlap;
  BlendArgs b;
  b.tiles = tiles;
  b.out = outp;
  b.H = H;
  b.W = W;
  b.scale = m->cfg.scale;
  b.tile = tile;
  b.overlap = overlap;
  b.tiles_x = num_tiles_1d(W, tile, stride);
  b.tiles_y = num_tiles_1d(H, tile, stride);
  b.u8_scale = m->cfg.img_range == 1.0f ? 255.0f : 1.0f;
  b.row_begin = row_begin;
  b.row_end = row_end < 0 ? H * m->cfg.scale : row_end;
  return launch_blend(b, (cudaStream_t)stream);
}

int ssr_model_upscale_tiled_u8(ssr_model_t* m, const uint8_t* frame, uint8_t* outp, int H, int W, int tile, int overlap,
                               int chunk_tiles, void* workspace, size_t workspace_bytes, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(frame && outp && workspace, SSR_E_INVALID, "null buffer");
  SSR_CHECK(tile > overlap && overlap >= 0, SSR_E_INVALID, "tile %d must exceed overlap %d", tile, overlap);
  size_t tiles_off, in_off, out_off;
  const size_t need = tiled_layout(m, H, W, tile, overlap, chunk_tiles, &tiles_off, &in_off, &out_off);
  SSR_CHECK(need <= workspace_bytes, SSR_E_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  float* tiles = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + tiles_off);
  SSR_TRY(ssr_model_tiles_u8(m, frame, tiles, H, W, tile, overlap, 0, -1, chunk_tiles, workspace, tiles_off, stream));
  return ssr_model_blend_tiles_u8(m, tiles, outp, H, W, tile, overlap, 0, -1, stream);
}

int ssr_model_upscale_tiled_u8_host(ssr_model_t* m, const uint8_t* frame_host, uint8_t* out_host, int H, int W, int tile,
                                    int overlap, int chunk_tiles, void* workspace, size_t workspace_bytes, void* stream) {
  SSR_TRY(check_ready(m));
  SSR_CHECK(frame_host && out_host && workspace, SSR_E_INVALID, "null buffer");
  SSR_CHECK(tile > overlap && overlap >= 0, SSR_E_INVALID, "tile %d must exceed overlap %d", tile, overlap);
  cudaStream_t s = (cudaStream_t)stream;
  size_t tiles_off, in_off, out_off;
  const size_t need = tiled_layout(m, H, W, tile, overlap, chunk_tiles, &tiles_off, &in_off, &out_off);
  SSR_CHECK(need <= workspace_bytes, SSR_E_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  uint8_t* din = reinterpret_cast<uint8_t*>(workspace) + in_off;
  uint8_t* dout = reinterpret_cast<uint8_t*>(workspace) + out_off;
  const int sc = m->cfg.scale;
  SSR_CUDA(cudaMemcpyAsync(din, frame_host, (size_t)H * W * 3, cudaMemcpyHostToDevice, s));
  SSR_TRY(ssr_model_upscale_tiled_u8(m, din, dout, H, W, tile, overlap, chunk_tiles, workspace, workspace_bytes, stream));
  SSR_CUDA(cudaMemcpyAsync(out_host, dout, (size_t)H * sc * W * sc * 3, cudaMemcpyDeviceToHost, s));
  SSR_CUDA(cudaStreamSynchronize(s));
  return SSR_OK;
}

}  // extern "C"
