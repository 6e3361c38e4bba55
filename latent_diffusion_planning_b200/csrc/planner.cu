// Planner handle: ConditionalUnet1D score network (reference networks/diffusion_nets_v2.py:104-169) and the
// reverse-diffusion loop around it (reference agent/ldp_agent.py:459-476), in two precisions:
//   fp32  - SIMT implicit GEMM + separate GroupNorm/FiLM kernels (parity gate 1e-5),
//   bf16  - one tcgen05 GEMM launch per convolution with GroupNorm/Mish/FiLM/residual fused into its epilogue and
//           the scheduler step fused into the last 1x1 convolution; a whole denoising step is one CUDA graph.
// Step- and batch-invariant work is hoisted out of the loop: the FiLM Dense of every block is split into a
// time part (table over the N timesteps, built once at create) and an observation part (one GEMM per act()).
#include <algorithm>
#include <cmath>

#include "net_common.h"

namespace ldp {

// ---------------------------------------------------------------------------------------------
// schedule (host)
// ---------------------------------------------------------------------------------------------
void ddpm_schedule_host(int n, std::vector<float>& betas, std::vector<float>& alphas, std::vector<float>& acp) {
  betas.resize(n);
  alphas.resize(n);
  acp.resize(n);
  auto alpha_bar = [](double s) {
    double c = std::cos((s + 0.008) / 1.008 * M_PI / 2.0);
    return c * c;
  };
  float run = 1.0f;
  for (int i = 0; i < n; ++i) {
    double b = 1.0 - alpha_bar((double)(i + 1) / n) / alpha_bar((double)i / n);
    betas[i] = (float)std::min(b, 0.999);
    alphas[i] = 1.0f - betas[i];
    run = run * alphas[i];                 // sequential fp32 cumprod (jnp.cumprod on f32)
    acp[i] = run;
  }
}

void ddpm_coef_host(int n, std::vector<float>& coef) {
  std::vector<float> betas, alphas, acp;
  ddpm_schedule_host(n, betas, alphas, acp);
  coef.assign((size_t)n * 8, 0.f);
  for (int t = 0; t < n; ++t) {
    // fp32 throughout, in the op order of FlaxDDPMScheduler.step
    float a_t = acp[t];
    float a_prev = t > 0 ? acp[t - 1] : 1.0f;      // the `t > 0` select of alpha_prod_t_prev
    float b_t = 1.0f - a_t, b_prev = 1.0f - a_prev;
    float c0 = (std::sqrt(a_prev) * betas[t]) / b_t;
    float ct = std::sqrt(alphas[t]) * b_prev / b_t;
    float var = (1.0f - a_prev) / (1.0f - a_t) * betas[t];
    var = std::max(var, 1e-20f);
    float* c = &coef[(size_t)t * 8];
    c[0] = 1.0f / std::sqrt(a_t);
    c[1] = std::sqrt(b_t);
    c[2] = c0;
    c[3] = ct;
    c[4] = t > 0 ? std::sqrt(var) : 0.f;            // the `t > 0` select of the noise term
    c[5] = std::sqrt(a_prev);
    c[6] = std::sqrt(b_prev);
  }
}

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
enum ConvKind { CONV_K = 0, CONV_DOWN = 1, CONV_UP = 2 };

struct CrbW {
  int cin = 0, cout = 0;
  bool proj = false;
  int film_off = 0;
  const float *c1w, *c1b, *g1s, *g1b, *fw, *fb, *c2w, *c2b, *g2s, *g2b, *rw, *rb;
};

// Packed weights + stage table of one convolution on the tcgen05 path (layouts described at get_packed).
struct ConvPack : PackedW {
  int n_acc = 1;
  int shift[5] = {0, 0, 0, 0, 0};
  int w_max = 1;
  int num_kb_main = 0;             // stages before the aux (residual projection) part
  bool tapacc = false;
};

struct PlanWs {
  int B = 0, T = 0;
  uint64_t last_use = 0;         // LRU stamp (ws_evict_lru)
  Arena arena;
  float* otab = nullptr;
  float* otab_q = nullptr;       // half2 (scale, shift) pairs, quad-transposed, for the tcgen05 epilogues (film_pack_kernel)
  __nv_bfloat16* cond_bf16 = nullptr;   // Mish(cond) in bf16: A operand of the observation-part FiLM GEMM (bf16 path)
  int ld_cb = 0;
  TcGemm otab_op;
  bool otab_ready = false;
  float* x_state = nullptr;
  float* eps_buf = nullptr;
  int32_t* step_dev = nullptr;
  unsigned int* done_counter = nullptr;
  DdpmCall* call_dev = nullptr;
  // fp32 path
  bool f32_ready = false;
  float *f_tmp = nullptr, *f_h1 = nullptr, *f_res = nullptr, *f_final = nullptr;
  std::vector<float*> f_out;       // one per CRB
  std::vector<float*> f_down, f_up;
  // bf16 path
  bool bf16_ready = false;
  __nv_bfloat16* x_bf16 = nullptr;
  int ld_xb = 0;
  std::vector<TcGemm> ops;         // one denoising step; last op = final 1x1 with the DDPM epilogue
  struct Tap { int id; ActBf16 act; long long rows; };
  std::vector<Tap> taps;           // named intermediates for per-layer parity tests (ldp_planner_read_activation)
  cudaGraphExec_t graph = nullptr;
  cudaGraph_t graph_src = nullptr;
  // the whole n-step loop as ONE graph (keyed by the step count): consecutive launches of the per-step graph are ordered by the
  // stream, i.e. by a full kernel boundary, while inside a graph every layer-to-layer edge is a programmatic (PDL) edge
  std::map<int, cudaGraphExec_t> loop_graphs;
  // persistent loop kernel (planner_loop.cu)
  int loop_state = 0;              // 0 not prepared, 1 ready, -1 unsupported for this shape
  std::vector<TcGemm> loop_ops;    // host copy of the loop kernel's layer table
  int* group_counter = nullptr;
  int n_groups = 0;
  long long* loop_dbg = nullptr;
  ~PlanWs() {
    if (graph) cudaGraphExecDestroy(graph);
    if (graph_src) cudaGraphDestroy(graph_src);
    for (auto& kv : loop_graphs)
      if (kv.second) cudaGraphExecDestroy(kv.second);
  }
};

}  // namespace ldp

using namespace ldp;

struct LdpPlanner {
  LdpUnetConfig cfg;
  Arena arena;
  float* blob = nullptr;
  const float *t0w, *t0b, *t1w, *t1b;
  std::vector<CrbW> crb;
  std::vector<const float*> down_w, down_b, up_w, up_b;
  std::vector<float*> up_bias2;     // [b|b] for the two output phases of the transposed conv
  const float *fcw, *fcb, *fgs, *fgb, *ow, *ob;
  int sum_c2 = 0;
  float* ttab = nullptr;     // [n_train][sum_c2]  time part of every FiLM Dense (+ its bias)
  float* wc_all = nullptr;   // [Dc][sum_c2]       observation part of every FiLM Dense
  float* coef = nullptr;     // [n_train][8]
  std::map<std::pair<int, int>, std::unique_ptr<PlanWs>> ws;
  uint64_t use_clock = 0;
  std::map<std::pair<int, int>, ConvPack> packed;   // (op id * 2 + layout, T_in)
  PackedW pw_otab;                  // wc_all^T packed K-major in bf16 (observation part of every FiLM Dense)
  bool otab_packed = false;
  bool use_graph = true;
};

namespace ldp {

static std::vector<std::tuple<int, int, bool>> block_plan(const LdpUnetConfig& c) {
  std::vector<std::tuple<int, int, bool>> b;
  int ch = c.input_dim;
  for (int i = 0; i < c.n_levels; ++i) {
    b.emplace_back(ch, c.down_dims[i], true);
    b.emplace_back(c.down_dims[i], c.down_dims[i], false);
    ch = c.down_dims[i];
  }
  int mid = c.down_dims[c.n_levels - 1];
  b.emplace_back(mid, mid, false);
  b.emplace_back(mid, mid, false);
  ch = mid;
  for (int i = c.n_levels - 2; i >= 0; --i) {
    int skip = c.down_dims[i + 1];           // h.pop(): deepest remaining skip
    b.emplace_back(ch + skip, c.down_dims[i], true);
    b.emplace_back(c.down_dims[i], c.down_dims[i], false);
    ch = c.down_dims[i];
  }
  return b;
}

static int64_t unet_param_count(const LdpUnetConfig& c) {
  int64_t n = 0;
  const int64_t ds = c.step_embed_dim, cond = ds + c.global_cond_dim, k = c.kernel_size;
  n += ds * ds * 4 + ds * 4 + ds * 4 * ds + ds;
  for (auto& [cin, cout, proj] : block_plan(c)) {
    n += k * cin * cout + cout + 2 * cout;
    n += cond * 2 * cout + 2 * cout;
    n += k * (int64_t)cout * cout + cout + 2 * cout;
    if (proj) n += (int64_t)cin * cout + cout;
  }
  for (int i = 0; i < c.n_levels - 1; ++i) n += 3 * (int64_t)c.down_dims[i] * c.down_dims[i] + c.down_dims[i];
  for (int i = 0; i < c.n_levels - 1; ++i) n += 4 * (int64_t)c.down_dims[i] * c.down_dims[i] + c.down_dims[i];
  const int64_t d0 = c.down_dims[0];
  n += k * d0 * d0 + d0 + 2 * d0 + d0 * c.input_dim + c.input_dim;
  return n;
}

static int validate_cfg(const LdpUnetConfig* c) {
  LDP_CHECK(c != nullptr, LDP_ERR_INVALID_ARG, "null config");
  LDP_CHECK(c->input_dim > 0 && c->global_cond_dim > 0 && c->step_embed_dim >= 4 && (c->step_embed_dim % 2) == 0,
            LDP_ERR_INVALID_ARG, "bad dims");
  LDP_CHECK(c->n_levels >= 1 && c->n_levels <= 6, LDP_ERR_INVALID_ARG, "n_levels must be 1..6");
  LDP_CHECK(c->kernel_size == 5, LDP_ERR_UNSUPPORTED, "kernel_size must be 5 (reference agent/ldp_agent.yaml:13)");
  LDP_CHECK(c->n_groups > 0 && c->n_train_steps > 0, LDP_ERR_INVALID_ARG, "bad n_groups / n_train_steps");
  for (int i = 0; i < c->n_levels; ++i)
    LDP_CHECK(c->down_dims[i] > 0 && c->down_dims[i] % c->n_groups == 0 && c->down_dims[i] % 8 == 0, LDP_ERR_INVALID_ARG,
              "down_dims must be positive multiples of n_groups and 8");
  return LDP_OK;
}

// ------------------------------- create -------------------------------------------------------
static int planner_build_tables(LdpPlanner* h) {
  const LdpUnetConfig& c = h->cfg;
  cudaStream_t s = 0;
  const int ds = c.step_embed_dim, n = c.n_train_steps, dc = c.global_cond_dim;
  float *sinus, *hid, *temb, *wt_all, *bias_all;
  Arena tmp;
  LDP_TRY(tmp.alloc_t(&sinus, (size_t)n * ds));
  LDP_TRY(tmp.alloc_t(&hid, (size_t)n * ds * 4));
  LDP_TRY(tmp.alloc_t(&temb, (size_t)n * ds));
  LDP_TRY(tmp.alloc_t(&wt_all, (size_t)ds * h->sum_c2));
  LDP_TRY(tmp.alloc_t(&bias_all, (size_t)h->sum_c2));
  LDP_TRY(h->arena.alloc_t(&h->wc_all, (size_t)dc * h->sum_c2));
  LDP_TRY(h->arena.alloc_t(&h->ttab, (size_t)n * h->sum_c2));
  // diffusion_step_encoder: sinusoid -> Dense(4d) -> Mish -> Dense(d)   (diffusion_nets_v2.py:120-127)
  LDP_TRY(launch_sinusoid_table(sinus, n, ds, /*cos_first=*/0, s));
  GemmF32 g;
  g.x1 = sinus; g.c1 = ds; g.ld1 = ds; g.w = h->t0w; g.ldw = ds * 4; g.bias = h->t0b; g.out = hid; g.ldo = ds * 4;
  g.m = n; g.n = ds * 4;
  LDP_TRY(launch_gemm_f32(g, s));
  g = GemmF32();
  g.x1 = hid; g.c1 = ds * 4; g.ld1 = ds * 4; g.a_act = 1; g.w = h->t1w; g.ldw = ds; g.bias = h->t1b; g.out = temb;
  g.ldo = ds; g.m = n; g.n = ds;
  LDP_TRY(launch_gemm_f32(g, s));
  // split every FiLM Dense (K = ds + Dc) into its time rows and observation rows, side by side over all blocks
  for (auto& b : h->crb) {
    const int n2 = 2 * b.cout;
    LDP_CUDA_OK(cudaMemcpy2DAsync(wt_all + b.film_off, (size_t)h->sum_c2 * 4, b.fw, (size_t)n2 * 4, (size_t)n2 * 4, ds,
                                  cudaMemcpyDeviceToDevice, s));
    LDP_CUDA_OK(cudaMemcpy2DAsync(h->wc_all + b.film_off, (size_t)h->sum_c2 * 4, b.fw + (size_t)ds * n2, (size_t)n2 * 4,
                                  (size_t)n2 * 4, dc, cudaMemcpyDeviceToDevice, s));
    LDP_CUDA_OK(cudaMemcpyAsync(bias_all + b.film_off, b.fb, (size_t)n2 * 4, cudaMemcpyDeviceToDevice, s));
  }
  // ttab[k] = Mish(temb_k) Wt + bias   (Mish is elementwise, so Dense(Mish([temb|cond])) separates exactly)
  g = GemmF32();
  g.x1 = temb; g.c1 = ds; g.ld1 = ds; g.a_act = 1; g.w = wt_all; g.ldw = h->sum_c2; g.bias = bias_all; g.out = h->ttab;
  g.ldo = h->sum_c2; g.m = n; g.n = h->sum_c2;
  LDP_TRY(launch_gemm_f32(g, s));
  std::vector<float> coef;
  ddpm_coef_host(n, coef);
  LDP_TRY(h->arena.alloc_t(&h->coef, coef.size()));
  LDP_CUDA_OK(cudaMemcpy(h->coef, coef.data(), coef.size() * 4, cudaMemcpyHostToDevice));
  for (size_t i = 0; i < h->up_w.size(); ++i) {
    const int d = c.down_dims[c.n_levels - 2 - (int)i];
    float* b2;
    LDP_TRY(h->arena.alloc_t(&b2, (size_t)2 * d));
    LDP_CUDA_OK(cudaMemcpy(b2, h->up_b[i], (size_t)d * 4, cudaMemcpyDeviceToDevice));
    LDP_CUDA_OK(cudaMemcpy(b2 + d, h->up_b[i], (size_t)d * 4, cudaMemcpyDeviceToDevice));
    h->up_bias2.push_back(b2);
  }
  LDP_CUDA_OK(cudaStreamSynchronize(s));
  return LDP_OK;
}

static int planner_create_impl(const LdpUnetConfig* cfg, const float* params_host, uint64_t n_params, LdpPlanner* h) {
  h->cfg = *cfg;
  const LdpUnetConfig& c = h->cfg;
  const int64_t expect = unet_param_count(c);
  LDP_CHECK((int64_t)n_params == expect, LDP_ERR_PARAM_COUNT,
            "planner weight blob has " + std::to_string(n_params) + " floats, config needs " + std::to_string(expect));
  LDP_TRY(h->arena.alloc_t(&h->blob, n_params, false));
  LDP_CUDA_OK(cudaMemcpy(h->blob, params_host, n_params * 4, cudaMemcpyHostToDevice));
  BlobWalker w{h->blob, 0};
  const int64_t ds = c.step_embed_dim, cond = ds + c.global_cond_dim, k = c.kernel_size;
  h->t0w = w.take(ds * ds * 4); h->t0b = w.take(ds * 4);
  h->t1w = w.take(ds * 4 * ds); h->t1b = w.take(ds);
  int film_off = 0;
  for (auto& [cin, cout, proj] : block_plan(c)) {
    CrbW b;
    b.cin = cin; b.cout = cout; b.proj = proj; b.film_off = film_off;
    film_off += 2 * cout;
    b.c1w = w.take(k * cin * cout); b.c1b = w.take(cout); b.g1s = w.take(cout); b.g1b = w.take(cout);
    b.fw = w.take(cond * 2 * cout); b.fb = w.take(2 * cout);
    b.c2w = w.take(k * (int64_t)cout * cout); b.c2b = w.take(cout); b.g2s = w.take(cout); b.g2b = w.take(cout);
    b.rw = proj ? w.take((int64_t)cin * cout) : nullptr;
    b.rb = proj ? w.take(cout) : nullptr;
    h->crb.push_back(b);
  }
  h->sum_c2 = film_off;
  LDP_CHECK(h->crb.size() <= 16, LDP_ERR_UNSUPPORTED, "at most 16 residual blocks (FilmBlocks table)");
  for (int i = 0; i < c.n_levels - 1; ++i) {
    int64_t d = c.down_dims[i];
    h->down_w.push_back(w.take(3 * d * d));
    h->down_b.push_back(w.take(d));
  }
  for (int i = 0; i < c.n_levels - 1; ++i) {
    int64_t d = c.down_dims[c.n_levels - 2 - i];
    h->up_w.push_back(w.take(4 * d * d));
    h->up_b.push_back(w.take(d));
  }
  const int64_t d0 = c.down_dims[0];
  h->fcw = w.take(k * d0 * d0); h->fcb = w.take(d0); h->fgs = w.take(d0); h->fgb = w.take(d0);
  h->ow = w.take(d0 * c.input_dim); h->ob = w.take(c.input_dim);
  LDP_CHECK((int64_t)w.pos == expect, LDP_ERR_PARAM_COUNT, "internal: blob walk mismatch");
  const char* env = getenv("LDP_NO_GRAPH");
  h->use_graph = !(env && env[0] == '1');
  return planner_build_tables(h);
}

static bool env_off(const char* name);

// ------------------------------- workspace -------------------------------------------------------
static int level_len(int T, int level) { return T >> level; }

static int get_ws(LdpPlanner* h, int B, int T, PlanWs** out) {
  const LdpUnetConfig& c = h->cfg;
  LDP_CHECK(B > 0 && T > 0, LDP_ERR_BAD_SHAPE, "B and T must be positive");
  LDP_CHECK(T % (1 << (c.n_levels - 1)) == 0, LDP_ERR_BAD_SHAPE,
            "T must be divisible by 2^(n_levels-1): the reference UNet's up path otherwise mismatches its skips");
  auto key = std::make_pair(B, T);
  auto it = h->ws.find(key);
  if (it != h->ws.end()) {
    it->second->last_use = ++h->use_clock;
    *out = it->second.get();
    return LDP_OK;
  }
  ws_evict_lru(h->ws);
  std::unique_ptr<PlanWs> w(new PlanWs());
  w->B = B; w->T = T; w->last_use = ++h->use_clock;
  LDP_TRY(w->arena.alloc_t(&w->otab, (size_t)B * h->sum_c2));
  LDP_TRY(w->arena.alloc_t(&w->otab_q, (size_t)B * h->sum_c2));
  LDP_TRY(w->arena.alloc_t(&w->x_state, (size_t)B * T * c.input_dim));
  LDP_TRY(w->arena.alloc_t(&w->eps_buf, (size_t)B * T * c.input_dim));
  LDP_TRY(w->arena.alloc_t(&w->step_dev, 4));
  LDP_TRY(w->arena.alloc_t(&w->done_counter, 4));
  LDP_TRY(w->arena.alloc_t(&w->call_dev, 1));
  *out = w.get();
  h->ws[key] = std::move(w);
  return LDP_OK;
}

static FilmBlocks film_blocks(const LdpPlanner* h) {
  FilmBlocks fb;
  fb.n = (int)h->crb.size();
  for (int i = 0; i < fb.n; ++i) {
    fb.film_off[i] = h->crb[i].film_off;
    fb.pair_off[i] = h->crb[i].film_off / 2;
    fb.C[i] = h->crb[i].cout;
  }
  return fb;
}

// Observation part of every block's FiLM Dense, once per act(): otab[b] = Mish(cond[b]) Wc  (B x Dc x sum_c2).
// fp32 path: SIMT FFMA GEMM (the 1e-5 parity instrument).  bf16 path: one tcgen05 GEMM (persistent tile loop, fp32
// accumulation and fp32 output) - on SIMT this contraction alone was 1 ms of a 41 ms sampling loop.
static int compute_otab(LdpPlanner* h, PlanWs* w, const float* cond, int precision, cudaStream_t s) {
  const int dc = h->cfg.global_cond_dim;
  if (precision == LDP_PREC_FP32 || env_off("LDP_OTAB_TC")) {
    GemmF32 g;
    g.x1 = cond; g.c1 = dc; g.ld1 = dc; g.a_act = 1;
    g.w = h->wc_all; g.ldw = h->sum_c2; g.out = w->otab; g.ldo = h->sum_c2; g.m = w->B; g.n = h->sum_c2;
    LDP_TRY(launch_gemm_f32(g, s));
    return launch_film_pack(w->otab, h->sum_c2, w->otab_q, w->B, film_blocks(h), s);
  }
  if (!h->otab_packed) {
    PackedW& pw = h->pw_otab;
    pw.kp = round_up(dc, 64);
    pw.n_pad = round_up(h->sum_c2, 128);
    LDP_TRY(h->arena.alloc_t(&pw.wt, (size_t)pw.n_pad * pw.kp));
    std::vector<int32_t> kmap(pw.kp);
    std::vector<TcStage> st(pw.kp / 64);
    for (int k = 0; k < pw.kp; ++k) kmap[k] = k < dc ? k : -1;
    for (int i = 0; i < pw.kp / 64; ++i) st[i] = make_stage(0, 0, 1, i * 64, 0, 0, i);
    Arena tmp;
    int32_t* map_dev;
    LDP_TRY(tmp.alloc_t(&map_dev, pw.kp));
    LDP_CUDA_OK(cudaMemcpy(map_dev, kmap.data(), (size_t)pw.kp * 4, cudaMemcpyHostToDevice));
    LDP_TRY(upload_stage_table(h->arena, st, &pw));
    LDP_TRY(launch_pack_wt_bf16(h->wc_all, h->sum_c2, h->sum_c2, map_dev, pw.kp, pw.wt, pw.kp, 0, pw.n_pad, 0));
    LDP_CUDA_OK(cudaDeviceSynchronize());
    h->otab_packed = true;
  }
  if (!w->otab_ready) {
    LDP_TRY(tc_driver_check());
    LDP_TRY(tc_gemm_init());
    const PackedW& pw = h->pw_otab;
    w->ld_cb = round_up(dc, 8);
    LDP_TRY(w->arena.alloc_t(&w->cond_bf16, (size_t)w->B * w->ld_cb));
    TcGemm& op = w->otab_op;
    op = TcGemm();
    uint64_t ad[4] = {(uint64_t)dc, 1, 1, (uint64_t)w->B};
    uint64_t as[3] = {(uint64_t)w->ld_cb * 2, (uint64_t)w->ld_cb * 2, (uint64_t)w->ld_cb * 2};
    uint32_t ab[4] = {64, 1, 1, 128};
    LDP_TRY(make_tmap_bf16(&op.map_a[0], w->cond_bf16, 4, ad, as, ab));
    for (int i = 1; i < 4; ++i) op.map_a[i] = op.map_a[0];
    uint64_t bd[2] = {(uint64_t)pw.kp, (uint64_t)pw.n_pad};
    uint64_t bs[1] = {(uint64_t)pw.kp * 2};
    uint32_t bb[2] = {64, 128};
    LDP_TRY(make_tmap_bf16(&op.map_b, pw.wt, 2, bd, bs, bb));
    op.kb = pw.kb_dev; op.num_kb = pw.num_kb; op.runs = pw.runs_dev; op.num_runs = pw.num_runs;
    tc_set_inline_runs(&op, pw.runs_host.data(), pw.num_runs);
    op.M = w->B; op.N = h->sum_c2; op.block_n = 128;
    op.items_per_tile = 128; op.rows_per_item = 1;
    op.mode = TC_EPI_PLAIN;
    op.out_f32 = w->otab; op.ld_out_f32 = h->sum_c2;
    w->otab_ready = true;
  }
  LDP_TRY(launch_cast_bf16(cond, dc, w->cond_bf16, w->ld_cb, w->B, dc, /*mish=*/1, s));
  LDP_TRY(launch_tc_gemm(w->otab_op, s));
  return launch_film_pack(w->otab, h->sum_c2, w->otab_q, w->B, film_blocks(h), s);
}

// ------------------------------- fp32 program -------------------------------------------------------
static int prepare_f32(LdpPlanner* h, PlanWs* w) {
  if (w->f32_ready) return LDP_OK;
  const LdpUnetConfig& c = h->cfg;
  size_t max_act = 0;
  {
    int lvl = 0;
    for (size_t i = 0; i < h->crb.size(); ++i) {
      // level of block i: down blocks 2l,2l+1 at level l; mid at last; up blocks mirror
      int nl = c.n_levels;
      if ((int)i < 2 * nl) lvl = (int)i / 2;
      else if ((int)i < 2 * nl + 2) lvl = nl - 1;
      else lvl = nl - 1 - ((int)i - 2 * nl - 2) / 2;
      size_t a = (size_t)w->B * level_len(w->T, lvl) * std::max(h->crb[i].cout, h->crb[i].cin);
      max_act = std::max(max_act, a);
      float* o;
      LDP_TRY(w->arena.alloc_t(&o, (size_t)w->B * level_len(w->T, lvl) * h->crb[i].cout));
      w->f_out.push_back(o);
    }
  }
  max_act = std::max(max_act, (size_t)w->B * w->T * c.down_dims[0]);
  LDP_TRY(w->arena.alloc_t(&w->f_tmp, max_act));
  LDP_TRY(w->arena.alloc_t(&w->f_h1, max_act));
  LDP_TRY(w->arena.alloc_t(&w->f_res, max_act));
  LDP_TRY(w->arena.alloc_t(&w->f_final, (size_t)w->B * w->T * c.down_dims[0]));
  for (int i = 0; i < c.n_levels - 1; ++i) {
    float* d;
    LDP_TRY(w->arena.alloc_t(&d, (size_t)w->B * level_len(w->T, i + 1) * c.down_dims[i]));
    w->f_down.push_back(d);
  }
  for (int i = 0; i < c.n_levels - 1; ++i) {
    int lvl = c.n_levels - 1 - i;               // input level of Upsample1d_i
    float* u;
    LDP_TRY(w->arena.alloc_t(&u, (size_t)w->B * level_len(w->T, lvl - 1) * c.down_dims[lvl - 1]));
    w->f_up.push_back(u);
  }
  w->f32_ready = true;
  return LDP_OK;
}

struct SrcF32 { const float* p; int c; int ld; };

static int conv_f32(const SrcF32& a, const SrcF32* b, int B, int t_in, int t_out, int taps, int stride, int pad, int dil,
                    const float* wgt, const float* bias, int cout, float* out, cudaStream_t s) {
  GemmF32 g;
  g.x1 = a.p; g.c1 = a.c; g.ld1 = a.ld;
  if (b) { g.x2 = b->p; g.c2 = b->c; g.ld2 = b->ld; }
  g.t_in = t_in; g.t_out = t_out; g.taps = taps; g.stride = stride; g.pad = pad; g.dil = dil;
  g.w = wgt; g.ldw = cout; g.bias = bias; g.out = out; g.ldo = cout; g.m = B * t_out; g.n = cout;
  return launch_gemm_f32(g, s);
}

static int crb_f32(LdpPlanner* h, PlanWs* w, int bi, const SrcF32& a, const SrcF32* b2, int Tl, StepRef step, float* out,
                   cudaStream_t s) {
  const CrbW& b = h->crb[bi];
  const int B = w->B, G = h->cfg.n_groups;
  LDP_TRY(conv_f32(a, b2, B, Tl, Tl, 5, 1, 2, 1, b.c1w, b.c1b, b.cout, w->f_tmp, s));
  GroupNormF32 n;
  n.x = w->f_tmp; n.ldx = b.cout; n.y = w->f_h1; n.ldy = b.cout; n.B = B; n.P = Tl; n.C = b.cout; n.G = G;
  n.gamma = b.g1s; n.beta = b.g1b; n.act = 1; n.film = 1; n.ttab = h->ttab; n.ld_ttab = h->sum_c2; n.step = step;
  n.otab = w->otab; n.ld_otab = h->sum_c2; n.film_off = b.film_off;
  LDP_TRY(launch_groupnorm_f32(n, s));
  SrcF32 h1{w->f_h1, b.cout, b.cout};
  LDP_TRY(conv_f32(h1, nullptr, B, Tl, Tl, 5, 1, 2, 1, b.c2w, b.c2b, b.cout, w->f_tmp, s));
  const float* res = a.p;
  int ldres = a.ld;
  if (b.proj) {
    LDP_TRY(conv_f32(a, b2, B, Tl, Tl, 1, 1, 0, 1, b.rw, b.rb, b.cout, w->f_res, s));
    res = w->f_res;
    ldres = b.cout;
  }
  n = GroupNormF32();
  n.x = w->f_tmp; n.ldx = b.cout; n.y = out; n.ldy = b.cout; n.B = B; n.P = Tl; n.C = b.cout; n.G = G;
  n.gamma = b.g2s; n.beta = b.g2b; n.act = 1; n.res = res; n.ldres = ldres;
  return launch_groupnorm_f32(n, s);
}

// eps = UNet(x, step, cond);  otab must already hold the observation part of the FiLM embeddings.
static int forward_f32(LdpPlanner* h, PlanWs* w, const float* x, StepRef step, float* eps, cudaStream_t s) {
  const LdpUnetConfig& c = h->cfg;
  LDP_TRY(prepare_f32(h, w));
  const int B = w->B, nl = c.n_levels;
  SrcF32 cur{x, c.input_dim, c.input_dim};
  int bi = 0;
  std::vector<SrcF32> skips;
  for (int l = 0; l < nl; ++l) {
    const int Tl = level_len(w->T, l);
    LDP_TRY(crb_f32(h, w, bi, cur, nullptr, Tl, step, w->f_out[bi], s));
    cur = SrcF32{w->f_out[bi], h->crb[bi].cout, h->crb[bi].cout}; ++bi;
    LDP_TRY(crb_f32(h, w, bi, cur, nullptr, Tl, step, w->f_out[bi], s));
    cur = SrcF32{w->f_out[bi], h->crb[bi].cout, h->crb[bi].cout}; ++bi;
    skips.push_back(cur);
    if (l < nl - 1) {
      // Downsample1d: Conv(k3, stride 2, 'SAME') -> pad_lo = 0 for even lengths
      const int d = c.down_dims[l];
      LDP_TRY(conv_f32(cur, nullptr, B, Tl, Tl / 2, 3, 2, 0, 1, h->down_w[l], h->down_b[l], d, w->f_down[l], s));
      cur = SrcF32{w->f_down[l], d, d};
    }
  }
  {
    const int Tl = level_len(w->T, nl - 1);
    for (int r = 0; r < 2; ++r) {
      LDP_TRY(crb_f32(h, w, bi, cur, nullptr, Tl, step, w->f_out[bi], s));
      cur = SrcF32{w->f_out[bi], h->crb[bi].cout, h->crb[bi].cout}; ++bi;
    }
  }
  for (int u = 0; u < nl - 1; ++u) {
    const int lvl = nl - 1 - u;
    const int Tl = level_len(w->T, lvl);
    SrcF32 skip = skips.back();
    skips.pop_back();
    LDP_TRY(crb_f32(h, w, bi, cur, &skip, Tl, step, w->f_out[bi], s));
    cur = SrcF32{w->f_out[bi], h->crb[bi].cout, h->crb[bi].cout}; ++bi;
    LDP_TRY(crb_f32(h, w, bi, cur, nullptr, Tl, step, w->f_out[bi], s));
    cur = SrcF32{w->f_out[bi], h->crb[bi].cout, h->crb[bi].cout}; ++bi;
    // Upsample1d: Flax ConvTranspose(k4, s2, 'SAME'), kernel not flipped
    const int d = c.down_dims[lvl - 1];
    LDP_TRY(conv_f32(cur, nullptr, B, Tl, 2 * Tl, 4, 1, 2, 2, h->up_w[u], h->up_b[u], d, w->f_up[u], s));
    cur = SrcF32{w->f_up[u], d, d};
  }
  const int d0 = c.down_dims[0];
  LDP_TRY(conv_f32(cur, nullptr, B, w->T, w->T, 5, 1, 2, 1, h->fcw, h->fcb, d0, w->f_tmp, s));
  GroupNormF32 n;
  n.x = w->f_tmp; n.ldx = d0; n.y = w->f_final; n.ldy = d0; n.B = B; n.P = w->T; n.C = d0; n.G = 8;
  n.gamma = h->fgs; n.beta = h->fgb; n.act = 1;
  LDP_CHECK(d0 % 8 == 0, LDP_ERR_UNSUPPORTED, "final Conv1dBlock uses 8 groups");
  LDP_TRY(launch_groupnorm_f32(n, s));
  SrcF32 f{w->f_final, d0, d0};
  return conv_f32(f, nullptr, B, w->T, w->T, 1, 1, 0, 1, h->ow, h->ob, c.input_dim, eps, s);
}

// ------------------------------- bf16 / tcgen05 program -------------------------------------------------------
// A convolution on the tcgen05 path.  Two K layouts:
//  * tap-accumulator (k-tap conv, stride 1): K = (source, 64-channel block, tap).  One pipeline stage loads the
//    un-shifted A tile of a channel block once plus the W tiles of all taps; tap j accumulates into its own TMEM
//    accumulator and the epilogue recombines out[t] = sum_j acc_j[t + j - pad] with warp shuffles.  Against the
//    per-tap layout this cuts the A traffic (the larger share of what an SM pulls from L2) by the tap count.
//  * per-tap (strided / transposed convs, or when k accumulators do not fit the 512 TMEM columns): K = (tap, source,
//    channel block); every tap re-fetches its shifted A tile through TMA (zero padding = out-of-bounds fill).
// An optional aux part appends the K blocks of a 1x1 convolution on other sources (the residual projection of a
// ConditionalResidualBlock1D) that accumulate into one more TMEM accumulator.
struct ConvDesc {
  int kind = CONV_K, taps_k = 1;
  const ActBf16* srcs = nullptr; int nsrc = 0;
  int t_in = 1;
  const float* wgt = nullptr; int cout = 0;
  const ActBf16* aux_srcs = nullptr; int aux_nsrc = 0; const float* aux_w = nullptr;   // 1x1 on the block input
  int group_width = 1 << 30;      // GroupNorm group width the N tile must contain (1<<30: no constraint)
  bool no_pair = false;
};

static bool env_off(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '0';
}

// N-tile width and K layout.  64-wide tiles double the CTA count of layers whose 128-wide grid would fill under half
// of the 148 SMs; GroupNorm groups must fit inside one tile; (accumulators x BN) must fit the 512 TMEM columns.
static void choose_tiling(int M, int N, const ConvDesc& d, int n_taps, int* bn_out, bool* tapacc_out) {
  const int aux = d.aux_nsrc > 0 ? 1 : 0;
  const int ctas128 = ceil_div(M, 128) * ceil_div(N, 128);
  const bool ok64 = N % 64 == 0 && d.group_width <= 64 && !env_off("LDP_BN64");
  const bool ok128 = d.group_width <= 128 || d.group_width == (1 << 30);
  bool tapacc = d.kind == CONV_K && n_taps >= 2 && !env_off("LDP_TAPACC");
  int bn = 128;
  if (tapacc) {
    const bool t128 = ok128 && (n_taps + aux) * 128 <= 512;
    const bool t64 = ok64 && (n_taps + aux) * 64 <= 512;
    if (t128 && (!t64 || ctas128 > 74)) bn = 128;
    else if (t64) bn = 64;
    else tapacc = false;
    // a 64-wide tap-accumulator grid that spills into a second wave loses to the one-wave per-tap grid (measured)
    if (tapacc && ceil_div(M, 128) * ceil_div(N, bn) > 148 && ctas128 <= 148 && ok128) tapacc = false;
  }
  if (!tapacc) bn = (ok64 && ctas128 <= 74) ? 64 : 128;
  *bn_out = bn;
  *tapacc_out = tapacc;
}

// Build (or fetch) the packed weights + stage table of one convolution at input length t_in.
static int get_packed(LdpPlanner* h, int op_id, const ConvDesc& d, bool tapacc, ConvPack** out) {
  auto key = std::make_pair(op_id * 2 + (tapacc ? 1 : 0), d.t_in);
  auto it = h->packed.find(key);
  if (it != h->packed.end()) {
    *out = &it->second;
    return LDP_OK;
  }
  int ctot = 0;
  for (int i = 0; i < d.nsrc; ++i) ctot += d.srcs[i].c;
  ConvPack pw;
  std::vector<TcStage> st;
  std::vector<int32_t> kmap, kmap2;     // kmap2: odd output phase of the transposed conv
  auto push64 = [&](int src_c, int c0, int row_base, int row_base2) {
    for (int i = 0; i < 64; ++i) {
      const bool ok = c0 + i < src_c;
      kmap.push_back(ok && row_base >= 0 ? row_base + c0 + i : -1);
      kmap2.push_back(ok && row_base2 >= 0 ? row_base2 + c0 + i : -1);
    }
  };
  if (d.kind == CONV_K && tapacc) {
    const int pad = d.taps_k / 2;
    std::vector<int> taps;
    for (int j = 0; j < d.taps_k; ++j)
      if (std::abs(j - pad) < d.t_in) taps.push_back(j);        // taps that only ever see padding are dropped
    pw.n_acc = (int)taps.size();
    pw.tapacc = true;
    for (size_t a = 0; a < taps.size(); ++a) pw.shift[a] = taps[a] - pad;
    pw.w_max = pw.n_acc;
    int coff = 0;
    for (int sidx = 0; sidx < d.nsrc; ++sidx) {
      for (int c0 = 0; c0 < d.srcs[sidx].c; c0 += 64) {
        st.push_back(make_stage(sidx, 0, pw.n_acc, c0, 0, 0, (int)kmap.size() / 64));
        for (int j : taps) push64(d.srcs[sidx].c, c0, j * ctot + coff, -1);
      }
      coff += d.srcs[sidx].c;
    }
  } else {
    auto push_sources = [&](int d1, int d2, int tap_row, int tap_row2) {
      int coff = 0;
      for (int sidx = 0; sidx < d.nsrc; ++sidx) {
        for (int c0 = 0; c0 < d.srcs[sidx].c; c0 += 64) {
          st.push_back(make_stage(sidx, 0, 1, c0, d1, d2, (int)kmap.size() / 64));
          push64(d.srcs[sidx].c, c0, tap_row >= 0 ? tap_row * ctot + coff : -1, tap_row2 >= 0 ? tap_row2 * ctot + coff : -1);
        }
        coff += d.srcs[sidx].c;
      }
    };
    if (d.kind == CONV_K) {
      const int pad = d.taps_k / 2;
      for (int j = 0; j < d.taps_k; ++j)
        if (std::abs(j - pad) < d.t_in) push_sources(j - pad, 0, j, -1);
    } else if (d.kind == CONV_DOWN) {
      // y[t] = sum_j W[j] x[2t+j] (pad_lo = 0 for even t_in); x viewed as (parity, t/2): tap j -> (j%2, t + j/2)
      const int t_out = d.t_in / 2;
      for (int j = 0; j < 3; ++j)
        if (j < 2 || t_out >= 2) push_sources(j % 2, j / 2, j, -1);
    } else {
      // y[2u] = W0 x[u-1] + W2 x[u];  y[2u+1] = W1 x[u] + W3 x[u+1]
      if (d.t_in >= 2) push_sources(-1, 0, 0, -1);
      push_sources(0, 0, 2, 1);
      if (d.t_in >= 2) push_sources(1, 0, -1, 3);
    }
  }
  const int kp_main = (int)kmap.size();
  pw.num_kb_main = (int)st.size();
  std::vector<int32_t> kmap_aux;
  {
    int coff = 0;
    for (int sidx = 0; sidx < d.aux_nsrc; ++sidx) {
      for (int c0 = 0; c0 < d.aux_srcs[sidx].c; c0 += 64) {
        st.push_back(make_stage(d.nsrc + sidx, pw.n_acc, 1, c0, 0, 0, (kp_main + (int)kmap_aux.size()) / 64));
        for (int i = 0; i < 64; ++i) kmap_aux.push_back(c0 + i < d.aux_srcs[sidx].c ? coff + c0 + i : -1);
      }
      coff += d.aux_srcs[sidx].c;
    }
  }
  LDP_CHECK(d.nsrc + d.aux_nsrc <= 4, LDP_ERR_UNSUPPORTED, "a convolution reads at most 4 activation tensors");
  const int n_out = d.kind == CONV_UP ? 2 * d.cout : d.cout;
  pw.kp = kp_main + (int)kmap_aux.size();
  pw.n_pad = round_up(n_out, 128);
  LDP_TRY(h->arena.alloc_t(&pw.wt, (size_t)pw.n_pad * pw.kp));
  LDP_TRY(upload_stage_table(h->arena, st, &pw));
  Arena tmp;
  int32_t* map_dev;
  LDP_TRY(tmp.alloc_t(&map_dev, (size_t)kp_main * 2 + kmap_aux.size() + 16));
  LDP_CUDA_OK(cudaMemcpy(map_dev, kmap.data(), (size_t)kp_main * 4, cudaMemcpyHostToDevice));
  if (d.kind == CONV_UP) {
    LDP_CUDA_OK(cudaMemcpy(map_dev + kp_main, kmap2.data(), (size_t)kp_main * 4, cudaMemcpyHostToDevice));
    LDP_TRY(launch_pack_wt_bf16(d.wgt, d.cout, d.cout, map_dev, kp_main, pw.wt, pw.kp, 0, d.cout, 0));
    LDP_TRY(launch_pack_wt_bf16(d.wgt, d.cout, d.cout, map_dev + kp_main, kp_main, pw.wt + (size_t)d.cout * pw.kp, pw.kp, 0,
                                pw.n_pad - d.cout, 0));
  } else {
    LDP_TRY(launch_pack_wt_bf16(d.wgt, d.cout, d.cout, map_dev, kp_main, pw.wt, pw.kp, 0, pw.n_pad, 0));
    if (!kmap_aux.empty()) {
      LDP_CUDA_OK(cudaMemcpy(map_dev + 2 * kp_main, kmap_aux.data(), kmap_aux.size() * 4, cudaMemcpyHostToDevice));
      LDP_TRY(launch_pack_wt_bf16(d.aux_w, d.cout, d.cout, map_dev + 2 * kp_main, (int)kmap_aux.size(), pw.wt, pw.kp, kp_main,
                                  pw.n_pad, 0));
    }
  }
  LDP_CUDA_OK(cudaDeviceSynchronize());
  h->packed[key] = pw;
  *out = &h->packed[key];
  return LDP_OK;
}

static int act_map(CUtensorMap* m, const ActBf16& a, int kind, int t_in, int B) {
  if (kind == CONV_DOWN) {
    const int t_out = t_in / 2;
    uint64_t dims[4] = {(uint64_t)a.c, 2, (uint64_t)t_out, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)a.ld * 2, (uint64_t)a.ld * 4, (uint64_t)t_in * a.ld * 2};
    uint32_t box[4] = {64, 1, (uint32_t)t_out, (uint32_t)(128 / t_out)};
    return make_tmap_bf16(m, a.p, 4, dims, str, box);
  }
  uint64_t dims[4] = {(uint64_t)a.c, (uint64_t)t_in, 1, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)a.ld * 2, (uint64_t)t_in * a.ld * 2, (uint64_t)t_in * a.ld * 2};
  uint32_t box[4] = {64, (uint32_t)t_in, 1, (uint32_t)(128 / t_in)};
  return make_tmap_bf16(m, a.p, 4, dims, str, box);
}

// Generic convolution op on the tcgen05 path; the caller fills in the epilogue afterwards.
static int conv_tc(LdpPlanner* h, PlanWs* w, int op_id, const ConvDesc& d, TcGemm* op) {
  const int rows = d.kind == CONV_DOWN ? d.t_in / 2 : d.t_in;
  LDP_CHECK(rows >= 1 && rows <= 32 && (rows & (rows - 1)) == 0, LDP_ERR_UNSUPPORTED,
            "bf16 path needs power-of-two level lengths <= 32 (use LDP_PREC_FP32 for other horizons)");
  for (int i = 0; i < d.nsrc; ++i)
    LDP_CHECK(d.srcs[i].ld % 8 == 0, LDP_ERR_INVALID_ARG, "activation pitch must be a multiple of 8 elements");
  const int M = w->B * rows, N = d.kind == CONV_UP ? 2 * d.cout : d.cout;
  int n_taps = 0;
  if (d.kind == CONV_K)
    for (int j = 0; j < d.taps_k; ++j) n_taps += std::abs(j - d.taps_k / 2) < d.t_in ? 1 : 0;
  int bn;
  bool tapacc;
  choose_tiling(M, N, d, n_taps, &bn, &tapacc);
  ConvPack* pw;
  LDP_TRY(get_packed(h, op_id, d, tapacc, &pw));
  *op = TcGemm();
  for (int i = 0; i < d.nsrc; ++i) LDP_TRY(act_map(&op->map_a[i], d.srcs[i], d.kind, d.t_in, w->B));
  for (int i = 0; i < d.aux_nsrc; ++i) LDP_TRY(act_map(&op->map_a[d.nsrc + i], d.aux_srcs[i], CONV_K, d.t_in, w->B));
  for (int i = d.nsrc + d.aux_nsrc; i < 4; ++i) op->map_a[i] = op->map_a[0];
  uint64_t bd[2] = {(uint64_t)pw->kp, (uint64_t)pw->n_pad};
  uint64_t bs[1] = {(uint64_t)pw->kp * 2};
  // CTA pairs (cta_group::2): each CTA of a pair fetches half of every W tile.  Measured: the tap-accumulator layout
  // (W is the larger share of a stage) gains ~4k cycles of main loop per layer; the per-tap layout gains nothing and
  // pays ~0.9 us of cluster launch + cluster barriers, so it stays single-CTA (LDP_PAIR=2 forces pairs everywhere).
  const char* pe = getenv("LDP_PAIR");
  const bool pair = !d.no_pair && !(pe && pe[0] == '0') && (tapacc || (pe && pe[0] == '2'));
  uint32_t bb[2] = {64, (uint32_t)(pair ? bn / 2 : bn)};
  LDP_TRY(make_tmap_bf16(&op->map_b, pw->wt, 2, bd, bs, bb));
  {
    uint32_t bt[2] = {64, 16};                       // extra W rows of a widened last N tile (DDPM epilogue; see TcGemm)
    LDP_TRY(make_tmap_bf16(&op->map_b_tail, pw->wt, 2, bd, bs, bt));
  }
  op->pair = pair ? 1 : 0;
  op->kb = pw->kb_dev;
  op->num_kb = pw->num_kb;
  op->runs = pw->runs_dev;
  op->num_runs = pw->num_runs;
  tc_set_inline_runs(op, pw->runs_host.data(), pw->num_runs);
  op->w_max = pw->w_max;
  op->kb_main = pw->num_kb_main;
  op->nw_main = pw->w_max;
  op->k_pad = pw->kp;
  op->wt_host_ref = pw->wt;
  op->n_pad = pw->n_pad;
  op->n_acc = pw->n_acc;
  for (int j = 0; j < 5; ++j) op->shift[j] = pw->shift[j];
  op->use_aux = d.aux_nsrc > 0 ? 1 : 0;
  op->M = M;
  op->N = N;
  op->block_n = bn;
  op->tiles_per_item = 1;
  op->rows_step = 0;
  op->items_per_tile = 128 / rows;
  op->rows_per_item = rows;
  return LDP_OK;
}

static void set_gn(TcGemm* op, const float* bias, const float* gamma, const float* beta, int cout, int groups,
                   __nv_bfloat16* out) {
  op->mode = TC_EPI_GN;
  op->bias = bias; op->gamma = gamma; op->beta = beta; op->eps = 1e-6f;
  op->group_width = cout / groups;
  op->out_bf16 = out; op->ld_out_bf16 = cout;
}

static int crb_tc(LdpPlanner* h, PlanWs* w, int bi, const ActBf16* srcs, int nsrc, int Tl, __nv_bfloat16* h1buf,
                  __nv_bfloat16* outbuf) {
  const CrbW& b = h->crb[bi];
  LDP_CHECK((b.cout / h->cfg.n_groups) % 32 == 0, LDP_ERR_UNSUPPORTED,
            "bf16 path needs GroupNorm group widths that are multiples of 32 channels");
  const int gw = b.cout / h->cfg.n_groups;
  TcGemm op;
  ConvDesc d;
  d.kind = CONV_K; d.taps_k = 5; d.srcs = srcs; d.nsrc = nsrc; d.t_in = Tl; d.wgt = b.c1w; d.cout = b.cout; d.group_width = gw;
  LDP_TRY(conv_tc(h, w, 4 * bi + 0, d, &op));
  set_gn(&op, b.c1b, b.g1s, b.g1b, b.cout, h->cfg.n_groups, h1buf);
  op.film = 1; op.ttab = h->ttab; op.ld_ttab = h->sum_c2; op.otab_q = w->otab_q; op.otab_B = w->B;
  op.film_off = b.film_off; op.film_c = b.cout;
  w->ops.push_back(op);
  ActBf16 h1{h1buf, b.cout, b.cout};
  d = ConvDesc();
  d.kind = CONV_K; d.taps_k = 5; d.srcs = &h1; d.nsrc = 1; d.t_in = Tl; d.wgt = b.c2w; d.cout = b.cout; d.group_width = gw;
  if (b.proj) {
    // conv2 on h1 + the 1x1 residual projection of the block input (one more accumulator) in ONE launch
    d.aux_srcs = srcs; d.aux_nsrc = nsrc; d.aux_w = b.rw;
  }
  LDP_TRY(conv_tc(h, w, 4 * bi + 1, d, &op));
  set_gn(&op, b.c2b, b.g2s, b.g2b, b.cout, h->cfg.n_groups, outbuf);
  if (b.proj) {
    op.bias_aux = b.rb;
  } else {
    op.res_bf16 = srcs[0].p; op.ld_res_bf16 = srcs[0].ld;
  }
  w->ops.push_back(op);
  return LDP_OK;
}

static int prepare_bf16(LdpPlanner* h, PlanWs* w) {
  if (w->bf16_ready) return LDP_OK;
  LDP_TRY(tc_driver_check());
  LDP_TRY(tc_gemm_init());
  const LdpUnetConfig& c = h->cfg;
  const int B = w->B, T = w->T, nl = c.n_levels;
  w->ld_xb = round_up(c.input_dim, 8);
  LDP_TRY(w->arena.alloc_t(&w->x_bf16, (size_t)B * T * w->ld_xb));
  size_t max_act = 0;
  for (int l = 0; l < nl; ++l) max_act = std::max(max_act, (size_t)B * level_len(T, l) * c.down_dims[l]);
  __nv_bfloat16* h1buf;
  LDP_TRY(w->arena.alloc_t(&h1buf, max_act));
  auto new_act = [&](int Tl, int ch, ActBf16* a) -> int {
    a->c = ch; a->ld = ch;
    return w->arena.alloc_t(&a->p, (size_t)B * Tl * ch);
  };
  w->ops.clear();
  ActBf16 cur{w->x_bf16, c.input_dim, w->ld_xb};
  std::vector<ActBf16> skips;
  int bi = 0;
  for (int l = 0; l < nl; ++l) {
    const int Tl = level_len(T, l);
    for (int r = 0; r < 2; ++r) {
      ActBf16 o;
      LDP_TRY(new_act(Tl, h->crb[bi].cout, &o));
      LDP_TRY(crb_tc(h, w, bi, &cur, 1, Tl, h1buf, o.p));
      w->taps.push_back({bi, o, (long long)B * Tl});
      cur = o; ++bi;
    }
    skips.push_back(cur);
    if (l < nl - 1) {
      const int d = c.down_dims[l];
      ActBf16 o;
      LDP_TRY(new_act(Tl / 2, d, &o));
      TcGemm op;
      ConvDesc cd;
      cd.kind = CONV_DOWN; cd.taps_k = 3; cd.srcs = &cur; cd.nsrc = 1; cd.t_in = Tl; cd.wgt = h->down_w[l]; cd.cout = d;
      cd.group_width = 32;
      LDP_TRY(conv_tc(h, w, 1000 + l, cd, &op));
      op.mode = TC_EPI_PLAIN; op.bias = h->down_b[l]; op.out_bf16 = o.p; op.ld_out_bf16 = d;
      w->ops.push_back(op);
      w->taps.push_back({100 + l, o, (long long)B * (Tl / 2)});
      cur = o;
    }
  }
  {
    const int Tl = level_len(T, nl - 1);
    for (int r = 0; r < 2; ++r) {
      ActBf16 o;
      LDP_TRY(new_act(Tl, h->crb[bi].cout, &o));
      LDP_TRY(crb_tc(h, w, bi, &cur, 1, Tl, h1buf, o.p));
      w->taps.push_back({bi, o, (long long)B * Tl});
      cur = o; ++bi;
    }
  }
  for (int u = 0; u < nl - 1; ++u) {
    const int lvl = nl - 1 - u;
    const int Tl = level_len(T, lvl);
    ActBf16 srcs[2] = {cur, skips.back()};
    skips.pop_back();
    ActBf16 o;
    LDP_TRY(new_act(Tl, h->crb[bi].cout, &o));
    LDP_TRY(crb_tc(h, w, bi, srcs, 2, Tl, h1buf, o.p));
    w->taps.push_back({bi, o, (long long)B * Tl});
    cur = o; ++bi;
    LDP_TRY(new_act(Tl, h->crb[bi].cout, &o));
    LDP_TRY(crb_tc(h, w, bi, &cur, 1, Tl, h1buf, o.p));
    w->taps.push_back({bi, o, (long long)B * Tl});
    cur = o; ++bi;
    const int d = c.down_dims[lvl - 1];
    LDP_TRY(new_act(2 * Tl, d, &o));
    TcGemm op;
    ConvDesc cd;
    cd.kind = CONV_UP; cd.taps_k = 4; cd.srcs = &cur; cd.nsrc = 1; cd.t_in = Tl; cd.wgt = h->up_w[u]; cd.cout = d;
    cd.group_width = 32;
    LDP_TRY(conv_tc(h, w, 2000 + u, cd, &op));
    op.mode = TC_EPI_PLAIN; op.bias = h->up_bias2[u]; op.out_bf16 = o.p; op.ld_out_bf16 = 2 * d;
    w->ops.push_back(op);
    w->taps.push_back({200 + u, o, (long long)B * 2 * Tl});
    cur = o;
  }
  const int d0 = c.down_dims[0];
  LDP_CHECK((d0 / 8) % 32 == 0, LDP_ERR_UNSUPPORTED, "bf16 path: final block group width must be a multiple of 32");
  ActBf16 f;
  LDP_TRY(new_act(T, d0, &f));
  TcGemm op;
  ConvDesc cd;
  cd.kind = CONV_K; cd.taps_k = 5; cd.srcs = &cur; cd.nsrc = 1; cd.t_in = T; cd.wgt = h->fcw; cd.cout = d0; cd.group_width = d0 / 8;
  LDP_TRY(conv_tc(h, w, 3000, cd, &op));
  set_gn(&op, h->fcb, h->fgs, h->fgb, d0, 8, f.p);
  w->ops.push_back(op);
  w->taps.push_back({300, f, (long long)B * T});
  cd = ConvDesc();
  cd.kind = CONV_K; cd.taps_k = 1; cd.srcs = &f; cd.nsrc = 1; cd.t_in = T; cd.wgt = h->ow; cd.cout = c.input_dim;
  cd.no_pair = true;                                   // the DDPM epilogue kernel is single-CTA
  LDP_TRY(conv_tc(h, w, 3001, cd, &op));
  op.mode = TC_EPI_DDPM;
  op.bias = h->ob;
  op.coef = h->coef;
  op.call_dev = w->call_dev;
  op.x_io = w->x_state; op.ld_x = c.input_dim;
  op.out_bf16 = w->x_bf16; op.ld_out_bf16 = w->ld_xb;
  op.step_dec = w->step_dev; op.done_counter = w->done_counter;      // the step counter advances inside this kernel
  w->ops.push_back(op);
  if (!env_off("LDP_L2PF")) {
    // every layer prefetches the next layer's packed weights into L2 (the last one: the first layer's, for the next step)
    const size_t n = w->ops.size();
    for (size_t i = 0; i < n; ++i) {
      const TcGemm& nx = w->ops[(i + 1) % n];
      w->ops[i].l2_prefetch = nx.wt_host_ref;
      w->ops[i].l2_prefetch_bytes = (unsigned)std::min<size_t>((size_t)nx.n_pad * nx.k_pad * 2, (size_t)64 << 20);
    }
  }
  w->bf16_ready = true;
  return LDP_OK;
}

static int run_ops_bf16(PlanWs* w, StepRef step, bool final_plain, float* eps_out, int D, cudaStream_t s) {
  for (size_t i = 0; i < w->ops.size(); ++i) {
    TcGemm op = w->ops[i];
    op.step = step;
    op.step.rows_per_t = step.rows ? op.rows_per_item : 1;
    if (i + 1 == w->ops.size() && final_plain) {
      op.mode = TC_EPI_PLAIN;
      op.out_f32 = eps_out; op.ld_out_f32 = D;
      op.out_bf16 = nullptr;
      op.x_io = nullptr;
      op.step_dec = nullptr;
    }
    LDP_TRY(launch_tc_gemm(op, s));
  }
  return LDP_OK;
}


// Persistent loop kernel: the ops of prepare_bf16 (CTA pairs undone: the loop kernel is single-CTA), with the group
// geometry filled in, as a device array.  Unsupported shapes leave loop_state = -1 and the caller uses the graph path.
static int prepare_loop(LdpPlanner* h, PlanWs* w) {
  if (w->loop_state != 0) return LDP_OK;
  w->loop_state = -1;
  const char* env = getenv("LDP_LOOP");
  if (!(env && env[0] == '1')) return LDP_OK;          // opt-in until it beats the per-layer graph path
  const LdpUnetConfig& c = h->cfg;
  const int t_deep = level_len(w->T, c.n_levels - 1);
  if (t_deep < 1 || t_deep > 128 || (128 % t_deep) != 0) return LDP_OK;
  const int spc = 128 / t_deep;                       // samples per group: 128 rows at the deepest level
  const int n_groups = ceil_div(w->B, spc);
  const int G = 8;
  int dev = 0, sms = 0;
  LDP_CUDA_OK(cudaGetDevice(&dev));
  LDP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (n_groups * G > sms) return LDP_OK;              // every CTA of a group must be resident: one CTA per SM
  std::vector<TcGemm> ops = w->ops;
  for (TcGemm& op : ops) {
    if (op.mode != TC_EPI_PLAIN && op.mode != TC_EPI_GN && op.mode != TC_EPI_DDPM) return LDP_OK;
    if (op.block_n != 64 && op.block_n != 128) return LDP_OK;
    if (op.mode == TC_EPI_DDPM && op.block_n != 128) return LDP_OK;
    op.allow_tail = 0;
    if (op.num_kb > 256 || op.tiles_per_item != 1) return LDP_OK;
    if ((spc * op.rows_per_item) % 128 != 0) return LDP_OK;
    if (op.pair) {
      uint64_t bd[2] = {(uint64_t)op.k_pad, (uint64_t)op.n_pad};
      uint64_t bs[1] = {(uint64_t)op.k_pad * 2};
      uint32_t bb[2] = {64, (uint32_t)op.block_n};
      LDP_TRY(make_tmap_bf16(&op.map_b, op.wt_host_ref, 2, bd, bs, bb));
      op.pair = 0;
    }
    LDP_TRY(tc_gemm_geometry(&op));
    op.num_stages = std::min(8, (196 * 1024) / (16384 + op.w_max * op.block_n * 128));
    if (op.num_stages < 2) return LDP_OK;
    op.tiles_m_group = spc * op.rows_per_item / 128;
    op.persistent = 0; op.acc_bufs = 1;
    op.step_dec = nullptr;
    op.epi_skip = getenv("LDP_LOOP_FLAGS") ? (atoi(getenv("LDP_LOOP_FLAGS")) >> 4 << 4) : 0; op.dbg = nullptr; op.dbg_stage = nullptr;
  }
  if (ops.size() > 36) return LDP_OK;
  w->loop_ops = ops;
  LDP_TRY(w->arena.alloc_t(&w->group_counter, (size_t)n_groups));
  LDP_TRY(w->arena.alloc_t(&w->loop_dbg, ops.size() * 8 + 8));
  w->n_groups = n_groups;
  w->loop_state = 1;
  return LDP_OK;
}

}  // namespace ldp

// ------------------------------- C ABI -------------------------------------------------------
extern "C" {

int64_t ldp_unet_param_count(const LdpUnetConfig* cfg) {
  if (validate_cfg(cfg) != LDP_OK) return -1;
  return unet_param_count(*cfg);
}

int ldp_planner_create(const LdpUnetConfig* cfg, const float* params_host, uint64_t n_params, LdpPlanner** out) {
  LDP_CHECK(out != nullptr && params_host != nullptr, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_TRY(validate_cfg(cfg));
  LdpPlanner* h = new LdpPlanner();
  int st = planner_create_impl(cfg, params_host, n_params, h);
  if (st != LDP_OK) {
    delete h;
    return st;
  }
  *out = h;
  return LDP_OK;
}

int ldp_planner_destroy(LdpPlanner* h) {
  if (!h) return LDP_OK;
  cudaDeviceSynchronize();
  delete h;
  return LDP_OK;
}

int ldp_unet_forward(LdpPlanner* h, int precision, const float* sample_dev, const int32_t* timesteps_dev, int timestep,
                     const float* cond_dev, int B, int T, float* eps_dev, void* cuda_stream) {
  LDP_CHECK(h && sample_dev && cond_dev && eps_dev, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_CHECK(timesteps_dev || (timestep >= 0 && timestep < h->cfg.n_train_steps), LDP_ERR_INVALID_ARG,
            "timestep outside [0, n_train_steps)");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  PlanWs* w;
  LDP_TRY(get_ws(h, B, T, &w));
  LDP_CHECK(precision == LDP_PREC_FP32 || precision == LDP_PREC_BF16, LDP_ERR_INVALID_ARG, "unknown precision");
  LDP_TRY(compute_otab(h, w, cond_dev, precision, s));
  StepRef step;
  step.rows = timesteps_dev;
  step.scalar = timestep;
  step.rows_per_t = 1;
  if (precision == LDP_PREC_FP32) return forward_f32(h, w, sample_dev, step, eps_dev, s);
  LDP_CHECK(precision == LDP_PREC_BF16, LDP_ERR_INVALID_ARG, "unknown precision");
  LDP_TRY(prepare_bf16(h, w));
  LDP_TRY(launch_cast_bf16(sample_dev, h->cfg.input_dim, w->x_bf16, w->ld_xb, (long long)B * T, h->cfg.input_dim, 0, s));
  return run_ops_bf16(w, step, /*final_plain=*/true, eps_dev, h->cfg.input_dim, s);
}

int ldp_planner_sample(LdpPlanner* h, int precision, int sampler, const float* x_T_dev, const float* cond_dev,
                       const float* noise_dev, uint64_t seed, int64_t row_offset, int B, int T, int n_steps,
                       float* x0_dev, void* cuda_stream) {
  LDP_CHECK(h && x_T_dev && cond_dev && x0_dev, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_CHECK(n_steps > 0 && n_steps <= h->cfg.n_train_steps, LDP_ERR_INVALID_ARG, "n_steps must be in [1, n_train_steps]");
  LDP_CHECK(sampler == LDP_SAMPLER_DDPM || sampler == LDP_SAMPLER_DDIM, LDP_ERR_INVALID_ARG, "unknown sampler");
  LDP_CHECK(precision == LDP_PREC_FP32 || precision == LDP_PREC_BF16, LDP_ERR_INVALID_ARG, "unknown precision");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const int D = h->cfg.input_dim;
  PlanWs* w;
  LDP_TRY(get_ws(h, B, T, &w));
  const size_t n = (size_t)B * T * D;
  LDP_TRY(compute_otab(h, w, cond_dev, precision, s));
  LDP_CUDA_OK(cudaMemcpyAsync(w->x_state, x_T_dev, n * 4, cudaMemcpyDeviceToDevice, s));
  DdpmCall call;
  call.noise = noise_dev;
  call.noise_step_stride = (long long)n;
  call.n_steps = n_steps;
  call.sampler = sampler;
  call.seed = seed;
  call.elem_offset = (long long)row_offset * T * D;
  call.row_len = D;                                  // row-structured Philox: (plan row, column quad)
  call.row_offset = (long long)row_offset * T;
  call.stream_id = 0;
  LDP_CUDA_OK(cudaMemcpyAsync(w->call_dev, &call, sizeof(call), cudaMemcpyHostToDevice, s));
  LDP_TRY(launch_set_i32(w->step_dev, n_steps - 1, s));
  StepRef step;
  step.dev = w->step_dev;
  if (precision == LDP_PREC_FP32) {
    for (int i = 0; i < n_steps; ++i) {
      LDP_TRY(forward_f32(h, w, w->x_state, step, w->eps_buf, s));
      DdpmStep d;
      d.coef = h->coef; d.step = step; d.eps = w->eps_buf; d.x = w->x_state; d.out = w->x_state;
      d.call_dev = w->call_dev; d.n = (long long)n;
      LDP_TRY(launch_ddpm_step(d, s));
      LDP_TRY(launch_add_i32(w->step_dev, -1, s));
    }
  } else {
    LDP_TRY(prepare_bf16(h, w));
    LDP_TRY(launch_cast_bf16(w->x_state, D, w->x_bf16, w->ld_xb, (long long)B * T, D, 0, s));
    LDP_TRY(prepare_loop(h, w));
    if (w->loop_state == 1) {
      // the whole loop in one persistent kernel; timestep of iteration i = n_steps - 1 - i
      LDP_CUDA_OK(cudaMemsetAsync(w->group_counter, 0, (size_t)w->n_groups * sizeof(int), s));
      PlannerLoop lp;
      // one loop launch uses the table at a time: uploads are stream-ordered behind the previous launch on this stream
      LDP_TRY(upload_planner_loop_layers(w->loop_ops.data(), (int)w->loop_ops.size(), s));
      lp.n_layers = (int)w->ops.size();
      lp.n_steps = n_steps; lp.t_first = n_steps - 1;
      lp.group_counter = w->group_counter; lp.group_ctas = 8;
      if (getenv("LDP_LOOP_FLAGS")) lp.flags = atoi(getenv("LDP_LOOP_FLAGS"));
      if (getenv("LDP_LOOP_DBG")) { lp.dbg = w->loop_dbg; lp.dbg_step = std::min(n_steps - 1, atoi(getenv("LDP_LOOP_DBG"))); }
      LDP_TRY(launch_planner_loop(lp, w->n_groups, s));
      if (lp.dbg) {
        std::vector<long long> hb(w->ops.size() * 8 + 8);
        LDP_CUDA_OK(cudaStreamSynchronize(s));
        LDP_CUDA_OK(cudaMemcpy(hb.data(), w->loop_dbg, hb.size() * 8, cudaMemcpyDeviceToHost));
        const long long t0 = hb[0];
        { const long long* d = &hb[w->ops.size() * 8]; fprintf(stderr, "last-layer epilogue (thread 64): tfull %lld | phase1 %lld bar %lld phase2 %lld bar %lld phase3 %lld\n", d[5] - t0, d[0] - d[5], d[1] - d[0], d[2] - d[1], d[3] - d[2], d[4] - d[3]); }
        for (size_t i = 0; i < w->ops.size(); ++i)
          fprintf(stderr, "loop layer %2zu: enter %8lld polled %8lld first-operands %8lld mma-issued %8lld epilogue-done %8lld tile0-done %8lld\n", i,
                  hb[i * 8 + 0] - t0, hb[i * 8 + 1] ? hb[i * 8 + 1] - t0 : 0, hb[i * 8 + 2] - t0, hb[i * 8 + 3] - t0, hb[i * 8 + 5] - t0, hb[i * 8 + 4] - t0);
      }
    } else
    if (h->use_graph && !w->graph) {
      // capture one denoising step (all GEMMs + the step-counter decrement) once per (B,T)
      cudaStream_t cs;
      LDP_CUDA_OK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      long long before = launch_count_get();
      cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
      int st = LDP_OK;
      if (e == cudaSuccess) {
        st = run_ops_bf16(w, step, false, nullptr, D, cs);     // the DDPM kernel decrements the step counter itself
        e = cudaStreamEndCapture(cs, &w->graph_src);
      }
      count_launch((int)(before - launch_count_get()));  // captured launches did not execute
      if (st != LDP_OK) { cudaStreamDestroy(cs); return st; }
      LDP_CUDA_OK(e);
      LDP_CUDA_OK(cudaGraphInstantiate(&w->graph, w->graph_src, 0));
      LDP_CUDA_OK(cudaStreamDestroy(cs));
    }
    static const bool whole_loop_graph = !env_off("LDP_LOOP_GRAPH");
    if (w->loop_state != 1 && h->use_graph && whole_loop_graph && n_steps >= 4) {
      auto it = w->loop_graphs.find(n_steps);
      if (it == w->loop_graphs.end()) {
        cudaStream_t cs;
        LDP_CUDA_OK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        const long long before = launch_count_get();
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        int st = LDP_OK;
        if (e == cudaSuccess) {
          for (int i = 0; i < n_steps && st == LDP_OK; ++i) st = run_ops_bf16(w, step, false, nullptr, D, cs);
          e = cudaStreamEndCapture(cs, &g);
        }
        count_launch((int)(before - launch_count_get()));
        cudaGraphExec_t ex = nullptr;
        if (st == LDP_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&ex, g, 0);
        if (g) cudaGraphDestroy(g);
        cudaStreamDestroy(cs);
        if (st != LDP_OK) return st;
        if (e != cudaSuccess) { cudaGetLastError(); ex = nullptr; }      // fall back to per-step graphs below
        it = w->loop_graphs.emplace(n_steps, ex).first;
      }
      if (it->second) {
        LDP_CUDA_OK(cudaGraphLaunch(it->second, s));
        count_launch((int)w->ops.size() * n_steps);
        LDP_CUDA_OK(cudaMemcpyAsync(x0_dev, w->x_state, n * 4, cudaMemcpyDeviceToDevice, s));
        return LDP_OK;
      }
    }
    for (int i = 0; i < n_steps && w->loop_state != 1; ++i) {
      if (w->graph) {
        LDP_CUDA_OK(cudaGraphLaunch(w->graph, s));
        count_launch((int)w->ops.size());
      } else {
        LDP_TRY(run_ops_bf16(w, step, false, nullptr, D, s));
      }
    }
  }
  LDP_CUDA_OK(cudaMemcpyAsync(x0_dev, w->x_state, n * 4, cudaMemcpyDeviceToDevice, s));
  return LDP_OK;
}


// Diagnostics for per-layer parity: the bf16 activation `tap_id` left behind by the last bf16 forward at (B, T), as f32.
int ldp_planner_read_activation(LdpPlanner* h, int B, int T, int tap_id, float* out_dev, int64_t max_elems, int32_t* rows_out,
                                int32_t* cols_out, void* cuda_stream) {
  LDP_CHECK(h && out_dev && rows_out && cols_out, LDP_ERR_INVALID_ARG, "null pointer");
  auto it = h->ws.find(std::make_pair(B, T));
  LDP_CHECK(it != h->ws.end() && it->second->bf16_ready, LDP_ERR_INVALID_ARG, "no bf16 forward has run at this (B, T)");
  for (const PlanWs::Tap& t : it->second->taps) {
    if (t.id != tap_id) continue;
    LDP_CHECK(t.rows * t.act.c <= max_elems, LDP_ERR_INVALID_ARG, "output buffer too small");
    *rows_out = (int32_t)t.rows;
    *cols_out = t.act.c;
    return launch_cast_f32_from_bf16(t.act.p, t.act.ld, out_dev, t.act.c, t.rows, t.act.c, (cudaStream_t)cuda_stream);
  }
  set_last_error("unknown activation id " + std::to_string(tap_id));
  return LDP_ERR_INVALID_ARG;
}

// Diagnostics: time every kernel of one bf16 denoising step in isolation (reps back-to-back launches, CUDA events).
int ldp_planner_profile_step(LdpPlanner* h, int B, int T, int reps, float* us_host, int32_t* meta_host, float* phases_host,
                             int max_ops, int* n_ops, void* cuda_stream) {
  LDP_CHECK(h && us_host && meta_host && n_ops && reps > 0, LDP_ERR_INVALID_ARG, "bad arguments");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  PlanWs* w;
  LDP_TRY(get_ws(h, B, T, &w));
  LDP_TRY(prepare_bf16(h, w));
  LDP_CHECK((int)w->ops.size() <= max_ops, LDP_ERR_INVALID_ARG, "output arrays too small");
  DdpmCall call;
  call.n_steps = 1;
  call.seed = 1;
  call.row_len = h->cfg.input_dim;
  LDP_CUDA_OK(cudaMemcpyAsync(w->call_dev, &call, sizeof(call), cudaMemcpyHostToDevice, s));
  LDP_TRY(launch_set_i32(w->step_dev, h->cfg.n_train_steps / 2, s));
  StepRef step;
  step.dev = w->step_dev;
  cudaEvent_t e0, e1;
  LDP_CUDA_OK(cudaEventCreate(&e0));
  LDP_CUDA_OK(cudaEventCreate(&e1));
  for (size_t i = 0; i < w->ops.size(); ++i) {
    TcGemm op = w->ops[i];
    op.step = step;
    op.step_dec = nullptr;                      // isolated launches must not advance the timestep
    for (int r = 0; r < 3; ++r) LDP_TRY(launch_tc_gemm(op, s));
    LDP_CUDA_OK(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; ++r) LDP_TRY(launch_tc_gemm(op, s));
    LDP_CUDA_OK(cudaEventRecord(e1, s));
    LDP_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    LDP_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    us_host[i] = ms * 1000.f / reps;
    if (phases_host) {          // one more launch with the in-kernel phase clocks; mean over CTAs
      LDP_TRY(tc_gemm_geometry(&op));
      const int ctas = op.grid_ctas;
      std::vector<long long> hb((size_t)ctas * 8 + 64);
      long long* db;
      Arena tmp;
      LDP_TRY(tmp.alloc_t(&db, hb.size() + 64));
      op.dbg = db;
      op.dbg_stage = db + (size_t)ctas * 8;
      LDP_TRY(launch_tc_gemm(op, s));
      LDP_CUDA_OK(cudaStreamSynchronize(s));
      LDP_CUDA_OK(cudaMemcpy(hb.data(), db, hb.size() * 8, cudaMemcpyDeviceToHost));
      for (int k = 1; k < 8; ++k) {
        double acc = 0;
        for (int c = 0; c < ctas; ++c) acc += (double)hb[(size_t)c * 8 + k];
        phases_host[8 * i + k] = (float)(acc / ctas);
      }
      phases_host[8 * i + 0] = (float)ctas;
      if (getenv("LDP_DBG_STAGES")) {
        fprintf(stderr, "op %zu stage arrivals (CTA 0):", i);
        for (int k = 0; k < 23; ++k) fprintf(stderr, " %lld", hb[(size_t)ctas * 8 + k]);
        fprintf(stderr, " | epilogue deltas:");
        for (int k = 24; k < 29; ++k) fprintf(stderr, " %lld", hb[(size_t)ctas * 8 + k] - hb[(size_t)ctas * 8 + 23]);
        fprintf(stderr, "\n");
      }
    }
    meta_host[4 * i + 0] = op.M;
    meta_host[4 * i + 1] = op.N;
    meta_host[4 * i + 2] = op.k_pad / 64;
    meta_host[4 * i + 3] = op.block_n | (op.mode << 16) | (op.use_aux << 24) | (op.n_acc << 25);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *n_ops = (int)w->ops.size();
  return LDP_OK;
}

}  // extern "C"
