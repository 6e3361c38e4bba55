// Inverse-dynamics handle: MLPDiffusion score network (reference networks/mlp_diffusion_nets.py:50-68:
// FourierFeatures -> MLP cond encoder -> MLPResNet over concat([a, s||s', cond])) and the action denoising
// loop around it (reference agent/ldp_agent.py:486-505).
//
// Hoisting: the first Dense of the MLPResNet acts on concat([a | s||s' | cond]); its three row blocks are
// separated:  cond part -> table over the N timesteps (batch-invariant, built at create),
//             s||s' part -> one GEMM per act() (step-invariant),
//             a part     -> K = action_dim (7 / 14), done per step in the small fused input kernel below.
#include <algorithm>

#include "net_common.h"

namespace ldp {

struct IdmBlockW {
  const float *ln_s, *ln_b, *w1, *b1, *w2, *b2;
};

struct IdmWs {
  int N = 0;
  uint64_t last_use = 0;         // LRU stamp (ws_evict_lru)
  Arena arena;
  float *spre = nullptr, *h = nullptr, *hn = nullptr, *u = nullptr, *a_state = nullptr, *eps_buf = nullptr;
  int32_t* step_dev = nullptr;
  unsigned int* done_counter = nullptr;
  DdpmCall* call_dev = nullptr;
  bool bf16_ready = false;
  __nv_bfloat16 *hn_b = nullptr, *u_b = nullptr;
  std::vector<TcGemm> ops;     // per block: up-projection, down-projection; last = output Dense with the DDPM epilogue
  cudaGraphExec_t graph = nullptr;
  cudaGraph_t graph_src = nullptr;
  ~IdmWs() {
    if (graph) cudaGraphExecDestroy(graph);
    if (graph_src) cudaGraphDestroy(graph_src);
  }
};

// h[row] = a[row] Wa + spre[row] + ctabH[step];  hn = LayerNorm(h) (first block's affine) as f32 and/or bf16.
// One warp per row; H <= 512.
__global__ void idm_input_kernel(const float* __restrict__ a, int A, const float* __restrict__ wa, const float* __restrict__ spre,
                                 const float* __restrict__ ctab_h, StepRef step, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, float* __restrict__ h, float* __restrict__ hn_f,
                                 __nv_bfloat16* __restrict__ hn_b, int N, int H) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* trow = ctab_h + (long long)step_of(step, row) * H;
  float av = lane < A ? a[(long long)row * A + lane] : 0.f;
  float v[16];
  const int per = H >> 5;
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i < per) {
      const int n = i * 32 + lane;
      float acc = spre[(long long)row * H + n] + trow[n];
      for (int j = 0; j < A; ++j) acc = fmaf(__shfl_sync(0xffffffffu, av, j), wa[j * H + n], acc);
      v[i] = acc;
      s += acc;
      ss += acc * acc;
      h[(long long)row * H + n] = acc;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const float mu = s / H, rs = rsqrtf(fmaxf(ss / H - mu * mu, 0.f) + eps);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i < per) {
      const int n = i * 32 + lane;
      float y = (v[i] - mu) * rs * gamma[n] + beta[n];
      if (hn_f) hn_f[(long long)row * H + n] = y;
      if (hn_b) hn_b[(long long)row * H + n] = __float2bfloat16(y);
    }
  }
}

}  // namespace ldp

using namespace ldp;

struct LdpIdm {
  LdpIdmConfig cfg;
  Arena arena;
  float* blob = nullptr;
  std::vector<const float*> cw, cb;     // cond MLP
  const float *w0, *b0, *wout, *bout;
  std::vector<IdmBlockW> blk;
  int cond_out = 0;
  float* ctab_h = nullptr;   // [n_train][H]: cond(t) Wc + b0
  float* coef = nullptr;
  std::map<int, std::unique_ptr<IdmWs>> ws;
  uint64_t use_clock = 0;
  // bf16 packed weights
  bool packed = false;
  std::vector<PackedW> pw1, pw2;
  PackedW pwout;
  bool use_graph = true;
  // persistent loop kernel (idm_loop.cu)
  bool loop_ready = false;
  IdmLoop loop;
  float* bsum = nullptr;     // [n_blocks][H] running sums of the down-projection biases
};

namespace ldp {

static int64_t idm_param_count(const LdpIdmConfig& c) {
  int64_t n = 0, ch = c.time_dim;
  for (int i = 0; i < c.n_cond_layers; ++i) {
    n += ch * c.cond_hidden[i] + c.cond_hidden[i];
    ch = c.cond_hidden[i];
  }
  const int64_t H = c.hidden_dim, in = c.action_dim + 2 * (int64_t)c.obs_dim + ch;
  n += in * H + H;
  n += (int64_t)c.n_blocks * (2 * H + H * 4 * H + 4 * H + 4 * H * H + H);
  n += H * c.action_dim + c.action_dim;
  return n;
}

static int idm_validate(const LdpIdmConfig* c) {
  LDP_CHECK(c != nullptr, LDP_ERR_INVALID_ARG, "null config");
  LDP_CHECK(c->obs_dim > 0 && c->action_dim > 0 && c->action_dim <= 32, LDP_ERR_INVALID_ARG, "action_dim must be in 1..32");
  LDP_CHECK(c->hidden_dim > 0 && c->hidden_dim % 32 == 0 && c->hidden_dim <= 512, LDP_ERR_UNSUPPORTED,
            "hidden_dim must be a multiple of 32, <= 512");
  LDP_CHECK(c->n_blocks >= 1 && c->time_dim >= 4 && c->time_dim % 2 == 0, LDP_ERR_INVALID_ARG, "bad n_blocks/time_dim");
  LDP_CHECK(c->n_cond_layers >= 1 && c->n_cond_layers <= 4 && c->n_train_steps > 0, LDP_ERR_INVALID_ARG, "bad cond MLP");
  return LDP_OK;
}

static int idm_create_impl(const LdpIdmConfig* cfg, const float* params_host, uint64_t n_params, LdpIdm* h) {
  h->cfg = *cfg;
  const LdpIdmConfig& c = h->cfg;
  const int64_t expect = idm_param_count(c);
  LDP_CHECK((int64_t)n_params == expect, LDP_ERR_PARAM_COUNT,
            "IDM weight blob has " + std::to_string(n_params) + " floats, config needs " + std::to_string(expect));
  LDP_TRY(h->arena.alloc_t(&h->blob, n_params, false));
  LDP_CUDA_OK(cudaMemcpy(h->blob, params_host, n_params * 4, cudaMemcpyHostToDevice));
  BlobWalker w{h->blob, 0};
  int64_t ch = c.time_dim;
  for (int i = 0; i < c.n_cond_layers; ++i) {
    h->cw.push_back(w.take(ch * c.cond_hidden[i]));
    h->cb.push_back(w.take(c.cond_hidden[i]));
    ch = c.cond_hidden[i];
  }
  h->cond_out = (int)ch;
  const int64_t H = c.hidden_dim, A = c.action_dim, D2 = 2 * (int64_t)c.obs_dim;
  h->w0 = w.take((A + D2 + ch) * H);
  h->b0 = w.take(H);
  for (int b = 0; b < c.n_blocks; ++b) {
    IdmBlockW k;
    k.ln_s = w.take(H); k.ln_b = w.take(H);
    k.w1 = w.take(H * 4 * H); k.b1 = w.take(4 * H);
    k.w2 = w.take(4 * H * H); k.b2 = w.take(H);
    h->blk.push_back(k);
  }
  h->wout = w.take(H * A);
  h->bout = w.take(A);
  LDP_CHECK((int64_t)w.pos == expect, LDP_ERR_PARAM_COUNT, "internal: blob walk mismatch");
  // cond table: FourierFeatures [cos|sin] -> MLP (mish between layers, none after the last; ldp_agent.yaml:30-34)
  cudaStream_t s = 0;
  const int n = c.n_train_steps;
  Arena tmp;
  float *bufa, *bufb;
  int maxw = c.time_dim;
  for (int i = 0; i < c.n_cond_layers; ++i) maxw = std::max(maxw, c.cond_hidden[i]);
  LDP_TRY(tmp.alloc_t(&bufa, (size_t)n * maxw));
  LDP_TRY(tmp.alloc_t(&bufb, (size_t)n * maxw));
  LDP_TRY(launch_sinusoid_table(bufa, n, c.time_dim, /*cos_first=*/1, s));
  int cin = c.time_dim;
  for (int i = 0; i < c.n_cond_layers; ++i) {
    GemmF32 g;
    g.x1 = bufa; g.c1 = cin; g.ld1 = cin; g.a_act = i > 0 ? 1 : 0;     // Mish on the previous layer's output
    g.w = h->cw[i]; g.ldw = c.cond_hidden[i]; g.bias = h->cb[i]; g.out = bufb; g.ldo = c.cond_hidden[i];
    g.m = n; g.n = c.cond_hidden[i];
    LDP_TRY(launch_gemm_f32(g, s));
    std::swap(bufa, bufb);
    cin = c.cond_hidden[i];
  }
  LDP_TRY(h->arena.alloc_t(&h->ctab_h, (size_t)n * H));
  GemmF32 g;
  g.x1 = bufa; g.c1 = cin; g.ld1 = cin; g.w = h->w0 + (A + D2) * H; g.ldw = (int)H; g.bias = h->b0;
  g.out = h->ctab_h; g.ldo = (int)H; g.m = n; g.n = (int)H;
  LDP_TRY(launch_gemm_f32(g, s));
  std::vector<float> coef;
  ddpm_coef_host(n, coef);
  LDP_TRY(h->arena.alloc_t(&h->coef, coef.size()));
  LDP_CUDA_OK(cudaMemcpy(h->coef, coef.data(), coef.size() * 4, cudaMemcpyHostToDevice));
  LDP_CUDA_OK(cudaStreamSynchronize(s));
  const char* env = getenv("LDP_NO_GRAPH");
  h->use_graph = !(env && env[0] == '1');
  return LDP_OK;
}

static int idm_get_ws(LdpIdm* h, int N, IdmWs** out) {
  LDP_CHECK(N > 0, LDP_ERR_BAD_SHAPE, "N must be positive");
  auto it = h->ws.find(N);
  if (it != h->ws.end()) {
    it->second->last_use = ++h->use_clock;
    *out = it->second.get();
    return LDP_OK;
  }
  ws_evict_lru(h->ws);
  const int H = h->cfg.hidden_dim, A = h->cfg.action_dim;
  std::unique_ptr<IdmWs> w(new IdmWs());
  w->N = N; w->last_use = ++h->use_clock;
  LDP_TRY(w->arena.alloc_t(&w->spre, (size_t)N * H));
  LDP_TRY(w->arena.alloc_t(&w->h, (size_t)N * H));
  LDP_TRY(w->arena.alloc_t(&w->hn, (size_t)N * H));
  LDP_TRY(w->arena.alloc_t(&w->u, (size_t)N * 4 * H));
  LDP_TRY(w->arena.alloc_t(&w->a_state, (size_t)N * A));
  LDP_TRY(w->arena.alloc_t(&w->eps_buf, (size_t)N * A));
  LDP_TRY(w->arena.alloc_t(&w->step_dev, 4));
  LDP_TRY(w->arena.alloc_t(&w->done_counter, 4));
  LDP_TRY(w->arena.alloc_t(&w->call_dev, 1));
  *out = w.get();
  h->ws[N] = std::move(w);
  return LDP_OK;
}

static int idm_spre(LdpIdm* h, IdmWs* w, const float* s_dev, cudaStream_t s) {
  const int H = h->cfg.hidden_dim, A = h->cfg.action_dim, D2 = 2 * h->cfg.obs_dim;
  GemmF32 g;
  g.x1 = s_dev; g.c1 = D2; g.ld1 = D2; g.w = h->w0 + (size_t)A * H; g.ldw = H; g.out = w->spre; g.ldo = H;
  g.m = w->N; g.n = H;
  return launch_gemm_f32(g, s);
}

static int idm_input(LdpIdm* h, IdmWs* w, const float* a, StepRef step, bool bf16, cudaStream_t s) {
  const int H = h->cfg.hidden_dim;
  idm_input_kernel<<<ceil_div(w->N, 8), 256, 0, s>>>(a, h->cfg.action_dim, h->w0, w->spre, h->ctab_h, step, h->blk[0].ln_s,
                                                     h->blk[0].ln_b, 1e-6f, w->h, bf16 ? nullptr : w->hn,
                                                     bf16 ? w->hn_b : nullptr, w->N, H);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(std::string("idm_input launch failed: ") + cudaGetErrorString(e));
    return LDP_ERR_CUDA;
  }
  count_launch();
  return LDP_OK;
}

static int idm_forward_f32(LdpIdm* h, IdmWs* w, const float* a, StepRef step, float* eps, cudaStream_t s) {
  const int H = h->cfg.hidden_dim, A = h->cfg.action_dim, nb = h->cfg.n_blocks;
  LDP_TRY(idm_input(h, w, a, step, false, s));
  for (int b = 0; b < nb; ++b) {
    const IdmBlockW& k = h->blk[b];
    if (b > 0) LDP_TRY(launch_layernorm_f32(w->h, w->hn, w->N, H, k.ln_s, k.ln_b, 1e-6f, 0, s));
    GemmF32 g;
    g.x1 = w->hn; g.c1 = H; g.ld1 = H; g.w = k.w1; g.ldw = 4 * H; g.bias = k.b1; g.act = 1; g.out = w->u; g.ldo = 4 * H;
    g.m = w->N; g.n = 4 * H;
    LDP_TRY(launch_gemm_f32(g, s));
    g = GemmF32();
    g.x1 = w->u; g.c1 = 4 * H; g.ld1 = 4 * H; g.w = k.w2; g.ldw = H; g.bias = k.b2; g.res = w->h; g.ldres = H;
    g.out = w->h; g.ldo = H; g.m = w->N; g.n = H;
    LDP_TRY(launch_gemm_f32(g, s));
  }
  LDP_TRY(launch_layernorm_f32(w->h, w->hn, w->N, H, nullptr, nullptr, 0.f, /*relu_instead=*/1, s));
  GemmF32 g;
  g.x1 = w->hn; g.c1 = H; g.ld1 = H; g.w = h->wout; g.ldw = A; g.bias = h->bout; g.out = eps; g.ldo = A; g.m = w->N; g.n = A;
  return launch_gemm_f32(g, s);
}

// ---- bf16 ----
static int pack_dense(LdpIdm* h, const float* wgt, int K, int N, int block_n, PackedW* pw) {
  pw->kp = round_up(K, 64);
  pw->n_pad = round_up(N, block_n);
  pw->num_kb = pw->kp / 64;
  LDP_TRY(h->arena.alloc_t(&pw->wt, (size_t)pw->n_pad * pw->kp));
  std::vector<int32_t> kmap(pw->kp);
  std::vector<TcKBlock> kb(pw->num_kb);
  for (int k = 0; k < pw->kp; ++k) kmap[k] = k < K ? k : -1;
  for (int i = 0; i < pw->num_kb; ++i) kb[i] = make_stage(0, 0, 1, i * 64, 0, 0, i);
  Arena tmp;
  int32_t* map_dev;
  LDP_TRY(tmp.alloc_t(&map_dev, pw->kp));
  LDP_CUDA_OK(cudaMemcpy(map_dev, kmap.data(), (size_t)pw->kp * 4, cudaMemcpyHostToDevice));
  LDP_TRY(upload_stage_table(h->arena, kb, pw));
  LDP_TRY(launch_pack_wt_bf16(wgt, N, N, map_dev, pw->kp, pw->wt, pw->kp, 0, pw->n_pad, 0));
  LDP_CUDA_OK(cudaDeviceSynchronize());
  return LDP_OK;
}

static int dense_op(const PackedW& pw, const __nv_bfloat16* a, int K, int lda, int M, int N, int block_n, TcGemm* op) {
  *op = TcGemm();
  uint64_t ad[4] = {(uint64_t)K, 1, 1, (uint64_t)M};
  uint64_t as[3] = {(uint64_t)lda * 2, (uint64_t)lda * 2, (uint64_t)lda * 2};
  uint32_t ab[4] = {64, 1, 1, 128};
  LDP_TRY(make_tmap_bf16(&op->map_a[0], a, 4, ad, as, ab));
  for (int i = 1; i < 4; ++i) op->map_a[i] = op->map_a[0];
  uint64_t bd[2] = {(uint64_t)pw.kp, (uint64_t)pw.n_pad};
  uint64_t bs[1] = {(uint64_t)pw.kp * 2};
  uint32_t bb[2] = {64, (uint32_t)block_n};
  LDP_TRY(make_tmap_bf16(&op->map_b, pw.wt, 2, bd, bs, bb));
  op->kb = pw.kb_dev; op->num_kb = pw.num_kb; op->runs = pw.runs_dev; op->num_runs = pw.num_runs; tc_set_inline_runs(op, pw.runs_host.data(), pw.num_runs); op->M = M; op->N = N; op->block_n = block_n;
  op->items_per_tile = 128; op->rows_per_item = 1;
  return LDP_OK;
}

static int idm_prepare_bf16(LdpIdm* h, IdmWs* w) {
  if (w->bf16_ready) return LDP_OK;
  const int H = h->cfg.hidden_dim, A = h->cfg.action_dim, nb = h->cfg.n_blocks;
  LDP_CHECK(H == 256, LDP_ERR_UNSUPPORTED, "bf16 IDM path needs hidden_dim == 256 (LayerNorm row must equal the 256-wide N tile)");
  LDP_TRY(tc_driver_check());
  LDP_TRY(tc_gemm_init());
  if (!h->packed) {
    h->pw1.resize(nb);
    h->pw2.resize(nb);
    for (int b = 0; b < nb; ++b) {
      LDP_TRY(pack_dense(h, h->blk[b].w1, H, 4 * H, 256, &h->pw1[b]));
      LDP_TRY(pack_dense(h, h->blk[b].w2, 4 * H, H, 256, &h->pw2[b]));
    }
    LDP_TRY(pack_dense(h, h->wout, H, A, 128, &h->pwout));
    h->packed = true;
  }
  LDP_TRY(w->arena.alloc_t(&w->hn_b, (size_t)w->N * H));
  LDP_TRY(w->arena.alloc_t(&w->u_b, (size_t)w->N * 4 * H));
  w->ops.clear();
  for (int b = 0; b < nb; ++b) {
    TcGemm op;
    LDP_TRY(dense_op(h->pw1[b], w->hn_b, H, H, w->N, 4 * H, 256, &op));
    op.mode = TC_EPI_PLAIN; op.bias = h->blk[b].b1; op.relu = 1; op.out_bf16 = w->u_b; op.ld_out_bf16 = 4 * H;
    w->ops.push_back(op);
    LDP_TRY(dense_op(h->pw2[b], w->u_b, 4 * H, 4 * H, w->N, H, 256, &op));
    op.mode = TC_EPI_LN; op.bias = h->blk[b].b2; op.res_f32 = w->h; op.ld_res_f32 = H; op.out_f32 = w->h; op.ld_out_f32 = H;
    op.out_bf16 = w->hn_b; op.ld_out_bf16 = H; op.eps = 1e-6f;
    if (b + 1 < nb) { op.gamma = h->blk[b + 1].ln_s; op.beta = h->blk[b + 1].ln_b; op.relu = 0; }
    else            { op.relu = 1; }      // after the last block: activations(x) = relu, no norm (mlp_diffusion_nets.py:46)
    w->ops.push_back(op);
  }
  TcGemm op;
  LDP_TRY(dense_op(h->pwout, w->hn_b, H, H, w->N, A, 128, &op));
  op.mode = TC_EPI_DDPM; op.bias = h->bout; op.coef = h->coef; op.call_dev = w->call_dev;
  op.x_io = w->a_state; op.ld_x = A;
  op.step_dec = w->step_dev; op.done_counter = w->done_counter;      // the step counter advances inside this kernel
  w->ops.push_back(op);
  w->bf16_ready = true;
  return LDP_OK;
}

// Weight-side arguments of the persistent loop kernel (maps over the packed weights, bias sums); built once per handle.
static int idm_prepare_loop(LdpIdm* h) {
  if (h->loop_ready) return LDP_OK;
  const int H = h->cfg.hidden_dim, A = h->cfg.action_dim, nb = h->cfg.n_blocks;
  IdmLoop& lp = h->loop;
  lp = IdmLoop();
  std::vector<float> sums((size_t)nb * H, 0.f), tmp(H);
  for (int b = 0; b < nb; ++b) {
    LDP_CUDA_OK(cudaMemcpy(tmp.data(), h->blk[b].b2, (size_t)H * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < H; ++i) sums[(size_t)b * H + i] = (b ? sums[(size_t)(b - 1) * H + i] : 0.f) + tmp[i];
  }
  LDP_TRY(h->arena.alloc_t(&h->bsum, sums.size()));
  LDP_CUDA_OK(cudaMemcpy(h->bsum, sums.data(), sums.size() * 4, cudaMemcpyHostToDevice));
  for (int b = 0; b < nb; ++b) {
    uint64_t d1[2] = {(uint64_t)h->pw1[b].kp, (uint64_t)h->pw1[b].n_pad};
    uint64_t s1[1] = {(uint64_t)h->pw1[b].kp * 2};
    uint32_t b1[2] = {64, 128};
    LDP_TRY(make_tmap_bf16(&lp.map_w1[b], h->pw1[b].wt, 2, d1, s1, b1));
    uint64_t d2[2] = {(uint64_t)h->pw2[b].kp, (uint64_t)h->pw2[b].n_pad};
    uint64_t s2[1] = {(uint64_t)h->pw2[b].kp * 2};
    uint32_t b2[2] = {64, 256};
    LDP_TRY(make_tmap_bf16(&lp.map_w2[b], h->pw2[b].wt, 2, d2, s2, b2));
    lp.b1[b] = h->blk[b].b1;
    lp.bsum[b] = h->bsum + (size_t)b * H;
    lp.ln_g[b] = h->blk[b].ln_s;
    lp.ln_b[b] = h->blk[b].ln_b;
  }
  for (int b = nb; b < 4; ++b) { lp.map_w1[b] = lp.map_w1[0]; lp.map_w2[b] = lp.map_w2[0]; }
  uint64_t d3[2] = {(uint64_t)h->pwout.kp, (uint64_t)h->pwout.n_pad};
  uint64_t s3[1] = {(uint64_t)h->pwout.kp * 2};
  uint32_t b3[2] = {64, 16};
  LDP_TRY(make_tmap_bf16(&lp.map_wout, h->pwout.wt, 2, d3, s3, b3));
  static unsigned long long gen = 0;
  lp.const_gen = ++gen;                       // a new handle may reuse a freed handle's addresses
  lp.n_blocks = nb; lp.A = A;
  lp.ctab_h = h->ctab_h; lp.wa = h->w0; lp.bout = h->bout; lp.coef = h->coef;
  h->loop_ready = true;
  return LDP_OK;
}

static int idm_run_bf16(LdpIdm* h, IdmWs* w, const float* a, StepRef step, bool final_plain, float* eps_out, cudaStream_t s) {
  LDP_TRY(idm_input(h, w, a, step, true, s));
  for (size_t i = 0; i < w->ops.size(); ++i) {
    TcGemm op = w->ops[i];
    op.step = step;
    if (i + 1 == w->ops.size() && final_plain) {
      op.mode = TC_EPI_PLAIN;
      op.out_f32 = eps_out; op.ld_out_f32 = h->cfg.action_dim;
      op.x_io = nullptr;
      op.step_dec = nullptr;
    }
    LDP_TRY(launch_tc_gemm(op, s));
  }
  return LDP_OK;
}

}  // namespace ldp

extern "C" {

int64_t ldp_idm_param_count(const LdpIdmConfig* cfg) {
  if (idm_validate(cfg) != LDP_OK) return -1;
  return idm_param_count(*cfg);
}

int ldp_idm_create(const LdpIdmConfig* cfg, const float* params_host, uint64_t n_params, LdpIdm** out) {
  LDP_CHECK(out != nullptr && params_host != nullptr, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_TRY(idm_validate(cfg));
  LdpIdm* h = new LdpIdm();
  int st = idm_create_impl(cfg, params_host, n_params, h);
  if (st != LDP_OK) {
    delete h;
    return st;
  }
  *out = h;
  return LDP_OK;
}

int ldp_idm_destroy(LdpIdm* h) {
  if (!h) return LDP_OK;
  cudaDeviceSynchronize();
  delete h;
  return LDP_OK;
}

int ldp_idm_forward(LdpIdm* h, int precision, const float* s_dev, const float* a_dev, const int32_t* timesteps_dev,
                    int timestep, int N, float* eps_dev, void* cuda_stream) {
  LDP_CHECK(h && s_dev && a_dev && eps_dev, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_CHECK(timesteps_dev || (timestep >= 0 && timestep < h->cfg.n_train_steps), LDP_ERR_INVALID_ARG,
            "timestep outside [0, n_train_steps)");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  IdmWs* w;
  LDP_TRY(idm_get_ws(h, N, &w));
  LDP_TRY(idm_spre(h, w, s_dev, s));
  StepRef step;
  step.rows = timesteps_dev;
  step.scalar = timestep;
  if (precision == LDP_PREC_FP32) return idm_forward_f32(h, w, a_dev, step, eps_dev, s);
  LDP_CHECK(precision == LDP_PREC_BF16, LDP_ERR_INVALID_ARG, "unknown precision");
  LDP_TRY(idm_prepare_bf16(h, w));
  return idm_run_bf16(h, w, a_dev, step, true, eps_dev, s);
}

int ldp_idm_sample(LdpIdm* h, int precision, int sampler, const float* s_dev, const float* a_T_dev, const float* noise_dev,
                   uint64_t seed, int64_t row_offset, int N, int n_steps, float* a0_dev, void* cuda_stream) {
  LDP_CHECK(h && s_dev && a_T_dev && a0_dev, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_CHECK(n_steps > 0 && n_steps <= h->cfg.n_train_steps, LDP_ERR_INVALID_ARG, "n_steps must be in [1, n_train_steps]");
  LDP_CHECK(sampler == LDP_SAMPLER_DDPM || sampler == LDP_SAMPLER_DDIM, LDP_ERR_INVALID_ARG, "unknown sampler");
  LDP_CHECK(precision == LDP_PREC_FP32 || precision == LDP_PREC_BF16, LDP_ERR_INVALID_ARG, "unknown precision");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const int A = h->cfg.action_dim;
  IdmWs* w;
  LDP_TRY(idm_get_ws(h, N, &w));
  const size_t n = (size_t)N * A;
  LDP_TRY(idm_spre(h, w, s_dev, s));
  LDP_CUDA_OK(cudaMemcpyAsync(w->a_state, a_T_dev, n * 4, cudaMemcpyDeviceToDevice, s));
  DdpmCall call;
  call.noise = noise_dev;
  call.noise_step_stride = (long long)n;
  call.n_steps = n_steps;
  call.sampler = sampler;
  call.seed = seed;
  call.elem_offset = (long long)row_offset * A;
  call.row_len = A;                                  // row-structured Philox: (transition row, column quad)
  call.row_offset = (long long)row_offset;
  call.stream_id = 1;
  LDP_CUDA_OK(cudaMemcpyAsync(w->call_dev, &call, sizeof(call), cudaMemcpyHostToDevice, s));
  LDP_TRY(launch_set_i32(w->step_dev, n_steps - 1, s));
  StepRef step;
  step.dev = w->step_dev;
  if (precision == LDP_PREC_FP32) {
    for (int i = 0; i < n_steps; ++i) {
      LDP_TRY(idm_forward_f32(h, w, w->a_state, step, w->eps_buf, s));
      DdpmStep d;
      d.coef = h->coef; d.step = step; d.eps = w->eps_buf; d.x = w->a_state; d.out = w->a_state;
      d.call_dev = w->call_dev; d.n = (long long)n;
      LDP_TRY(launch_ddpm_step(d, s));
      LDP_TRY(launch_add_i32(w->step_dev, -1, s));
    }
  } else {
    LDP_TRY(idm_prepare_bf16(h, w));
    const char* le = getenv("LDP_IDM_LOOP");
    if (!(le && le[0] == '0') && h->cfg.hidden_dim == 256 && A <= 16 && h->cfg.n_blocks <= 4) {
      // the whole reverse loop in one persistent launch (idm_loop.cu)
      LDP_TRY(idm_prepare_loop(h));
      IdmLoop lp = h->loop;
      lp.N = N; lp.n_steps = n_steps; lp.t_first = n_steps - 1;
      lp.spre = w->spre; lp.a_state = w->a_state; lp.call = call;
      const char* de = getenv("LDP_IDM_LOOP_DBG");
      long long* dbg_dev = nullptr;
      Arena dbg_arena;
      if (de && de[0] == '1') { LDP_TRY(dbg_arena.alloc_t(&dbg_dev, 16)); lp.dbg = dbg_dev; }
      LDP_TRY(launch_idm_loop(lp, s));
      count_launch();
      if (dbg_dev) {
        long long hb[16];
        LDP_CUDA_OK(cudaStreamSynchronize(s));
        LDP_CUDA_OK(cudaMemcpy(hb, dbg_dev, sizeof(hb), cudaMemcpyDeviceToHost));
        fprintf(stderr, "idm loop (CTA 0, %d steps): MMA thread total %lld cycles, waiting: weights %lld, U drained %lld, operand chunk %lld, "
                "LayerNorm %lld | epilogue thread: input %lld, LayerNorm %lld, chunk drain work %lld, wait chunk %lld, wait block %lld, output+step %lld\n",
                n_steps, hb[4], hb[0], hb[1], hb[2], hb[3], hb[8], hb[9], hb[10], hb[11], hb[12], hb[13]);
      }
      LDP_CUDA_OK(cudaMemcpyAsync(a0_dev, w->a_state, n * 4, cudaMemcpyDeviceToDevice, s));
      return LDP_OK;
    }
    if (h->use_graph && !w->graph) {
      cudaStream_t cs;
      LDP_CUDA_OK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      long long before = launch_count_get();
      cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
      int st = LDP_OK;
      if (e == cudaSuccess) {
        st = idm_run_bf16(h, w, w->a_state, step, false, nullptr, cs);     // the DDPM kernel decrements the step counter
        e = cudaStreamEndCapture(cs, &w->graph_src);
      }
      count_launch((int)(before - launch_count_get()));
      if (st != LDP_OK) { cudaStreamDestroy(cs); return st; }
      LDP_CUDA_OK(e);
      LDP_CUDA_OK(cudaGraphInstantiate(&w->graph, w->graph_src, 0));
      LDP_CUDA_OK(cudaStreamDestroy(cs));
    }
    for (int i = 0; i < n_steps; ++i) {
      if (w->graph) {
        LDP_CUDA_OK(cudaGraphLaunch(w->graph, s));
        count_launch((int)w->ops.size() + 1);
      } else {
        LDP_TRY(idm_run_bf16(h, w, w->a_state, step, false, nullptr, s));
      }
    }
  }
  LDP_CUDA_OK(cudaMemcpyAsync(a0_dev, w->a_state, n * 4, cudaMemcpyDeviceToDevice, s));
  return LDP_OK;
}

}  // extern "C"
