// C ABI entry points that are not tied to a network handle: error text, device check, the DDPM schedule and
// standalone scheduler ops (reference: diffusers FlaxDDPMScheduler as used at agent/ldp_agent.py:119,:471,:637-650),
// the Philox generator, and a plain dense GEMM on the tcgen05 path for tests / roofline measurement.
#include <map>
#include <mutex>

#include "net_common.h"

namespace ldp {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

// device-resident coefficient / acp tables per n_train_steps (built on first use)
struct SchedDev {
  float* coef = nullptr;
  float* acp = nullptr;
};
static std::mutex g_sched_mu;
static std::map<std::pair<int, int>, SchedDev> g_sched;   // (device, n)

static int get_sched(int n, SchedDev* out) {
  int dev = 0;
  LDP_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_sched_mu);
  auto key = std::make_pair(dev, n);
  auto it = g_sched.find(key);
  if (it == g_sched.end()) {
    std::vector<float> coef, betas, alphas, acp;
    ddpm_coef_host(n, coef);
    ddpm_schedule_host(n, betas, alphas, acp);
    SchedDev sd;
    LDP_CUDA_OK(cudaMalloc(&sd.coef, coef.size() * 4));
    LDP_CUDA_OK(cudaMalloc(&sd.acp, acp.size() * 4));
    LDP_CUDA_OK(cudaMemcpy(sd.coef, coef.data(), coef.size() * 4, cudaMemcpyHostToDevice));
    LDP_CUDA_OK(cudaMemcpy(sd.acp, acp.data(), acp.size() * 4, cudaMemcpyHostToDevice));
    it = g_sched.emplace(key, sd).first;
  }
  *out = it->second;
  return LDP_OK;
}

}  // namespace ldp

using namespace ldp;

extern "C" {

const char* ldp_last_error(void) { return get_last_error(); }
int ldp_version(void) { return 100; }

int ldp_device_check(void) {
  int dev = 0;
  LDP_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  LDP_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  LDP_CHECK(prop.major == 10, LDP_ERR_NO_DEVICE,
            std::string("libldp_b200 is built for sm_100a only; current device is sm_") + std::to_string(prop.major) +
                std::to_string(prop.minor));
  return tc_driver_check();
}

int ldp_ddpm_schedule(int n, float* betas_host, float* alphas_host, float* acp_host) {
  LDP_CHECK(n > 0 && betas_host && alphas_host && acp_host, LDP_ERR_INVALID_ARG, "bad arguments");
  std::vector<float> b, a, c;
  ddpm_schedule_host(n, b, a, c);
  for (int i = 0; i < n; ++i) {
    betas_host[i] = b[i];
    alphas_host[i] = a[i];
    acp_host[i] = c[i];
  }
  return LDP_OK;
}

int ldp_ddpm_step(int n_train_steps, int t, int sampler, const float* eps_dev, const float* x_dev, const float* noise_dev,
                  uint64_t seed, uint32_t stream_id, float* x_prev_dev, int64_t n, void* cuda_stream) {
  LDP_CHECK(n_train_steps > 0 && t >= 0 && t < n_train_steps, LDP_ERR_INVALID_ARG, "timestep outside [0, n_train_steps)");
  LDP_CHECK(eps_dev && x_dev && x_prev_dev && n > 0, LDP_ERR_INVALID_ARG, "bad arguments");
  LDP_CHECK(sampler == LDP_SAMPLER_DDPM || sampler == LDP_SAMPLER_DDIM, LDP_ERR_INVALID_ARG, "unknown sampler");
  SchedDev sd;
  LDP_TRY(get_sched(n_train_steps, &sd));
  DdpmStep d;
  d.coef = sd.coef;
  d.step.scalar = t;
  d.eps = eps_dev; d.x = x_dev; d.out = x_prev_dev;
  d.call.noise = noise_dev;
  d.call.noise_step_stride = 0;
  d.call.n_steps = t + 1;            // (n_steps-1-t) == 0 -> noise_dev itself
  d.call.sampler = sampler;
  d.call.seed = seed;
  d.call.stream_id = stream_id;
  d.n = n;
  return launch_ddpm_step(d, (cudaStream_t)cuda_stream);
}

int ldp_ddpm_add_noise(int n_train_steps, const float* x0_dev, const float* noise_dev, const int32_t* t_dev,
                       float* out_dev, int64_t rows, int64_t row_len, void* cuda_stream) {
  LDP_CHECK(n_train_steps > 0 && x0_dev && noise_dev && t_dev && out_dev && rows > 0 && row_len > 0, LDP_ERR_INVALID_ARG,
            "bad arguments");
  SchedDev sd;
  LDP_TRY(get_sched(n_train_steps, &sd));
  return launch_add_noise(sd.acp, x0_dev, noise_dev, t_dev, out_dev, rows, row_len, (cudaStream_t)cuda_stream);
}

int ldp_philox_normal(uint64_t seed, uint32_t stream_id, uint32_t step, float* out_dev, int64_t n, void* cuda_stream) {
  LDP_CHECK(out_dev && n > 0, LDP_ERR_INVALID_ARG, "bad arguments");
  return launch_philox_normal(seed, stream_id, step, out_dev, n, (cudaStream_t)cuda_stream);
}

int ldp_philox_normal_rows(uint64_t seed, uint32_t stream_id, uint32_t step, int64_t row0, int64_t rows, int row_len,
                           float* out_dev, void* cuda_stream) {
  LDP_CHECK(out_dev && rows > 0 && row_len > 0 && row0 >= 0, LDP_ERR_INVALID_ARG, "bad arguments");
  return launch_philox_normal_rows(seed, stream_id, step, row0, rows, row_len, out_dev, (cudaStream_t)cuda_stream);
}

int ldp_jax_random(const uint32_t* keys_dev, int n_keys, int64_t n, int mode, void* out_dev, void* cuda_stream) {
  LDP_CHECK(mode == 0 || mode == 1, LDP_ERR_INVALID_ARG, "mode must be 0 (bits) or 1 (normal)");
  return launch_jax_random(keys_dev, n_keys, n, mode, out_dev, (cudaStream_t)cuda_stream);
}

int ldp_tc_geometry(int M, int N, int block_n, int epilogue, int pair, int n_acc, int32_t* out) {
  LDP_CHECK(out != nullptr && M > 0 && N > 0, LDP_ERR_INVALID_ARG, "bad arguments");
  LDP_CHECK(block_n == 64 || block_n == 128 || block_n == 256, LDP_ERR_INVALID_ARG, "block_n must be 64, 128 or 256");
  LDP_CHECK(epilogue >= TC_EPI_PLAIN && epilogue <= TC_EPI_LN && n_acc >= 1 && n_acc <= 5, LDP_ERR_INVALID_ARG, "bad epilogue / n_acc");
  TcGemm p;
  p.M = M; p.N = N; p.block_n = block_n; p.mode = epilogue; p.pair = pair ? 1 : 0; p.n_acc = n_acc; p.w_max = n_acc;
  LDP_TRY(tc_gemm_geometry(&p));
  const int32_t v[8] = {p.tiles_m, p.tiles_n, p.grid_ctas, p.persistent, p.acc_bufs, p.tmem_cols, p.n_tail, p.acc_stride};
  for (int i = 0; i < 8; ++i) out[i] = v[i];
  return LDP_OK;
}

int ldp_tc_dense(const float* a_dev, const float* w_host, const float* bias_host, float* c_dev, int M, int K, int N,
                 void* cuda_stream) {
  LDP_CHECK(a_dev && w_host && c_dev && M > 0 && K > 0 && N > 0, LDP_ERR_INVALID_ARG, "bad arguments");
  LDP_TRY(tc_driver_check());
  LDP_TRY(tc_gemm_init());
  cudaStream_t s = (cudaStream_t)cuda_stream;
  Arena tmp;
  const int kp = round_up(K, 64), n_pad = round_up(N, 128), lda = round_up(K, 8);
  float *w_dev, *bias_dev = nullptr;
  __nv_bfloat16 *wt, *a_bf;
  int32_t* map_dev;
  TcKBlock* kb_dev;
  LDP_TRY(tmp.alloc_t(&w_dev, (size_t)K * N, false));
  LDP_TRY(tmp.alloc_t(&wt, (size_t)n_pad * kp));
  LDP_TRY(tmp.alloc_t(&a_bf, (size_t)M * lda));
  LDP_TRY(tmp.alloc_t(&map_dev, kp));
  LDP_TRY(tmp.alloc_t(&kb_dev, kp / 64));
  LDP_CUDA_OK(cudaMemcpy(w_dev, w_host, (size_t)K * N * 4, cudaMemcpyHostToDevice));
  if (bias_host) {
    LDP_TRY(tmp.alloc_t(&bias_dev, N));
    LDP_CUDA_OK(cudaMemcpy(bias_dev, bias_host, (size_t)N * 4, cudaMemcpyHostToDevice));
  }
  std::vector<int32_t> kmap(kp);
  std::vector<TcKBlock> kb(kp / 64);
  for (int k = 0; k < kp; ++k) kmap[k] = k < K ? k : -1;
  for (int i = 0; i < kp / 64; ++i) kb[i] = make_stage(0, 0, 1, i * 64, 0, 0, i);
  LDP_CUDA_OK(cudaMemcpy(map_dev, kmap.data(), (size_t)kp * 4, cudaMemcpyHostToDevice));
  LDP_CUDA_OK(cudaMemcpy(kb_dev, kb.data(), kb.size() * sizeof(TcKBlock), cudaMemcpyHostToDevice));
  LDP_TRY(launch_pack_wt_bf16(w_dev, N, N, map_dev, kp, wt, kp, 0, n_pad, s));
  LDP_TRY(launch_cast_bf16(a_dev, K, a_bf, lda, M, K, 0, s));
  TcGemm op;
  uint64_t ad[4] = {(uint64_t)K, 1, 1, (uint64_t)M};
  uint64_t as[3] = {(uint64_t)lda * 2, (uint64_t)lda * 2, (uint64_t)lda * 2};
  uint32_t ab[4] = {64, 1, 1, 128};
  LDP_TRY(make_tmap_bf16(&op.map_a[0], a_bf, 4, ad, as, ab));
  for (int i = 1; i < 4; ++i) op.map_a[i] = op.map_a[0];
  uint64_t bd[2] = {(uint64_t)kp, (uint64_t)n_pad};
  uint64_t bs[1] = {(uint64_t)kp * 2};
  uint32_t bb[2] = {64, 128};
  LDP_TRY(make_tmap_bf16(&op.map_b, wt, 2, bd, bs, bb));
  TcRun run;
  run.src_acc = kb[0].src_acc; run.c0 = 0; run.d12 = 0; run.wk = 0; run.count = kp / 64;
  TcRun* run_dev;
  LDP_TRY(tmp.alloc_t(&run_dev, 1));
  LDP_CUDA_OK(cudaMemcpy(run_dev, &run, sizeof(run), cudaMemcpyHostToDevice));
  op.kb = kb_dev; op.num_kb = kp / 64; op.runs = run_dev; op.num_runs = 1; tc_set_inline_runs(&op, &run, 1); op.M = M; op.N = N; op.block_n = 128;
  op.items_per_tile = 128; op.rows_per_item = 1;
  op.mode = TC_EPI_PLAIN; op.bias = bias_dev; op.out_f32 = c_dev; op.ld_out_f32 = N;
  LDP_TRY(launch_tc_gemm(op, s));
  LDP_CUDA_OK(cudaStreamSynchronize(s));    // temporaries die with `tmp`
  return LDP_OK;
}

int64_t ldp_launch_count(void) { return launch_count_get(); }
void ldp_launch_count_reset(void) { launch_count_reset(); }

}  // extern "C"
