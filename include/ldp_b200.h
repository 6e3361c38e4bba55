/* libldp_b200 - C ABI of the B200-native LDP hot path.
 *
 * The reference (amberxie88/latent_diffusion_planning) is pure Python/JAX and has no FFI layer; its seam for
 * this path is a handful of Python calls on Flax pytrees.  Each entry point below replaces one of those calls
 * (reference file:line cited per function).  INTEGRATION.md shows the ctypes stub a maintainer would add on the
 * reference side.
 *
 * Conventions
 *  - plain C: pointers + sizes only, no torch / CUDA C++ types (streams travel as void*).
 *  - every function returns an int status (LDP_OK == 0, negative on error); ldp_last_error() returns the message
 *    of the last failure on the calling thread.  Nothing throws across this boundary.
 *  - "dev" pointers are CUDA device pointers owned by the caller; "host" pointers are host memory.
 *    The library owns only what lives inside its opaque handles (packed weights, tables, workspace, CUDA graphs).
 *  - all calls are stream-ordered on the given stream and never call cudaDeviceSynchronize(); a handle must be
 *    used by one stream at a time.
 *  - tensors are row-major, channels-last, exactly as in the reference: trajectories (B,T,D), images NHWC.
 *  - weights arrive as ONE float32 host blob: the network's tensors in canonical order, each in Flax layout
 *    (Dense kernel (in,out); Conv kernel (k..,in,out)).  The canonical order is documented per create call and
 *    implemented by latent_diffusion_planning_b200/params.py (unet_spec / idm_spec / vae_encoder_spec).
 */
#ifndef LDP_B200_H_
#define LDP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDP_API __attribute__((visibility("default")))

enum {
  LDP_OK = 0,
  LDP_ERR_INVALID_ARG = -1,   /* null pointer, bad enum, non-positive size */
  LDP_ERR_BAD_SHAPE = -2,     /* shape inconsistent with the handle's config */
  LDP_ERR_UNSUPPORTED = -3,   /* valid in the reference but not supported by this path (see message) */
  LDP_ERR_CUDA = -4,          /* CUDA runtime/driver error; message has the call and the CUDA error string */
  LDP_ERR_PARAM_COUNT = -5,   /* weight blob length does not match the config's canonical spec */
  LDP_ERR_NO_DEVICE = -6      /* no sm_100 device / driver entry point missing */
};

/* compute precision of a call */
enum {
  LDP_PREC_FP32 = 0,  /* fp32 SIMT path: every contraction in fp32 FFMA (parity gate 1e-5) */
  LDP_PREC_BF16 = 1   /* tcgen05 path: bf16 operands, fp32 accumulate in TMEM, fp32 norms/epilogues (gate 1e-2) */
};

/* sampler of the reverse loop */
enum {
  LDP_SAMPLER_DDPM = 0,  /* diffusers FlaxDDPMScheduler.step semantics (reference agent/ldp_agent.py:471,:498) */
  LDP_SAMPLER_DDIM = 1   /* eta = 0; NOT in the reference (BASELINE config #5) - spec in DESIGN.md */
};

LDP_API const char* ldp_last_error(void);
LDP_API int ldp_version(void);
/* Device sanity: returns LDP_OK iff the current device is sm_100 and the TMA driver entry point resolves. */
LDP_API int ldp_device_check(void);

/* ---------------------------------------------------------------------------------------------------------
 * DDPM schedule  -  replaces FlaxDDPMScheduler(...).create_state()  (reference agent/ldp_agent.py:637-650)
 * beta_schedule 'squaredcos_cap_v2', clip_sample=True, prediction_type 'epsilon', variance 'fixed_small'.
 * Host-side: fills three float32 arrays of length n (betas, alphas, alphas_cumprod).
 * --------------------------------------------------------------------------------------------------------- */
LDP_API int ldp_ddpm_schedule(int n_train_steps, float* betas_host, float* alphas_host, float* alphas_cumprod_host);

/* scheduler.step(state, model_output, t, sample, key).prev_sample   (agent/ldp_agent.py:471, :498)
 * x_prev = c0(t) clip((x - sqrt(1-acp_t) eps)/sqrt(acp_t), -1, 1) + ct(t) x + [t>0] sigma_t z
 * noise_dev: n floats of N(0,1) (the injected stand-in for jax.random inside step), or NULL to draw
 * Philox4x32-10 normals with (seed, stream, step=t, element index).  In-place (x_prev_dev == x_dev) allowed. */
LDP_API int ldp_ddpm_step(int n_train_steps, int t, int sampler, const float* eps_dev, const float* x_dev,
                          const float* noise_dev, uint64_t seed, uint32_t stream_id, float* x_prev_dev,
                          int64_t n, void* cuda_stream);

/* scheduler.add_noise(state, x0, noise, t)   (agent/ldp_agent.py:119, :136)
 * rows x row_len elements; t_dev holds one int32 timestep per row (t broadcast from the left). */
LDP_API int ldp_ddpm_add_noise(int n_train_steps, const float* x0_dev, const float* noise_dev, const int32_t* t_dev,
                               float* out_dev, int64_t rows, int64_t row_len, void* cuda_stream);

/* Philox normals (the generator ldp_ddpm_step uses when noise_dev == NULL: flat element index), exposed for tests. */
LDP_API int ldp_philox_normal(uint64_t seed, uint32_t stream_id, uint32_t step, float* out_dev, int64_t n,
                              void* cuda_stream);
/* The row-structured normals the fused sampling loops (ldp_planner_sample / ldp_idm_sample) draw at reverse step
 * `step` when noise_dev == NULL: out (rows,row_len); element (r,c) is component c&3 of the Philox4x32-10 quad with
 * counter (c>>2, row0+r, step, stream_id) and key seed, through a Box-Muller on MUFU approximations.  Keyed by the
 * GLOBAL row, so a batch sharded over ranks (row_offset) draws exactly what the unsharded batch draws. */
LDP_API int ldp_philox_normal_rows(uint64_t seed, uint32_t stream_id, uint32_t step, int64_t row0, int64_t rows,
                                   int row_len, float* out_dev, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------------
 * Planner score network  -  ConditionalUnet1D  (reference networks/diffusion_nets_v2.py:104-169)
 * Canonical weight order (params.unet_spec): time-MLP Dense_0, Dense_1 {kernel,bias}; then for each
 * ConditionalResidualBlock1D_i in creation order: conv1 {kernel,bias}, gn1 {scale,bias}, FiLM Dense {kernel,bias},
 * conv2 {kernel,bias}, gn2 {scale,bias}, [residual 1x1 conv {kernel,bias}]; Downsample1d_i; Upsample1d_i;
 * final Conv1dBlock {conv, gn}; final 1x1 Conv.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct LdpUnetConfig {
  int32_t input_dim;         /* D  (planner.input_dim = obs_dim, agent/ldp_agent.py:569) */
  int32_t global_cond_dim;   /* Dc = obs_horizon * D (agent/ldp_agent.py:570,573-574) */
  int32_t step_embed_dim;    /* 256 */
  int32_t n_levels;          /* len(down_dims) <= 6 */
  int32_t down_dims[6];      /* [256,512,1024] */
  int32_t kernel_size;       /* 5 */
  int32_t n_groups;          /* 8 */
  int32_t n_train_steps;     /* planner_n_diffusion_steps = 100 (time/FiLM tables are precomputed for t in [0,n)) */
} LdpUnetConfig;

typedef struct LdpPlanner LdpPlanner;

LDP_API int ldp_planner_create(const LdpUnetConfig* cfg, const float* params_host, uint64_t n_params,
                               LdpPlanner** out);
LDP_API int ldp_planner_destroy(LdpPlanner* h);
/* Number of float32 values the config's canonical blob must hold. */
LDP_API int64_t ldp_unet_param_count(const LdpUnetConfig* cfg);

/* planner_state.apply_fn({"params": p}, sample, timestep, global_cond)   (agent/ldp_agent.py:123, :470)
 * sample_dev (B,T,D) f32, cond_dev (B,Dc) f32 -> eps_dev (B,T,D) f32.
 * timesteps_dev: B int32 per-row timesteps (training call, :116-123) or NULL to use the scalar `timestep`. */
LDP_API int ldp_unet_forward(LdpPlanner* h, int precision, const float* sample_dev, const int32_t* timesteps_dev,
                             int timestep, const float* cond_dev, int B, int T, float* eps_dev, void* cuda_stream);

/* HOT LOOP 1  -  the fori_loop at agent/ldp_agent.py:465-476:
 *   for i in 0..n_steps-1: k = n_steps-1-i; eps = UNet(x,k,cond); x = step(eps,k,x,z_i)
 * x_T_dev (B,T,D) start noise; noise_dev (n_steps,B,T,D) injected z_i or NULL (Philox: seed, stream 0, step k,
 * keyed by global row (row_offset*T + local row) and column quad - see ldp_philox_normal_rows - so that a batch
 * sharded over ranks draws the same numbers as unsharded).
 * The per-step eps-predict -> x0 -> clip -> posterior mean -> add-noise update runs in the epilogue of the UNet's
 * last GEMM (bf16 path); the 100 steps replay a captured CUDA graph.  x0_dev (B,T,D) may alias x_T_dev. */
LDP_API int ldp_planner_sample(LdpPlanner* h, int precision, int sampler, const float* x_T_dev, const float* cond_dev,
                               const float* noise_dev, uint64_t seed, int64_t row_offset, int B, int T, int n_steps,
                               float* x0_dev, void* cuda_stream);

/* Diagnostics for per-layer parity (SURVEY.md 8c protocol; no reference counterpart - in the reference these are the
 * values of `x` between the statements of ConditionalUnet1D.__call__, networks/diffusion_nets_v2.py:137-167): after a
 * bf16 ldp_unet_forward at (B, T), copies one intermediate activation to out_dev as float32 (rows, cols) row-major.
 * tap_id: i in [0, n_blocks) = output of ConditionalResidualBlock1D_i; 100 + l = Downsample1d_l; 200 + u =
 * Upsample1d_u; 300 = the final Conv1dBlock. */
LDP_API int ldp_planner_read_activation(LdpPlanner* h, int B, int T, int tap_id, float* out_dev, int64_t max_elems,
                                        int32_t* rows_out, int32_t* cols_out, void* cuda_stream);

/* Diagnostics (no reference counterpart): times every kernel of one bf16 denoising step in isolation - `reps`
 * back-to-back launches between two CUDA events per kernel.  us_host[i] = microseconds per launch of kernel i;
 * meta_host[4i..4i+3] = {M, N, K/64, block_n | epilogue << 16 | aux << 24 | accumulators << 25}.  phases_host (may
 * be NULL): [8i] = CTAs, [8i+1..8i+7] = mean SM-clock cycles from kernel entry to {prologue done, dependency
 * resolved, first operands landed, last MMA issued, accumulators complete, epilogue done, exit}.  *n_ops = kernels. */
LDP_API int ldp_planner_profile_step(LdpPlanner* h, int B, int T, int reps, float* us_host, int32_t* meta_host,
                                     float* phases_host, int max_ops, int* n_ops, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------------
 * Inverse-dynamics score network  -  MLPDiffusion(FourierFeatures -> MLP -> MLPResNet)
 * (reference networks/mlp_diffusion_nets.py:8-68, networks/diffusion.py:7-22, networks/mlp_nets.py:49-97)
 * Canonical weight order (params.idm_spec): cond MLP Dense_i {kernel,bias}; MLPResNet Dense_0; per block
 * LayerNorm {scale,bias}, Dense_0, Dense_1; MLPResNet Dense_1.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct LdpIdmConfig {
  int32_t obs_dim;        /* D: s||s' has 2D columns */
  int32_t action_dim;     /* A */
  int32_t hidden_dim;     /* 256 */
  int32_t n_blocks;       /* 3 */
  int32_t time_dim;       /* 256 (FourierFeatures output_size) */
  int32_t n_cond_layers;  /* 2 */
  int32_t cond_hidden[4]; /* [256,256] */
  int32_t n_train_steps;  /* idm_n_diffusion_steps = 100 */
} LdpIdmConfig;

typedef struct LdpIdm LdpIdm;

LDP_API int ldp_idm_create(const LdpIdmConfig* cfg, const float* params_host, uint64_t n_params, LdpIdm** out);
LDP_API int ldp_idm_destroy(LdpIdm* h);
LDP_API int64_t ldp_idm_param_count(const LdpIdmConfig* cfg);

/* idm_state.apply_fn({"params": p}, s_sprime, noisy_action, t)   (agent/ldp_agent.py:137, :380, :497)
 * s_dev (N,2D), a_dev (N,A) -> eps_dev (N,A); timesteps_dev N int32 or NULL + scalar. */
LDP_API int ldp_idm_forward(LdpIdm* h, int precision, const float* s_dev, const float* a_dev,
                            const int32_t* timesteps_dev, int timestep, int N, float* eps_dev, void* cuda_stream);

/* HOT LOOP 2  -  agent/ldp_agent.py:492-503 (same code at :375-386, :416-427).  noise_dev (n_steps,N,A) or NULL
 * (Philox stream 1).  a0_dev may alias a_T_dev. */
LDP_API int ldp_idm_sample(LdpIdm* h, int precision, int sampler, const float* s_dev, const float* a_T_dev,
                           const float* noise_dev, uint64_t seed, int64_t row_offset, int N, int n_steps,
                           float* a0_dev, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------------
 * VAE encoder  -  FlaxAutoencoderKL.encode(x).latent_dist.mean   (agent/ldp_agent.py:59, process_sdvae_data.py:70-73)
 * Canonical weight order: params.vae_encoder_spec.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct LdpVaeConfig {
  int32_t in_channels;         /* 3 */
  int32_t latent_channels;     /* 4 */
  int32_t n_blocks;            /* len(block_out_channels) <= 8 */
  int32_t block_out_channels[8];
  int32_t layers_per_block;    /* 2 */
  int32_t norm_num_groups;     /* 32 */
  int32_t image_size;          /* 64 (square) */
} LdpVaeConfig;

typedef struct LdpVae LdpVae;

LDP_API int ldp_vae_create(const LdpVaeConfig* cfg, const float* params_host, uint64_t n_params, LdpVae** out);
LDP_API int ldp_vae_destroy(LdpVae* h);
LDP_API int64_t ldp_vae_param_count(const LdpVaeConfig* cfg);

/* images_dev: (B,S,S,3) NHWC.  pixel_format 0: uint8 0..255 (normalised in-kernel as x/255*2-1, utils/data_utils.py:11
 * with min 0 / max 255; == process_sdvae_data.py:89-90); 1: float32 already in [-1,1].
 * latent_dev: (B,h,w,latent_channels) float32 = latent_dist.mean, then, if lat_max > lat_min, the reference's
 * normalize_obs (z-min)/(max-min)*2-1 (agent/ldp_agent.py:62) is fused into the last kernel. */
LDP_API int ldp_vae_encode(LdpVae* h, int precision, const void* images_dev, int pixel_format, int B,
                           float lat_min, float lat_max, float* latent_dev, void* cuda_stream);

/* Decoder: FlaxAutoencoderKL.decode(z).sample (reference agent/ldp_agent.py:66-85 `vae_decode`, the plan_viz of
 * sample_viz :483).  Same config struct (in_channels = channels of the decoded image); weights in the order of
 * params.py:vae_decoder_spec (post_quant_conv, decoder/conv_in, mid block, up blocks, conv_norm_out, conv_out).
 * latent_dev: (B,h,w,latent_channels) float32, already un-normalised (the caller applies unnormalize_obs :82);
 * images_dev: (B,S,S,in_channels) float32 NHWC (the reference views the same values as NCHW). */
LDP_API int64_t ldp_vae_decoder_param_count(const LdpVaeConfig* cfg);
LDP_API int ldp_vae_decoder_create(const LdpVaeConfig* cfg, const float* params_host, uint64_t n_params, LdpVae** out);
LDP_API int ldp_vae_decode(LdpVae* h, int precision, const float* latent_dev, int B, float* images_dev, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------------
 * Low-level operator exposed for tests and roofline measurement: C[M,N] = A[M,K] W[K,N] + bias on the
 * tcgen05 path (bf16 operands, fp32 accumulate).  a_dev (M,K) f32, w_host (K,N) f32 Flax Dense layout,
 * bias_host (N) or NULL, c_dev (M,N) f32.  Synchronous w.r.t. weight packing; compute is stream-ordered. */
LDP_API int ldp_tc_dense(const float* a_dev, const float* w_host, const float* bias_host, float* c_dev, int M, int K,
                         int N, void* cuda_stream);

/* Host-only: the tile / grid / tensor-memory geometry a tcgen05 launch of this shape gets (no device needed; the CPU test suite checks the
 * decisions for the benchmarked shapes).  epilogue: 0 plain, 1 GroupNorm, 2 DDPM, 3 LayerNorm; pair: 1 = cta_group::2 CTA pairs;
 * n_acc: tap accumulators (1..5).  out[8] = {tiles_m, tiles_n, grid_ctas, persistent, acc_bufs, tmem_cols, n_tail, acc_stride}. */
LDP_API int ldp_tc_geometry(int M, int N, int block_n, int epilogue, int pair, int n_acc, int32_t* out);

/* Counters: number of kernels this library launched on the calling thread since the last reset (bench.py's
 * gpu_launches; graph replays count the kernels inside the graph). */
LDP_API int64_t ldp_launch_count(void);
LDP_API void ldp_launch_count_reset(void);

/* jax.random on the device (jax 0.4.26 default threefry2x32 generator; row N4's optional RNG-stream compatibility):
 * for each of n_keys raw keys (keys_dev: [n_keys][2] uint32) the n values `jax.random.bits(key, (n,), uint32)` (mode 0)
 * or `jax.random.normal(key, (n,), float32)` (mode 1) -> out_dev [n_keys][n].  With the key threading of
 * agent/ldp_agent.py:461-503 (restated in latent_diffusion_planning_b200/jax_random.py) the sampling loops can be fed
 * the noise the reference would draw from the same PRNGKey.  Bits are exact; normal differs from XLA's erf_inv by ~1 ulp. */
LDP_API int ldp_jax_random(const uint32_t* keys_dev, int n_keys, int64_t n, int mode, void* out_dev, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------------
 * Training objective (next-row N1)  -  LDPAgent.update (agent/ldp_agent.py:229-277)
 * The caller owns flat float32 device buffers in the canonical spec order (params.unet_spec / params.idm_spec):
 * parameters, gradients (same layout; the functions ADD into them, zero them first) and the two Adam moments.
 * The data-parallel exchange is therefore one all-reduce over the gradient buffer (reference: GSPMD mean over the
 * global batch, train_bc.py:73).  The trainer handle owns activations only.
 * precision LDP_PREC_FP32: every contraction in fp32 FFMA (the reference trains in float32; parity gate).
 * precision LDP_PREC_BF16: forward, data-gradient and weight-gradient contractions on tcgen05 with bf16 operands and
 * fp32 accumulation (BASELINE config #4 "bf16"); parameters, gradients, norms, activations and Adam stay fp32.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct LdpTrainer LdpTrainer;

LDP_API int ldp_unet_trainer_create(const LdpUnetConfig* cfg, LdpTrainer** out);
LDP_API int ldp_idm_trainer_create(const LdpIdmConfig* cfg, LdpTrainer** out);
LDP_API int ldp_trainer_destroy(LdpTrainer* h);

/* planner_loss (agent/ldp_agent.py:113-126): noisy = add_noise(x0, noise, t); eps = UNet(noisy, t, cond);
 * loss = mean((eps - noise)^2).  x0_dev, noise_dev (B,T,D); t_dev (B,) int32; cond_dev (B,Dc).
 * *loss_dev += loss;  grads_dev += loss_weight * d loss / d params. */
LDP_API int ldp_unet_loss_grad(LdpTrainer* h, int precision, const float* params_dev, float* grads_dev, const float* x0_dev,
                               const float* noise_dev, const int32_t* t_dev, const float* cond_dev, int B, int T,
                               float loss_weight, float* loss_dev, void* cuda_stream);

/* idm_loss (agent/ldp_agent.py:128-139): s_dev (N,2D) = [s | s'], a0_dev, noise_dev (N,A), t_dev (N,) int32. */
LDP_API int ldp_idm_loss_grad(LdpTrainer* h, int precision, const float* params_dev, float* grads_dev, const float* s_dev,
                              const float* a0_dev, const float* noise_dev, const int32_t* t_dev, int N,
                              float loss_weight, float* loss_dev, void* cuda_stream);

/* Data-parallel exchange of `update` (reference train_bc.py:70-78: the global batch is sharded over devices and jax.grad's
 * result is averaged by GSPMD; here every rank holds an equal shard and all-reduces the flat gradient buffer).  The backward
 * pass finishes the gradient buffer range by range, in reverse layer order: ldp_trainer_grad_buckets returns those ranges
 * (float offsets / lengths into grads_dev) for the LAST ldp_*_loss_grad call on this handle, in completion order;
 * has_event[i] = 1: the range is complete as soon as the event the step records for it has fired - ldp_trainer_wait_bucket
 * makes `cuda_stream` wait for exactly that (cudaStreamWaitEvent), so a communication stream can start the bucket's
 * all-reduce while the rest of the backward pass still runs; has_event[i] = 0: complete when the call itself is. */
LDP_API int ldp_trainer_grad_buckets(LdpTrainer* h, int64_t* offsets, int64_t* lengths, int32_t* has_event, int max_buckets,
                                     int* n_buckets);
LDP_API int ldp_trainer_wait_bucket(LdpTrainer* h, int bucket, void* cuda_stream);

/* optax.adam step (agent/ldp_agent.py:580-600; optax 0.2.2 scale_by_adam, eps outside the square root, eps_root 0):
 * g = grads * grad_scale; mu = b1 mu + (1-b1) g; nu = b2 nu + (1-b2) g^2;
 * params -= lr * (mu / (1-b1^count)) / (sqrt(nu / (1-b2^count)) + eps);  count is 1-based. */
LDP_API int ldp_adam_update(float* params_dev, const float* grads_dev, float* mu_dev, float* nu_dev, uint64_t n,
                            float lr, float b1, float b2, float eps, int64_t count, float grad_scale,
                            void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* LDP_B200_H_ */
