// fp32 SIMT kernels: the exact-arithmetic path (parity gate 1e-5 vs the fp64 oracle) and the small
// elementwise / table / packing kernels both precisions share.
#include <cuda_fp16.h>

#include "kernels.h"

namespace ldp {

static thread_local long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
long long launch_count_get() { return g_launches; }
void launch_count_reset() { g_launches = 0; }

#define LDP_LAUNCH_OK()                                                                                     \
  do {                                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                                    \
    if (_e != cudaSuccess) {                                                                                \
      set_last_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                     std::to_string(__LINE__) + ")");                                                       \
      return LDP_ERR_CUDA;                                                                                  \
    }                                                                                                       \
    count_launch();                                                                                         \
  } while (0)

// ------------------------------------------------------------------------------------------------
// implicit-GEMM, fp32 FFMA.  64x64 tile, K step 16, 256 threads, 4x4 register tile per thread.
// ------------------------------------------------------------------------------------------------
constexpr int GM = 64, GN = 64, GK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmF32 p) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int ctot = p.c1 + p.c2;
  const int K = p.taps * ctot;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int r = idx / GK, kk = idx % GK;
      int m = m0 + r, k = k0 + kk;
      float v = 0.f;
      if (m < p.m && k < K) {
        int j = k / ctot, c = k - j * ctot;
        int b = m / p.t_out, t = m - b * p.t_out;
        int num = t * p.stride + j - p.pad;
        if (num >= 0 && (num % p.dil) == 0) {
          int ti = num / p.dil;
          if (ti < p.t_in) {
            long long row = (long long)b * p.t_in + ti;
            v = (c < p.c1) ? p.x1[row * p.ld1 + c] : p.x2[row * p.ld2 + (c - p.c1)];
            if (p.a_act == 1) v = mish_f<false>(v);
          }
        }
      }
      As[kk][r] = v;
    }
    if (p.w_mode == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int idx = tid + i * 256;
        int kk = idx / GN, nn = idx % GN;
        int k = k0 + kk, n = n0 + nn;
        Bs[kk][nn] = (k < K && n < p.n) ? p.w[(long long)k * p.ldw + n] : 0.f;
      }
    } else {
      // data-gradient view of a forward kernel W[taps][w_ctot][ldw]: B(k = (j', co), n) = W[taps-1-j'][w_coff+n][co]
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int idx = tid + i * 256;
        int nn = idx / GK, kk = idx % GK;
        int k = k0 + kk, n = n0 + nn;
        float v = 0.f;
        if (k < K && n < p.n) {
          int j = k / ctot, co = k - j * ctot;
          v = p.w[((long long)(p.taps - 1 - j) * p.w_ctot + p.w_coff + n) * p.ldw + co];
        }
        Bs[kk][nn] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.m) continue;
    const float* tabrow = p.tab ? p.tab + (long long)step_of(p.step, m) * p.ld_tab : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.n) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (tabrow) v += tabrow[n];
      if (p.act == 1) v = fmaxf(v, 0.f);
      if (p.res) v += p.res[(long long)m * p.ldres + n];
      p.out[(long long)m * p.ldo + n] = v;
    }
  }
}

int launch_gemm_f32(const GemmF32& p, cudaStream_t s) {
  LDP_CHECK(p.x1 && p.w && p.out && p.m > 0 && p.n > 0, LDP_ERR_INVALID_ARG, "gemm_f32: bad arguments");
  dim3 grid(ceil_div(p.n, GN), ceil_div(p.m, GM));
  gemm_f32_kernel<<<grid, 256, 0, s>>>(p);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (+Mish/SiLU) (+FiLM) (+residual), fp32.  One block per (sample, group).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) sh[0] = t;
  }
  __syncthreads();
  return sh[0];
}

__global__ void __launch_bounds__(256) groupnorm_f32_kernel(const GroupNormF32 p) {
  __shared__ float sh[32];
  const int b = blockIdx.x / p.G, g = blockIdx.x % p.G;
  const int gw = p.C / p.G;
  const int cnt = p.P * gw;
  const float* xb = p.x + (long long)b * p.P * p.ldx + g * gw;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int pos = i / gw, c = i - pos * gw;
    float v = xb[(long long)pos * p.ldx + c];
    s += v;
    ss += v * v;
  }
  s = block_sum(s, sh);
  ss = block_sum(ss, sh);
  const float mean = s / cnt;
  const float var = fmaxf(ss / cnt - mean * mean, 0.f);
  const float rstd = rsqrtf(var + p.eps);
  const float* trow = nullptr;
  const float* orow = nullptr;
  if (p.film) {
    trow = p.ttab ? p.ttab + (long long)step_of(p.step, b) * p.ld_ttab + p.film_off : nullptr;
    orow = p.otab + (long long)b * p.ld_otab + p.film_off;
  }
  float* yb = p.y + (long long)b * p.P * p.ldy + g * gw;
  const float* rb = p.res ? p.res + (long long)b * p.P * p.ldres + g * gw : nullptr;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int pos = i / gw, c = i - pos * gw;
    int ch = g * gw + c;
    float v = (xb[(long long)pos * p.ldx + c] - mean) * rstd * p.gamma[ch] + p.beta[ch];
    if (p.act == 1) v = mish_f<false>(v);
    else if (p.act == 2) v = v / (1.f + expf(-v));
    if (p.film) v = ((trow ? trow[ch] : 0.f) + orow[ch]) * v + ((trow ? trow[p.C + ch] : 0.f) + orow[p.C + ch]);
    if (rb) v += rb[(long long)pos * p.ldres + c];
    yb[(long long)pos * p.ldy + c] = v;
  }
}

int launch_groupnorm_f32(const GroupNormF32& p, cudaStream_t s) {
  LDP_CHECK(p.x && p.y && p.B > 0 && p.G > 0 && p.C % p.G == 0, LDP_ERR_INVALID_ARG, "groupnorm_f32: bad arguments");
  groupnorm_f32_kernel<<<p.B * p.G, 256, 0, s>>>(p);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// LayerNorm over the last dim, one warp per row (Flax nn.LayerNorm: eps 1e-6, fast variance).
__global__ void layernorm_f32_kernel(const float* x, float* y, int rows, int C, const float* gamma, const float* beta,
                                     float eps, int relu_instead) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * C;
  float* yr = y + (long long)row * C;
  if (relu_instead) {
    for (int c = lane; c < C; c += 32) yr[c] = fmaxf(xr[c], 0.f);
    return;
  }
  float s = 0.f, ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    float v = xr[c];
    s += v;
    ss += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  float mean = s / C, var = fmaxf(ss / C - mean * mean, 0.f), rstd = rsqrtf(var + eps);
  for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}

int launch_layernorm_f32(const float* x, float* y, int rows, int C, const float* gamma, const float* beta, float eps,
                         int relu_instead, cudaStream_t s) {
  layernorm_f32_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(x, y, rows, C, gamma, beta, eps, relu_instead);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// ------------------------------------------------------------------------------------------------
// sinusoid tables: f_j = exp(-j ln(10000)/(half-1)) in fp32 as the reference does
// (networks/diffusion_nets_v2.py:25-30, networks/diffusion.py:17-22)
// ------------------------------------------------------------------------------------------------
__global__ void sinusoid_table_kernel(float* out, int n_steps, int dim, int cos_first) {
  int half = dim / 2;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_steps * half) return;
  int k = idx / half, j = idx % half;
  float scale = -(logf(10000.f) / (float)(half - 1));
  float f = expf((float)j * scale);
  float arg = (float)k * f;
  float sv = sinf(arg), cv = cosf(arg);
  float* row = out + (long long)k * dim;
  if (cos_first) { row[j] = cv; row[half + j] = sv; }
  else           { row[j] = sv; row[half + j] = cv; }
}

int launch_sinusoid_table(float* out, int n_steps, int dim, int cos_first, cudaStream_t s) {
  int n = n_steps * (dim / 2);
  sinusoid_table_kernel<<<ceil_div(n, 256), 256, 0, s>>>(out, n_steps, dim, cos_first);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// ------------------------------------------------------------------------------------------------
// scheduler kernels
// ------------------------------------------------------------------------------------------------
__global__ void ddpm_step_kernel(const DdpmStep p) {
  const int t = step_of(p.step, 0);
  const float* cf = p.coef + t * 8;
  const float inv_sa = cf[0], s1a = cf[1], c0 = cf[2], ct = cf[3], sigma = cf[4], sap = cf[5], s1ap = cf[6];
  const DdpmCall call = p.call_dev ? *p.call_dev : p.call;
  const float* noise = call.noise ? call.noise + (long long)(call.n_steps - 1 - t) * call.noise_step_stride : nullptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    float e = p.eps[i], x = p.x[i];
    float x0 = fminf(fmaxf((x - s1a * e) * inv_sa, -1.f), 1.f);
    float o;
    if (call.sampler == LDP_SAMPLER_DDIM) {
      o = sap * x0 + s1ap * e;
    } else {
      o = c0 * x0 + ct * x;
      if (t > 0) {
        float z;
        if (noise) {
          z = noise[i];
        } else if (call.row_len > 0) {       // row-structured noise of the sampling loops (same draw as the fused epilogue)
          const long long row = i / call.row_len;
          const int col = (int)(i - row * call.row_len);
          float z4[4];
          philox_normal4_rows(call.seed, call.stream_id, (uint32_t)t, (uint32_t)(call.row_offset + row), (uint32_t)(col >> 2), z4);
          z = (col & 3) == 0 ? z4[0] : ((col & 3) == 1 ? z4[1] : ((col & 3) == 2 ? z4[2] : z4[3]));
        } else {
          z = philox_normal(call.seed, call.stream_id, (uint32_t)t, (unsigned long long)(call.elem_offset + i));
        }
        o = fmaf(sigma, z, o);
      }
    }
    p.out[i] = o;
  }
}

int launch_ddpm_step(const DdpmStep& p, cudaStream_t s) {
  LDP_CHECK(p.coef && p.eps && p.x && p.out && p.n > 0, LDP_ERR_INVALID_ARG, "ddpm_step: bad arguments");
  int blocks = (int)((p.n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ddpm_step_kernel<<<blocks, 256, 0, s>>>(p);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

__global__ void add_noise_kernel(const float* acp, const float* x0, const float* noise, const int32_t* t, float* out,
                                 long long rows, long long row_len) {
  long long n = rows * row_len;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / row_len;
    float a = acp[t[r]];
    out[i] = sqrtf(a) * x0[i] + sqrtf(1.f - a) * noise[i];
  }
}

int launch_add_noise(const float* acp, const float* x0, const float* noise, const int32_t* t, float* out, long long rows,
                     long long row_len, cudaStream_t s) {
  long long n = rows * row_len;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  add_noise_kernel<<<blocks, 256, 0, s>>>(acp, x0, noise, t, out, rows, row_len);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

__global__ void philox_normal_kernel(unsigned long long seed, uint32_t stream_id, uint32_t step, float* out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = philox_normal(seed, stream_id, step, (unsigned long long)i);
}

int launch_philox_normal(unsigned long long seed, uint32_t stream_id, uint32_t step, float* out, long long n,
                         cudaStream_t s) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  philox_normal_kernel<<<blocks, 256, 0, s>>>(seed, stream_id, step, out, n);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

__global__ void philox_normal_rows_kernel(unsigned long long seed, uint32_t stream_id, uint32_t step, long long row0,
                                          long long rows, int row_len, float* out) {
  const int nq = (row_len + 3) >> 2;
  const long long total = rows * nq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nq;
    const int g = (int)(i - r * nq);
    float z[4];
    philox_normal4_rows(seed, stream_id, step, (uint32_t)(row0 + r), (uint32_t)g, z);
    for (int k = 0; k < 4; ++k)
      if (4 * g + k < row_len) out[r * row_len + 4 * g + k] = z[k];
  }
}

int launch_philox_normal_rows(unsigned long long seed, uint32_t stream_id, uint32_t step, long long row0, long long rows,
                              int row_len, float* out, cudaStream_t s) {
  const long long total = rows * ((row_len + 3) >> 2);
  int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  philox_normal_rows_kernel<<<blocks, 256, 0, s>>>(seed, stream_id, step, row0, rows, row_len, out);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

__global__ void transpose_quads_kernel(const float* __restrict__ in, int ld, float4* __restrict__ out, int rows, int nq) {
  // tile of 32 rows x 32 quads through shared memory so that both sides are coalesced
  __shared__ float4 t[32][33];
  const int r0 = blockIdx.x * 32, q0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, q = q0 + threadIdx.x;
    if (r < rows && q < nq) t[i][threadIdx.x] = *reinterpret_cast<const float4*>(in + (long long)r * ld + 4 * q);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int q = q0 + i, r = r0 + threadIdx.x;
    if (r < rows && q < nq) out[(long long)q * rows + r] = t[threadIdx.x][i];
  }
}

int launch_transpose_quads(const float* in, int ld, float* out, int rows, int cols, cudaStream_t s) {
  const int nq = cols / 4;
  dim3 grid((rows + 31) / 32, (nq + 31) / 32), block(32, 8);
  transpose_quads_kernel<<<grid, block, 0, s>>>(in, ld, reinterpret_cast<float4*>(out), rows, nq);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// FiLM observation part for the tcgen05 epilogues: out[(pair / 4) * rows + r] = 4 x half2(scale, shift) of pairs pair .. pair+3, where
// pair = blk.pair_off + c pairs scale = in[r][blk.film_off + c] with shift = in[r][blk.film_off + blk.C + c] (c < blk.C, C % 4 == 0).
// Half the bytes and a third of the instructions of the float4 (scale quad | shift quad) form the epilogue prefetch used to read.
__global__ void film_pack_kernel(const float* __restrict__ in, int ld, uint4* __restrict__ out, int rows, int nq, FilmBlocks fb) {
  __shared__ uint4 t[32][33];
  const int r0 = blockIdx.x * 32, q0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, q = q0 + threadIdx.x;
    if (r < rows && q < nq) {
      const int pair = 4 * q;
      int b = 0;
      while (b + 1 < fb.n && fb.pair_off[b + 1] <= pair) ++b;
      const int c = pair - fb.pair_off[b];
      const float4 sc = *reinterpret_cast<const float4*>(in + (long long)r * ld + fb.film_off[b] + c);
      const float4 sh = *reinterpret_cast<const float4*>(in + (long long)r * ld + fb.film_off[b] + fb.C[b] + c);
      __half2 h0 = __floats2half2_rn(sc.x, sh.x), h1 = __floats2half2_rn(sc.y, sh.y);
      __half2 h2 = __floats2half2_rn(sc.z, sh.z), h3 = __floats2half2_rn(sc.w, sh.w);
      t[i][threadIdx.x] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                     *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int q = q0 + i, r = r0 + threadIdx.x;
    if (r < rows && q < nq) out[(long long)q * rows + r] = t[threadIdx.x][i];
  }
}

int launch_film_pack(const float* in, int ld, void* out, int rows, const FilmBlocks& fb, cudaStream_t s) {
  const int nq = (fb.pair_off[fb.n - 1] + fb.C[fb.n - 1]) / 4;
  dim3 grid((rows + 31) / 32, (nq + 31) / 32), block(32, 8);
  film_pack_kernel<<<grid, block, 0, s>>>(in, ld, reinterpret_cast<uint4*>(out), rows, nq, fb);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

__global__ void add_i32_kernel(int32_t* p, int delta) { *p += delta; }
__global__ void set_i32_kernel(int32_t* p, int v) { *p = v; }
int launch_add_i32(int32_t* p, int delta, cudaStream_t s) {
  add_i32_kernel<<<1, 1, 0, s>>>(p, delta);
  LDP_LAUNCH_OK();
  return LDP_OK;
}
int launch_set_i32(int32_t* p, int v, cudaStream_t s) {
  set_i32_kernel<<<1, 1, 0, s>>>(p, v);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// ------------------------------------------------------------------------------------------------
// casts / packing
// ------------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* in, int ld_in, __nv_bfloat16* out, int ld_out, long long rows, int cols,
                                 int mish) {
  long long n = rows * ld_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / ld_out;
    int c = (int)(i - r * ld_out);
    float v = 0.f;
    if (c < cols) {
      v = in[r * ld_in + c];
      if (mish) v = mish_f<false>(v);
    }
    out[i] = __float2bfloat16(v);
  }
}

int launch_cast_bf16(const float* in, int ld_in, __nv_bfloat16* out, int ld_out, long long rows, int cols, int mish,
                     cudaStream_t s) {
  long long n = rows * ld_out;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_bf16_kernel<<<blocks, 256, 0, s>>>(in, ld_in, out, ld_out, rows, cols, mish);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

__global__ void cast_f32_from_bf16_kernel(const __nv_bfloat16* in, int ld_in, float* out, int ld_out, long long rows, int cols) {
  long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cols;
    int c = (int)(i - r * cols);
    out[r * ld_out + c] = __bfloat162float(in[r * ld_in + c]);
  }
}

int launch_cast_f32_from_bf16(const __nv_bfloat16* in, int ld_in, float* out, int ld_out, long long rows, int cols,
                              cudaStream_t s) {
  long long n = rows * cols;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_f32_from_bf16_kernel<<<blocks, 256, 0, s>>>(in, ld_in, out, ld_out, rows, cols);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// Weight packing for the tcgen05 path: W^T as [n_pad][kp] bf16 (K contiguous), K gathered through `map`.
__global__ void pack_wt_bf16_kernel(const float* src, int ld_src, int n_src, const int32_t* map, int kp,
                                    __nv_bfloat16* dst, int ld_dst, int k_off, int n_pad) {
  __shared__ float tile[32][33];
  int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {   // i: k within tile; threadIdx.x: n (coalesced reads)
    int k = k0 + i, n = n0 + threadIdx.x;
    float v = 0.f;
    if (k < kp && n < n_src) {
      int sk = map[k];
      if (sk >= 0) v = src[(long long)sk * ld_src + n];
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {   // i: n within tile; threadIdx.x: k (coalesced writes)
    int n = n0 + i, k = k0 + threadIdx.x;
    if (n < n_pad && k < kp) dst[(long long)n * ld_dst + k_off + k] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

int launch_pack_wt_bf16(const float* src, int ld_src, int n_src, const int32_t* map, int kp, __nv_bfloat16* dst,
                        int ld_dst, int k_off, int n_pad, cudaStream_t s) {
  dim3 grid(ceil_div(kp, 32), ceil_div(n_pad, 32));
  pack_wt_bf16_kernel<<<grid, dim3(32, 8), 0, s>>>(src, ld_src, n_src, map, kp, dst, ld_dst, k_off, n_pad);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

}  // namespace ldp

// ------------------------------------------------------------------------------------------------
// jax.random on the device (jax 0.4.26, threefry2x32, non-partitionable): random_bits / normal for a batch of keys.
// Counter layout of `threefry_2x32(key, iota(n))`: counters padded to even length with a zero, first half -> x0,
// second half -> x1, outputs concatenated.  One thread per counter pair.
// ------------------------------------------------------------------------------------------------
namespace ldp {

__device__ __forceinline__ void threefry2x32_dev(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int r0 = (i & 1) ? 17 : 13, r1 = (i & 1) ? 29 : 15, r2 = (i & 1) ? 16 : 26, r3 = (i & 1) ? 24 : 6;
    x0 += x1; x1 = __funnelshift_l(x1, x1, r0); x1 ^= x0;
    x0 += x1; x1 = __funnelshift_l(x1, x1, r1); x1 ^= x0;
    x0 += x1; x1 = __funnelshift_l(x1, x1, r2); x1 ^= x0;
    x0 += x1; x1 = __funnelshift_l(x1, x1, r3); x1 ^= x0;
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
}

// jax.random.normal float32: u = bits -> [0,1); v = max(lo, u * (1 - lo) + lo), lo = nextafter(-1, 0); sqrt(2) erfinv(v)
__device__ __forceinline__ float jax_normal_from_bits(uint32_t bits) {
  const float u = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float lo = -0.99999994f;
  const float v = fmaxf(lo, __fadd_rn(__fmul_rn(u, 1.0f - lo), lo));
  return __fmul_rn(1.41421354f, erfinvf(v));
}

__global__ void jax_random_kernel(const uint32_t* __restrict__ keys, long long n, int mode, void* __restrict__ out) {
  const long long half = (n + 1) >> 1;
  const uint32_t k0 = keys[2 * blockIdx.y], k1 = keys[2 * blockIdx.y + 1];
  uint32_t* ob = reinterpret_cast<uint32_t*>(out) + (long long)blockIdx.y * n;
  float* of = reinterpret_cast<float*>(out) + (long long)blockIdx.y * n;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < half; p += (long long)gridDim.x * blockDim.x) {
    uint32_t x0 = (uint32_t)p;
    uint32_t x1 = (half + p < n) ? (uint32_t)(half + p) : 0u;     // the pad element of an odd-length counter vector
    threefry2x32_dev(k0, k1, x0, x1);
    if (mode == 0) {
      ob[p] = x0;
      if (half + p < n) ob[half + p] = x1;
    } else {
      of[p] = jax_normal_from_bits(x0);
      if (half + p < n) of[half + p] = jax_normal_from_bits(x1);
    }
  }
}

int launch_jax_random(const uint32_t* keys_dev, int n_keys, long long n, int mode, void* out, cudaStream_t s) {
  LDP_CHECK(keys_dev && out && n_keys > 0 && n > 0 && n < (1ll << 32), LDP_ERR_INVALID_ARG, "jax_random: bad arguments");
  const long long half = (n + 1) >> 1;
  const int bx = (int)std::min<long long>((half + 255) / 256, 148 * 16);
  jax_random_kernel<<<dim3(bx, n_keys), 256, 0, s>>>(keys_dev, n, mode, out);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

}  // namespace ldp
