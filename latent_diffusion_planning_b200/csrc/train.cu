// Next-row N1: the training objective of LDPAgent.update - denoising losses of the planner
// (reference agent/ldp_agent.py:113-126) and of the IDM (:128-139), their parameter gradients, and the optax.adam
// update (:580-600).  fp32 throughout (the reference trains in float32): implicit-GEMM SIMT kernels for
// forward / data-gradient (simt.cu gemm_f32, w_mode 1) and the weight-gradient kernel below, plus the backward
// kernels of GroupNorm+Mish+FiLM, LayerNorm, Mish, ReLU and the MSE loss.
//
// Ownership: the caller (PyTorch) owns the flat parameter / gradient / Adam-moment buffers in the canonical
// params.py spec order, so the data-parallel exchange is one all-reduce over the gradient buffer.  The trainer handle
// owns only activations.
#include <algorithm>
#include <cmath>
#include <deque>
#include <map>
#include <tuple>

#include "net_common.h"

namespace ldp {

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

// Programmatic dependent launch for the small operand kernels of the bf16 path: launched with the stream-serialisation
// attribute they may become resident while their predecessor drains; `pdl_enter()` is the first thing they execute -
// wait until everything before them in the stream is complete and visible, then let their own dependent (the next
// operand kernel, or the GEMM, whose producer waits on its own) start its prologue.  Nothing is read or written
// before the wait, so shared scratch buffers are safe.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
static bool train_use_pdl() {
  static const bool on = !(getenv("LDP_TRAIN_PDL") && getenv("LDP_TRAIN_PDL")[0] == '0');
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = train_use_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dW[(j,c)][n] += sum_m A(m,(j,c)) dY[m][n];  dbias[n] += sum_m dY[m][n]
// A is the same implicit-GEMM gather as the forward (kernels.h GemmF32).  64x64 tile of (k, n), m in steps of 16,
// grid.z splits m; partial tiles are combined with atomics into the (caller-zeroed) gradient buffer.
// ------------------------------------------------------------------------------------------------
struct WgradF32 {
  const float* x1 = nullptr; int c1 = 0, ld1 = 0;
  const float* x2 = nullptr; int c2 = 0, ld2 = 0;
  int t_in = 1, t_out = 1, taps = 1, stride = 1, pad = 0, dil = 1;
  const float* dy = nullptr; int lddy = 0;
  float* dw = nullptr; int ldw = 0;
  float* dbias = nullptr;
  int m = 0, n = 0, m_chunk = 0;
};

constexpr int WK = 64, WN = 64, WM = 16;

__global__ void __launch_bounds__(256) wgrad_f32_kernel(const WgradF32 p) {
  __shared__ float As[WM][WK + 4];
  __shared__ float Bs[WM][WN + 4];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.y * WK, n0 = blockIdx.x * WN;
  const int ctot = p.c1 + p.c2;
  const int K = p.taps * ctot;
  const int ty = tid / 16, tx = tid % 16;
  const int m_lo = blockIdx.z * p.m_chunk, m_hi = min(p.m, m_lo + p.m_chunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;

  for (int mb = m_lo; mb < m_hi; mb += WM) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int mm = idx / WK, kk = idx % WK;
      int m = mb + mm, k = k0 + kk;
      float v = 0.f;
      if (m < m_hi && k < K) {
        int j = k / ctot, c = k - j * ctot;
        int b = m / p.t_out, t = m - b * p.t_out;
        int num = t * p.stride + j - p.pad;
        if (num >= 0 && (num % p.dil) == 0) {
          int ti = num / p.dil;
          if (ti < p.t_in) {
            long long row = (long long)b * p.t_in + ti;
            v = (c < p.c1) ? p.x1[row * p.ld1 + c] : p.x2[row * p.ld2 + (c - p.c1)];
          }
        }
      }
      As[mm][kk] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int mm = idx / WN, nn = idx % WN;
      int m = mb + mm, n = n0 + nn;
      Bs[mm][nn] = (m < m_hi && n < p.n) ? p.dy[(long long)m * p.lddy + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WM; ++mm) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[mm][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[mm][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (p.dbias && blockIdx.y == 0 && tid < WN) {
#pragma unroll
      for (int mm = 0; mm < WM; ++mm) bsum += Bs[mm][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int k = k0 + ty * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < p.n) atomicAdd(p.dw + (long long)k * p.ldw + n, acc[i][j]);
    }
  }
  if (p.dbias && blockIdx.y == 0 && tid < WN && n0 + tid < p.n) atomicAdd(p.dbias + n0 + tid, bsum);
}

static int launch_wgrad_f32(WgradF32 p, cudaStream_t s) {
  LDP_CHECK(p.x1 && p.dy && p.dw && p.m > 0 && p.n > 0, LDP_ERR_INVALID_ARG, "wgrad_f32: bad arguments");
  const int K = p.taps * (p.c1 + p.c2);
  const int gx = ceil_div(p.n, WN), gy = ceil_div(K, WK);
  int splits = std::max(1, std::min(ceil_div(p.m, 4 * WM), ceil_div(2 * 148, gx * gy)));
  p.m_chunk = round_up(ceil_div(p.m, splits), WM);
  splits = ceil_div(p.m, p.m_chunk);
  wgrad_f32_kernel<<<dim3(gx, gy, splits), 256, 0, s>>>(p);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// ------------------------------------------------------------------------------------------------
// Mish'(x) with n = e^x (e^x + 2):  w = n/(n+2) = tanh(softplus(x));  dw/dx = 4 e (e+1) / (n+2)^2
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mish_and_grad(float x, float* y, float* dydx) {
  if (x > 20.f) { *y = x; *dydx = 1.f; return; }
  float e = expf(x);
  float n = e * (e + 2.f);
  float w = n / (n + 2.f);
  *y = x * w;
  *dydx = w + x * (4.f * e * (e + 1.f)) / ((n + 2.f) * (n + 2.f));
}

// ------------------------------------------------------------------------------------------------
// backward of  y = [scale *] Mish(GN(x)) [+ shift] [+ res]   (forward: simt.cu groupnorm_f32_kernel)
// One block per (sample, group).  dynamic smem: 2*cnt (dxhat, xhat) + 4*gw floats (per-channel sums).
// ------------------------------------------------------------------------------------------------
struct GnBwdF32 {
  const float* x = nullptr;      // pre-norm input (B,P,C)
  const float* dy = nullptr;     // (B,P,C)
  float* dx = nullptr;           // (B,P,C), written
  int B = 0, P = 0, C = 0, G = 0;
  const float* gamma = nullptr; const float* beta = nullptr;
  float* dgamma = nullptr; float* dbeta = nullptr;    // atomics
  float eps = 1e-6f;
  const float* film = nullptr;   // (B,2C) [scale | shift] or null
  float* dfilm = nullptr;        // (B,2C), written
  float* dres = nullptr; int dres_acc = 0;   // residual branch gradient: (+)= dy
};

__device__ __forceinline__ float block_sum_t(float v, float* sh) {
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

__global__ void __launch_bounds__(256) gn_bwd_f32_kernel(const GnBwdF32 p) {
  extern __shared__ float dsm[];
  __shared__ float sh[32];
  const int b = blockIdx.x / p.G, g = blockIdx.x % p.G;
  const int gw = p.C / p.G;
  const int cnt = p.P * gw;
  float* s_dxh = dsm;             // dxhat = dn * gamma
  float* s_xh = dsm + cnt;        // xhat
  float* s_ch = dsm + 2 * cnt;    // [4][gw]: dgamma, dbeta, dscale, dshift
  for (int i = threadIdx.x; i < 4 * gw; i += blockDim.x) s_ch[i] = 0.f;
  const long long base = (long long)b * p.P * p.C + g * gw;
  const float* xb = p.x + base;
  const float* dyb = p.dy + base;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int pos = i / gw, c = i - pos * gw;
    float v = xb[(long long)pos * p.C + c];
    s += v;
    ss += v * v;
  }
  s = block_sum_t(s, sh);
  ss = block_sum_t(ss, sh);
  const float mean = s / cnt;
  const float rstd = rsqrtf(fmaxf(ss / cnt - mean * mean, 0.f) + p.eps);
  const float* frow = p.film ? p.film + (long long)b * 2 * p.C : nullptr;
  float s1 = 0.f, s2 = 0.f;
  // when the group width divides the block size a thread always sees the same channel: per-channel sums stay in
  // registers and are combined once; otherwise shared-memory atomics per element
  const bool fixed_c = (blockDim.x % gw) == 0;
  float a_dg = 0.f, a_db = 0.f, a_ds = 0.f, a_dh = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int pos = i / gw, c = i - pos * gw, ch = g * gw + c;
    float xh = (xb[(long long)pos * p.C + c] - mean) * rstd;
    float nrm = xh * p.gamma[ch] + p.beta[ch];
    float mv, md;
    mish_and_grad(nrm, &mv, &md);
    float dyv = dyb[(long long)pos * p.C + c];
    float dm = dyv;
    if (frow) {
      if (fixed_c) { a_ds += dyv * mv; a_dh += dyv; }
      else { atomicAdd(&s_ch[2 * gw + c], dyv * mv); atomicAdd(&s_ch[3 * gw + c], dyv); }
      dm = dyv * frow[ch];
    }
    float dn = dm * md;
    if (fixed_c) { a_dg += dn * xh; a_db += dn; }
    else { atomicAdd(&s_ch[c], dn * xh); atomicAdd(&s_ch[gw + c], dn); }
    float dxh = dn * p.gamma[ch];
    s_dxh[i] = dxh;
    s_xh[i] = xh;
    s1 += dxh;
    s2 += dxh * xh;
    if (p.dres) {
      float* d = p.dres + base + (long long)pos * p.C + c;
      *d = p.dres_acc ? *d + dyv : dyv;
    }
  }
  if (fixed_c && (int)threadIdx.x < cnt + (int)blockDim.x) {
    // threads c, c + gw, c + 2 gw, ... own channel c: a handful of atomics per channel instead of one per element
    const int c = threadIdx.x % gw;
    atomicAdd(&s_ch[c], a_dg);
    atomicAdd(&s_ch[gw + c], a_db);
    if (frow) { atomicAdd(&s_ch[2 * gw + c], a_ds); atomicAdd(&s_ch[3 * gw + c], a_dh); }
  }
  s1 = block_sum_t(s1, sh) / cnt;
  s2 = block_sum_t(s2, sh) / cnt;
  float* dxb = p.dx + base;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int pos = i / gw, c = i - pos * gw;
    dxb[(long long)pos * p.C + c] = rstd * (s_dxh[i] - s1 - s_xh[i] * s2);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < gw; c += blockDim.x) {
    int ch = g * gw + c;
    atomicAdd(p.dgamma + ch, s_ch[c]);
    atomicAdd(p.dbeta + ch, s_ch[gw + c]);
    if (p.dfilm) {
      p.dfilm[(long long)b * 2 * p.C + ch] = s_ch[2 * gw + c];
      p.dfilm[(long long)b * 2 * p.C + p.C + ch] = s_ch[3 * gw + c];
    }
  }
}

static int launch_gn_bwd_f32(const GnBwdF32& p, cudaStream_t s) {
  LDP_CHECK(p.x && p.dy && p.dx && p.B > 0 && p.G > 0 && p.C % p.G == 0, LDP_ERR_INVALID_ARG, "gn_bwd: bad arguments");
  const int gw = p.C / p.G;
  const size_t smem = ((size_t)2 * p.P * gw + 4 * gw) * sizeof(float);
  LDP_CHECK(smem <= 48 * 1024, LDP_ERR_UNSUPPORTED, "gn_bwd: group too large for shared memory");
  gn_bwd_f32_kernel<<<p.B * p.G, 256, smem, s>>>(p);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  dx = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat)) [+ add];  one warp per row,
// per-block shared partials for dgamma / dbeta, then atomics.  dynamic smem: 2*C floats.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_bwd_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         const float* __restrict__ add, float* __restrict__ dx, int rows,
                                                         int C, const float* __restrict__ gamma, float* dgamma,
                                                         float* dbeta, float eps, int rows_per_block) {
  extern __shared__ float dsm[];
  float* s_dg = dsm;
  float* s_db = dsm + C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) dsm[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int row = r0 + warp; row < r1; row += nw) {
    const float* xr = x + (long long)row * C;
    const float* dyr = dy + (long long)row * C;
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
    const float mean = s / C, rstd = rsqrtf(fmaxf(ss / C - mean * mean, 0.f) + eps);
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      float xh = (xr[c] - mean) * rstd, d = dyr[c];
      atomicAdd(&s_dg[c], d * xh);
      atomicAdd(&s_db[c], d);
      float dxh = d * gamma[c];
      s1 += dxh;
      s2 += dxh * xh;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 /= C;
    s2 /= C;
    for (int c = lane; c < C; c += 32) {
      float xh = (xr[c] - mean) * rstd;
      float v = rstd * (dyr[c] * gamma[c] - s1 - xh * s2);
      if (add) v += add[(long long)row * C + c];
      dx[(long long)row * C + c] = v;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(dgamma + c, s_dg[c]);
    atomicAdd(dbeta + c, s_db[c]);
  }
}

static int launch_ln_bwd_f32(const float* x, const float* dy, const float* add, float* dx, int rows, int C,
                             const float* gamma, float* dgamma, float* dbeta, float eps, cudaStream_t s) {
  const int rpb = 32;
  ln_bwd_f32_kernel<<<ceil_div(rows, rpb), 256, (size_t)2 * C * sizeof(float), s>>>(x, dy, add, dx, rows, C, gamma, dgamma,
                                                                                     dbeta, eps, rpb);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

// ------------------------------------------------------------------------------------------------
// elementwise pieces
// ------------------------------------------------------------------------------------------------
#define LDP_GRID_STRIDE(i, n) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

static int ew_blocks(long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 8); }

// 2-D views: element (r, c) of a (rows, cols) tensor with row pitch ld
__global__ void mish_fwd_kernel(const float* x, int ldx, float* y, int ldy, long long rows, int cols) {
  LDP_GRID_STRIDE(i, rows * cols) {
    long long r = i / cols; int c = (int)(i - r * cols);
    y[r * ldy + c] = mish_f<false>(x[r * ldx + c]);
  }
}
// dx (=|+=) dy * Mish'(x)
__global__ void mish_bwd_kernel(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx, long long rows,
                                int cols, int acc) {
  LDP_GRID_STRIDE(i, rows * cols) {
    long long r = i / cols; int c = (int)(i - r * cols);
    float y, d;
    mish_and_grad(x[r * ldx + c], &y, &d);
    float v = dy[r * lddy + c] * d;
    float* o = dx + r * lddx + c;
    *o = acc ? *o + v : v;
  }
}
__global__ void relu_fwd_kernel(const float* x, float* y, long long n) {
  LDP_GRID_STRIDE(i, n) y[i] = fmaxf(x[i], 0.f);
}
// dx (=|+=) dy * (x > 0)    (x may be the ReLU output as well as its input)
__global__ void relu_bwd_kernel(const float* x, const float* dy, float* dx, long long n, int acc) {
  LDP_GRID_STRIDE(i, n) {
    float v = x[i] > 0.f ? dy[i] : 0.f;
    dx[i] = acc ? dx[i] + v : v;
  }
}
__global__ void gather_rows_kernel(const float* table, int cols, const int32_t* idx, float* out, int ldo, long long rows) {
  LDP_GRID_STRIDE(i, rows * cols) {
    long long r = i / cols; int c = (int)(i - r * cols);
    out[r * ldo + c] = table[(long long)idx[r] * cols + c];
  }
}
// loss += sum((eps - target)^2) / n;   d_eps = weight * 2 (eps - target) / n
__global__ void __launch_bounds__(256) mse_loss_grad_kernel(const float* eps, const float* target, float* d_eps, long long n,
                                                            float weight, float* loss) {
  __shared__ float sh[32];
  float acc = 0.f;
  const float inv = 1.f / (float)n;
  LDP_GRID_STRIDE(i, n) {
    float d = eps[i] - target[i];
    acc += d * d;
    d_eps[i] = weight * 2.f * d * inv;
  }
  acc = block_sum_t(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv);
}

// optax.adam (scale_by_adam + scale(-lr)), count = 1-based step:  mu = b1 mu + (1-b1) g;  nu = b2 nu + (1-b2) g^2;
// p -= lr * (mu / (1-b1^count)) / (sqrt(nu / (1-b2^count)) + eps)
__device__ __forceinline__ void adam_one(float& p, float g, float& mu, float& nu, float lr, float b1, float b2, float eps,
                                         float bc1, float bc2, float gscale) {
  float gv = g * gscale;
  float m = b1 * mu + (1.f - b1) * gv;
  float v = b2 * nu + (1.f - b2) * gv * gv;
  mu = m;
  nu = v;
  p -= lr * (m / bc1) / (sqrtf(v / bc2) + eps);
}
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu,
                                                   float* __restrict__ nu, long long n, float lr, float b1, float b2,
                                                   float eps, float bc1, float bc2, float gscale) {
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(mu);
  float4* v4 = reinterpret_cast<float4*>(nu);
  // two independent float4 quadruples per iteration: 8 loads of 16 bytes in flight per thread
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    const long long j = i + stride;
    float4 pa = __ldcs(p4 + i), ga = __ldcs(g4 + i), ma = __ldcs(m4 + i), va = __ldcs(v4 + i);
    float4 pb = __ldcs(p4 + j), gb = __ldcs(g4 + j), mb = __ldcs(m4 + j), vb = __ldcs(v4 + j);
    adam_one(pa.x, ga.x, ma.x, va.x, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pa.y, ga.y, ma.y, va.y, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pa.z, ga.z, ma.z, va.z, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pa.w, ga.w, ma.w, va.w, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pb.x, gb.x, mb.x, vb.x, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pb.y, gb.y, mb.y, vb.y, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pb.z, gb.z, mb.z, vb.z, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pb.w, gb.w, mb.w, vb.w, lr, b1, b2, eps, bc1, bc2, gscale);
    __stcs(p4 + i, pa); __stcs(m4 + i, ma); __stcs(v4 + i, va);
    __stcs(p4 + j, pb); __stcs(m4 + j, mb); __stcs(v4 + j, vb);
  }
  for (; i < n4; i += stride) {
    float4 pv = __ldcs(p4 + i), gv = __ldcs(g4 + i), mv = __ldcs(m4 + i), vv = __ldcs(v4 + i);
    adam_one(pv.x, gv.x, mv.x, vv.x, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pv.y, gv.y, mv.y, vv.y, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pv.z, gv.z, mv.z, vv.z, lr, b1, b2, eps, bc1, bc2, gscale);
    adam_one(pv.w, gv.w, mv.w, vv.w, lr, b1, b2, eps, bc1, bc2, gscale);
    __stcs(p4 + i, pv); __stcs(m4 + i, mv); __stcs(v4 + i, vv);
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    adam_one(p[i], g[i], mu[i], nu[i], lr, b1, b2, eps, bc1, bc2, gscale);
}

// ------------------------------------------------------------------------------------------------
// bf16 / tcgen05 path: every contraction (forward, data gradient, weight gradient) becomes one dense
// C[M][N] = A[M][K] W[K][N] on tc_gemm_kernel (PLAIN epilogue, fp32 accumulate and output).  The operands are
// materialised in bf16 by the kernels below: the implicit-GEMM gather (im2col), its transpose for the weight
// gradient (the reduction runs over rows there), and the K-major weight packs.
// ------------------------------------------------------------------------------------------------
struct Im2col {
  const float* x1 = nullptr; int c1 = 0, ld1 = 0;
  const float* x2 = nullptr; int c2 = 0, ld2 = 0;
  int t_in = 1, t_out = 1, taps = 1, stride = 1, pad = 0, dil = 1;
  int m = 0;
};
__device__ __forceinline__ float im2col_at(const Im2col& p, int m, int k, int ctot) {
  int j = k / ctot, c = k - j * ctot;
  int b = m / p.t_out, t = m - b * p.t_out;
  int num = t * p.stride + j - p.pad;
  if (num < 0 || (num % p.dil) != 0) return 0.f;
  int ti = num / p.dil;
  if (ti >= p.t_in) return 0.f;
  long long row = (long long)b * p.t_in + ti;
  return (c < p.c1) ? p.x1[row * p.ld1 + c] : p.x2[row * p.ld2 + (c - p.c1)];
}
// ---- operand builders.  Each GEMM needs two of them (an A part and a W part); both run in ONE launch (`prep_kernel`),
// the leading blocks on the A part and the rest on the W part - at batch 256 the step is bound by kernel boundaries,
// not by bytes.  Grid-stride parts take (block id, block count); tiled parts take their 32x32 tile coordinates.

// A[m][k] for k < kp (zero beyond K).  v8: 8 consecutive k per thread - needs c1, c2, ld1, ld2 multiples of 8 (a group
// never straddles a tap or a source) and 16-byte aligned inputs.
__device__ __forceinline__ void im2col_body(const Im2col& p, __nv_bfloat16* __restrict__ out, int kp, int v8, long long bid,
                                            long long nb) {
  const int ctot = p.c1 + p.c2, K = p.taps * ctot;
  if (v8) {
    const int kp8 = kp >> 3;
    for (long long i = bid * blockDim.x + threadIdx.x; i < (long long)p.m * kp8; i += nb * blockDim.x) {
      int m = (int)(i / kp8), k = (int)(i - (long long)m * kp8) << 3;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (k < K) {
        int j = k / ctot, c = k - j * ctot;
        int bi = m / p.t_out, t = m - bi * p.t_out;
        int num = t * p.stride + j - p.pad;
        if (num >= 0 && (num % p.dil) == 0) {
          int ti = num / p.dil;
          if (ti < p.t_in) {
            long long row = (long long)bi * p.t_in + ti;
            const float* src = (c < p.c1) ? p.x1 + row * p.ld1 + c : p.x2 + row * p.ld2 + (c - p.c1);
            a = *reinterpret_cast<const float4*>(src);
            b = *reinterpret_cast<const float4*>(src + 4);
          }
        }
      }
      __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w),
                             __floats2bfloat162_rn(b.x, b.y), __floats2bfloat162_rn(b.z, b.w)};
      *reinterpret_cast<uint4*>(out + (long long)m * kp + k) = *reinterpret_cast<uint4*>(o);
    }
  } else {
    for (long long i = bid * blockDim.x + threadIdx.x; i < (long long)p.m * kp; i += nb * blockDim.x) {
      int m = (int)(i / kp), k = (int)(i - (long long)m * kp);
      out[i] = __float2bfloat16(k < K ? im2col_at(p, m, k, ctot) : 0.f);
    }
  }
}

// At[k][m] for m < mp (zero beyond M): the weight gradient reduces over rows.  32x32 tile (bx over k, by over m)
__device__ __forceinline__ void im2col_t_body(const Im2col& p, __nv_bfloat16* __restrict__ out, int mp, int bx, int by) {
  __shared__ float tile[32][33];
  const int ctot = p.c1 + p.c2, K = p.taps * ctot;
  const int k0 = bx * 32, m0 = by * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  {
    // this thread's k (tap j, channel c) is fixed; only the row changes
    const int k = k0 + tx;
    const bool kok = k < K;
    const int j = kok ? k / ctot : 0, c = kok ? k - j * ctot : 0;
    const float* src = (c < p.c1) ? p.x1 + c : p.x2 + (c - p.c1);
    const int ld = (c < p.c1) ? p.ld1 : p.ld2;
    for (int r = ty; r < 32; r += 8) {
      const int m = m0 + r;
      float v = 0.f;
      if (kok && m < p.m) {
        const int b = m / p.t_out, t = m - b * p.t_out;
        const int num = t * p.stride + j - p.pad;
        if (num >= 0 && (num % p.dil) == 0) {
          const int ti = num / p.dil;
          if (ti < p.t_in) v = src[((long long)b * p.t_in + ti) * ld];
        }
      }
      tile[r][tx] = v;
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int k = k0 + r, m = m0 + tx;
    if (k < K && m < mp) out[(long long)k * mp + m] = __float2bfloat16(tile[tx][r]);
  }
}

// dst[n][k] = bf16(src[k*ld + n]) for n < n_pad, k < kp (zeros outside K x N): the K-major "W^T" operand.  colsum, if
// given, also receives sum_k src[k][n] (the bias gradient when src is dY).  64x64 tile (bx over n, by over k; kp and
// n_pad are multiples of 64): float4 reads along n, 16-byte bf16 writes along k.
constexpr int PT = 64;
__device__ __forceinline__ void pack_t_body(const float* __restrict__ src, int ld, int K, int N, __nv_bfloat16* __restrict__ dst,
                                            int kp, int n_pad, float* __restrict__ colsum, int bx, int by) {
  __shared__ float tile[PT][PT + 1];
  const int n0 = bx * PT, k0 = by * PT;
  const int tid = threadIdx.x;
  const bool vec = (ld & 3) == 0 && (((uintptr_t)src) & 15) == 0;
  {
    const int c4 = (tid & 15) * 4, r0 = tid >> 4;          // 16 threads x float4 cover 64 columns; 16 rows per pass
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int r = r0 + pass * 16, k = k0 + r, n = n0 + c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        const float* sp = src + (long long)k * ld + n;
        if (vec && n + 3 < N) {
          v = *reinterpret_cast<const float4*>(sp);
        } else {
          if (n < N) v.x = sp[0];
          if (n + 1 < N) v.y = sp[1];
          if (n + 2 < N) v.z = sp[2];
          if (n + 3 < N) v.w = sp[3];
        }
      }
      tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
    }
  }
  __syncthreads();
  {
    const int k8 = (tid & 7) * 8, nr0 = tid >> 3;          // 8 threads x 8 bf16 cover 64 k; 32 n rows per pass
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int nr = nr0 + pass * 32, n = n0 + nr;
      if (n < n_pad) {
        __nv_bfloat162 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __floats2bfloat162_rn(tile[k8 + 2 * i][nr], tile[k8 + 2 * i + 1][nr]);
        *reinterpret_cast<uint4*>(dst + (long long)n * kp + k0 + k8) = *reinterpret_cast<uint4*>(o);
      }
    }
  }
  if (colsum && tid < PT && n0 + tid < N) {
    float acc = 0.f;
#pragma unroll 8
    for (int r = 0; r < PT; ++r) acc += tile[r][tid];
    atomicAdd(colsum + n0 + tid, acc);
  }
}

// data-gradient operand of a forward kernel W[taps][ctot][cout]:  dst[n = c][k = (j', co)] = W[taps-1-j'][c][co]
// (v8: 8 consecutive k per thread, cout a multiple of 8)
__device__ __forceinline__ void pack_dgrad_body(const float* __restrict__ w, int taps, int ctot, int cout,
                                                __nv_bfloat16* __restrict__ dst, int kp, int n_pad, int v8, long long bid,
                                                long long nb) {
  const int K = taps * cout;
  if (v8) {
    const int kp8 = kp >> 3;
    for (long long i = bid * blockDim.x + threadIdx.x; i < (long long)n_pad * kp8; i += nb * blockDim.x) {
      int n = (int)(i / kp8), k = (int)(i - (long long)n * kp8) << 3;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (n < ctot && k < K) {
        int j = k / cout, co = k - j * cout;
        const float* src = w + ((long long)(taps - 1 - j) * ctot + n) * cout + co;
        a = *reinterpret_cast<const float4*>(src);
        b = *reinterpret_cast<const float4*>(src + 4);
      }
      __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w),
                             __floats2bfloat162_rn(b.x, b.y), __floats2bfloat162_rn(b.z, b.w)};
      *reinterpret_cast<uint4*>(dst + (long long)n * kp + k) = *reinterpret_cast<uint4*>(o);
    }
  } else {
    for (long long i = bid * blockDim.x + threadIdx.x; i < (long long)n_pad * kp; i += nb * blockDim.x) {
      int n = (int)(i / kp), k = (int)(i - (long long)n * kp);
      float v = 0.f;
      if (n < ctot && k < K) {
        int j = k / cout, co = k - j * cout;
        v = w[((long long)(taps - 1 - j) * ctot + n) * cout + co];
      }
      dst[i] = __float2bfloat16(v);
    }
  }
}

struct PrepArgs {
  // A part: a_kind 0 row-major im2col (grid-stride over a_blocks), 1 transposed im2col (tiles a_gx x a_gy, a_blocks = product)
  Im2col q;
  __nv_bfloat16* a_out = nullptr;
  int a_kind = 0, a_ld = 0, a_v8 = 0, a_blocks = 0, a_gx = 1;
  // W part: w_kind 0 transpose-pack of src[K][N] (tiles w_gx x w_gy), 1 data-gradient view of a conv kernel (grid-stride)
  const float* w_src = nullptr;
  __nv_bfloat16* w_out = nullptr;
  int w_kind = 0, w_ld = 0, w_K = 0, w_N = 0, w_kp = 0, w_npad = 0, w_v8 = 0, w_blocks = 0, w_gx = 1;
  int taps = 1, ctot = 0, cout = 0;
  float* colsum = nullptr;
};

__global__ void __launch_bounds__(256) prep_kernel(const PrepArgs a) {
  pdl_enter();
  int b = blockIdx.x;
  if (b < a.a_blocks) {
    if (a.a_kind == 0) im2col_body(a.q, a.a_out, a.a_ld, a.a_v8, b, a.a_blocks);
    else im2col_t_body(a.q, a.a_out, a.a_ld, b % a.a_gx, b / a.a_gx);
  } else {
    b -= a.a_blocks;
    if (a.w_kind == 0) pack_t_body(a.w_src, a.w_ld, a.w_K, a.w_N, a.w_out, a.w_kp, a.w_npad, a.colsum, b % a.w_gx, b / a.w_gx);
    else pack_dgrad_body(a.w_src, a.taps, a.ctot, a.cout, a.w_out, a.w_kp, a.w_npad, a.w_v8, b, a.w_blocks);
  }
}

static int launch_prep(const PrepArgs& a, cudaStream_t s) {
  LDP_CUDA_OK(launch_pdl(prep_kernel, dim3((unsigned)(a.a_blocks + a.w_blocks)), dim3(256), s, a));
  count_launch();
  return LDP_OK;
}

static bool im2col_v8(const Im2col& q) {
  return (q.c1 % 8) == 0 && (q.c2 % 8) == 0 && (q.ld1 % 8) == 0 && (q.c2 == 0 || (q.ld2 % 8) == 0) &&
         (((uintptr_t)q.x1 | (uintptr_t)q.x2) & 15) == 0;
}
static int stride_blocks(long long items) { return (int)std::min<long long>((items + 255) / 256, 148 * 4); }

// ------------------------------------------------------------------------------------------------
// tensors with gradients
// ------------------------------------------------------------------------------------------------
struct Tn {
  float* v = nullptr;
  float* g = nullptr;     // null: no gradient wanted (data)
  int rows = 0, c = 0, ld = 0;
  bool gset = false;      // gradient already holds a contribution in this backward pass
};

struct Geo { int t_in = 1, t_out = 1, taps = 1, stride = 1, pad = 0, dil = 1; };

struct TrainWs {
  Arena arena;
  std::deque<Tn> pool;
  size_t cursor = 0;      // tensors are handed out in a fixed order, so a second pass reuses the first pass's buffers
  bool built = false;
  int key_a = 0, key_b = 0;
  Tn* get(int rows, int c, bool grad) {
    if (cursor < pool.size()) {
      Tn* t = &pool[cursor++];
      t->gset = false;
      return (t->rows == rows && t->c == c && (t->g != nullptr) == grad) ? t : nullptr;
    }
    Tn t;
    t.rows = rows; t.c = c; t.ld = c;
    if (arena.alloc_t(&t.v, (size_t)rows * c) != LDP_OK) return nullptr;
    if (grad && arena.alloc_t(&t.g, (size_t)rows * c) != LDP_OK) return nullptr;
    pool.push_back(t);
    ++cursor;
    return &pool.back();
  }
  void rewind() { cursor = 0; }
  void reset() { arena.release(); pool.clear(); cursor = 0; }
};

// growable device scratch (operands of the bf16 path; reused by consecutive GEMMs, stream order keeps that safe)
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return LDP_OK;
    if (p) { cudaDeviceSynchronize(); cudaFree(p); p = nullptr; cap = 0; }
    bytes = bytes + bytes / 4;
    LDP_CUDA_OK(cudaMalloc(&p, bytes));
    cap = bytes;
    return LDP_OK;
  }
  ~Scratch() { if (p) cudaFree(p); }
};

// Dense bf16 GEMMs on tc_gemm_kernel.  Ops (two tensor maps each) are cached by their pointers and shapes: the
// workspace and scratch buffers keep their addresses from step to step.
struct TcDense {
  Arena arena;
  TcStage* kb_dev = nullptr;     // identity stage table, shared by every op
  int kb_cap = 0;
  struct Key {
    const void *a, *w, *out, *res, *bias; int M, kp, N, lda, ldo, relu;
    bool operator<(const Key& o) const {
      return std::tie(a, w, out, res, bias, M, kp, N, lda, ldo, relu) <
             std::tie(o.a, o.w, o.out, o.res, o.bias, o.M, o.kp, o.N, o.lda, o.ldo, o.relu);
    }
  };
  std::map<Key, TcGemm> ops;
  struct EpiMaps { CUtensorMap m[3]; };
  std::deque<EpiMaps> epi_host;
  std::deque<TcRun> run_host;
  // Forget every cached op and its device-side tables (called when the step's shape changes: the workspace is rebuilt then, so the
  // pointer-keyed entries can never be hit again and would only accumulate).  The caller has synchronised the device.
  void reset() {
    ops.clear();
    epi_host.clear();
    run_host.clear();
    arena.release();
    kb_dev = nullptr;
    kb_cap = 0;
  }
  int init(int max_kb) {
    if (kb_dev && max_kb <= kb_cap) return LDP_OK;
    LDP_TRY(tc_driver_check());
    LDP_TRY(tc_gemm_init());
    std::vector<TcStage> kb(max_kb);
    for (int i = 0; i < max_kb; ++i) kb[i] = make_stage(0, 0, 1, i * 64, 0, 0, i);
    LDP_TRY(arena.alloc_t(&kb_dev, max_kb));
    LDP_CUDA_OK(cudaMemcpy(kb_dev, kb.data(), kb.size() * sizeof(TcStage), cudaMemcpyHostToDevice));
    kb_cap = max_kb;
    ops.clear();
    return LDP_OK;
  }
  // out[M][N] (ldo) = A[M][kp] (lda, bf16) . Wt[n_pad][kp]^T + bias, optional relu, optional + res (may alias out)
  int gemm(const __nv_bfloat16* a, int lda, int M, int kp, const __nv_bfloat16* wt, int n_pad, int N, const float* bias,
           int relu, const float* res, int ldres, float* out, int ldo, cudaStream_t s) {
    LDP_TRY(init(std::max(kp / 64, 256)));
    Key key{a, wt, out, res, bias, M, kp, N, lda, ldo, relu};
    auto it = ops.find(key);
    if (it == ops.end()) {
      TcGemm op;
      uint64_t ad[4] = {(uint64_t)kp, 1, 1, (uint64_t)M};
      uint64_t as[3] = {(uint64_t)lda * 2, (uint64_t)lda * 2, (uint64_t)lda * 2};
      uint32_t ab[4] = {64, 1, 1, 128};
      LDP_TRY(make_tmap_bf16(&op.map_a[0], a, 4, ad, as, ab));
      for (int i = 1; i < 4; ++i) op.map_a[i] = op.map_a[0];
      uint64_t bd[2] = {(uint64_t)kp, (uint64_t)n_pad};
      uint64_t bs[1] = {(uint64_t)kp * 2};
      // 128-wide tiles leave most SMs idle on the small-M layers (batch 256: 16 x 2 tiles): halve the tile there
      static const int force_bn = getenv("LDP_TRAIN_BN") ? atoi(getenv("LDP_TRAIN_BN")) : 0;
      const int bn = force_bn ? force_bn : ((ceil_div(M, 128) * ceil_div(N, 128) < 120 && N > 64) ? 64 : 128);
      uint32_t bb[2] = {64, (uint32_t)bn};
      LDP_TRY(make_tmap_bf16(&op.map_b, wt, 2, bd, bs, bb));
      op.block_n = bn;
      run_host.emplace_back();           // stable address: under graph capture the copy below is a node that re-reads its source
      TcRun& run = run_host.back();
      run.src_acc = make_stage(0, 0, 1, 0, 0, 0, 0).src_acc; run.count = kp / 64;
      TcRun* run_dev;
      LDP_TRY(arena.alloc_t(&run_dev, 1, false));      // (no legacy-stream memset racing the copy on `s`)
      LDP_CUDA_OK(cudaMemcpyAsync(run_dev, &run, sizeof(run), cudaMemcpyHostToDevice, s));
      op.kb = kb_dev; op.num_kb = kp / 64; op.runs = run_dev; op.num_runs = 1;
      tc_set_inline_runs(&op, &run, 1);
      op.M = M; op.N = N; op.items_per_tile = 128; op.rows_per_item = 1;
      op.mode = TC_EPI_PLAIN; op.bias = bias; op.relu = relu;
      op.res_f32 = res; op.ld_res_f32 = ldres; op.out_f32 = out; op.ld_out_f32 = ldo;
      // output / residual as TMA boxes (TcGemm::epi_tma) where the shapes allow: LDP_TRAIN_EPI_TMA=0 keeps the row-per-thread accesses
      static const bool epi_tma = !(getenv("LDP_TRAIN_EPI_TMA") && getenv("LDP_TRAIN_EPI_TMA")[0] == '0');
      if (epi_tma) {
        // ops may be created while the step is being captured into a graph: the copy node then reads its source at every replay, so
        // the host copies of the maps live as long as this object
        epi_host.emplace_back();
        CUtensorMap* host = epi_host.back().m;
        int bits = 0;
        LDP_TRY(tc_build_epi_maps(op, (size_t)M, host, &bits));
        if (bits) {
          CUtensorMap* dev;
          // no zero fill: Arena's cudaMemset runs on the legacy stream, which a non-blocking `s` does not order against - it could land
          // after the copy below
          LDP_TRY(arena.alloc_t(&dev, 3, false));
          LDP_CUDA_OK(cudaMemcpyAsync(dev, host, 3 * sizeof(CUtensorMap), cudaMemcpyHostToDevice, s));
          op.epi_maps = dev;
          op.epi_tma = bits;
        }
      }
      it = ops.emplace(key, op).first;
    }
    return launch_tc_gemm(it->second, s);
  }
};

struct TrainCtx {
  cudaStream_t s = 0;
  int prec = LDP_PREC_FP32;
  TcDense* tc = nullptr;
  Scratch *sa = nullptr, *sw = nullptr;   // A-operand and W-operand scratch
  // bf16 path: weight gradients run on a side stream with their own scratch, next to the data gradients on `s` - the
  // two only share dY as an input, and most GEMMs at batch 256 fill a fraction of the SMs
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  Scratch *sa2 = nullptr, *sw2 = nullptr;
};

#define LDP_TN(var, ws, rows, c, grad)                                                            \
  Tn* var = (ws).get((rows), (c), (grad));                                                        \
  LDP_CHECK(var != nullptr, LDP_ERR_CUDA, "trainer workspace allocation failed")

static Im2col im2col_of(const Tn& x1, const Tn* x2, const Geo& g, int m, bool grad_values = false) {
  Im2col q;
  q.x1 = grad_values ? x1.g : x1.v; q.c1 = x1.c; q.ld1 = x1.ld;
  if (x2) { q.x2 = x2->v; q.c2 = x2->c; q.ld2 = x2->ld; }
  q.t_in = g.t_in; q.t_out = g.t_out; q.taps = g.taps; q.stride = g.stride; q.pad = g.pad; q.dil = g.dil; q.m = m;
  return q;
}

// y = conv(x1 [| x2]) + bias [relu] [+ res]
static int conv_fwd(const Tn& x1, const Tn* x2, const Geo& g, const float* w, const float* bias, int cout, int act,
                    const Tn* res, Tn* y, TrainCtx& cx) {
  cudaStream_t s = cx.s;
  if (cx.prec == LDP_PREC_BF16) {
    const int K = g.taps * (x1.c + (x2 ? x2->c : 0)), kp = round_up(K, 64), n_pad = round_up(cout, 128), M = y->rows;
    LDP_TRY(cx.sa->ensure((size_t)M * kp * 2));
    LDP_TRY(cx.sw->ensure((size_t)n_pad * kp * 2));
    __nv_bfloat16* A = (__nv_bfloat16*)cx.sa->p;
    __nv_bfloat16* W = (__nv_bfloat16*)cx.sw->p;
    PrepArgs pa;
    pa.q = im2col_of(x1, x2, g, M); pa.a_out = A; pa.a_kind = 0; pa.a_ld = kp; pa.a_v8 = im2col_v8(pa.q) ? 1 : 0;
    pa.a_blocks = stride_blocks((long long)M * (pa.a_v8 ? kp / 8 : kp));
    pa.w_kind = 0; pa.w_src = w; pa.w_ld = cout; pa.w_K = K; pa.w_N = cout; pa.w_out = W; pa.w_kp = kp; pa.w_npad = n_pad;
    pa.w_gx = n_pad / PT; pa.w_blocks = pa.w_gx * (kp / PT);
    LDP_TRY(launch_prep(pa, s));
    return cx.tc->gemm(A, kp, M, kp, W, n_pad, cout, bias, act, res ? res->v : nullptr, res ? res->ld : 0, y->v, y->ld, s);
  }
  GemmF32 p;
  p.x1 = x1.v; p.c1 = x1.c; p.ld1 = x1.ld;
  if (x2) { p.x2 = x2->v; p.c2 = x2->c; p.ld2 = x2->ld; }
  p.t_in = g.t_in; p.t_out = g.t_out; p.taps = g.taps; p.stride = g.stride; p.pad = g.pad; p.dil = g.dil;
  p.w = w; p.ldw = cout; p.bias = bias; p.act = act;
  if (res) { p.res = res->v; p.ldres = res->ld; }
  p.out = y->v; p.ldo = y->ld; p.m = y->rows; p.n = cout;
  return launch_gemm_f32(p, s);
}

// data gradient of one source: src.g (=|+=) dY (*) W^T restricted to the source's channel range
static int conv_dgrad(Tn* src, int coff, int ctot, const Geo& g, const float* w, int cout, const Tn& y, TrainCtx& cx) {
  if (!src->g) return LDP_OK;
  cudaStream_t s = cx.s;
  Geo gt;                       // the transposed geometry: a convolution over dY
  gt.t_in = g.t_out; gt.t_out = g.t_in; gt.taps = g.taps; gt.stride = g.dil; gt.dil = g.stride; gt.pad = g.taps - 1 - g.pad;
  const bool acc = src->gset;
  src->gset = true;
  if (cx.prec == LDP_PREC_BF16) {
    const int K = g.taps * cout, kp = round_up(K, 64), n_pad = round_up(ctot, 128) + 128, M = src->rows;
    LDP_TRY(cx.sa->ensure((size_t)M * kp * 2));
    LDP_TRY(cx.sw->ensure((size_t)n_pad * kp * 2));
    __nv_bfloat16* A = (__nv_bfloat16*)cx.sa->p;
    __nv_bfloat16* W = (__nv_bfloat16*)cx.sw->p;
    Tn dy = y;
    dy.c = cout;
    PrepArgs pa;
    pa.q = im2col_of(dy, nullptr, gt, M, true); pa.a_out = A; pa.a_kind = 0; pa.a_ld = kp; pa.a_v8 = im2col_v8(pa.q) ? 1 : 0;
    pa.a_blocks = stride_blocks((long long)M * (pa.a_v8 ? kp / 8 : kp));
    pa.w_kind = 1; pa.w_src = w; pa.taps = g.taps; pa.ctot = ctot; pa.cout = cout; pa.w_out = W; pa.w_kp = kp; pa.w_npad = n_pad;
    pa.w_v8 = ((cout % 8) == 0 && ((uintptr_t)w & 15) == 0) ? 1 : 0;
    pa.w_blocks = stride_blocks((long long)n_pad * (pa.w_v8 ? kp / 8 : kp));
    LDP_TRY(launch_prep(pa, s));
    // rows [coff, coff + src->c) of the pack are this source's channels
    return cx.tc->gemm(A, kp, M, kp, W + (size_t)coff * kp, round_up(src->c, 128), src->c, nullptr, 0,
                       acc ? src->g : nullptr, src->ld, src->g, src->ld, s);
  }
  GemmF32 p;
  p.x1 = y.g; p.c1 = cout; p.ld1 = y.ld;
  p.t_in = gt.t_in; p.t_out = gt.t_out; p.taps = gt.taps; p.stride = gt.stride; p.dil = gt.dil; p.pad = gt.pad;
  p.w = w; p.ldw = cout; p.w_mode = 1; p.w_ctot = ctot; p.w_coff = coff;
  if (acc) { p.res = src->g; p.ldres = src->ld; }
  p.out = src->g; p.ldo = src->ld; p.m = src->rows; p.n = src->c;
  return launch_gemm_f32(p, s);
}

// weight and bias gradients (added into dw / db), then the data gradients of the sources
static int conv_wgrad(const Tn& x1, const Tn* x2, const Geo& g, float* dw, float* db, int cout, const Tn& y, TrainCtx& cx) {
  cudaStream_t s = cx.s;
  if (cx.prec == LDP_PREC_BF16) {
    const int K = g.taps * (x1.c + (x2 ? x2->c : 0)), m = y.rows, mp = round_up(m, 64), n_pad = round_up(cout, 128);
    LDP_TRY(cx.sa->ensure((size_t)K * mp * 2));
    LDP_TRY(cx.sw->ensure((size_t)n_pad * mp * 2));
    __nv_bfloat16* At = (__nv_bfloat16*)cx.sa->p;
    __nv_bfloat16* Yt = (__nv_bfloat16*)cx.sw->p;
    PrepArgs pa;
    pa.q = im2col_of(x1, x2, g, m); pa.a_out = At; pa.a_kind = 1; pa.a_ld = mp; pa.a_gx = ceil_div(K, 32);
    pa.a_blocks = pa.a_gx * (mp / 32);
    pa.w_kind = 0; pa.w_src = y.g; pa.w_ld = y.ld; pa.w_K = m; pa.w_N = cout; pa.w_out = Yt; pa.w_kp = mp; pa.w_npad = n_pad;
    pa.w_gx = n_pad / PT; pa.w_blocks = pa.w_gx * (mp / PT); pa.colsum = db;
    LDP_TRY(launch_prep(pa, s));
    LDP_TRY(cx.tc->gemm(At, mp, K, mp, Yt, n_pad, cout, nullptr, 0, dw, cout, dw, cout, s));
    return LDP_OK;
  }
  WgradF32 q;
  q.x1 = x1.v; q.c1 = x1.c; q.ld1 = x1.ld;
  if (x2) { q.x2 = x2->v; q.c2 = x2->c; q.ld2 = x2->ld; }
  q.t_in = g.t_in; q.t_out = g.t_out; q.taps = g.taps; q.stride = g.stride; q.pad = g.pad; q.dil = g.dil;
  q.dy = y.g; q.lddy = y.ld; q.dw = dw; q.ldw = cout; q.dbias = db; q.m = y.rows; q.n = cout;
  return launch_wgrad_f32(q, s);
}

// the weight gradient of one layer, on the side stream when there is one (dY is complete on `s` at this point)
static int conv_wgrad_side(const Tn& x1, const Tn* x2, const Geo& g, float* dw, float* db, int cout, const Tn& y, TrainCtx& cx) {
  if (cx.prec != LDP_PREC_BF16 || !cx.s2) return conv_wgrad(x1, x2, g, dw, db, cout, y, cx);
  LDP_CUDA_OK(cudaEventRecord(cx.ev_fork, cx.s));
  LDP_CUDA_OK(cudaStreamWaitEvent(cx.s2, cx.ev_fork, 0));
  TrainCtx c2 = cx;
  c2.s = cx.s2; c2.sa = cx.sa2; c2.sw = cx.sw2; c2.s2 = nullptr;
  return conv_wgrad(x1, x2, g, dw, db, cout, y, c2);
}
// a context on the side stream, ordered after everything issued on `s` so far (falls back to `cx` itself)
static int fork_side(TrainCtx& cx, TrainCtx* c2) {
  *c2 = cx;
  if (cx.prec != LDP_PREC_BF16 || !cx.s2) return LDP_OK;
  LDP_CUDA_OK(cudaEventRecord(cx.ev_fork, cx.s));
  LDP_CUDA_OK(cudaStreamWaitEvent(cx.s2, cx.ev_fork, 0));
  c2->s = cx.s2; c2->sa = cx.sa2; c2->sw = cx.sw2; c2->s2 = nullptr;
  return LDP_OK;
}
// the main stream waits for everything issued on the side stream so far
static int join_side(TrainCtx& cx) {
  if (cx.prec != LDP_PREC_BF16 || !cx.s2) return LDP_OK;
  LDP_CUDA_OK(cudaEventRecord(cx.ev_join, cx.s2));
  LDP_CUDA_OK(cudaStreamWaitEvent(cx.s, cx.ev_join, 0));
  return LDP_OK;
}

static int conv_bwd(Tn* x1, Tn* x2, const Geo& g, const float* w, float* dw, float* db, int cout, const Tn& y,
                    TrainCtx& cx) {
  LDP_TRY(conv_wgrad_side(*x1, x2, g, dw, db, cout, y, cx));
  const int ctot = x1->c + (x2 ? x2->c : 0);
  LDP_TRY(conv_dgrad(x1, 0, ctot, g, w, cout, y, cx));
  if (x2) LDP_TRY(conv_dgrad(x2, x1->c, ctot, g, w, cout, y, cx));
  return LDP_OK;
}

static int mish_fwd(const Tn& x, Tn* y, cudaStream_t s) {
  mish_fwd_kernel<<<ew_blocks((long long)x.rows * x.c), 256, 0, s>>>(x.v, x.ld, y->v, y->ld, x.rows, x.c);
  LDP_LAUNCH_OK();
  return LDP_OK;
}
static int mish_bwd(Tn* x, const Tn& y, cudaStream_t s) {
  if (!x->g) return LDP_OK;
  mish_bwd_kernel<<<ew_blocks((long long)x->rows * x->c), 256, 0, s>>>(x->v, x->ld, y.g, y.ld, x->g, x->ld, x->rows, x->c,
                                                                       x->gset ? 1 : 0);
  LDP_LAUNCH_OK();
  x->gset = true;
  return LDP_OK;
}

static int mse(const Tn& eps, const float* target, float weight, float* loss_dev, cudaStream_t s) {
  const long long n = (long long)eps.rows * eps.c;
  mse_loss_grad_kernel<<<ew_blocks(n), 256, 0, s>>>(eps.v, target, eps.g, n, weight, loss_dev);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

static int gather_rows(const float* table, int cols, const int32_t* idx, float* out, int ldo, int rows, cudaStream_t s) {
  gather_rows_kernel<<<ew_blocks((long long)rows * cols), 256, 0, s>>>(table, cols, idx, out, ldo, rows);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

}  // namespace ldp

using namespace ldp;

// ================================================================================================
// trainer handle
// ================================================================================================
struct LdpTrainer {
  int kind = 0;                 // 0 planner UNet, 1 IDM
  LdpUnetConfig ucfg{};
  LdpIdmConfig icfg{};
  Arena arena;
  float* acp = nullptr;         // alphas_cumprod f32 [n_train]
  float* time_table = nullptr;  // sinusoid / Fourier features of every timestep [n_train][dim]
  int64_t n_params = 0;
  TrainWs ws;
  TcDense tc;
  Scratch scratch_a, scratch_w, scratch_a2, scratch_w2;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // Captured steps (bf16 path).  A step is ~750 launches plus ~120 cross-stream events: issued one by one the host is
  // the bottleneck, replayed as a graph it is not.  Inputs are staged into fixed buffers so the graph's pointers hold.
  struct GraphKey {
    int a, b; const void *params, *grads, *loss; float weight;
    bool operator<(const GraphKey& o) const {
      return std::tie(a, b, params, grads, loss, weight) < std::tie(o.a, o.b, o.params, o.grads, o.loss, o.weight);
    }
  };
  struct GraphSlot { int warm = 0; cudaGraphExec_t exec = nullptr; };
  std::map<GraphKey, GraphSlot> graphs;
  Scratch stage[4];
  bool use_graph = true;
  cudaStream_t cap_stream = nullptr;
  // Gradient buckets for the data-parallel exchange: contiguous ranges of the flat gradient buffer in the order the backward
  // pass completes them (reverse layer order).  A bucket with ev >= 0 has its own event, recorded inside the step (an external
  // event-record node when the step is a captured graph) once every kernel that writes the range - data gradients on the main
  // stream, weight gradients on the side stream - has been issued; ev = -1: complete when the whole call is (stream order).
  struct Bucket { int64_t off, len; int ev; };
  std::vector<Bucket> buckets;
  std::vector<cudaEvent_t> bucket_ev;
  int64_t bucket_floats = 6 << 20;                    // close a bucket every ~24 MB (LDP_BUCKET_MB)
  void drop_graphs() {
    for (auto& kv : graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs.clear();
  }
  LdpTrainer() {
    const char* ge = getenv("LDP_TRAIN_GRAPH");
    use_graph = !(ge && ge[0] == '0');
    if (const char* be = getenv("LDP_BUCKET_MB")) bucket_floats = std::max<int64_t>(1024, (int64_t)(atof(be) * (1 << 18)));
    const char* e = getenv("LDP_TRAIN_STREAMS");
    if (!(e && e[0] == '1')) {
      cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
    }
  }
  ~LdpTrainer() {
    drop_graphs();
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (side) { cudaStreamSynchronize(side); cudaStreamDestroy(side); }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (cudaEvent_t e : bucket_ev) cudaEventDestroy(e);
  }
};

namespace ldp {

static int trainer_tables(LdpTrainer* h, int n_train, int dim, int cos_first) {
  std::vector<float> betas, alphas, acp;
  ddpm_schedule_host(n_train, betas, alphas, acp);
  LDP_TRY(h->arena.alloc_t(&h->acp, n_train));
  LDP_CUDA_OK(cudaMemcpy(h->acp, acp.data(), (size_t)n_train * 4, cudaMemcpyHostToDevice));
  LDP_TRY(h->arena.alloc_t(&h->time_table, (size_t)n_train * dim));
  LDP_TRY(launch_sinusoid_table(h->time_table, n_train, dim, cos_first, 0));
  LDP_CUDA_OK(cudaStreamSynchronize(0));
  return LDP_OK;
}

// parameter cursor over the canonical blob: value pointer and the matching gradient pointer
struct PG { const float* w; float* g; };
struct PWalk {
  const float* p; float* g; int64_t pos = 0;
  PG take(int64_t n) { PG r{p + pos, g + pos}; pos += n; return r; }
};

static int gn_fwd(const Tn& x, int B, int P, int G, PG gamma, PG beta, const Tn* film, const Tn* res, Tn* y, cudaStream_t s) {
  GroupNormF32 n;
  n.x = x.v; n.ldx = x.ld; n.y = y->v; n.ldy = y->ld; n.B = B; n.P = P; n.C = x.c; n.G = G;
  n.gamma = gamma.w; n.beta = beta.w; n.act = 1;
  if (film) { n.film = 1; n.otab = film->v; n.ld_otab = film->ld; n.film_off = 0; }
  if (res) { n.res = res->v; n.ldres = res->ld; }
  return launch_groupnorm_f32(n, s);
}

static int gn_bwd(Tn* x, int B, int P, int G, PG gamma, PG beta, Tn* film, Tn* res, const Tn& y, cudaStream_t s) {
  GnBwdF32 q;
  q.x = x->v; q.dy = y.g; q.dx = x->g; q.B = B; q.P = P; q.C = x->c; q.G = G;
  q.gamma = gamma.w; q.beta = beta.w; q.dgamma = gamma.g; q.dbeta = beta.g;
  if (film) { q.film = film->v; q.dfilm = film->g; film->gset = true; }
  if (res && res->g) { q.dres = res->g; q.dres_acc = res->gset ? 1 : 0; res->gset = true; }
  x->gset = true;
  return launch_gn_bwd_f32(q, s);
}

// ------------------------------------------------------------------------------------------------
// planner: loss = mean((UNet(add_noise(x0, noise, t), t, cond) - noise)^2)   (agent/ldp_agent.py:113-126)
// ------------------------------------------------------------------------------------------------
struct CrbT {
  Tn *x1, *x2, *c1, *e, *h1, *c2, *r, *out;
  PG c1w, c1b, g1s, g1b, fw, fb, c2w, c2b, g2s, g2b, rw, rb;
  int cout, Tl;
  bool proj;
};

static int unet_loss_grad(LdpTrainer* h, int prec, const float* params, float* grads, const float* x0,
                          const float* noise, const int32_t* t, const float* cond, int B, int T, float weight,
                          float* loss_dev, cudaStream_t s) {
  TrainCtx cx{s, prec, &h->tc, &h->scratch_a, &h->scratch_w, h->side, h->ev_fork, h->ev_join, &h->scratch_a2, &h->scratch_w2};
  const LdpUnetConfig& c = h->ucfg;
  const int nl = c.n_levels, ds = c.step_embed_dim, dc = c.global_cond_dim, cd = ds + dc, D = c.input_dim, G = c.n_groups;
  LDP_CHECK(T > 0 && (T % (1 << (nl - 1))) == 0, LDP_ERR_UNSUPPORTED, "T must be a multiple of 2^(n_levels-1)");
  TrainWs& ws = h->ws;
  if (ws.key_a != B || ws.key_b != T) { ws.reset(); ws.key_a = B; ws.key_b = T; }
  ws.rewind();
  PWalk pw{params, grads};
  const int64_t k = c.kernel_size;
  PG t0w = pw.take((int64_t)ds * ds * 4), t0b = pw.take(ds * 4), t1w = pw.take((int64_t)ds * 4 * ds), t1b = pw.take(ds);

  // ---- forward
  LDP_TN(xin, ws, B * T, D, false);
  LDP_TRY(launch_add_noise(h->acp, x0, noise, t, xin->v, B, (long long)T * D, s));
  LDP_TN(sin_in, ws, B, ds, false);
  LDP_TRY(gather_rows(h->time_table, ds, t, sin_in->v, ds, B, s));
  LDP_TN(hid, ws, B, ds * 4, true);
  LDP_TN(hm, ws, B, ds * 4, true);
  LDP_TN(gbuf, ws, B, cd, true);
  LDP_TN(mg, ws, B, cd, true);
  Geo dense;
  LDP_TRY(conv_fwd(*sin_in, nullptr, dense, t0w.w, t0b.w, ds * 4, 0, nullptr, hid, cx));
  LDP_TRY(mish_fwd(*hid, hm, s));
  Tn temb = *gbuf;            // view: the first ds columns of g
  temb.c = ds;
  LDP_TRY(conv_fwd(*hm, nullptr, dense, t1w.w, t1b.w, ds, 0, nullptr, &temb, cx));
  LDP_CUDA_OK(cudaMemcpy2DAsync(gbuf->v + ds, (size_t)cd * 4, cond, (size_t)dc * 4, (size_t)dc * 4, B,
                                cudaMemcpyDeviceToDevice, s));
  LDP_TRY(mish_fwd(*gbuf, mg, s));

  std::vector<CrbT> blocks;
  auto crb_fwd = [&](Tn* x1, Tn* x2, int cout, bool proj, int Tl, Tn** out) -> int {
    CrbT b{};
    const int cin = x1->c + (x2 ? x2->c : 0);
    b.x1 = x1; b.x2 = x2; b.cout = cout; b.Tl = Tl; b.proj = proj;
    b.c1w = pw.take(k * cin * cout); b.c1b = pw.take(cout); b.g1s = pw.take(cout); b.g1b = pw.take(cout);
    b.fw = pw.take((int64_t)cd * 2 * cout); b.fb = pw.take(2 * cout);
    b.c2w = pw.take(k * (int64_t)cout * cout); b.c2b = pw.take(cout); b.g2s = pw.take(cout); b.g2b = pw.take(cout);
    if (proj) { b.rw = pw.take((int64_t)cin * cout); b.rb = pw.take(cout); }
    LDP_CHECK(proj || (!x2 && cin == cout), LDP_ERR_INVALID_ARG, "identity residual needs cin == cout");
    const int rows = B * Tl;
    Geo g5; g5.t_in = Tl; g5.t_out = Tl; g5.taps = (int)k; g5.pad = (int)k / 2;
    Geo g1; g1.t_in = Tl; g1.t_out = Tl;
    LDP_TN(c1, ws, rows, cout, true);
    LDP_TN(e, ws, B, 2 * cout, true);
    LDP_TN(h1, ws, rows, cout, true);
    LDP_TN(c2, ws, rows, cout, true);
    LDP_TN(o, ws, rows, cout, true);
    b.c1 = c1; b.e = e; b.h1 = h1; b.c2 = c2; b.out = o; b.r = nullptr;
    // the FiLM Dense and the residual projection only need the block's inputs: side stream, next to the first conv
    TrainCtx cside;
    LDP_TRY(fork_side(cx, &cside));
    LDP_TRY(conv_fwd(*mg, nullptr, dense, b.fw.w, b.fb.w, 2 * cout, 0, nullptr, e, cside));
    const Tn* res = x1;
    if (proj) {
      LDP_TN(r, ws, rows, cout, true);
      b.r = r;
      LDP_TRY(conv_fwd(*x1, x2, g1, b.rw.w, b.rb.w, cout, 0, nullptr, r, cside));
      res = r;
    }
    LDP_TRY(conv_fwd(*x1, x2, g5, b.c1w.w, b.c1b.w, cout, 0, nullptr, c1, cx));
    LDP_TRY(join_side(cx));
    LDP_TRY(gn_fwd(*c1, B, Tl, G, b.g1s, b.g1b, e, nullptr, h1, s));
    LDP_TRY(conv_fwd(*h1, nullptr, g5, b.c2w.w, b.c2b.w, cout, 0, nullptr, c2, cx));
    LDP_TRY(gn_fwd(*c2, B, Tl, G, b.g2s, b.g2b, nullptr, res, o, s));
    blocks.push_back(b);
    *out = o;
    return LDP_OK;
  };

  // the UNet's module creation order (params.py unet_spec): all CRBs, then Downsample1d_*, Upsample1d_*, final block.
  // CRB parameters are taken inside crb_fwd in that order; the resampling convs come later in the blob, so their
  // offsets are computed up front.
  int64_t crb_total = 0;
  {
    int ch = D;
    auto add = [&](int cin, int cout, bool proj) {
      crb_total += k * cin * cout + cout + 2 * cout + (int64_t)cd * 2 * cout + 2 * cout + k * (int64_t)cout * cout + cout +
                   2 * cout + (proj ? (int64_t)cin * cout + cout : 0);
    };
    for (int i = 0; i < nl; ++i) { add(ch, c.down_dims[i], true); add(c.down_dims[i], c.down_dims[i], false); ch = c.down_dims[i]; }
    add(ch, ch, false); add(ch, ch, false);
    for (int i = nl - 2; i >= 0; --i) { add(ch + c.down_dims[i + 1], c.down_dims[i], true); add(c.down_dims[i], c.down_dims[i], false); ch = c.down_dims[i]; }
  }
  const int64_t head = pw.pos;
  PWalk tail{params, grads, head + crb_total};
  std::vector<PG> down_w, down_b, up_w, up_b;
  for (int i = 0; i < nl - 1; ++i) { int64_t d = c.down_dims[i]; down_w.push_back(tail.take(3 * d * d)); down_b.push_back(tail.take(d)); }
  for (int i = 0; i < nl - 1; ++i) { int64_t d = c.down_dims[nl - 2 - i]; up_w.push_back(tail.take(4 * d * d)); up_b.push_back(tail.take(d)); }
  const int64_t d0 = c.down_dims[0];
  PG fcw = tail.take(k * d0 * d0), fcb = tail.take(d0), fgs = tail.take(d0), fgb = tail.take(d0);
  PG ow = tail.take(d0 * D), ob = tail.take(D);
  LDP_CHECK(tail.pos == h->n_params, LDP_ERR_PARAM_COUNT, "internal: trainer blob walk mismatch");

  struct Resamp { Tn* x; Tn* y; Geo g; PG w, b; int cout; };
  std::vector<Resamp> resamp;
  std::vector<Tn*> skips;
  Tn* cur = xin;
  for (int l = 0; l < nl; ++l) {
    const int Tl = T >> l, d = c.down_dims[l];
    LDP_TRY(crb_fwd(cur, nullptr, d, true, Tl, &cur));
    LDP_TRY(crb_fwd(cur, nullptr, d, false, Tl, &cur));
    skips.push_back(cur);
    if (l < nl - 1) {
      Geo g; g.t_in = Tl; g.t_out = Tl / 2; g.taps = 3; g.stride = 2; g.pad = 0;
      LDP_TN(y, ws, B * (Tl / 2), d, true);
      LDP_TRY(conv_fwd(*cur, nullptr, g, down_w[l].w, down_b[l].w, d, 0, nullptr, y, cx));
      resamp.push_back({cur, y, g, down_w[l], down_b[l], d});
      cur = y;
    }
  }
  {
    const int Tl = T >> (nl - 1), d = c.down_dims[nl - 1];
    LDP_TRY(crb_fwd(cur, nullptr, d, false, Tl, &cur));
    LDP_TRY(crb_fwd(cur, nullptr, d, false, Tl, &cur));
  }
  for (int u = 0; u < nl - 1; ++u) {
    const int lvl = nl - 1 - u, Tl = T >> lvl, d = c.down_dims[lvl - 1];
    Tn* skip = skips.back();
    skips.pop_back();
    LDP_TRY(crb_fwd(cur, skip, d, true, Tl, &cur));
    LDP_TRY(crb_fwd(cur, nullptr, d, false, Tl, &cur));
    Geo g; g.t_in = Tl; g.t_out = 2 * Tl; g.taps = 4; g.stride = 1; g.pad = 2; g.dil = 2;
    LDP_TN(y, ws, B * 2 * Tl, d, true);
    LDP_TRY(conv_fwd(*cur, nullptr, g, up_w[u].w, up_b[u].w, d, 0, nullptr, y, cx));
    resamp.push_back({cur, y, g, up_w[u], up_b[u], d});
    cur = y;
  }
  LDP_CHECK(pw.pos == head + crb_total, LDP_ERR_PARAM_COUNT, "internal: CRB blob walk mismatch");
  Geo g5; g5.t_in = T; g5.t_out = T; g5.taps = (int)k; g5.pad = (int)k / 2;
  Geo g1; g1.t_in = T; g1.t_out = T;
  LDP_CHECK(d0 % 8 == 0, LDP_ERR_UNSUPPORTED, "final Conv1dBlock uses 8 groups");
  LDP_TN(fc, ws, B * T, (int)d0, true);
  LDP_TN(ff, ws, B * T, (int)d0, true);
  LDP_TN(eps, ws, B * T, D, true);
  Tn* final_in = cur;
  LDP_TRY(conv_fwd(*final_in, nullptr, g5, fcw.w, fcb.w, (int)d0, 0, nullptr, fc, cx));
  LDP_TRY(gn_fwd(*fc, B, T, 8, fgs, fgb, nullptr, nullptr, ff, s));
  LDP_TRY(conv_fwd(*ff, nullptr, g1, ow.w, ob.w, D, 0, nullptr, eps, cx));

  // ---- loss and backward
  LDP_TRY(mse(*eps, noise, weight, loss_dev, s));
  eps->gset = true;
  LDP_TRY(conv_bwd(ff, nullptr, g1, ow.w, ow.g, ob.g, D, *eps, cx));
  LDP_TRY(gn_bwd(fc, B, T, 8, fgs, fgb, nullptr, nullptr, *ff, s));
  LDP_TRY(conv_bwd(final_in, nullptr, g5, fcw.w, fcw.g, fcb.g, (int)d0, *fc, cx));

  auto crb_bwd = [&](CrbT& b) -> int {
    const int Tl = b.Tl, cout = b.cout;
    Geo g5b; g5b.t_in = Tl; g5b.t_out = Tl; g5b.taps = (int)k; g5b.pad = (int)k / 2;
    Geo g1b; g1b.t_in = Tl; g1b.t_out = Tl;
    Tn* res = b.proj ? b.r : b.x1;
    LDP_TRY(gn_bwd(b.c2, B, Tl, G, b.g2s, b.g2b, nullptr, res, *b.out, s));
    LDP_TRY(conv_bwd(b.h1, nullptr, g5b, b.c2w.w, b.c2w.g, b.c2b.g, cout, *b.c2, cx));
    LDP_TRY(gn_bwd(b.c1, B, Tl, G, b.g1s, b.g1b, b.e, nullptr, *b.h1, s));
    // FiLM Dense backward entirely on the side stream: Mish(g)'s gradient is only ever accumulated there (in order)
    TrainCtx cside;
    LDP_TRY(fork_side(cx, &cside));
    LDP_TRY(conv_bwd(mg, nullptr, dense, b.fw.w, b.fw.g, b.fb.g, 2 * cout, *b.e, cside));
    LDP_TRY(conv_bwd(b.x1, b.x2, g5b, b.c1w.w, b.c1w.g, b.c1b.g, cout, *b.c1, cx));
    if (b.proj) LDP_TRY(conv_bwd(b.x1, b.x2, g1b, b.rw.w, b.rw.g, b.rb.g, cout, *b.r, cx));
    return LDP_OK;
  };
  // Gradient buckets (data-parallel all-reduce overlapped with the rest of the backward pass): the CRB parameters lie in the
  // blob in forward order and are finished in reverse, so "the blocks finished since the last bucket" is one contiguous range.
  h->buckets.clear();
  int n_ev = 0;
  int64_t bucket_hi = head + crb_total;               // end of the open bucket
  auto block_off = [&](int bi) { return (int64_t)(blocks[bi].c1w.g - grads); };
  auto close_bucket = [&](int bi) -> int {             // called after crb_bwd(blocks[bi]); closes [offset of block bi, bucket_hi)
    const int64_t lo = block_off(bi);
    if (bi == 0 || bucket_hi - lo < h->bucket_floats) return LDP_OK;
    LDP_TRY(join_side(cx));                            // weight gradients of these layers run on the side stream
    if ((int)h->bucket_ev.size() <= n_ev) {
      cudaEvent_t e;
      LDP_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->bucket_ev.push_back(e);
    }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    LDP_CUDA_OK(cudaStreamIsCapturing(s, &cs));
    if (cs == cudaStreamCaptureStatusActive) LDP_CUDA_OK(cudaEventRecordWithFlags(h->bucket_ev[n_ev], s, cudaEventRecordExternal));
    else LDP_CUDA_OK(cudaEventRecord(h->bucket_ev[n_ev], s));
    h->buckets.push_back({lo, bucket_hi - lo, n_ev});
    ++n_ev;
    bucket_hi = lo;
    return LDP_OK;
  };
  // reverse creation order: resampling convs are interleaved with the blocks exactly as in the forward
  {
    int bi = (int)blocks.size() - 1;
    int ri = (int)resamp.size() - 1;
    for (int u = nl - 2; u >= 0; --u) {          // up path, last to first
      Resamp& r = resamp[ri--];
      LDP_TRY(conv_bwd(r.x, nullptr, r.g, r.w.w, r.w.g, r.b.g, r.cout, *r.y, cx));
      LDP_TRY(crb_bwd(blocks[bi])); LDP_TRY(close_bucket(bi--));
      LDP_TRY(crb_bwd(blocks[bi])); LDP_TRY(close_bucket(bi--));
    }
    LDP_TRY(crb_bwd(blocks[bi])); LDP_TRY(close_bucket(bi--));              // mid
    LDP_TRY(crb_bwd(blocks[bi])); LDP_TRY(close_bucket(bi--));
    for (int l = nl - 1; l >= 0; --l) {          // down path
      if (l < nl - 1) {
        Resamp& r = resamp[ri--];
        LDP_TRY(conv_bwd(r.x, nullptr, r.g, r.w.w, r.w.g, r.b.g, r.cout, *r.y, cx));
      }
      LDP_TRY(crb_bwd(blocks[bi])); LDP_TRY(close_bucket(bi--));
      LDP_TRY(crb_bwd(blocks[bi])); LDP_TRY(close_bucket(bi--));
    }
  }
  // what is left completes with the call: the time MLP + the first blocks, and the resampling / final convolutions behind the CRBs
  h->buckets.push_back({0, bucket_hi, -1});
  h->buckets.push_back({head + crb_total, h->n_params - (head + crb_total), -1});
  // conditioning path: mg = Mish(g); g[:, :ds] = Dense_1(Mish(Dense_0(sinusoid)))
  LDP_TRY(join_side(cx));
  LDP_TRY(mish_bwd(gbuf, *mg, s));
  Tn temb_g = *gbuf;
  temb_g.c = ds;
  LDP_TRY(conv_bwd(hm, nullptr, dense, t1w.w, t1w.g, t1b.g, ds, temb_g, cx));
  LDP_TRY(mish_bwd(hid, *hm, s));
  LDP_TRY(conv_bwd(sin_in, nullptr, dense, t0w.w, t0w.g, t0b.g, ds * 4, *hid, cx));
  return join_side(cx);
}

// ------------------------------------------------------------------------------------------------
// IDM: loss = mean((MLPDiffusion(s||s', add_noise(a0, noise, t), t) - noise)^2)   (agent/ldp_agent.py:128-139)
// ------------------------------------------------------------------------------------------------
static int idm_loss_grad_impl(LdpTrainer* h, int prec, const float* params, float* grads, const float* sdev, const float* a0,
                              const float* noise, const int32_t* t, int N, float weight, float* loss_dev, cudaStream_t s);
static int idm_loss_grad(LdpTrainer* h, int prec, const float* params, float* grads, const float* sdev, const float* a0,
                         const float* noise, const int32_t* t, int N, float weight, float* loss_dev, cudaStream_t s) {
  h->buckets.assign(1, LdpTrainer::Bucket{0, h->n_params, -1});      // 7.6 MB: one exchange
  return idm_loss_grad_impl(h, prec, params, grads, sdev, a0, noise, t, N, weight, loss_dev, s);
}
static int idm_loss_grad_impl(LdpTrainer* h, int prec, const float* params, float* grads, const float* sdev, const float* a0,
                         const float* noise, const int32_t* t, int N, float weight, float* loss_dev, cudaStream_t s) {
  TrainCtx cx{s, prec, &h->tc, &h->scratch_a, &h->scratch_w, h->side, h->ev_fork, h->ev_join, &h->scratch_a2, &h->scratch_w2};
  const LdpIdmConfig& c = h->icfg;
  const int A = c.action_dim, S2 = 2 * c.obs_dim, H = c.hidden_dim, td = c.time_dim;
  TrainWs& ws = h->ws;
  if (ws.key_a != N) { ws.reset(); ws.key_a = N; }
  ws.rewind();
  PWalk pw{params, grads};
  Geo dense;
  const int co = c.cond_hidden[c.n_cond_layers - 1], in_dim = A + S2 + co;

  // ---- forward
  LDP_TN(ff, ws, N, td, false);
  LDP_TRY(gather_rows(h->time_table, td, t, ff->v, td, N, s));
  LDP_TN(xin, ws, N, in_dim, true);       // [noisy a | s||s' | cond]; only the cond columns carry a gradient
  LDP_TN(a_noisy, ws, N, A, false);
  LDP_TRY(launch_add_noise(h->acp, a0, noise, t, a_noisy->v, N, A, s));
  LDP_CUDA_OK(cudaMemcpy2DAsync(xin->v, (size_t)in_dim * 4, a_noisy->v, (size_t)A * 4, (size_t)A * 4, N,
                                cudaMemcpyDeviceToDevice, s));
  LDP_CUDA_OK(cudaMemcpy2DAsync(xin->v + A, (size_t)in_dim * 4, sdev, (size_t)S2 * 4, (size_t)S2 * 4, N,
                                cudaMemcpyDeviceToDevice, s));
  struct CondL { Tn* x; Tn* y; Tn* ym; PG w, b; int n; };
  std::vector<CondL> cl;
  Tn cond_view = *xin;                    // view of the cond columns of xin
  cond_view.v = xin->v + A + S2; cond_view.g = xin->g + A + S2; cond_view.c = co;
  {
    Tn* cur = ff;
    int ch = td;
    for (int i = 0; i < c.n_cond_layers; ++i) {
      const int n = c.cond_hidden[i];
      CondL L{};
      L.x = cur; L.n = n;
      L.w = pw.take((int64_t)ch * n); L.b = pw.take(n);
      const bool last = i + 1 == c.n_cond_layers;
      if (last) {
        L.y = nullptr; L.ym = nullptr;
        LDP_TRY(conv_fwd(*cur, nullptr, dense, L.w.w, L.b.w, n, 0, nullptr, &cond_view, cx));
      } else {
        LDP_TN(y, ws, N, n, true);
        LDP_TN(ym, ws, N, n, true);
        L.y = y; L.ym = ym;
        LDP_TRY(conv_fwd(*cur, nullptr, dense, L.w.w, L.b.w, n, 0, nullptr, y, cx));
        LDP_TRY(mish_fwd(*y, ym, s));
        cur = ym;
      }
      cl.push_back(L);
      ch = n;
    }
  }
  PG w0 = pw.take((int64_t)in_dim * H), b0 = pw.take(H);
  LDP_TN(h0, ws, N, H, true);
  LDP_TRY(conv_fwd(*xin, nullptr, dense, w0.w, b0.w, H, 0, nullptr, h0, cx));
  struct BlkT { Tn *hin, *hn, *u, *hout; PG ln_s, ln_b, w1, b1, w2, b2; };
  std::vector<BlkT> blk;
  Tn* hc = h0;
  for (int b = 0; b < c.n_blocks; ++b) {
    BlkT q{};
    q.hin = hc;
    q.ln_s = pw.take(H); q.ln_b = pw.take(H);
    q.w1 = pw.take((int64_t)H * 4 * H); q.b1 = pw.take(4 * H);
    q.w2 = pw.take((int64_t)4 * H * H); q.b2 = pw.take(H);
    LDP_TN(hn, ws, N, H, true);
    LDP_TN(u, ws, N, 4 * H, true);
    LDP_TN(ho, ws, N, H, true);
    q.hn = hn; q.u = u; q.hout = ho;
    LDP_TRY(launch_layernorm_f32(hc->v, hn->v, N, H, q.ln_s.w, q.ln_b.w, 1e-6f, 0, s));
    LDP_TRY(conv_fwd(*hn, nullptr, dense, q.w1.w, q.b1.w, 4 * H, 1, nullptr, u, cx));
    LDP_TRY(conv_fwd(*u, nullptr, dense, q.w2.w, q.b2.w, H, 0, hc, ho, cx));
    blk.push_back(q);
    hc = ho;
  }
  PG wout = pw.take((int64_t)H * A), bout = pw.take(A);
  LDP_CHECK(pw.pos == h->n_params, LDP_ERR_PARAM_COUNT, "internal: trainer blob walk mismatch");
  LDP_TN(hr, ws, N, H, true);
  LDP_TN(eps, ws, N, A, true);
  relu_fwd_kernel<<<ew_blocks((long long)N * H), 256, 0, s>>>(hc->v, hr->v, (long long)N * H);
  LDP_LAUNCH_OK();
  LDP_TRY(conv_fwd(*hr, nullptr, dense, wout.w, bout.w, A, 0, nullptr, eps, cx));

  // ---- loss and backward
  LDP_TRY(mse(*eps, noise, weight, loss_dev, s));
  eps->gset = true;
  LDP_TRY(conv_bwd(hr, nullptr, dense, wout.w, wout.g, bout.g, A, *eps, cx));
  relu_bwd_kernel<<<ew_blocks((long long)N * H), 256, 0, s>>>(hc->v, hr->g, hc->g, (long long)N * H, 0);
  LDP_LAUNCH_OK();
  hc->gset = true;
  for (int b = c.n_blocks - 1; b >= 0; --b) {
    BlkT& q = blk[b];
    // hout = Dense_1(u) + hin
    LDP_TRY(conv_bwd(q.u, nullptr, dense, q.w2.w, q.w2.g, q.b2.g, H, *q.hout, cx));
    relu_bwd_kernel<<<ew_blocks((long long)N * 4 * H), 256, 0, s>>>(q.u->v, q.u->g, q.u->g, (long long)N * 4 * H, 0);
    LDP_LAUNCH_OK();
    LDP_TRY(conv_bwd(q.hn, nullptr, dense, q.w1.w, q.w1.g, q.b1.g, 4 * H, *q.u, cx));
    LDP_TRY(launch_ln_bwd_f32(q.hin->v, q.hn->g, q.hout->g, q.hin->g, N, H, q.ln_s.w, q.ln_s.g, q.ln_b.g, 1e-6f, s));
    q.hin->gset = true;
  }
  // Dense_0: weight gradient over the whole input, data gradient only for the cond columns
  {
    LDP_TRY(conv_wgrad_side(*xin, nullptr, dense, w0.g, b0.g, H, *h0, cx));
    LDP_TRY(conv_dgrad(&cond_view, A + S2, in_dim, dense, w0.w, H, *h0, cx));
  }
  for (int i = c.n_cond_layers - 1; i >= 0; --i) {
    CondL& L = cl[i];
    const bool last = i + 1 == c.n_cond_layers;
    if (last) {
      LDP_TRY(conv_bwd(L.x, nullptr, dense, L.w.w, L.w.g, L.b.g, L.n, cond_view, cx));
    } else {
      LDP_TRY(mish_bwd(L.y, *L.ym, s));
      LDP_TRY(conv_bwd(L.x, nullptr, dense, L.w.w, L.w.g, L.b.g, L.n, *L.y, cx));
    }
  }
  return join_side(cx);
}

}  // namespace ldp

namespace ldp {

// Runs `body` (which only enqueues work on `s` and the trainer's side stream): directly the first time a key is seen
// (workspace, scratch and tensor-map allocations happen then), captured into a graph the second time, replayed after.
template <typename F>
static int run_step(LdpTrainer* h, int prec, const LdpTrainer::GraphKey& key, cudaStream_t s, F&& body) {
  if (prec != LDP_PREC_BF16 || !h->use_graph) return body(s);
  // one shape at a time: the workspace is rebuilt when the shape changes, which would leave older graphs dangling
  if (!h->graphs.empty() && h->graphs.find(key) == h->graphs.end()) {
    h->drop_graphs();
    cudaDeviceSynchronize();
    h->tc.reset();
  }
  LdpTrainer::GraphSlot& slot = h->graphs[key];
  if (slot.exec) {
    LDP_CUDA_OK(cudaGraphLaunch(slot.exec, s));
    return LDP_OK;
  }
  // two direct passes first: the first grows the scratch buffers layer by layer (tensor maps built early in it point at
  // buffers that were reallocated later), the second builds every tensor map against the final addresses
  if (slot.warm < 2) {
    ++slot.warm;
    return body(s);
  }
  // capture on a private stream (the caller's may be the legacy default stream, which cannot be captured)
  if (!h->cap_stream) LDP_CUDA_OK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  const long long before = launch_count_get();
  cudaError_t e = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal);
  int st = LDP_OK;
  if (e == cudaSuccess) {
    st = body(h->cap_stream);
    e = cudaStreamEndCapture(h->cap_stream, &graph);
  }
  count_launch((int)(before - launch_count_get()));     // captured launches did not execute
  if (st == LDP_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&slot.exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  if (st != LDP_OK || e != cudaSuccess || !slot.exec) {
    slot.exec = nullptr;
    cudaGetLastError();
    h->use_graph = false;                               // fall back to direct launches for good
    return body(s);
  }
  LDP_CUDA_OK(cudaGraphLaunch(slot.exec, s));
  return LDP_OK;
}

// copy a caller tensor into the trainer's fixed staging buffer i
static int stage_in(LdpTrainer* h, int i, const void* src, size_t bytes, cudaStream_t s, const void** out) {
  LDP_TRY(h->stage[i].ensure(bytes));
  LDP_CUDA_OK(cudaMemcpyAsync(h->stage[i].p, src, bytes, cudaMemcpyDeviceToDevice, s));
  *out = h->stage[i].p;
  return LDP_OK;
}

}  // namespace ldp

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int ldp_unet_trainer_create(const LdpUnetConfig* cfg, LdpTrainer** out) {
  LDP_CHECK(cfg && out, LDP_ERR_INVALID_ARG, "null argument");
  const int64_t n = ldp_unet_param_count(cfg);
  LDP_CHECK(n > 0, LDP_ERR_INVALID_ARG, "bad planner config");
  LDP_CHECK(cfg->kernel_size == 5 && cfg->n_levels >= 1 && cfg->n_levels <= 6, LDP_ERR_UNSUPPORTED, "bad planner config");
  std::unique_ptr<LdpTrainer> h(new LdpTrainer());
  h->kind = 0; h->ucfg = *cfg; h->n_params = n;
  LDP_TRY(trainer_tables(h.get(), cfg->n_train_steps, cfg->step_embed_dim, /*cos_first=*/0));
  *out = h.release();
  return LDP_OK;
}

int ldp_idm_trainer_create(const LdpIdmConfig* cfg, LdpTrainer** out) {
  LDP_CHECK(cfg && out, LDP_ERR_INVALID_ARG, "null argument");
  const int64_t n = ldp_idm_param_count(cfg);
  LDP_CHECK(n > 0, LDP_ERR_INVALID_ARG, "bad IDM config");
  std::unique_ptr<LdpTrainer> h(new LdpTrainer());
  h->kind = 1; h->icfg = *cfg; h->n_params = n;
  LDP_TRY(trainer_tables(h.get(), cfg->n_train_steps, cfg->time_dim, /*cos_first=*/1));
  *out = h.release();
  return LDP_OK;
}

int ldp_trainer_destroy(LdpTrainer* h) {
  delete h;
  return LDP_OK;
}

int ldp_unet_loss_grad(LdpTrainer* h, int precision, const float* params_dev, float* grads_dev, const float* x0_dev,
                       const float* noise_dev, const int32_t* t_dev, const float* cond_dev, int B, int T,
                       float loss_weight, float* loss_dev, void* cuda_stream) {
  LDP_CHECK(h && h->kind == 0, LDP_ERR_INVALID_ARG, "not a planner trainer handle");
  LDP_CHECK(params_dev && grads_dev && x0_dev && noise_dev && t_dev && cond_dev && loss_dev && B > 0 && T > 0,
            LDP_ERR_INVALID_ARG, "bad arguments");
  LDP_CHECK(precision == LDP_PREC_FP32 || precision == LDP_PREC_BF16, LDP_ERR_INVALID_ARG, "bad precision");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  if (precision == LDP_PREC_BF16 && h->use_graph) {
    const size_t nx = (size_t)B * T * h->ucfg.input_dim * 4;
    const void *x0s, *zs, *ts, *cs;
    LDP_TRY(stage_in(h, 0, x0_dev, nx, s, &x0s));
    LDP_TRY(stage_in(h, 1, noise_dev, nx, s, &zs));
    LDP_TRY(stage_in(h, 2, t_dev, (size_t)B * 4, s, &ts));
    LDP_TRY(stage_in(h, 3, cond_dev, (size_t)B * h->ucfg.global_cond_dim * 4, s, &cs));
    x0_dev = (const float*)x0s; noise_dev = (const float*)zs; t_dev = (const int32_t*)ts; cond_dev = (const float*)cs;
  }
  LdpTrainer::GraphKey key{B, T, params_dev, grads_dev, loss_dev, loss_weight};
  return run_step(h, precision, key, s, [&](cudaStream_t st) {
    return unet_loss_grad(h, precision, params_dev, grads_dev, x0_dev, noise_dev, t_dev, cond_dev, B, T, loss_weight, loss_dev, st);
  });
}

int ldp_idm_loss_grad(LdpTrainer* h, int precision, const float* params_dev, float* grads_dev, const float* s_dev, const float* a0_dev,
                      const float* noise_dev, const int32_t* t_dev, int N, float loss_weight, float* loss_dev,
                      void* cuda_stream) {
  LDP_CHECK(h && h->kind == 1, LDP_ERR_INVALID_ARG, "not an IDM trainer handle");
  LDP_CHECK(params_dev && grads_dev && s_dev && a0_dev && noise_dev && t_dev && loss_dev && N > 0, LDP_ERR_INVALID_ARG,
            "bad arguments");
  LDP_CHECK(precision == LDP_PREC_FP32 || precision == LDP_PREC_BF16, LDP_ERR_INVALID_ARG, "bad precision");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  if (precision == LDP_PREC_BF16 && h->use_graph) {
    const void *ss, *as, *zs, *ts;
    LDP_TRY(stage_in(h, 0, s_dev, (size_t)N * 2 * h->icfg.obs_dim * 4, s, &ss));
    LDP_TRY(stage_in(h, 1, a0_dev, (size_t)N * h->icfg.action_dim * 4, s, &as));
    LDP_TRY(stage_in(h, 2, noise_dev, (size_t)N * h->icfg.action_dim * 4, s, &zs));
    LDP_TRY(stage_in(h, 3, t_dev, (size_t)N * 4, s, &ts));
    s_dev = (const float*)ss; a0_dev = (const float*)as; noise_dev = (const float*)zs; t_dev = (const int32_t*)ts;
  }
  LdpTrainer::GraphKey key{N, 0, params_dev, grads_dev, loss_dev, loss_weight};
  return run_step(h, precision, key, s, [&](cudaStream_t st) {
    return idm_loss_grad(h, precision, params_dev, grads_dev, s_dev, a0_dev, noise_dev, t_dev, N, loss_weight, loss_dev, st);
  });
}

int ldp_trainer_grad_buckets(LdpTrainer* h, int64_t* offsets, int64_t* lengths, int32_t* has_event, int max_buckets, int* n_buckets) {
  LDP_CHECK(h && offsets && lengths && has_event && n_buckets, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_CHECK((int)h->buckets.size() <= max_buckets, LDP_ERR_INVALID_ARG, "output arrays too small");
  for (size_t i = 0; i < h->buckets.size(); ++i) {
    offsets[i] = h->buckets[i].off;
    lengths[i] = h->buckets[i].len;
    has_event[i] = h->buckets[i].ev >= 0 ? 1 : 0;
  }
  *n_buckets = (int)h->buckets.size();
  return LDP_OK;
}

int ldp_trainer_wait_bucket(LdpTrainer* h, int bucket, void* cuda_stream) {
  LDP_CHECK(h && bucket >= 0 && bucket < (int)h->buckets.size() && h->buckets[bucket].ev >= 0, LDP_ERR_INVALID_ARG,
            "bucket has no event (it completes with the loss/gradient call itself)");
  LDP_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)cuda_stream, h->bucket_ev[h->buckets[bucket].ev], 0));
  return LDP_OK;
}

int ldp_adam_update(float* params_dev, const float* grads_dev, float* mu_dev, float* nu_dev, uint64_t n, float lr,
                    float b1, float b2, float eps, int64_t count, float grad_scale, void* cuda_stream) {
  LDP_CHECK(params_dev && grads_dev && mu_dev && nu_dev && n > 0 && count >= 1, LDP_ERR_INVALID_ARG, "bad arguments");
  const float bc1 = (float)(1.0 - std::pow((double)b1, (double)count));
  const float bc2 = (float)(1.0 - std::pow((double)b2, (double)count));
  LDP_CHECK((((uintptr_t)params_dev | (uintptr_t)grads_dev | (uintptr_t)mu_dev | (uintptr_t)nu_dev) & 15) == 0,
            LDP_ERR_INVALID_ARG, "adam: buffers must be 16-byte aligned");
  adam_kernel<<<148 * 16, 256, 0, (cudaStream_t)cuda_stream>>>(params_dev, grads_dev, mu_dev, nu_dev,
                                                                              (long long)n, lr, b1, b2, eps, bc1, bc2,
                                                                              grad_scale);
  LDP_LAUNCH_OK();
  return LDP_OK;
}

}  // extern "C"
