// VAE encoder handle: FlaxAutoencoderKL.encode(x).latent_dist.mean (diffusers 0.27.2, un-vendored; call sites reference
// agent/ldp_agent.py:46-64 and process_sdvae_data.py:70-73, config model/stable_vae_model.yaml:4-16).
//
//   conv_in 3x3 -> [ResnetBlock x layers_per_block, Downsample (pad (0,1), 3x3 stride 2)] per block (no downsample after
//   the last) -> mid (Resnet, 1-head self-attention, Resnet) -> GroupNorm -> swish -> conv_out 3x3 -> quant_conv 1x1 ->
//   first `latent_channels` channels (= the mean) -> optional (z - min)/(max - min)*2 - 1   (agent/ldp_agent.py:62)
//
// Data layout: NHWC throughout, exactly the reference's.  The residual stream lives in HBM as float32 plus a bf16 copy
// (the A operand of the next tensor-core convolution); GroupNorm reads float32 and writes the normalised, swish-ed
// activation as bf16.  Images are processed in chunks of <= 256 so that the 64x64x128 level (2 MB per image in f32)
// stays a few hundred MB.
//
// Kernels:
//  * every 3x3 / 1x1 / strided convolution and the attention projections: the tcgen05 implicit-GEMM kernel of
//    tc_gemm.cu - a tile is 128 output pixels (whole image rows), a tap is a TMA coordinate offset into the
//    (C, W, H, B) tensor map, the zero padding is TMA's out-of-bounds fill, the stride-2 convolution is a tensor map
//    with traversal stride 2;
//  * GroupNorm(32, eps 1e-6) + swish: vectorised float4 loads, per-thread partial sums, shared-memory + one global
//    atomic per (block, group), then an elementwise apply pass (warp-uniform statistics);
//  * conv_in (3 input channels - too thin for TMA / tensor cores), the 64-token attention core and the 8-channel
//    quant_conv: small SIMT kernels.
//  * LDP_PREC_FP32: the same program with every contraction in fp32 FFMA (direct convolution kernel) - parity gate.
#include <algorithm>
#include <cmath>

#include "net_common.h"

namespace ldp {

// ------------------------------------------------------------------------------------------------
// SIMT kernels
// ------------------------------------------------------------------------------------------------
// conv_in: 3x3, pad 1, Cin = in_ch (3) -> C0.  One thread per (pixel, 4 output channels).  Pixel normalisation
// x/255*2-1 (utils/data_utils.py:11; process_sdvae_data.py:89-90) is applied on load for uint8 input.
template <typename PixT>
__global__ void __launch_bounds__(256) vae_conv_in_kernel(const PixT* __restrict__ img, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out_f32,
                                                          __nv_bfloat16* __restrict__ out_bf16, int B, int S, int cin, int c0) {
  extern __shared__ float ws[];                       // [9*cin][c0]
  for (int i = threadIdx.x; i < 9 * cin * c0; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int cq = c0 / 4;
  const long long total = (long long)B * S * S * cq;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx % cq);
    const long long pix = idx / cq;
    const int x = (int)(pix % S), y = (int)((pix / S) % S);
    const long long b = pix / ((long long)S * S);
    float4 acc = *reinterpret_cast<const float4*>(bias + 4 * q);
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= S) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        if (xx < 0 || xx >= S) continue;
        const PixT* ip = img + ((b * S + yy) * S + xx) * cin;
        for (int ci = 0; ci < cin; ++ci) {
          float v = (float)ip[ci];
          if (sizeof(PixT) == 1) v = v / 255.f * 2.f - 1.f;
          const float4 wv = *reinterpret_cast<const float4*>(ws + ((dy * 3 + dx) * cin + ci) * c0 + 4 * q);
          acc.x = fmaf(v, wv.x, acc.x); acc.y = fmaf(v, wv.y, acc.y); acc.z = fmaf(v, wv.z, acc.z); acc.w = fmaf(v, wv.w, acc.w);
        }
      }
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + pix * c0 + 4 * q) = acc;
    if (out_bf16) {
      uint2 u;
      u.x = pack_bf16x2(acc.x, acc.y);
      u.y = pack_bf16x2(acc.z, acc.w);
      *reinterpret_cast<uint2*>(out_bf16 + pix * c0 + 4 * q) = u;
    }
  }
}

// conv_in on the tensor cores (bf16 path): im2col of the 3x3xCin input window into a K-major bf16 matrix [pixels][64] that the
// tcgen05 GEMM reads through TMA like any dense operand.  uint8 pixels are written as x - 128 - integers -128..127 are exact in
// bf16 - and the reference's x/255*2-1 (utils/data_utils.py:11, process_sdvae_data.py:89-90) is folded into the packed weights:
//   sum_taps w (x/127.5 - 1) = sum_taps (w/127.5) (x - 128)  +  sum_{valid taps} (0.5/127.5) sum_ci w
// where the second term needs "is this tap inside the image" (zero padding pads the NORMALISED image): columns [9 Cin, 9 Cin + 9)
// hold those nine indicators.  (Centring keeps the operand in [-1, 1] like the normalised pixel: with raw 0..255 the two terms
// are each ~3x the result and their bf16 weight rounding does not cancel.)  float32 input (already normalised) is written as
// is and the indicator weights are zero.
template <typename PixT>
__global__ void __launch_bounds__(256) vae_im2col_in_kernel(const PixT* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                            long long npix, int S, int cin) {
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < npix; pix += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % S), y = (int)((pix / S) % S);
    const long long b = pix / ((long long)S * S);
    __align__(16) __nv_bfloat16 row[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) row[i] = __float2bfloat16(0.f);
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        const bool ok = yy >= 0 && yy < S && xx >= 0 && xx < S;
        const int tap = dy * 3 + dx;
        if (ok) {
          const PixT* ip = img + ((b * S + yy) * S + xx) * cin;
          for (int ci = 0; ci < cin; ++ci) row[tap * cin + ci] = __float2bfloat16(sizeof(PixT) == 1 ? (float)ip[ci] - 128.f : (float)ip[ci]);
          row[9 * cin + tap] = __float2bfloat16(1.f);
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + pix * 64);
    const uint4* src = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = src[i];
  }
}

// Direct NHWC convolution in fp32 (parity path): out[b,y,x,co] = bias[co] + sum in[b, y*s+dy-pad, x*s+dx-pad, ci] w[dy,dx,ci,co]
// (+ res).  One thread per (8 consecutive output pixels of a row, output channel); weights are read coalesced over co.
__global__ void __launch_bounds__(128) vae_conv_f32_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                           const float* __restrict__ bias, const float* __restrict__ res,
                                                           float* __restrict__ out, int B, int Hin, int Win, int Hout, int Wout,
                                                           int cin, int cout, int k, int stride, int pad) {
  const int co = blockIdx.y * blockDim.x + threadIdx.x;
  const int xg = (Wout + 7) / 8;
  const long long job = blockIdx.x;                  // (b, y, xgroup)
  const int x0 = (int)(job % xg) * 8;
  const int y = (int)((job / xg) % Hout);
  const long long b = job / ((long long)xg * Hout);
  if (co >= cout || b >= B) return;
  float acc[8];
  const float bv = bias ? bias[co] : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = bv;
  for (int dy = 0; dy < k; ++dy) {
    const int yy = y * stride + dy - pad;
    if (yy < 0 || yy >= Hin) continue;
    for (int dx = 0; dx < k; ++dx) {
      const float* wp = w + (long long)((dy * k + dx) * cin) * cout + co;
      for (int ci = 0; ci < cin; ++ci) {
        const float wv = wp[(long long)ci * cout];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int xx = (x0 + i) * stride + dx - pad;
          if (x0 + i < Wout && xx >= 0 && xx < Win) acc[i] = fmaf(in[((b * Hin + yy) * Win + xx) * cin + ci], wv, acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (x0 + i < Wout) {
      const long long o = ((b * Hout + y) * Wout + x0 + i) * cout + co;
      out[o] = acc[i] + (res ? res[o] : 0.f);
    }
  }
}

// GroupNorm statistics, deterministic (same image -> same bits whatever the batch): x (B, P, C) f32.
// Pass 1, grid (B, slabs): a thread owns one channel quad (requires 256 % (C/4) == 0), accumulates (sum, sum sq) over its
// pixels of the slab; the block combines its threads per group in a fixed order and writes part[b][slab][g].
// Pass 2: one thread per (b, g) adds the slabs in order and stores (mean, rstd).
__global__ void __launch_bounds__(256) vae_gn_stats_kernel(const float* __restrict__ x, float* __restrict__ part, int P, int C,
                                                           int G) {
  __shared__ float ps[256][8];
  const int b = blockIdx.x;
  const int cq = C / 4, cpg = C / G;
  const long long total = (long long)P * cq;
  long long per = (total + gridDim.y - 1) / gridDim.y;
  per = (per + cq - 1) / cq * cq;                              // whole pixels per slab: a thread's quad index is fixed
  const long long lo = blockIdx.y * per, hi = min(total, lo + per);
  const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * P * C);
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float4 v = xb[i];
    s[0] += v.x; ss[0] = fmaf(v.x, v.x, ss[0]);
    s[1] += v.y; ss[1] = fmaf(v.y, v.y, ss[1]);
    s[2] += v.z; ss[2] = fmaf(v.z, v.z, ss[2]);
    s[3] += v.w; ss[3] = fmaf(v.w, v.w, ss[3]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ps[threadIdx.x][k] = s[k];
    ps[threadIdx.x][4 + k] = ss[k];
  }
  __syncthreads();
  if (threadIdx.x < G) {
    // thread t holds quad t % cq (lo is a multiple of cq and 256 % cq == 0): visit only the threads / components of
    // group g, in a fixed order
    const int g = threadIdx.x;
    const int q_lo = (g * cpg) >> 2, q_hi = ((g + 1) * cpg + 3) >> 2;
    float a = 0.f, a2 = 0.f;
    for (int q = q_lo; q < q_hi; ++q) {
      for (int t = q; t < 256; t += cq) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if ((4 * q + k) / cpg == g) {
            a += ps[t][k];
            a2 += ps[t][4 + k];
          }
        }
      }
    }
    float* o = part + (((long long)b * gridDim.y + blockIdx.y) * G + g) * 2;
    o[0] = a;
    o[1] = a2;
  }
}

__global__ void vae_gn_final_kernel(const float* __restrict__ part, float* __restrict__ mr, int n_bg, int G, int slabs, float inv_n,
                                    float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // (b, g)
  if (i >= n_bg) return;
  const int b = i / G, g = i % G;
  float a = 0.f, a2 = 0.f;
  for (int sl = 0; sl < slabs; ++sl) {
    const float* p = part + (((long long)b * slabs + sl) * G + g) * 2;
    a += p[0];
    a2 += p[1];
  }
  const float mu = a * inv_n;
  mr[2 * i] = mu;
  mr[2 * i + 1] = rsqrtf(fmaxf(a2 * inv_n - mu * mu, 0.f) + eps);
}

// y = act((x - mean) rstd gamma + beta); act: 0 none, 1 swish.  Output f32 and/or bf16.
// A pure streaming pass (16 B in, 8-24 B out per thread-iteration), so it must run at HBM speed: every thread keeps one
// channel quad for the whole grid-stride loop (the stride is a multiple of C/4), i.e. gamma/beta/group are loop
// invariants and the only per-iteration index work is one 32-bit division for the image index.  (The first version
// did two 64-bit divisions and eight scalar parameter loads per float4 and reached 2.5 TB/s.)
__global__ void __launch_bounds__(256) vae_gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mr,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* __restrict__ y_f32, __nv_bfloat16* __restrict__ y_bf16,
                                                           long long B, int P, int C, int G, int act) {
  const int cq = C / 4, cpg = C / G;
  const unsigned nthreads = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned rows = (unsigned)(B * P);
  if (nthreads % cq == 0 && cpg % 4 == 0) {
    const unsigned q = tid % cq, row0 = tid / cq, rstep = nthreads / cq;
    const int g = (4 * q) / cpg;
    const float4 ga = reinterpret_cast<const float4*>(gamma)[q], be = reinterpret_cast<const float4*>(beta)[q];
    const float4* x4 = reinterpret_cast<const float4*>(x);
#pragma unroll 4
    for (unsigned row = row0; row < rows; row += rstep) {
      const unsigned b = row / (unsigned)P;
      const size_t i = (size_t)row * cq + q;
      const float4 v = __ldcs(x4 + i);
      const float2 m = __ldg(reinterpret_cast<const float2*>(mr) + (b * G + g));
      float o0 = fmaf((v.x - m.x) * m.y, ga.x, be.x), o1 = fmaf((v.y - m.x) * m.y, ga.y, be.y);
      float o2 = fmaf((v.z - m.x) * m.y, ga.z, be.z), o3 = fmaf((v.w - m.x) * m.y, ga.w, be.w);
      if (act == 1) {
        o0 = __fdividef(o0, 1.f + __expf(-o0)); o1 = __fdividef(o1, 1.f + __expf(-o1));
        o2 = __fdividef(o2, 1.f + __expf(-o2)); o3 = __fdividef(o3, 1.f + __expf(-o3));
      }
      if (y_f32) reinterpret_cast<float4*>(y_f32)[i] = make_float4(o0, o1, o2, o3);
      if (y_bf16) {
        uint2 u;
        u.x = pack_bf16x2(o0, o1);
        u.y = pack_bf16x2(o2, o3);
        reinterpret_cast<uint2*>(y_bf16)[i] = u;
      }
    }
    return;
  }
  const long long total = B * P * cq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq);
    const long long b = i / ((long long)P * cq);
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float in[4] = {v.x, v.y, v.z, v.w}, o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = 4 * q + k, g = c / cpg;
      const float2 m = *reinterpret_cast<const float2*>(mr + (b * G + g) * 2);
      float t = fmaf((in[k] - m.x) * m.y, gamma[c], beta[c]);
      if (act == 1) t = t / (1.f + __expf(-t));
      o[k] = t;
    }
    if (y_f32) reinterpret_cast<float4*>(y_f32)[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (y_bf16) {
      uint2 u;
      u.x = pack_bf16x2(o[0], o[1]);
      u.y = pack_bf16x2(o[2], o[3]);
      reinterpret_cast<uint2*>(y_bf16)[i] = u;
    }
  }
}

// Image-major variant of the pass above (the one the tensor-core path uses): a block owns `rows_per_block` consecutive pixels of ONE image,
// a thread one channel quad of every (256 / (C/4))-th pixel of them.  mean / rstd / gamma / beta are loop invariants (the grid-stride
// version re-derived the image index with an integer division and re-read the statistics per float4: 85 instructions per float4,
// 47 % issue-active, 4.6 TB/s; ncu r2c); same arithmetic per element, so the results are bit-identical.  Four independent 16-byte loads
// are in flight per thread.
template <int ACT>
__global__ void __launch_bounds__(256) vae_gn_apply_img_kernel(const float* __restrict__ x, const float* __restrict__ part,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* __restrict__ y_f32, __nv_bfloat16* __restrict__ y_bf16, int P, int C,
                                                               int G, int rows_per_block, int slabs, float inv_n, float eps) {
  const int cq = C >> 2, cpg = C / G;
  const int q = threadIdx.x % cq, rl = threadIdx.x / cq, rpb = 256 / cq;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_block, r_end = min(P, r_begin + rows_per_block);
  // (mean, rstd) of this image's groups from the per-slab partial sums, in the fixed slab order of vae_gn_final_kernel (whose launch this
  // replaces: same values, bit for bit)
  __shared__ float2 stat[64];
  if (threadIdx.x < G) {
    float a = 0.f, a2 = 0.f;
    for (int sl = 0; sl < slabs; ++sl) {
      const float* pp = part + (((long long)b * slabs + sl) * G + threadIdx.x) * 2;
      a += pp[0];
      a2 += pp[1];
    }
    const float mu = a * inv_n;
    stat[threadIdx.x] = make_float2(mu, rsqrtf(fmaxf(a2 * inv_n - mu * mu, 0.f) + eps));
  }
  __syncthreads();
  const float2 m = stat[(4 * q) / cpg];
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + q), be = __ldg(reinterpret_cast<const float4*>(beta) + q);
  const size_t img0 = (size_t)b * P * cq;
  const float4* x4 = reinterpret_cast<const float4*>(x) + img0 + q;
  auto emit = [&](int r, const float4 v) {
    float o0 = fmaf((v.x - m.x) * m.y, ga.x, be.x), o1 = fmaf((v.y - m.x) * m.y, ga.y, be.y);      // same expression as the grid-stride pass
    float o2 = fmaf((v.z - m.x) * m.y, ga.z, be.z), o3 = fmaf((v.w - m.x) * m.y, ga.w, be.w);
    if (ACT == 1) {
      o0 = __fdividef(o0, 1.f + __expf(-o0)); o1 = __fdividef(o1, 1.f + __expf(-o1));
      o2 = __fdividef(o2, 1.f + __expf(-o2)); o3 = __fdividef(o3, 1.f + __expf(-o3));
    }
    const size_t i = img0 + (size_t)r * cq + q;
    if (y_f32) reinterpret_cast<float4*>(y_f32)[i] = make_float4(o0, o1, o2, o3);
    if (y_bf16) {
      uint2 u;
      u.x = pack_bf16x2(o0, o1);
      u.y = pack_bf16x2(o2, o3);
      reinterpret_cast<uint2*>(y_bf16)[i] = u;
    }
  };
  int r = r_begin + rl;
  for (; r + 3 * rpb < r_end; r += 4 * rpb) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldcs(x4 + (size_t)(r + k * rpb) * cq);
#pragma unroll
    for (int k = 0; k < 4; ++k) emit(r + k * rpb, v[k]);
  }
  for (; r < r_end; r += rpb) emit(r, __ldcs(x4 + (size_t)r * cq));
}

// One-head self-attention core over L = h*w tokens (FlaxAttentionBlock): scores = (q C^-1/4)(k C^-1/4)^T, softmax, @ v.
// qkv: (B, L, 3C) f32 [q | k | v];  out (B, L, C) f32 and/or bf16.  One block per image; L <= 64.
// Scores: the block walks C in chunks of 32 channels staged TRANSPOSED in shared memory ([channel][token], so that a thread's four
// tokens are one 16-byte load), every thread owns a 4x4 patch of the L x L score matrix: 2 LDS.128 per 16 FMA (the first version read
// eight scalars and was shared-memory-issue-bound).  Softmax: four threads per row.  P V: the probabilities are kept transposed
// ([key][query]); a thread owns one channel pair position and 16 queries, i.e. 4 LDS.128 + 1 coalesced global load per 16 FMA.
__global__ void __launch_bounds__(256) vae_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out_f32,
                                                       __nv_bfloat16* __restrict__ out_bf16, int L, int C) {
  __shared__ __align__(16) float sc[64][68];         // probabilities, transposed: sc[key j][query i]
  __shared__ __align__(16) float qs[32][68], ks[32][68];   // [channel of the chunk][token]
  const long long b = blockIdx.x;
  const float* base = qkv + b * L * 3 * C;
  const float scale = rsqrtf(sqrtf((float)C));
  const float s2 = scale * scale;
  const int ti = (threadIdx.x >> 4) * 4, tj = (threadIdx.x & 15) * 4;     // 16 x 16 threads x (4 x 4) = 64 x 64
  float acc[4][4] = {};
  // a thread stages tokens (tid >> 5) + 8 r, channel (tid & 31) of every chunk; the next chunk's 16 values are fetched into registers
  // before the current one is consumed (one block per image and 8 warps: nothing else hides the L2 latency)
  const int si = threadIdx.x >> 5, scn = threadIdx.x & 31;
  float qn[8], kn[8];
  auto fetch = [&](int c0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = si + 8 * r;
      const bool ok = c0 + scn < C && i < L;
      qn[r] = ok ? __ldg(base + (long long)i * 3 * C + c0 + scn) : 0.f;
      kn[r] = ok ? __ldg(base + (long long)i * 3 * C + C + c0 + scn) : 0.f;
    }
  };
  fetch(0);
  for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      qs[scn][si + 8 * r] = qn[r];
      ks[scn][si + 8 * r] = kn[r];
    }
    __syncthreads();
    if (c0 + 32 < C) fetch(c0 + 32);
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const float4 q4 = *reinterpret_cast<const float4*>(&qs[c][ti]);
      const float4 k4 = *reinterpret_cast<const float4*>(&ks[c][tj]);
      const float qa[4] = {q4.x, q4.y, q4.z, q4.w}, kb[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(qa[u], kb[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) sc[tj + v][ti + u] = (ti + u < L && tj + v < L) ? acc[u][v] * s2 : -INFINITY;   // padded keys: weight 0
  __syncthreads();
  {
    // softmax over the keys of query i: four threads per query (adjacent lanes), 16 keys each
    const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
    float mx = -INFINITY;
    for (int j = part * 16; j < part * 16 + 16; ++j) mx = fmaxf(mx, sc[j][i]);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
    for (int j = part * 16; j < part * 16 + 16; ++j) {
      const float e = i < L ? expf(sc[j][i] - mx) : 0.f;       // exp(-inf - mx) = 0 for padded keys
      sc[j][i] = e;
      sum += e;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = i < L ? 1.f / sum : 0.f;
    for (int j = part * 16; j < part * 16 + 16; ++j) sc[j][i] *= inv;
  }
  __syncthreads();
  // out[i][c] = sum_j p[i][j] v[j][c]: thread = (channel lane, 16-query block); channels strided by 64 so that a warp reads 32 consecutive floats of v
  const int cl = threadIdx.x & 63, qb = (threadIdx.x >> 6) * 16;
  for (int c = cl; c < C; c += 64) {
    float o[16] = {};
    for (int j0 = 0; j0 < L; j0 += 8) {
      float vv[8];                                     // eight independent loads in flight (one per iteration was a serial chain of L2 round trips)
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) vv[jj] = j0 + jj < L ? __ldg(base + (long long)(j0 + jj) * 3 * C + 2 * C + c) : 0.f;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = min(j0 + jj, 63);                // (rows beyond L hold zeros for valid queries)
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const float4 p4 = *reinterpret_cast<const float4*>(&sc[j][qb + 4 * u4]);
          o[4 * u4] = fmaf(p4.x, vv[jj], o[4 * u4]); o[4 * u4 + 1] = fmaf(p4.y, vv[jj], o[4 * u4 + 1]);
          o[4 * u4 + 2] = fmaf(p4.z, vv[jj], o[4 * u4 + 2]); o[4 * u4 + 3] = fmaf(p4.w, vv[jj], o[4 * u4 + 3]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (qb + u < L) {
        if (out_f32) out_f32[(b * L + qb + u) * C + c] = o[u];
        if (out_bf16) out_bf16[(b * L + qb + u) * C + c] = __float2bfloat16(o[u]);
      }
    }
  }
}

// quant_conv (1x1, 2L -> 2L), keep the first L channels (the mean), optional latent normalisation.
__global__ void vae_post_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                float* __restrict__ out, long long npix, int c2, int lat, float lat_min, float lat_max) {
  const long long total = npix * lat;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % lat);
    const long long p = i / lat;
    float acc = bias[c];
    for (int k = 0; k < c2; ++k) acc = fmaf(x[p * c2 + k], w[k * c2 + c], acc);
    if (lat_max > lat_min) acc = (acc - lat_min) / (lat_max - lat_min) * 2.f - 1.f;
    out[i] = acc;
  }
}

#define VAE_LAUNCH_OK(name)                                                                     \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess) {                                                                    \
      set_last_error(std::string(name " launch failed: ") + cudaGetErrorString(_e));           \
      return LDP_ERR_CUDA;                                                                      \
    }                                                                                           \
    count_launch();                                                                             \
  } while (0)

struct ConvW { const float* w = nullptr; const float* b = nullptr; int k = 3, cin = 0, cout = 0; };
struct ResW { const float *n1s, *n1b, *n2s, *n2b; ConvW c1, c2, sc; bool has_sc = false; };

// a packed conv on the tcgen05 path
struct VaeTcConv {
  PackedW pw;
  bool ready = false;
};

struct VaeWs {
  int Bc = 0;
  uint64_t last_use = 0;                             // LRU stamp (ws_evict_lru)
  Arena arena;
  float *S = nullptr, *Hf = nullptr, *stats = nullptr, *part = nullptr, *qkv = nullptr;
  float* Gf = nullptr;                               // fp32 path: GN output
  __nv_bfloat16 *Sb = nullptr, *Gb = nullptr;        // tc path
  bool tc_ready = false;
  std::vector<TcGemm> ops;
  TcGemm op_in[2];                                   // conv_in as a dense tcgen05 GEMM over the im2col matrix (uint8 / float pixels)
  bool op_in_ok = false;
  __nv_bfloat16* cols = nullptr;                     // im2col matrix [Bc * S * S][64]: Gb when that is large enough, else its own buffer
};

}  // namespace ldp

using namespace ldp;

struct LdpVae {
  LdpVaeConfig cfg;
  Arena arena;
  float* blob = nullptr;
  ConvW conv_in, conv_out, quant;
  std::vector<std::vector<ResW>> blocks;
  std::vector<ConvW> down;
  ResW mid0, mid1;
  const float *ag_s, *ag_b, *nos, *nob;
  ConvW aq, ak, av, ap;
  float* wqkv = nullptr; float* bqkv = nullptr;      // [C][3C], [3C] concatenated projections
  std::map<int, std::unique_ptr<VaeWs>> ws;
  uint64_t use_clock = 0;
  std::map<int, PackedW> packed;                     // conv id -> packed weights (tc path)
  PackedW pw_in[2];                                  // conv_in as a K = 64 dense GEMM: [0] uint8 pixels (normalisation folded in), [1] float pixels
  bool pw_in_ready[2] = {false, false};
  int64_t n_params = 0;
  bool is_decoder = false;                           // decoder handle: blocks = up blocks, down = upsampler convs, quant = post_quant_conv
};

namespace ldp {

static int64_t vae_param_count(const LdpVaeConfig& c) {
  auto conv = [](int64_t k, int64_t ci, int64_t co) { return k * k * ci * co + co; };
  auto res = [&](int64_t ci, int64_t co) { return 2 * ci + conv(3, ci, co) + 2 * co + conv(3, co, co) + (ci != co ? conv(1, ci, co) : 0); };
  int64_t n = conv(3, c.in_channels, c.block_out_channels[0]);
  int64_t ch = c.block_out_channels[0];
  for (int i = 0; i < c.n_blocks; ++i) {
    for (int j = 0; j < c.layers_per_block; ++j) {
      n += res(ch, c.block_out_channels[i]);
      ch = c.block_out_channels[i];
    }
    if (i != c.n_blocks - 1) n += conv(3, ch, ch);
  }
  n += res(ch, ch) + 2 * ch + 4 * (ch * ch + ch) + res(ch, ch);
  n += 2 * ch + conv(3, ch, 2 * c.latent_channels) + conv(1, 2 * c.latent_channels, 2 * c.latent_channels);
  return n;
}

static int vae_validate(const LdpVaeConfig* c) {
  LDP_CHECK(c != nullptr, LDP_ERR_INVALID_ARG, "null config");
  LDP_CHECK(c->in_channels >= 1 && c->in_channels <= 4 && c->latent_channels >= 1 && c->latent_channels <= 16, LDP_ERR_INVALID_ARG,
            "in_channels must be 1..4, latent_channels 1..16");
  LDP_CHECK(c->n_blocks >= 1 && c->n_blocks <= 8 && c->layers_per_block >= 1, LDP_ERR_INVALID_ARG, "bad block structure");
  LDP_CHECK(c->norm_num_groups >= 1 && c->norm_num_groups <= 64, LDP_ERR_INVALID_ARG, "norm_num_groups must be 1..64");
  const int S = c->image_size;
  LDP_CHECK(S >= 1 && S <= 128 && (S & (S - 1)) == 0 && (S >> (c->n_blocks - 1)) >= 1, LDP_ERR_UNSUPPORTED,
            "image_size must be a power of two <= 128 and survive n_blocks-1 halvings");
  const int hw = (S >> (c->n_blocks - 1));
  LDP_CHECK(hw * hw <= 64, LDP_ERR_UNSUPPORTED, "the attention core handles at most 64 tokens (8x8 latent grid)");
  for (int i = 0; i < c->n_blocks; ++i)
    LDP_CHECK(c->block_out_channels[i] > 0 && c->block_out_channels[i] % c->norm_num_groups == 0 && c->block_out_channels[i] % 8 == 0,
              LDP_ERR_INVALID_ARG, "block_out_channels must be multiples of norm_num_groups and 8");
  return LDP_OK;
}

static ConvW take_conv(BlobWalker& w, int k, int ci, int co) {
  ConvW c;
  c.k = k; c.cin = ci; c.cout = co;
  c.w = w.take((uint64_t)k * k * ci * co);
  c.b = w.take(co);
  return c;
}
static ResW take_res(BlobWalker& w, int ci, int co) {
  ResW r;
  r.n1s = w.take(ci); r.n1b = w.take(ci);
  r.c1 = take_conv(w, 3, ci, co);
  r.n2s = w.take(co); r.n2b = w.take(co);
  r.c2 = take_conv(w, 3, co, co);
  r.has_sc = ci != co;
  if (r.has_sc) r.sc = take_conv(w, 1, ci, co);
  return r;
}

static int vae_create_impl(const LdpVaeConfig* cfg, const float* params_host, uint64_t n_params, LdpVae* h) {
  h->cfg = *cfg;
  const LdpVaeConfig& c = h->cfg;
  const int64_t expect = vae_param_count(c);
  LDP_CHECK((int64_t)n_params == expect, LDP_ERR_PARAM_COUNT,
            "VAE weight blob has " + std::to_string(n_params) + " floats, config needs " + std::to_string(expect));
  LDP_TRY(h->arena.alloc_t(&h->blob, n_params, false));
  LDP_CUDA_OK(cudaMemcpy(h->blob, params_host, n_params * 4, cudaMemcpyHostToDevice));
  BlobWalker w{h->blob, 0};
  h->conv_in = take_conv(w, 3, c.in_channels, c.block_out_channels[0]);
  int ch = c.block_out_channels[0];
  for (int i = 0; i < c.n_blocks; ++i) {
    std::vector<ResW> rs;
    for (int j = 0; j < c.layers_per_block; ++j) {
      rs.push_back(take_res(w, ch, c.block_out_channels[i]));
      ch = c.block_out_channels[i];
    }
    h->blocks.push_back(rs);
    if (i != c.n_blocks - 1) h->down.push_back(take_conv(w, 3, ch, ch));
  }
  h->mid0 = take_res(w, ch, ch);
  h->ag_s = w.take(ch); h->ag_b = w.take(ch);
  h->aq.w = w.take((uint64_t)ch * ch); h->aq.b = w.take(ch);
  h->ak.w = w.take((uint64_t)ch * ch); h->ak.b = w.take(ch);
  h->av.w = w.take((uint64_t)ch * ch); h->av.b = w.take(ch);
  h->ap = ConvW();
  h->ap.k = 1; h->ap.cin = ch; h->ap.cout = ch;
  h->ap.w = w.take((uint64_t)ch * ch); h->ap.b = w.take(ch);
  h->mid1 = take_res(w, ch, ch);
  h->nos = w.take(ch); h->nob = w.take(ch);
  h->conv_out = take_conv(w, 3, ch, 2 * c.latent_channels);
  h->quant = take_conv(w, 1, 2 * c.latent_channels, 2 * c.latent_channels);
  LDP_CHECK((int64_t)w.pos == expect, LDP_ERR_PARAM_COUNT, "internal: blob walk mismatch");
  // concatenated q|k|v projection: one GEMM with N = 3C
  LDP_TRY(h->arena.alloc_t(&h->wqkv, (size_t)ch * 3 * ch));
  LDP_TRY(h->arena.alloc_t(&h->bqkv, (size_t)3 * ch));
  const float* ws[3] = {h->aq.w, h->ak.w, h->av.w};
  const float* bs[3] = {h->aq.b, h->ak.b, h->av.b};
  for (int i = 0; i < 3; ++i) {
    LDP_CUDA_OK(cudaMemcpy2D(h->wqkv + (size_t)i * ch, (size_t)3 * ch * 4, ws[i], (size_t)ch * 4, (size_t)ch * 4, ch,
                             cudaMemcpyDeviceToDevice));
    LDP_CUDA_OK(cudaMemcpy(h->bqkv + (size_t)i * ch, bs[i], (size_t)ch * 4, cudaMemcpyDeviceToDevice));
  }
  return LDP_OK;
}

// largest (pixels x channels) any activation tensor reaches per image
static size_t vae_max_act(const LdpVae* h) {
  const LdpVaeConfig& c = h->cfg;
  size_t max_act = 0;
  if (!h->is_decoder) {
    int S = c.image_size;
    for (int i = 0; i < c.n_blocks; ++i) {
      const int cin = i == 0 ? c.block_out_channels[0] : c.block_out_channels[i - 1];
      max_act = std::max(max_act, (size_t)S * S * std::max(cin, c.block_out_channels[i]));
      if (i != c.n_blocks - 1) S /= 2;
    }
    const int cl = c.block_out_channels[c.n_blocks - 1];
    max_act = std::max(max_act, (size_t)S * S * 3 * cl);
  } else {
    int S = c.image_size >> (c.n_blocks - 1);
    int ch = c.block_out_channels[c.n_blocks - 1];
    max_act = (size_t)S * S * 3 * ch;
    for (int i = 0; i < c.n_blocks; ++i) {
      const int co = c.block_out_channels[c.n_blocks - 1 - i];
      max_act = std::max(max_act, (size_t)S * S * std::max(ch, co));
      ch = co;
      if (i != c.n_blocks - 1) {
        S *= 2;
        max_act = std::max(max_act, (size_t)S * S * co);
      }
    }
  }
  return max_act;
}

static int vae_get_ws(LdpVae* h, int Bc, VaeWs** out) {
  auto it = h->ws.find(Bc);
  if (it != h->ws.end()) {
    it->second->last_use = ++h->use_clock;
    *out = it->second.get();
    return LDP_OK;
  }
  ws_evict_lru(h->ws);
  const LdpVaeConfig& c = h->cfg;
  std::unique_ptr<VaeWs> w(new VaeWs());
  w->Bc = Bc; w->last_use = ++h->use_clock;
  const size_t max_act = vae_max_act(h);
  const int S = c.image_size >> (c.n_blocks - 1);
  const int cl = c.block_out_channels[c.n_blocks - 1];
  LDP_TRY(w->arena.alloc_t(&w->S, (size_t)Bc * max_act, false));
  LDP_TRY(w->arena.alloc_t(&w->Hf, (size_t)Bc * max_act, false));
  LDP_TRY(w->arena.alloc_t(&w->stats, (size_t)Bc * c.norm_num_groups * 2));
  LDP_TRY(w->arena.alloc_t(&w->part, (size_t)Bc * 64 * c.norm_num_groups * 2));
  LDP_TRY(w->arena.alloc_t(&w->qkv, (size_t)Bc * S * S * 3 * cl, false));
  *out = w.get();
  h->ws[Bc] = std::move(w);
  return LDP_OK;
}

// ---- shared pieces ----
static int vae_gn(LdpVae* h, VaeWs* w, const float* x, int nimg, int P, int C, const float* gamma, const float* beta, int act,
                  float* y_f32, __nv_bfloat16* y_bf16, cudaStream_t s, int fused_slabs = 0) {
  const int G = h->cfg.norm_num_groups;
  LDP_CHECK(256 % (C / 4) == 0, LDP_ERR_UNSUPPORTED, "VAE GroupNorm needs C/4 to divide 256 (C in {8,...,1024} powers of two)");
  const long long per_img = (long long)P * C / 4;
  // fused_slabs > 0: the convolution that produced x already wrote part[image][tile][group] (tc_epilogue.cuh)
  const int slabs = fused_slabs > 0 ? fused_slabs : (int)std::min<long long>(std::max<long long>(1, per_img / 8192), 64);
  if (fused_slabs == 0) {
    vae_gn_stats_kernel<<<dim3(nimg, slabs), 256, 0, s>>>(x, w->part, P, C, G);
    VAE_LAUNCH_OK("vae_gn_stats");
  }
  static const bool img_major = !(getenv("LDP_VAE_GN_IMG") && getenv("LDP_VAE_GN_IMG")[0] == '0');
  const int cq = C / 4, cpg = C / G;
  const float inv_n = 1.f / ((float)P * (C / G));
  const bool fused_final = img_major && cpg % 4 == 0 && nimg <= 65535 && (act == 0 || act == 1) && G <= 64;
  if (!fused_final) {
    vae_gn_final_kernel<<<(nimg * G + 127) / 128, 128, 0, s>>>(w->part, w->stats, nimg * G, G, slabs, inv_n, 1e-6f);
    VAE_LAUNCH_OK("vae_gn_final");
  }
  if (fused_final) {
    // ~16 float4 per thread, but enough blocks to fill the machine a few times over
    const int rpb = 256 / cq;
    int rows_per_block = rpb * 16;
    while (rows_per_block > rpb * 4 && (long long)nimg * ceil_div(P, rows_per_block) < 148 * 8) rows_per_block >>= 1;
    const dim3 grid(ceil_div(P, rows_per_block), nimg);
    // the finalise step of the statistics (sum over slabs -> mean, rstd) runs inside the apply pass: 22 fewer launches per encoder pass
    if (act == 1) vae_gn_apply_img_kernel<1><<<grid, 256, 0, s>>>(x, w->part, gamma, beta, y_f32, y_bf16, P, C, G, rows_per_block, slabs, inv_n, 1e-6f);
    else vae_gn_apply_img_kernel<0><<<grid, 256, 0, s>>>(x, w->part, gamma, beta, y_f32, y_bf16, P, C, G, rows_per_block, slabs, inv_n, 1e-6f);
    VAE_LAUNCH_OK("vae_gn_apply_img");
    return LDP_OK;
  }
  const long long total = (long long)nimg * per_img;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  vae_gn_apply_kernel<<<blocks, 256, 0, s>>>(x, w->stats, gamma, beta, y_f32, y_bf16, nimg, P, C, G, act);
  VAE_LAUNCH_OK("vae_gn_apply");
  return LDP_OK;
}

static int vae_conv_f32(const ConvW& cw, const float* in, const float* res, float* out, int nimg, int Hin, int stride, int pad,
                        cudaStream_t s) {
  const int Hout = stride == 2 ? Hin / 2 : Hin;
  const int xg = (Hout + 7) / 8;
  dim3 grid((unsigned)((long long)nimg * Hout * xg), (cw.cout + 127) / 128);
  vae_conv_f32_kernel<<<grid, 128, 0, s>>>(in, cw.w, cw.b, res, out, nimg, Hin, Hin, Hout, Hout, cw.cin, cw.cout, cw.k, stride, pad);
  VAE_LAUNCH_OK("vae_conv_f32");
  return LDP_OK;
}

// ---- fp32 program ----
static int vae_res_f32(LdpVae* h, VaeWs* w, const ResW& r, int nimg, int S, cudaStream_t s) {
  const int P = S * S;
  LDP_TRY(vae_gn(h, w, w->S, nimg, P, r.c1.cin, r.n1s, r.n1b, 1, w->Gf, nullptr, s));
  LDP_TRY(vae_conv_f32(r.c1, w->Gf, nullptr, w->Hf, nimg, S, 1, 1, s));
  LDP_TRY(vae_gn(h, w, w->Hf, nimg, P, r.c1.cout, r.n2s, r.n2b, 1, w->Gf, nullptr, s));
  if (r.has_sc) {
    // shortcut(x) into Hf (free now), then S = conv2(Gf) + Hf
    LDP_TRY(vae_conv_f32(r.sc, w->S, nullptr, w->Hf, nimg, S, 1, 0, s));
    LDP_TRY(vae_conv_f32(r.c2, w->Gf, w->Hf, w->S, nimg, S, 1, 1, s));
  } else {
    LDP_TRY(vae_conv_f32(r.c2, w->Gf, w->S, w->S, nimg, S, 1, 1, s));     // in-place residual: each thread reads res[o] then writes out[o]
  }
  return LDP_OK;
}

static int vae_forward_f32(LdpVae* h, VaeWs* w, const void* images, int fmt, int nimg, float lat_min, float lat_max, float* out,
                           cudaStream_t s) {
  const LdpVaeConfig& c = h->cfg;
  if (!w->Gf) LDP_TRY(w->arena.alloc_t(&w->Gf, (size_t)w->Bc * vae_max_act(h), false));
  int S = c.image_size;
  const int c0 = c.block_out_channels[0];
  const size_t smem = (size_t)9 * c.in_channels * c0 * 4;
  const long long tot = (long long)nimg * S * S * c0 / 4;
  const int blocks = (int)std::min<long long>((tot + 255) / 256, 148 * 8);
  if (fmt == 0)
    vae_conv_in_kernel<uint8_t><<<blocks, 256, smem, s>>>((const uint8_t*)images, h->conv_in.w, h->conv_in.b, w->S, nullptr, nimg, S,
                                                           c.in_channels, c0);
  else
    vae_conv_in_kernel<float><<<blocks, 256, smem, s>>>((const float*)images, h->conv_in.w, h->conv_in.b, w->S, nullptr, nimg, S,
                                                         c.in_channels, c0);
  VAE_LAUNCH_OK("vae_conv_in");
  for (int i = 0; i < c.n_blocks; ++i) {
    for (auto& r : h->blocks[i]) LDP_TRY(vae_res_f32(h, w, r, nimg, S, s));
    if (i != c.n_blocks - 1) {
      // FlaxDownsample2D: pad H,W by (0,1), conv 3x3 stride 2 VALID  ==  pad 0 on the low side, zero beyond the high edge
      LDP_TRY(vae_conv_f32(h->down[i], w->S, nullptr, w->Hf, nimg, S, 2, 0, s));
      std::swap(w->S, w->Hf);
      S /= 2;
    }
  }
  const int ch = c.block_out_channels[c.n_blocks - 1], L = S * S;
  LDP_TRY(vae_res_f32(h, w, h->mid0, nimg, S, s));
  {
    LDP_TRY(vae_gn(h, w, w->S, nimg, L, ch, h->ag_s, h->ag_b, 0, w->Gf, nullptr, s));
    GemmF32 g;
    g.x1 = w->Gf; g.c1 = ch; g.ld1 = ch; g.w = h->wqkv; g.ldw = 3 * ch; g.bias = h->bqkv; g.out = w->qkv; g.ldo = 3 * ch;
    g.m = nimg * L; g.n = 3 * ch;
    LDP_TRY(launch_gemm_f32(g, s));
    vae_attn_kernel<<<nimg, 256, 0, s>>>(w->qkv, w->Gf, nullptr, L, ch);
    VAE_LAUNCH_OK("vae_attn");
    g = GemmF32();
    g.x1 = w->Gf; g.c1 = ch; g.ld1 = ch; g.w = h->ap.w; g.ldw = ch; g.bias = h->ap.b; g.res = w->S; g.ldres = ch;
    g.out = w->S; g.ldo = ch; g.m = nimg * L; g.n = ch;
    LDP_TRY(launch_gemm_f32(g, s));
  }
  LDP_TRY(vae_res_f32(h, w, h->mid1, nimg, S, s));
  LDP_TRY(vae_gn(h, w, w->S, nimg, L, ch, h->nos, h->nob, 1, w->Gf, nullptr, s));
  LDP_TRY(vae_conv_f32(h->conv_out, w->Gf, nullptr, w->Hf, nimg, S, 1, 1, s));
  const long long npix = (long long)nimg * L;
  vae_post_kernel<<<(int)std::min<long long>((npix * c.latent_channels + 255) / 256, 4096), 256, 0, s>>>(
      w->Hf, h->quant.w, h->quant.b, out, npix, 2 * c.latent_channels, c.latent_channels, lat_min, lat_max);
  VAE_LAUNCH_OK("vae_post");
  return LDP_OK;
}

// ---- tcgen05 program ----
// Tile geometry of a level: 128 output pixels = wb x hb pixels of ib images (whole rows, so that a tile's rows are
// consecutive NHWC pixels).
struct TileGeo { int wb, hb, ib, tiles_per_img; };
static TileGeo tile_geo(int S) {
  TileGeo g;
  g.wb = S;
  g.hb = std::min(S, 128 / S);
  g.ib = 128 / (g.wb * g.hb);
  g.tiles_per_img = g.ib > 1 ? 1 : S / g.hb;
  return g;
}

// share_rows: the stage order of TcGemm::a_rows == 256 - a stage is (kw, channel block) with the three kh taps as its W tiles
static int vae_pack(LdpVae* h, int id, const float* wgt, int k, int pad, int cin, int cout, PackedW** out, bool share_rows = false) {
  if (share_rows) id += 1 << 20;                     // its own cache entry: the K order of the packed weights differs
  auto it = h->packed.find(id);
  if (it != h->packed.end()) {
    *out = &it->second;
    return LDP_OK;
  }
  PackedW pw;
  std::vector<TcStage> st;
  std::vector<int32_t> kmap;
  if (share_rows) {
    for (int dx = 0; dx < k; ++dx)
      for (int c0 = 0; c0 < cin; c0 += 64) {
        st.push_back(make_stage(0, 0, k, c0, dx - pad, -pad, (int)kmap.size() / 64));      // A box starts one image row above the tile
        for (int dy = 0; dy < k; ++dy)
          for (int i = 0; i < 64; ++i) kmap.push_back(c0 + i < cin ? (dy * k + dx) * cin + c0 + i : -1);
      }
  } else
  for (int dy = 0; dy < k; ++dy)
    for (int dx = 0; dx < k; ++dx)
      for (int c0 = 0; c0 < cin; c0 += 64) {
        st.push_back(make_stage(0, 0, 1, c0, dx - pad, dy - pad, (int)kmap.size() / 64));
        for (int i = 0; i < 64; ++i) kmap.push_back(c0 + i < cin ? (dy * k + dx) * cin + c0 + i : -1);
      }
  pw.kp = (int)kmap.size();
  pw.n_pad = round_up(cout, 128);
  LDP_TRY(h->arena.alloc_t(&pw.wt, (size_t)pw.n_pad * pw.kp));
  LDP_TRY(upload_stage_table(h->arena, st, &pw));
  Arena tmp;
  int32_t* map_dev;
  LDP_TRY(tmp.alloc_t(&map_dev, kmap.size()));
  LDP_CUDA_OK(cudaMemcpy(map_dev, kmap.data(), kmap.size() * 4, cudaMemcpyHostToDevice));
  LDP_TRY(launch_pack_wt_bf16(wgt, cout, cout, map_dev, pw.kp, pw.wt, pw.kp, 0, pw.n_pad, 0));
  LDP_CUDA_OK(cudaDeviceSynchronize());
  h->packed[id] = pw;
  *out = &h->packed[id];
  return LDP_OK;
}

// Outputs and the f32 residual of a PLAIN-epilogue op as TMA boxes (TcGemm::epi_tma): maps live in device memory.
// LDP_VAE_EPI_TMA=0 keeps the row-per-thread vector accesses.
static int vae_epi_maps(Arena& arena, TcGemm* op, size_t rows) {
  static const bool on = !(getenv("LDP_VAE_EPI_TMA") && getenv("LDP_VAE_EPI_TMA")[0] == '0');
  if (!on) return LDP_OK;
  CUtensorMap host[3];
  int bits = 0;
  // 256-wide tiles: 2 KB staging buffers (16-column f32 boxes) buy the CTA pairs a fifth ring stage (LDP_VAE_EPI_HALF=0: 4 KB everywhere)
  static const bool half_ok = !(getenv("LDP_VAE_EPI_HALF") && getenv("LDP_VAE_EPI_HALF")[0] == '0');
  LDP_TRY(tc_build_epi_maps(*op, rows, host, &bits, (half_ok && op->block_n == 256) || op->a_rows == 256));
  if (!bits) return LDP_OK;
  CUtensorMap* dev;
  LDP_TRY(arena.alloc_t(&dev, 3));
  LDP_CUDA_OK(cudaMemcpy(dev, host, sizeof(host), cudaMemcpyHostToDevice));
  op->epi_maps = dev;
  op->epi_tma = bits;
  return LDP_OK;
}

// conv on `in` (bf16, NHWC at resolution S_in) -> PLAIN epilogue.  stride 2 = the Downsample conv: taps address
// (2x + dx, 2y + dy) with dx, dy >= 0 (pad low 0), the row/column beyond the high edge is TMA out-of-bounds zero fill.
static int vae_conv_tc(LdpVae* h, VaeWs* w, int id, const ConvW& cw, const float* bias, const __nv_bfloat16* in, int S_in,
                       int stride, TcGemm* op) {
  PackedW* pw;
  const int pad = (cw.k == 3 && stride == 1) ? 1 : 0;
  const int S_out = S_in / stride;
  const TileGeo g = tile_geo(S_out);
  // 256-wide tiles where the channel count allows: one A tile feeds twice the MMA work (per-tap stages of 48 KB with
  // 512 cycles of MMAs instead of 32 KB with 256), which is what the latency-bound per-tap ring needs
  static const bool bn256 = !(getenv("LDP_VAE_BN256") && getenv("LDP_VAE_BN256")[0] == '0');
  const int bn = (bn256 && cw.cout % 256 == 0) ? 256 : (cw.cout > 64 ? 128 : 64);
  static const bool pair_ok = !(getenv("LDP_VAE_PAIR") && getenv("LDP_VAE_PAIR")[0] == '0');
  const int tiles_m = ceil_div(w->Bc * S_out * S_out, 128), tiles_n = ceil_div(cw.cout, bn);
  const bool pair = pair_ok && bn >= 128 && tiles_m % 2 == 0 && tiles_m * tiles_n > 148;
  // Shared tap rows (TcGemm::a_rows = 256): 3x3 stride-1 convolutions whose tile is exactly two whole rows of one image, as CTA pairs
  // at BN = 128 (a stage is 32 KB of A + 3 x 8 KB of W; three of them fit next to 2 KB staging buffers).  LDP_VAE_SHARE_ROWS=0: per-tap stages.
  static const bool share_ok = !(getenv("LDP_VAE_SHARE_ROWS") && getenv("LDP_VAE_SHARE_ROWS")[0] == '0');
  const bool share_rows = share_ok && pair && bn == 128 && cw.k == 3 && stride == 1 && g.ib == 1 && g.hb == 2 && g.wb == S_out &&
                          g.wb * g.hb == 128 && cw.cin % 64 == 0;
  LDP_TRY(vae_pack(h, id, cw.w, cw.k, pad, cw.cin, cw.cout, &pw, share_rows));
  *op = TcGemm();
  uint64_t dims[4] = {(uint64_t)cw.cin, (uint64_t)S_in, (uint64_t)S_in, (uint64_t)w->Bc};
  uint64_t str[3] = {(uint64_t)cw.cin * 2, (uint64_t)S_in * cw.cin * 2, (uint64_t)S_in * S_in * cw.cin * 2};
  uint32_t box[4] = {64, (uint32_t)(g.wb * stride), (uint32_t)((share_rows ? g.hb + 2 : g.hb) * stride), (uint32_t)g.ib};
  uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
  LDP_TRY(make_tmap_bf16_strided(&op->map_a[0], in, 4, dims, str, box, es));
  for (int i = 1; i < 4; ++i) op->map_a[i] = op->map_a[0];
  if (share_rows) {
    op->a_rows = 256;
    op->a_tap_shift16 = (g.wb * 128) >> 4;           // one image row of the box = wb rows of 128 bytes
    op->taps_same_acc = 1;
  }
  uint64_t bd[2] = {(uint64_t)pw->kp, (uint64_t)pw->n_pad};
  uint64_t bs[1] = {(uint64_t)pw->kp * 2};
  // CTA pairs (cta_group::2) where the launch is persistent: the convolutions are bound by L2 -> SM operand delivery, and in a
  // pair each CTA fetches only half of every W tile (LDP_VAE_PAIR=0: single CTAs)
  uint32_t bb[2] = {64, (uint32_t)(pair ? bn / 2 : bn)};
  LDP_TRY(make_tmap_bf16(&op->map_b, pw->wt, 2, bd, bs, bb));
  op->pair = pair ? 1 : 0;
  op->kb = pw->kb_dev; op->num_kb = pw->num_kb; op->runs = pw->runs_dev; op->num_runs = pw->num_runs; tc_set_inline_runs(op, pw->runs_host.data(), pw->num_runs); op->w_max = share_rows ? cw.k : 1; op->k_pad = pw->kp;
  op->M = w->Bc * S_out * S_out; op->N = cw.cout; op->block_n = bn;
  op->tiles_per_item = g.tiles_per_img; op->rows_step = g.hb * stride; op->items_per_tile = g.ib;
  op->rows_per_item = 1;
  op->mode = TC_EPI_PLAIN;
  op->bias = bias;
  return LDP_OK;
}

// conv_in as a dense K = 64 GEMM (see vae_im2col_in_kernel): packed weights [n_pad][64] bf16 for uint8 (fmt 0) or float (fmt 1) pixels
static int vae_pack_conv_in(LdpVae* h, int fmt) {
  if (h->pw_in_ready[fmt]) return LDP_OK;
  const int cin = h->cfg.in_channels, c0 = h->cfg.block_out_channels[0], kc = 9 * cin;
  std::vector<float> w((size_t)kc * c0), full((size_t)64 * c0, 0.f);
  LDP_CUDA_OK(cudaMemcpy(w.data(), h->conv_in.w, w.size() * 4, cudaMemcpyDeviceToHost));
  for (int k = 0; k < kc; ++k)
    for (int n = 0; n < c0; ++n) full[(size_t)k * c0 + n] = fmt == 0 ? w[(size_t)k * c0 + n] / 127.5f : w[(size_t)k * c0 + n];
  if (fmt == 0)
    for (int tap = 0; tap < 9; ++tap)
      for (int n = 0; n < c0; ++n) {
        float sum = 0.f;
        for (int ci = 0; ci < cin; ++ci) sum += w[(size_t)(tap * cin + ci) * c0 + n];
        full[(size_t)(kc + tap) * c0 + n] = sum * (0.5f / 127.5f);
      }
  PackedW& pw = h->pw_in[fmt];
  pw.kp = 64;
  pw.n_pad = round_up(c0, 128);
  LDP_TRY(h->arena.alloc_t(&pw.wt, (size_t)pw.n_pad * pw.kp));
  std::vector<TcStage> st(1, make_stage(0, 0, 1, 0, 0, 0, 0));
  LDP_TRY(upload_stage_table(h->arena, st, &pw));
  Arena tmp;
  float* full_dev;
  int32_t* map_dev;
  std::vector<int32_t> kmap(64);
  for (int k = 0; k < 64; ++k) kmap[k] = k;
  LDP_TRY(tmp.alloc_t(&full_dev, full.size()));
  LDP_TRY(tmp.alloc_t(&map_dev, 64));
  LDP_CUDA_OK(cudaMemcpy(full_dev, full.data(), full.size() * 4, cudaMemcpyHostToDevice));
  LDP_CUDA_OK(cudaMemcpy(map_dev, kmap.data(), 64 * 4, cudaMemcpyHostToDevice));
  LDP_TRY(launch_pack_wt_bf16(full_dev, c0, c0, map_dev, 64, pw.wt, 64, 0, pw.n_pad, 0));
  LDP_CUDA_OK(cudaDeviceSynchronize());
  h->pw_in_ready[fmt] = true;
  return LDP_OK;
}

static int vae_conv_in_op(LdpVae* h, VaeWs* w, int fmt, const __nv_bfloat16* cols, float* of32, __nv_bfloat16* obf, TcGemm* op) {
  LDP_TRY(vae_pack_conv_in(h, fmt));
  const PackedW& pw = h->pw_in[fmt];
  const LdpVaeConfig& c = h->cfg;
  const int S = c.image_size, c0 = c.block_out_channels[0], G = c.norm_num_groups, cpg = c0 / G;
  const TileGeo g = tile_geo(S);
  const int P = S * S;
  *op = TcGemm();
  if (g.ib == 1) {                  // a tile = 128 consecutive pixels of one image
    uint64_t dims[4] = {64, 128, (uint64_t)g.tiles_per_img, (uint64_t)w->Bc};
    uint64_t str[3] = {128, 128 * 128, (uint64_t)P * 128};
    uint32_t box[4] = {64, 128, 1, 1};
    LDP_TRY(make_tmap_bf16(&op->map_a[0], cols, 4, dims, str, box));
    op->tiles_per_item = g.tiles_per_img; op->rows_step = 1; op->items_per_tile = 1;
  } else {                          // a tile = all P pixels of ib images
    uint64_t dims[4] = {64, (uint64_t)P, 1, (uint64_t)w->Bc};
    uint64_t str[3] = {128, (uint64_t)P * 128, (uint64_t)P * 128};
    uint32_t box[4] = {64, (uint32_t)P, 1, (uint32_t)g.ib};
    LDP_TRY(make_tmap_bf16(&op->map_a[0], cols, 4, dims, str, box));
    op->tiles_per_item = 1; op->rows_step = 0; op->items_per_tile = g.ib;
  }
  for (int i = 1; i < 4; ++i) op->map_a[i] = op->map_a[0];
  const int bn = c0 > 64 ? 128 : 64;
  uint64_t bd[2] = {(uint64_t)pw.kp, (uint64_t)pw.n_pad};
  uint64_t bs[1] = {(uint64_t)pw.kp * 2};
  uint32_t bb[2] = {64, (uint32_t)bn};
  LDP_TRY(make_tmap_bf16(&op->map_b, pw.wt, 2, bd, bs, bb));
  op->kb = pw.kb_dev; op->num_kb = pw.num_kb; op->runs = pw.runs_dev; op->num_runs = pw.num_runs;
  tc_set_inline_runs(op, pw.runs_host.data(), pw.num_runs);
  op->w_max = 1; op->k_pad = pw.kp;
  op->M = w->Bc * P; op->N = c0; op->block_n = bn;
  op->rows_per_item = 1;
  op->mode = TC_EPI_PLAIN;
  op->bias = h->conv_in.b;
  op->out_f32 = of32; op->ld_out_f32 = c0;
  op->out_bf16 = obf; op->ld_out_bf16 = c0;
  const bool fuse = c0 % G == 0 && (cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) && P >= 64 && g.ib <= 2 && g.tiles_per_img <= 64;
  if (fuse) {                       // partial GroupNorm sums of the first resnet's norm1 come out of this epilogue too
    op->gn_part = w->part; op->gn_cpg = cpg; op->gn_G = G; op->gn_slabs = g.tiles_per_img; op->gn_imgs_per_tile = g.ib;
  }
  LDP_TRY(vae_epi_maps(w->arena, op, (size_t)w->Bc * P));
  return LDP_OK;
}

struct VaeBufs { float *S, *Hf; __nv_bfloat16 *Sb, *Gb; };

// Walks the encoder once.  build = true: creates the TcGemm ops (tensor maps bound to the workspace buffers) in
// execution order; build = false: runs the program for `nimg` images, consuming the ops in the same order.
static int vae_walk_tc(LdpVae* h, VaeWs* w, bool build, const void* images, int fmt, int nimg, float lat_min, float lat_max,
                       float* out, cudaStream_t s) {
  const LdpVaeConfig& c = h->cfg;
  VaeBufs b{w->S, w->Hf, w->Sb, w->Gb};
  // LDP_VAE_DBG=1: per-CTA wait-cycle sums of every persistent convolution of the third replay, printed to stderr
  static long long* dbg_store = nullptr;
  static int dbg_calls = 0;
  long long* dbg_buf = nullptr;
  if (!build && getenv("LDP_VAE_DBG") && ++dbg_calls == 3) {
    if (!dbg_store) LDP_CUDA_OK(cudaMalloc(&dbg_store, (size_t)64 * 148 * 8 * sizeof(long long)));
    LDP_CUDA_OK(cudaMemsetAsync(dbg_store, 0, (size_t)64 * 148 * 8 * sizeof(long long), s));
    dbg_buf = dbg_store;
  }
  size_t oi = 0;
  int conv_id = 0;
  // emit(op builder) / run(op)
  // gn_next: the f32 output is the input of a GroupNorm -> its statistics are accumulated in this epilogue
  static const bool fuse_gn = !(getenv("LDP_VAE_FUSE_GN") && getenv("LDP_VAE_FUSE_GN")[0] == '0');
  int fused_slabs = 0;            // > 0 after a conv that wrote the partial sums of its output
  auto conv = [&](const ConvW& cw, const __nv_bfloat16* in, int S_in, int stride, const float* res, float* of32,
                  __nv_bfloat16* obf, bool gn_next = false) -> int {
    const int S_out = S_in / stride;
    const TileGeo tg = tile_geo(S_out);
    const int G = c.norm_num_groups, cpg = cw.cout / G;
    const bool fuse = fuse_gn && gn_next && of32 != nullptr && cw.cout % G == 0 && (cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) &&
                      S_out * S_out >= 64 && tg.ib <= 2 && tg.tiles_per_img <= 64;
    fused_slabs = fuse ? tg.tiles_per_img : 0;
    if (build) {
      TcGemm op;
      LDP_TRY(vae_conv_tc(h, w, conv_id, cw, cw.b, in, S_in, stride, &op));
      op.res_f32 = res; op.ld_res_f32 = cw.cout;
      op.out_f32 = of32; op.ld_out_f32 = cw.cout;
      op.out_bf16 = obf; op.ld_out_bf16 = cw.cout;
      if (fuse) {
        op.gn_part = w->part; op.gn_cpg = cpg; op.gn_G = G; op.gn_slabs = tg.tiles_per_img; op.gn_imgs_per_tile = tg.ib;
      }
      LDP_TRY(vae_epi_maps(w->arena, &op, (size_t)w->Bc * S_out * S_out));
      w->ops.push_back(op);
    } else {
      TcGemm op = w->ops[oi];
      op.M = nimg * S_out * S_out;
      if (dbg_buf && oi < 64) op.dbg = dbg_buf + (size_t)oi * 148 * 8;
      LDP_TRY(launch_tc_gemm(op, s));
    }
    ++oi;
    ++conv_id;
    return LDP_OK;
  };
  auto gn = [&](const float* x, int P, int C, const float* gs_, const float* gb_, int act, __nv_bfloat16* y) -> int {
    const int slabs = fused_slabs;      // set by the conv that produced x (0: separate statistics kernel)
    fused_slabs = 0;
    if (build) return LDP_OK;
    return vae_gn(h, w, x, nimg, P, C, gs_, gb_, act, nullptr, y, s, slabs);
  };
  // sb_out: the bf16 copy of the block's output is read only by a shortcut convolution or the Downsample convolution; everywhere
  // else (the next consumer is a GroupNorm, which reads the f32 stream) it is not written
  auto resnet = [&](const ResW& r, int S, bool sb_out) -> int {
    const int P = S * S;
    __nv_bfloat16* sb = sb_out ? b.Sb : nullptr;
    LDP_TRY(gn(b.S, P, r.c1.cin, r.n1s, r.n1b, 1, b.Gb));
    LDP_TRY(conv(r.c1, b.Gb, S, 1, nullptr, b.Hf, nullptr, true));
    LDP_TRY(gn(b.Hf, P, r.c1.cout, r.n2s, r.n2b, 1, b.Gb));
    if (r.has_sc) {
      LDP_TRY(conv(r.sc, b.Sb, S, 1, nullptr, b.Hf, nullptr));            // Hf is free again after the second GroupNorm
      LDP_TRY(conv(r.c2, b.Gb, S, 1, b.Hf, b.S, sb, true));
    } else {
      LDP_TRY(conv(r.c2, b.Gb, S, 1, b.S, b.S, sb, true));                 // in place: a thread reads its residual row, then writes it
    }
    return LDP_OK;
  };
  int S = c.image_size;
  const int c0 = c.block_out_channels[0];
  static const bool in_tc = !(getenv("LDP_VAE_IN_TC") && getenv("LDP_VAE_IN_TC")[0] == '0');
  const bool conv_in_tc = in_tc && 9 * c.in_channels + 9 <= 64 && ((S * S) % 128 == 0 || 128 % (S * S) == 0);
  if (build) {
    w->op_in_ok = false;
    if (conv_in_tc) {
      // the im2col matrix lives in Gb (which the first GroupNorm overwrites only after this GEMM has read it) when Gb holds 64
      // values per pixel, i.e. for c0 >= 64; narrow test encoders get their own buffer
      if (vae_max_act(h) >= (size_t)S * S * 64) w->cols = b.Gb;
      else LDP_TRY(w->arena.alloc_t(&w->cols, (size_t)w->Bc * S * S * 64));
      // (the bf16 copy of conv_in's output would only feed a shortcut convolution of the first resnet)
      __nv_bfloat16* in_sb = (!h->blocks[0].empty() && h->blocks[0][0].has_sc) ? b.Sb : nullptr;
      for (int f = 0; f < 2; ++f) LDP_TRY(vae_conv_in_op(h, w, f, w->cols, b.S, in_sb, &w->op_in[f]));
      w->op_in_ok = true;
    }
  } else if (w->op_in_ok) {
    const long long npix = (long long)nimg * S * S;
    const int blocks = (int)std::min<long long>((npix + 255) / 256, 148 * 16);
    if (fmt == 0) vae_im2col_in_kernel<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)images, w->cols, npix, S, c.in_channels);
    else vae_im2col_in_kernel<float><<<blocks, 256, 0, s>>>((const float*)images, w->cols, npix, S, c.in_channels);
    VAE_LAUNCH_OK("vae_im2col_in");
    TcGemm op = w->op_in[fmt];
    op.M = (int)npix;
    LDP_TRY(launch_tc_gemm(op, s));
    fused_slabs = op.gn_part ? op.gn_slabs : 0;
  } else {
    const size_t smem = (size_t)9 * c.in_channels * c0 * 4;
    const long long tot = (long long)nimg * S * S * c0 / 4;
    const int blocks = (int)std::min<long long>((tot + 255) / 256, 148 * 8);
    if (fmt == 0)
      vae_conv_in_kernel<uint8_t><<<blocks, 256, smem, s>>>((const uint8_t*)images, h->conv_in.w, h->conv_in.b, b.S, b.Sb, nimg, S,
                                                             c.in_channels, c0);
    else
      vae_conv_in_kernel<float><<<blocks, 256, smem, s>>>((const float*)images, h->conv_in.w, h->conv_in.b, b.S, b.Sb, nimg, S,
                                                           c.in_channels, c0);
    VAE_LAUNCH_OK("vae_conv_in");
  }
  for (int i = 0; i < c.n_blocks; ++i) {
    const size_t nres = h->blocks[i].size();
    for (size_t j = 0; j < nres; ++j) {
      const bool last = j + 1 == nres;
      const bool sb_out = last ? (i != c.n_blocks - 1) : h->blocks[i][j + 1].has_sc;
      LDP_TRY(resnet(h->blocks[i][j], S, sb_out));
    }
    if (i != c.n_blocks - 1) {
      const bool next_sc = !h->blocks[i + 1].empty() && h->blocks[i + 1][0].has_sc;
      LDP_TRY(conv(h->down[i], b.Sb, S, 2, nullptr, b.Hf, next_sc ? b.Gb : nullptr, true));
      std::swap(b.S, b.Hf);
      std::swap(b.Sb, b.Gb);
      S /= 2;
    }
  }
  const int ch = c.block_out_channels[c.n_blocks - 1], L = S * S;
  LDP_TRY(resnet(h->mid0, S, false));
  {
    LDP_TRY(gn(b.S, L, ch, h->ag_s, h->ag_b, 0, b.Gb));
    ConvW qkv;
    qkv.k = 1; qkv.cin = ch; qkv.cout = 3 * ch; qkv.w = h->wqkv; qkv.b = h->bqkv;
    LDP_TRY(conv(qkv, b.Gb, S, 1, nullptr, w->qkv, nullptr));
    if (!build) {
      vae_attn_kernel<<<nimg, 256, 0, s>>>(w->qkv, nullptr, b.Gb, L, ch);
      VAE_LAUNCH_OK("vae_attn");
    }
    LDP_TRY(conv(h->ap, b.Gb, S, 1, b.S, b.S, nullptr, true));
  }
  LDP_TRY(resnet(h->mid1, S, false));
  LDP_TRY(gn(b.S, L, ch, h->nos, h->nob, 1, b.Gb));
  LDP_TRY(conv(h->conv_out, b.Gb, S, 1, nullptr, b.Hf, nullptr));
  if (!build) {
    const long long npix = (long long)nimg * L;
    vae_post_kernel<<<(int)std::min<long long>((npix * c.latent_channels + 255) / 256, 4096), 256, 0, s>>>(
        b.Hf, h->quant.w, h->quant.b, out, npix, 2 * c.latent_channels, c.latent_channels, lat_min, lat_max);
    VAE_LAUNCH_OK("vae_post");
  }
  if (dbg_buf) {
    LDP_CUDA_OK(cudaStreamSynchronize(s));
    std::vector<long long> d((size_t)64 * 148 * 8);
    LDP_CUDA_OK(cudaMemcpy(d.data(), dbg_buf, d.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    fprintf(stderr, "# op M N kb bn pair | per-CTA mean cycles: total producer_wait_empty issuer_wait_acc issuer_wait_operands epilogue_wait_acc\n");
    for (int o = 0; o < (int)std::min<size_t>((size_t)oi, 64); ++o) {
      const TcGemm& op = w->ops[o];
      double sum[8] = {0};
      int n = 0;
      for (int cta = 0; cta < 148; ++cta) {
        const long long* r = &d[((size_t)o * 148 + cta) * 8];
        if (r[7] == 0) continue;
        ++n;
        for (int i = 0; i < 8; ++i) sum[i] += (double)r[i];
      }
      if (n == 0) continue;
      fprintf(stderr, "%2d %8d %4d %3d %3d %d | %9.0f %9.0f %9.0f %9.0f %9.0f  (%d ctas)\n", o, nimg * 0 + op.M, op.N, op.num_kb, op.block_n, op.pair,
              sum[7] / n, sum[2] / n, sum[3] / n, sum[4] / n, sum[5] / n, n);
    }
  }
  return LDP_OK;
}

static int vae_prepare_tc(LdpVae* h, VaeWs* w) {
  if (w->tc_ready) return LDP_OK;
  LDP_TRY(tc_driver_check());
  LDP_TRY(tc_gemm_init());
  const size_t max_act = vae_max_act(h);
  LDP_TRY(w->arena.alloc_t(&w->Sb, (size_t)w->Bc * max_act));
  LDP_TRY(w->arena.alloc_t(&w->Gb, (size_t)w->Bc * max_act));
  w->ops.clear();
  LDP_TRY(vae_walk_tc(h, w, true, nullptr, 0, w->Bc, 0.f, 0.f, nullptr, 0));
  w->tc_ready = true;
  return LDP_OK;
}


// ------------------------------------------------------------------------------------------------
// Decoder: FlaxAutoencoderKL.decode(z).sample (reference agent/ldp_agent.py:66-85, the plan_viz of sample_viz :483).
// post_quant_conv 1x1 -> conv_in -> mid (resnet, attention, resnet) -> up blocks over the reversed channel list
// (layers_per_block + 1 resnets, nearest x2 + conv 3x3 on all but the last) -> GroupNorm -> swish -> conv_out.
// Same handle type and building blocks as the encoder: `blocks` holds the up blocks, `down` the upsampler convs, `quant`
// the post_quant_conv; the 4-channel conv_in and the 3-channel conv_out run on the SIMT convolution (K = 36 / N = 3 are
// no tensor-core shapes), everything between on the tcgen05 path.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) vae_upsample2x_kernel(const T* __restrict__ in, T* __restrict__ out, long long nimg, int S,
                                                             int C) {
  // out (n, 2S, 2S, C) = in (n, S, S, C) nearest: out[y][x] = in[y / 2][x / 2]; 16-byte vectors along C
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V, S2 = 2 * S;
  const long long total = nimg * S2 * S2 * cv;
  const uint4* i4 = reinterpret_cast<const uint4*>(in);
  uint4* o4 = reinterpret_cast<uint4*>(out);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cv);
    const long long pix = i / cv;
    const int x = (int)(pix % S2), y = (int)((pix / S2) % S2);
    const long long b = pix / ((long long)S2 * S2);
    o4[i] = i4[((b * S + (y >> 1)) * S + (x >> 1)) * cv + q];
  }
}

static int64_t vae_dec_param_count(const LdpVaeConfig& c) {
  auto conv = [](int64_t k, int64_t ci, int64_t co) { return k * k * ci * co + co; };
  auto res = [&](int64_t ci, int64_t co) { return 2 * ci + conv(3, ci, co) + 2 * co + conv(3, co, co) + (ci != co ? conv(1, ci, co) : 0); };
  int64_t ch = c.block_out_channels[c.n_blocks - 1];
  int64_t n = conv(1, c.latent_channels, c.latent_channels) + conv(3, c.latent_channels, ch);
  n += res(ch, ch) + 2 * ch + 4 * (ch * ch + ch) + res(ch, ch);
  for (int i = 0; i < c.n_blocks; ++i) {
    const int64_t co = c.block_out_channels[c.n_blocks - 1 - i];
    for (int j = 0; j < c.layers_per_block + 1; ++j) {
      n += res(ch, co);
      ch = co;
    }
    if (i != c.n_blocks - 1) n += conv(3, co, co);
  }
  n += 2 * ch + conv(3, ch, c.in_channels);
  return n;
}

static int vae_dec_create_impl(const LdpVaeConfig* cfg, const float* params_host, uint64_t n_params, LdpVae* h) {
  h->cfg = *cfg;
  h->is_decoder = true;
  const LdpVaeConfig& c = h->cfg;
  const int64_t expect = vae_dec_param_count(c);
  LDP_CHECK((int64_t)n_params == expect, LDP_ERR_PARAM_COUNT,
            "VAE decoder weight blob has " + std::to_string(n_params) + " floats, config needs " + std::to_string(expect));
  LDP_TRY(h->arena.alloc_t(&h->blob, n_params, false));
  LDP_CUDA_OK(cudaMemcpy(h->blob, params_host, n_params * 4, cudaMemcpyHostToDevice));
  BlobWalker w{h->blob, 0};
  int ch = c.block_out_channels[c.n_blocks - 1];
  h->quant = take_conv(w, 1, c.latent_channels, c.latent_channels);
  h->conv_in = take_conv(w, 3, c.latent_channels, ch);
  h->mid0 = take_res(w, ch, ch);
  h->ag_s = w.take(ch); h->ag_b = w.take(ch);
  h->aq.w = w.take((uint64_t)ch * ch); h->aq.b = w.take(ch);
  h->ak.w = w.take((uint64_t)ch * ch); h->ak.b = w.take(ch);
  h->av.w = w.take((uint64_t)ch * ch); h->av.b = w.take(ch);
  h->ap = ConvW();
  h->ap.k = 1; h->ap.cin = ch; h->ap.cout = ch;
  h->ap.w = w.take((uint64_t)ch * ch); h->ap.b = w.take(ch);
  h->mid1 = take_res(w, ch, ch);
  const int c_mid = ch;
  for (int i = 0; i < c.n_blocks; ++i) {
    const int co = c.block_out_channels[c.n_blocks - 1 - i];
    std::vector<ResW> rs;
    for (int j = 0; j < c.layers_per_block + 1; ++j) {
      rs.push_back(take_res(w, ch, co));
      ch = co;
    }
    h->blocks.push_back(rs);
    if (i != c.n_blocks - 1) h->down.push_back(take_conv(w, 3, co, co));
  }
  h->nos = w.take(ch); h->nob = w.take(ch);
  h->conv_out = take_conv(w, 3, ch, c.in_channels);
  LDP_CHECK((int64_t)w.pos == expect, LDP_ERR_PARAM_COUNT, "internal: decoder blob walk mismatch");
  LDP_TRY(h->arena.alloc_t(&h->wqkv, (size_t)c_mid * 3 * c_mid));
  LDP_TRY(h->arena.alloc_t(&h->bqkv, (size_t)3 * c_mid));
  const float* ws[3] = {h->aq.w, h->ak.w, h->av.w};
  const float* bs[3] = {h->aq.b, h->ak.b, h->av.b};
  for (int i = 0; i < 3; ++i) {
    LDP_CUDA_OK(cudaMemcpy2D(h->wqkv + (size_t)i * c_mid, (size_t)3 * c_mid * 4, ws[i], (size_t)c_mid * 4, (size_t)c_mid * 4, c_mid,
                             cudaMemcpyDeviceToDevice));
    LDP_CUDA_OK(cudaMemcpy(h->bqkv + (size_t)i * c_mid, bs[i], (size_t)c_mid * 4, cudaMemcpyDeviceToDevice));
  }
  return LDP_OK;
}

template <typename T>
static int vae_upsample(const T* in, T* out, int nimg, int S, int C, cudaStream_t s) {
  const long long total = (long long)nimg * 4 * S * S * (C / (16 / (int)sizeof(T)));
  vae_upsample2x_kernel<T><<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, s>>>(in, out, nimg, S, C);
  VAE_LAUNCH_OK("vae_upsample2x");
  return LDP_OK;
}

// fp32 program (parity instrument)
static int vae_decode_f32(LdpVae* h, VaeWs* w, const float* z, int nimg, float* out, cudaStream_t s) {
  const LdpVaeConfig& c = h->cfg;
  if (!w->Gf) LDP_TRY(w->arena.alloc_t(&w->Gf, (size_t)w->Bc * vae_max_act(h), false));
  int S = c.image_size >> (c.n_blocks - 1);
  const int ch = c.block_out_channels[c.n_blocks - 1], L = S * S;
  LDP_TRY(vae_conv_f32(h->quant, z, nullptr, w->Gf, nimg, S, 1, 0, s));
  LDP_TRY(vae_conv_f32(h->conv_in, w->Gf, nullptr, w->S, nimg, S, 1, 1, s));
  LDP_TRY(vae_res_f32(h, w, h->mid0, nimg, S, s));
  {
    LDP_TRY(vae_gn(h, w, w->S, nimg, L, ch, h->ag_s, h->ag_b, 0, w->Gf, nullptr, s));
    GemmF32 g;
    g.x1 = w->Gf; g.c1 = ch; g.ld1 = ch; g.w = h->wqkv; g.ldw = 3 * ch; g.bias = h->bqkv; g.out = w->qkv; g.ldo = 3 * ch;
    g.m = nimg * L; g.n = 3 * ch;
    LDP_TRY(launch_gemm_f32(g, s));
    vae_attn_kernel<<<nimg, 256, 0, s>>>(w->qkv, w->Gf, nullptr, L, ch);
    VAE_LAUNCH_OK("vae_attn");
    g = GemmF32();
    g.x1 = w->Gf; g.c1 = ch; g.ld1 = ch; g.w = h->ap.w; g.ldw = ch; g.bias = h->ap.b; g.res = w->S; g.ldres = ch;
    g.out = w->S; g.ldo = ch; g.m = nimg * L; g.n = ch;
    LDP_TRY(launch_gemm_f32(g, s));
  }
  LDP_TRY(vae_res_f32(h, w, h->mid1, nimg, S, s));
  for (int i = 0; i < c.n_blocks; ++i) {
    for (auto& r : h->blocks[i]) LDP_TRY(vae_res_f32(h, w, r, nimg, S, s));
    if (i != c.n_blocks - 1) {
      const int co = h->down[i].cin;
      LDP_TRY(vae_upsample<float>(w->S, w->Gf, nimg, S, co, s));
      S *= 2;
      LDP_TRY(vae_conv_f32(h->down[i], w->Gf, nullptr, w->S, nimg, S, 1, 1, s));
    }
  }
  const int c0 = h->conv_out.cin;
  LDP_TRY(vae_gn(h, w, w->S, nimg, S * S, c0, h->nos, h->nob, 1, w->Gf, nullptr, s));
  return vae_conv_f32(h->conv_out, w->Gf, nullptr, out, nimg, S, 1, 1, s);
}

// tcgen05 program: same two-phase walk as the encoder (build the ops once, then replay)
static int vae_dec_walk_tc(LdpVae* h, VaeWs* w, bool build, const float* z, int nimg, float* out, cudaStream_t s) {
  const LdpVaeConfig& c = h->cfg;
  VaeBufs b{w->S, w->Hf, w->Sb, w->Gb};
  size_t oi = 0;
  int conv_id = 0;
  static const bool fuse_gn = !(getenv("LDP_VAE_FUSE_GN") && getenv("LDP_VAE_FUSE_GN")[0] == '0');
  int fused_slabs = 0;
  auto conv = [&](const ConvW& cw, const __nv_bfloat16* in, int S_in, const float* res, float* of32, __nv_bfloat16* obf,
                  bool gn_next) -> int {
    const TileGeo tg = tile_geo(S_in);
    const int G = c.norm_num_groups, cpg = cw.cout / G;
    const bool fuse = fuse_gn && gn_next && of32 != nullptr && cw.cout % G == 0 && (cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) &&
                      S_in * S_in >= 64 && tg.ib <= 2 && tg.tiles_per_img <= 64;
    fused_slabs = fuse ? tg.tiles_per_img : 0;
    if (build) {
      TcGemm op;
      LDP_TRY(vae_conv_tc(h, w, conv_id, cw, cw.b, in, S_in, 1, &op));
      op.res_f32 = res; op.ld_res_f32 = cw.cout;
      op.out_f32 = of32; op.ld_out_f32 = cw.cout;
      op.out_bf16 = obf; op.ld_out_bf16 = cw.cout;
      if (fuse) {
        op.gn_part = w->part; op.gn_cpg = cpg; op.gn_G = G; op.gn_slabs = tg.tiles_per_img; op.gn_imgs_per_tile = tg.ib;
      }
      LDP_TRY(vae_epi_maps(w->arena, &op, (size_t)w->Bc * S_in * S_in));
      w->ops.push_back(op);
    } else {
      TcGemm op = w->ops[oi];
      op.M = nimg * S_in * S_in;
      LDP_TRY(launch_tc_gemm(op, s));
    }
    ++oi;
    ++conv_id;
    return LDP_OK;
  };
  auto gn = [&](const float* x, int P, int C, const float* gs_, const float* gb_, int act, float* yf, __nv_bfloat16* y) -> int {
    const int slabs = fused_slabs;
    fused_slabs = 0;
    if (build) return LDP_OK;
    return vae_gn(h, w, x, nimg, P, C, gs_, gb_, act, yf, y, s, slabs);
  };
  auto resnet = [&](const ResW& r, int S) -> int {
    const int P = S * S;
    LDP_TRY(gn(b.S, P, r.c1.cin, r.n1s, r.n1b, 1, nullptr, b.Gb));
    LDP_TRY(conv(r.c1, b.Gb, S, nullptr, b.Hf, nullptr, true));
    LDP_TRY(gn(b.Hf, P, r.c1.cout, r.n2s, r.n2b, 1, nullptr, b.Gb));
    if (r.has_sc) {
      LDP_TRY(conv(r.sc, b.Sb, S, nullptr, b.Hf, nullptr, false));
      LDP_TRY(conv(r.c2, b.Gb, S, b.Hf, b.S, b.Sb, true));
    } else {
      LDP_TRY(conv(r.c2, b.Gb, S, b.S, b.S, b.Sb, true));
    }
    return LDP_OK;
  };
  int S = c.image_size >> (c.n_blocks - 1);
  const int ch = c.block_out_channels[c.n_blocks - 1], L = S * S;
  if (!build) {
    // 4-channel head on the SIMT convolution, then the bf16 copy the first resnet's shortcut / the tensor maps read
    LDP_TRY(vae_conv_f32(h->quant, z, nullptr, b.Hf, nimg, S, 1, 0, s));
    LDP_TRY(vae_conv_f32(h->conv_in, b.Hf, nullptr, b.S, nimg, S, 1, 1, s));
    LDP_TRY(launch_cast_bf16(b.S, ch, b.Sb, ch, (long long)nimg * L, ch, 0, s));
  }
  LDP_TRY(resnet(h->mid0, S));
  {
    LDP_TRY(gn(b.S, L, ch, h->ag_s, h->ag_b, 0, nullptr, b.Gb));
    ConvW qkv;
    qkv.k = 1; qkv.cin = ch; qkv.cout = 3 * ch; qkv.w = h->wqkv; qkv.b = h->bqkv;
    LDP_TRY(conv(qkv, b.Gb, S, nullptr, w->qkv, nullptr, false));
    if (!build) {
      vae_attn_kernel<<<nimg, 256, 0, s>>>(w->qkv, nullptr, b.Gb, L, ch);
      VAE_LAUNCH_OK("vae_attn");
    }
    LDP_TRY(conv(h->ap, b.Gb, S, b.S, b.S, b.Sb, true));
  }
  LDP_TRY(resnet(h->mid1, S));
  for (int i = 0; i < c.n_blocks; ++i) {
    for (auto& r : h->blocks[i]) LDP_TRY(resnet(r, S));
    if (i != c.n_blocks - 1) {
      const int co = h->down[i].cin;
      if (!build) LDP_TRY(vae_upsample<__nv_bfloat16>(b.Sb, b.Gb, nimg, S, co, s));
      S *= 2;
      LDP_TRY(conv(h->down[i], b.Gb, S, nullptr, b.S, b.Sb, true));
    }
  }
  const int c0 = h->conv_out.cin;
  LDP_TRY(gn(b.S, S * S, c0, h->nos, h->nob, 1, b.Hf, nullptr));
  if (!build) LDP_TRY(vae_conv_f32(h->conv_out, b.Hf, nullptr, out, nimg, S, 1, 1, s));
  return LDP_OK;
}

static int vae_dec_prepare_tc(LdpVae* h, VaeWs* w) {
  if (w->tc_ready) return LDP_OK;
  LDP_TRY(tc_driver_check());
  LDP_TRY(tc_gemm_init());
  const size_t max_act = vae_max_act(h);
  LDP_TRY(w->arena.alloc_t(&w->Sb, (size_t)w->Bc * max_act));
  LDP_TRY(w->arena.alloc_t(&w->Gb, (size_t)w->Bc * max_act));
  w->ops.clear();
  LDP_TRY(vae_dec_walk_tc(h, w, true, nullptr, w->Bc, nullptr, 0));
  w->tc_ready = true;
  return LDP_OK;
}

}  // namespace ldp

extern "C" {

int64_t ldp_vae_param_count(const LdpVaeConfig* cfg) {
  if (vae_validate(cfg) != LDP_OK) return -1;
  return vae_param_count(*cfg);
}

int ldp_vae_create(const LdpVaeConfig* cfg, const float* params_host, uint64_t n_params, LdpVae** out) {
  LDP_CHECK(out != nullptr && params_host != nullptr, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_TRY(vae_validate(cfg));
  LdpVae* h = new LdpVae();
  int st = vae_create_impl(cfg, params_host, n_params, h);
  if (st != LDP_OK) {
    delete h;
    return st;
  }
  *out = h;
  return LDP_OK;
}

int ldp_vae_destroy(LdpVae* h) {
  if (!h) return LDP_OK;
  cudaDeviceSynchronize();
  delete h;
  return LDP_OK;
}

int ldp_vae_encode(LdpVae* h, int precision, const void* images_dev, int pixel_format, int B, float lat_min, float lat_max,
                   float* latent_dev, void* cuda_stream) {
  LDP_CHECK(h && images_dev && latent_dev, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_CHECK(!h->is_decoder, LDP_ERR_INVALID_ARG, "handle is a decoder (use ldp_vae_create)");
  LDP_CHECK(B > 0, LDP_ERR_BAD_SHAPE, "B must be positive");
  LDP_CHECK(pixel_format == 0 || pixel_format == 1, LDP_ERR_INVALID_ARG, "pixel_format must be 0 (uint8) or 1 (float32)");
  LDP_CHECK(precision == LDP_PREC_FP32 || precision == LDP_PREC_BF16, LDP_ERR_INVALID_ARG, "unknown precision");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const LdpVaeConfig& c = h->cfg;
  const int S = c.image_size, hw = S >> (c.n_blocks - 1);
  const size_t px_bytes = pixel_format == 0 ? 1 : 4;
  // 592 = 4 x 148 images: at 64x64 pixels every level of a 4-block encoder is then a whole number of tiles per SM (level 3: 2 M tiles x
  // 2 N tiles per CTA pair), so the persistent convolutions have no partial last wave; 256 -> 296 -> 592 images: +2.3 % / +4.1 % images/s
  static const int chunk_max = []() { const char* e = getenv("LDP_VAE_CHUNK"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 592; }();
  const int chunk = std::min(B, chunk_max);
  VaeWs* w;
  LDP_TRY(vae_get_ws(h, chunk, &w));
  if (precision == LDP_PREC_BF16) LDP_TRY(vae_prepare_tc(h, w));
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = std::min(chunk, B - b0);
    const void* img = (const char*)images_dev + (size_t)b0 * S * S * c.in_channels * px_bytes;
    float* out = latent_dev + (size_t)b0 * hw * hw * c.latent_channels;
    if (precision == LDP_PREC_FP32) LDP_TRY(vae_forward_f32(h, w, img, pixel_format, n, lat_min, lat_max, out, s));
    else LDP_TRY(vae_walk_tc(h, w, false, img, pixel_format, n, lat_min, lat_max, out, s));
  }
  return LDP_OK;
}

int64_t ldp_vae_decoder_param_count(const LdpVaeConfig* cfg) {
  if (vae_validate(cfg) != LDP_OK) return -1;
  return vae_dec_param_count(*cfg);
}

int ldp_vae_decoder_create(const LdpVaeConfig* cfg, const float* params_host, uint64_t n_params, LdpVae** out) {
  LDP_CHECK(out != nullptr && params_host != nullptr, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_TRY(vae_validate(cfg));
  LdpVae* h = new LdpVae();
  int st = vae_dec_create_impl(cfg, params_host, n_params, h);
  if (st != LDP_OK) {
    delete h;
    return st;
  }
  *out = h;
  return LDP_OK;
}

int ldp_vae_decode(LdpVae* h, int precision, const float* latent_dev, int B, float* images_dev, void* cuda_stream) {
  LDP_CHECK(h && latent_dev && images_dev, LDP_ERR_INVALID_ARG, "null pointer");
  LDP_CHECK(h->is_decoder, LDP_ERR_INVALID_ARG, "handle is an encoder (use ldp_vae_decoder_create)");
  LDP_CHECK(B > 0, LDP_ERR_BAD_SHAPE, "B must be positive");
  LDP_CHECK(precision == LDP_PREC_FP32 || precision == LDP_PREC_BF16, LDP_ERR_INVALID_ARG, "unknown precision");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const LdpVaeConfig& c = h->cfg;
  const int S = c.image_size, hw = S >> (c.n_blocks - 1);
  // 296 = 2 x 148 frames per pass (128 / 148 / 256 / 296: 12.8 / 13.4 / 13.4 / 13.5 - 13.7 k frames/s)
  static const int dec_chunk_max = []() { const char* e = getenv("LDP_VAE_DEC_CHUNK"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 296; }();
  const int chunk = std::min(B, dec_chunk_max);
  VaeWs* w;
  LDP_TRY(vae_get_ws(h, chunk, &w));
  if (precision == LDP_PREC_BF16) LDP_TRY(vae_dec_prepare_tc(h, w));
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = std::min(chunk, B - b0);
    const float* z = latent_dev + (size_t)b0 * hw * hw * c.latent_channels;
    float* out = images_dev + (size_t)b0 * S * S * c.in_channels;
    if (precision == LDP_PREC_FP32) LDP_TRY(vae_decode_f32(h, w, z, n, out, s));
    else LDP_TRY(vae_dec_walk_tc(h, w, false, z, n, out, s));
  }
  return LDP_OK;
}

}  // extern "C"
