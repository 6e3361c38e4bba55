// tcgen05 implicit-GEMM kernel for sm_100a: the one tensor-core kernel behind every contraction of the
// planner UNet (k5 / strided / transposed / 1x1 1-D convolutions), the IDM MLP and the VAE encoder convs.
//
//   C[128 x BN tile] = sum over 64-wide K blocks  A_tile(kb) [128 x 64 bf16]  x  W^T_tile(kb) [BN x 64 bf16]
//
// * A tiles are fetched by TMA straight out of the channels-last activation tensor through 4-D tensor maps
//   (channels, d1, d2, items): a convolution tap is just a coordinate offset, and the zero padding of the
//   convolution is TMA's out-of-bounds zero fill - no im2col buffer, no halo copies.
// * W^T tiles come from weights pre-packed K-major ([N_pad][K_pad] bf16) at handle creation.
// * Both land in 128B-swizzled shared memory; one elected thread issues tcgen05.mma (M=128, N=BN, K=16)
//   with the fp32 accumulator in TMEM; a 4-stage mbarrier ring overlaps TMA with MMA.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue
//   (each owns one 32-lane quarter of TMEM; thread i owns output row 32*quarter+i).
// * Fused epilogues (all fp32 math on the accumulator, read with tcgen05.ld):
//     PLAIN  bias (+ReLU) (+residual) -> f32 and/or bf16
//     GN     bias -> GroupNorm over (rows of one sample x group channels) -> Mish -> [FiLM] -> [+residual] -> bf16
//            (the tile owns whole samples and whole groups, so the statistics never leave the CTA)
//     DDPM   bias -> eps; x0 = clip((x - s eps)/a); x <- c0 x0 + ct x + sigma z   (the scheduler step of the
//            reverse-diffusion loop, reference agent/ldp_agent.py:470-471, fused around the score-net's last GEMM)
//     LN     h = acc + bias + residual -> f32;  LayerNorm(h) (or ReLU(h)) -> bf16   (IDM MLPResNet block)
#include "kernels.h"

namespace ldp {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KB
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_KB_SMEM = 320;

template <int BN>
struct TcSmem {
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int TOTAL = TC_STAGES * STAGE_BYTES + 1024;   // + alignment slack
};

// ---- epilogue helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32], bool vec_ok, int nvalid) {
  if (vec_ok && nvalid == 32) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u;
      u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
      u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
      u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
      u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
      d4[j] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) dst[i] = __float2bfloat16(v[i]);
  }
}

__device__ __forceinline__ void store_f32x32(float* dst, const float (&v)[32], bool vec_ok, int nvalid) {
  if (vec_ok && nvalid == 32) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) dst[i] = v[i];
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_plain(const TcGemm& p, uint32_t taddr, int m, int n0) {
  const bool row_ok = m < p.M;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    float v[32];
    tmem_ld_32x32(taddr + c * 32, v);
    const int nb = n0 + c * 32;
    const int nvalid = min(32, p.N - nb);
    if (nvalid <= 0) continue;           // uniform across the warp
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      int n = nb + i;
      float x = v[i];
      if (i < nvalid) {
        if (p.bias) x += __ldg(p.bias + n);
        if (p.relu) x = fmaxf(x, 0.f);
        if (row_ok) {
          if (p.res_f32) x += p.res_f32[(long long)m * p.ld_res_f32 + n];
          if (p.res_bf16) x += __bfloat162float(p.res_bf16[(long long)m * p.ld_res_bf16 + n]);
        }
      }
      v[i] = x;
    }
    if (row_ok) {
      if (p.out_f32) store_f32x32(p.out_f32 + (long long)m * p.ld_out_f32 + nb, v, (p.ld_out_f32 & 3) == 0, nvalid);
      if (p.out_bf16) store_bf16x32(p.out_bf16 + (long long)m * p.ld_out_bf16 + nb, v, (p.ld_out_bf16 & 7) == 0, nvalid);
    }
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_gn(const TcGemm& p, uint32_t taddr, int m, int n0) {
  constexpr int NC = BN / 32;
  const int T = p.rows_per_item;
  const int cpg = p.group_width >> 5;                    // 32-column chunks per group: 1, 2, 4 (or 8)
  const int nchunks = min(NC, (p.N - n0) >> 5);          // N and group widths are multiples of 32 here
  const bool row_ok = m < p.M;
  float cs[NC], css[NC];
  // pass 1: per-chunk sums of (acc + bias) and its square over this thread's row
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    cs[c] = 0.f;
    css[c] = 0.f;
    if (c < nchunks) {
      float v[32];
      tmem_ld_32x32(taddr + c * 32, v);
      const float* bp = p.bias + n0 + c * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = v[i] + __ldg(bp + i);
        cs[c] += x;
        css[c] += x * x;
      }
    }
  }
  // group totals (same value replicated on every chunk of the group), then across the T rows of the sample
  float mean[NC], rstd[NC];
  const float inv_cnt = 1.f / (float)(T * p.group_width);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int c2 = 0; c2 < NC; ++c2) {
      bool same = (c2 / cpg) == (c / cpg);
      s += same ? cs[c2] : 0.f;
      ss += same ? css[c2] : 0.f;
    }
    for (int off = 1; off < T; off <<= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    float mu = s * inv_cnt;
    float var = fmaxf(ss * inv_cnt - mu * mu, 0.f);
    mean[c] = mu;
    rstd[c] = rsqrtf(var + p.eps);
  }
  // pass 2: normalise -> Mish -> FiLM -> residual -> bf16
  const int b = m / T;
  const float* trow = nullptr;
  const float* orow = nullptr;
  if (p.film) {
    trow = p.ttab + (long long)step_of(p.step, row_ok ? m : 0) * p.ld_ttab + p.film_off;
    orow = p.otab + (long long)(row_ok ? b : 0) * p.ld_otab + p.film_off;
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < nchunks) {
      float v[32];
      tmem_ld_32x32(taddr + c * 32, v);
      const int nb = n0 + c * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        int n = nb + i;
        float x = v[i] + __ldg(p.bias + n);
        x = (x - mean[c]) * rstd[c] * __ldg(p.gamma + n) + __ldg(p.beta + n);
        x = mish_f<true>(x);
        if (p.film) x = (__ldg(trow + n) + orow[n]) * x + (__ldg(trow + p.film_c + n) + orow[p.film_c + n]);
        v[i] = x;
      }
      if (p.use_aux) {
        float r[32];
        tmem_ld_32x32(taddr + BN + c * 32, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += r[i] + __ldg(p.bias_aux + nb + i);
      } else if (p.res_bf16 && row_ok) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res_bf16 + (long long)m * p.ld_res_bf16 + nb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u = rp[j];
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 f = __bfloat1622float2(h[q]);
            v[8 * j + 2 * q] += f.x;
            v[8 * j + 2 * q + 1] += f.y;
          }
        }
      }
      if (row_ok) {
        if (p.out_bf16) store_bf16x32(p.out_bf16 + (long long)m * p.ld_out_bf16 + nb, v, true, 32);
        if (p.out_f32) store_f32x32(p.out_f32 + (long long)m * p.ld_out_f32 + nb, v, true, 32);
      }
    }
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_ddpm(const TcGemm& p, uint32_t taddr, int m, int n0) {
  const bool row_ok = m < p.M;
  const int t = step_of(p.step, 0);
  const float* cf = p.coef + t * 8;
  const float inv_sa = cf[0], s1a = cf[1], c0 = cf[2], ct = cf[3], sigma = cf[4], sap = cf[5], s1ap = cf[6];
  const DdpmCall call = p.call_dev ? *p.call_dev : p.call;
  const float* noise = call.noise ? call.noise + (long long)(call.n_steps - 1 - t) * call.noise_step_stride : nullptr;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    float v[32];
    tmem_ld_32x32(taddr + c * 32, v);
    const int nb = n0 + c * 32;
    const int nvalid = min(32, p.N - nb);
    if (nvalid <= 0 || !row_ok) continue;
    float* xr = p.x_io + (long long)m * p.ld_x + nb;
    const long long e0 = (long long)m * p.N + nb;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < nvalid) {
        float e = v[i] + __ldg(p.bias + nb + i);
        float x = xr[i];
        float x0 = fminf(fmaxf((x - s1a * e) * inv_sa, -1.f), 1.f);
        float o;
        if (call.sampler == LDP_SAMPLER_DDIM) {
          o = sap * x0 + s1ap * e;
        } else {
          o = c0 * x0 + ct * x;
          if (t > 0) {
            float z = noise ? noise[e0 + i]
                            : philox_normal(call.seed, call.stream_id, (uint32_t)t,
                                            (unsigned long long)(call.elem_offset + e0 + i));
            o += sigma * z;
          }
        }
        xr[i] = o;
        v[i] = o;
      }
    }
    if (p.out_bf16) store_bf16x32(p.out_bf16 + (long long)m * p.ld_out_bf16 + nb, v, false, nvalid);
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_ln(const TcGemm& p, uint32_t taddr, int m, int n0) {
  // requires N == BN (the whole feature row lives in this tile)
  const bool row_ok = m < p.M;
  float s = 0.f, ss = 0.f;
  float* hrow = p.out_f32 + (long long)(row_ok ? m : 0) * p.ld_out_f32;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    float v[32];
    tmem_ld_32x32(taddr + c * 32, v);
    const int nb = n0 + c * 32;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x = v[i] + __ldg(p.bias + nb + i);
      if (p.res_f32 && row_ok) x += p.res_f32[(long long)m * p.ld_res_f32 + nb + i];
      s += x;
      ss += x * x;
      v[i] = x;
    }
    if (row_ok) store_f32x32(hrow + nb, v, (p.ld_out_f32 & 3) == 0, 32);
  }
  const float mu = s / (float)BN;
  const float rs = rsqrtf(fmaxf(ss / (float)BN - mu * mu, 0.f) + p.eps);
  if (!row_ok || !p.out_bf16) return;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    float v[32];
    const int nb = n0 + c * 32;
    const float4* h4 = reinterpret_cast<const float4*>(hrow + nb);   // written above by this same thread
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 f = h4[j];
      v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x = v[i];
      v[i] = p.relu ? fmaxf(x, 0.f) : (x - mu) * rs * __ldg(p.gamma + nb + i) + __ldg(p.beta + nb + i);
    }
    store_bf16x32(p.out_bf16 + (long long)m * p.ld_out_bf16 + nb, v, (p.ld_out_bf16 & 7) == 0, 32);
  }
}

// ---- the kernel --------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemm p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_accum;
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) TcKBlock kb_s[TC_MAX_KB_SMEM];     // K-block table staged once per CTA

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int n0 = tile_n * BN;
  constexpr uint32_t NCOLS_MAIN = BN;
  const uint32_t ncols = p.use_aux ? 2 * BN : BN;     // 128 / 256 / 512: powers of two >= 32

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_accum), 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.map_b);
    tma_prefetch_desc(&p.map_a[0]);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), ncols);
    tmem_relinquish();
  }
  const bool kb_in_smem = p.num_kb <= TC_MAX_KB_SMEM;
  if (kb_in_smem)
    for (int i = threadIdx.x; i < p.num_kb; i += TC_THREADS) kb_s[i] = p.kb[i];
  const TcKBlock* kbt = kb_in_smem ? kb_s : p.kb;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int q = tile_m / p.tiles_per_item, r = tile_m - q * p.tiles_per_item;
      const int c2_base = r * p.rows_step, c3 = q * p.items_per_tile;
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const TcKBlock e = kbt[kb];
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        const uint32_t bar = smem_u32(&bar_full[stage]);
        const uint32_t sa = smem_base + stage * TcSmem<BN>::STAGE_BYTES;
        const uint32_t sb = sa + TC_A_BYTES;
        mbar_arrive_expect_tx(bar, TcSmem<BN>::STAGE_BYTES);
        tma_load_4d(sa, &p.map_a[e.src_acc & 0xff], bar, e.c0, e.d1, c2_base + e.d2, c3);
        tma_load_2d(sb, &p.map_b, bar, kb * TC_BK, n0);
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, BN);
      uint32_t stage = 0, phase = 0;
      uint32_t started = 0;    // bit a set once accumulator a has received its first MMA
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const uint32_t acc = ((uint32_t)kbt[kb].src_acc >> 8) & 0xffu;
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * TcSmem<BN>::STAGE_BYTES;
        const uint64_t da = umma_desc_sw128(sa);
        const uint64_t db = umma_desc_sw128(sa + TC_A_BYTES);
        const uint32_t d_tmem = tmem_base + acc * NCOLS_MAIN;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128B swizzle row: +2 in the (addr >> 4) field
          umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, ((started >> acc) & 1u) | (k > 0 ? 1u : 0u));
        }
        started |= 1u << acc;
        umma_commit(smem_u32(&bar_empty[stage]));       // frees the smem stage when these MMAs retire
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(smem_u32(&bar_accum));                // accumulator(s) complete
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    const int m = tile_m * TC_BM + quarter * 32 + lane;
    mbar_wait(smem_u32(&bar_accum), 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    switch (p.mode) {
      case TC_EPI_PLAIN: epilogue_plain<BN>(p, taddr, m, n0); break;
      case TC_EPI_GN:    epilogue_gn<BN>(p, taddr, m, n0); break;
      case TC_EPI_DDPM:  epilogue_ddpm<BN>(p, taddr, m, n0); break;
      default:           epilogue_ln<BN>(p, taddr, m, n0); break;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    else (void)cudaGetLastError();
  }
  return fn;
}

int tc_driver_check() {
  LDP_CHECK(get_encode_fn() != nullptr, LDP_ERR_NO_DEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  return LDP_OK;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_fn();
  LDP_CHECK(fn != nullptr, LDP_ERR_NO_DEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  LDP_CHECK(rank >= 2 && rank <= 4, LDP_ERR_INVALID_ARG, "tensor map rank must be 2..4");
  cuuint64_t gdim[4];
  cuuint64_t gstr[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      LDP_CHECK((gstr[i - 1] & 15) == 0, LDP_ERR_INVALID_ARG, "tensor map strides must be multiples of 16 bytes");
    }
  }
  LDP_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, LDP_ERR_INVALID_ARG, "tensor map base must be 16B aligned");
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (rank " +
                   std::to_string(rank) + ", dim0 " + std::to_string(dims[0]) + ", box0 " + std::to_string(box[0]) + ")");
    return LDP_ERR_CUDA;
  }
  return LDP_OK;
}

int tc_gemm_init() {
  static bool done = false;
  if (done) return LDP_OK;
  LDP_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<128>::TOTAL));
  LDP_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<256>::TOTAL));
  done = true;
  return LDP_OK;
}

template <int BN>
static int launch_tc_gemm_bn(const TcGemm& p, cudaStream_t s) {
  LDP_TRY(tc_gemm_init());
  dim3 grid(ceil_div(p.M, TC_BM), ceil_div(p.N, BN));
  tc_gemm_kernel<BN><<<grid, TC_THREADS, TcSmem<BN>::TOTAL, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(std::string("tc_gemm launch failed: ") + cudaGetErrorString(e));
    return LDP_ERR_CUDA;
  }
  count_launch();
  return LDP_OK;
}

int launch_tc_gemm(const TcGemm& p, cudaStream_t s) {
  LDP_CHECK(p.kb && p.num_kb > 0 && p.M > 0 && p.N > 0, LDP_ERR_INVALID_ARG, "tc_gemm: bad arguments");
  LDP_CHECK(p.block_n == 128 || p.block_n == 256, LDP_ERR_INVALID_ARG, "tc_gemm: block_n must be 128 or 256");
  if (p.mode == TC_EPI_GN) {
    LDP_CHECK(p.group_width % 32 == 0 && p.group_width <= p.block_n && p.N % p.group_width == 0, LDP_ERR_UNSUPPORTED,
              "tc_gemm GN epilogue: group width must be a multiple of 32 and fit the N tile");
    LDP_CHECK(p.rows_per_item >= 1 && p.rows_per_item <= 32 && (p.rows_per_item & (p.rows_per_item - 1)) == 0,
              LDP_ERR_UNSUPPORTED, "tc_gemm GN epilogue: rows per sample must be a power of two <= 32");
    LDP_CHECK(p.bias && p.gamma && p.beta, LDP_ERR_INVALID_ARG, "tc_gemm GN epilogue: bias/gamma/beta required");
  }
  if (p.mode == TC_EPI_LN) LDP_CHECK(p.N == p.block_n && p.out_f32 && p.bias, LDP_ERR_UNSUPPORTED, "tc_gemm LN epilogue: N must equal block_n");
  return p.block_n == 128 ? launch_tc_gemm_bn<128>(p, s) : launch_tc_gemm_bn<256>(p, s);
}

}  // namespace ldp
