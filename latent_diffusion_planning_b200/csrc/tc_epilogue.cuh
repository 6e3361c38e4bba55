// Device-side pieces shared by the tcgen05 kernels (tc_gemm.cu: one launch per layer; planner_loop.cu: the persistent
// reverse-diffusion kernel): tile geometry, per-CTA epilogue staging, and the fused epilogues.  See tc_gemm.cu for the
// description of the layouts.
#pragma once
#include <cuda_fp16.h>

#include "kernels.h"

namespace ldp {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KB
// Thread geometry: warp 0 = TMA producer, warp 1 = MMA issuer, then the epilogue warps.  A warp may only read the TMEM
// lane quarter (warp index mod 4), so the epilogue warps come in groups of four (one per quarter); group g owns column
// slice g of the tile.  BN = 128 runs 16 epilogue warps with one 32-column chunk per thread: the epilogue is a long
// dependent chain (TMEM load -> statistics -> barrier -> MUFU-heavy activation -> FiLM / residual loads -> store), and
// four resident warps per scheduler are what hides its latencies (8 warps took 13-15k cycles per tile, as long as
// the main loop itself; profiles/).
template <int BN>
struct TcGeo {
  static constexpr int EPI_WARPS = BN == 64 ? 8 : 16;
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int THREADS = 64 + EPI_THREADS;
  static constexpr int PARTS = EPI_WARPS / 4;              // column slices of the tile
  static constexpr int CPP = BN / 32 / PARTS;              // 32-column chunks per thread: 1 (BN 64, 128), 2 (BN 256)
};
constexpr int TC_MAX_KB_SMEM = 256;
constexpr int TC_MAX_RUNS = 64;            // (taps x sources) + aux sources of any layer
constexpr int TC_MAX_STAGES = 8;

constexpr int TC_SMEM_RING = 196 * 1024;        // shared-memory budget of the TMA ring (static smem takes <= 20 KB more)
constexpr int TC_SMEM_RING_PLAIN = 204 * 1024;  // PLAIN epilogue kernels (ring + TMA-epilogue staging): their static part is <= 19.1 KB

template <int BN>
struct TcSmem {
  static constexpr int B_BYTES = BN * TC_BK * 2;
};
// ring geometry of a launch: a stage holds one A tile and w_max W tiles
static inline int tc_stage_bytes(int bn, int w_max) { return TC_A_BYTES + w_max * bn * TC_BK * 2; }
static inline int tc_num_stages(int bn, int w_max) {
  int n = TC_SMEM_RING / tc_stage_bytes(bn, w_max);
  return n > TC_MAX_STAGES ? TC_MAX_STAGES : n;
}

// Per-CTA staging of everything the epilogue needs per output column (filled while the main loop runs).
template <int BN>
struct EpiSmem {
  float bias[BN];
  float bias2[BN];      // bias of the aux accumulator (residual 1x1 projection)
  float gamma[BN];
  float beta[BN];
  float fscale[BN];     // FiLM scale / shift, time part (valid when the whole launch shares one timestep)
  float fshift[BN];
  uint32_t film_t[BN];  // the same as half2(scale, shift): what the prefetch adds to the observation part
  float2 part[TC_BM][BN / 32];   // per-row, per-32-column-chunk (sum, sum of squares)
};

template <int BN>
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TcGeo<BN>::EPI_THREADS) : "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Mish(y) = y tanh(softplus(y)) = y n/(n+2) = y - 2y/(n+2), n = e^y (e^y + 2).  Written as y - 2 y r with
// r = 1/(e (e+2) + 2): when e^y overflows, r = 0 and the result is y (no clamp needed); 7 instructions, 2 of them MUFU.
__device__ __forceinline__ float mish_fast(float y) {
  const float e = ex2_approx(y * 1.4426950408889634f);
  const float r = rcp_approx(fmaf(e, e + 2.f, 2.f));
  return fmaf(-2.f * r, y, y);
}
// swish(y) = y / (1 + e^-y)
__device__ __forceinline__ float swish_fast(float y) {
  const float e = ex2_approx(fminf(-y, 80.f) * 1.4426950408889634f);
  return y * rcp_approx(1.f + e);
}

// ---- vector load/store helpers (32 consecutive columns of one row) -------------------------------------
// Activations written earlier by other SMs (residual rows, x) are read with ld.global.cg: inside the persistent
// reverse-diffusion kernel there is no kernel boundary to invalidate L1 between a layer's writes and a later layer's reads.
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  The epilogues own one row per thread, so every warp-level
// access touches 32 different lines whatever its width; the LSU cost is per line touched, i.e. per instruction, and a
// 32-byte access halves the instruction count of a 16-byte one (ncu on the VAE's residual convolution: `lg throttle`
// was the top stall of the epilogue warps, 16 data bytes per sector).
struct U32x8 { uint32_t r[8]; };
__device__ __forceinline__ void st_global_256(void* p, const U32x8& u) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(u.r[0]), "r"(u.r[1]), "r"(u.r[2]), "r"(u.r[3]),
               "r"(u.r[4]), "r"(u.r[5]), "r"(u.r[6]), "r"(u.r[7]) : "memory");
}
__device__ __forceinline__ U32x8 ld_global_cg_256(const void* p) {
  U32x8 u;
  asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(u.r[0]), "=r"(u.r[1]), "=r"(u.r[2]), "=r"(u.r[3]),
               "=r"(u.r[4]), "=r"(u.r[5]), "=r"(u.r[6]), "=r"(u.r[7]) : "l"(p) : "memory");
  return u;
}
__device__ __forceinline__ bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32], bool vec_ok, int nvalid) {
  if (vec_ok && nvalid == 32 && aligned32(dst)) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      U32x8 u;
#pragma unroll
      for (int q = 0; q < 8; ++q) u.r[q] = pack_bf16x2(v[16 * j + 2 * q], v[16 * j + 2 * q + 1]);
      st_global_256(dst + 16 * j, u);
    }
  } else if (vec_ok && nvalid == 32) {
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
  if (vec_ok && nvalid == 32 && aligned32(dst)) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      U32x8 u;
#pragma unroll
      for (int q = 0; q < 8; ++q) u.r[q] = __float_as_uint(v[8 * j + q]);
      st_global_256(dst + 8 * j, u);
    }
  } else if (vec_ok && nvalid == 32) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) dst[i] = v[i];
  }
}

__device__ __forceinline__ void add_f32x32(float (&v)[32], const float* src, bool vec_ok, int nvalid) {
  if (vec_ok && nvalid == 32 && aligned32(src)) {
    U32x8 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = ld_global_cg_256(src + 8 * j);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 8; ++q) v[8 * j + q] += __uint_as_float(u[j].r[q]);
  } else if (vec_ok && nvalid == 32) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 f = __ldcg(s4 + j);
      v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) v[i] += __ldcg(src + i);
  }
}

__device__ __forceinline__ void add_bf16x32(float (&v)[32], const __nv_bfloat16* src, bool vec_ok, int nvalid) {
  if (vec_ok && nvalid == 32 && aligned32(src)) {
    U32x8 u[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) u[j] = ld_global_cg_256(src + 16 * j);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[j].r[q]));
        v[16 * j + 2 * q] += f.x;
        v[16 * j + 2 * q + 1] += f.y;
      }
  } else if (vec_ok && nvalid == 32) {
    const uint4* rp = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u = __ldcg(rp + j);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 f = __bfloat1622float2(h[q]);
        v[8 * j + 2 * q] += f.x;
        v[8 * j + 2 * q + 1] += f.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) v[i] += __bfloat162float(__ldcg(src + i));
  }
}

// v[i] += s[i] for 32 consecutive floats of a shared-memory vector (LDS.128, broadcast across the warp)
__device__ __forceinline__ void add_smem32(float (&v)[32], const float* s) {
  const float4* s4 = reinterpret_cast<const float4*>(s);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 f = s4[j];
    v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
  }
}

// v = sum_j acc_j[row + shift_j][32-column chunk c]  (the recombination of a k-tap convolution computed as k un-shifted
// GEMMs).  Thread `lane` of the warp owns TMEM lane = output row; rows of one sample are adjacent lanes, so a shifted
// row is a warp shuffle away, and rows that fall outside the sample contribute zero (the convolution's zero padding).
template <int BN>
__device__ __forceinline__ void load_acc_chunk(const TcGemm& p, uint32_t taddr, int c, int lane, float (&v)[32]) {
  tmem_ld_32x32(taddr + c * 32, v);
  if ((p.n_acc == 1 && p.shift[0] == 0) || (p.epi_skip & 16)) return;       // uniform: dense GEMM / per-tap mode
  const int T = p.rows_per_item, t = lane & (T - 1);
  {
    const int s = p.shift[0];
    if (s != 0) {
      const bool ok = (unsigned)(t + s) < (unsigned)T;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float q = __shfl_sync(0xffffffffu, v[i], (lane + s) & 31);
        v[i] = ok ? q : 0.f;
      }
    }
  }
#pragma unroll 1
  for (int j = 1; j < p.n_acc; ++j) {
    float r[32];
    tmem_ld_32x32(taddr + j * BN + c * 32, r);
    const int s = p.shift[j];
    if (s != 0) {                                     // uniform
      const bool ok = (unsigned)(t + s) < (unsigned)T;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float q = __shfl_sync(0xffffffffu, r[i], (lane + s) & 31);
        v[i] += ok ? q : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += r[i];
    }
  }
}

// ---- epilogues: thread owns row `m`; this warp covers chunks [c_begin, c_begin + CPP) of the N tile ------------
// Per-group (sum, sum of squares) of one 32-column chunk over the 32 rows of this warp, NGR groups per chunk
// (32 / NGR channels each).  Fixed xor-shuffle tree: deterministic.  Lane 0 returns the totals through `out`.
template <int NGR>
__device__ __forceinline__ void chunk_group_sums(const float (&v)[32], bool row_ok, float2* out, int lane) {
  constexpr int W = 32 / NGR;
  float s[NGR], ss[NGR];
#pragma unroll
  for (int g = 0; g < NGR; ++g) {
    s[g] = 0.f; ss[g] = 0.f;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const float x = row_ok ? v[g * W + i] : 0.f;
      s[g] += x;
      ss[g] = fmaf(x, x, ss[g]);
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
    for (int g = 0; g < NGR; ++g) {
      s[g] += __shfl_xor_sync(0xffffffffu, s[g], off);
      ss[g] += __shfl_xor_sync(0xffffffffu, ss[g], off);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int g = 0; g < NGR; ++g) out[g] = make_float2(s[g], ss[g]);
  }
}

// combine the lane quarters of each image in a fixed order and publish this tile's partial GroupNorm sums
template <int BN>
__device__ __forceinline__ void gn_part_publish(const TcGemm& p, EpiSmem<BN>& es, int n0, int tile_m) {
  constexpr int NC = BN / 32;
  float2* gscr = &es.part[0][0];
  epi_bar<BN>();
  const int ngr = 32 / p.gn_cpg, ipt = p.gn_imgs_per_tile, qpi = 4 / ipt;      // groups per chunk, images per tile, quarters per image
  const int nchunks = min(NC, (p.N - n0 + 31) >> 5);
  const int q = tile_m / p.tiles_per_item, r = tile_m - q * p.tiles_per_item;
  for (int t = (int)threadIdx.x - 64; t < nchunks * ngr * ipt; t += TcGeo<BN>::EPI_THREADS) {
    const int img = t / (nchunks * ngr), cg = t - img * (nchunks * ngr);
    const int c = cg / ngr, g = cg - c * ngr;
    float a = 0.f, a2 = 0.f;
    for (int qq = img * qpi; qq < (img + 1) * qpi; ++qq) {
      const float2 f = gscr[(qq * NC + c) * 8 + g];
      a += f.x;
      a2 += f.y;
    }
    const long long b = (long long)q * ipt + img;
    float* dst = p.gn_part + ((b * p.gn_slabs + r) * p.gn_G + (n0 + c * 32) / p.gn_cpg + g) * 2;
    dst[0] = a;
    dst[1] = a2;
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_plain(const TcGemm& p, EpiSmem<BN>& es, uint32_t taddr, int m, int n0,
                                               int c_begin, int lane, int tile_m = 0, int quarter = 0) {
  constexpr int CPP = TcGeo<BN>::CPP, NC = BN / 32;
  const bool row_ok = m < p.M;
  const bool vf = (p.ld_out_f32 & 3) == 0, vb = (p.ld_out_bf16 & 7) == 0;
  const bool vrf = (p.ld_res_f32 & 3) == 0, vrb = (p.ld_res_bf16 & 7) == 0;
  float2* gscr = &es.part[0][0];              // scratch [quarter][chunk][8]: `part` is unused by this epilogue otherwise
#pragma unroll 1
  for (int cc = 0; cc < CPP; ++cc) {
    const int c = c_begin + cc;
    const int nb = n0 + c * 32;
    const int nvalid = min(32, p.N - nb);
    if (nvalid <= 0) continue;           // uniform across the warp
    float v[32];
    load_acc_chunk<BN>(p, taddr, c, lane, v);
    add_smem32(v, es.bias + c * 32);
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (row_ok && !(p.epi_skip & 1)) {
      if (p.res_f32) add_f32x32(v, p.res_f32 + (long long)m * p.ld_res_f32 + nb, vrf, nvalid);
      if (p.res_bf16) add_bf16x32(v, p.res_bf16 + (long long)m * p.ld_res_bf16 + nb, vrb, nvalid);
      if (p.out_f32) store_f32x32(p.out_f32 + (long long)m * p.ld_out_f32 + nb, v, vf, nvalid);
      if (p.out_bf16) store_bf16x32(p.out_bf16 + (long long)m * p.ld_out_bf16 + nb, v, vb, nvalid);
    }
    if (p.gn_part && !(p.epi_skip & 2)) {                      // uniform
      float2* o = gscr + (quarter * NC + c) * 8;
      if (p.gn_cpg == 4) chunk_group_sums<8>(v, row_ok, o, lane);
      else if (p.gn_cpg == 8) chunk_group_sums<4>(v, row_ok, o, lane);
      else if (p.gn_cpg == 16) chunk_group_sums<2>(v, row_ok, o, lane);
      else chunk_group_sums<1>(v, row_ok, o, lane);
    }
  }
  if (p.gn_part) gn_part_publish<BN>(p, es, n0, tile_m);
}

// ---- PLAIN epilogue through TMA (TcGemm::epi_tma) ------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const void* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ float4 lds_128(uint32_t a) {
  float4 f;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(a) : "memory");
  return f;
}
__device__ __forceinline__ void sts_128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

constexpr int TC_EPI_BUF = 4096;     // per epilogue warp: 32 rows x 128 bytes (TcGemm::epi_tma bit 3: 2048 = 32 rows x 64 bytes)

// The warp (lane quarter `quarter`, chunks [c_begin, c_begin + CPP)) handles 32 rows x 32 columns at a time through `buf`:
//   [res_f32 box arrives by TMA (issued by the caller for the first box of the tile, before the accumulator wait)] -> v = acc + bias (+ res)
//   -> v to buf (f32, 128B swizzle: 16-byte unit u of row r sits at u ^ (r & 7), conflict-free for st.shared.v4) -> TMA store
//   -> bf16(v) to buf (64B swizzle: unit u of row r at u ^ ((r >> 1) & 3)) -> TMA store.
// HALF: the f32 boxes are 16 columns wide (64-byte rows, 64B swizzle, two per chunk) so that a warp's buffer is 2 KB: the 32 KB that
// gives back to the operand ring are a whole stage for the BN = 256 pairs.
// Lane 0 owns the bulk groups: before the buffer is rewritten it waits until the previous store has READ it.  Rows beyond M are
// written too (they exist: the maps cover the workspace's full chunk) but stay out of the GroupNorm sums.
template <int BN, bool HALF>
__device__ __forceinline__ void epilogue_plain_tma(const TcGemm& p, EpiSmem<BN>& es, uint32_t taddr, int m, int n0, int c_begin,
                                                   int lane, int tile_m, int quarter, uint32_t buf, uint32_t bar_res,
                                                   uint32_t& res_phase) {
  constexpr int CPP = TcGeo<BN>::CPP, NC = BN / 32;
  constexpr int NH = HALF ? 2 : 1, UNITS = HALF ? 4 : 8;          // f32 boxes per chunk, 16-byte units per box row
  constexpr uint32_t BOX_BYTES = HALF ? 2048u : 4096u;
  const bool row_ok = m < p.M;
  const int row0 = tile_m * TC_BM + quarter * 32;
  const bool vf = (p.ld_out_f32 & 3) == 0, vb = (p.ld_out_bf16 & 7) == 0, vrf = (p.ld_res_f32 & 3) == 0;
  float2* gscr = &es.part[0][0];
  const uint32_t rowb = buf + (uint32_t)lane * 64u, swb = (uint32_t)((lane >> 1) & 3);
  const uint32_t rowf = HALF ? rowb : buf + (uint32_t)lane * 128u, swf = HALF ? swb : (uint32_t)(lane & 7);
  bool first_box = true;                 // its residual was requested by the caller
  const bool stats_early = HALF && p.gn_part && (p.epi_tma & 1) && !(p.epi_tma & 4) && !p.res_f32;
#pragma unroll 1
  for (int cc = 0; cc < CPP; ++cc) {
    const int c = c_begin + cc;
    const int nb = n0 + c * 32;
    if (nb >= p.N) continue;             // uniform across the warp (N is a multiple of 32 here)
    float v[32];
    tmem_ld_32x32(taddr + c * 32, v);
    add_smem32(v, es.bias + c * 32);
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (!(p.epi_tma & 4) && p.res_f32 && row_ok) add_f32x32(v, p.res_f32 + (long long)m * p.ld_res_f32 + nb, vrf, 32);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int col = nb + h * 16;       // (NH == 1: h = 0)
      if (p.epi_tma & 4) {
        if (!first_box && lane == 0) {
          bulk_wait_read0();
          mbar_arrive_expect_tx(bar_res, BOX_BYTES);
          tma_load_2d(buf, &p.epi_maps[2], bar_res, col, row0);
        }
        first_box = false;
        mbar_wait(bar_res, res_phase);
        res_phase ^= 1u;
#pragma unroll
        for (int j = 0; j < UNITS; ++j) {
          const float4 f = lds_128(rowf + (((uint32_t)j ^ swf) << 4));
          const int i = h * 16 + 4 * j;
          v[i] += f.x; v[i + 1] += f.y; v[i + 2] += f.z; v[i + 3] += f.w;
        }
      }
      if (p.epi_tma & 1) {
        if (!(p.epi_tma & 4) && lane == 0) bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < UNITS; ++j) {
          const int i = h * 16 + 4 * j;
          sts_128(rowf + (((uint32_t)j ^ swf) << 4), __float_as_uint(v[i]), __float_as_uint(v[i + 1]), __float_as_uint(v[i + 2]),
                  __float_as_uint(v[i + 3]));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.epi_maps[0], buf, col, row0);
          bulk_commit();
        }
      }
      // HALF without a residual: v is final, so the GroupNorm sums run here - under the first box's store, whose read the second box
      // has to wait for (the store queues behind the producer's operand loads in the TMA unit: ~1.5k cycles)
      if (HALF && h == 0 && stats_early) {
        float2* o = gscr + (quarter * NC + c) * 8;
        if (p.gn_cpg == 4) chunk_group_sums<8>(v, row_ok, o, lane);
        else if (p.gn_cpg == 8) chunk_group_sums<4>(v, row_ok, o, lane);
        else if (p.gn_cpg == 16) chunk_group_sums<2>(v, row_ok, o, lane);
        else chunk_group_sums<1>(v, row_ok, o, lane);
      }
    }
    if (!(p.epi_tma & 1) && p.out_f32 && row_ok) store_f32x32(p.out_f32 + (long long)m * p.ld_out_f32 + nb, v, vf, 32);
    if (p.epi_tma & 2) {
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts_128(rowb + (((uint32_t)j ^ swb) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&p.epi_maps[1], buf, nb, row0);
        bulk_commit();
      }
    } else if (p.out_bf16 && row_ok) {
      store_bf16x32(p.out_bf16 + (long long)m * p.ld_out_bf16 + nb, v, vb, 32);
    }
    if (p.gn_part && !stats_early) {      // uniform
      float2* o = gscr + (quarter * NC + c) * 8;
      if (p.gn_cpg == 4) chunk_group_sums<8>(v, row_ok, o, lane);
      else if (p.gn_cpg == 8) chunk_group_sums<4>(v, row_ok, o, lane);
      else if (p.gn_cpg == 16) chunk_group_sums<2>(v, row_ok, o, lane);
      else chunk_group_sums<1>(v, row_ok, o, lane);
    }
  }
  if (p.gn_part) gn_part_publish<BN>(p, es, n0, tile_m);
}

// Operands of the GN epilogue that do not depend on the accumulator - the FiLM scale/shift of this thread's sample and
// the residual row - are fetched into registers BEFORE the epilogue warps block on the accumulator barrier: the
// row-per-thread access pattern is uncoalesced (6k cycles per tile when issued after the main loop), but the epilogue
// warps are idle for the whole main loop, so issuing the loads up front hides them completely.  FiLM pairs are kept as
// half2 (scale, shift): O(1) values, 2^-11 relative rounding, far below the bf16 rounding of the activations.
template <int BN>
struct GnPrefetch {
  uint32_t film[TcGeo<BN>::CPP][32];   // half2(scale, shift) per column
};

template <int BN>
__device__ __forceinline__ void gn_prefetch(const TcGemm& p, const EpiSmem<BN>& es, int m, int n0, int c_begin,
                                            GnPrefetch<BN>& pf) {
  constexpr int NC = BN / 32, CPP = TcGeo<BN>::CPP;
  const int nchunks = min(NC, (p.N - n0) >> 5);
  const bool row_ok = m < p.M;
  const int mm = row_ok ? m : 0;
  const int b = mm / p.rows_per_item;
  // observation part of the FiLM embedding: half2 (scale, shift) pairs, quad-transposed [pair quad][sample] uint4 - the lanes of a
  // warp are consecutive rows = consecutive (or equal) samples, so one load instruction touches one or two 64-byte runs instead
  // of one line per lane; the time part is added in half2 (staged as half2 in shared memory, or per row for per-row timesteps).
  // One LDG.128 + one LDS.128 + four HADD2 per four columns: the prefetch runs while the MMA issuer needs the issue slots.
  const uint4* oq = reinterpret_cast<const uint4*>(p.otab_q);
  const float* trow = (p.film && p.step.rows) ? p.ttab + (long long)step_of(p.step, mm) * p.ld_ttab + p.film_off : nullptr;
  const bool film = p.film && !(p.epi_skip & 2);
#pragma unroll
  for (int cc = 0; cc < CPP; ++cc) {
    const int c = c_begin + cc;
    const int nb = n0 + c * 32;
    if (film && c < nchunks) {
      const uint4* o4 = oq + (long long)(((p.film_off >> 1) + nb) >> 2) * p.otab_B + b;
      const uint4* t4 = reinterpret_cast<const uint4*>(es.film_t + c * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 o = __ldg(o4 + (long long)j * p.otab_B);
        uint4 t;
        if (trow) {
          const float4 t1 = __ldg(reinterpret_cast<const float4*>(trow + nb) + j);
          const float4 t2 = __ldg(reinterpret_cast<const float4*>(trow + p.film_c + nb) + j);
          __half2 h0 = __floats2half2_rn(t1.x, t2.x), h1 = __floats2half2_rn(t1.y, t2.y);
          __half2 h2 = __floats2half2_rn(t1.z, t2.z), h3 = __floats2half2_rn(t1.w, t2.w);
          t = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
                         *reinterpret_cast<uint32_t*>(&h3));
        } else {
          t = t4[j];
        }
        const uint32_t ov[4] = {o.x, o.y, o.z, o.w}, tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&ov[q]), *reinterpret_cast<const __half2*>(&tv[q]));
          pf.film[cc][4 * j + q] = *reinterpret_cast<const uint32_t*>(&sum);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) pf.film[cc][i] = 0x00003c00u;     // half2(1, 0): identity
    }
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_gn(const TcGemm& p, EpiSmem<BN>& es, uint32_t taddr, int m, int n0, int c_begin,
                                            int row, int lane, const GnPrefetch<BN>& pf) {
  constexpr int NC = BN / 32, CPP = TcGeo<BN>::CPP;
  const int T = p.rows_per_item;
  const int cpg = p.group_width >> 5;                    // chunks per group: 1, 2 or 4
  const int nchunks = min(NC, (p.N - n0) >> 5);          // N and the group widths are multiples of 32 here
  const bool row_ok = m < p.M;
  // one TMEM pass: acc + bias stays in registers; (sum, sum sq) per 32-column chunk goes to shared memory
  float v[CPP][32];
#pragma unroll
  for (int cc = 0; cc < CPP; ++cc) {
    const int c = c_begin + cc;
    float s = 0.f, ss = 0.f;
    if (c < nchunks) {
      load_acc_chunk<BN>(p, taddr, c, lane, v[cc]);
      add_smem32(v[cc], es.bias + c * 32);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s += v[cc][i];
        ss = fmaf(v[cc][i], v[cc][i], ss);
      }
    }
    es.part[row][c] = make_float2(s, ss);
  }
  const bool dbg_t = p.dbg_stage && threadIdx.x == 64 && blockIdx.x == 0 && blockIdx.y == 0;
  if (dbg_t) p.dbg_stage[24] = clock64();          // accumulators read, partial statistics written
  epi_bar<BN>();
  if (dbg_t) p.dbg_stage[25] = clock64();          // statistics barrier passed
  // The residual row is fetched here rather than with the FiLM prefetch: its ~1k cycles of L2 latency hide behind the
  // normalise + Mish phase below, and its 16 registers are not live while the tap accumulators are being recombined
  // (v + the second accumulator + FiLM already fill the 96-register budget there).
  uint4 res[CPP][4];
  {
    const bool have_res = p.res_bf16 && !p.use_aux && row_ok && !(p.epi_skip & 4);
#pragma unroll
    for (int cc = 0; cc < CPP; ++cc) {
      const int c = c_begin + cc;
      if (have_res && c < nchunks) {
        const __nv_bfloat16* rsrc = p.res_bf16 + (long long)m * p.ld_res_bf16 + n0 + c * 32;
        if (aligned32(rsrc)) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const U32x8 u = ld_global_cg_256(rsrc + 16 * j);
            res[cc][2 * j] = make_uint4(u.r[0], u.r[1], u.r[2], u.r[3]);
            res[cc][2 * j + 1] = make_uint4(u.r[4], u.r[5], u.r[6], u.r[7]);
          }
        } else {
          const uint4* rp = reinterpret_cast<const uint4*>(rsrc);
#pragma unroll
          for (int j = 0; j < 4; ++j) res[cc][j] = __ldcg(rp + j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) res[cc][j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  // group statistics: chunks of the group (from smem) x the T rows of the sample (adjacent lanes)
  float mean[CPP], rstd[CPP];
  const float inv_cnt = 1.f / (float)(T * p.group_width);
#pragma unroll
  for (int cc = 0; cc < CPP; ++cc) {
    const int c = c_begin + cc;
    const int g0 = (c / cpg) * cpg;
    float s = 0.f, ss = 0.f;
    for (int c2 = g0; c2 < g0 + cpg; ++c2) {
      float2 q = es.part[row][c2];
      s += q.x;
      ss += q.y;
    }
    for (int off = 1; off < T; off <<= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    const float mu = s * inv_cnt;
    mean[cc] = mu;
    rstd[cc] = rsqrtf(fmaxf(ss * inv_cnt - mu * mu, 0.f) + p.eps);
  }
  // normalise -> activation -> FiLM -> residual -> store
#pragma unroll
  for (int cc = 0; cc < CPP; ++cc) {
    const int c = c_begin + cc;
    if (c >= nchunks) continue;
    const int nb = n0 + c * 32;
    const float4* g4 = reinterpret_cast<const float4*>(es.gamma + c * 32);
    const float4* be4 = reinterpret_cast<const float4*>(es.beta + c * 32);
    const float rs = rstd[cc], nmr = -mean[cc] * rstd[cc];
    float(&w)[32] = v[cc];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 gg = g4[j], be = be4[j];
      w[4 * j + 0] = fmaf(fmaf(w[4 * j + 0], rs, nmr), gg.x, be.x);
      w[4 * j + 1] = fmaf(fmaf(w[4 * j + 1], rs, nmr), gg.y, be.y);
      w[4 * j + 2] = fmaf(fmaf(w[4 * j + 2], rs, nmr), gg.z, be.z);
      w[4 * j + 3] = fmaf(fmaf(w[4 * j + 3], rs, nmr), gg.w, be.w);
    }
    if (p.epi_skip & 8) {
    } else if (p.gn_act == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) w[i] = swish_fast(w[i]);
    } else if (p.gn_act == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) w[i] = mish_fast(w[i]);
    }
    if (dbg_t) p.dbg_stage[26] = clock64();        // normalised + activation
    if (p.film) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pf.film[cc][i]));
        w[i] = fmaf(f.x, w[i], f.y);
      }
    }
    if (p.use_aux) {
      float r[32];
      tmem_ld_32x32(taddr + p.n_acc * BN + c * 32, r);
      add_smem32(r, es.bias2 + c * 32);
#pragma unroll
      for (int i = 0; i < 32; ++i) w[i] += r[i];
    } else if (p.res_bf16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = res[cc][j];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(h[q]);
          w[8 * j + 2 * q] += f.x;
          w[8 * j + 2 * q + 1] += f.y;
        }
      }
    }
    if (dbg_t) p.dbg_stage[27] = clock64();        // FiLM + residual applied
    if (row_ok && !((p.epi_skip & 1) && w[0] != 12345.678f)) {
      if (p.out_bf16) store_bf16x32(p.out_bf16 + (long long)m * p.ld_out_bf16 + nb, w, true, 32);
      if (p.out_f32) store_f32x32(p.out_f32 + (long long)m * p.ld_out_f32 + nb, w, true, 32);
    }
    if (dbg_t) p.dbg_stage[28] = clock64();        // stores issued
  }
}

// DDPM / DDIM update fused behind the score net's last GEMM.  Phase 1: every epilogue thread drops its row of
// eps = acc + bias into a padded shared-memory tile (the pipeline buffers, idle by now).  Phase 2: the tile is
// walked row-major, one thread per group of 4 consecutive columns, so x, the injected noise and the bf16 copy of x
// are accessed coalesced and one Philox call serves four elements (row-structured quads, see DdpmCall).
template <int BN>
__device__ __forceinline__ void epilogue_ddpm(const TcGemm& p, EpiSmem<BN>& es, uint32_t taddr, float* tile, int tile_m,
                                              int n0, int c_begin, int row, int et, int lane, int t_override = -1, long long* dbg = nullptr,
                                              int n_tail = 0) {
  constexpr int CPP = TcGeo<BN>::CPP;
  const int TS = BN + n_tail + 1;            // odd row pitch of the transposition tile (n_tail: widened last N tile, 0 or 16)
#pragma unroll 1
  for (int cc = 0; cc < CPP; ++cc) {
    const int c = c_begin + cc;
    if (n0 + c * 32 >= p.N) continue;          // uniform
    float v[32];
    load_acc_chunk<BN>(p, taddr, c, lane, v);
    add_smem32(v, es.bias + c * 32);
    float* trow = tile + row * TS + c * 32;
#pragma unroll
    for (int i = 0; i < 32; ++i) trow[i] = v[i];
  }
  if (n_tail > 0 && c_begin == 0) {            // uniform per warp: the first column slice also drains the 16 extra columns
    float v[32];
    tmem_ld_32x32(taddr + BN, v);              // columns [BN, BN + 32) lie inside the power-of-two TMEM allocation
    float* trow = tile + row * TS + BN;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + BN + i;
      trow[i] = v[i] + ((n < p.N && p.bias) ? __ldg(p.bias + n) : 0.f);
    }
  }
  epi_bar<BN>();
  const int t = t_override >= 0 ? t_override : step_of(p.step, 0);
  const float* cf = p.coef + t * 8;
  const float inv_sa = cf[0], s1a = cf[1], c0 = cf[2], ct = cf[3], sigma = cf[4], sap = cf[5], s1ap = cf[6];
  const DdpmCall call = p.call_dev ? *p.call_dev : p.call;
  const float* noise = call.noise ? call.noise + (long long)(call.n_steps - 1 - t) * call.noise_step_stride : nullptr;
  const bool ddim = call.sampler == LDP_SAMPLER_DDIM;
  const bool add_noise = !ddim && t > 0;
  const int nv = min(BN + n_tail, p.N - n0);
  const int nq = (nv + 3) >> 2;                                  // column quads of this tile (n0 is a multiple of 4)
  const int rows_here = min(TC_BM, p.M - tile_m * TC_BM);
  const int total = rows_here * nq;
  const bool vb = p.out_bf16 != nullptr && (p.ld_out_bf16 & 3) == 0;
  // The loop below is issue-bound (Philox + Box-Muller + index arithmetic: ~250 instructions per quad, 9 quads per thread),
  // not load-bound - prefetching x ahead of the accumulator barrier changed nothing - so the index division is a multiply:
  // floor(idx / nq) = (idx * ceil(2^20 / nq)) >> 20 exactly for idx < 4608, nq <= 36.
  const uint32_t nq_inv = ((1u << 20) + (uint32_t)nq - 1u) / (uint32_t)nq;
#pragma unroll 1
  for (int idx = et; idx < total; idx += TcGeo<BN>::EPI_THREADS) {
    const int r = (int)(((uint32_t)idx * nq_inv) >> 20), g = idx - r * nq;
    const int m = tile_m * TC_BM + r;
    const int cb = g * 4;
    const int cnt = min(4, nv - cb);
    const long long e0 = (long long)m * p.N + n0 + cb;
    float* xr = p.x_io + (long long)m * p.ld_x + n0 + cb;
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (add_noise) {
      if (noise) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < cnt) z[i] = noise[e0 + i];
      } else {
        philox_normal4_rows(call.seed, call.stream_id, (uint32_t)t, (uint32_t)(call.row_offset + m), (uint32_t)((n0 + cb) >> 2), z);
      }
    }
    // all four x loads first: ld.global.cg is a strong access that the compiler keeps in program order with the stores
    // below, so loading inside the per-element loop serialises four L2 round trips per iteration
    float xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = i < cnt ? __ldcg(xr + i) : 0.f;
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i] = 0.f;
      if (i < cnt) {
        const float e = tile[r * TS + cb + i];
        const float x = xv[i];
        const float x0 = fminf(fmaxf((x - s1a * e) * inv_sa, -1.f), 1.f);
        float y;
        if (ddim) {
          y = sap * x0 + s1ap * e;
        } else {
          y = c0 * x0 + ct * x;
          if (add_noise) y = fmaf(sigma, z[i], y);
        }
        o[i] = y;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < cnt) xr[i] = o[i];
    if (p.out_bf16) {
      __nv_bfloat16* ob = p.out_bf16 + (long long)m * p.ld_out_bf16 + n0 + cb;
      if (vb && cnt == 4) {
        uint2 u;
        u.x = pack_bf16x2(o[0], o[1]);
        u.y = pack_bf16x2(o[2], o[3]);
        *reinterpret_cast<uint2*>(ob) = u;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < cnt) ob[i] = __float2bfloat16(o[i]);
      }
    }
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_ln(const TcGemm& p, EpiSmem<BN>& es, uint32_t taddr, int m, int n0, int c_begin,
                                            int row, int part, int lane) {
  // requires N == BN: the whole feature row lives in this tile, split over the two warps of the lane quarter
  constexpr int CPP = TcGeo<BN>::CPP;
  const bool row_ok = m < p.M;
  float s = 0.f, ss = 0.f;
  float* hrow = p.out_f32 + (long long)(row_ok ? m : 0) * p.ld_out_f32;
  const bool vf = (p.ld_out_f32 & 3) == 0, vr = (p.ld_res_f32 & 3) == 0;
#pragma unroll 1
  for (int cc = 0; cc < CPP; ++cc) {
    const int c = c_begin + cc;
    const int nb = n0 + c * 32;
    float v[32];
    load_acc_chunk<BN>(p, taddr, c, lane, v);
    add_smem32(v, es.bias + c * 32);
    if (p.res_f32 && row_ok) add_f32x32(v, p.res_f32 + (long long)m * p.ld_res_f32 + nb, vr, 32);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      s += v[i];
      ss = fmaf(v[i], v[i], ss);
    }
    if (row_ok) store_f32x32(hrow + nb, v, vf, 32);
  }
  es.part[row][part] = make_float2(s, ss);
  epi_bar<BN>();
  float ts = 0.f, tss = 0.f;
#pragma unroll
  for (int q = 0; q < TcGeo<BN>::PARTS; ++q) {
    const float2 pq = es.part[row][q];
    ts += pq.x;
    tss += pq.y;
  }
  const float mu = ts / (float)BN;
  const float rs = rsqrtf(fmaxf(tss / (float)BN - mu * mu, 0.f) + p.eps);
  if (!row_ok || !p.out_bf16) return;
#pragma unroll 1
  for (int cc = 0; cc < CPP; ++cc) {
    const int c = c_begin + cc;
    const int nb = n0 + c * 32;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    add_f32x32(v, hrow + nb, vf, 32);                     // h written above by this same thread
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaf((v[i] - mu) * rs, es.gamma[c * 32 + i], es.beta[c * 32 + i]);
    }
    store_bf16x32(p.out_bf16 + (long long)m * p.ld_out_bf16 + nb, v, (p.ld_out_bf16 & 7) == 0, 32);
  }
}

}  // namespace ldp
