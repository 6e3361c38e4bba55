// The whole IDM reverse-diffusion loop (reference agent/ldp_agent.py:492-503 around networks/mlp_diffusion_nets.py:32-68)
// as ONE persistent launch: n_steps x { h = a Wa + spre + ctab[t];  3 x { h += relu(LN(h) W1 + b1) W2 + b2 };  eps = relu(h) Wout + b;
// a <- DDPM/DDIM step(eps, t, a) }.
//
// Transition rows are independent, so a CTA owns 128 rows for the whole loop and no CTA ever waits for another: there is
// no kernel boundary, no grid barrier and no activation traffic to global memory inside the loop - the per-layer path
// (idm.cu: 8 launches per step, ~12 us each, all fixed cost) spent 100 us per step on 9 us of tensor work.
//
//   * residual stream h: fp32, TMEM columns [0, 256) - it IS the accumulator of the down-projection (h += u W2 is the MMA's own
//     accumulate), initialised per step with tcgen05.st;
//   * the 1024-wide hidden layer never exists as a whole: it is produced in 8 chunks of 128 columns (TMEM columns [256, 512), two
//     buffers), each drained by the epilogue warps (bias, ReLU, bf16) into a 128B-swizzled K-major shared-memory tile that is
//     the A operand of the down-projection's next two K blocks - chunk j's drain overlaps the MMAs of chunks j-1 / j+1;
//   * LayerNorm (two passes over TMEM, statistics exchanged between the two column halves of a row through shared memory)
//     writes the next up-projection's A operand (4 K blocks) in the same swizzled layout;
//   * weights (3 MB bf16 per step, L2 resident) stream through a 3-slot TMA ring of 32 KB tiles in exactly the order the MMA
//     issuer consumes them;
//   * the output Dense is one N = 16 MMA group; each row's thread pair keeps its action vector in registers across all steps
//     and applies the scheduler step (Philox noise keyed by the global row, identical to the per-layer path's draws).
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..17 = epilogue: warp w reads TMEM lane quarter w % 4 and column
// quarter (w - 2) / 4 of whatever is being drained, so a thread holds 64 columns of its row of h (LayerNorm runs on registers
// after ONE TMEM pass; the four column quarters of a row exchange their sums through shared memory) and 32 columns of a hidden
// chunk.  (8 epilogue warps with 128 / 64 columns per thread ran the serial phases - input, 3 x LayerNorm - in 13k cycles each:
// 2 warps per scheduler cannot hide the TMEM / L2 latencies of a dependent chain; measured with the in-kernel wait clocks.)
#include <cstddef>

#include "net_common.h"
#include "tc_epilogue.cuh"

namespace ldp {

constexpr int IL_THREADS = 576;
constexpr int IL_EPI = 512;
constexpr uint32_t IL_HN = 0;                    // 4 K blocks x 16 KB: LayerNorm output, A operand of the up-projection
constexpr uint32_t IL_UB = 65536;                // 2 buffers x (2 K blocks x 16 KB): hidden chunk, A operand of the down-projection
constexpr uint32_t IL_RING = 131072;             // 3 slots x 32 KB weight tiles
constexpr uint32_t IL_SLOT = 32768;
constexpr int IL_NSLOT = 3;
constexpr uint32_t IL_SMEM = IL_RING + IL_NSLOT * IL_SLOT;     // 224 KB
constexpr uint32_t IL_COL_H = 0, IL_COL_U = 256;  // TMEM columns

// Per-column constants of the network.  With 225 KB of shared memory the L1 data cache is a few KB, so a warp-uniform __ldg
// of a bias / LayerNorm / Wa vector is an L2 round trip (measured: every epilogue phase was 3-4x slower than its instruction
// count); the constant cache serves warp-uniform addresses at full rate.  Uploaded (device to device) when the handle changes.
struct IdmConst {
  float b1[4][1024];
  float bsum[4][256];
  float ln_g[4][256];
  float ln_b[4][256];
  float wa[16][256];
  float bout[16];
};
__constant__ IdmConst c_idm;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void il_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(IL_EPI) : "memory"); }

// 32 consecutive K elements (bf16) of row `row` at K offset `k0` (multiple of 32) of a [128][64 * nkb] K-major operand made of
// 16 KB K blocks in the canonical 128B-swizzle layout: 16-byte chunk index ^= row % 8.
__device__ __forceinline__ void st_operand32(uint32_t base, int row, int k0, const float (&v)[32]) {
  const uint32_t rowb = base + (uint32_t)(k0 >> 6) * 16384u + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
  const uint32_t c0 = (uint32_t)((k0 & 63) >> 3);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t addr = rowb + (((c0 + q) ^ (uint32_t)(row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[8 * q], v[8 * q + 1])),
                 "r"(pack_bf16x2(v[8 * q + 2], v[8 * q + 3])), "r"(pack_bf16x2(v[8 * q + 4], v[8 * q + 5])),
                 "r"(pack_bf16x2(v[8 * q + 6], v[8 * q + 7])) : "memory");
  }
}

__device__ __forceinline__ void add_vec32(float (&v)[32], const float* __restrict__ g) {
  const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 f = __ldg(g4 + j);
    v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
  }
}

template <bool DBG>
__global__ void __launch_bounds__(IL_THREADS, 1) idm_loop_kernel(const __grid_constant__ IdmLoop p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full[IL_NSLOT], w_empty[IL_NSLOT];
  __shared__ __align__(8) uint64_t u_full[2], u_empty[2], ub_full[2], ub_empty[2];
  __shared__ __align__(8) uint64_t h_full, hn_full, out_full;
  __shared__ uint32_t tmem_holder;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.n_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < IL_NSLOT; ++s) { mbar_init(smem_u32(&w_full[s]), 1); mbar_init(smem_u32(&w_empty[s]), 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(smem_u32(&u_full[x]), 1);
      mbar_init(smem_u32(&u_empty[x]), IL_EPI / 32);
      mbar_init(smem_u32(&ub_full[x]), IL_EPI / 32);
      mbar_init(smem_u32(&ub_empty[x]), 1);
    }
    mbar_init(smem_u32(&h_full), 1);
    mbar_init(smem_u32(&hn_full), IL_EPI / 32);
    mbar_init(smem_u32(&out_full), 1);
    fence_mbar_init();
    for (int b = 0; b < nb; ++b) { tma_prefetch_desc(&p.map_w1[b]); tma_prefetch_desc(&p.map_w2[b]); }
    tma_prefetch_desc(&p.map_wout);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer: weights, in the order the MMA issuer consumes them =====================
    if (elect_one()) {
      uint32_t slot = 0, round = 0;
      auto acquire = [&](uint32_t bytes) -> uint32_t {
        mbar_wait(smem_u32(&w_empty[slot]), (round & 1u) ^ 1u);
        mbar_arrive_expect_tx(smem_u32(&w_full[slot]), bytes);
        return sbase + IL_RING + slot * IL_SLOT;
      };
      auto advance = [&]() { if (++slot == IL_NSLOT) { slot = 0; ++round; } };
      for (int it = 0; it < p.n_steps; ++it) {
        for (int b = 0; b < nb; ++b) {
          for (int j = 0; j <= 8; ++j) {
            if (j < 8) {                              // up-projection chunk j: W1 rows [128 j, +128), 4 K blocks in two slots
              for (int s = 0; s < 2; ++s) {
                const uint32_t dst = acquire(IL_SLOT);
                const uint32_t bar = smem_u32(&w_full[slot]);
                tma_load_2d(dst, &p.map_w1[b], bar, (2 * s) * 64, 128 * j);
                tma_load_2d(dst + 16384u, &p.map_w1[b], bar, (2 * s + 1) * 64, 128 * j);
                advance();
              }
            }
            if (j >= 1) {                             // down-projection chunk j-1: W2 K blocks 2(j-1), 2(j-1)+1, all 256 rows
              for (int s = 0; s < 2; ++s) {
                const uint32_t dst = acquire(IL_SLOT);
                tma_load_2d(dst, &p.map_w2[b], smem_u32(&w_full[slot]), (2 * (j - 1) + s) * 64, 0);
                advance();
              }
            }
          }
        }
        const uint32_t dst = acquire(8192u);          // output Dense: 16 rows x 4 K blocks
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(dst + (uint32_t)kb * 2048u, &p.map_wout, smem_u32(&w_full[slot]), kb * 64, 0);
        advance();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t IDESC_UP = umma_idesc_bf16(128, 128), IDESC_DOWN = umma_idesc_bf16(128, 256), IDESC_OUT = umma_idesc_bf16(128, 16);
      uint32_t slot = 0, round = 0;
      uint32_t n_u[2] = {0, 0}, n_ub[2] = {0, 0}, n_hn = 0;
      long long wt[4] = {0, 0, 0, 0};                 // diagnostics: cycles waiting for {weights, U drained, operand chunk, LayerNorm}
      const bool dbg = DBG && p.dbg != nullptr && blockIdx.x == 0;
      const long long t_begin = clock64();
#define IL_TIMED(idx, stmt) do { if (dbg) { const long long _t = clock64(); stmt; wt[idx] += clock64() - _t; } else { stmt; } } while (0)
      auto wait_w = [&]() -> uint32_t {
        IL_TIMED(0, mbar_wait(smem_u32(&w_full[slot]), round & 1u));
        tc_fence_after();
        return sbase + IL_RING + slot * IL_SLOT;
      };
      auto release_w = [&]() {
        umma_commit(smem_u32(&w_empty[slot]));
        if (++slot == IL_NSLOT) { slot = 0; ++round; }
      };
      auto up = [&](int j) {
        const int x = j & 1;
        if (n_u[x] > 0) { IL_TIMED(1, mbar_wait(smem_u32(&u_empty[x]), (n_u[x] - 1u) & 1u)); tc_fence_after(); }
        ++n_u[x];
        const uint32_t d = tmem + IL_COL_U + (uint32_t)x * 128u;
        for (int s = 0; s < 2; ++s) {
          const uint32_t w = wait_w();
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t da = umma_desc_sw128(sbase + IL_HN + (uint32_t)(2 * s + kk) * 16384u);
            const uint64_t db = umma_desc_sw128(w + (uint32_t)kk * 16384u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(d, da + 2 * k, db + 2 * k, IDESC_UP, (s | kk | k) ? 1u : 0u);
          }
          release_w();
        }
        umma_commit(smem_u32(&u_full[x]));
      };
      auto down = [&](int j) {
        const int x = j & 1;
        IL_TIMED(2, mbar_wait(smem_u32(&ub_full[x]), n_ub[x] & 1u));
        ++n_ub[x];
        tc_fence_after();
        for (int s = 0; s < 2; ++s) {
          const uint32_t w = wait_w();
          const uint64_t da = umma_desc_sw128(sbase + IL_UB + (uint32_t)x * 32768u + (uint32_t)s * 16384u);
          const uint64_t db = umma_desc_sw128(w);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem + IL_COL_H, da + 2 * k, db + 2 * k, IDESC_DOWN, 1u);
          release_w();
        }
        umma_commit(smem_u32(&ub_empty[x]));
      };
      for (int it = 0; it < p.n_steps; ++it) {
        for (int b = 0; b < nb; ++b) {
          IL_TIMED(3, mbar_wait(smem_u32(&hn_full), n_hn & 1u));      // LayerNorm output of this block is in shared memory
          ++n_hn;
          tc_fence_after();
          for (int j = 0; j <= 8; ++j) {
            if (j < 8) up(j);
            if (j >= 1) down(j - 1);
          }
          umma_commit(smem_u32(&h_full));
        }
        IL_TIMED(3, mbar_wait(smem_u32(&hn_full), n_hn & 1u));
        ++n_hn;
        tc_fence_after();
        const uint32_t w = wait_w();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const uint64_t da = umma_desc_sw128(sbase + IL_HN + (uint32_t)kb * 16384u);
          const uint64_t db = umma_desc_sw128(w + (uint32_t)kb * 2048u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem + IL_COL_U, da + 2 * k, db + 2 * k, IDESC_OUT, (kb | k) ? 1u : 0u);
        }
        release_w();
        umma_commit(smem_u32(&out_full));
      }
      if (dbg) {
        for (int i = 0; i < 4; ++i) p.dbg[i] = wt[i];
        p.dbg[4] = clock64() - t_begin;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2, quarter = warp & 3, part = ew >> 2;
    const int row = quarter * 32 + lane;
    const long long grow = (long long)blockIdx.x * 128 + row;
    const bool row_ok = grow < p.N;
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    const int A = p.A;
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (row_ok && j < A) ? p.a_state[grow * A + j] : 0.f;
    uint32_t n_uf[2] = {0, 0}, n_ube[2] = {0, 0}, n_h = 0, n_out = 0;
    // diagnostics (DBG build only): cycles of thread 64 in {input stage, LayerNorm, chunk drain work, wait chunk acc, wait block, output+step}
    long long ph[6] = {0, 0, 0, 0, 0, 0};
    const bool edbg = DBG && p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
#define IL_PH(idx, t0) do { if (DBG && edbg) { const long long _n = clock64(); ph[idx] += _n - (t0); (t0) = _n; } } while (0)
    float2* scratch = reinterpret_cast<float2*>(smem_raw + (sbase - smem_u32(smem_raw)) + IL_UB + 32768u);   // [128 rows][4 column quarters], in the second operand buffer (the transposition tiles of the input stage end at 72 KB)
    const int hcol = part * 64;                        // this thread's 64 columns of h

    // this thread's 64 columns of h (TMEM, + bias) -> LayerNorm / ReLU -> bf16 A operand of the next GEMM.  Two passes over TMEM,
    // 32 columns live at a time: with 18 warps the register budget is 96 per thread and the action vector stays resident.
    auto norm_to_operand = [&](int bias_blk, int ln_blk, bool relu_only) {      // bias_blk < 0: no bias
      float rstd = 1.f, nm = 0.f;
      if (!relu_only) {
        float s = 0.f, ss = 0.f;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld_32x32(tlane + IL_COL_H + (uint32_t)(hcol + c * 32), v);
          if (bias_blk >= 0) {
            const float4* b4 = reinterpret_cast<const float4*>(&c_idm.bsum[bias_blk][hcol + c * 32]);      // LDC.128: 4 values per constant load
#pragma unroll
            for (int q = 0; q < 8; ++q) { const float4 f = b4[q]; v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w; }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) { s += v[i]; ss = fmaf(v[i], v[i], ss); }
        }
        scratch[row * 4 + part] = make_float2(s, ss);
        il_epi_bar();
        s = 0.f; ss = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {                  // same order in all four threads of the row: identical statistics
          const float2 o = scratch[row * 4 + q];
          s += o.x; ss += o.y;
        }
        const float mean = s * (1.f / 256.f);
        rstd = rsqrtf(fmaxf(ss * (1.f / 256.f) - mean * mean, 0.f) + 1e-6f);
        nm = -mean * rstd;
      }
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float v[32];
        const int col = hcol + c * 32;
        tmem_ld_32x32(tlane + IL_COL_H + (uint32_t)col, v);
        if (bias_blk >= 0) {
          const float4* b4 = reinterpret_cast<const float4*>(&c_idm.bsum[bias_blk][col]);
#pragma unroll
          for (int q = 0; q < 8; ++q) { const float4 f = b4[q]; v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w; }
        }
        if (relu_only) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        } else {
          const float4* g4 = reinterpret_cast<const float4*>(&c_idm.ln_g[ln_blk][col]);
          const float4* e4 = reinterpret_cast<const float4*>(&c_idm.ln_b[ln_blk][col]);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 g = g4[q], be = e4[q];
            v[4 * q] = fmaf(fmaf(v[4 * q], rstd, nm), g.x, be.x);
            v[4 * q + 1] = fmaf(fmaf(v[4 * q + 1], rstd, nm), g.y, be.y);
            v[4 * q + 2] = fmaf(fmaf(v[4 * q + 2], rstd, nm), g.z, be.z);
            v[4 * q + 3] = fmaf(fmaf(v[4 * q + 3], rstd, nm), g.w, be.w);
          }
        }
        st_operand32(sbase + IL_HN, row, col, v);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&hn_full));       // one arrival per warp: 512 per-thread arrivals on one mbarrier cost ~1.5k cycles
    };

    const DdpmCall call = p.call;
    const bool ddim = call.sampler == LDP_SAMPLER_DDIM;
    for (int it = 0; it < p.n_steps; ++it) {
      const int t = p.t_first - it;
      long long tc = DBG ? clock64() : 0;
      // ---- input: h = a Wa + spre + ctab[t]  (fp32, written to TMEM), then LayerNorm of block 0 ----
      {
        const float* ct = p.ctab_h + (long long)t * 256 + hcol;
        // spre rows are read coalesced (8 lanes x 16 B = one 128-byte line per row, 4 rows per instruction) and transposed through
        // a per-warp shared-memory tile (pitch 36 floats: conflict-free both ways) in the operand buffers, which are idle here:
        // one thread per row reading its own 128 bytes costs a wavefront per lane per instruction (8k wavefronts, 20k cycles per step).
        float* tile = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw))) + ew * (32 * 36);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {                  // 32 columns at a time (register budget: 96 with 18 warps)
          float v[32];
          {
            const int lr = lane >> 3, lc = (lane & 7) * 4;
            const long long r0 = (long long)blockIdx.x * 128 + quarter * 32;
            float4 f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const long long gr = r0 + k * 4 + lr;
              f[k] = __ldg(reinterpret_cast<const float4*>(p.spre + (gr < p.N ? gr : 0) * 256 + hcol + c * 32 + lc));
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(tile + (k * 4 + lr) * 36 + lc) = f[k];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g = *reinterpret_cast<const float4*>(tile + lane * 36 + 4 * j);
              v[4 * j] = g.x; v[4 * j + 1] = g.y; v[4 * j + 2] = g.z; v[4 * j + 3] = g.w;
            }
          }
          add_vec32(v, ct + c * 32);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < A) {
              const float aj = a[j];
              const float4* w4 = reinterpret_cast<const float4*>(&c_idm.wa[j][hcol + c * 32]);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 f = w4[q];
                v[4 * q] = fmaf(aj, f.x, v[4 * q]); v[4 * q + 1] = fmaf(aj, f.y, v[4 * q + 1]);
                v[4 * q + 2] = fmaf(aj, f.z, v[4 * q + 2]); v[4 * q + 3] = fmaf(aj, f.w, v[4 * q + 3]);
              }
            }
          }
          tmem_st_32x32(tlane + IL_COL_H + (uint32_t)(hcol + c * 32), v);
        }
        tmem_st_wait();
        IL_PH(0, tc);
        norm_to_operand(-1, 0, false);
        IL_PH(1, tc);
      }
      // ---- residual blocks ----
      for (int b = 0; b < nb; ++b) {
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          const int x = j & 1;
          mbar_wait(smem_u32(&u_full[x]), n_uf[x] & 1u);
          ++n_uf[x];
          tc_fence_after();
          if (n_ube[x] > 0) mbar_wait(smem_u32(&ub_empty[x]), (n_ube[x] - 1u) & 1u);   // the MMAs that read this buffer have retired
          ++n_ube[x];
          IL_PH(3, tc);
          {
            float v[32];
            const int col = part * 32;                             // column inside the 128-wide chunk
            tmem_ld_32x32(tlane + IL_COL_U + (uint32_t)(x * 128 + col), v);
            const float4* b4 = reinterpret_cast<const float4*>(&c_idm.b1[b][j * 128 + col]);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 f = b4[q];
              v[4 * q] = fmaxf(v[4 * q] + f.x, 0.f); v[4 * q + 1] = fmaxf(v[4 * q + 1] + f.y, 0.f);
              v[4 * q + 2] = fmaxf(v[4 * q + 2] + f.z, 0.f); v[4 * q + 3] = fmaxf(v[4 * q + 3] + f.w, 0.f);
            }
            st_operand32(sbase + IL_UB + (uint32_t)x * 32768u, row, col, v);
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(smem_u32(&u_empty[x]));                    // accumulator chunk drained
            mbar_arrive(smem_u32(&ub_full[x]));                    // operand chunk written
          }
          IL_PH(2, tc);
        }
        mbar_wait(smem_u32(&h_full), n_h & 1u);
        ++n_h;
        tc_fence_after();
        IL_PH(4, tc);
        if (b + 1 < nb) norm_to_operand(b, b + 1, false);
        else norm_to_operand(b, 0, true);                         // after the last block: activations = relu, no norm
        IL_PH(1, tc);
      }
      // ---- output Dense + scheduler step on this row's action vector ----
      mbar_wait(smem_u32(&out_full), n_out & 1u);
      ++n_out;
      tc_fence_after();
      {
        float e[16];
        tmem_ld_32x16(tlane + IL_COL_U, e);
        tc_fence_before();
        const float* cf = p.coef + t * 8;
        const float inv_sa = cf[0], s1a = cf[1], c0 = cf[2], ctc = cf[3], sigma = cf[4], sap = cf[5], s1ap = cf[6];
        const bool add_noise = !ddim && t > 0;
        const float* noise = call.noise ? call.noise + (long long)(call.n_steps - 1 - t) * call.noise_step_stride + grow * A : nullptr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q * 4 < A) {
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            if (add_noise) {
              if (noise) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (q * 4 + i < A && row_ok) z[i] = noise[q * 4 + i];
              } else {
                philox_normal4_rows(call.seed, call.stream_id, (uint32_t)t, (uint32_t)(call.row_offset + grow), (uint32_t)q, z);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = q * 4 + i;
              if (j < A) {
                const float eps = e[j] + c_idm.bout[j];
                const float x = a[j];
                const float x0 = fminf(fmaxf((x - s1a * eps) * inv_sa, -1.f), 1.f);
                float y;
                if (ddim) {
                  y = sap * x0 + s1ap * eps;
                } else {
                  y = c0 * x0 + ctc * x;
                  if (add_noise) y = fmaf(sigma, z[i], y);
                }
                a[j] = y;
              }
            }
          }
        }
      }
      IL_PH(5, tc);
    }
    if (DBG && edbg) {
      for (int i = 0; i < 6; ++i) p.dbg[8 + i] = ph[i];
    }
    if (part == 0 && row_ok) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < A) p.a_state[grow * A + j] = a[j];
    }

  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int launch_idm_loop(const IdmLoop& p, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    LDP_CUDA_OK(cudaFuncSetAttribute(idm_loop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IL_SMEM + 1024));
    LDP_CUDA_OK(cudaFuncSetAttribute(idm_loop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IL_SMEM + 1024));
    attr_done = true;
  }
  LDP_CHECK(p.N > 0 && p.A >= 1 && p.A <= 16 && p.n_blocks >= 1 && p.n_blocks <= 4 && p.n_steps >= 1, LDP_ERR_UNSUPPORTED,
            "idm loop kernel: unsupported shape");
  {
    // constants of this handle -> constant memory (stream-ordered device-to-device copies; skipped while the handle stays the same)
    static const void* uploaded_for = nullptr;
    static unsigned long long uploaded_gen = 0;
    if (uploaded_for != (const void*)p.wa || uploaded_gen != p.const_gen) {
      const size_t H = 256;
      for (int b = 0; b < p.n_blocks; ++b) {
        LDP_CUDA_OK(cudaMemcpyToSymbolAsync(c_idm, p.b1[b], 1024 * 4, offsetof(IdmConst, b1) + (size_t)b * 1024 * 4, cudaMemcpyDeviceToDevice, s));
        LDP_CUDA_OK(cudaMemcpyToSymbolAsync(c_idm, p.bsum[b], H * 4, offsetof(IdmConst, bsum) + (size_t)b * H * 4, cudaMemcpyDeviceToDevice, s));
        LDP_CUDA_OK(cudaMemcpyToSymbolAsync(c_idm, p.ln_g[b], H * 4, offsetof(IdmConst, ln_g) + (size_t)b * H * 4, cudaMemcpyDeviceToDevice, s));
        LDP_CUDA_OK(cudaMemcpyToSymbolAsync(c_idm, p.ln_b[b], H * 4, offsetof(IdmConst, ln_b) + (size_t)b * H * 4, cudaMemcpyDeviceToDevice, s));
      }
      LDP_CUDA_OK(cudaMemcpyToSymbolAsync(c_idm, p.wa, (size_t)p.A * H * 4, offsetof(IdmConst, wa), cudaMemcpyDeviceToDevice, s));
      LDP_CUDA_OK(cudaMemcpyToSymbolAsync(c_idm, p.bout, (size_t)p.A * 4, offsetof(IdmConst, bout), cudaMemcpyDeviceToDevice, s));
      uploaded_for = p.wa;
      uploaded_gen = p.const_gen;
    }
  }
  if (p.dbg) idm_loop_kernel<true><<<ceil_div(p.N, 128), IL_THREADS, IL_SMEM + 1024, s>>>(p);
  else idm_loop_kernel<false><<<ceil_div(p.N, 128), IL_THREADS, IL_SMEM + 1024, s>>>(p);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(std::string("idm loop kernel launch failed: ") + cudaGetErrorString(e));
    return LDP_ERR_CUDA;
  }
  return LDP_OK;
}

}  // namespace ldp
