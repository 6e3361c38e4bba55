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
//   with the fp32 accumulator in TMEM; an mbarrier ring (6 stages at BN=128) overlaps TMA with MMA.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue
//   (two warps per 32-lane TMEM quarter, each taking half of the tile's columns; thread i owns output row
//   32*quarter+i).  Per-column vectors (bias, norm affine, FiLM time part) are staged in shared memory while
//   the main loop runs.
// * One template instantiation per (BN, epilogue): each kernel carries only its own epilogue, written as compact
//   loops over 32-column chunks (an earlier all-in-one, fully unrolled kernel was 16.6k SASS instructions and
//   spent 85 % of its cycles in instruction-fetch stalls; see profiles/).
// * Fused epilogues (all fp32 math on the accumulator, read with tcgen05.ld):
//     PLAIN  bias (+ReLU) (+residual) -> f32 and/or bf16
//     GN     bias -> GroupNorm over (rows of one sample x group channels) -> Mish|swish -> [FiLM] -> [+residual] -> bf16
//            (the tile owns whole samples and whole groups, so the statistics never leave the CTA; one TMEM
//            pass, the thread's 64 accumulator values stay in registers between statistics and normalisation)
//     DDPM   bias -> eps; x0 = clip((x - s eps)/a); x <- c0 x0 + ct x + sigma z   (the scheduler step of the
//            reverse-diffusion loop, reference agent/ldp_agent.py:470-471, fused around the score-net's last GEMM;
//            the tile is transposed through shared memory so that x / z / x_bf16 are accessed coalesced)
//     LN     h = acc + bias + residual -> f32;  LayerNorm(h) (or ReLU(h)) -> bf16   (IDM MLPResNet block)
// * Programmatic dependent launch: the prologue (barrier init, TMEM allocation, descriptor prefetch) runs before
//   griddepcontrol.wait, i.e. it overlaps the tail of the previous layer's kernel inside the captured graph.
#include "tc_epilogue.cuh"

namespace ldp {

// MMA with the 64-bit shared-memory descriptors given as (lo, hi) words: hi is a constant, lo advances by adds.
template <bool PAIR>
__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                          uint32_t accumulate) {
  if (PAIR) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

}  // namespace ldp

#include "tc_gemm_v1.cuh"

namespace ldp {

// ---- the kernel --------------------------------------------------------------------------------
template <int BN, int MODE, bool PAIR, bool PERSIST>
__global__ void __launch_bounds__(TcGeo<BN>::THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemm p) {
  extern __shared__ uint8_t smem_raw[];
  // PAIR: two CTAs of a cluster (adjacent M tiles, same N tile) run one cta_group::2 MMA (M = 256): each CTA loads its
  // own A tile but only its half of every W tile - the peer's half is read over the SM-to-SM path, not through this
  // SM's L2 ingest port, which is what bounds the single-CTA kernel.
  constexpr int B_BYTES = PAIR ? TcSmem<BN>::B_BYTES / 2 : TcSmem<BN>::B_BYTES;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int STAGES = p.num_stages;
  // A box of a stage: 128 rows, or 256 in the persistent kernel's shared-tap-row mode (TcGemm::a_rows) - a compile-time constant elsewhere
  const uint32_t a_bytes = PERSIST ? (uint32_t)p.a_rows * (TC_BK * 2) : (uint32_t)TC_A_BYTES;
  const uint32_t stage_bytes = a_bytes + (uint32_t)p.w_max * B_BYTES + (MODE == TC_EPI_DDPM ? (uint32_t)p.n_tail * (TC_BK * 2) : 0u);
  // DDPM: this CTA owns the widened last N tile (BN + n_tail columns)
  const bool tail_tile = MODE == TC_EPI_DDPM && p.n_tail > 0 && blockIdx.y + 1 == gridDim.y;
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2];              // accumulator buffer b complete (MMA -> epilogue)
  __shared__ __align__(8) uint64_t bar_tempty[2];             // accumulator buffer b drained (epilogue -> MMA)
  __shared__ __align__(8) uint64_t bar_res[16];               // TMA epilogue: residual box of epilogue warp w landed
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) TcRun runs_s[TC_MAX_RUNS];         // run-length stage table staged once per CTA
  __shared__ __align__(16) EpiSmem<BN> es;
  __shared__ long long ts[12];                                // phase timestamps / wait-cycle sums (diagnostics, only when p.dbg != nullptr)
  __shared__ long long tk[24];                                // arrival time of the first 24 stages at the MMA issuer
  if (p.dbg && threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) ts[i] = 0;
    for (int i = 0; i < 24; ++i) tk[i] = 0;
    ts[0] = clock64();
  }

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Tiles.  !PERSIST: one tile per CTA, (tile_m, tile_n) = (blockIdx.x, blockIdx.y) - the planner's layers.
  // PERSIST (launches with more tiles than SMs, i.e. the VAE convolutions): 1-D grid of one CTA per SM, CTA c walks
  // tiles c, c + gridDim.x, ... (tile t -> (t % tiles_m, t / tiles_m)); with two accumulator buffers in TMEM the
  // epilogue of one tile overlaps the main loop of the next, and the smem ring simply keeps running across tiles.
  // PAIR + PERSIST: the cluster (CTA pair) is the scheduling unit: pair c walks tile pairs c, c + gridDim.x / 2, ...; tile pair t ->
  // M tiles (2 (t % (tiles_m / 2)), + 1) of N tile t / (tiles_m / 2); tiles_m is even (tc_gemm_geometry).
  constexpr bool PP = PAIR && PERSIST;
  const int tiles_m_s = PP ? p.tiles_m >> 1 : p.tiles_m;     // M extent of the tile index the loops walk
  const int num_tiles = tiles_m_s * p.tiles_n;
  const int tile_first = PP ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PP ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t ncols = (uint32_t)p.tmem_cols;        // power of two in [32, 512] covering acc_bufs * (n_acc + aux) * BN columns

  // ---- prologue: touches only constants, shared memory and TMEM -> may overlap the previous kernel (PDL) ----
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);                 // PAIR: the leader's arrive.expect_tx covers the bytes of both CTAs
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);
      // PAIR + PERSIST: one arrival per epilogue warp of both CTAs, all on the leader's barrier (the issuer lives there)
      mbar_init(smem_u32(&bar_tempty[b]), PP ? 2 * TcGeo<BN>::EPI_WARPS : TcGeo<BN>::EPI_THREADS);
    }
    if (MODE == TC_EPI_PLAIN)
      for (int w = 0; w < 16; ++w) mbar_init(smem_u32(&bar_res[w]), 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.map_b);
    tma_prefetch_desc(&p.map_a[0]);
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_2sm(smem_u32(&tmem_holder), ncols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(smem_u32(&tmem_holder), ncols);
      tmem_relinquish();
    }
  }
  if (threadIdx.x < p.num_runs * 2)                  // num_runs <= TC_MAX_RUNS (host check); 2 x 16 bytes per run
    reinterpret_cast<uint4*>(runs_s)[threadIdx.x] = reinterpret_cast<const uint4*>(p.runs)[threadIdx.x];
  tc_fence_before();
  if (PAIR) {
    __syncwarp();
    cluster_sync_all();
    __syncthreads();          // implied by the cluster barrier; spelled out because compute-sanitizer's racecheck does not model barrier.cluster
  } else {
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  griddep_launch();
  if (p.dbg && threadIdx.x == 0) ts[1] = clock64();           // prologue done

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
      asm volatile("" : "+r"(full0), "+r"(empty0));
      griddep_wait();                                   // activations of the previous layer are complete from here on
      if (p.dbg) ts[2] = clock64();                     // dependency resolved
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
      const int tile_m = PERSIST ? (PP ? 2 * (tile % tiles_m_s) + (int)cta_rank : tile % tiles_m_s) : (int)blockIdx.x;
      const int n0 = (PERSIST ? tile / tiles_m_s : (int)blockIdx.y) * BN;
      const int q = tile_m / p.tiles_per_item, r = tile_m - q * p.tiles_per_item;
      const int c2_base = r * p.rows_step, c3 = q * p.items_per_tile;
      // Runs of stages that differ only by their channel block: the per-stage work is a barrier wait, the byte-count
      // arrive and the TMA instructions with two coordinates advanced by adds (this thread is a scalar in-order
      // stream; decoding a table entry per stage cost more cycles than the stage's MMAs take).
      // run fields come from the kernel parameters (constant bank -> uniform registers, no register-to-uniform moves
      // in the per-stage loop) when the table fits there, else from shared memory; the loop body is instantiated twice
      auto produce_run = [&](int r_src_acc, int r_c0, int r_d12, int r_wk, int r_count) {
        const uint32_t nw = (uint32_t)(r_src_acc >> 16) & 0xffu;
        const int d1 = (int)(short)(r_d12 & 0xffff), c2 = c2_base + (r_d12 >> 16);
        const CUtensorMap* map_a = &p.map_a[r_src_acc & 0xff];
        const uint32_t tx = (PAIR ? 2u : 1u) * (a_bytes + nw * (uint32_t)B_BYTES) + (tail_tile ? (uint32_t)p.n_tail * (TC_BK * 2) : 0u);
        int c0 = r_c0, wkc = r_wk * TC_BK;
        for (int i = 0; i < r_count; ++i) {
          if (PERSIST && p.dbg) {
            const long long t = clock64();
            mbar_wait(empty0 + 8u * stage, phase ^ 1u);
            ts[7] += clock64() - t;                     // producer: cycles waiting for a free stage
          } else {
            mbar_wait(empty0 + 8u * stage, phase ^ 1u);
          }
          const uint32_t bar = full0 + 8u * stage;
          const uint32_t sa = smem_base + stage * stage_bytes;
          if (PAIR) {
            const uint32_t bar_leader = mapa_shared(bar, 0);
            if (leader) mbar_arrive_expect_tx(bar, tx);
            tma_load_4d_2sm(sa, map_a, bar_leader, c0, d1, c2, c3);
            for (uint32_t j = 0; j < nw; ++j)
              tma_load_2d_2sm(sa + a_bytes + j * B_BYTES, &p.map_b, bar_leader, wkc + (int)j * TC_BK, n0 + (int)cta_rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(bar, tx);
            tma_load_4d(sa, map_a, bar, c0, d1, c2, c3);
            if (nw == 1) {
              tma_load_2d(sa + a_bytes, &p.map_b, bar, wkc, n0);
              if (MODE == TC_EPI_DDPM && tail_tile) tma_load_2d(sa + a_bytes + B_BYTES, &p.map_b_tail, bar, wkc, n0 + BN);
            } else {
              for (uint32_t j = 0; j < nw; ++j) tma_load_2d(sa + a_bytes + j * B_BYTES, &p.map_b, bar, wkc + (int)j * TC_BK, n0);
            }
          }
          c0 += TC_BK;
          wkc += (int)nw * TC_BK;
          if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1u; }
        }
      };
      if (p.num_runs_c > 0) {
        for (int ri = 0; ri < p.num_runs_c; ++ri)
          produce_run(p.runs_c[ri].src_acc, p.runs_c[ri].c0, p.runs_c[ri].d12, p.runs_c[ri].wk, p.runs_c[ri].count);
      } else {
        for (int ri = 0; ri < p.num_runs; ++ri) {
          const TcRun e = runs_s[ri];
          produce_run(e.src_acc, e.c0, e.d12, e.wk, e.count);
        }
      }
      if (!PERSIST) break;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (PAIR: the leader CTA issues for both) =====================
    // The issuing thread is a scalar in-order stream (~5-10 cycles per instruction) and the tensor pipe queues only a
    // few MMAs ahead of it, so every instruction between two MMAs of consecutive stages shows up as idle tensor time
    // (scripts/mma_ubench.cu).  The loop therefore never reads the stage table (the table's shape is two uniform
    // segments: kb_main stages of nw_main W tiles, then the aux stages), keeps barrier addresses and the descriptor
    // words in registers and advances them by adds.
    if (leader && elect_one()) {
      constexpr uint32_t idesc_full = umma_idesc_bf16(PAIR ? 2 * TC_BM : TC_BM, BN);
      const uint32_t idesc = (MODE == TC_EPI_DDPM && tail_tile) ? umma_idesc_bf16(TC_BM, BN + p.n_tail) : idesc_full;
      constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      constexpr uint32_t B_STEP = (uint32_t)B_BYTES >> 4;
      const uint32_t stage_step = stage_bytes >> 4;
      const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
      uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
      asm volatile("" : "+r"(full0), "+r"(empty0));     // keep them in registers: recomputing costs an S2UR per stage
      const int kb_main = p.kb_main > 0 ? p.kb_main : p.num_kb;
      const uint32_t nw_main = p.kb_main > 0 ? (uint32_t)p.nw_main : (uint32_t)p.w_max;
      const int kb_aux = p.num_kb - kb_main;
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++it) {
      const int buf = PERSIST ? it % p.acc_bufs : 0;
      const uint32_t use = PERSIST ? (uint32_t)(it / p.acc_bufs) : 0u;
      if (PERSIST) {
        const long long tw = p.dbg ? clock64() : 0;
        if (PP) mbar_wait_cluster(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);   // both CTAs' epilogues drained this buffer
        else mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);              // epilogue drained this buffer (first use: passes)
        if (p.dbg) ts[8] += clock64() - tw;                                       // issuer: cycles waiting for an accumulator buffer
        tc_fence_after();
      }
      const uint32_t acc_base = tmem_base + (uint32_t)buf * (uint32_t)p.acc_stride;
      uint32_t accf = 0;
      if (nw_main == 1) {
        // one W tile per stage (per-tap convolutions, dense layers): 4 MMAs per stage, so the loop body is kept to the
        // barrier wait, four MMAs on running 64-bit descriptors, the commit and a handful of adds
        uint64_t da = ((uint64_t)DESC_HI << 32) | (uint64_t)(a_lo0 + stage * stage_step);
        uint32_t fb = full0 + 8u * stage, eb = empty0 + 8u * stage;
        const uint64_t da_wrap = (uint64_t)((uint32_t)STAGES * stage_step);
        const bool dbg_on = p.dbg != nullptr;
        for (int kb = 0; kb < kb_main; ++kb) {
          if (PERSIST && dbg_on) {
            const long long t = clock64();
            mbar_wait(fb, phase);
            ts[9] += clock64() - t;                       // issuer: cycles waiting for operands
          } else {
            mbar_wait(fb, phase);
          }
          tc_fence_after();
          if (dbg_on && kb == 0) ts[3] = clock64();
          const uint64_t db = da + (a_bytes >> 4);
          if (PAIR) {
            umma_bf16_ss_2sm(acc_base, da, db, idesc, accf);
            umma_bf16_ss_2sm(acc_base, da + 2, db + 2, idesc, 1u);
            umma_bf16_ss_2sm(acc_base, da + 4, db + 4, idesc, 1u);
            umma_bf16_ss_2sm(acc_base, da + 6, db + 6, idesc, 1u);
            umma_commit_2sm(eb, 3);
          } else {
            umma_bf16_ss(acc_base, da, db, idesc, accf);
            umma_bf16_ss(acc_base, da + 2, db + 2, idesc, 1u);
            umma_bf16_ss(acc_base, da + 4, db + 4, idesc, 1u);
            umma_bf16_ss(acc_base, da + 6, db + 6, idesc, 1u);
            umma_commit(eb);
          }
          accf = 1u;
          da += stage_step; fb += 8u; eb += 8u;
          if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1u; da -= da_wrap; fb = full0; eb = empty0; }
        }
      } else
      for (int kb = 0; kb < kb_main; ++kb) {
        mbar_wait(full0 + 8u * stage, phase);
        tc_fence_after();
        if (p.dbg && kb == 0) ts[3] = clock64();        // first operands landed
        uint32_t a_lo = a_lo0 + stage * stage_step;
        uint32_t b_lo = a_lo + (a_bytes >> 4);
        uint32_t d = acc_base;
        // the A tile is shared by the taps' accumulators; shared-tap-row mode (PERSIST only): one accumulator, the A window moves instead
        const bool same_acc = PERSIST && p.taps_same_acc != 0;
        const uint32_t a_step = same_acc ? (uint32_t)p.a_tap_shift16 : 0u, d_step = same_acc ? 0u : (uint32_t)BN;
        for (uint32_t j = 0; j < nw_main; ++j, b_lo += B_STEP, d += d_step, a_lo += a_step) {
          umma_lohi<PAIR>(d, a_lo, b_lo, DESC_HI, idesc, (same_acc && j > 0) ? 1u : accf);
          umma_lohi<PAIR>(d, a_lo + 2, b_lo + 2, DESC_HI, idesc, 1u);          // +16 bf16 = 32 bytes along K inside the swizzle row
          umma_lohi<PAIR>(d, a_lo + 4, b_lo + 4, DESC_HI, idesc, 1u);
          umma_lohi<PAIR>(d, a_lo + 6, b_lo + 6, DESC_HI, idesc, 1u);
        }
        if (PAIR) umma_commit_2sm(empty0 + 8u * stage, 3);           // frees the stage in both CTAs
        else umma_commit(empty0 + 8u * stage);                       // frees the smem stage when these MMAs retire
        accf = 1u;
        if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1u; }
      }
      accf = 0;
      const uint32_t d_aux = acc_base + (uint32_t)p.n_acc * BN;
      for (int kb = 0; kb < kb_aux; ++kb) {
        mbar_wait(full0 + 8u * stage, phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + stage * stage_step;
        const uint32_t b_lo = a_lo + (a_bytes >> 4);
        umma_lohi<PAIR>(d_aux, a_lo, b_lo, DESC_HI, idesc, accf);
        umma_lohi<PAIR>(d_aux, a_lo + 2, b_lo + 2, DESC_HI, idesc, 1u);
        umma_lohi<PAIR>(d_aux, a_lo + 4, b_lo + 4, DESC_HI, idesc, 1u);
        umma_lohi<PAIR>(d_aux, a_lo + 6, b_lo + 6, DESC_HI, idesc, 1u);
        if (PAIR) umma_commit_2sm(empty0 + 8u * stage, 3);
        else umma_commit(empty0 + 8u * stage);
        accf = 1u;
        if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1u; }
      }
      if (PAIR) umma_commit_2sm(smem_u32(&bar_tfull[buf]), 3);
      else umma_commit(smem_u32(&bar_tfull[buf]));      // accumulator(s) of this tile complete
      if (p.dbg && it == 0) ts[4] = clock64();          // all MMAs of the first tile issued
      if (!PERSIST) break;
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int ew = warp - 2;
    if (p.l2_prefetch_bytes != 0 && threadIdx.x == 64) {
      // weights never depend on earlier kernels: issued ahead of griddepcontrol.wait
      const unsigned nctas = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
      const unsigned chunk = (((p.l2_prefetch_bytes + nctas - 1) / nctas) + 4095u) & ~4095u;
      const unsigned lo = cta * chunk, hi = min(p.l2_prefetch_bytes, lo + chunk);
      const char* base = reinterpret_cast<const char*>(p.l2_prefetch);
      for (unsigned off = lo; off < hi; off += 4096u) {
        const unsigned n = min(4096u, hi - off) & ~15u;
        if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(n) : "memory");
      }
    }
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    const int part = ew >> 2;                            // which column slice of the tile
    const int row = quarter * 32 + lane;
    griddep_wait();                                      // the step counter / residuals / x belong to earlier kernels
    // The epilogue warps have nothing to do until the accumulators are complete except staging per-column vectors and prefetching
    // FiLM pairs (~200 instructions each).  Doing that right away competes with the TMA producer and the MMA issuer - single
    // threads whose every instruction is on the critical path while the ring fills - for issue slots (measured: the FiLM
    // prefetch alone stretched the main loop of a conv1 layer by 1.5 - 1.9 k cycles).  Sleep through the pipeline fill instead.
    if (MODE == TC_EPI_GN && p.epi_sleep_ns > 0) __nanosleep(p.epi_sleep_ns);
    int it = 0;
    const uint32_t epi_box = (p.epi_tma & 8) ? TC_EPI_BUF / 2 : TC_EPI_BUF;                         // TMA epilogue: per-warp staging buffer,
    const uint32_t epi_buf = smem_base + (uint32_t)STAGES * stage_bytes + (uint32_t)ew * epi_box;   // behind the ring
    const uint32_t my_bar_res = smem_u32(&bar_res[ew & 15]);
    uint32_t res_phase = 0;
    int staged_n0 = -1;
    for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++it) {
    const int tile_m = PERSIST ? (PP ? 2 * (tile % tiles_m_s) + (int)cta_rank : tile % tiles_m_s) : (int)blockIdx.x;
    const int n0 = (PERSIST ? tile / tiles_m_s : (int)blockIdx.y) * BN;
    const int buf = PERSIST ? it % p.acc_bufs : 0;
    const uint32_t use = PERSIST ? (uint32_t)(it / p.acc_bufs) : 0u;
    const int m = tile_m * TC_BM + row;
    // stage the per-column vectors while the main loop runs; the persistent kernel does it only when the N tile changes (uniform across
    // the CTA; the barrier that ends the previous tile already separates this write from that tile's reads)
    if (!PERSIST || n0 != staged_n0) {
      staged_n0 = n0;
      const int et = threadIdx.x - 64;
      const bool uniform_step = p.film && p.step.rows == nullptr;
      const float* trow = uniform_step ? p.ttab + (long long)step_of(p.step, 0) * p.ld_ttab + p.film_off : nullptr;
      for (int i = et; i < BN; i += TcGeo<BN>::EPI_THREADS) {
        const int n = n0 + i;
        const bool ok = n < p.N;
        es.bias[i] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
        if (MODE == TC_EPI_GN || MODE == TC_EPI_LN) {
          es.bias2[i] = (ok && p.bias_aux) ? __ldg(p.bias_aux + n) : 0.f;
          es.gamma[i] = (ok && p.gamma) ? __ldg(p.gamma + n) : 0.f;
          es.beta[i] = (ok && p.beta) ? __ldg(p.beta + n) : 0.f;
          es.fscale[i] = (ok && trow) ? __ldg(trow + n) : 0.f;
          es.fshift[i] = (ok && trow) ? __ldg(trow + p.film_c + n) : 0.f;
          const __half2 ft = __floats2half2_rn(es.fscale[i], es.fshift[i]);
          es.film_t[i] = *reinterpret_cast<const uint32_t*>(&ft);
        }
      }
      epi_bar<BN>();
    }
    const int c_begin = part * TcGeo<BN>::CPP;
    if (MODE == TC_EPI_PLAIN && (p.epi_tma & 4) && lane == 0 && n0 + c_begin * 32 < p.N) {
      // the residual box of this warp's first chunk: in flight while the main loop runs
      bulk_wait_read0();                                 // the previous tile's stores have read the buffer
      mbar_arrive_expect_tx(my_bar_res, epi_box);
      tma_load_2d(epi_buf, &p.epi_maps[2], my_bar_res, n0 + c_begin * 32, tile_m * TC_BM + quarter * 32);
    }
    GnPrefetch<MODE == TC_EPI_GN ? BN : 64> pf;
    if constexpr (MODE == TC_EPI_GN) gn_prefetch<BN>(p, es, m, n0, c_begin, pf);
    {
      const long long tw = (PERSIST && p.dbg) ? clock64() : 0;
      mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
      if (PERSIST && p.dbg && threadIdx.x == 64) ts[10] += clock64() - tw;   // epilogue: cycles waiting for accumulators
    }
    tc_fence_after();
    if (p.dbg && threadIdx.x == 64 && it == 0) ts[5] = clock64();   // accumulators complete
    const uint32_t taddr = tmem_base + (uint32_t)buf * (uint32_t)p.acc_stride + ((uint32_t)(quarter * 32) << 16);
    if constexpr (MODE == TC_EPI_PLAIN) {
      if (p.epi_tma & 8) epilogue_plain_tma<BN, true>(p, es, taddr, m, n0, c_begin, lane, tile_m, quarter, epi_buf, my_bar_res, res_phase);
      else if (p.epi_tma) epilogue_plain_tma<BN, false>(p, es, taddr, m, n0, c_begin, lane, tile_m, quarter, epi_buf, my_bar_res, res_phase);
      else epilogue_plain<BN>(p, es, taddr, m, n0, c_begin, lane, tile_m, quarter);
    }
    else if constexpr (MODE == TC_EPI_GN) epilogue_gn<BN>(p, es, taddr, m, n0, c_begin, row, lane, pf);
    else if constexpr (MODE == TC_EPI_DDPM)
      epilogue_ddpm<BN>(p, es, taddr, reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))), tile_m, n0,
                        c_begin, row, (int)threadIdx.x - 64, lane, -1, nullptr, tail_tile ? p.n_tail : 0);
    else epilogue_ln<BN>(p, es, taddr, m, n0, c_begin, row, part, lane);
    if (!PERSIST) {
      if (MODE == TC_EPI_PLAIN && p.epi_tma && lane == 0) bulk_wait_read0();
      break;
    }
    tc_fence_before();                                   // hand the accumulator buffer back, protect `es`
    if (PP) {
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&bar_tempty[buf]), 0));
    } else {
      mbar_arrive(smem_u32(&bar_tempty[buf]));
    }
    epi_bar<BN>();
    }
    if (PERSIST && MODE == TC_EPI_PLAIN && p.epi_tma && lane == 0) bulk_wait_read0();   // shared memory must outlive the last store's read
  }

  if (p.dbg && threadIdx.x == 64) ts[6] = clock64();     // this warp's epilogue done
  if (MODE == TC_EPI_DDPM && p.step_dec != nullptr) {
    // every CTA has read the step counter by now; the last one to get here moves it to the next timestep
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned total = gridDim.x * gridDim.y;
      if (atomicAdd(p.done_counter, 1u) == total - 1u) {
        *reinterpret_cast<volatile int32_t*>(p.step_dec) = *reinterpret_cast<volatile int32_t*>(p.step_dec) - 1;
        *reinterpret_cast<volatile unsigned int*>(p.done_counter) = 0u;
        __threadfence();
      }
    }
  }
  tc_fence_before();
  if (PAIR) {
    __syncwarp();
    cluster_sync_relaxed();   // the peer's MMAs read this CTA's shared memory and arrive on its barriers until here
    if (warp == 1) tmem_dealloc_2sm(tmem_base, ncols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
  }
  if (p.dbg && threadIdx.x == 0) {
    long long* d = p.dbg + (long long)(blockIdx.y * gridDim.x + blockIdx.x) * 8;
    const long long t0 = ts[0];
    for (int i = 1; i < 7; ++i) d[i] = ts[i] ? ts[i] - t0 : 0;
    d[7] = clock64() - t0;
    if (PERSIST) { d[2] = ts[7]; d[3] = ts[8]; d[4] = ts[9]; d[5] = ts[10]; }    // wait-cycle sums instead of first-tile stamps
    if (p.dbg_stage && blockIdx.x == 0 && blockIdx.y == 0)
      for (int i = 0; i < 23; ++i) p.dbg_stage[i] = tk[i] ? tk[i] - t0 : 0;
    d[0] = (long long)(__cvta_generic_to_shared(&ts[0]) & 0) + (long long)blockIdx.x;   // tile id
  }
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

int make_tmap_bf16_strided(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled fn = get_encode_fn();
  LDP_CHECK(fn != nullptr, LDP_ERR_NO_DEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  LDP_CHECK(rank >= 2 && rank <= 4, LDP_ERR_INVALID_ARG, "tensor map rank must be 2..4");
  cuuint64_t gdim[4];
  cuuint64_t gstr[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
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

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap_bf16_strided(out, base, rank, dims, strides_bytes, box, nullptr);
}

int make_tmap_epi(CUtensorMap* out, const void* base, bool f32, uint64_t cols, uint64_t rows, uint64_t ld, int box_cols) {
  PFN_encodeTiled fn = get_encode_fn();
  LDP_CHECK(fn != nullptr, LDP_ERR_NO_DEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  const uint64_t es = f32 ? 4 : 2;
  LDP_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * es) % 16 == 0 && cols % 32 == 0, LDP_ERR_INVALID_ARG,
            "epilogue tensor map: base / leading dimension must be 16-byte aligned, columns a multiple of 32");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * es};
  cuuint32_t bx[2] = {(cuuint32_t)box_cols, 32}, estr[2] = {1, 1};
  const uint64_t row_bytes = (uint64_t)box_cols * es;
  LDP_CHECK(row_bytes == 64 || row_bytes == 128, LDP_ERR_INVALID_ARG, "epilogue tensor map: box rows must be 64 or 128 bytes");
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, bx,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (epilogue map) failed with CUresult " + std::to_string((int)r));
    return LDP_ERR_CUDA;
  }
  return LDP_OK;
}

// Host side of TcGemm::epi_tma: which of out_f32 / out_bf16 / res_f32 can move as 32 x 32 boxes, and their maps (the caller copies the
// three maps to device memory and sets op->epi_maps / op->epi_tma = *bits).
int tc_build_epi_maps(const TcGemm& op, size_t rows, CUtensorMap host[3], int* bits, bool half) {
  *bits = 0;
  const int fcols = half ? 16 : 32;
  memset(host, 0, 3 * sizeof(CUtensorMap));
  if (op.mode != TC_EPI_PLAIN || op.N % 32 != 0 || op.n_acc != 1 || op.shift[0] != 0 || op.use_aux) return LDP_OK;
  auto ok = [](const void* p) { return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (ok(op.out_f32) && op.ld_out_f32 % 4 == 0) { LDP_TRY(make_tmap_epi(&host[0], op.out_f32, true, op.N, rows, op.ld_out_f32, fcols)); *bits |= 1; }
  if (ok(op.out_bf16) && op.ld_out_bf16 % 8 == 0) { LDP_TRY(make_tmap_epi(&host[1], op.out_bf16, false, op.N, rows, op.ld_out_bf16, 32)); *bits |= 2; }
  if (ok(op.res_f32) && op.ld_res_f32 % 4 == 0) { LDP_TRY(make_tmap_epi(&host[2], op.res_f32, true, op.N, rows, op.ld_res_f32, fcols)); *bits |= 4; }
  if (*bits && half) *bits |= 8;
  return LDP_OK;
}

template <int MODE>
static constexpr int ring_budget() { return MODE == TC_EPI_PLAIN ? TC_SMEM_RING_PLAIN : TC_SMEM_RING; }

template <int BN, int MODE>
static int set_smem_attr() {
  constexpr int TC_SMEM_RING = ring_budget<MODE>();        // (shadows the global: every attribute below uses this kernel family's budget)
  LDP_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<BN, MODE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TC_SMEM_RING + 1024));
  constexpr bool kPairable = (MODE == TC_EPI_PLAIN || MODE == TC_EPI_GN) && BN <= 128;
  if (kPairable)
    LDP_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<BN, MODE, kPairable, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TC_SMEM_RING + 1024));
  constexpr bool kPersistable = MODE == TC_EPI_PLAIN;
  if (kPersistable)
    LDP_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<BN, MODE, false, kPersistable>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TC_SMEM_RING + 1024));
  LDP_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel_v1<BN, MODE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TC_SMEM_RING + 1024));
  if (kPairable)
    LDP_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel_v1<BN, MODE, kPairable, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TC_SMEM_RING + 1024));
  if (kPersistable)
    LDP_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel_v1<BN, MODE, false, kPersistable>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TC_SMEM_RING + 1024));
  constexpr bool kPairPersist = MODE == TC_EPI_PLAIN && BN >= 128;
  if (kPairPersist)
    LDP_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<BN, MODE, kPairPersist, kPairPersist>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TC_SMEM_RING + 1024));
  return LDP_OK;
}

static bool g_use_pdl = true;

int tc_gemm_init() {
  static bool done = false;
  if (done) return LDP_OK;
  LDP_TRY((set_smem_attr<64, TC_EPI_PLAIN>()));
  LDP_TRY((set_smem_attr<64, TC_EPI_GN>()));
  LDP_TRY((set_smem_attr<128, TC_EPI_PLAIN>()));
  LDP_TRY((set_smem_attr<128, TC_EPI_GN>()));
  LDP_TRY((set_smem_attr<128, TC_EPI_DDPM>()));
  LDP_TRY((set_smem_attr<256, TC_EPI_PLAIN>()));
  LDP_TRY((set_smem_attr<256, TC_EPI_LN>()));
  const char* env = getenv("LDP_NO_PDL");
  g_use_pdl = !(env && env[0] == '1');
  done = true;
  return LDP_OK;
}

// Tile / grid / TMEM geometry of a launch (also used by the profiling entry point to size its per-CTA buffers).
int tc_gemm_geometry(TcGemm* p) {
  const int bn = p->block_n;
  const int n_acc_total = p->n_acc + (p->use_aux ? 1 : 0);
  LDP_CHECK(n_acc_total * bn <= 512, LDP_ERR_UNSUPPORTED, "tc_gemm: accumulators exceed the 512 TMEM columns");
  const int tm = ceil_div(p->M, TC_BM);
  p->tiles_m = p->pair ? round_up(tm, 2) : tm;
  p->tiles_n = ceil_div(p->N, bn);
  p->n_tail = 0;
  if (p->mode == TC_EPI_DDPM && p->allow_tail && !p->pair && p->tiles_n >= 2 && p->n_acc == 1 && !p->use_aux &&
      p->N - (p->tiles_n - 1) * bn <= 16 && p->w_max == 1) {
    p->tiles_n -= 1;                 // the last N tile takes the remaining <= 16 columns as well
    p->n_tail = 16;
  }
  const int tiles = p->tiles_m * p->tiles_n;
  static int persist = -1;
  if (persist < 0) { const char* e = getenv("LDP_PERSIST"); persist = (e && e[0] == '0') ? 0 : 1; }
  // pairs + persistence (the VAE convolutions): PLAIN epilogue, BN >= 128; a paired launch that is not persistent needs BN <= 128
  // (the shared-tap-row layout exists only in the persistent kernel: such a launch is persistent whatever its tile count)
  // ... and so is a paired 256-wide launch (no plain-grid kernel is instantiated for it): ops are built for a full chunk of images and
  // replayed for the last, smaller one
  const bool persistent = (persist && tiles > 148 && p->mode == TC_EPI_PLAIN && (!p->pair || bn >= 128)) || p->a_rows == 256 ||
                          (p->pair && bn == 256 && p->mode == TC_EPI_PLAIN);
  p->grid_ctas = persistent ? std::min(148, tiles) : tiles;
  p->persistent = persistent ? 1 : 0;
  p->acc_stride = n_acc_total * bn + p->n_tail;
  p->acc_bufs = (persistent && 2 * p->acc_stride <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p->acc_bufs * p->acc_stride) cols <<= 1;
  p->tmem_cols = cols;
  return LDP_OK;
}

template <int BN, int MODE, bool PAIR, bool PERSIST>
static int launch_tc_gemm_inst(const TcGemm& p_in, cudaStream_t s) {
  LDP_TRY(tc_gemm_init());
  TcGemm p = p_in;
  const int n_acc_total = p.n_acc + (p.use_aux ? 1 : 0);
  LDP_CHECK(p.n_acc >= 1 && p.n_acc <= 5 && n_acc_total * BN <= 512, LDP_ERR_UNSUPPORTED,
            "tc_gemm: accumulators exceed the 512 TMEM columns");
  LDP_CHECK(p.w_max >= 1 && p.w_max <= 5, LDP_ERR_INVALID_ARG, "tc_gemm: w_max must be 1..5");
  LDP_TRY(tc_gemm_geometry(&p));
  LDP_CHECK((p.pair != 0) == PAIR && (p.persistent != 0) == PERSIST, LDP_ERR_INVALID_ARG,
            "tc_gemm: pair / persistent flags do not match the kernel");
  {
    static int skip = -1;                      // diagnostics: LDP_EPI_SKIP bit mask disables parts of the GN epilogue
    if (skip < 0) { const char* e = getenv("LDP_EPI_SKIP"); skip = e ? atoi(e) : 0; }
    p.epi_skip = skip;
    static int sleep_ns = -1;
    if (sleep_ns < 0) { const char* e = getenv("LDP_EPI_SLEEP"); sleep_ns = e ? atoi(e) : 1000; }
    p.epi_sleep_ns = (unsigned)sleep_ns;

  }
  LDP_CHECK(p.a_rows == 128 || (PERSIST && p.a_rows == 256 && p.taps_same_acc && p.n_acc == 1), LDP_ERR_INVALID_ARG,
            "tc_gemm: 256-row A boxes exist only in the persistent kernel's shared-tap-row mode");
  const int a_bytes = p.a_rows * TC_BK * 2;
  const int stage_bytes = a_bytes + p.w_max * (PAIR ? BN / 2 : BN) * TC_BK * 2 + p.n_tail * TC_BK * 2;
  const int epi_bytes = p.epi_tma ? TcGeo<BN>::EPI_WARPS * ((p.epi_tma & 8) ? TC_EPI_BUF / 2 : TC_EPI_BUF) : 0;
  if (p.epi_tma)
    LDP_CHECK(MODE == TC_EPI_PLAIN && p.epi_maps && p.N % 32 == 0 && p.n_acc == 1 && p.shift[0] == 0 && !p.use_aux, LDP_ERR_INVALID_ARG,
              "tc_gemm: the TMA epilogue needs the PLAIN epilogue, one accumulator and N % 32 == 0");
  p.num_stages = std::min(TC_MAX_STAGES, (ring_budget<MODE>() - epi_bytes) / stage_bytes);
  LDP_CHECK(p.num_stages >= 2, LDP_ERR_UNSUPPORTED, "tc_gemm: stage does not fit the shared-memory ring twice");
  if (MODE == TC_EPI_DDPM)
    LDP_CHECK(p.num_stages * stage_bytes >= TC_BM * (BN + p.n_tail + 1) * 4, LDP_ERR_UNSUPPORTED,
              "tc_gemm DDPM epilogue: transposition tile does not fit the ring");
  if (p.n_acc > 1 || p.shift[0] != 0)
    LDP_CHECK(p.rows_per_item >= 1 && p.rows_per_item <= 32 && (p.rows_per_item & (p.rows_per_item - 1)) == 0,
              LDP_ERR_UNSUPPORTED, "tc_gemm: shifted accumulators need power-of-two rows per sample <= 32");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = PERSIST ? dim3(p.grid_ctas, 1, 1) : dim3(p.tiles_m, p.tiles_n, 1);
  cfg.blockDim = dim3(TcGeo<BN>::THREADS, 1, 1);
  cfg.dynamicSmemBytes = p.num_stages * stage_bytes + epi_bytes + 1024;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (PAIR && PERSIST) {
    // one CTA pair per TPC: ask how many clusters fit (74 on a full B200) instead of assuming it
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t q = cfg;
      q.gridDim = dim3(148, 1, 1);
      q.attrs = attr + (g_use_pdl ? 1 : 0);
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, tc_gemm_kernel<BN, MODE, PAIR, PERSIST>, &q) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        n = 74;
      }
      max_clusters = std::min(n, 74);
    }
    cfg.gridDim = dim3(2 * std::min(max_clusters, p.tiles_m / 2 * p.tiles_n), 1, 1);
  }
  // ops that use none of the later features (TMA epilogue, paired persistence, shared tap rows) run on the v1 body (tc_gemm_v1.cuh)
  const bool v2 = p.epi_tma != 0 || (PAIR && PERSIST) || p.a_rows != 128;
  cudaError_t e;
  if constexpr (PAIR && PERSIST) {
    e = cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, MODE, PAIR, PERSIST>, p);
  } else {
    e = v2 ? cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, MODE, PAIR, PERSIST>, p)
           : cudaLaunchKernelEx(&cfg, tc_gemm_kernel_v1<BN, MODE, PAIR, PERSIST>, p);
  }
  if (e != cudaSuccess) {
    set_last_error(std::string("tc_gemm launch failed: ") + cudaGetErrorString(e));
    return LDP_ERR_CUDA;
  }
  count_launch();
  return LDP_OK;
}

void tc_set_inline_runs(TcGemm* op, const TcRun* runs_host, int n) {
  op->num_runs_c = 0;
  if (n <= 0 || n > TC_RUNS_INLINE) return;
  for (int i = 0; i < n; ++i) op->runs_c[i] = runs_host[i];
  op->num_runs_c = n;
}

int launch_tc_gemm(const TcGemm& p, cudaStream_t s) {
  LDP_CHECK(p.kb && p.num_kb > 0 && p.M > 0 && p.N > 0, LDP_ERR_INVALID_ARG, "tc_gemm: bad arguments");
  LDP_CHECK(p.runs && p.num_runs > 0 && p.num_runs <= TC_MAX_RUNS, LDP_ERR_UNSUPPORTED, "tc_gemm: run table missing or larger than TC_MAX_RUNS");
  LDP_CHECK(p.block_n == 64 || p.block_n == 128 || p.block_n == 256, LDP_ERR_INVALID_ARG,
            "tc_gemm: block_n must be 64, 128 or 256");
  if (p.mode == TC_EPI_GN) {
    LDP_CHECK(p.group_width % 32 == 0 && p.group_width <= p.block_n && p.N % p.group_width == 0, LDP_ERR_UNSUPPORTED,
              "tc_gemm GN epilogue: group width must be a multiple of 32 and fit the N tile");
    LDP_CHECK(p.rows_per_item >= 1 && p.rows_per_item <= 32 && (p.rows_per_item & (p.rows_per_item - 1)) == 0,
              LDP_ERR_UNSUPPORTED, "tc_gemm GN epilogue: rows per sample must be a power of two <= 32");
    LDP_CHECK(p.bias && p.gamma && p.beta, LDP_ERR_INVALID_ARG, "tc_gemm GN epilogue: bias/gamma/beta required");
  }
  if (p.mode == TC_EPI_LN) LDP_CHECK(p.N == p.block_n && p.out_f32 && p.bias, LDP_ERR_UNSUPPORTED, "tc_gemm LN epilogue: N must equal block_n");
  if (p.mode == TC_EPI_DDPM) LDP_CHECK(p.x_io && p.coef, LDP_ERR_INVALID_ARG, "tc_gemm DDPM epilogue: x / coefficient table required");
  const bool pair = p.pair != 0;
  TcGemm g = p;
  LDP_TRY(tc_gemm_geometry(&g));
  if (pair) LDP_CHECK((p.mode == TC_EPI_PLAIN || p.mode == TC_EPI_GN) && (p.block_n <= 128 || g.persistent), LDP_ERR_UNSUPPORTED,
                      "tc_gemm: pair mode exists for the PLAIN / GN epilogues at BN <= 128 (BN 256: persistent PLAIN launches only)");
  const int key = p.block_n * 8 + p.mode;
  if (pair && g.persistent) {
    switch (key) {
      case 128 * 8 + TC_EPI_PLAIN: return launch_tc_gemm_inst<128, TC_EPI_PLAIN, true, true>(p, s);
      case 256 * 8 + TC_EPI_PLAIN: return launch_tc_gemm_inst<256, TC_EPI_PLAIN, true, true>(p, s);
      default: break;
    }
  } else if (pair) {
    switch (key) {
      case 64 * 8 + TC_EPI_PLAIN:  return launch_tc_gemm_inst<64, TC_EPI_PLAIN, true, false>(p, s);
      case 64 * 8 + TC_EPI_GN:     return launch_tc_gemm_inst<64, TC_EPI_GN, true, false>(p, s);
      case 128 * 8 + TC_EPI_PLAIN: return launch_tc_gemm_inst<128, TC_EPI_PLAIN, true, false>(p, s);
      case 128 * 8 + TC_EPI_GN:    return launch_tc_gemm_inst<128, TC_EPI_GN, true, false>(p, s);
      default: break;
    }
  }
  if (g.persistent) {
    switch (key) {
      case 64 * 8 + TC_EPI_PLAIN:  return launch_tc_gemm_inst<64, TC_EPI_PLAIN, false, true>(p, s);
      case 128 * 8 + TC_EPI_PLAIN: return launch_tc_gemm_inst<128, TC_EPI_PLAIN, false, true>(p, s);
      case 256 * 8 + TC_EPI_PLAIN: return launch_tc_gemm_inst<256, TC_EPI_PLAIN, false, true>(p, s);
      default: break;
    }
  }
  switch (key) {
    case 64 * 8 + TC_EPI_PLAIN:  return launch_tc_gemm_inst<64, TC_EPI_PLAIN, false, false>(p, s);
    case 64 * 8 + TC_EPI_GN:     return launch_tc_gemm_inst<64, TC_EPI_GN, false, false>(p, s);
    case 128 * 8 + TC_EPI_PLAIN: return launch_tc_gemm_inst<128, TC_EPI_PLAIN, false, false>(p, s);
    case 128 * 8 + TC_EPI_GN:    return launch_tc_gemm_inst<128, TC_EPI_GN, false, false>(p, s);
    case 128 * 8 + TC_EPI_DDPM:  return launch_tc_gemm_inst<128, TC_EPI_DDPM, false, false>(p, s);
    case 256 * 8 + TC_EPI_PLAIN: return launch_tc_gemm_inst<256, TC_EPI_PLAIN, false, false>(p, s);
    case 256 * 8 + TC_EPI_LN:    return launch_tc_gemm_inst<256, TC_EPI_LN, false, false>(p, s);
    default: break;
  }
  set_last_error("tc_gemm: no kernel instantiated for block_n " + std::to_string(p.block_n) + " with epilogue " +
                 std::to_string(p.mode));
  return LDP_ERR_UNSUPPORTED;
}

}  // namespace ldp
