// tc_gemm_kernel_v1: the kernel exactly as the planner / IDM paths were tuned with it (round 2, before the VAE work added CTA-pair
// persistence, the TMA epilogue and shared tap rows to tc_gemm_kernel).  The single-thread TMA-producer and MMA-issuer loops are
// sensitive to instruction scheduling: those additions - all behind compile-time-false branches for the planner's instantiations -
// still reshuffled the generated code and cost the 100-step loop 3.7 % (38.4 -> 39.8 ms, bisected commit by commit on one box).
// launch_tc_gemm therefore runs every op that needs none of the new features on this body; tc_gemm_kernel serves the rest.
// Included by tc_gemm.cu only (after umma_lohi).
#pragma once

namespace ldp {

// ---- the kernel --------------------------------------------------------------------------------
template <int BN, int MODE, bool PAIR, bool PERSIST>
__global__ void __launch_bounds__(TcGeo<BN>::THREADS, 1) tc_gemm_kernel_v1(const __grid_constant__ TcGemm p) {
  extern __shared__ uint8_t smem_raw[];
  // PAIR: two CTAs of a cluster (adjacent M tiles, same N tile) run one cta_group::2 MMA (M = 256): each CTA loads its
  // own A tile but only its half of every W tile - the peer's half is read over the SM-to-SM path, not through this
  // SM's L2 ingest port, which is what bounds the single-CTA kernel.
  constexpr int B_BYTES = PAIR ? TcSmem<BN>::B_BYTES / 2 : TcSmem<BN>::B_BYTES;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int STAGES = p.num_stages;
  const uint32_t stage_bytes = TC_A_BYTES + (uint32_t)p.w_max * B_BYTES + (MODE == TC_EPI_DDPM ? (uint32_t)p.n_tail * (TC_BK * 2) : 0u);
  // DDPM: this CTA owns the widened last N tile (BN + n_tail columns)
  const bool tail_tile = MODE == TC_EPI_DDPM && p.n_tail > 0 && blockIdx.y + 1 == gridDim.y;
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2];              // accumulator buffer b complete (MMA -> epilogue)
  __shared__ __align__(8) uint64_t bar_tempty[2];             // accumulator buffer b drained (epilogue -> MMA)
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) TcRun runs_s[TC_MAX_RUNS];         // run-length stage table staged once per CTA
  __shared__ __align__(16) EpiSmem<BN> es;
  __shared__ long long ts[8];                                 // phase timestamps (diagnostics, only when p.dbg != nullptr)
  __shared__ long long tk[24];                                // arrival time of the first 24 stages at the MMA issuer
  if (p.dbg && threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) ts[i] = 0;
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
  const int num_tiles = p.tiles_m * p.tiles_n;
  const uint32_t ncols = (uint32_t)p.tmem_cols;        // power of two in [32, 512] covering acc_bufs * (n_acc + aux) * BN columns

  // ---- prologue: touches only constants, shared memory and TMEM -> may overlap the previous kernel (PDL) ----
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);                 // PAIR: the leader's arrive.expect_tx covers the bytes of both CTAs
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);
      mbar_init(smem_u32(&bar_tempty[b]), TcGeo<BN>::EPI_THREADS);
    }
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
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tile_m = PERSIST ? tile % p.tiles_m : (int)blockIdx.x;
      const int n0 = (PERSIST ? tile / p.tiles_m : (int)blockIdx.y) * BN;
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
        const uint32_t tx = (PAIR ? 2u : 1u) * (TC_A_BYTES + nw * (uint32_t)B_BYTES) + (tail_tile ? (uint32_t)p.n_tail * (TC_BK * 2) : 0u);
        int c0 = r_c0, wkc = r_wk * TC_BK;
        for (int i = 0; i < r_count; ++i) {
          mbar_wait(empty0 + 8u * stage, phase ^ 1u);
          const uint32_t bar = full0 + 8u * stage;
          const uint32_t sa = smem_base + stage * stage_bytes;
          if (PAIR) {
            const uint32_t bar_leader = mapa_shared(bar, 0);
            if (leader) mbar_arrive_expect_tx(bar, tx);
            tma_load_4d_2sm(sa, map_a, bar_leader, c0, d1, c2, c3);
            for (uint32_t j = 0; j < nw; ++j)
              tma_load_2d_2sm(sa + TC_A_BYTES + j * B_BYTES, &p.map_b, bar_leader, wkc + (int)j * TC_BK, n0 + (int)cta_rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(bar, tx);
            tma_load_4d(sa, map_a, bar, c0, d1, c2, c3);
            if (nw == 1) {
              tma_load_2d(sa + TC_A_BYTES, &p.map_b, bar, wkc, n0);
              if (MODE == TC_EPI_DDPM && tail_tile) tma_load_2d(sa + TC_A_BYTES + B_BYTES, &p.map_b_tail, bar, wkc, n0 + BN);
            } else {
              for (uint32_t j = 0; j < nw; ++j) tma_load_2d(sa + TC_A_BYTES + j * B_BYTES, &p.map_b, bar, wkc + (int)j * TC_BK, n0);
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
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = PERSIST ? it % p.acc_bufs : 0;
      const uint32_t use = PERSIST ? (uint32_t)(it / p.acc_bufs) : 0u;
      if (PERSIST) {
        mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);     // epilogue drained this buffer (first use: passes)
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
          mbar_wait(fb, phase);
          tc_fence_after();
          if (dbg_on && kb == 0) ts[3] = clock64();
          const uint64_t db = da + (TC_A_BYTES >> 4);
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
        const uint32_t a_lo = a_lo0 + stage * stage_step;
        uint32_t b_lo = a_lo + (TC_A_BYTES >> 4);
        uint32_t d = acc_base;
        for (uint32_t j = 0; j < nw_main; ++j, b_lo += B_STEP, d += BN) {     // the A tile is shared by the taps' accumulators
          umma_lohi<PAIR>(d, a_lo, b_lo, DESC_HI, idesc, accf);
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
        const uint32_t b_lo = a_lo + (TC_A_BYTES >> 4);
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int tile_m = PERSIST ? tile % p.tiles_m : (int)blockIdx.x;
    const int n0 = (PERSIST ? tile / p.tiles_m : (int)blockIdx.y) * BN;
    const int buf = PERSIST ? it % p.acc_bufs : 0;
    const uint32_t use = PERSIST ? (uint32_t)(it / p.acc_bufs) : 0u;
    const int m = tile_m * TC_BM + row;
    // stage the per-column vectors while the main loop runs
    {
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
    GnPrefetch<MODE == TC_EPI_GN ? BN : 64> pf;
    if constexpr (MODE == TC_EPI_GN) gn_prefetch<BN>(p, es, m, n0, c_begin, pf);
    mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
    tc_fence_after();
    if (p.dbg && threadIdx.x == 64 && it == 0) ts[5] = clock64();   // accumulators complete
    const uint32_t taddr = tmem_base + (uint32_t)buf * (uint32_t)p.acc_stride + ((uint32_t)(quarter * 32) << 16);
    if constexpr (MODE == TC_EPI_PLAIN) epilogue_plain<BN>(p, es, taddr, m, n0, c_begin, lane, tile_m, quarter);
    else if constexpr (MODE == TC_EPI_GN) epilogue_gn<BN>(p, es, taddr, m, n0, c_begin, row, lane, pf);
    else if constexpr (MODE == TC_EPI_DDPM)
      epilogue_ddpm<BN>(p, es, taddr, reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))), tile_m, n0,
                        c_begin, row, (int)threadIdx.x - 64, lane, -1, nullptr, tail_tile ? p.n_tail : 0);
    else epilogue_ln<BN>(p, es, taddr, m, n0, c_begin, row, part, lane);
    if (!PERSIST) break;
    tc_fence_before();                                   // hand the accumulator buffer back, protect `es`
    mbar_arrive(smem_u32(&bar_tempty[buf]));
    epi_bar<BN>();
    }
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
    if (p.dbg_stage && blockIdx.x == 0 && blockIdx.y == 0)
      for (int i = 0; i < 23; ++i) p.dbg_stage[i] = tk[i] ? tk[i] - t0 : 0;
    d[0] = (long long)(__cvta_generic_to_shared(&ts[0]) & 0) + (long long)blockIdx.x;   // tile id
  }
}


}  // namespace ldp
