// The planner's whole reverse-diffusion loop (reference agent/ldp_agent.py:459-476) as ONE persistent kernel.
//
// Why: with one launch per layer, a denoising step at B = 1024 is 31 kernels of 7-20 us each, and in every one of them
// the tensor pipe idles while the kernel launches, allocates TMEM, waits for its first operands, runs its epilogue and
// tears down (in-kernel phase clocks: ~12k of the ~36k cycles of a deep-level layer are main loop).  Plans are
// independent, so nothing forces a device-wide barrier between layers: a layer's tile only needs the previous layer's
// tiles of the SAME samples.
//
// Layout: the batch is cut into groups of `spc` samples (spc * T_deepest = 128 rows, i.e. 64 samples at T = 8); a group
// is served by G = 8 CTAs (one per SM) that walk all layers of all denoising steps together.  In every layer the group
// owns (spc * T_l / 128) M tiles x (N / BN) N tiles - exactly 8 tiles at every level of the benchmark UNet (4x2, 2x4,
// 1x8), one per CTA; CTA r of the group takes tiles r, r + G, ....  Layer l+1 of a group starts as soon as the G CTAs of
// that group have finished layer l: a monotonically increasing arrival counter per group in global memory
// (red.release.gpu by one thread per CTA after the epilogue; the TMA producer polls it with ld.acquire.gpu before it
// issues the first activation load of the next layer).  Weights do not depend on the previous layer, so the producer
// streams the W tiles of the next layer's first stages into the (idle) shared-memory ring while the epilogue of the
// current layer is still running; only the activation tiles wait for the counter.
//
// Per CTA: warp 0 = TMA producer, warp 1 = MMA issuer (TMEM allocated once for the whole loop), warps 2-17 = epilogue
// (the fused epilogues of tc_epilogue.cuh, shared with the per-layer kernel).  Layer parameters (TcGemm, with its
// tensor maps) live in a device array; each role keeps its own shared-memory copy of the current layer's parameters so
// that the roles can run ahead of each other without a CTA-wide barrier.
//
// Activations written by one CTA are read by the other CTAs of its group through TMA (async proxy) and ld.global.cg:
// writer = st.global ... bar.sync (epilogue warps) ... fence.proxy.async + red.release.gpu; reader = ld.acquire.gpu +
// fence.proxy.async before the TMA loads.
#ifdef LDP_LOOP_BACKOFF
#define LDP_MBAR_BACKOFF LDP_LOOP_BACKOFF
#endif
#include "tc_epilogue.cuh"

namespace ldp {

// The ops of one denoising step, uploaded before the launch.  Constant memory keeps every parameter access a uniform
// constant-bank operand (as in the per-layer kernel, where the parameters are kernel arguments) and lets the TMA unit
// fetch the tensor maps from it.
constexpr int PL_MAX_LAYERS = 36;
__constant__ __align__(64) uint8_t c_layers_raw[PL_MAX_LAYERS * sizeof(TcGemm)];     // TcGemm has default member initialisers
#define c_layers (reinterpret_cast<const TcGemm*>(c_layers_raw))

constexpr int PL_EPI_WARPS = 16;
constexpr int PL_THREADS = 64 + PL_EPI_WARPS * 32;

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_all_bar() { asm volatile("bar.sync 2, %0;" ::"n"(PL_EPI_WARPS * 32) : "memory"); }

// MMA with the 64-bit shared-memory descriptors given as (lo, hi) words: hi is a launch constant, lo advances.
__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct TileIter {      // the tiles of one layer that belong to this CTA, in the order every role walks them
  int tm_g, ntiles, group, r, G, M;
  __device__ __forceinline__ TileIter(const TcGemm& p, int group_, int r_, int G_)
      : tm_g(p.tiles_m_group), ntiles(p.tiles_m_group * p.tiles_n), group(group_), r(r_), G(G_), M(p.M) {}
  __device__ __forceinline__ bool valid(int t, int* tile_m, int* tile_n) const {
    *tile_m = group * tm_g + t % tm_g;
    *tile_n = t / tm_g;
    return *tile_m * TC_BM < M;
  }
};

template <int BN, int MODE>
__device__ __noinline__ void epilogue_tile(int layer, uint8_t* es_raw, uint32_t tmem_base, uint8_t* ring, int tile_m,
                                              int n0, int warp, int lane, uint32_t bar_tfull, uint32_t tfull_parity, int flags, int timestep, long long* dbg) {
  const TcGemm& p = c_layers[layer];              // formed here so that the accesses stay constant-bank loads
  EpiSmem<BN>& es = *reinterpret_cast<EpiSmem<BN>*>(es_raw);
  const int ew = warp - 2;
  const int quarter = warp & 3;
  const int part = ew >> 2;
  const int row = quarter * 32 + lane;
  const int m = tile_m * TC_BM + row;
  {
    const int et = threadIdx.x - 64;
    const bool uniform_step = p.film && p.step.rows == nullptr;
    const float* trow = uniform_step ? p.ttab + (long long)timestep * p.ld_ttab + p.film_off : nullptr;
    for (int i = et; i < BN; i += TcGeo<BN>::EPI_THREADS) {
      const int n = n0 + i;
      const bool ok = n < p.N;
      es.bias[i] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
      if (MODE == TC_EPI_GN) {
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
  mbar_wait(bar_tfull, tfull_parity);
  tc_fence_after();
  if (dbg && threadIdx.x == 64) dbg[5] = clock64();
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
  if constexpr (MODE == TC_EPI_PLAIN) epilogue_plain<BN>(p, es, taddr, m, n0, c_begin, lane);
  else if constexpr (MODE == TC_EPI_GN) epilogue_gn<BN>(p, es, taddr, m, n0, c_begin, row, lane, pf);
  else epilogue_ddpm<BN>(p, es, taddr, reinterpret_cast<float*>(ring), tile_m, n0, c_begin, row, (int)threadIdx.x - 64, lane, timestep, dbg);
  if (flags & 2) { __threadfence(); fence_proxy_async_all(); }
  tc_fence_before();
  epi_bar<BN>();      // all of this tile's accumulator reads and global stores are issued; `es` may be rewritten
}

__global__ void __launch_bounds__(PL_THREADS, 1) planner_loop_kernel(const __grid_constant__ PlannerLoop lp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull;                 // accumulators of the current tile complete (MMA -> epilogue, producer)
  __shared__ __align__(8) uint64_t bar_tempty;                // accumulators drained (epilogue -> MMA, producer)
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) TcStage kb_s[TC_MAX_KB_SMEM];
  __shared__ __align__(16) uint8_t es_raw[sizeof(EpiSmem<128>)];

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = lp.group_ctas;
  const int group = blockIdx.x / G, r = blockIdx.x - group * G;
  int* counter = lp.group_counter + group;
  const bool dbg_cta = lp.dbg != nullptr && blockIdx.x == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_MAX_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_tfull), 1);
    mbar_init(smem_u32(&bar_tempty), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
  const uint32_t tfull = smem_u32(&bar_tfull), tempty = smem_u32(&bar_tempty);

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t use_bits = 0;          // bit s: parity of the number of times ring slot s has been filled
    int tile_seq = 0;               // tiles of this CTA handed to the pipeline so far
    bool prev_ddpm = false;         // the previous tile's epilogue uses the ring as its transposition buffer
    int q = 0;                      // layers completed by the group before the current one (over all steps)
    for (int it = 0; it < lp.n_steps; ++it) {
      for (int l = 0; l < lp.n_layers; ++l, ++q) {
        const TcGemm* gl = &c_layers[l];
        __syncwarp();
        {
          const int nkb = gl->num_kb;
          const uint4* src = reinterpret_cast<const uint4*>(gl->kb);
          uint4* dst = reinterpret_cast<uint4*>(kb_s);
          for (int i = lane; i < nkb; i += 32) dst[i] = __ldg(src + i);
        }
        __syncwarp();
        if (elect_one()) {
          const TcGemm& p = *gl;
          const int BN = p.block_n;
          const uint32_t b_bytes = (uint32_t)BN * TC_BK * 2;
          const uint32_t stage_bytes = TC_A_BYTES + (uint32_t)p.w_max * b_bytes;
          const uint32_t S = (uint32_t)p.num_stages;
          const int num_kb = p.num_kb;
          const CUtensorMap* map_a = gl->map_a;            // tensor maps are read from global memory by the TMA unit
          const CUtensorMap* map_b = &gl->map_b;
          const bool dbg_now = dbg_cta && it == lp.dbg_step;
          if (dbg_now) lp.dbg[l * 8 + 0] = clock64();
          TileIter ti(p, group, r, G);
          uint32_t stage = 0;
          bool first = true;
          for (int t = r; t < ti.ntiles; t += G) {
            int tile_m, tile_n;
            if (!ti.valid(t, &tile_m, &tile_n)) continue;
            // The ring changes geometry from layer to layer: a tile's loads start only when the previous tile's MMAs have
            // consumed every stage (and, after a DDPM tile, when its epilogue has released the ring).  Waiting for the
            // previous tile on EVERY tile also keeps this thread at most one phase ahead of bar_tfull / bar_tempty, which
            // is what makes waiting on a phase parity unambiguous.
            if (tile_seq > 0) mbar_wait(tfull, (uint32_t)(tile_seq - 1) & 1u);
            if (prev_ddpm) mbar_wait(tempty, (uint32_t)(tile_seq - 1) & 1u);
            const int n0 = tile_n * BN;
            const int qq = tile_m / p.tiles_per_item, rr = tile_m - qq * p.tiles_per_item;
            const int c2_base = rr * p.rows_step, c3 = qq * p.items_per_tile;
            int kb0 = 0;
            if (first && q > 0 && (lp.flags & 1)) {
              const int target = G * q;
              while (ld_acquire_gpu(counter) < target) {}
              fence_proxy_async_all();
            } else if (first && q > 0) {
              // weights first (they do not depend on the previous layer), then wait for the group, then activations
              const int npre = num_kb < (int)S ? num_kb : (int)S;
              for (int kb = 0; kb < npre; ++kb) {
                const TcStage e = kb_s[kb];
                const uint32_t nw = (uint32_t)(e.src_acc >> 16) & 0xffu;
                const uint32_t par = (use_bits >> kb) & 1u;
                mbar_wait(empty0 + 8u * kb, par ^ 1u);
                const uint32_t bar = full0 + 8u * kb;
                const uint32_t sa = smem_base + (uint32_t)kb * stage_bytes;
                mbar_arrive_expect_tx(bar, TC_A_BYTES + nw * b_bytes);
                for (uint32_t j = 0; j < nw; ++j) tma_load_2d(sa + TC_A_BYTES + j * b_bytes, map_b, bar, (e.wk + (int)j) * TC_BK, n0);
              }
              const int target = G * q;
              if (ld_acquire_gpu(counter) < target) {
                const long long t0 = clock64();
                while (ld_acquire_gpu(counter) < target) {
                  if (lp.flags & 4) __nanosleep(200);
                  if (clock64() - t0 > 4000000000ll) {     // ~2 s: a lost arrival shows up as a CUDA error, not a hung GPU
                    printf("ldp_b200: planner loop group barrier timeout (block %d layer %d iteration %d)\n", blockIdx.x, l, it);
                    __trap();
                  }
                }
              }
              fence_proxy_async_all();
              if (dbg_now) lp.dbg[l * 8 + 1] = clock64();
              for (int kb = 0; kb < npre; ++kb) {
                const TcStage e = kb_s[kb];
                const int d1 = (int)(short)(e.d12 & 0xffff), d2 = e.d12 >> 16;
                tma_load_4d(smem_base + (uint32_t)kb * stage_bytes, map_a + (e.src_acc & 0xff), full0 + 8u * kb, e.c0, d1,
                            c2_base + d2, c3);
                use_bits ^= 1u << kb;
              }
              stage = (uint32_t)npre == S ? 0u : (uint32_t)npre;
              kb0 = npre;
            }
            for (int kb = kb0; kb < num_kb; ++kb) {
              const TcStage e = kb_s[kb];
              const uint32_t nw = (uint32_t)(e.src_acc >> 16) & 0xffu;
              const int d1 = (int)(short)(e.d12 & 0xffff), d2 = e.d12 >> 16;
              const uint32_t par = (use_bits >> stage) & 1u;
              mbar_wait(empty0 + 8u * stage, par ^ 1u);
              const uint32_t bar = full0 + 8u * stage;
              const uint32_t sa = smem_base + stage * stage_bytes;
              mbar_arrive_expect_tx(bar, TC_A_BYTES + nw * b_bytes);
              tma_load_4d(sa, map_a + (e.src_acc & 0xff), bar, e.c0, d1, c2_base + d2, c3);
              for (uint32_t j = 0; j < nw; ++j) tma_load_2d(sa + TC_A_BYTES + j * b_bytes, map_b, bar, (e.wk + (int)j) * TC_BK, n0);
              use_bits ^= 1u << stage;
              if (++stage == S) stage = 0;
            }
            first = false;
            prev_ddpm = p.mode == TC_EPI_DDPM;
            ++tile_seq;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t use_bits = 0;
    int tile_seq = 0;
    constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);     // SBO 1024 B, version 1, SWIZZLE_128B
    for (int it = 0; it < lp.n_steps; ++it) {
      for (int l = 0; l < lp.n_layers; ++l) {
        __syncwarp();
        if (elect_one()) {
          const TcGemm& p = c_layers[l];
          const int BN = p.block_n;
          const uint32_t b_step = ((uint32_t)BN * TC_BK * 2) >> 4;                     // descriptor units (16 B)
          const uint32_t stage_step = (TC_A_BYTES + (uint32_t)p.w_max * BN * TC_BK * 2) >> 4;
          const uint32_t S = (uint32_t)p.num_stages;
          const uint32_t idesc = BN == 64 ? umma_idesc_bf16(TC_BM, 64) : umma_idesc_bf16(TC_BM, 128);
          const int kb_main = p.kb_main > 0 ? p.kb_main : p.num_kb;
          const uint32_t nw_main = p.kb_main > 0 ? (uint32_t)p.nw_main : (uint32_t)p.w_max;
          const int kb_aux = p.num_kb - kb_main;
          const uint32_t d_aux = tmem_base + (uint32_t)p.n_acc * BN;
          const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
          const bool dbg_now = dbg_cta && it == lp.dbg_step;
          TileIter ti(p, group, r, G);
          uint32_t stage = 0;
          for (int t = r; t < ti.ntiles; t += G) {
            int tile_m, tile_n;
            if (!ti.valid(t, &tile_m, &tile_n)) continue;
            if (tile_seq > 0) {
              mbar_wait(tempty, (uint32_t)(tile_seq - 1) & 1u);        // the epilogue has drained the accumulators
              tc_fence_after();
            }
            uint32_t accf = 0;
            for (int kb = 0; kb < kb_main; ++kb) {
              const uint32_t par = (use_bits >> stage) & 1u;
              mbar_wait(full0 + 8u * stage, par);
              tc_fence_after();
              if (dbg_now && kb == 0 && t == r) lp.dbg[l * 8 + 2] = clock64();
              const uint32_t a_lo = a_lo0 + stage * stage_step;
              uint32_t b_lo = a_lo + (TC_A_BYTES >> 4);
              uint32_t d = tmem_base;
              for (uint32_t j = 0; j < nw_main; ++j, b_lo += b_step, d += BN) {
                umma_lohi(d, a_lo, b_lo, DESC_HI, idesc, accf);
                umma_lohi(d, a_lo + 2, b_lo + 2, DESC_HI, idesc, 1u);
                umma_lohi(d, a_lo + 4, b_lo + 4, DESC_HI, idesc, 1u);
                umma_lohi(d, a_lo + 6, b_lo + 6, DESC_HI, idesc, 1u);
              }
              umma_commit(empty0 + 8u * stage);
              accf = 1u;
              use_bits ^= 1u << stage;
              if (++stage == S) stage = 0;
            }
            accf = 0;
            for (int kb = 0; kb < kb_aux; ++kb) {
              const uint32_t par = (use_bits >> stage) & 1u;
              mbar_wait(full0 + 8u * stage, par);
              tc_fence_after();
              const uint32_t a_lo = a_lo0 + stage * stage_step;
              const uint32_t b_lo = a_lo + (TC_A_BYTES >> 4);
              umma_lohi(d_aux, a_lo, b_lo, DESC_HI, idesc, accf);
              umma_lohi(d_aux, a_lo + 2, b_lo + 2, DESC_HI, idesc, 1u);
              umma_lohi(d_aux, a_lo + 4, b_lo + 4, DESC_HI, idesc, 1u);
              umma_lohi(d_aux, a_lo + 6, b_lo + 6, DESC_HI, idesc, 1u);
              umma_commit(empty0 + 8u * stage);
              accf = 1u;
              use_bits ^= 1u << stage;
              if (++stage == S) stage = 0;
            }
            umma_commit(tfull);
            if (dbg_now && t == r) lp.dbg[l * 8 + 3] = clock64();
            ++tile_seq;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    int tile_seq = 0;
    for (int it = 0; it < lp.n_steps; ++it) {
      const int timestep = lp.t_first - it;
      for (int l = 0; l < lp.n_layers; ++l) {
        epi_all_bar();                                   // every epilogue warp is done with the previous layer (`es` is free)
        const TcGemm& p = c_layers[l];
        const int BN = p.block_n;
        const bool active = ew < (BN == 64 ? 8 : 16);
        const bool dbg_now = dbg_cta && it == lp.dbg_step && threadIdx.x == 64;
        TileIter ti(p, group, r, G);
        bool had_tile = false;
        for (int t = r; t < ti.ntiles; t += G) {
          int tile_m, tile_n;
          if (!ti.valid(t, &tile_m, &tile_n)) continue;
          had_tile = true;
          const uint32_t par = (uint32_t)tile_seq & 1u;
          if (active) {
            const int n0 = tile_n * BN;
            if (BN == 64) {
              if (p.mode == TC_EPI_GN) epilogue_tile<64, TC_EPI_GN>(l, es_raw, tmem_base, ring, tile_m, n0, warp, lane, tfull, par, lp.flags, timestep, (dbg_cta && it == lp.dbg_step && t == r) ? lp.dbg + lp.n_layers * 8 : nullptr);
              else epilogue_tile<64, TC_EPI_PLAIN>(l, es_raw, tmem_base, ring, tile_m, n0, warp, lane, tfull, par, lp.flags, timestep, (dbg_cta && it == lp.dbg_step && t == r) ? lp.dbg + lp.n_layers * 8 : nullptr);
            } else {
              if (p.mode == TC_EPI_GN) epilogue_tile<128, TC_EPI_GN>(l, es_raw, tmem_base, ring, tile_m, n0, warp, lane, tfull, par, lp.flags, timestep, (dbg_cta && it == lp.dbg_step && t == r) ? lp.dbg + lp.n_layers * 8 : nullptr);
              else if (p.mode == TC_EPI_DDPM) epilogue_tile<128, TC_EPI_DDPM>(l, es_raw, tmem_base, ring, tile_m, n0, warp, lane, tfull, par, lp.flags, timestep, (dbg_cta && it == lp.dbg_step && t == r) ? lp.dbg + lp.n_layers * 8 : nullptr);
              else epilogue_tile<128, TC_EPI_PLAIN>(l, es_raw, tmem_base, ring, tile_m, n0, warp, lane, tfull, par, lp.flags, timestep, (dbg_cta && it == lp.dbg_step && t == r) ? lp.dbg + lp.n_layers * 8 : nullptr);
            }
            if (threadIdx.x == 64) {
              mbar_arrive(tempty);
              if (dbg_now && t == r) lp.dbg[l * 8 + 4] = clock64();
            }
          }
          ++tile_seq;
        }
        // this CTA's part of the layer is written: tell the group (the bar.sync inside epilogue_tile ordered every
        // epilogue thread's stores before this thread; the release makes them visible device-wide, cumulatively)
        if (threadIdx.x == 64) {
          if (dbg_now) lp.dbg[l * 8 + 5] = clock64();
          // The counter is a running sum, so no CTA may arrive for layer l before the whole group has arrived for layer
          // l-1.  With a tile that is implied (its operands waited for the counter); without one, wait here.
          if (!had_tile) {
            const int target = G * (it * lp.n_layers + l);
            const long long t0 = clock64();
            while (ld_acquire_gpu(counter) < target) {
              if (clock64() - t0 > 4000000000ll) {
                printf("ldp_b200: planner loop idle-CTA barrier timeout (block %d layer %d iteration %d)\n", blockIdx.x, l, it);
                __trap();
              }
            }
          }
          fence_proxy_async_all();
          red_release_gpu_add(counter, 1);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int upload_planner_loop_layers(const TcGemm* layers_host, int n_layers, cudaStream_t s) {
  LDP_CHECK(n_layers > 0 && n_layers <= PL_MAX_LAYERS, LDP_ERR_UNSUPPORTED, "planner loop: too many layers for the constant table");
  LDP_CUDA_OK(cudaMemcpyToSymbolAsync(c_layers_raw, layers_host, (size_t)n_layers * sizeof(TcGemm), 0, cudaMemcpyHostToDevice, s));
  return LDP_OK;
}

int launch_planner_loop(const PlannerLoop& lp, int n_groups, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    LDP_CUDA_OK(cudaFuncSetAttribute(planner_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_RING + 1024));
    attr_done = true;
  }
  LDP_CHECK(lp.n_layers > 0 && lp.n_layers <= PL_MAX_LAYERS && lp.n_steps > 0 && lp.group_counter && lp.group_ctas > 0 && n_groups > 0,
            LDP_ERR_INVALID_ARG, "planner loop: bad arguments");
  planner_loop_kernel<<<dim3(n_groups * lp.group_ctas), dim3(PL_THREADS), TC_SMEM_RING + 1024, s>>>(lp);
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, planner_loop_kernel);
      set_last_error(std::string("planner loop launch failed: ") + cudaGetErrorString(e) + " (regs " + std::to_string(fa.numRegs) +
                     ", static smem " + std::to_string(fa.sharedSizeBytes) + ", max threads " + std::to_string(fa.maxThreadsPerBlock) +
                     ", local " + std::to_string(fa.localSizeBytes) + ")");
      return LDP_ERR_CUDA;
    }
  }
  count_launch();
  return LDP_OK;
}

}  // namespace ldp
