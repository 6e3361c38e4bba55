// Microbenchmark (diagnostics, not product): what main-loop rate can a tcgen05 pipeline sustain on one SM as a function
// of the MMA shape, the operand sources (A from shared memory or TMEM), CTA pairs, and the TMA bytes written into
// shared memory per stage?  Mirrors the ring of tc_gemm.cu (TMA producer thread, MMA issuer thread, mbarrier ring).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_bin/mma_ubench scripts/mma_ubench.cu
//   scripts/_bin/mma_ubench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../latent_diffusion_planning_b200/csrc/common.cuh"

using namespace ldp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Cfg {
  CUtensorMap map16;   // box {64, 128}: 16 KB
  CUtensorMap map8;    // box {64, 64}: 8 KB
  int iters;           // stages processed per CTA
  int a_loads;         // 16 KB A loads per stage (0/1)
  int b_tiles;         // B tiles loaded per stage (each N x 64 bf16; pair: N/2 x 64 per CTA)
  int groups;          // groups of 4 MMAs (K = 64) per stage; group j uses B tile j % b_tiles, accumulator j % n_acc
  int n_acc;
  int stages;          // ring depth
  int stage_bytes;
  int rows_total;      // rows of the global source matrix
  int no_load;         // producer only arrives on the full barrier (no TMA traffic)
  int free_run;        // MMA issuer does not wait for the full barriers
  int commit_every;    // commit to the empty barriers of the last n stages only every n stages
  int pure;            // no producer, no waits: 1 = MMAs only, |2 = fence::after per stage, |4 = commit per stage (nobody waits)
  int a_cp;            // A from TMEM: issue tcgen05.cp smem->TMEM per stage (4 x 128x256b)
  long long* out;      // [ctas][2]: cycles of the main loop, 0
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b_2sm(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

template <int N, bool PAIR, bool ATMEM>
__global__ void __launch_bounds__(128, 1) ubench(const __grid_constant__ Cfg p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_done;
  __shared__ uint32_t tmem_holder;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  constexpr int B_BYTES = (PAIR ? N / 2 : N) * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_done), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_2sm(smem_u32(&tmem_holder), 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(smem_u32(&tmem_holder), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) { __syncwarp(); cluster_sync_all(); } else { __syncthreads(); }
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const uint32_t tx = p.no_load ? 0u : (uint32_t)p.a_loads * 16384u + (uint32_t)p.b_tiles * B_BYTES;

  if (warp == 0 && !p.pure) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      int row = (blockIdx.x * 977) % (p.rows_total - 4096);
      for (int it = 0; it < p.iters; ++it) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        const uint32_t bar = smem_u32(&bar_full[stage]);
        const uint32_t sa = smem_base + stage * p.stage_bytes;
        if (PAIR) {
          const uint32_t bl = mapa_shared(bar, 0);
          if (leader && tx) mbar_arrive_expect_tx(bar, 2u * tx);
          if (p.a_loads && tx) tma_load_2d_2sm(sa, &p.map16, bl, 0, row);
          for (int j = 0; j < (tx ? p.b_tiles : 0); ++j) {
            if (N == 256) tma_load_2d_2sm(sa + 16384 + j * B_BYTES, &p.map16, bl, 0, row + 128 * (j + 1));
            else if (N == 128) tma_load_2d_2sm(sa + 16384 + j * B_BYTES, &p.map8, bl, 0, row + 128 * (j + 1));
          }
          if (leader && !tx) mbar_arrive(bar);
          if (!leader && !tx) {}
        } else {
          if (tx) mbar_arrive_expect_tx(bar, tx); else mbar_arrive(bar);
          if (p.a_loads && tx) tma_load_2d(sa, &p.map16, bar, 0, row);
          for (int j = 0; j < (tx ? p.b_tiles : 0); ++j)
            for (int h = 0; h < N / 128; ++h)
              tma_load_2d(sa + 16384 + j * B_BYTES + h * 16384, &p.map16, bar, 0, row + 128 * (j * 2 + h + 1));
          if (N == 64)
            for (int j = 0; j < (tx ? p.b_tiles : 0); ++j) tma_load_2d(sa + 16384 + j * B_BYTES, &p.map8, bar, 0, row + 128 * (j + 1));
        }
        row += 1024;
        if (row >= p.rows_total - 4096) row -= (p.rows_total - 4096);
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, N);
      uint32_t stage = 0, phase = 0;
      long long t0 = 0;
      for (int it = 0; it < p.iters; ++it) {
        if (!p.free_run && !p.pure) mbar_wait(smem_u32(&bar_full[stage]), phase);
        if (!p.pure || (p.pure & 2)) tc_fence_after();
        if (it == 0) t0 = clock64();
        const uint32_t sa = smem_base + stage * p.stage_bytes;
        const uint64_t da = umma_desc_sw128(sa);
        const uint32_t a_t = tmem_base + 384 + (it & 1) * 32;        // A slices in TMEM: 8 columns per K=16 step
        if (ATMEM && p.a_cp) {
          for (int k = 0; k < 4; ++k) {
            if (PAIR) tmem_cp_128x256b_2sm(a_t + k * 8, da + 2 * k);
            else tmem_cp_128x256b(a_t + k * 8, da + 2 * k);
          }
        }
        {
          uint64_t db = umma_desc_sw128(sa + 16384);
          uint32_t d_tmem = tmem_base;
          int bt = 0, ac = 0;
          for (int g = 0; g < p.groups; ++g) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (ATMEM) {
                if (PAIR) umma_bf16_ts_2sm(d_tmem, a_t + k * 8, db + 2 * k, idesc, 1u);
                else umma_bf16_ts(d_tmem, a_t + k * 8, db + 2 * k, idesc, 1u);
              } else {
                if (PAIR) umma_bf16_ss_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                else umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, 1u);
              }
            }
            db += B_BYTES >> 4; d_tmem += N;
            if (++bt == p.b_tiles) { bt = 0; db = umma_desc_sw128(sa + 16384); }
            if (++ac == p.n_acc) { ac = 0; d_tmem = tmem_base; }
          }
        }
        if (p.pure && !(p.pure & 4)) {
        } else if (p.commit_every <= 1) {
          if (PAIR) umma_commit_2sm(smem_u32(&bar_empty[stage]), 3);
          else umma_commit(smem_u32(&bar_empty[stage]));
        } else if ((it + 1) % p.commit_every == 0) {
          for (int b = 0; b < p.commit_every; ++b) {
            const uint32_t st = (stage + p.stages - b) % p.stages;
            if (PAIR) umma_commit_2sm(smem_u32(&bar_empty[st]), 3);
            else umma_commit(smem_u32(&bar_empty[st]));
          }
        }
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
      }
      if (PAIR) umma_commit_2sm(smem_u32(&bar_done), 1); else umma_commit(smem_u32(&bar_done));
      mbar_wait(smem_u32(&bar_done), 0);
      const long long t1 = clock64();
      p.out[blockIdx.x * 2] = t1 - t0;
    }
  }
  tc_fence_before();
  if (PAIR) { __syncwarp(); cluster_sync_all(); if (warp == 1) tmem_dealloc_2sm(tmem_base, 512); }
  else { __syncthreads(); if (warp == 1) tmem_dealloc(tmem_base, 512); }
}


// Completion timeline of a pure MMA stream (no loads, no waits): group g = 4 MMAs (K = 64) committed to its own barrier;
// an observer thread records when each group completes.  Shows any start-up ramp of the tensor pipe.
template <int N>
__global__ void __launch_bounds__(128, 1) ramp(long long* out, int groups, int spin_before) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[64];
  __shared__ uint32_t tmem_holder;
  __shared__ long long t0s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 64; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_mbar_init();
    t0s = 0;
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_holder), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  if (warp == 1 && elect_one()) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    long long t = clock64();
    while (clock64() - t < spin_before) {}
    *(volatile long long*)&t0s = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint64_t da = umma_desc_sw128(smem_base + (g & 3) * 49152);
      const uint64_t db = umma_desc_sw128(smem_base + (g & 3) * 49152 + 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + (g & 1) * N, da + 2 * k, db + 2 * k, idesc, 1u);
      if (g < 64) umma_commit(smem_u32(&bars[g]));
    }
  } else if (warp == 2 && elect_one()) {
    while (*(volatile long long*)&t0s == 0) {}
    const long long t0 = *(volatile long long*)&t0s;
    for (int g = 0; g < 64 && g < groups; ++g) {
      mbar_wait(smem_u32(&bars[g]), 0);
      out[blockIdx.x * 64 + g] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int N>
static void run_ramp(int grid, int groups, int spin, long long* dev) {
  CK(cudaFuncSetAttribute(ramp<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaMemset(dev, 0, 148 * 64 * 8));
  for (int rep = 0; rep < 2; ++rep) ramp<N><<<grid, 128, 200 * 1024>>>(dev, groups, spin);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148 * 64);
  CK(cudaMemcpy(h.data(), dev, h.size() * 8, cudaMemcpyDeviceToHost));
  printf("ramp N=%d grid=%d groups=%d spin=%d: completion clocks of groups (CTA 0):", N, grid, groups, spin);
  for (int g = 0; g < 64 && g < groups; ++g) printf(" %lld", h[g]);
  printf("\n   last CTA:");
  for (int g = 0; g < 64 && g < groups; g += 4) printf(" %lld", h[(grid - 1) * 64 + g]);
  printf("\n");
}

// Queue-depth probe: groups of `gsize` MMAs (4 per accumulator switch) separated by a busy-wait of `gap` clocks in the
// issuing thread.  If the tensor pipe queues issued MMAs, time per group = max(gap + issue, gsize * T_mma); if issue is
// synchronous with execution, time per group = gap + gsize * T_mma.
template <int N, bool PAIR>
__global__ void __launch_bounds__(128, 1) qprobe(long long* out, int groups, int gsize, int gap, int same_acc, int gap_pos = -1) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_done, bar_ready, bar_scratch[8];
  __shared__ uint32_t tmem_holder;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const bool leader = !PAIR || cluster_ctarank() == 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar_done), 1); mbar_init(smem_u32(&bar_ready), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar_scratch[i]), 1); fence_mbar_init(); mbar_arrive(smem_u32(&bar_ready)); }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_2sm(smem_u32(&tmem_holder), 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(smem_u32(&tmem_holder), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) { __syncwarp(); cluster_sync_all(); } else { __syncthreads(); }
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  if (warp == 1 && leader && elect_one()) {
    constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, N);
    const uint64_t da = umma_desc_sw128(smem_base);
    const uint64_t db = umma_desc_sw128(smem_base + 16384);
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (gap_pos == -2 || gap_pos <= -4) mbar_wait(smem_u32(&bar_ready), 0);
      if (gap_pos == -3 || gap_pos <= -4) tc_fence_after();
      for (int i = 0; i < gsize; ++i) {
        const uint32_t d = tmem_base + (same_acc ? 0 : ((g + (i >> 2)) & 1) * N);
        if (PAIR) umma_bf16_ss_2sm(d, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1u);
        else umma_bf16_ss(d, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1u);
        if (i == gap_pos) { const long long t = clock64(); while (clock64() - t < gap) {} }
      }
      if (gap_pos == -5) { if (PAIR) umma_commit_2sm(smem_u32(&bar_scratch[g & 7]), 1); else umma_commit(smem_u32(&bar_scratch[g & 7])); }
      if (gap > 0 && gap_pos == -1) { const long long t = clock64(); while (clock64() - t < gap) {} }
    }
    const long long t1 = clock64();
    if (PAIR) umma_commit_2sm(smem_u32(&bar_done), 1); else umma_commit(smem_u32(&bar_done));
    mbar_wait(smem_u32(&bar_done), 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  if (PAIR) { __syncwarp(); cluster_sync_all(); if (warp == 1) tmem_dealloc_2sm(tmem_base, 512); }
  else { __syncthreads(); if (warp == 1) tmem_dealloc(tmem_base, 512); }
}

template <int N, bool PAIR>
static void run_qprobe(int grid, int groups, int gsize, int gap, int same_acc, long long* dev, int gap_pos = -1) {
  CK(cudaFuncSetAttribute(qprobe<N, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(128, 1, 1); cfg.dynamicSmemBytes = 100 * 1024;
  cudaLaunchAttribute attr[1]; int na = 0;
  if (PAIR) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; ++na; }
  cfg.attrs = attr; cfg.numAttrs = na;
  CK(cudaMemset(dev, 0, 148 * 16));
  for (int rep = 0; rep < 2; ++rep) CK(cudaLaunchKernelEx(&cfg, qprobe<N, PAIR>, dev, groups, gsize, gap, same_acc, gap_pos));
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, dev, 16, cudaMemcpyDeviceToHost));
  printf("qprobe pos=%2d N=%d pair=%d gsize=%2d gap=%4d same_acc=%d: issue %7.1f clk/group, complete %7.1f clk/group (MMA floor %d)\n", gap_pos, N, (int)PAIR,
         gsize, gap, same_acc, (double)h[0] / groups, (double)h[1] / groups, gsize * N / 2);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled enc;
static void make_map(CUtensorMap* m, void* base, int rows, int box_rows) {
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t str[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

template <int N, bool PAIR, bool ATMEM>
static void run(const char* name, Cfg c, int grid) {
  constexpr int B_BYTES = (PAIR ? N / 2 : N) * 128;
  c.stage_bytes = 16384 + c.b_tiles * B_BYTES;
  if (c.stage_bytes < 16384 + B_BYTES) c.stage_bytes = 16384 + B_BYTES;
  c.stages = std::min(8, (196 * 1024) / c.stage_bytes);
  const int smem = c.stages * c.stage_bytes + 1024;
  CK(cudaFuncSetAttribute(ubench<N, PAIR, ATMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(128, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  CK(cudaMemset(c.out, 0, grid * 16));
  for (int rep = 0; rep < 2; ++rep) CK(cudaLaunchKernelEx(&cfg, ubench<N, PAIR, ATMEM>, c));
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(grid * 2);
  CK(cudaMemcpy(h.data(), c.out, grid * 16, cudaMemcpyDeviceToHost));
  double sum = 0, mx = 0; int n = 0;
  for (int i = 0; i < grid; ++i) if (h[2 * i] > 0) { sum += h[2 * i]; mx = std::max(mx, (double)h[2 * i]); ++n; }
  const double cyc = sum / n;
  const double mmas = (double)c.iters * c.groups * 4;
  const double floor_clk = (N / 128.0) * 64.0;                 // per MMA (per SM) at 8192 flop/clk/SM
  const double tma_bytes = (double)c.iters * (c.a_loads * 16384.0 + c.b_tiles * (double)B_BYTES);
  printf("%-46s N=%3d pair=%d atmem=%d grid=%3d stage=%3dKB x%d groups=%d: %7.1f clk/MMA (floor %5.1f, %5.1f%% of peak)  TMA %5.1f B/clk/SM  max/mean %.2f\n",
         name, N, (int)PAIR, (int)ATMEM, grid, c.stage_bytes / 1024, c.stages, c.groups, cyc / mmas, floor_clk,
         100.0 * floor_clk / (cyc / mmas), tma_bytes / cyc, mx / cyc);
}

int main(int argc, char** argv) {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
  enc = (PFN_encodeTiled)ptr;
  const int rows = 1 << 18;                      // 256k rows x 128 B = 32 MB (L2 resident)
  void* src;
  CK(cudaMalloc(&src, (size_t)rows * 128));
  {
    std::vector<uint16_t> h((size_t)rows * 64);
    uint32_t s = 12345;
    for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (uint16_t)(0x3c00 + ((s >> 20) & 0x1ff)) | ((s >> 31) << 15); }   // ~ +-1..2 bf16
    CK(cudaMemcpy(src, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  }
  Cfg c = {};
  make_map(&c.map16, src, rows, 128);
  make_map(&c.map8, src, rows, 64);
  c.rows_total = rows;
  CK(cudaMalloc(&c.out, 4096));
  c.iters = 64;
  const int G = argc > 1 ? atoi(argv[1]) : 128;
  if (argc > 2) {
    long long* dev;
    CK(cudaMalloc(&dev, 148 * 64 * 8));
    for (int pos : {-1, -2, -3, -4, -5}) {
      run_qprobe<128, false>(G, 128, 4, 0, 1, dev, pos);
      run_qprobe<128, true>(G, 128, 4, 0, 1, dev, pos);
      run_qprobe<256, false>(G, 128, 4, 0, 1, dev, pos);
      run_qprobe<128, false>(G, 128, 12, 0, 0, dev, pos);
    }
    run_ramp<128>(1, 64, 0, dev);
    run_ramp<128>(128, 64, 0, dev);
    run_ramp<128>(128, 256, 0, dev);
    run_ramp<256>(128, 64, 0, dev);
    run_ramp<128>(128, 64, 100000, dev);
    run_ramp<128>(148, 64, 0, dev);
    return 0;
  }
  auto set = [&](int a_loads, int b_tiles, int groups, int n_acc, int a_cp) { c.a_loads = a_loads; c.b_tiles = b_tiles; c.groups = groups; c.n_acc = n_acc; c.a_cp = a_cp; };
  // SS, single CTA
  set(1, 1, 1, 1, 0); run<128, false, false>("per-tap: A16+B16 per 4 MMA", c, G);
  set(1, 3, 3, 3, 0); run<128, false, false>("tapacc3: A16+3xB16 per 12 MMA", c, G);
  set(1, 1, 4, 1, 0); run<128, false, false>("light TMA: A16+B16 per 16 MMA", c, G);
  set(1, 1, 16, 1, 0); run<128, false, false>("very light TMA: 32KB per 64 MMA", c, G);
  set(1, 1, 1, 1, 0); run<256, false, false>("N256 per-tap: A16+B32 per 4 MMA", c, G);
  set(1, 1, 16, 1, 0); run<256, false, false>("N256 very light TMA", c, G);
  set(1, 1, 1, 1, 0); run<64, false, false>("N64 per-tap: A16+B8 per 4 MMA", c, G);
  set(1, 3, 3, 3, 0); run<64, false, false>("N64 tapacc3", c, G);
  set(1, 1, 16, 1, 0); run<64, false, false>("N64 very light TMA", c, G);
  // TS (A in TMEM), single CTA
  set(1, 1, 1, 1, 0); run<128, false, true>("TS per-tap (no cp): A16+B16 per 4 MMA", c, G);
  set(1, 1, 1, 1, 1); run<128, false, true>("TS per-tap (+cp)", c, G);
  set(1, 3, 3, 3, 1); run<128, false, true>("TS tapacc3 (+cp)", c, G);
  set(1, 3, 3, 3, 0); run<128, false, true>("TS tapacc3 (no cp)", c, G);
  set(1, 1, 16, 1, 0); run<128, false, true>("TS very light TMA", c, G);
  set(1, 1, 1, 1, 1); run<256, false, true>("TS N256 per-tap (+cp)", c, G);
  set(1, 1, 16, 1, 0); run<256, false, true>("TS N256 very light TMA", c, G);
  // where does the per-stage bubble come from?
  c.no_load = 1;
  set(1, 1, 1, 1, 0); run<128, false, false>("NO LOAD per-tap (4 MMA/stage)", c, G);
  set(1, 3, 3, 3, 0); run<128, false, false>("NO LOAD tapacc3 (12 MMA/stage)", c, G);
  set(1, 1, 1, 1, 0); run<256, false, false>("NO LOAD N256 (4 MMA/stage)", c, G);
  set(1, 1, 1, 1, 0); run<128, true, false>("NO LOAD pair per-tap", c, G);
  c.no_load = 0;
  for (int pure : {1, 3, 5, 7}) {
    c.pure = pure;
    char nm[64];
    snprintf(nm, 64, "PURE mode %d per-tap", pure);
    set(1, 1, 1, 1, 0); run<128, false, false>(nm, c, G);
    snprintf(nm, 64, "PURE mode %d N256", pure);
    set(1, 1, 1, 1, 0); run<256, false, false>(nm, c, G);
    snprintf(nm, 64, "PURE mode %d pair N128", pure);
    set(1, 1, 1, 1, 0); run<128, true, false>(nm, c, G);
  }
  c.pure = 0;
  c.commit_every = 2;
  set(1, 1, 1, 1, 0); run<128, false, false>("COMMIT/2 per-tap", c, G);
  set(1, 1, 1, 1, 0); run<256, false, false>("COMMIT/2 N256", c, G);
  c.commit_every = 3;
  set(1, 1, 1, 1, 0); run<128, false, false>("COMMIT/3 per-tap", c, G);
  c.commit_every = 0;
  // pairs
  set(1, 1, 1, 1, 0); run<128, true, false>("pair per-tap: A16+B8 per 4 MMA", c, G);
  set(1, 3, 3, 3, 0); run<128, true, false>("pair tapacc3: A16+3xB8 per 12 MMA", c, G);
  set(1, 1, 16, 1, 0); run<128, true, false>("pair very light TMA", c, G);
  set(1, 1, 1, 1, 0); run<256, true, false>("pair N256: A16+B16 per 4 MMA", c, G);
  set(1, 1, 16, 1, 0); run<256, true, false>("pair N256 very light TMA", c, G);
  set(1, 3, 3, 3, 1); run<128, true, true>("pair TS tapacc3 (+cp)", c, G);
  set(1, 1, 1, 1, 1); run<256, true, true>("pair TS N256 (+cp)", c, G);
  set(1, 1, 16, 1, 0); run<256, true, true>("pair TS N256 very light", c, G);
  printf("done\n");
  return 0;
}
