// Shared device/host helpers for libldp_b200: error plumbing, sm_100a PTX wrappers (mbarrier, TMA,
// tcgen05/TMEM), and the small math pieces every kernel uses (Mish, Philox4x32-10 + Box-Muller).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/ldp_b200.h"

namespace ldp {

// ------------------------------------------------------------------------------------------------
// error handling: no exceptions cross the C ABI; the last message is kept per thread.
// ------------------------------------------------------------------------------------------------
void set_last_error(const std::string& msg);
const char* get_last_error();

#define LDP_CUDA_OK(expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::ldp::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +        \
                            __FILE__ + ":" + std::to_string(__LINE__) + ")");                  \
      return LDP_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define LDP_CHECK(cond, code, msg)                                                             \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      ::ldp::set_last_error(std::string(msg) + " [" #cond "] (" + __FILE__ + ":" +             \
                            std::to_string(__LINE__) + ")");                                   \
      return (code);                                                                           \
    }                                                                                          \
  } while (0)

#define LDP_TRY(expr)                                                                          \
  do {                                                                                         \
    int _s = (expr);                                                                           \
    if (_s != LDP_OK) return _s;                                                               \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ------------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------------
// Mish(x) = x tanh(softplus(x)) (reference networks/diffusion_nets_v2.py:11-14).
// tanh(log(1+e^x)) = n/(n+2) with n = e^x (e^x + 2): one exp, no cancellation for x << 0.
template <bool kFast>
__device__ __forceinline__ float mish_f(float x) {
  if (x > 20.f) return x;
  float e = kFast ? __expf(x) : expf(x);
  float n = e * (e + 2.f);
  return kFast ? x * __fdividef(n, n + 2.f) : x * (n / (n + 2.f));
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }

// Philox4x32-10 (counter-based; results independent of launch geometry and of how rows are sharded
// over GPUs).  Mirrors oracle/ldp_oracle.py:philox4x32.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    // one IMAD.WIDE per product (hi and lo words together) instead of IMAD.HI + IMAD
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c.x, p1 = (unsigned long long)0xCD9E8D57u * c.z;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// Standard normal for flat element index e of (stream, step): quad q = e/4 shares one Philox call;
// lanes 0/1 are the cos/sin of the first Box-Muller pair, 2/3 of the second.
__device__ __forceinline__ float philox_normal(unsigned long long seed, uint32_t stream, uint32_t step,
                                               unsigned long long e) {
  unsigned long long q = e >> 2;
  uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), step, stream),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t lane = (uint32_t)(e & 3ull);
  uint32_t ra = lane < 2 ? r.x : r.z, rb = lane < 2 ? r.y : r.w;
  float u0 = ((float)ra + 0.5f) * 2.3283064365386963e-10f;   // (r+0.5) 2^-32
  float u1 = ((float)rb + 0.5f) * 2.3283064365386963e-10f;
  // (float)ra may round up to 2^32 -> u0 == 1 -> log = 0 -> z = 0 (harmless); never 0 -> no inf.
  float rad = sqrtf(-2.f * logf(u0));
  float s, c;
  sincospif(2.f * u1, &s, &c);
  return rad * ((lane & 1u) ? s : c);
}

// Row-structured noise of the sampling loops: the four normals of Philox quad (quad, row) at (step, stream), i.e.
// columns 4*quad .. 4*quad+3 of global row `row`.  MUFU-based Box-Muller (lg2 / rsqrt / sin / cos approximations:
// absolute error ~1e-6, irrelevant for noise but ~8x fewer instructions than logf/sincospif).
// Mirrors oracle/ldp_oracle.py:philox_normal_rows.
__device__ __forceinline__ void philox_normal4_rows(unsigned long long seed, uint32_t stream, uint32_t step, uint32_t row,
                                                    uint32_t quad, float* z) {
  const uint4 r = philox4x32_10(make_uint4(quad, row, step, stream), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;
  const float u0 = ((float)r.x + 0.5f) * k, u1 = ((float)r.y + 0.5f) * k;
  const float u2 = ((float)r.z + 0.5f) * k, u3 = ((float)r.w + 0.5f) * k;
  const float ra = sqrtf(fmaxf(-1.3862943611198906f * __log2f(u0), 0.f));   // sqrt(-2 ln u) = sqrt(-2 ln2 log2 u)
  const float rb = sqrtf(fmaxf(-1.3862943611198906f * __log2f(u2), 0.f));
  float s, c;
  __sincosf(6.283185307179586f * u1, &s, &c);
  z[0] = ra * c; z[1] = ra * s;
  __sincosf(6.283185307179586f * u3, &s, &c);
  z[2] = rb * c; z[3] = rb * s;
}

// ------------------------------------------------------------------------------------------------
// sm_100a PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a wrong descriptor / byte count shows up as a trap (-> CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {   // ~2 s at 1.9 GHz
      printf("ldp_b200: mbarrier wait timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// --- TMEM / tcgen05 ---
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base+i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory operand descriptor, K-major tile of [rows][64] bf16 (128-byte rows), SWIZZLE_128B:
// start address >> 4 | LBO (ignored for swizzled K-major; 1) << 16 | SBO = 1024 B (8 rows) >> 4 << 32 |
// version 1 << 46 | layout SWIZZLE_128B (2) << 61.   (cf. cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16: D fp32, A/B bf16, both K-major, M x N tile (cf. cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// --- CTA pair (cta_group::2): two CTAs of a cluster run one M=256 MMA; each holds its own 128 rows of A, half of the
// B tile, and the accumulator of its own rows.  Only the leader (cluster rank 0) issues MMAs and owns the "full" barriers.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Cluster barrier without memory ordering: enough to keep both CTAs of a pair alive until neither touches the other's
// shared memory / TMEM any more.  (arrive.release makes every thread wait for its outstanding global stores first - at
// the end of a kernel that is ~1k cycles of MEMBAR for ordering that kernel completion provides anyway.)
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier of this CTA whose arrivals come from both CTAs of the pair
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000ll) {
      printf("ldp_b200: cluster mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const void* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const void* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs of this thread completed) on the barrier at the same shared-memory offset in
// every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace ldp
