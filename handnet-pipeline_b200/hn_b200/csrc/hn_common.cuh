// Shared device/host helpers for libhandnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../../include/handnet_b200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void hn_set_error(const char* fmt, ...);

#define HN_CHECK_CUDA(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      hn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
      return HN_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define HN_REQUIRE(cond, ...)                                                                  \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      hn_set_error(__VA_ARGS__);                                                               \
      return HN_ERR_ARG;                                                                       \
    }                                                                                          \
  } while (0)

#define HN_LAUNCH_CHECK() HN_CHECK_CUDA(cudaGetLastError())

static inline int hn_div_up(int a, int b) { return (a + b - 1) / b; }
int hn_num_sms();
void hn_count_launch();

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t hn_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t hn_pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float hn_bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float hn_bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ float hn_round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// 32-byte global accesses (sm_100: LDG/STG.256): one full sector per lane and instruction.  p must be 32-byte aligned.
__device__ __forceinline__ void hn_ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void hn_stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ float4 hn_lds128(uint32_t smem_addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr));
  return v;
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void hn_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hn_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hn_mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void hn_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hn_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void hn_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hn_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool hn_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(hn_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug turns into a trap (reported as a CUDA error) instead of a hang.  The report is kept
// out of line so that the polling loop stays a handful of instructions.
static __device__ __noinline__ void hn_mbar_timeout(const void* bar, uint32_t parity) {
  printf("hn: mbarrier wait timed out (block %d thread %d bar %p parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar,
         parity);
  __trap();
}
__device__ __forceinline__ void hn_mbar_wait(uint64_t* bar, uint32_t parity) {
  if (hn_mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!hn_mbar_try_wait(bar, parity)) {
    // time-based bound (~2 s of SM clock): a try_wait may suspend the thread for a while, so a spin count alone says little
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) hn_mbar_timeout(bar, parity);
    }
  }
}
// One lane of a CONVERGED warp (all 32 lanes must execute this).  Code guarded by the result is known to the
// compiler to run in a single thread, so tcgen05 / TMA operands go to uniform registers without a per-lane
// "waterfall" loop around every instruction (what `if (lane == 0)` compiles to).
__device__ __forceinline__ bool hn_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMA ----------------------------------------------------------------------------------
__device__ __forceinline__ void hn_tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void hn_tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(hn_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(hn_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void hn_tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(hn_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(hn_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void hn_tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(hn_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(hn_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}

__device__ __forceinline__ void hn_tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(hn_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(hn_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ void hn_tma_load_2d_mcast(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                     uint16_t cta_mask) {
  // the box lands at the same CTA-relative smem offset in every CTA of cta_mask and signals the mbarrier at the
  // same offset there
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(hn_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(hn_smem_u32(bar)), "r"(c0), "r"(c1),
        "h"(cta_mask)
      : "memory");
}

// ---- clusters -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hn_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void hn_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------
__device__ __forceinline__ void hn_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hn_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void hn_tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hn_smem_u32(smem_dst)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void hn_tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread for the CTA.
__device__ __forceinline__ void hn_umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void hn_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hn_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void hn_umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(hn_smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// Four K=16 steps over one 64-wide (128-byte, SWIZZLE_128B) K block: the start-address field advances by 32 bytes
// (2 in descriptor units) per step.  `accumulate` = 0 overwrites the accumulator with the first product.
__device__ __forceinline__ void hn_umma_bf16_x4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %5, %6, %3, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %7, %8, %3, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %9, %10, %3, 1;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "l"(desc_a + 2), "l"(desc_b + 2),
        "l"(desc_a + 4), "l"(desc_b + 4), "l"(desc_a + 6), "l"(desc_b + 6)
      : "memory");
}
// tcgen05.commit on an mbarrier given by its shared-memory address; CS > 1 signals the barrier at the same offset in
// every CTA of the cluster.
template <int CS>
__device__ __forceinline__ void hn_umma_commit_addr(uint32_t bar_addr) {
  if constexpr (CS == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar_addr), "h"((uint16_t)((1u << CS) - 1u)) : "memory");
  }
}

// ---- CTA pairs (tcgen05 cta_group::2) -------------------------------------------------------------------------------
// A pair = the two CTAs of a cluster of 2 (same TPC).  The leader (cluster rank 0) issues tcgen05.mma.cta_group::2 for
// both: M = 256 (each CTA's own 128 rows of A), B split in halves (each CTA's shared memory holds N/2 rows of the weight
// tile), accumulators in each CTA's own TMEM.  Operand loads of BOTH CTAs signal the leader's mbarrier.
constexpr uint32_t HN_PEER_BIT_MASK = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address (-> rank 0)

template <int COLS>
__device__ __forceinline__ void hn_tmem_alloc_pair(uint32_t* smem_dst) {      // one warp of EACH CTA, same smem offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hn_smem_u32(smem_dst)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void hn_tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// TMA loads whose completion bytes go to the mbarrier at `bar`'s offset in the LEADER CTA (executed by either CTA)
__device__ __forceinline__ void hn_tma_load_2d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(hn_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(hn_smem_u32(bar) & HN_PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void hn_tma_load_3d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(hn_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(hn_smem_u32(bar) & HN_PEER_BIT_MASK), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
// arrive on the mbarrier at `bar`'s offset in the leader CTA's shared memory (from either CTA of the pair)
__device__ __forceinline__ void hn_mbar_arrive_leader(uint64_t* bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(hn_smem_u32(bar)));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// Four K=16 steps of a pair MMA (see hn_umma_bf16_x4); issued by ONE thread of the leader CTA
__device__ __forceinline__ void hn_umma_bf16_x4_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %5, %6, %3, 1;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %7, %8, %3, 1;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %9, %10, %3, 1;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "l"(desc_a + 2), "l"(desc_b + 2),
        "l"(desc_a + 4), "l"(desc_b + 4), "l"(desc_a + 6), "l"(desc_b + 6)
      : "memory");
}
// tcgen05.commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void hn_umma_commit_pair(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar_addr), "h"((uint16_t)3) : "memory");
}
// Instruction descriptor of a pair MMA: bf16 A/B (K-major both), fp32 accumulate, M = 256 over the two CTAs
__host__ __device__ constexpr uint32_t hn_umma_idesc_bf16_pair(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(256 >> 4) << 24);
}

__device__ __forceinline__ void hn_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (taddr.lane + t).
__device__ __forceinline__ void hn_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void hn_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Instruction descriptor for kind::f16 with bf16 A/B (K-major both), fp32 accumulate, M=128.
__host__ __device__ constexpr uint32_t hn_umma_idesc_bf16(int n) {
  return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | (uint32_t(n >> 3) << 17) |
         (uint32_t(128 >> 4) << 24);
}
// Shared-memory matrix descriptor: K-major tile, rows of 128 bytes, SWIZZLE_128B, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t hn_umma_smem_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr >> 4) & 0x3FFFu) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}

#endif  // __CUDACC__
