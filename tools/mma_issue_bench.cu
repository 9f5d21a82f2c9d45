// Micro-benchmark: how fast can ONE thread issue tcgen05.mma (SS operands, M=128, K=16) back to back, as a function
// of N and of how often it commits?  Operands are whatever is in shared memory; only timing matters.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/mma_issue_bench tools/mma_issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../handnet-pipeline_b200/hn_b200/csrc/hn_common.cuh"
void hn_set_error(const char*, ...) {}

__device__ __forceinline__ void wait_ptx(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(hn_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void wait_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(hn_smem_u32(bar)), "r"(parity) : "memory");
  }
}


template <int N>
__global__ void __launch_bounds__(128, 1) bench(int iters, int mmas_per_commit, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) hn_mbar_init(&bar[i], (mode == 0 || mode >= 3) ? 1 : iters); hn_mbar_init_fence(); }
  if (threadIdx.x < 32) hn_tmem_alloc<256>(&slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  hn_tc_fence_before();
  __syncthreads();
  hn_tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = hn_umma_idesc_bf16(N);
    const uint64_t da = hn_umma_smem_desc(hn_smem_u32(smem));
    const uint64_t db = hn_umma_smem_desc(hn_smem_u32(smem) + 16384);
    long long t0 = clock64();
    if (mode == 0) {          // ring of 8 commits in flight, like a smem pipeline
      for (int it = 0; it < iters; ++it) {
        const int b = it & 7;
        if (it >= 8) hn_mbar_wait(&bar[b], ((it >> 3) - 1) & 1);
        for (int k = 0; k < mmas_per_commit; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
        hn_umma_commit(&bar[b]);
      }
    } else if (mode == 3 || mode == 4) {   // ring of 8 with a PTX-only try_wait loop (3) / test_wait (4)
      for (int it = 0; it < iters; ++it) {
        const int b = it & 7;
        if (it >= 8) { if (mode == 3) wait_ptx(&bar[b], ((it >> 3) - 1) & 1); else wait_test(&bar[b], ((it >> 3) - 1) & 1); }
        for (int k = 0; k < mmas_per_commit; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
        hn_umma_commit(&bar[b]);
      }
    } else if (mode == 5) {   // waits only (barriers completed by plain arrives), no MMAs: cost of a satisfied wait
      for (int it = 0; it < iters; ++it) {
        const int b = it & 7;
        if (it >= 8) hn_mbar_wait(&bar[b], ((it >> 3) - 1) & 1);
        hn_mbar_arrive(&bar[b]);
      }
    } else if (mode == 1) {   // commits but never wait inside the loop (one barrier expecting `iters` arrivals)
      for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < mmas_per_commit; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
        hn_umma_commit(&bar[0]);
      }
    } else {                  // MMAs only, one commit at the very end
      for (int it = 0; it < iters; ++it)
        for (int k = 0; k < mmas_per_commit; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
      for (int it = 0; it < iters; ++it) hn_umma_commit(&bar[0]);
    }
    long long t1 = clock64();
    if (mode == 0 || mode >= 3) { for (int it = iters - 8; it < iters; ++it) hn_mbar_wait(&bar[it & 7], (it >> 3) & 1); }
    else hn_mbar_wait(&bar[0], 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) hn_tmem_dealloc<256>(tmem);
}

template <int N>
void run(int iters, int mpc, int mode) {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  bench<N><<<1, 128, 70000>>>(iters, mpc, mode, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("mode %d N=%3d mmas/commit=%2d iters=%d: issue loop %.1f cyc/MMA, incl. drain %.1f cyc/MMA  (%s)\n", mode, N, mpc, iters,
         (double)h[0] / (iters * mpc), (double)h[1] / (iters * mpc), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int mode : {0, 3, 4, 5})
    for (int mpc : {1, 4}) {
      run<64>(2000, mpc, mode);
      run<256>(2000, mpc, mode);
    }
  return 0;
}
