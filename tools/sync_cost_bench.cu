// Micro-benchmark 3: cost of the synchronisation instructions around tcgen05.mma, issued from a converged warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/sync_cost_bench tools/sync_cost_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../handnet-pipeline_b200/hn_b200/csrc/hn_common.cuh"
void hn_set_error(const char*, ...) {}

template <int mode, int mmas, int NN>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ volatile uint32_t flag;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    hn_mbar_init(&bar[0], 1);           // completed once below -> parity 0 waits succeed immediately
    hn_mbar_init(&bar[1], 1u << 20);    // never completes: target of commits
    hn_mbar_init(&bar[2], 1);
    hn_mbar_init_fence();
    hn_mbar_arrive(&bar[0]);
    flag = 1;
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) hn_tmem_alloc<256>(&slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  hn_tc_fence_before();
  __syncthreads();
  hn_tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = hn_umma_idesc_bf16(NN);
  if (warp == 1) {
    const uint32_t base = hn_smem_u32(smem);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if constexpr (mode == 0) {                       // satisfied try_wait, whole warp
        acc += hn_mbar_try_wait(&bar[0], 0);
      } else if constexpr (mode == 1) {                // satisfied try_wait, one lane
        if (hn_elect_one()) acc += hn_mbar_try_wait(&bar[0], 0);
        __syncwarp();
      } else if constexpr (mode == 2) {                // tcgen05.commit alone
        if (hn_elect_one()) hn_umma_commit(&bar[1]);
        __syncwarp();
      } else if constexpr (mode == 3) {                // volatile shared-memory flag poll
        acc += flag;
      } else if constexpr (mode == 4) {                // mmas + commit, never waiting
        if (hn_elect_one()) {
          const uint64_t da = hn_umma_smem_desc(base + (it & 3) * 16384);
          const uint64_t db = hn_umma_smem_desc(base + 65536);
#pragma unroll
          for (int k = 0; k < mmas; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
          hn_umma_commit(&bar[1]);
        }
        __syncwarp();
      } else if constexpr (mode == 5) {                // mmas + commit + satisfied try_wait
        acc += hn_mbar_try_wait(&bar[0], 0);
        if (hn_elect_one()) {
          const uint64_t da = hn_umma_smem_desc(base + (it & 3) * 16384);
          const uint64_t db = hn_umma_smem_desc(base + 65536);
#pragma unroll
          for (int k = 0; k < mmas; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
          hn_umma_commit(&bar[1]);
        }
        __syncwarp();
      } else if constexpr (mode == 6) {                // mmas + satisfied try_wait, no commit
        acc += hn_mbar_try_wait(&bar[0], 0);
        if (hn_elect_one()) {
          const uint64_t da = hn_umma_smem_desc(base + (it & 3) * 16384);
          const uint64_t db = hn_umma_smem_desc(base + 65536);
#pragma unroll
          for (int k = 0; k < mmas; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
        }
        __syncwarp();
      } else if constexpr (mode == 8 || mode == 9) {   // MMAs whose A descriptor starts 1 (8) or 0,1,2 (9) rows into the swizzle atom
        if (hn_elect_one()) {
          const uint64_t db = hn_umma_smem_desc(base + 65536);
#pragma unroll
          for (int k = 0; k < mmas; ++k) {
            const int rows = mode == 8 ? 1 : (k / 4) % 3;
            const uint64_t da = hn_umma_smem_desc(base + (it & 1) * 32768 + rows * 128);
            hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
          }
        }
        __syncwarp();
      } else if constexpr (mode == 7) {                // mmas + fence::after_thread_sync only
        hn_tc_fence_after();
        if (hn_elect_one()) {
          const uint64_t da = hn_umma_smem_desc(base + (it & 3) * 16384);
          const uint64_t db = hn_umma_smem_desc(base + 65536);
#pragma unroll
          for (int k = 0; k < mmas; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
        }
        __syncwarp();
      }
    }
    long long t1 = clock64();
    if (threadIdx.x == 32) { out[0] = t1 - t0; out[1] = acc; }
  }
  __syncthreads();
  if (warp == 1) {
    if (hn_elect_one()) hn_umma_commit(&bar[2]);
    __syncwarp();
    hn_mbar_wait(&bar[2], 0);
  }
  __syncthreads();
  if (warp == 0) hn_tmem_dealloc<256>(tmem);
}

template <int mode, int mmas, int NN = 64>
void run(long long* d, const char* name) {
  const int iters = 4000;
  cudaFuncSetAttribute(bench<mode, mmas, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  bench<mode, mmas, NN><<<148, 128, 100000>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d %-34s mmas=%2d: %7.1f cyc/iter (%s)\n", NN, name, mode < 4 ? 0 : mmas, (double)h[0] / iters, cudaGetErrorString(e));
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<0, 4>(d, "try_wait (satisfied), 32 lanes");
  run<1, 4>(d, "try_wait (satisfied), 1 lane");
  run<2, 4>(d, "tcgen05.commit");
  run<3, 4>(d, "ld.volatile.shared flag");
  run<4, 4>(d, "MMAs(N=64) + commit"); run<4, 8>(d, "MMAs(N=64) + commit"); run<4, 12>(d, "MMAs(N=64) + commit");
  run<5, 4>(d, "try_wait + MMAs + commit"); run<5, 8>(d, "try_wait + MMAs + commit"); run<5, 12>(d, "try_wait + MMAs + commit");
  run<6, 4>(d, "try_wait + MMAs"); run<6, 8>(d, "try_wait + MMAs"); run<6, 12>(d, "try_wait + MMAs");
  run<8, 12>(d, "MMAs, A descriptor +1 row");
  run<9, 12>(d, "MMAs, A descriptor +0/1/2 rows");
  run<7, 12, 128>(d, "fence::after + MMAs");
  run<8, 12, 128>(d, "MMAs, A descriptor +1 row");
  run<9, 12, 128>(d, "MMAs, A descriptor +0/1/2 rows");
  run<7, 12, 256>(d, "fence::after + MMAs");
  run<9, 12, 256>(d, "MMAs, A descriptor +0/1/2 rows");
  run<7, 4>(d, "fence::after + MMAs"); run<7, 8>(d, "fence::after + MMAs"); run<7, 12>(d, "fence::after + MMAs");
  return 0;
}
