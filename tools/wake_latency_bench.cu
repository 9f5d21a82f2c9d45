// Micro-benchmark 4: how long after an mbarrier phase completes does a thread blocked in mbarrier.try_wait resume?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/wake_latency_bench tools/wake_latency_bench.cu
// Ping-pong between warp 1 (signals bar[0], waits bar[1]) and warp 2 (waits bar[0], signals bar[1]); the signal of
// warp 1 is a plain mbarrier.arrive (mode 0), a tcgen05.commit with nothing pending (mode 1), or a tcgen05.commit
// behind MMAS tcgen05.mma (N=64) (mode 2).  Reports cycles per round trip.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../handnet-pipeline_b200/hn_b200/csrc/hn_common.cuh"
void hn_set_error(const char*, ...) {}

template <int MODE, int MMAS, int POLL_LANES>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    hn_mbar_init(&bar[0], 1); hn_mbar_init(&bar[1], 1); hn_mbar_init(&bar[2], 1);
    hn_mbar_init_fence();
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0) hn_tmem_alloc<64>(&slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  hn_tc_fence_before();
  __syncthreads();
  hn_tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = hn_umma_idesc_bf16(64);
  auto wait = [&](uint64_t* b, uint32_t parity) {
    if (POLL_LANES == 32) { hn_mbar_wait(b, parity); }
    else { if (lane == 0) hn_mbar_wait(b, parity); __syncwarp(); }
  };
  if (warp == 1) {
    const uint32_t base = hn_smem_u32(smem);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (hn_elect_one()) {
        if (MODE == 0) hn_mbar_arrive(&bar[0]);
        else {
          if (MODE == 2) {
            const uint64_t da = hn_umma_smem_desc(base + (it & 3) * 16384);
            const uint64_t db = hn_umma_smem_desc(base + 65536);
#pragma unroll
            for (int k = 0; k < MMAS; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
          }
          hn_umma_commit(&bar[0]);
        }
      }
      __syncwarp();
      wait(&bar[1], ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 32) out[0] = t1 - t0;
  } else if (warp == 2) {
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      wait(&bar[0], ph);
      ph ^= 1;
      if (hn_elect_one()) hn_mbar_arrive(&bar[1]);
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 1) {
    if (hn_elect_one()) hn_umma_commit(&bar[2]);
    __syncwarp();
    hn_mbar_wait(&bar[2], 0);
  }
  __syncthreads();
  if (warp == 0) hn_tmem_dealloc<64>(tmem);
}

template <int MODE, int MMAS, int POLL>
void run(long long* d, const char* name) {
  const int iters = 2000;
  cudaFuncSetAttribute(bench<MODE, MMAS, POLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  bench<MODE, MMAS, POLL><<<148, 128, 100000>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-52s poll lanes %2d: %7.1f cyc/round trip (%s)\n", name, POLL, (double)h / iters, cudaGetErrorString(e));
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<0, 0, 32>(d, "arrive -> wait, arrive -> wait");
  run<0, 0, 1>(d, "arrive -> wait, arrive -> wait");
  run<1, 0, 32>(d, "tcgen05.commit (idle pipe) -> wait, arrive -> wait");
  run<1, 0, 1>(d, "tcgen05.commit (idle pipe) -> wait, arrive -> wait");
  run<2, 4, 32>(d, "4 MMAs + commit -> wait, arrive -> wait");
  run<2, 12, 32>(d, "12 MMAs + commit -> wait, arrive -> wait");
  run<2, 12, 1>(d, "12 MMAs + commit -> wait, arrive -> wait");
  return 0;
}
