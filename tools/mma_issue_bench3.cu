// Micro-benchmark 3: does a tcgen05.mma (SS, M=128, K=16, bf16) cost more when every MMA reads FRESH operands, the way
// the conv kernel's resident-weights path issues them (9 taps x 4 k-slices per 64-channel chunk: A tile of kernel row t
// read from row sx on, weight tile of tap t*3+sx), than in the steady loop of mma_issue_bench2 (one A tile, one B tile)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/mma_issue_bench3 tools/mma_issue_bench3.cu
// variant bits: 1 = a different weight tile per tap, 2 = A start shifted by sx rows (128 B), 4 = a different A tile per
// kernel row.  0 = everything reads the same A and B tile.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../handnet-pipeline_b200/hn_b200/csrc/hn_common.cuh"
void hn_set_error(const char*, ...) {}

constexpr int A_SLOT = 17408;     // 136 rows x 128 B (1024-byte aligned)

template <int N>
__global__ void __launch_bounds__(128, 1) bench(int iters, int variant, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[2];
  __shared__ uint32_t slot;
  constexpr int B_TILE = N * 128;
  constexpr int TOTAL = 3 * A_SLOT + 9 * B_TILE;
  for (int i = threadIdx.x; i < TOTAL / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { hn_mbar_init(&full[0], 1); hn_mbar_init_fence(); }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) hn_tmem_alloc<256>(&slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  hn_tc_fence_before();
  __syncthreads();
  hn_tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = hn_umma_idesc_bf16(N);
  if (warp == 1) {
    const uint32_t base = hn_smem_u32(smem);
    const uint32_t bbase = base + 3 * A_SLOT;
    const uint32_t vb = (variant & 1) ? B_TILE : 0, vs = (variant & 2) ? 128 : 0, va = (variant & 4) ? A_SLOT : 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (hn_elect_one()) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
#pragma unroll
          for (int sx = 0; sx < 3; ++sx) {
            const uint64_t da = hn_umma_smem_desc(base + t * va + sx * vs);
            const uint64_t db = hn_umma_smem_desc(bbase + (t * 3 + sx) * vb);
#pragma unroll
            for (int k = 0; k < 4; ++k) hn_umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, 1);
          }
        }
      }
      __syncwarp();
    }
    if (hn_elect_one()) hn_umma_commit(&full[0]);
    __syncwarp();
    hn_mbar_wait(&full[0], 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
  }
  hn_tc_fence_before();
  __syncthreads();
  if (warp == 0) hn_tmem_dealloc<256>(tmem);
}

template <int N>
void run(int iters, int variant) {
  long long* d; cudaMalloc(&d, 16);
  const int smem = 3 * A_SLOT + 9 * N * 128 + 2048;
  cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N><<<148, 128, smem>>>(iters, variant, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d variant %d (%s%s%s): %.1f cycles per MMA (%s)\n", N, variant, (variant & 1) ? "B per tap " : "", (variant & 2) ? "A row shift " : "",
         (variant & 4) ? "A per kernel row" : "", (double)h[0] / (iters * 36.0), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int v : {0, 1, 2, 4, 7}) { run<64>(2000, v); }
  for (int v : {0, 1, 2, 4, 7}) { run<128>(2000, v); }
  for (int v : {0, 7}) { run<32>(2000, v); run<16>(2000, v); }
  return 0;
}
