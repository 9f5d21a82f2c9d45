// Micro-benchmark 5: how fast can one elected lane issue TMA tile loads, as a function of the box size?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/tma_issue_bench tools/tma_issue_bench.cu -lcuda
// A producer warp issues `iters` loads of a [ROWS x 64] bf16 box (SWIZZLE_128B) into a ring of 8 slots; a consumer warp
// frees each slot as soon as it has landed.  Reports cycles per load in steady state (source: a 64 MiB matrix, so
// most of it comes from L2 after the first pass) and the cycles the issuing thread spends per load when the ring
// never blocks (first 8 loads).
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../handnet-pipeline_b200/hn_b200/csrc/hn_common.cuh"
void hn_set_error(const char*, ...) {}

template <int ROWS, int P>
__global__ void __launch_bounds__(32 * (P + 1), 1) bench(const __grid_constant__ CUtensorMap tm, int iters, int total_rows, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8], empty[8];
  constexpr int SLOT = ((ROWS * 128 + 1023) / 1024) * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { hn_mbar_init(&full[i], 1); hn_mbar_init(&empty[i], 1); }
    hn_mbar_init_fence();
  }
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp < P) {
    int stage = warp; uint32_t phase = 0;
    int row = (blockIdx.x * 7919) % (total_rows - ROWS);
    long long t0 = clock64(), t8 = 0;
    for (int it = warp; it < iters; it += P) {
      hn_mbar_wait(&empty[stage], phase ^ 1);
      if (hn_elect_one()) {
        hn_mbar_expect_tx(&full[stage], ROWS * 128);
        hn_tma_load_2d(smem + stage * SLOT, &tm, &full[stage], 0, row);
      }
      __syncwarp();
      row += ROWS * 148; if (row >= total_rows - ROWS) row -= (total_rows - ROWS);
      stage += P; if (stage >= 8) { stage -= 8; phase ^= 1; }
      if (it == 7) t8 = clock64();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t8 - t0; }
  } else if (warp == P) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      hn_mbar_wait(&full[stage], phase);
      if (hn_elect_one()) hn_mbar_arrive(&empty[stage]);
      __syncwarp();
      if (++stage == 8) { stage = 0; phase ^= 1; }
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int ROWS, int P = 1>
void run(void* buf, int total_rows, long long* d, PFN_encodeTiled enc, int grid) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, ROWS}, ones[2] = {1, 1};
  enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int iters = 2000;
  cudaFuncSetAttribute(bench<ROWS, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220000);
  bench<ROWS, P><<<grid, 32 * (P + 1), 8 * (((ROWS * 128 + 1023) / 1024) * 1024) + 2048>>>(tm, iters, total_rows, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("producers %d grid %3d box %3d rows (%6d B): %7.1f cyc/load steady (%5.1f B/clk/SM), %6.1f cyc/issue unblocked (%s)\n", P, grid, ROWS, ROWS * 128,
         (double)h[0] / iters, ROWS * 128.0 * iters / h[0], (double)h[1] / 8, cudaGetErrorString(e));
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)p;
  const int total_rows = 512 * 1024;   // 64 MiB
  void* buf; cudaMalloc(&buf, (size_t)total_rows * 128); cudaMemset(buf, 0, (size_t)total_rows * 128);
  long long* d; cudaMalloc(&d, 148 * 16);
  for (int grid : {1, 148}) {
    run<16>(buf, total_rows, d, enc, grid); run<32>(buf, total_rows, d, enc, grid); run<64>(buf, total_rows, d, enc, grid);
    run<128>(buf, total_rows, d, enc, grid); run<136>(buf, total_rows, d, enc, grid); run<192>(buf, total_rows, d, enc, grid);
    run<256>(buf, total_rows, d, enc, grid);
  }
  run<64, 2>(buf, total_rows, d, enc, 148); run<64, 4>(buf, total_rows, d, enc, 148);
  run<136, 2>(buf, total_rows, d, enc, 148); run<136, 4>(buf, total_rows, d, enc, 148);
  return 0;
}
