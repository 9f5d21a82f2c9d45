// Does tcgen05.mma accept a K-major, un-swizzled A operand whose core matrices OVERLAP in shared memory?
//
// A direct 7x7 / stride-2 stem reads, for output pixel ox, the 8 input pixels 2*ox-4 .. 2*ox+3 of a canvas row (4 channels,
// 8 bytes each): consecutive output pixels' windows start 16 bytes apart and overlap by 48 bytes.  In the un-swizzled K-major
// canonical layout element (row r, 16-byte K chunk j) of an operand lives at start + (r % 8) * 16 + (r / 8) * SBO + j * LBO.
// With SBO = 128 and LBO = 16 that is start + 16 * r + 16 * j: the im2col matrix of the window IS the canvas row, and one
// 2 KB run of it serves 128 output pixels (the row-pair stem moves 4x that through TMA: 128 rows of 64 bytes).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/umma_overlap_test tools/umma_overlap_test.cu
// Prints the max |error| against a CPU reference for (a) an ordinary non-overlapping un-swizzled A tile (harness check) and
// (b) the overlapping one.  (Descriptor fields: LBO = stride between K chunks, SBO = stride between 8-row groups; the
// other order reads out of bounds.)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../handnet-pipeline_b200/hn_b200/csrc/hn_common.cuh"
void hn_set_error(const char*, ...) {}

constexpr int M = 128, N = 64, K = 32;      // two K = 16 MMAs

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((addr >> 4) & 0x3FFFu) | (uint64_t((lbo >> 4) & 0x3FFFu) << 16) | (uint64_t((sbo >> 4) & 0x3FFFu) << 32) |
         (uint64_t(1) << 46);
}
__device__ __forceinline__ void umma1(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

// a_bytes: the A operand region copied verbatim to shared memory; (a_lbo, a_sbo): its descriptor strides; a_kstep: bytes
// between the starts of the two K = 16 halves.  B: canonical un-swizzled [N/8][8 rows][16 B] per K chunk, chunks 1024 B apart.
__global__ void __launch_bounds__(128) test_kernel(const uint8_t* a_src, int a_bytes, uint32_t a_lbo, uint32_t a_sbo,
                                                   uint32_t a_kstep, const uint8_t* b_src, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sa = smem;
  uint8_t* sb = smem + 8192;
  for (int i = threadIdx.x; i < a_bytes; i += 128) sa[i] = a_src[i];
  for (int i = threadIdx.x; i < N * K * 2; i += 128) sb[i] = b_src[i];
  if (threadIdx.x == 0) {
    hn_mbar_init(&bar, 1);
    hn_mbar_init_fence();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
  if (threadIdx.x < 32) hn_tmem_alloc<64>(&tmem_slot);
  hn_tc_fence_before();
  __syncthreads();
  hn_tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    if (hn_elect_one()) {
      const uint32_t idesc = hn_umma_idesc_bf16(N);
      const uint32_t a0 = hn_smem_u32(sa), b0 = hn_smem_u32(sb);
      umma1(tmem, desc_noswz(a0, a_lbo, a_sbo), desc_noswz(b0, 1024, 128), idesc, 0u);
      umma1(tmem, desc_noswz(a0 + a_kstep, a_lbo, a_sbo), desc_noswz(b0 + 2048, 1024, 128), idesc, 1u);
      hn_umma_commit_addr<1>(hn_smem_u32(&bar));
    }
    __syncwarp();
  }
  hn_mbar_wait(&bar, 0);
  hn_tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t r[32];
  for (int c = 0; c < 2; ++c) {
    hn_tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, r);
    hn_tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c * 32 + j] = __uint_as_float(r[j]);
  }
  hn_tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) hn_tmem_dealloc<64>(tmem);
}

static float bf(uint16_t v) { uint32_t u = (uint32_t)v << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t tobf(float f) { __nv_bfloat16 h = __float2bfloat16(f); uint16_t v; memcpy(&v, &h, 2); return v; }

int main() {
  srand(7);
  // logical operands
  std::vector<uint16_t> canvas(M * 8 + 24 + 64);                  // A[r][k] = canvas[r * 8 + k], k < 32
  for (auto& v : canvas) v = tobf((rand() % 2001 - 1000) / 1000.0f);
  std::vector<uint16_t> B(N * K);
  for (auto& v : B) v = tobf((rand() % 2001 - 1000) / 1000.0f);
  std::vector<float> ref(M * N);
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += bf(canvas[r * 8 + k]) * bf(B[n * K + k]);
      ref[r * N + n] = acc;
    }
  // B in the canonical un-swizzled K-major layout: chunk j (8 k), row n -> j * 1024 + (n / 8) * 128 + (n % 8) * 16
  std::vector<uint8_t> bsm(N * K * 2);
  for (int n = 0; n < N; ++n)
    for (int j = 0; j < K / 8; ++j) memcpy(&bsm[j * 1024 + (n / 8) * 128 + (n % 8) * 16], &B[n * K + j * 8], 16);
  // (a) ordinary A: chunk j, row r -> j * 2048 + (r / 8) * 128 + (r % 8) * 16
  std::vector<uint8_t> a_std(M * K * 2);
  for (int r = 0; r < M; ++r)
    for (int j = 0; j < K / 8; ++j) memcpy(&a_std[j * 2048 + (r / 8) * 128 + (r % 8) * 16], &canvas[r * 8 + j * 8], 16);
  uint8_t *d_a, *d_b;
  float* d_out;
  cudaMalloc(&d_a, 16384);
  cudaMalloc(&d_b, bsm.size());
  cudaMalloc(&d_out, M * N * 4);
  cudaMemcpy(d_b, bsm.data(), bsm.size(), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  std::vector<float> out(M * N);
  auto run = [&](const char* name, const void* a, int bytes, uint32_t lbo, uint32_t sbo, uint32_t kstep) {
    cudaMemset(d_out, 0, M * N * 4);
    cudaMemcpy(d_a, a, bytes, cudaMemcpyHostToDevice);
    test_kernel<<<1, 128, 16384>>>(d_a, bytes, lbo, sbo, kstep, d_b, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-60s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(out.data(), d_out, M * N * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < M * N; ++i) err = fmax(err, fabs(out[i] - ref[i]));
    printf("%-60s max |err| = %.3e  %s\n", name, err, err < 1e-3 ? "OK" : "MISMATCH");
  };
  run("(a) ordinary tile, LBO = 2048 (K), SBO = 128 (M)", a_std.data(), (int)a_std.size(), 2048, 128, 4096);
  run("(b) overlapping windows, LBO = 16 (K), SBO = 128 (M)", canvas.data(), M * 16 + 48, 16, 128, 32);
  return 0;
}
