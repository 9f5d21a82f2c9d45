// Micro-benchmark 2: tcgen05 issue / synchronisation costs with a CONVERGED issuing warp (elect.sync), the way the
// conv kernel issues since round 1 session 2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/mma_issue_bench2 tools/mma_issue_bench2.cu
// (a) back-to-back tcgen05.mma (SS, M=128, K=16) for N = 16..256: cycles per MMA
// (b) ring of D stages: [wait full] 4*KG MMAs, commit(empty); a second warp turns empty -> full (no data movement):
//     cycles per k-step as a function of D, N and KG  (the pipeline skeleton of the conv kernel)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../handnet-pipeline_b200/hn_b200/csrc/hn_common.cuh"
void hn_set_error(const char*, ...) {}

template <int N>
__global__ void __launch_bounds__(128, 1) bench(int iters, int mmas, int depth, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[16], empty[16];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { hn_mbar_init(&full[i], 1); hn_mbar_init(&empty[i], 1); }
    hn_mbar_init_fence();
  }
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
    long long t0 = clock64();
    if (mode == 0) {           // (a) MMAs only
      for (int it = 0; it < iters; ++it) {
        if (hn_elect_one()) {
          const uint64_t da = hn_umma_smem_desc(base + (it & 3) * 16384);
          const uint64_t db = hn_umma_smem_desc(base + 65536);
#pragma unroll 4
          for (int k = 0; k < mmas; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
        }
        __syncwarp();
      }
      if (hn_elect_one()) hn_umma_commit(&full[0]);
      __syncwarp();
      hn_mbar_wait(&full[0], 0);
    } else if (mode == 3) {    // (c) ring consumer that peeks at the NEXT stage's barrier before issuing this stage's MMAs
      int stage = 0; uint32_t phase = 0;
      hn_mbar_wait(&full[0], 0);
      for (int it = 0; it < iters; ++it) {
        int nstage = stage + 1; uint32_t nphase = phase;
        if (nstage == depth) { nstage = 0; nphase ^= 1; }
        const bool ready = (it + 1 < iters) ? hn_mbar_try_wait(&full[nstage], nphase) : true;
        hn_tc_fence_after();
        if (hn_elect_one()) {
          const uint64_t da = hn_umma_smem_desc(base + (stage & 3) * 16384);
          const uint64_t db = hn_umma_smem_desc(base + 65536);
#pragma unroll 4
          for (int k = 0; k < mmas; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
          hn_umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (!ready) hn_mbar_wait(&full[nstage], nphase);
        stage = nstage; phase = nphase;
      }
    } else {                   // (b) consumer of a ring
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        hn_mbar_wait(&full[stage], phase);
        hn_tc_fence_after();
        if (hn_elect_one()) {
          const uint64_t da = hn_umma_smem_desc(base + (stage & 3) * 16384);
          const uint64_t db = hn_umma_smem_desc(base + 65536);
          if (mode == 1) {
#pragma unroll 4
            for (int k = 0; k < mmas; ++k) hn_umma_bf16(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
          }
          hn_umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == depth) { stage = 0; phase ^= 1; }
      }
      // drain
      for (int s = 0; s < depth; ++s) {
        // the producer has stopped; wait until the last commits have landed by observing the empty barriers' phases
      }
    }
    long long t1 = clock64();
    if (threadIdx.x == 32) { out[0] = t1 - t0; }
  } else if (warp == 2 && mode == 3) {   // producer with the same early peek
    int stage = 0; uint32_t phase = 0;
    bool ready = true;                       // first pass: all slots are free
    for (int it = 0; it < iters; ++it) {
      if (!ready) hn_mbar_wait(&empty[stage], phase ^ 1);
      int nstage = stage + 1; uint32_t nphase = phase;
      if (nstage == depth) { nstage = 0; nphase ^= 1; }
      ready = hn_mbar_try_wait(&empty[nstage], nphase ^ 1);
      if (hn_elect_one()) hn_mbar_arrive(&full[stage]);
      __syncwarp();
      stage = nstage; phase = nphase;
    }
  } else if (warp == 2 && mode != 0) {   // producer: empty -> full, no data
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      hn_mbar_wait(&empty[stage], phase ^ 1);
      if (hn_elect_one()) hn_mbar_arrive(&full[stage]);
      __syncwarp();
      if (++stage == depth) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  // let outstanding MMAs finish before TMEM goes away
  if (warp == 1) {
    if (hn_elect_one()) hn_umma_commit(&full[15]);
    __syncwarp();
    hn_mbar_wait(&full[15], 0);
  }
  __syncthreads();
  if (warp == 0) hn_tmem_dealloc<256>(tmem);
}

template <int N>
void run(int grid, int iters, int mmas, int depth, int mode) {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  bench<N><<<grid, 128, 100000>>>(iters, mmas, depth, mode, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (mode == 0)
    printf("grid %3d  MMA only        N=%3d mmas/iter=%2d: %.1f cyc/MMA (%s)\n", grid, N, mmas, (double)h[0] / (iters * (double)mmas), cudaGetErrorString(e));
  else
    printf("grid %3d  ring depth %2d %s N=%3d mmas/step=%2d: %.1f cyc/step (%s)\n", grid, depth, mode == 1 ? "with MMAs" : (mode == 3 ? "peek+MMAs" : "no MMAs  "), N, mmas,
           (double)h[0] / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<16>(grid, 4000, 8, 0, 0); run<32>(grid, 4000, 8, 0, 0); run<64>(grid, 4000, 8, 0, 0);
    run<128>(grid, 4000, 8, 0, 0); run<256>(grid, 4000, 8, 0, 0);
  }
  for (int depth : {4}) {
    run<64>(148, 4000, 4, depth, 2);
    run<64>(148, 4000, 4, depth, 1);
    run<64>(148, 4000, 8, depth, 1);
    run<64>(148, 4000, 12, depth, 1);
    run<128>(148, 4000, 4, depth, 1);
    run<128>(148, 4000, 8, depth, 1);
    run<256>(148, 4000, 4, depth, 1);
  }
  for (int depth : {3, 4, 8}) {
    run<64>(148, 4000, 4, depth, 3);
    run<64>(148, 4000, 8, depth, 3);
    run<64>(148, 4000, 12, depth, 3);
    run<128>(148, 4000, 4, depth, 3);
    run<128>(148, 4000, 8, depth, 3);
    run<128>(148, 4000, 12, depth, 3);
    run<256>(148, 4000, 4, depth, 3);
  }
  return 0;
}
