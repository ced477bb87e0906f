// Experiment: tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SS operands) cycles per MMA as a function of N, of the
// number of independent accumulators the chain alternates over, of the swizzle span and of a row-shifted A start.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I b-cosification_b200/csrc -o scripts/exp/_bin/mma_rate scripts/exp/mma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "bcosk_common.cuh"
using namespace bcosk;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int nacc, int row_bytes, int shift_rows, int count, int kadv,
                                                      long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(1u, 128, N);
    const uint32_t a0 = smem_u32(smem) + shift_rows * row_bytes;   // A region: first 48 KB
    const uint32_t b0 = smem_u32(smem) + 48 * 1024;                // B region
    const uint64_t da = umma_smem_desc_kmajor(a0, row_bytes), db = umma_smem_desc_kmajor(b0, row_bytes);
    const uint32_t kstep = kadv ? 2u : 0u;                         // +32 bytes in descriptor units of 16 bytes
    const uint32_t acc1 = nacc > 1 ? N : 0, acc2 = nacc > 2 ? 2 * N : 0, acc3 = nacc > 2 ? 3 * N : acc1;
    const long long t0 = clock64();
    // 8 MMAs per iteration, no per-MMA integer work: descriptors differ by constants only
    umma_f16(tmem, da, db, idesc, 0);
    umma_f16(tmem + acc1, da + kstep, db + kstep, idesc, 0);
    umma_f16(tmem + acc2, da + 2 * kstep, db + 2 * kstep, idesc, 0);
    umma_f16(tmem + acc3, da + 3 * kstep, db + 3 * kstep, idesc, 0);
#pragma unroll 1
    for (int i = 4; i < count; i += 8) {
      umma_f16(tmem, da, db, idesc, 1);
      umma_f16(tmem + acc1, da + kstep, db + kstep, idesc, 1);
      umma_f16(tmem + acc2, da + 2 * kstep, db + 2 * kstep, idesc, 1);
      umma_f16(tmem + acc3, da + 3 * kstep, db + 3 * kstep, idesc, 1);
      umma_f16(tmem, da, db, idesc, 1);
      umma_f16(tmem + acc1, da + kstep, db + kstep, idesc, 1);
      umma_f16(tmem + acc2, da + 2 * kstep, db + 2 * kstep, idesc, 1);
      umma_f16(tmem + acc3, da + 3 * kstep, db + 3 * kstep, idesc, 1);
    }
    umma_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int count = 4 + 8 * 64;
  printf("%5s %5s %6s %6s %5s | %10s %10s\n", "N", "nacc", "swz", "shift", "kadv", "issue/mma", "total/mma");
  const int Ns[] = {32, 64, 128, 256};
  for (int N : Ns)
    for (int nacc : {1, 2, 4})
      for (int rb : {128, 64})
        for (int shift : {0, 1, 3})
          for (int kadv : {0, 1}) {
            if (nacc * N > 512) continue;
            if (kadv && rb == 64) continue;
            if (shift == 3 && nacc != 1) continue;
            for (int rep = 0; rep < 2; ++rep) {
              rate_kernel<<<1, 128, 96 * 1024>>>(N, nacc, rb, shift, count, kadv, d);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h[2];
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%5d %5d %6d %6d %5d | %10.1f %10.1f\n", N, nacc, rb, shift, kadv, (double)h[0] / count, (double)h[1] / count);
          }
  return 0;
}
