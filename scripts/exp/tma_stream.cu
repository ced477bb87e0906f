// Experiment: HBM bandwidth of the epilogue access pattern of the 1x1 launches, without any math.
// A persistent CTA per SM streams [128 rows x BOXC columns] boxes of a row-major [M, N] 16-bit tensor: one TMA load
// (the residual / gain tile) and two TMA stores (y and gain) per tile, DEPTH tiles in flight.
// Variables: box width (64 / 128 / 256 columns of N = 256), tile order (n fastest over the grid, or all n tiles of a
// row block on one CTA), pipeline depth.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I b-cosification_b200/csrc -I include -o scripts/exp/_bin/tma_stream scripts/exp/tma_stream.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include "bcosk_common.cuh"
using namespace bcosk;
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap t_in,
                                                       const __grid_constant__ CUtensorMap t_o1,
                                                       const __grid_constant__ CUtensorMap t_o2, int m_tiles, int n_tiles,
                                                       int boxc, int depth, int order, int nstores) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[8];
  const int tile_bytes = 128 * boxc * 2;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int total = m_tiles * n_tiles;
  auto tile_of = [&](int i) -> int {
    if (order == 0) { const int t = blockIdx.x + i * gridDim.x; return t < total ? t : -1; }
    const int g = blockIdx.x + (i / n_tiles) * gridDim.x;
    return g < m_tiles ? g * n_tiles + i % n_tiles : -1;
  };
  auto issue_load = [&](int i) {
    const int t = tile_of(i);
    if (t < 0) return;
    const int slot = i % depth;
    mbar_arrive_expect_tx(&full[slot], tile_bytes);
    for (int b = 0; b < boxc / 64; ++b)
      tma_load_2d(smem + slot * tile_bytes + b * 16384, &t_in, &full[slot], (t % n_tiles) * boxc + b * 64, (t / n_tiles) * 128);
  };
  for (int i = 0; i < depth - 1; ++i) issue_load(i);
  for (int i = 0;; ++i) {
    const int t = tile_of(i);
    if (t < 0) break;
    const int slot = i % depth;
    // slot (i + depth - 1) % depth was stored from at iteration i - 1: its reads must be done before it is refilled
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    issue_load(i + depth - 1);
    mbar_wait(&full[slot], (i / depth) & 1);
    for (int b = 0; b < boxc / 64; ++b) {
      const uint32_t src = smem_u32(smem + slot * tile_bytes + b * 16384);
      tma_store_2d_addr(&t_o1, src, (t % n_tiles) * boxc + b * 64, (t / n_tiles) * 128);
      if (nstores > 1) tma_store_2d_addr(&t_o2, src, (t % n_tiles) * boxc + b * 64, (t / n_tiles) * 128);
    }
    tma_store_commit();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  EncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  const long long M = 802816, N = 256;
  uint16_t *in, *o1, *o2;
  cudaMalloc(&in, M * N * 2); cudaMalloc(&o1, M * N * 2); cudaMalloc(&o2, M * N * 2);
  cudaMemset(in, 1, M * N * 2);
  auto mk = [&](CUtensorMap* tm, void* base) {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t st[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap ti, t1, t2;
  mk(&ti, in); mk(&t1, o1); mk(&t2, o2);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("%5s %6s %6s %7s %5s | %8s %9s\n", "boxc", "order", "depth", "stores", "ctas", "ms", "GB/s");
  for (int nstores : {2, 1})
    for (int boxc : {64, 128, 256})
      for (int order : {0, 1})
        for (int depth : {2, 3, 4, 6})
          for (int ctas : {148, 296}) {
            const int tile_bytes = 128 * boxc * 2;
            if ((long long)depth * tile_bytes * (ctas / 148) > 200 * 1024) continue;
            const int n_tiles = N / boxc, m_tiles = M / 128;
            if (order == 1 && n_tiles == 1) continue;
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
              cudaEventRecord(e0);
              stream_kernel<<<ctas, 64, depth * tile_bytes>>>(ti, t1, t2, m_tiles, n_tiles, boxc, depth, order, nstores);
              cudaEventRecord(e1);
              cudaError_t e = cudaEventSynchronize(e1);
              if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
              float ms; cudaEventElapsedTime(&ms, e0, e1);
              if (ms < best) best = ms;
            }
            const double bytes = (double)M * N * 2 * (1 + nstores);
            printf("%5d %6d %6d %7d %5d | %8.3f %9.0f\n", boxc, order, depth, nstores, ctas, best, bytes / best / 1e6);
          }
  return 0;
}
