// Experiment: which 3-D tiled TMA stores (smem -> global) are legal on sm_100a?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_store3d scripts/exp/tma_store3d.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void store_kernel(const __grid_constant__ CUtensorMap tm, int c1, int c2, int rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // tile [rows][128 B], 16-byte units XOR-swizzled by row; value = row index (bf16 bits = row + 1)
  for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int unit = (c / 8) ^ (r & 7);
    reinterpret_cast<uint16_t*>(smem)[r * 64 + unit * 8 + (c % 8)] = (uint16_t)(r + 1);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tm), "r"(sa), "r"(0),
                 "r"(c1), "r"(c2)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main() {
  EncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  struct Case { int X, R, boxX, c1, c2; const char* what; };
  std::vector<Case> cases = {
      {200, 24, 128, 0, 0, "box < dim, in bounds"},
      {200, 24, 128, 150, 1, "box < dim, clipped high"},
      {200, 24, 128, -40, 2, "box < dim, negative start"},
      {12, 24, 128, 0, 0, "box > dim"},
      {12, 24, 128, -5, 1, "box > dim, negative start"},
      {112, 512, 128, -3, 7, "stem-like: dim 112, negative start"},
      {112, 512, 128, 100, 7, "stem-like: clipped high"},
  };
  for (auto& cs : cases) {
    const int C = 64;
    size_t n = (size_t)C * cs.X * cs.R;
    uint16_t* d;
    cudaMalloc(&d, n * 2);
    cudaMemset(d, 0, n * 2);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)cs.X, (cuuint64_t)cs.R};
    cuuint64_t st[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * cs.X};
    cuuint32_t box[3] = {64, (cuuint32_t)cs.boxX, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", cs.what, (int)r); cudaFree(d); continue; }
    cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    store_kernel<<<1, 128, cs.boxX * 128, 0>>>(tm, cs.c1, cs.c2, cs.boxX);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-40s launch error: %s\n", cs.what, cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> h(n);
    cudaMemcpy(h.data(), d, n * 2, cudaMemcpyDeviceToHost);
    long bad = 0, written = 0;
    for (int rr = 0; rr < cs.R; ++rr)
      for (int x = 0; x < cs.X; ++x)
        for (int c = 0; c < C; ++c) {
          uint16_t v = h[((size_t)rr * cs.X + x) * C + c];
          int r_tile = x - cs.c1;
          uint16_t exp = (rr == cs.c2 && r_tile >= 0 && r_tile < cs.boxX) ? (uint16_t)(r_tile + 1) : 0;
          bad += v != exp;
          written += v != 0;
        }
    printf("%-40s ok: wrote %ld elements, %ld mismatches\n", cs.what, written, bad);
    cudaFree(d);
  }
  return 0;
}
