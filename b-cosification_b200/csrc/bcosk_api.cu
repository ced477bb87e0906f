// bcosk_api.cu -- error reporting, device query and TMA descriptor construction for libbcosk.so.
//
// The CUDA driver's tensor-map encoders are resolved at run time through cudaGetDriverEntryPoint, so the
// library links against cudart only (it must load on a box without libcuda, e.g. the CPU-only CI).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "bcosk_host.h"

namespace bcosk {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static int resolve_driver() {
  if (g_encode_tiled && g_encode_im2col) return BCOSK_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
    return set_error(BCOSK_EUNSUPPORTED, "cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
    return set_error(BCOSK_EUNSUPPORTED, "cuTensorMapEncodeIm2col unavailable (%s)", cudaGetErrorString(e));
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  cudaDriverGetVersion(&g_driver_version);
  return BCOSK_OK;
}

static CUtensorMapSwizzle swizzle_enum(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                      : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                     : (bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
}

int make_im2col_map_nhwc(CUtensorMap* map, const void* base, int nb, int h, int w, int c, int lo_w, int lo_h, int up_w,
                         int up_h, int stride_w, int stride_h, int ch_per_pixel, int pixels, int swizzle_bytes) {
  int rc = resolve_driver();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(BCOSK_EINVAL, "im2col map: base not 16B aligned");
  if (ch_per_pixel * 2 != swizzle_bytes) return set_error(BCOSK_EINVAL, "im2col map: box row must equal the swizzle span");
  if (lo_w < -128 || lo_w > 127 || lo_h < -128 || lo_h > 127 || up_w < -128 || up_w > 127 || up_h < -128 || up_h > 127)
    return set_error(BCOSK_EUNSUPPORTED, "im2col map: corner out of the 8-bit range");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nb};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  int lower[2] = {lo_w, lo_h};
  int upper[2] = {up_w, up_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
  CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                               upper, (cuuint32_t)ch_per_pixel, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle_enum(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(BCOSK_ECUDA, "cuTensorMapEncodeIm2col failed: CUresult %d (nb=%d h=%d w=%d c=%d lo=(%d,%d) up=(%d,%d) s=(%d,%d))",
                     (int)r, nb, h, w, c, lo_w, lo_h, up_w, up_h, stride_w, stride_h);
  // Drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB; the published work-around
  // (CUTLASS cute/atom/copy_traits_sm90_im2col.hpp) clears bit 21 of the second descriptor word.
  if (g_driver_version <= 13010 && (unsigned long long)nb * h * w * c * 2 < 131072ull)
    reinterpret_cast<unsigned long long*>(map)[1] &= ~(1ull << 21);
  return BCOSK_OK;
}

int make_tiled_map_2d(CUtensorMap* map, const void* base, long long cols, long long rows, int box_cols, int box_rows,
                      int swizzle_bytes) {
  int rc = resolve_driver();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(BCOSK_EINVAL, "tiled map: base not 16B aligned");
  if ((cols * 2) % 16 != 0) return set_error(BCOSK_EINVAL, "tiled map: row pitch not a multiple of 16 bytes");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(swizzle_bytes),
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(BCOSK_ECUDA, "cuTensorMapEncodeTiled failed: CUresult %d (cols=%lld rows=%lld box=%dx%d)", (int)r, cols,
                     rows, box_cols, box_rows);
  return BCOSK_OK;
}

// General tiled map (up to 3-D) with explicit byte strides; elem_bytes 2 (16-bit) or 4 (fp32).
int make_tiled_map_nd(CUtensorMap* map, const void* base, int elem_bytes, int rank, const long long* dims,
                      const long long* strides_bytes, const int* box, int swizzle_bytes) {
  int rc = resolve_driver();
  if (rc) return rc;
  if (rank < 2 || rank > 4) return set_error(BCOSK_EINVAL, "tiled map: rank");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(BCOSK_EINVAL, "tiled map: base not 16B aligned");
  cuuint64_t d[4], st[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    d[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides_bytes[i] % 16 != 0) return set_error(BCOSK_EINVAL, "tiled map: stride not a multiple of 16 bytes");
    st[i] = (cuuint64_t)strides_bytes[i];
  }
  CUresult r = g_encode_tiled(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                              (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle_enum(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(BCOSK_ECUDA, "cuTensorMapEncodeTiled (rank %d) failed: CUresult %d (dims %lld %lld %lld box %d %d %d)", rank,
                     (int)r, dims[0], dims[1], rank > 2 ? dims[2] : 0ll, box[0], box[1], rank > 2 ? box[2] : 0);
  return BCOSK_OK;
}

}  // namespace bcosk

extern "C" const char* bcosk_last_error(void) { return bcosk::g_err; }

extern "C" int bcosk_version(void) { return 100; }

extern "C" int bcosk_sizeof_igemm_params(void) { return (int)sizeof(bcosk_igemm_params); }

extern "C" int bcosk_sizeof_wgrad_params(void) { return (int)sizeof(bcosk_wgrad_params); }

extern "C" int bcosk_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
