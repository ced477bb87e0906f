// bcosk_host.h -- host-side helpers shared by the translation units of libbcosk.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/bcosk.h"

namespace bcosk {

// records the message (thread-local) and returns `code`
int set_error(int code, const char* fmt, ...);

#define BCOSK_CUDA_CHECK(expr)                                                                         \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return ::bcosk::set_error(BCOSK_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

// TMA descriptor for im2col loads of an NHWC 16-bit tensor seen as (C, W, H, N).
//   box = [pixels x ch_per_pixel]; swizzle_bytes = 128 or 64 (must equal ch_per_pixel * 2)
int make_im2col_map_nhwc(CUtensorMap* map, const void* base, int nb, int h, int w, int c, int lo_w, int lo_h, int up_w,
                         int up_h, int stride_w, int stride_h, int ch_per_pixel, int pixels, int swizzle_bytes);

// TMA descriptor for tiled loads of a row-major [rows][cols] 16-bit matrix; box = [box_rows x box_cols]
int make_tiled_map_2d(CUtensorMap* map, const void* base, long long cols, long long rows, int box_cols, int box_rows,
                      int swizzle_bytes);

// General tiled descriptor (rank 2 or 3) with explicit byte strides; elem_bytes 2 or 4.  dims[0] is the contiguous one.
int make_tiled_map_nd(CUtensorMap* map, const void* base, int elem_bytes, int rank, const long long* dims,
                      const long long* strides_bytes, const int* box, int swizzle_bytes);

}  // namespace bcosk
