// bcosk_layout.cu -- layout / dtype bridges and per-channel helpers for the module-level (un-fused) path.
//
// The reference's modules exchange NCHW fp32 tensors; the kernels work on NHWC 16-bit (precision planes).  These
// bandwidth kernels are the bridge, plus eval/train-mode uncentered batch norm and the logit layer on NCHW fp32.
// Thread = pixel, looping over 8-channel groups: reads of one channel plane are coalesced across the warp.
#include "../../include/bcosk.h"
#include "bcosk_common.cuh"
#include "bcosk_host.h"

namespace bcosk {

static inline unsigned nblocks(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

template <typename T>
__global__ void nchw_to_nhwc16_kernel(const float* __restrict__ x, int nb, int c, long long hw, T* __restrict__ out, int cp,
                                      int planes, const float* __restrict__ mul, int mul_ld, float* __restrict__ sq) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over nb*hw
  if (pix >= (long long)nb * hw) return;
  const long long img = pix / hw, sp = pix - img * hw;
  const float* src = x + img * c * hw + sp;
  T* dst = out + pix * ((long long)planes * cp);
  float sqacc = 0.f;
  for (int g = 0; g < cp / 8; ++g) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = g * 8 + i;
      float v = (ch < c) ? __ldg(src + (long long)ch * hw) : 0.f;
      if (mul != nullptr && ch < c) v *= __ldg(mul + pix * mul_ld + ch);
      f[i] = v;
    }
    float r[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[i] = f[i]; acc[i] = 0.f; }
    for (int pl = 0; pl < planes; ++pl) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        w[k] = Cvt<T>::pack2(r[2 * k], r[2 * k + 1]);
        const float2 q = Cvt<T>::unpack2(w[k]);
        r[2 * k] -= q.x; r[2 * k + 1] -= q.y;
        acc[2 * k] += q.x; acc[2 * k + 1] += q.y;
      }
      *reinterpret_cast<uint4*>(dst + (long long)pl * cp + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sqacc = fmaf(acc[i], acc[i], sqacc);
  }
  if (sq != nullptr) sq[pix] = sqacc;
}

template <typename T>
__global__ void nhwc_to_nchw_f32_kernel(const void* __restrict__ y, int y_f32, int nb, int c, long long hw, int ld,
                                        int planes, float* __restrict__ out) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)nb * hw) return;
  const long long img = pix / hw, sp = pix - img * hw;
  float* dst = out + img * c * hw + sp;
  const int cpl = ld / (y_f32 ? 1 : planes);   // channels per plane (physical)
  for (int ch0 = 0; ch0 < c; ch0 += 8) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
    if (y_f32) {
      const float* src = reinterpret_cast<const float*>(y) + pix * ld + ch0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (ch0 + i < c) f[i] = __ldg(src + i);
    } else {
      const T* src = reinterpret_cast<const T*>(y) + pix * ld + ch0;
      for (int pl = 0; pl < planes; ++pl) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (long long)pl * cpl));
        float2 q;
        q = Cvt<T>::unpack2(u.x); f[0] += q.x; f[1] += q.y;
        q = Cvt<T>::unpack2(u.y); f[2] += q.x; f[3] += q.y;
        q = Cvt<T>::unpack2(u.z); f[4] += q.x; f[5] += q.y;
        q = Cvt<T>::unpack2(u.w); f[6] += q.x; f[7] += q.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (ch0 + i < c) dst[(long long)(ch0 + i) * hw] = f[i];
  }
}

// y = relu?((x * alpha[c] + beta[c]) * smul + sadd) on NCHW fp32 (alpha/beta optional)
__global__ void scale_bias_nchw_kernel(const float* __restrict__ x, long long total, int c, long long hw,
                                       const float* __restrict__ alpha, const float* __restrict__ beta, float smul,
                                       float sadd, int relu, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = (int)((i / hw) % c);
  float v = __ldg(x + i);
  if (alpha != nullptr) v *= __ldg(alpha + ch);
  if (beta != nullptr) v += __ldg(beta + ch);
  v = fmaf(v, smul, sadd);
  out[i] = (relu && v < 0.f) ? 0.f : v;
}

// per-channel mean and biased variance over (N, H, W) of an NCHW fp32 tensor: one block per channel
__global__ void channel_stats_nchw_kernel(const float* __restrict__ x, int nb, int c, long long hw, float* __restrict__ mean,
                                          float* __restrict__ var) {
  const int ch = blockIdx.x;
  double s = 0.0, s2 = 0.0;
  for (int n = 0; n < nb; ++n) {
    const float* src = x + ((long long)n * c + ch) * hw;
    for (long long i = threadIdx.x; i < hw; i += blockDim.x) {
      const double v = (double)__ldg(src + i);
      s += v;
      s2 += v * v;
    }
  }
  __shared__ double sh[2][32];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = s2; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    s = lane < nw ? sh[0][lane] : 0.0;
    s2 = lane < nw ? sh[1][lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      const double cnt = (double)nb * (double)hw;
      const double m = s / cnt;
      mean[ch] = (float)m;
      var[ch] = (float)fmax(s2 / cnt - m * m, 0.0);
    }
  }
}

}  // namespace bcosk

using namespace bcosk;

static inline cudaStream_t S2(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" int bcosk_nchw_to_nhwc16(const float* x, int32_t nb, int32_t c, int32_t h, int32_t w, void* out, int32_t cp,
                                    int32_t planes, int32_t dtype, const float* mul, int32_t mul_ld, float* sq,
                                    void* stream) {
  if (!x || !out || cp % 8 || cp < c || planes < 1) return set_error(BCOSK_EINVAL, "nchw_to_nhwc16: bad argument");
  const long long hw = (long long)h * w, n = (long long)nb * hw;
  if (dtype == BCOSK_DTYPE_BF16)
    nchw_to_nhwc16_kernel<__nv_bfloat16><<<nblocks(n, 128), 128, 0, S2(stream)>>>(
        x, nb, c, hw, reinterpret_cast<__nv_bfloat16*>(out), cp, planes, mul, mul_ld, sq);
  else if (dtype == BCOSK_DTYPE_F16)
    nchw_to_nhwc16_kernel<__half><<<nblocks(n, 128), 128, 0, S2(stream)>>>(x, nb, c, hw, reinterpret_cast<__half*>(out), cp,
                                                                           planes, mul, mul_ld, sq);
  else
    return set_error(BCOSK_EINVAL, "nchw_to_nhwc16: dtype");
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_nhwc_to_nchw_f32(const void* y, int32_t y_f32, int32_t nb, int32_t c, int32_t h, int32_t w, int32_t ld,
                                      int32_t planes, int32_t dtype, float* out, void* stream) {
  if (!y || !out || (!y_f32 && (ld / planes) % 8)) return set_error(BCOSK_EINVAL, "nhwc_to_nchw_f32: bad argument");
  const long long hw = (long long)h * w, n = (long long)nb * hw;
  if (dtype == BCOSK_DTYPE_F16)
    nhwc_to_nchw_f32_kernel<__half><<<nblocks(n, 128), 128, 0, S2(stream)>>>(y, y_f32, nb, c, hw, ld, planes, out);
  else
    nhwc_to_nchw_f32_kernel<__nv_bfloat16><<<nblocks(n, 128), 128, 0, S2(stream)>>>(y, y_f32, nb, c, hw, ld, planes, out);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_scale_bias_nchw(const float* x, int32_t nb, int32_t c, int64_t hw, const float* alpha, const float* beta,
                                     float smul, float sadd, int32_t relu, float* out, void* stream) {
  if (!x || !out || c < 1) return set_error(BCOSK_EINVAL, "scale_bias_nchw: bad argument");
  const long long total = (long long)nb * c * hw;
  scale_bias_nchw_kernel<<<nblocks(total, 256), 256, 0, S2(stream)>>>(x, total, c, hw, alpha, beta, smul, sadd, relu, out);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_channel_stats_nchw(const float* x, int32_t nb, int32_t c, int64_t hw, float* mean, float* var,
                                        void* stream) {
  if (!x || !mean || !var) return set_error(BCOSK_EINVAL, "channel_stats_nchw: bad argument");
  channel_stats_nchw_kernel<<<c, 256, 0, S2(stream)>>>(x, nb, c, hw, mean, var);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

// ------------------------------------------------------------------------------------------------
// Small strided movers of the module-level path (no eager torch indexing on the path):
//   zero_insert_nhwc      dst[img, s*p, s*q, :] = src[img, p, q, :]      (the zero-inserted gradient of a strided k x k conv; the
//                         other positions of dst are zero from allocation and never written)
//   nhwc_scatter_nchw_f32 out[img, ch, s*p, s*q] = y[img, p, q, ch]      (data gradient of a strided 1x1 conv: it lives on the sampled
//                         input positions only; out is zero elsewhere)
//   pixel_sqsum_nchw_f32  sq[img, pix] = sum_ch x[img, ch, pix]^2        (calc_patch_norms' first step on the reference's own layout)
// ------------------------------------------------------------------------------------------------
namespace bcosk {

__global__ void zero_insert_nhwc_kernel(const uint4* __restrict__ src, int oh, int ow, int vec_per_pix, uint4* __restrict__ dst, int h, int w,
                                        int stride, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long pix = idx / vec_per_pix;
  const int v = (int)(idx - pix * vec_per_pix);
  const long long img = pix / ((long long)oh * ow);
  const int r = (int)(pix - img * oh * ow);
  const int p = r / ow, q = r - p * ow;
  dst[((img * h + (long long)p * stride) * w + (long long)q * stride) * vec_per_pix + v] = src[idx];
}

template <typename T>
__global__ void nhwc_scatter_nchw_f32_kernel(const void* __restrict__ y, int y_f32, int nb, int c, int oh, int ow, int ld, int planes,
                                             float* __restrict__ out, int H, int W, int stride) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ohw = (long long)oh * ow;
  if (pix >= (long long)nb * ohw) return;
  const long long img = pix / ohw;
  const int r = (int)(pix - img * ohw);
  const int p = r / ow, q = r - p * ow;
  float* dst = out + img * c * (long long)H * W + (long long)p * stride * W + (long long)q * stride;
  const int cpl = ld / (y_f32 ? 1 : planes);
  for (int ch = 0; ch < c; ++ch) {
    float f = 0.f;
    if (y_f32) f = __ldg(reinterpret_cast<const float*>(y) + pix * ld + ch);
    else
      for (int pl = 0; pl < planes; ++pl) f += (float)reinterpret_cast<const T*>(y)[pix * ld + (long long)pl * cpl + ch];
    dst[(long long)ch * H * W] = f;
  }
}

__global__ void pixel_sqsum_nchw_f32_kernel(const float* __restrict__ x, int c, long long hw, long long total, float* __restrict__ sq) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const long long img = pix / hw, sp = pix - img * hw;
  const float* src = x + img * c * hw + sp;
  float acc = 0.f;
  for (int ch = 0; ch < c; ++ch) {
    const float v = __ldg(src + (long long)ch * hw);
    acc = fmaf(v, v, acc);
  }
  sq[pix] = acc;
}

}  // namespace bcosk

extern "C" int bcosk_zero_insert_nhwc(const void* src, int32_t nb, int32_t oh, int32_t ow, int32_t row_elems, void* dst, int32_t h, int32_t w,
                                      int32_t stride, void* stream) {
  using namespace bcosk;
  if (!src || !dst || nb < 1 || row_elems % 8 || stride < 1 || (oh - 1) * stride >= h || (ow - 1) * stride >= w)
    return set_error(BCOSK_EINVAL, "zero_insert_nhwc: bad argument");
  const int vpp = row_elems / 8;
  const long long total = (long long)nb * oh * ow * vpp;
  zero_insert_nhwc_kernel<<<nblocks(total, 256), 256, 0, S2(stream)>>>(reinterpret_cast<const uint4*>(src), oh, ow, vpp, reinterpret_cast<uint4*>(dst), h,
                                                                      w, stride, total);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_nhwc_scatter_nchw_f32(const void* y, int32_t y_f32, int32_t nb, int32_t c, int32_t oh, int32_t ow, int32_t ld,
                                           int32_t planes, int32_t dtype, float* out, int32_t h, int32_t w, int32_t stride, void* stream) {
  using namespace bcosk;
  if (!y || !out || nb < 1 || c < 1 || stride < 1 || (oh - 1) * stride >= h || (ow - 1) * stride >= w)
    return set_error(BCOSK_EINVAL, "nhwc_scatter_nchw_f32: bad argument");
  const long long n = (long long)nb * oh * ow;
  if (dtype == BCOSK_DTYPE_F16)
    nhwc_scatter_nchw_f32_kernel<__half><<<nblocks(n, 128), 128, 0, S2(stream)>>>(y, y_f32, nb, c, oh, ow, ld, planes, out, h, w, stride);
  else
    nhwc_scatter_nchw_f32_kernel<__nv_bfloat16><<<nblocks(n, 128), 128, 0, S2(stream)>>>(y, y_f32, nb, c, oh, ow, ld, planes, out, h, w, stride);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_pixel_sqsum_nchw_f32(const float* x, int32_t nb, int32_t c, int64_t hw, float* sq, void* stream) {
  using namespace bcosk;
  if (!x || !sq || nb < 1 || c < 1 || hw < 1) return set_error(BCOSK_EINVAL, "pixel_sqsum_nchw_f32: bad argument");
  const long long total = (long long)nb * hw;
  pixel_sqsum_nchw_f32_kernel<<<nblocks(total, 256), 256, 0, S2(stream)>>>(x, c, hw, total, sq);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

// ------------------------------------------------------------------------------------------------
// Seed of a fused trunk's explanation pass from a gradient computed outside the plan (the attention-pool head of the CLIP
// encoders runs on the module-level path): g NCHW fp32 -> the last block's two gradient tensors in NHWC 16-bit planes,
//   out1 = g * seed_scale * mul1               (gradient x gain of the block's last conv: its `ghat`)
//   out2 = g * seed_scale [* mul2], masked by the block's ReLU bits   (identity / downsample branch)
// Thread = pixel, 8 channels per step (the same access pattern as nchw_to_nhwc16).
// ------------------------------------------------------------------------------------------------
namespace bcosk {

template <typename T>
__device__ __forceinline__ void load8_side(const void* base, bool f32, long long row, int ld, int ch0, float (&m)[8]) {
  if (f32) {
    const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + row * ld + ch0);
    const float4 a = __ldg(q), b = __ldg(q + 1);
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(base) + row * ld + ch0));
    float2 f;
    f = Cvt<T>::unpack2(u.x); m[0] = f.x; m[1] = f.y;
    f = Cvt<T>::unpack2(u.y); m[2] = f.x; m[3] = f.y;
    f = Cvt<T>::unpack2(u.z); m[4] = f.x; m[5] = f.y;
    f = Cvt<T>::unpack2(u.w); m[6] = f.x; m[7] = f.y;
  }
}

template <typename T>
__device__ __forceinline__ void store8_planes(T* dst, int planes, int plane_stride, const float (&v)[8]) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = v[i];
  for (int pl = 0; pl < planes; ++pl) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      w[k] = Cvt<T>::pack2(r[2 * k], r[2 * k + 1]);
      const float2 q = Cvt<T>::unpack2(w[k]);
      r[2 * k] -= q.x; r[2 * k + 1] -= q.y;
    }
    *reinterpret_cast<uint4*>(dst + (long long)pl * plane_stride) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <typename T>
__global__ void seed_from_nchw_kernel(const float* __restrict__ g, int nb, int c, long long hw, float seed_scale,
                                      const void* __restrict__ mul1, int mul1_f32, T* __restrict__ out1,
                                      const uint32_t* __restrict__ mask2, const void* __restrict__ mul2, int mul2_f32,
                                      T* __restrict__ out2, int planes) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)nb * hw) return;
  const long long img = pix / hw, sp = pix - img * hw;
  const float* src = g + img * c * hw + sp;
  const int ld = planes * c;
  for (int ch0 = 0; ch0 < c; ch0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(src + (long long)(ch0 + i) * hw) * seed_scale;
    if (out1 != nullptr) {
      float o[8];
      if (mul1 != nullptr) {
        float m[8];
        load8_side<T>(mul1, mul1_f32 != 0, pix, c, ch0, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = v[i] * m[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = v[i];
      }
      store8_planes<T>(out1 + pix * ld + ch0, planes, c, o);
    }
    if (out2 != nullptr) {
      float o[8];
      if (mul2 != nullptr) {
        float m[8];
        load8_side<T>(mul2, mul2_f32 != 0, pix, c, ch0, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = v[i] * m[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = v[i];
      }
      if (mask2 != nullptr) {
        const uint32_t mb = __ldg(mask2 + pix * ((c + 31) / 32) + (ch0 >> 5)) >> (ch0 & 31);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ((mb >> i) & 1u) ? o[i] : 0.f;
      }
      store8_planes<T>(out2 + pix * ld + ch0, planes, c, o);
    }
  }
}

}  // namespace bcosk

extern "C" int bcosk_seed_from_nchw(const float* g, int32_t nb, int32_t c, int32_t h, int32_t w, float seed_scale, const void* mul1,
                                    int32_t mul1_f32, void* out1, const uint32_t* mask2, const void* mul2, int32_t mul2_f32,
                                    void* out2, int32_t planes, int32_t dtype, void* stream) {
  using namespace bcosk;
  if (!g || (!out1 && !out2) || nb < 1 || c < 8 || c % 8 || planes < 1 || planes > 3)
    return set_error(BCOSK_EINVAL, "seed_from_nchw: bad argument");
  const long long hw = (long long)h * w, n = (long long)nb * hw;
  if (dtype == BCOSK_DTYPE_BF16)
    seed_from_nchw_kernel<__nv_bfloat16><<<nblocks(n, 128), 128, 0, S2(stream)>>>(
        g, nb, c, hw, seed_scale, mul1, mul1_f32, reinterpret_cast<__nv_bfloat16*>(out1), mask2, mul2, mul2_f32,
        reinterpret_cast<__nv_bfloat16*>(out2), planes);
  else if (dtype == BCOSK_DTYPE_F16)
    seed_from_nchw_kernel<__half><<<nblocks(n, 128), 128, 0, S2(stream)>>>(g, nb, c, hw, seed_scale, mul1, mul1_f32,
                                                                         reinterpret_cast<__half*>(out1), mask2, mul2, mul2_f32,
                                                                         reinterpret_cast<__half*>(out2), planes);
  else
    return set_error(BCOSK_EINVAL, "seed_from_nchw: dtype");
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

// ------------------------------------------------------------------------------------------------
// MaxOut for the module-level path (bcosconv2d.py:166-170, bcoslinear.py:107-110): the linear map has O*M units,
// unit c = o*M + m; the output keeps the largest of the M candidates (first one on ties, like torch.max) and the
// B-cos scale is computed from it.  The explanation gradient reaches only that unit.
// ------------------------------------------------------------------------------------------------
namespace bcosk {

__global__ void maxout_bcos_fwd_kernel(const float* __restrict__ lin, const float* __restrict__ inv_norm, long long rows, int o,
                                       int m, int scale_mode, float b_exp, float* __restrict__ y, float* __restrict__ gain,
                                       uint8_t* __restrict__ amax) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * o) return;
  const long long r = idx / o;
  const float* src = lin + idx * m;            // (r * o + oo) * m
  float best = __ldg(src);
  int bi = 0;
  for (int k = 1; k < m; ++k) {
    const float v = __ldg(src + k);
    if (v > best) { best = v; bi = k; }
  }
  float s = 1.f;
  if (scale_mode == BCOSK_SCALE_B2) s = fabsf(best) * __ldg(inv_norm + r);
  else if (scale_mode == BCOSK_SCALE_POW) s = powf(fabsf(best) * __ldg(inv_norm + r) + 1e-6f, b_exp - 1.f);
  y[idx] = best * s;
  if (gain != nullptr) gain[idx] = s;
  if (amax != nullptr) amax[idx] = (uint8_t)bi;
}

// dst[r, pl, oo*m + k] = (k == amax[r, oo]) ? src[r, pl, oo] : 0      (16-bit planes)
template <typename T>
__global__ void maxout_scatter_kernel(const T* __restrict__ src, const uint8_t* __restrict__ amax, long long rows, int o,
                                      int m, int planes, T* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = rows * planes * o * m;
  if (idx >= total) return;
  const int om = o * m;
  const int c = (int)(idx % om);
  const long long rp = idx / om;               // r * planes + pl
  const long long r = rp / planes;
  const int oo = c / m, k = c - oo * m;
  dst[idx] = (k == (int)__ldg(amax + r * o + oo)) ? src[rp * o + oo] : T(0.f);
}

}  // namespace bcosk

extern "C" int bcosk_maxout_bcos_fwd(const float* lin, const float* inv_norm, int64_t rows, int32_t o, int32_t m,
                                     int32_t scale_mode, float b_exp, float* y, float* gain, uint8_t* amax, void* stream) {
  using namespace bcosk;
  if (!lin || !y || rows < 1 || o < 1 || m < 1 || m > 255) return set_error(BCOSK_EINVAL, "maxout_bcos_fwd: bad argument");
  if (scale_mode != BCOSK_SCALE_NONE && !inv_norm) return set_error(BCOSK_EINVAL, "maxout_bcos_fwd: inv_norm required");
  const long long n = rows * o;
  maxout_bcos_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      lin, inv_norm, rows, o, m, scale_mode, b_exp, y, gain, amax);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_maxout_scatter(const void* src, const uint8_t* amax, int64_t rows, int32_t o, int32_t m, int32_t planes,
                                    int32_t dtype, void* dst, void* stream) {
  using namespace bcosk;
  if (!src || !amax || !dst || rows < 1 || o < 1 || m < 1 || planes < 1) return set_error(BCOSK_EINVAL, "maxout_scatter: bad argument");
  const long long n = rows * planes * o * m;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == BCOSK_DTYPE_BF16)
    maxout_scatter_kernel<__nv_bfloat16><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(src), amax, rows, o, m, planes, reinterpret_cast<__nv_bfloat16*>(dst));
  else if (dtype == BCOSK_DTYPE_F16)
    maxout_scatter_kernel<__half><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const __half*>(src), amax, rows, o, m, planes, reinterpret_cast<__half*>(dst));
  else
    return set_error(BCOSK_EINVAL, "maxout_scatter: dtype");
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}
