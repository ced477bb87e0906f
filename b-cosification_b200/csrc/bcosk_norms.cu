// bcosk_norms.cu -- group / position normalisation of NCHW fp32 tensors with detachable statistics (module-level path).
// Reference: bcos/modules/norms/uncentered_norms/groupnorm_uncentered.py:21-61 (group_norm_uncentered),
// bcos/modules/norms/centered_norms.py:93-138 (DetachableGroupNorm2d), :251-297 (DetachablePositionNorm2d),
// bcos/modules/norms/uncentered_norms/posnorm_uncentered.py:39-58 (PositionNormUncentered2d).
// All four divide by sqrt(var + eps) with var the CENTRED biased variance; the "centred" ones also subtract the mean.
// In explanation mode the variance is a constant (detached), the mean stays in the graph.
// Bandwidth kernels: the statistics passes re-read the group / pixel column from L2, 16-byte accesses where the
// geometry allows, consecutive lanes on consecutive addresses.
#include <cooperative_groups.h>

#include "../../include/bcosk.h"
#include "bcosk_common.cuh"
#include "bcosk_host.h"

namespace bcosk {
namespace cg = cooperative_groups;

constexpr int GN_THREADS = 512;

// block-wide sum, result broadcast to every thread; `red` holds one float per warp
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();                       // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = lane < nw ? red[lane] : 0.f;
  return warp_sum(t);
}

// ---------------------------------------------------------------- group norm: one CTA per (image, group)
// The group is the contiguous run x[(n*C + g*Cg)*HW ... + Cg*HW).  VEC = 4 needs HW % 4 == 0 (then a float4 never
// straddles two channels and every run start is 16-byte aligned).
template <int VEC>
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_fwd_kernel(const float* __restrict__ x, int c, int hw, int groups, const float* __restrict__ w,
                     const float* __restrict__ b, float eps, int centred, float* __restrict__ y, float* __restrict__ rstd) {
  __shared__ float red[GN_THREADS / 32];
  const int cg = c / groups;
  const int g = blockIdx.x % groups;
  const long long len = (long long)cg * hw;
  const float* src = x + (long long)blockIdx.x * len;
  float* dst = y + (long long)blockIdx.x * len;
  const long long nv = len / VEC;
  float s = 0.f;
  if (VEC == 4) {
    for (long long i = threadIdx.x; i < nv; i += GN_THREADS) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      s += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (long long i = threadIdx.x; i < nv; i += GN_THREADS) s += __ldg(src + i);
  }
  const float mean = block_sum(s, red) / (float)len;
  float q = 0.f;
  if (VEC == 4) {
    for (long long i = threadIdx.x; i < nv; i += GN_THREADS) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
      q = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, q))));
    }
  } else {
    for (long long i = threadIdx.x; i < nv; i += GN_THREADS) {
      const float a = __ldg(src + i) - mean;
      q = fmaf(a, a, q);
    }
  }
  const float var = block_sum(q, red) / (float)len;     // biased, like var(unbiased=False)
  const float r = 1.0f / sqrtf(var + eps);
  const float sub = centred ? mean : 0.f;
  if (rstd != nullptr && threadIdx.x == 0) rstd[blockIdx.x] = r;
  if (VEC == 4) {
    const int hv = hw / 4;
    for (long long i = threadIdx.x; i < nv; i += GN_THREADS) {
      const int ch = g * cg + (int)((unsigned)i / (unsigned)hv);
      const float ww = (w ? __ldg(w + ch) : 1.f) * r, bb = b ? __ldg(b + ch) : 0.f;
      float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      v.x = fmaf(v.x - sub, ww, bb); v.y = fmaf(v.y - sub, ww, bb);
      v.z = fmaf(v.z - sub, ww, bb); v.w = fmaf(v.w - sub, ww, bb);
      reinterpret_cast<float4*>(dst)[i] = v;
    }
  } else {
    for (long long i = threadIdx.x; i < nv; i += GN_THREADS) {
      const int ch = g * cg + (int)((unsigned)i / (unsigned)hw);
      const float ww = (w ? __ldg(w + ch) : 1.f) * r, bb = b ? __ldg(b + ch) : 0.f;
      dst[i] = fmaf(__ldg(src + i) - sub, ww, bb);
    }
  }
}

// explanation backward (variance detached): uncentred gx = w*gy*rstd ; centred gx = (w*gy - mean_group(w*gy)) * rstd
template <int VEC>
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_explain_bwd_kernel(const float* __restrict__ gy, int c, int hw, int groups, const float* __restrict__ w,
                             const float* __restrict__ rstd, int centred, float* __restrict__ gx) {
  __shared__ float red[GN_THREADS / 32];
  const int cg = c / groups;
  const int g = blockIdx.x % groups;
  const long long len = (long long)cg * hw;
  const float* src = gy + (long long)blockIdx.x * len;
  float* dst = gx + (long long)blockIdx.x * len;
  const long long nv = len / VEC;
  const int hv = hw / VEC;
  float m = 0.f;
  if (centred) {
    float s = 0.f;
    for (long long i = threadIdx.x; i < nv; i += GN_THREADS) {
      const float ww = w ? __ldg(w + g * cg + (int)((unsigned)i / (unsigned)hv)) : 1.f;
      if (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        s = fmaf((v.x + v.y) + (v.z + v.w), ww, s);
      } else {
        s = fmaf(__ldg(src + i), ww, s);
      }
    }
    m = block_sum(s, red) / (float)len;
  }
  const float r = __ldg(rstd + blockIdx.x);
  for (long long i = threadIdx.x; i < nv; i += GN_THREADS) {
    const float ww = w ? __ldg(w + g * cg + (int)((unsigned)i / (unsigned)hv)) : 1.f;
    if (VEC == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      v.x = (v.x * ww - m) * r; v.y = (v.y * ww - m) * r; v.z = (v.z * ww - m) * r; v.w = (v.w * ww - m) * r;
      reinterpret_cast<float4*>(dst)[i] = v;
    } else {
      dst[i] = (__ldg(src + i) * ww - m) * r;
    }
  }
}

// ---------------------------------------------------------------- large groups: one thread-block CLUSTER per (image, group)
// A group of several MB (GN-LayerNorm of a 256 x 56 x 56 map = 3.2 MB) read three times by ONE CTA streams from HBM three
// times once all resident groups exceed L2.  Here the group is split over the S CTAs of a cluster (2 CTAs per SM): every
// CTA keeps the head of its chunk in shared memory (GNC_SMEM_FLOATS), re-reads only the tail from L2, and the partial sums
// are exchanged through distributed shared memory.  len % (4*S) == 0 and hw % 4 == 0 are required (host checks).
constexpr int GNC_SMEM_FLOATS = 24 * 1024;      // 96 KB: two CTAs per SM, so one CTA's streaming pass overlaps its neighbour's shared-memory passes

__device__ __forceinline__ float cluster_sum(float v, float* red, float* slot, cg::cluster_group& cluster) {
  const float local = block_sum(v, red);
  if (threadIdx.x == 0) *slot = local;
  cluster.sync();
  float tot = 0.f;
  const unsigned n = cluster.num_blocks();
  for (unsigned r = 0; r < n; ++r) tot += *cluster.map_shared_rank(slot, r);
  return tot;
}

// BWD == false: forward (mean, centred variance, write); BWD == true: explanation backward (x = gy, rstd read)
// 32-bit element indices (the host keeps len < 2^31); the streaming pass issues GNC_UNROLL independent 16-byte loads per
// thread before consuming them (one CTA per SM: the bytes in flight have to come from instruction-level parallelism).
constexpr int GNC_THREADS = 512;
constexpr int GNC_UNROLL = 4;

__device__ __forceinline__ float4 scale4(float4 v, float s) { v.x *= s; v.y *= s; v.z *= s; v.w *= s; return v; }

template <bool BWD>
__global__ void __launch_bounds__(GNC_THREADS, 2)
groupnorm_cluster_kernel(const float* __restrict__ x, int c, int hw, int groups, const float* __restrict__ w,
                         const float* __restrict__ b, float eps, int centred, float* __restrict__ y, float* __restrict__ rstd) {
  extern __shared__ float4 cache[];
  __shared__ float red[GNC_THREADS / 32];
  __shared__ float slots[2];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned S = cluster.num_blocks(), rank = cluster.block_rank();
  const unsigned grp = blockIdx.x / S;                  // (image, group)
  const unsigned cg_ = (unsigned)(c / groups);
  const unsigned g = grp % (unsigned)groups;
  const unsigned len = cg_ * (unsigned)hw;
  const unsigned nv = len / 4 / S;                      // float4 per CTA
  const unsigned v0 = rank * nv;                        // first float4 of this CTA inside the group
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)grp * len) + v0;
  float4* dst = reinterpret_cast<float4*>(y + (size_t)grp * len) + v0;
  const unsigned ncache = nv < GNC_SMEM_FLOATS / 4 ? nv : GNC_SMEM_FLOATS / 4;
  const unsigned hv = (unsigned)hw / 4;
  const float* wg = w ? w + g * cg_ : nullptr;
  float s = 0.f;
  for (unsigned i0 = threadIdx.x; i0 < nv; i0 += GNC_THREADS * GNC_UNROLL) {
    float4 v[GNC_UNROLL];
#pragma unroll
    for (int u = 0; u < GNC_UNROLL; ++u) {
      const unsigned i = i0 + u * GNC_THREADS;
      v[u] = i < nv ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < GNC_UNROLL; ++u) {
      const unsigned i = i0 + u * GNC_THREADS;
      if (BWD && wg != nullptr && i < nv) v[u] = scale4(v[u], __ldg(wg + (v0 + i) / hv));
      if (i < ncache) cache[i] = v[u];
      s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
  }
  float mean = 0.f, r;
  if (!BWD || centred) mean = cluster_sum(s, red, &slots[0], cluster) / (float)len;
  if (!BWD) {
    float q = 0.f;
    for (unsigned i = threadIdx.x; i < nv; i += GNC_THREADS) {
      const float4 v = i < ncache ? cache[i] : __ldg(src + i);
      const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
      q = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, q))));
    }
    const float var = cluster_sum(q, red, &slots[1], cluster) / (float)len;
    r = 1.0f / sqrtf(var + eps);
    if (rstd != nullptr && rank == 0 && threadIdx.x == 0) rstd[grp] = r;
  } else {
    r = __ldg(rstd + grp);
  }
  const float sub = centred ? mean : 0.f;
  for (unsigned i = threadIdx.x; i < nv; i += GNC_THREADS) {
    const unsigned ch = (v0 + i) / hv;                  // channel inside the group
    float4 v;
    if (i < ncache) {
      v = cache[i];
    } else {
      v = __ldg(src + i);
      if (BWD && wg != nullptr) v = scale4(v, __ldg(wg + ch));
    }
    if (BWD) {
      v.x = (v.x - sub) * r; v.y = (v.y - sub) * r; v.z = (v.z - sub) * r; v.w = (v.w - sub) * r;
    } else {
      const float ww = (wg ? __ldg(wg + ch) : 1.f) * r, bb = b ? __ldg(b + g * cg_ + ch) : 0.f;
      v.x = fmaf(v.x - sub, ww, bb); v.y = fmaf(v.y - sub, ww, bb);
      v.z = fmaf(v.z - sub, ww, bb); v.w = fmaf(v.w - sub, ww, bb);
    }
    dst[i] = v;
  }
  cluster.sync();                                       // nobody leaves while a peer may still read its slots
}

// ---------------------------------------------------------------- position norm: statistics over C per pixel
// CTA = 32 pixels (threadIdx.x, consecutive addresses) x PN_SLICES channel slices (threadIdx.y); grid (ceil(HW/32), N).
constexpr int PN_SLICES = 16;

__device__ __forceinline__ float slice_sum(float v, float (*red)[33]) {
  __syncthreads();
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < PN_SLICES; ++k) t += red[k][threadIdx.x];
  return t;
}

// backward == 0: y = w[c]*(x - centred*mean)*rstd + b[c], rstd[n*HW + p] saved
// backward == 1: x is the incoming gradient, y = (w[c]*x - centred*mean_c(w*x)) * rstd[n*HW + p]  (variance detached)
// NC > 0: every thread keeps its NC = ceil(C / PN_SLICES) channel values in registers, so the tensor is read ONCE
// (C <= 16*NC); NC == 0: any C, the column is re-read (from L1/L2) for every pass.
template <bool BWD, int NC>
__global__ void __launch_bounds__(32 * PN_SLICES, NC <= 16 ? 3 : 1)
positionnorm_kernel(const float* __restrict__ x, int c, int hw, const float* __restrict__ w, const float* __restrict__ b,
                    float eps, int centred, float* __restrict__ y, float* __restrict__ rstd) {
  __shared__ float red[PN_SLICES][33];
  const int p = blockIdx.x * 32 + threadIdx.x;
  const bool ok = p < hw;
  const long long base = (long long)blockIdx.y * c * hw + p;
  const float* src = x + base;
  float* dst = y + base;
  constexpr int NR = NC > 0 ? NC : 1;
  float val[NR];
  float s = 0.f;
  if (NC > 0) {
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int ch = threadIdx.y + k * PN_SLICES;
      float v = (ok && ch < c) ? __ldg(src + (long long)ch * hw) : 0.f;
      if (BWD && w != nullptr && ch < c) v *= __ldg(w + ch);
      val[k] = v;
      s += v;
    }
  } else if (ok && (!BWD || centred)) {
    for (int ch = threadIdx.y; ch < c; ch += PN_SLICES) {
      const float v = __ldg(src + (long long)ch * hw);
      s += BWD ? v * (w ? __ldg(w + ch) : 1.f) : v;
    }
  }
  float mean = 0.f, r;
  if (!BWD || centred) mean = slice_sum(s, red) / (float)c;
  if (!BWD) {
    float q = 0.f;
    if (NC > 0) {
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const float a = (threadIdx.y + k * PN_SLICES < c) ? val[k] - mean : 0.f;
        q = fmaf(a, a, q);
      }
    } else if (ok) {
      for (int ch = threadIdx.y; ch < c; ch += PN_SLICES) {
        const float a = __ldg(src + (long long)ch * hw) - mean;
        q = fmaf(a, a, q);
      }
    }
    const float var = slice_sum(q, red) / (float)c;
    r = 1.0f / sqrtf(var + eps);
    if (ok && rstd != nullptr && threadIdx.y == 0) rstd[(long long)blockIdx.y * hw + p] = r;
  } else {
    r = ok ? __ldg(rstd + (long long)blockIdx.y * hw + p) : 0.f;
  }
  if (!ok) return;
  const float sub = centred ? mean : 0.f;
  if (NC > 0) {
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int ch = threadIdx.y + k * PN_SLICES;
      if (ch < c) {
        if (BWD) dst[(long long)ch * hw] = (val[k] - sub) * r;
        else dst[(long long)ch * hw] = fmaf(val[k] - sub, (w ? __ldg(w + ch) : 1.f) * r, b ? __ldg(b + ch) : 0.f);
      }
    }
  } else {
    for (int ch = threadIdx.y; ch < c; ch += PN_SLICES) {
      const float v = __ldg(src + (long long)ch * hw);
      const float ww = w ? __ldg(w + ch) : 1.f;
      if (BWD) dst[(long long)ch * hw] = (v * ww - sub) * r;
      else dst[(long long)ch * hw] = fmaf(v - sub, ww * r, b ? __ldg(b + ch) : 0.f);
    }
  }
}

template <bool BWD>
static void launch_positionnorm(dim3 grid, dim3 block, cudaStream_t st, const float* x, int c, int hw, const float* w,
                                const float* b, float eps, int centred, float* y, float* rstd) {
  if (BWD && !centred) positionnorm_kernel<BWD, 0><<<grid, block, 0, st>>>(x, c, hw, w, b, eps, centred, y, rstd);   // one streaming pass
  else if (c <= 4 * PN_SLICES) positionnorm_kernel<BWD, 4><<<grid, block, 0, st>>>(x, c, hw, w, b, eps, centred, y, rstd);
  else if (c <= 16 * PN_SLICES) positionnorm_kernel<BWD, 16><<<grid, block, 0, st>>>(x, c, hw, w, b, eps, centred, y, rstd);
  else if (c <= 64 * PN_SLICES) positionnorm_kernel<BWD, 64><<<grid, block, 0, st>>>(x, c, hw, w, b, eps, centred, y, rstd);
  else positionnorm_kernel<BWD, 0><<<grid, block, 0, st>>>(x, c, hw, w, b, eps, centred, y, rstd);
}

}  // namespace bcosk

using namespace bcosk;
static inline cudaStream_t S4(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Large groups (> GNC_MIN_BYTES) go to the cluster kernel: S CTAs per group, S the smallest power of two <= 8 whose chunk
// fits the shared-memory cache (or 8), provided the float4 count divides evenly.
constexpr long long GNC_MIN_BYTES = 512 * 1024;
template <bool BWD>
static int launch_groupnorm_cluster(const float* x, int nb, int c, int hw, int groups, const float* w, const float* b, float eps,
                                    int centred, float* y, float* rstd, cudaStream_t st, bool* done) {
  *done = false;
  const long long len = (long long)(c / groups) * hw;
  if (len * 4 < GNC_MIN_BYTES || len >= (1LL << 31) || hw % 4 != 0 || !aligned16(x) || !aligned16(y)) return BCOSK_OK;
  int S = 1;
  while (S < 8 && len / S > GNC_SMEM_FLOATS) S *= 2;
  if (len % (4LL * S) != 0) return BCOSK_OK;
  const void* fn = (const void*)groupnorm_cluster_kernel<BWD>;
  const size_t smem = (size_t)GNC_SMEM_FLOATS * sizeof(float);
  BCOSK_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per (function, device)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((long long)nb * groups * S));
  cfg.blockDim = dim3(GNC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  BCOSK_CUDA_CHECK(cudaLaunchKernelEx(&cfg, groupnorm_cluster_kernel<BWD>, x, c, hw, groups, w, b, eps, centred, y, rstd));
  *done = true;
  return BCOSK_OK;
}


extern "C" int bcosk_groupnorm_fwd(const float* x, int32_t nb, int32_t c, int64_t hw, int32_t groups, const float* w,
                                   const float* b, float eps, int32_t centred, float* y, float* rstd, void* stream) {
  if (!x || !y || nb < 0 || c < 1 || hw < 1 || groups < 1) return set_error(BCOSK_EINVAL, "groupnorm_fwd: bad argument");
  if (c % groups != 0) return set_error(BCOSK_EINVAL, "groupnorm_fwd: channels %d not divisible by groups %d", c, groups);
  if ((int64_t)(c / groups) * hw > 0x7fffffffLL || (int64_t)nb * groups > 0x7fffffffLL)
    return set_error(BCOSK_EUNSUPPORTED, "groupnorm_fwd: too large");
  if (nb == 0) return BCOSK_OK;
  bool done = false;
  const int rc = launch_groupnorm_cluster<false>(x, nb, c, (int)hw, groups, w, b, eps, centred, y, rstd, S4(stream), &done);
  if (rc != BCOSK_OK || done) return rc;
  const unsigned grid = (unsigned)(nb * groups);
  if (hw % 4 == 0 && aligned16(x) && aligned16(y))
    groupnorm_fwd_kernel<4><<<grid, GN_THREADS, 0, S4(stream)>>>(x, c, (int)hw, groups, w, b, eps, centred, y, rstd);
  else
    groupnorm_fwd_kernel<1><<<grid, GN_THREADS, 0, S4(stream)>>>(x, c, (int)hw, groups, w, b, eps, centred, y, rstd);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_groupnorm_explain_bwd(const float* gy, int32_t nb, int32_t c, int64_t hw, int32_t groups, const float* w,
                                           const float* rstd, int32_t centred, float* gx, void* stream) {
  if (!gy || !gx || !rstd || nb < 0 || c < 1 || hw < 1 || groups < 1) return set_error(BCOSK_EINVAL, "groupnorm_explain_bwd: bad argument");
  if (c % groups != 0) return set_error(BCOSK_EINVAL, "groupnorm_explain_bwd: channels %d not divisible by groups %d", c, groups);
  if ((int64_t)(c / groups) * hw > 0x7fffffffLL || (int64_t)nb * groups > 0x7fffffffLL)
    return set_error(BCOSK_EUNSUPPORTED, "groupnorm_explain_bwd: too large");
  if (nb == 0) return BCOSK_OK;
  bool done = false;
  const int rc = launch_groupnorm_cluster<true>(gy, nb, c, (int)hw, groups, w, nullptr, 0.f, centred, gx,
                                                const_cast<float*>(rstd), S4(stream), &done);
  if (rc != BCOSK_OK || done) return rc;
  const unsigned grid = (unsigned)(nb * groups);
  if (hw % 4 == 0 && aligned16(gy) && aligned16(gx))
    groupnorm_explain_bwd_kernel<4><<<grid, GN_THREADS, 0, S4(stream)>>>(gy, c, (int)hw, groups, w, rstd, centred, gx);
  else
    groupnorm_explain_bwd_kernel<1><<<grid, GN_THREADS, 0, S4(stream)>>>(gy, c, (int)hw, groups, w, rstd, centred, gx);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_positionnorm_fwd(const float* x, int32_t nb, int32_t c, int64_t hw, const float* w, const float* b, float eps,
                                      int32_t centred, float* y, float* rstd, void* stream) {
  if (!x || !y || nb < 0 || c < 1 || hw < 1) return set_error(BCOSK_EINVAL, "positionnorm_fwd: bad argument");
  if (hw > 0x7fffffffLL || nb > 65535) return set_error(BCOSK_EUNSUPPORTED, "positionnorm_fwd: too large");
  if (nb == 0) return BCOSK_OK;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)nb), block(32, PN_SLICES);
  launch_positionnorm<false>(grid, block, S4(stream), x, c, (int)hw, w, b, eps, centred, y, rstd);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_positionnorm_explain_bwd(const float* gy, int32_t nb, int32_t c, int64_t hw, const float* w, const float* rstd,
                                              int32_t centred, float* gx, void* stream) {
  if (!gy || !gx || !rstd || nb < 0 || c < 1 || hw < 1) return set_error(BCOSK_EINVAL, "positionnorm_explain_bwd: bad argument");
  if (hw > 0x7fffffffLL || nb > 65535) return set_error(BCOSK_EUNSUPPORTED, "positionnorm_explain_bwd: too large");
  if (nb == 0) return BCOSK_OK;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)nb), block(32, PN_SLICES);
  launch_positionnorm<true>(grid, block, S4(stream), gy, c, (int)hw, w, nullptr, 0.f, centred, gx, const_cast<float*>(rstd));
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}
