// bcosk_dense.cu -- bandwidth kernels of the fused DenseNet plan (engine/densenet.py) for sm_100a.
//
// A dense block keeps ONE feature tensor F [pixels][planes * C_total] (NHWC 16-bit precision planes); every layer's 3x3 conv
// writes its `growth` new channels straight into its slice of F (bcosk_igemm with a column offset), so the torch.cat of the
// reference (torchvision densenet.py _DenseLayer / _DenseBlock, B-cosified by bcosify.py:74-113) never copies anything.
// Every consumer normalises the channels it reads with ITS OWN uncentred BN (batchnorm_uncentered.py:49-58, eval mode) + ReLU:
//
//   dense_bn_relu_fwd        t = relu(F[:, :c] * alpha) as dense plane rows + sum t^2 per pixel + ReLU bits
//   dense_bn_relu_bwd        G[:, :c] (+)= g * alpha * mask   (explanation / plain backward; G: fp32 feature-gradient tensor)
//   dense_slice_cast         ghat = G[:, col0 : col0 + c] (* gain) -> one 16-bit plane, the A operand of a data gradient
//   copy_rows_2d             strided device copy (pool output -> the first channels of the next block's F)
//
// One warp per pixel row, 16-byte vectors, shuffle reductions, fp32 arithmetic.
#include "../../include/bcosk.h"
#include "bcosk_common.cuh"
#include "bcosk_host.h"

namespace bcosk {

static inline cudaStream_t SD(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__device__ __forceinline__ void d_unpack8(const uint4& u, float (&f)[8]) {
  float2 q;
  q = Cvt<T>::unpack2(u.x); f[0] = q.x; f[1] = q.y;
  q = Cvt<T>::unpack2(u.y); f[2] = q.x; f[3] = q.y;
  q = Cvt<T>::unpack2(u.z); f[4] = q.x; f[5] = q.y;
  q = Cvt<T>::unpack2(u.w); f[6] = q.x; f[7] = q.y;
}
__device__ __forceinline__ void d_load8_f32(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// x rows: [rows][planes * x_pstride] (channels [0, c) of every plane are read); y rows: [rows][planes * c] dense.
template <typename T>
__global__ void dense_bn_relu_fwd_kernel(const T* __restrict__ x, long long rows, int c, int planes, int x_ld, int x_pstride,
                                         const float* __restrict__ alpha, int relu, T* __restrict__ y, float* __restrict__ sq,
                                         uint32_t* __restrict__ maskbits) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, nvec = c >> 3;
  const T* xr = x + row * (long long)x_ld;
  T* yr = y + row * (long long)planes * c;
  float sacc = 0.f;
#pragma unroll 2
  for (int v0 = 0; v0 < nvec; v0 += 32) {        // whole warp iterates together (shuffles below)
    const int v = v0 + lane;
    uint32_t bits = 0;
    if (v < nvec) {
      float f[8];
      uint4 u[3];                                  // all plane loads in flight before the first use
      u[0] = __ldg(reinterpret_cast<const uint4*>(xr + v * 8));
#pragma unroll
      for (int pl = 1; pl < 3; ++pl)
        u[pl] = pl < planes ? __ldg(reinterpret_cast<const uint4*>(xr + (size_t)pl * x_pstride + v * 8)) : make_uint4(0u, 0u, 0u, 0u);
      d_unpack8<T>(u[0], f);
#pragma unroll
      for (int pl = 1; pl < 3; ++pl) {
        float g[8];
        d_unpack8<T>(u[pl], g);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += g[i];
      }
      float a[8];
      d_load8_f32(alpha + v * 8, a);
      float r[8], acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float t = f[i] * a[i];
        const bool pos = !relu || t > 0.f;
        bits |= (pos ? 1u : 0u) << i;
        r[i] = pos ? t : 0.f;
        acc[i] = 0.f;
      }
      for (int pl = 0; pl < planes; ++pl) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          w[k] = Cvt<T>::pack2(r[2 * k], r[2 * k + 1]);
          const float2 q = Cvt<T>::unpack2(w[k]);
          r[2 * k] -= q.x; r[2 * k + 1] -= q.y;
          acc[2 * k] += q.x; acc[2 * k + 1] += q.y;
        }
        *reinterpret_cast<uint4*>(yr + (size_t)pl * c + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sacc = fmaf(acc[i], acc[i], sacc);
    }
    if (maskbits != nullptr) {
      // lanes 4k .. 4k+3 hold the four 8-channel groups of mask word k of this 256-channel step
      uint32_t wbits = bits << ((lane & 3) * 8);
      wbits |= __shfl_xor_sync(0xffffffffu, wbits, 1);
      wbits |= __shfl_xor_sync(0xffffffffu, wbits, 2);
      const int word = (v0 >> 2) + (lane >> 2);
      if ((lane & 3) == 0 && word < (c + 31) / 32) maskbits[row * ((c + 31) / 32) + word] = wbits;
    }
  }
  sacc = warp_sum(sacc);
  if (lane == 0 && sq != nullptr) sq[row] = sacc;
}

// G rows: [rows][g_ld] fp32, channels [0, c) updated: G = (accumulate ? G : 0) + g * alpha * mask.  g: [rows][c] fp32 or 16-bit.
template <typename T>
__global__ void dense_bn_relu_bwd_kernel(const void* __restrict__ g, int g_f32, long long rows, int c, const float* __restrict__ alpha,
                                         const uint32_t* __restrict__ maskbits, float* __restrict__ G, int g_ld, int accumulate) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (row, 8-channel group)
  const int nvec = c >> 3;
  if (idx >= rows * nvec) return;
  const long long row = idx / nvec;
  const int v = (int)(idx - row * nvec);
  float f[8];
  if (g_f32) d_load8_f32(reinterpret_cast<const float*>(g) + row * c + v * 8, f);
  else d_unpack8<T>(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(g) + row * c + v * 8)), f);
  float a[8];
  d_load8_f32(alpha + v * 8, a);
  uint32_t mb = 0xffu;
  if (maskbits != nullptr) mb = (__ldg(maskbits + row * ((c + 31) / 32) + (v >> 2)) >> ((v & 3) * 8)) & 0xffu;
  float* gp = G + row * (long long)g_ld + v * 8;
  float o[8];
  if (accumulate) {
    const float4 p0 = *reinterpret_cast<const float4*>(gp), p1 = *reinterpret_cast<const float4*>(gp + 4);
    o[0] = p0.x; o[1] = p0.y; o[2] = p0.z; o[3] = p0.w; o[4] = p1.x; o[5] = p1.y; o[6] = p1.z; o[7] = p1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] += ((mb >> i) & 1u) ? f[i] * a[i] : 0.f;
  reinterpret_cast<float4*>(gp)[0] = make_float4(o[0], o[1], o[2], o[3]);
  reinterpret_cast<float4*>(gp)[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// out [rows][c] one 16-bit plane = G[rows][col0 : col0 + c] (fp32, row pitch g_ld) * gain[rows][c] (optional, 16-bit or fp32) * scale
template <typename T>
__global__ void dense_slice_cast_kernel(const float* __restrict__ G, int g_ld, int col0, long long rows, int c, const void* __restrict__ gain,
                                        int gain_f32, float scale, T* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nvec = c >> 3;
  if (idx >= rows * nvec) return;
  const long long row = idx / nvec;
  const int v = (int)(idx - row * nvec);
  float f[8];
  d_load8_f32(G + row * (long long)g_ld + col0 + v * 8, f);
  if (gain != nullptr) {
    float gn[8];
    if (gain_f32) d_load8_f32(reinterpret_cast<const float*>(gain) + row * c + v * 8, gn);
    else d_unpack8<T>(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(gain) + row * c + v * 8)), gn);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] *= gn[i];
  }
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = Cvt<T>::pack2(f[2 * k] * scale, f[2 * k + 1] * scale);
  *reinterpret_cast<uint4*>(out + row * c + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------------
// Stem patch matrix for the contract-mode plans with uint8 input (engine/base.py _stem_fwd_im2col): the k x k / stride conv over the
// 6-channel normalised [x, 1-x] input is LINEAR in the raw byte v of every colour: xn_c = v/255/std_c - mean_c/std_c and
// xn_{c+3} = -v/255/std_{c+3} + (1 - mean_{c+3})/std_{c+3}.  So the conv equals a GEMM over rows of (v_R, v_G, v_B, 1) per tap with
// folded weights - the bytes are EXACT in one 16-bit plane (only the weights need precision planes: 2 plane products instead of 3),
// the "1" column carries the constant term exactly where the tap lies inside the image (zero padding applies to xn, not to v), and
// the 49-tap window is gathered once here instead of 16 x 3 times by the implicit GEMM (the stem was L2-bandwidth bound).
//   out  [pixels][kp]: column tap*4 + {0,1,2} = v * a_scale, tap*4 + 3 = 1 for in-image taps, zeros elsewhere (kp % 64 == 0)
//   inv_norm [pixels] = 1 / sqrt(sum over in-image taps and the 6 channels of xn^2 + 1e-6)      (calc_patch_norms bcosconv2d.py:196-231)
// One warp per output pixel, lane = 16-byte vector (two taps).
// ------------------------------------------------------------------------------------------------
struct SF6 { float v[6]; };

// One CTA per (image, output row): the k input rows of the three colour planes are staged in shared memory (bytes) together with
// the per-pixel sum of the six normalised channels squared; every thread then assembles 16-byte output vectors (two taps) from
// shared memory - the 512-byte rows of the patch matrix leave fully coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
stem_im2col_u8_kernel(const uint8_t* __restrict__ x, int H, int W, int k, int stride, int pad, int op, int oq, SF6 mean, SF6 istd,
                      float a_scale, T* __restrict__ out, int kp, float* __restrict__ inv_norm) {
  extern __shared__ uint8_t sm_raw[];
  const int img = blockIdx.y, p = blockIdx.x;
  uint8_t* sb = sm_raw;                                              // [k][3][W] bytes
  float* sq = reinterpret_cast<float*>(sm_raw + ((k * 3 * W + 15) / 16) * 16);   // [k][W] sum_c xn^2 (0 for rows outside the image)
  const uint8_t* xi = x + (size_t)img * 3 * H * W;
  const int iy0 = p * stride - pad;
  for (int e = threadIdx.x; e < k * 3 * W; e += blockDim.x) {
    const int dy = e / (3 * W), r = e - dy * 3 * W, c = r / W, ix = r - c * W;
    const int iy = iy0 + dy;
    sb[e] = (iy >= 0 && iy < H) ? xi[((size_t)c * H + iy) * W + ix] : (uint8_t)0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < k * W; e += blockDim.x) {
    const int dy = e / W, ix = e - dy * W;
    const int iy = iy0 + dy;
    float acc = 0.f;
    if (iy >= 0 && iy < H) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float xv = (float)sb[(dy * 3 + c) * W + ix] * (1.0f / 255.0f);
        const float a = (xv - mean.v[c]) * istd.v[c], b = ((1.0f - xv) - mean.v[c + 3]) * istd.v[c + 3];
        acc = fmaf(a, a, acc);
        acc = fmaf(b, b, acc);
      }
    }
    sq[e] = acc;
  }
  __syncthreads();
  const int nvec = kp >> 3, ntap = k * k;
  const long long row0 = ((long long)img * op + p) * oq;
  {
    // blockDim % nvec == 0: every thread owns ONE vector position (two taps) for all the pixels it writes - the tap arithmetic
    // (runtime divisions) is done once per thread, not once per vector
    const int v0 = threadIdx.x % nvec, qstep = blockDim.x / nvec;
    int off[2][3], dxs[2];
    bool live[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = v0 * 2 + h;
      const int dy = t / k, dx = t - dy * k;
      const int iy = iy0 + dy;
      live[h] = t < ntap && iy >= 0 && iy < H;
      dxs[h] = dx - pad;
#pragma unroll
      for (int c = 0; c < 3; ++c) off[h][c] = (dy * 3 + c) * W;
    }
    for (int q = threadIdx.x / nvec; q < oq; q += qstep) {
      float f[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ix = q * stride + dxs[h];
        float r = 0.f, g = 0.f, b = 0.f, one = 0.f;
        if (live[h] && ix >= 0 && ix < W) {
          r = (float)sb[off[h][0] + ix]; g = (float)sb[off[h][1] + ix]; b = (float)sb[off[h][2] + ix];
          one = 1.f;
        }
        f[h * 4 + 0] = r * a_scale; f[h * 4 + 1] = g * a_scale; f[h * 4 + 2] = b * a_scale; f[h * 4 + 3] = one;
      }
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = Cvt<T>::pack2(f[2 * i], f[2 * i + 1]);
      *reinterpret_cast<uint4*>(out + (row0 + q) * (long long)kp + v0 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  if (inv_norm != nullptr) {
    for (int q = threadIdx.x; q < oq; q += blockDim.x) {
      float acc = 0.f;
      for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) {
          const int ix = q * stride - pad + dx;
          if (ix >= 0 && ix < W) acc += sq[dy * W + ix];
        }
      inv_norm[row0 + q] = 1.0f / sqrtf(acc + 1e-6f);
    }
  }
}

// fp16 patch matrix with a_scale = 2^-6 (what the plans use): no conversion instructions at all.  The k input rows are staged as one
// word per pixel (R | G << 8 | B << 16 | 64 << 24; zero outside the image, zero columns for the padding), PRMT interleaves two of its
// bytes with 0x64 into a half2 whose lanes are 1024 + byte, and one HFMA2 (x 2^-6, - 16) turns that into byte / 64 exactly - the
// indicator byte 64 becomes 1.0.  (The generic kernel below spends ~220 instructions per 16-byte vector, 85 % issue-active, on
// I2F / F2F conversions: profiles/r02_parity_mode.md.)
__global__ void __launch_bounds__(256)
stem_im2col_u8_h64_kernel(const uint8_t* __restrict__ x, int H, int W, int k, int stride, int pad, int op, int oq, SF6 mean, SF6 istd,
                          __half* __restrict__ out, int kp, float* __restrict__ inv_norm) {
  extern __shared__ uint8_t sm_raw[];
  const int img = blockIdx.y, p = blockIdx.x;
  const int WP = W + 2 * pad;
  uint32_t* sw = reinterpret_cast<uint32_t*>(sm_raw);                 // [k][WP] packed pixels
  float* sq = reinterpret_cast<float*>(sw + k * WP);                  // [k][W] sum_c xn^2 (0 for rows outside the image)
  const uint8_t* xi = x + (size_t)img * 3 * H * W;
  const int iy0 = p * stride - pad;
  for (int e = threadIdx.x; e < k * WP; e += blockDim.x) {
    const int dy = e / WP, xx = e - dy * WP;
    const int iy = iy0 + dy, ix = xx - pad;
    uint32_t wv = 0u;
    float acc = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const uint32_t r = xi[((size_t)0 * H + iy) * W + ix], g = xi[((size_t)1 * H + iy) * W + ix], b = xi[((size_t)2 * H + iy) * W + ix];
      wv = r | (g << 8) | (b << 16) | (64u << 24);
      const uint32_t cv[3] = {r, g, b};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float xv = (float)cv[c] * (1.0f / 255.0f);
        const float a = (xv - mean.v[c]) * istd.v[c], bb = ((1.0f - xv) - mean.v[c + 3]) * istd.v[c + 3];
        acc = fmaf(a, a, acc);
        acc = fmaf(bb, bb, acc);
      }
    }
    sw[e] = wv;
    if (ix >= 0 && ix < W) sq[dy * W + ix] = acc;
  }
  __syncthreads();
  const int nvec = kp >> 3, ntap = k * k;
  const long long row0 = ((long long)img * op + p) * oq;
  {
    const int v0 = threadIdx.x % nvec, qstep = blockDim.x / nvec;
    int base[2];
    bool live[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = v0 * 2 + h;
      const int dy = t / k, dx = t - dy * k;
      live[h] = t < ntap;
      base[h] = live[h] ? dy * WP + dx : 0;
    }
    const __half2 sc = __floats2half2_rn(0.015625f, 0.015625f), off = __floats2half2_rn(-16.f, -16.f);
    for (int q = threadIdx.x / nvec; q < oq; q += qstep) {
      uint32_t w[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t px = live[h] ? sw[base[h] + q * stride] : 0u;
        uint32_t lo = __byte_perm(px, 0x64646464u, 0x4140), hi = __byte_perm(px, 0x64646464u, 0x4342);
        __half2 a = __hfma2(*reinterpret_cast<__half2*>(&lo), sc, off), b = __hfma2(*reinterpret_cast<__half2*>(&hi), sc, off);
        w[2 * h] = *reinterpret_cast<uint32_t*>(&a);
        w[2 * h + 1] = *reinterpret_cast<uint32_t*>(&b);
      }
      *reinterpret_cast<uint4*>(out + (row0 + q) * (long long)kp + v0 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  if (inv_norm != nullptr) {
    for (int q = threadIdx.x; q < oq; q += blockDim.x) {
      float acc = 0.f;
      for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) {
          const int ix = q * stride - pad + dx;
          if (ix >= 0 && ix < W) acc += sq[dy * W + ix];
        }
      inv_norm[row0 + q] = 1.0f / sqrtf(acc + 1e-6f);
    }
  }
}

}  // namespace bcosk

using namespace bcosk;

extern "C" int bcosk_stem_im2col_u8(const uint8_t* x, int32_t nb, int32_t h, int32_t w, int32_t k, int32_t stride, int32_t pad, const float* mean6,
                                    const float* inv_std6, float a_scale, void* out, int32_t kp, float* inv_norm, int32_t dtype, void* stream) {
  if (!x || !out || !mean6 || !inv_std6 || nb < 1 || k < 1 || stride < 1 || kp % 64 || kp < k * k * 4 || 256 % (kp / 8) != 0)
    return set_error(BCOSK_EINVAL, "stem_im2col_u8: bad argument (kp must be 64, 128 or 256 and >= 4 k^2)");
  const int op = (h + 2 * pad - k) / stride + 1, oq = (w + 2 * pad - k) / stride + 1;
  if (nb > 65535) return set_error(BCOSK_EUNSUPPORTED, "stem_im2col_u8: batch too large for the grid");
  SF6 m, s;
  for (int i = 0; i < 6; ++i) { m.v[i] = mean6[i]; s.v[i] = inv_std6[i]; }
  const size_t smem = (size_t)((k * 3 * w + 15) / 16) * 16 + (size_t)k * w * sizeof(float);
  if (smem > 48 * 1024) return set_error(BCOSK_EUNSUPPORTED, "stem_im2col_u8: image too wide for the shared-memory rows");
  dim3 grid(op, nb);
  if (dtype == BCOSK_DTYPE_F16 && a_scale == 0.015625f) {
    const size_t smem_h = (size_t)k * (w + 2 * pad) * 4 + (size_t)k * w * sizeof(float);
    if (smem_h > 48 * 1024) return set_error(BCOSK_EUNSUPPORTED, "stem_im2col_u8: image too wide for the shared-memory rows");
    stem_im2col_u8_h64_kernel<<<grid, 256, smem_h, SD(stream)>>>(x, h, w, k, stride, pad, op, oq, m, s, reinterpret_cast<__half*>(out), kp, inv_norm);
    BCOSK_CUDA_CHECK(cudaGetLastError());
    return BCOSK_OK;
  }
  if (dtype == BCOSK_DTYPE_BF16)
    stem_im2col_u8_kernel<__nv_bfloat16><<<grid, 256, smem, SD(stream)>>>(x, h, w, k, stride, pad, op, oq, m, s, a_scale, reinterpret_cast<__nv_bfloat16*>(out), kp, inv_norm);
  else if (dtype == BCOSK_DTYPE_F16)
    stem_im2col_u8_kernel<__half><<<grid, 256, smem, SD(stream)>>>(x, h, w, k, stride, pad, op, oq, m, s, a_scale, reinterpret_cast<__half*>(out), kp, inv_norm);
  else
    return set_error(BCOSK_EINVAL, "stem_im2col_u8: dtype");
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}


extern "C" int bcosk_dense_bn_relu_fwd(const void* x, int64_t rows, int32_t c, int32_t planes, int32_t x_ld, int32_t x_plane_stride,
                                       const float* alpha, int32_t relu, void* y, float* sq, uint32_t* maskbits, int32_t dtype, void* stream) {
  if (!x || !y || !alpha || rows < 1 || c < 8 || c % 8 || planes < 1 || planes > 3 || x_ld % 8 || x_plane_stride % 8)
    return set_error(BCOSK_EINVAL, "dense_bn_relu_fwd: bad argument");
  if (maskbits && c % 32) return set_error(BCOSK_EINVAL, "dense_bn_relu_fwd: mask bits need c % 32 == 0");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (dtype == BCOSK_DTYPE_BF16)
    dense_bn_relu_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, SD(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), rows, c, planes, x_ld,
                                                                        x_plane_stride, alpha, relu, reinterpret_cast<__nv_bfloat16*>(y), sq, maskbits);
  else if (dtype == BCOSK_DTYPE_F16)
    dense_bn_relu_fwd_kernel<__half><<<grid, 256, 0, SD(stream)>>>(reinterpret_cast<const __half*>(x), rows, c, planes, x_ld, x_plane_stride, alpha,
                                                                 relu, reinterpret_cast<__half*>(y), sq, maskbits);
  else
    return set_error(BCOSK_EINVAL, "dense_bn_relu_fwd: dtype");
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_dense_bn_relu_bwd(const void* g, int32_t g_f32, int64_t rows, int32_t c, const float* alpha, const uint32_t* maskbits,
                                       float* G, int32_t g_ld, int32_t accumulate, int32_t dtype, void* stream) {
  if (!g || !G || !alpha || rows < 1 || c < 8 || c % 8 || g_ld % 4 || g_ld < c) return set_error(BCOSK_EINVAL, "dense_bn_relu_bwd: bad argument");
  const long long n = rows * (c / 8);
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == BCOSK_DTYPE_BF16)
    dense_bn_relu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, SD(stream)>>>(g, g_f32, rows, c, alpha, maskbits, G, g_ld, accumulate);
  else if (dtype == BCOSK_DTYPE_F16)
    dense_bn_relu_bwd_kernel<__half><<<grid, 256, 0, SD(stream)>>>(g, g_f32, rows, c, alpha, maskbits, G, g_ld, accumulate);
  else
    return set_error(BCOSK_EINVAL, "dense_bn_relu_bwd: dtype");
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_dense_slice_cast(const float* G, int32_t g_ld, int32_t col0, int64_t rows, int32_t c, const void* gain, int32_t gain_f32,
                                      float scale, void* out, int32_t dtype, void* stream) {
  if (!G || !out || rows < 1 || c < 8 || c % 8 || col0 % 4 || g_ld % 4 || col0 + c > g_ld)
    return set_error(BCOSK_EINVAL, "dense_slice_cast: bad argument");
  const long long n = rows * (c / 8);
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == BCOSK_DTYPE_BF16)
    dense_slice_cast_kernel<__nv_bfloat16><<<grid, 256, 0, SD(stream)>>>(G, g_ld, col0, rows, c, gain, gain_f32, scale, reinterpret_cast<__nv_bfloat16*>(out));
  else if (dtype == BCOSK_DTYPE_F16)
    dense_slice_cast_kernel<__half><<<grid, 256, 0, SD(stream)>>>(G, g_ld, col0, rows, c, gain, gain_f32, scale, reinterpret_cast<__half*>(out));
  else
    return set_error(BCOSK_EINVAL, "dense_slice_cast: dtype");
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_copy_rows_2d(void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes, int64_t width_bytes,
                                  int64_t rows, void* stream) {
  if (!dst || !src || rows < 1 || width_bytes < 1) return set_error(BCOSK_EINVAL, "copy_rows_2d: bad argument");
  BCOSK_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes, (size_t)rows,
                                     cudaMemcpyDeviceToDevice, SD(stream)));
  return BCOSK_OK;
}
