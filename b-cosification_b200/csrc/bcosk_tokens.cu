// bcosk_tokens.cu -- token-model kernels of the B-cosified ViT path (fp32 I/O, module-level path):
// detachable LayerNorm, detached-gate GELU, softmax attention with frozen (detached) probabilities.
// Reference: bcos/modules/norms/centered_norms.py:187-224, bcosify_vit.py:27-32, bcos/models/vit.py:143-158.
#include "../../include/bcosk.h"
#include "bcosk_common.cuh"
#include "bcosk_host.h"

namespace bcosk {

// ---------------------------------------------------------------- LayerNorm: one warp per row
// fwd: y = w * (x - mean) / sqrt(var + eps) + b ; rstd[row] saved for the explanation backward
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, long long rows, int d, const float* __restrict__ w,
                                     const float* __restrict__ b, float eps, float* __restrict__ y, float* __restrict__ rstd) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* src = x + row * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += __ldg(src + i);
  const float mean = warp_sum(s) / (float)d;
  float v = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float c = __ldg(src + i) - mean;
    v = fmaf(c, c, v);
  }
  const float var = warp_sum(v) / (float)d;      // biased, like torch.var_mean(unbiased=False)
  const float r = 1.0f / sqrtf(var + eps);
  for (int i = lane; i < d; i += 32) {
    float o = (__ldg(src + i) - mean) * r;
    if (w != nullptr) o *= __ldg(w + i);
    if (b != nullptr) o += __ldg(b + i);
    y[row * d + i] = o;
  }
  if (rstd != nullptr && lane == 0) rstd[row] = r;
}
// explanation backward (variance detached, mean in graph): gx = (w*gy - mean_d(w*gy)) * rstd
__global__ void layernorm_explain_bwd_kernel(const float* __restrict__ gy, long long rows, int d, const float* __restrict__ w,
                                             const float* __restrict__ rstd, float* __restrict__ gx) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* src = gy + row * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += __ldg(src + i) * (w ? __ldg(w + i) : 1.f);
  const float m = warp_sum(s) / (float)d;
  const float r = __ldg(rstd + row);
  for (int i = lane; i < d; i += 32) gx[row * d + i] = (__ldg(src + i) * (w ? __ldg(w + i) : 1.f) - m) * r;
}


// ---------------------------------------------------------------- per-row L2 normalisation (attn_unpool head)
// y = x / ||x||_2 per row of d values, inv[row] = 1 / ||x||_2 saved for the explanation backward; one warp per row
__global__ void l2norm_rows_kernel(const float* __restrict__ x, long long rows, int d, float* __restrict__ y,
                                   float* __restrict__ inv) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* src = x + row * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float v = __ldg(src + i);
    s = fmaf(v, v, s);
  }
  const float r = 1.0f / sqrtf(warp_sum(s));
  for (int i = lane; i < d; i += 32) y[row * d + i] = __ldg(src + i) * r;
  if (inv != nullptr && lane == 0) inv[row] = r;
}
// y = x * s[row]   (explanation backward of the above with the norm detached)
__global__ void row_scale_kernel(const float* __restrict__ x, long long rows, int d, const float* __restrict__ s,
                                 float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  y[i] = __ldg(x + i) * __ldg(s + i / d);
}

// ---------------------------------------------------------------- GELU with detachable gate
// mode 0: y = x * gate(x); mode 1 (explanation backward): y = g * gate(x)
__global__ void gelu_gate_kernel(const float* __restrict__ x, const float* __restrict__ g, long long n, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = __ldg(x + i);
  const float gate = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
  y[i] = (g ? __ldg(g + i) : v) * gate;
}

// ---------------------------------------------------------------- attention with frozen probabilities
// One CTA per (batch, head).  qkv: [B, N, 3*H*D] fp32 (q | k | v blocks, head h at h*D), D = 64, N <= 208.
//   BWD = false: out[B, N, H*D]   = softmax(q k^T * scale) v
//   BWD = true : out[B, N, 3*H*D] : v-block = P^T g  (q,k blocks untouched = zero); g: [B, N, H*D]
// P (N x N fp32) lives in shared memory; K / V / g are staged in a second buffer.
constexpr int ATT_D = 64;
template <bool BWD>
__global__ void __launch_bounds__(256, 1)
attention_kernel(const float* __restrict__ qkv, const float* __restrict__ g, int n, int heads, float scale,
                 float* __restrict__ out) {
  extern __shared__ float sm[];
  float* P = sm;                       // [n][n]
  float* buf = sm + (size_t)n * n;     // [n][ATT_D + 1]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * heads * ATT_D;
  const float* q = qkv + (size_t)b * n * ld + h * ATT_D;
  const float* k = q + heads * ATT_D;
  const float* v = k + heads * ATT_D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  constexpr int LDB = ATT_D + 1;
  // phase 1: K -> buf
  for (int i = threadIdx.x; i < n * ATT_D; i += blockDim.x) buf[(i / ATT_D) * LDB + (i % ATT_D)] = __ldg(k + (size_t)(i / ATT_D) * ld + (i % ATT_D));
  __syncthreads();
  // phase 2: P rows (one warp per query row)
  for (int i = warp; i < n; i += nw) {
    float qi[2] = {__ldg(q + (size_t)i * ld + lane), __ldg(q + (size_t)i * ld + lane + 32)};
    float mx = -INFINITY;
    for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      float s = 0.f;
      // each lane computes the full 64-dim dot product for its key j (q broadcast through shuffles)
#pragma unroll
      for (int dd = 0; dd < 32; ++dd) {
        const float q0 = __shfl_sync(0xffffffffu, qi[0], dd), q1 = __shfl_sync(0xffffffffu, qi[1], dd);
        if (j < n) s = fmaf(q0, buf[j * LDB + dd], fmaf(q1, buf[j * LDB + dd + 32], s));
      }
      s *= scale;
      if (j < n) { P[(size_t)i * n + j] = s; mx = fmaxf(mx, s); }
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float e = expf(P[(size_t)i * n + j] - mx);
      P[(size_t)i * n + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < n; j += 32) P[(size_t)i * n + j] *= inv;
  }
  __syncthreads();
  // phase 3: V (forward) or g (backward) -> buf
  if (!BWD) {
    for (int i = threadIdx.x; i < n * ATT_D; i += blockDim.x) buf[(i / ATT_D) * LDB + (i % ATT_D)] = __ldg(v + (size_t)(i / ATT_D) * ld + (i % ATT_D));
  } else {
    const float* gh = g + (size_t)b * n * heads * ATT_D + h * ATT_D;
    for (int i = threadIdx.x; i < n * ATT_D; i += blockDim.x)
      buf[(i / ATT_D) * LDB + (i % ATT_D)] = __ldg(gh + (size_t)(i / ATT_D) * heads * ATT_D + (i % ATT_D));
  }
  __syncthreads();
  // phase 4: out[i, :] = sum_j P[i, j] * buf[j, :]   (forward)   /   out[j, :] = sum_i P[i, j] * buf[i, :]   (backward)
  for (int r = warp; r < n; r += nw) {
    float a0 = 0.f, a1 = 0.f;
    for (int t = 0; t < n; ++t) {
      const float pw = BWD ? P[(size_t)t * n + r] : P[(size_t)r * n + t];
      a0 = fmaf(pw, buf[t * LDB + lane], a0);
      a1 = fmaf(pw, buf[t * LDB + lane + 32], a1);
    }
    if (!BWD) {
      float* o = out + ((size_t)b * n + r) * heads * ATT_D + h * ATT_D;
      o[lane] = a0; o[lane + 32] = a1;
    } else {
      float* o = out + ((size_t)b * n + r) * ld + 2 * heads * ATT_D + h * ATT_D;
      o[lane] = a0; o[lane + 32] = a1;
    }
  }
}

}  // namespace bcosk

using namespace bcosk;
static inline cudaStream_t S3(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" int bcosk_layernorm_fwd(const float* x, int64_t rows, int32_t d, const float* w, const float* b, float eps, float* y,
                                   float* rstd, void* stream) {
  if (!x || !y || d < 1) return set_error(BCOSK_EINVAL, "layernorm_fwd: bad argument");
  layernorm_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, S3(stream)>>>(x, rows, d, w, b, eps, y, rstd);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_layernorm_explain_bwd(const float* gy, int64_t rows, int32_t d, const float* w, const float* rstd, float* gx,
                                           void* stream) {
  if (!gy || !gx || !rstd) return set_error(BCOSK_EINVAL, "layernorm_explain_bwd: bad argument");
  layernorm_explain_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, S3(stream)>>>(gy, rows, d, w, rstd, gx);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_l2norm_rows(const float* x, int64_t rows, int32_t d, float* y, float* inv, void* stream) {
  if (!x || !y || d < 1 || rows < 0) return set_error(BCOSK_EINVAL, "l2norm_rows: bad argument");
  if (rows == 0) return BCOSK_OK;
  l2norm_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, S3(stream)>>>(x, rows, d, y, inv);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_row_scale(const float* x, int64_t rows, int32_t d, const float* s, float* y, void* stream) {
  if (!x || !y || !s || d < 1 || rows < 0) return set_error(BCOSK_EINVAL, "row_scale: bad argument");
  if (rows == 0) return BCOSK_OK;
  row_scale_kernel<<<(unsigned)((rows * d + 255) / 256), 256, 0, S3(stream)>>>(x, rows, d, s, y);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_gelu_gate(const float* x, const float* g, int64_t n, float* y, void* stream) {
  if (!x || !y) return set_error(BCOSK_EINVAL, "gelu_gate: bad argument");
  gelu_gate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S3(stream)>>>(x, g, n, y);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

extern "C" int bcosk_attention(const float* qkv, const float* g, int32_t batch, int32_t n, int32_t heads, int32_t dim_head,
                               float scale, int32_t backward, float* out, void* stream) {
  if (!qkv || !out || (backward && !g)) return set_error(BCOSK_EINVAL, "attention: null pointer");
  if (dim_head != ATT_D) return set_error(BCOSK_EUNSUPPORTED, "attention: dim_head must be 64");
  const size_t smem = ((size_t)n * n + (size_t)n * (ATT_D + 1)) * sizeof(float);
  if (smem > 227 * 1024) return set_error(BCOSK_EUNSUPPORTED, "attention: sequence too long for the shared-memory kernel (n <= 208)");
  const void* fn = backward ? (const void*)attention_kernel<true> : (const void*)attention_kernel<false>;
  BCOSK_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per (function, device)
  if (backward) attention_kernel<true><<<batch * heads, 256, smem, S3(stream)>>>(qkv, g, n, heads, scale, out);
  else attention_kernel<false><<<batch * heads, 256, smem, S3(stream)>>>(qkv, g, n, heads, scale, out);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}
