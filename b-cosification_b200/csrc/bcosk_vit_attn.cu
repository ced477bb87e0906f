// bcosk_vit_attn.cu -- tensor-core attention of the fused SimpleViT plan for sm_100a (tcgen05 / TMEM / TMA).
//
// Reference: Attention.forward bcos/models/vit.py:143-158 (softmax(q k^T / sqrt(d)) v per head; in explanation mode q and k
// are detached, so the backward is P^T g only).  Operands are the plan's plane rows: qkv [images * n][planes * 3 * heads * 64]
// (q | k | v blocks of every precision plane), value = sum of planes.
//
// Forward, one CTA per (image, head, 128-query tile):
//   TMA          Q tile [128][64] and K, V [nk][64] of every plane as SWIZZLE_128B boxes (nk = n rounded up to 16; rows past the
//                image belong to the next image or are zero-filled: their probabilities are forced to 0)
//   tcgen05.mma  S[128][nk] = sum over plane pairs (q0 k0 + q0 k1 + q1 k0) of Q K^T, fp32 in TMEM (K = 64: four K = 16 steps)
//   4 warps      thread = query row (TMEM lane): row maximum, e = exp2((s - max) * scale * log2 e), row sum; e is written as
//                16-bit planes into shared memory in the K-major SWIZZLE_128B layout of an A operand (over the dead Q / K boxes)
//   tcgen05.mma  O[128][64] = sum over plane pairs of E V, V fed as an MN-major B operand straight from its [token][64] box
//   epilogue     O / row sum -> planes -> 128-byte row segments of out
//
// Backward (explanation mode), one CTA per (image, head): S and E for ALL queries (two 128-row tiles in the 512 TMEM columns),
// E stored [query][key] and read back as an MN-major A operand (E^T), g / row sum converted to one 16-bit plane [query][64]
// (MN-major B):  gv[key][64] = sum over queries E[query][key] * g[query] / sum[query].
#include <cuda.h>
#include <cfloat>
#include <cstring>

#include "../../include/bcosk.h"
#include "bcosk_common.cuh"
#include "bcosk_host.h"

namespace bcosk {
namespace vta {
constexpr int BM = 128;
constexpr int DH = 64;
constexpr int BOX = BM * 128;          // [128 rows][128 bytes]
constexpr int THREADS_FWD = 192;       // TMA warp, MMA warp, 4 softmax / epilogue warps
constexpr int THREADS_BWD = 320;       // TMA warp, MMA warp, 8 softmax / epilogue warps (two query tiles)
constexpr int TAIL = 256;
}  // namespace vta

__device__ __forceinline__ uint64_t vta_desc_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  // MN-major operand, SWIZZLE_128B: a 128-byte row holds 64 consecutive MN elements of one K index, 8 K rows form a swizzle
  // atom (SBO = 1024 bytes), the next group of 64 MN elements starts lbo_bytes further (same form as bcosk_wgrad.cu)
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= 2ull << 61;
  return d;
}

__constant__ int c_seg_a[3][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}, {0, 0, 1, 0, 2, 1}};
__constant__ int c_seg_b[3][6] = {{0, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 1, 0, 2, 0, 1}};
__device__ __forceinline__ int num_segs(int planes) { return planes == 1 ? 1 : (planes == 2 ? 3 : 6); }

// 8 values -> `planes` 16-bit planes, one 16-byte unit per plane at addr + pl * plane_stride (shared memory)
template <typename T>
__device__ __forceinline__ void sts_unit_planes(uint32_t addr, uint32_t plane_stride, int planes, const float (&v)[8]) {
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
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + pl * plane_stride), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
}

template <typename T>
__global__ void __launch_bounds__(vta::THREADS_FWD, 1)
vit_attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, T* __restrict__ out,
                       int planes, int n, int nk, int heads, int mtiles, float scale_log2e, uint32_t fmt, uint32_t region_bytes) {
  using namespace vta;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x % mtiles;
  const int bh = blockIdx.x / mtiles;
  const int b = bh / heads, h = bh - b * heads;
  const int hd = heads * DH, pst = 3 * hd;
  const uint32_t kvb = (uint32_t)nk * 128u;             // bytes of one K / V plane box
  const int kchunks = (nk + 63) >> 6;
  uint8_t* tail = smem + region_bytes + planes * kvb;
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_p = bar_qk + 3;
  uint64_t* bar_o = bar_qk + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_qk + 8);

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      const int row0 = b * n;
      mbar_arrive_expect_tx(bar_qk, (uint32_t)planes * (BOX + kvb));
      for (int pl = 0; pl < planes; ++pl) tma_load_2d(smem + pl * BOX, &tmap_q, bar_qk, pl * pst + h * DH, row0 + mt * BM);
      for (int pl = 0; pl < planes; ++pl) tma_load_2d(smem + planes * BOX + pl * kvb, &tmap_kv, bar_qk, pl * pst + hd + h * DH, row0);
      mbar_arrive_expect_tx(bar_v, (uint32_t)planes * kvb);
      for (int pl = 0; pl < planes; ++pl) tma_load_2d(smem + region_bytes + pl * kvb, &tmap_kv, bar_v, pl * pst + 2 * hd + h * DH, row0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int nseg = num_segs(planes);
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_f16(fmt, BM, (uint32_t)nk);
      uint32_t acc = 0;
      for (int s = 0; s < nseg; ++s) {
        const uint64_t da = umma_smem_desc_kmajor(smem_u32(smem + c_seg_a[planes - 1][s] * BOX), 128);
        const uint64_t db = umma_smem_desc_kmajor(smem_u32(smem + planes * BOX + c_seg_b[planes - 1][s] * kvb), 128);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          umma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_s, acc);
          acc = 1;
        }
      }
      umma_commit(bar_s);
      mbar_wait(bar_p, 0);
      mbar_wait(bar_v, 0);
      tc_fence_after();
      // O = E V: A = E (K-major, 64-key chunks), B = V as MN-major [token rows][64 columns]
      const uint32_t idesc_o = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((uint32_t)(DH >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      acc = 0;
      for (int s = 0; s < nseg; ++s) {
        const uint32_t e_base = smem_u32(smem + c_seg_a[planes - 1][s] * kchunks * BOX);
        const uint32_t v_base = smem_u32(smem + region_bytes + c_seg_b[planes - 1][s] * kvb);
        for (int k = 0; k < nk / 16; ++k) {
          const uint64_t da = umma_smem_desc_kmajor(e_base + (k >> 2) * BOX, 128) + (uint64_t)(2 * (k & 3));
          const uint64_t db = vta_desc_mnmajor(v_base + k * 2048, kvb);
          umma_f16(tmem_base, da, db, idesc_o, acc);
          acc = 1;
        }
      }
      umma_commit(bar_o);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int i = mt * BM + row;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int nch = (nk + 31) >> 5;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -FLT_MAX;
    for (int c = 0; c < nch; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32(t_row + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        if (c * 32 + jj < n) mx = fmaxf(mx, __uint_as_float(raw[jj]));
    }
    float sum = 0.f;
    const uint32_t e_plane = (uint32_t)kchunks * BOX;
    const uint32_t e_row = smem_u32(smem) + row * 128;
    for (int c = 0; c < nch; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32(t_row + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float e[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int j = c * 32 + u * 8 + t;
          e[t] = j < n ? exp2f((__uint_as_float(raw[u * 8 + t]) - mx) * scale_log2e) : 0.f;
          sum += e[t];
        }
        const int j0 = c * 32 + u * 8;
        const uint32_t addr = e_row + (j0 >> 6) * BOX + ((((j0 & 63) >> 3) ^ (row & 7)) << 4);
        sts_unit_planes<T>(addr, e_plane, planes, e);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_p);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    T* orow = out + ((size_t)b * n + i) * ((size_t)planes * hd) + h * DH;
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32(t_row + c * 32, raw);
      tmem_ld_wait();
      if (i < n) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float r[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) r[t] = __uint_as_float(raw[u * 8 + t]) * inv;
          for (int pl = 0; pl < planes; ++pl) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              w[k] = Cvt<T>::pack2(r[2 * k], r[2 * k + 1]);
              const float2 q = Cvt<T>::unpack2(w[k]);
              r[2 * k] -= q.x; r[2 * k + 1] -= q.y;
            }
            *reinterpret_cast<uint4*>(orow + (size_t)pl * hd + c * 32 + u * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------ backward (explanation mode)
template <typename T>
__global__ void __launch_bounds__(vta::THREADS_BWD, 1)
vit_attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, const float* __restrict__ g,
                       T* __restrict__ gv, int planes, int n, int nk, int heads, float scale_log2e, uint32_t fmt, uint32_t region_bytes) {
  using namespace vta;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
  const int hd = heads * DH, pst = 3 * hd;
  const uint32_t kvb = (uint32_t)nk * 128u;
  const int mtiles = (n + BM - 1) / BM;                  // 1 or 2 query tiles
  const uint32_t qrows = (uint32_t)mtiles * BM;          // query rows held in shared memory
  const uint32_t e_chunk = qrows * 128u;                 // E: [64-key chunk][query row][128 bytes]
  const uint32_t g_off = region_bytes;                   // g: [query row][128 bytes], one 16-bit plane
  uint8_t* tail = smem + region_bytes + qrows * 128u;
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_s = bar_qk + 1;
  uint64_t* bar_p = bar_qk + 2;
  uint64_t* bar_o = bar_qk + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_qk + 8);

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_o, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      const int row0 = b * n;
      mbar_arrive_expect_tx(bar_qk, (uint32_t)planes * (mtiles * BOX + kvb));
      for (int t = 0; t < mtiles; ++t)
        for (int pl = 0; pl < planes; ++pl)
          tma_load_2d(smem + (t * planes + pl) * BOX, &tmap_q, bar_qk, pl * pst + h * DH, row0 + t * BM);
      for (int pl = 0; pl < planes; ++pl)
        tma_load_2d(smem + mtiles * planes * BOX + pl * kvb, &tmap_kv, bar_qk, pl * pst + hd + h * DH, row0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int nseg = num_segs(planes);
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_f16(fmt, BM, (uint32_t)nk);
      for (int t = 0; t < mtiles; ++t) {
        uint32_t acc = 0;
        for (int s = 0; s < nseg; ++s) {
          const uint64_t da = umma_smem_desc_kmajor(smem_u32(smem + (t * planes + c_seg_a[planes - 1][s]) * BOX), 128);
          const uint64_t db = umma_smem_desc_kmajor(smem_u32(smem + mtiles * planes * BOX + c_seg_b[planes - 1][s] * kvb), 128);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) {
            umma_f16(tmem_base + t * 256, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_s, acc);
            acc = 1;
          }
        }
      }
      umma_commit(bar_s);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      // gv tile jt: D[128 keys][64] = sum over queries E^T g;  A = E^T (MN-major: rows = queries, 64 keys per 128-byte row,
      // the next 64 keys e_chunk bytes further), B = g (MN-major, 64 columns)
      const uint32_t idesc_o = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(DH >> 3) << 17) |
                               ((uint32_t)(BM >> 4) << 24);
      const int ksteps = (n + 15) >> 4;                  // query rows in steps of 16 (rows >= n hold zeros)
      const int jtiles = (nk + BM - 1) / BM;
      for (int jt = 0; jt < jtiles; ++jt) {
        uint32_t acc = 0;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t da = vta_desc_mnmajor(smem_u32(smem) + (2 * jt) * e_chunk + k * 2048, e_chunk);
          const uint64_t db = vta_desc_mnmajor(smem_u32(smem) + g_off + k * 2048, qrows * 128u);
          umma_f16(tmem_base + jt * 256, da, db, idesc_o, acc);
          acc = 1;
        }
      }
      umma_commit(bar_o);
    }
  } else {
    const int ew = warp - 2;                             // 0..7
    const int tile = ew >> 2;                            // query tile handled in the softmax phase / key tile in the epilogue
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + tile * 256;
    const int nch = (nk + 31) >> 5;
    const int i = tile * BM + row;                       // query index
    mbar_wait(bar_s, 0);
    tc_fence_after();
    if (tile < mtiles) {
      float mx = -FLT_MAX;
      for (int c = 0; c < nch; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(t_row + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          if (c * 32 + jj < n) mx = fmaxf(mx, __uint_as_float(raw[jj]));
      }
      float sum = 0.f;
      const uint32_t e_row = smem_u32(smem) + (uint32_t)i * 128u;
      const bool live = i < n;
      for (int c = 0; c < nch; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(t_row + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float e[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int j = c * 32 + u * 8 + t;
            e[t] = (live && j < n) ? exp2f((__uint_as_float(raw[u * 8 + t]) - mx) * scale_log2e) : 0.f;
            sum += e[t];
          }
          const int j0 = c * 32 + u * 8;
          const uint32_t addr = e_row + (j0 >> 6) * e_chunk + ((((j0 & 63) >> 3) ^ (i & 7)) << 4);
          sts_unit_planes<T>(addr, 0, 1, e);
        }
      }
      // g row / row sum -> one 16-bit plane, MN-major B rows
      const float inv = live ? 1.0f / sum : 0.f;
      const float* grow = g + ((size_t)b * n + (live ? i : 0)) * hd + h * DH;
      const uint32_t g_row = smem_u32(smem) + g_off + (uint32_t)i * 128u;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float v[8];
        const float4 a = __ldg(reinterpret_cast<const float4*>(grow + u * 8)), c4 = __ldg(reinterpret_cast<const float4*>(grow + u * 8) + 1);
        v[0] = a.x * inv; v[1] = a.y * inv; v[2] = a.z * inv; v[3] = a.w * inv;
        v[4] = c4.x * inv; v[5] = c4.y * inv; v[6] = c4.z * inv; v[7] = c4.w * inv;
        sts_unit_planes<T>(g_row + ((u ^ (i & 7)) << 4), 0, 1, v);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_p);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const int j = tile * BM + row;                       // key index
    const int jtiles = (nk + BM - 1) / BM;
    if (tile < jtiles) {
      T* orow = gv + ((size_t)b * n + (j < n ? j : 0)) * hd + h * DH;
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(t_row + c * 32, raw);
        tmem_ld_wait();
        if (j < n) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w[k] = Cvt<T>::pack2(__uint_as_float(raw[u * 8 + 2 * k]), __uint_as_float(raw[u * 8 + 2 * k + 1]));
            *reinterpret_cast<uint4*>(orow + c * 32 + u * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace bcosk

using namespace bcosk;

extern "C" int bcosk_vit_attention_tc(const void* qkv, int32_t planes, const float* g, int32_t batch, int32_t n, int32_t heads,
                                      int32_t dim_head, float scale, int32_t backward, void* out, int32_t dtype, void* stream) {
  using namespace vta;
  if (!qkv || !out || (backward && !g) || planes < 1 || planes > 3 || batch < 1 || heads < 1)
    return set_error(BCOSK_EINVAL, "vit_attention_tc: bad argument");
  if (dim_head != DH) return set_error(BCOSK_EUNSUPPORTED, "vit_attention_tc: dim_head must be 64");
  if (n < 1 || n > 256) return set_error(BCOSK_EUNSUPPORTED, "vit_attention_tc: 1 <= n <= 256 tokens");
  if (dtype != BCOSK_DTYPE_BF16 && dtype != BCOSK_DTYPE_F16) return set_error(BCOSK_EINVAL, "vit_attention_tc: dtype");
  const int nk = (n + 15) / 16 * 16;
  const int kchunks = (nk + 63) / 64;
  const int mtiles = (n + BM - 1) / BM;
  const long long rows = (long long)batch * n;
  const long long ld = (long long)planes * 3 * heads * DH;
  CUtensorMap mq, mkv;
  int rc = make_tiled_map_2d(&mq, qkv, ld, rows, DH, BM, 128);
  if (rc) return rc;
  rc = make_tiled_map_2d(&mkv, qkv, ld, rows, DH, nk, 128);
  if (rc) return rc;
  const float scale_log2e = scale * 1.4426950408889634f;
  const uint32_t fmt = (uint32_t)dtype;                  // 0 = fp16, 1 = bf16 (the UMMA format codes)
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t kvb = (uint32_t)nk * 128u;
  if (!backward) {
    const uint32_t qk_b = (uint32_t)planes * (BOX + kvb), e_b = (uint32_t)planes * kchunks * BOX;
    const uint32_t region = qk_b > e_b ? qk_b : e_b;
    const size_t smem = (size_t)region + (size_t)planes * kvb + TAIL;
    if (smem > 227 * 1024) return set_error(BCOSK_EUNSUPPORTED, "vit_attention_tc: shared memory");
    const unsigned grid = (unsigned)((long long)batch * heads * mtiles);
    if (dtype == BCOSK_DTYPE_F16) {
      BCOSK_CUDA_CHECK(cudaFuncSetAttribute(vit_attn_tc_fwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      vit_attn_tc_fwd_kernel<__half><<<grid, THREADS_FWD, smem, st>>>(mq, mkv, reinterpret_cast<__half*>(out), planes, n, nk, heads, mtiles,
                                                                      scale_log2e, fmt, region);
    } else {
      BCOSK_CUDA_CHECK(cudaFuncSetAttribute(vit_attn_tc_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      vit_attn_tc_fwd_kernel<__nv_bfloat16><<<grid, THREADS_FWD, smem, st>>>(mq, mkv, reinterpret_cast<__nv_bfloat16*>(out), planes, n, nk,
                                                                             heads, mtiles, scale_log2e, fmt, region);
    }
  } else {
    const uint32_t qrows = (uint32_t)mtiles * BM;
    const uint32_t qk_b = (uint32_t)planes * (mtiles * BOX + kvb);
    const uint32_t e_b = (uint32_t)(2 * ((nk + BM - 1) / BM)) * qrows * 128u;      // whole 128-key tiles: 2 chunks each
    const uint32_t region = qk_b > e_b ? qk_b : e_b;
    const size_t smem = (size_t)region + (size_t)qrows * 128u + TAIL;
    if (smem > 227 * 1024) return set_error(BCOSK_EUNSUPPORTED, "vit_attention_tc: shared memory (backward)");
    const unsigned grid = (unsigned)((long long)batch * heads);
    if (dtype == BCOSK_DTYPE_F16) {
      BCOSK_CUDA_CHECK(cudaFuncSetAttribute(vit_attn_tc_bwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      vit_attn_tc_bwd_kernel<__half><<<grid, THREADS_BWD, smem, st>>>(mq, mkv, g, reinterpret_cast<__half*>(out), planes, n, nk, heads,
                                                                      scale_log2e, fmt, region);
    } else {
      BCOSK_CUDA_CHECK(cudaFuncSetAttribute(vit_attn_tc_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      vit_attn_tc_bwd_kernel<__nv_bfloat16><<<grid, THREADS_BWD, smem, st>>>(mq, mkv, g, reinterpret_cast<__nv_bfloat16*>(out), planes, n, nk,
                                                                             heads, scale_log2e, fmt, region);
    }
  }
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}
