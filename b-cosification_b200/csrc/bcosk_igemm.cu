// bcosk_igemm.cu -- implicit-GEMM B-cos convolution / linear, forward and explain-dgrad, for sm_100a.
//
//   D[m, n] = sum_k A[m, k] * B[n, k],  fp32 accumulation in TMEM
//     A: NHWC 16-bit activations gathered by TMA *im2col* loads (one [128 pixel x kch channel] box per
//        filter tap / channel chunk, zero fill for padding and tails) into 128B- (or 64B-) swizzled smem
//     B: packed K-major weights, TMA tiled loads
//     MMA: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16, issued by one elected thread
//   warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue (tcgen05.ld).
//   Non-persistent; two CTAs are co-resident per SM (BN <= 128) so one CTA's epilogue overlaps the other's
//   main loop.  Epilogues are documented in include/bcosk.h (struct bcosk_igemm_params).
#include <cuda.h>

#include "../../include/bcosk.h"
#include "bcosk_common.cuh"
#include "bcosk_host.h"

namespace bcosk {

constexpr int BM = 128;       // CTA tile M (= UMMA M, one TMEM lane per output pixel)
constexpr int STAGE_K = 64;   // K elements per pipeline stage (128 bytes of 16-bit data per row)
constexpr int A_STAGE_BYTES = BM * STAGE_K * 2;
constexpr int NUM_THREADS = 192;

template <int BN, bool HP = false> struct TileCfg {
  static constexpr int kStages = (BN == 128) ? 3 : 4;
  static constexpr int kMinBlocks = (BN <= 128) ? 2 : 1;
  static constexpr int kBStageBytes = BN * STAGE_K * 2;
  // HP (high-precision accumulation): two TMEM accumulators that the epilogue warps drain every pipeline stage
  static constexpr int kTmemCols = (BN < 32 ? 32 : BN) * (HP ? 2 : 1);
  // stages + 1 KB alignment slack + barriers/params
  static constexpr int kSmemBytes = kStages * (A_STAGE_BYTES + kBStageBytes) + 1024 + 256 + 2 * BN * 4;
};

struct RowInfo {
  int m, img, p, q;
  bool valid;
};

// 16-bit row segment (32 columns starting at `ptr`) -> 32 floats, adding over precision planes
template <typename T>
__device__ __forceinline__ void load32_planes(const T* ptr, int planes, int plane_stride, int ncols, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
  for (int pl = 0; pl < planes; ++pl) {
    const uint4* src = reinterpret_cast<const uint4*>(ptr + (size_t)pl * plane_stride);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (g * 8 < ncols) {
        uint4 u = __ldg(src + g);
        float2 f;
        f = Cvt<T>::unpack2(u.x); v[g * 8 + 0] += f.x; v[g * 8 + 1] += f.y;
        f = Cvt<T>::unpack2(u.y); v[g * 8 + 2] += f.x; v[g * 8 + 3] += f.y;
        f = Cvt<T>::unpack2(u.z); v[g * 8 + 4] += f.x; v[g * 8 + 5] += f.y;
        f = Cvt<T>::unpack2(u.w); v[g * 8 + 6] += f.x; v[g * 8 + 7] += f.y;
      }
    }
  }
}

// fp32 row segment (32 columns) -> v
__device__ __forceinline__ void load32_f32(const float* ptr, int ncols, float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    if (g * 4 < ncols) {
      float4 f = __ldg(reinterpret_cast<const float4*>(ptr) + g);
      v[g * 4 + 0] = f.x; v[g * 4 + 1] = f.y; v[g * 4 + 2] = f.z; v[g * 4 + 3] = f.w;
    } else {
      v[g * 4 + 0] = v[g * 4 + 1] = v[g * 4 + 2] = v[g * 4 + 3] = 0.f;
    }
  }
}

// store 32 floats as 16-bit precision planes: plane 0 = rn(v), plane 1 = rn(v - plane0), ...
// returns in `v` the value actually representable by the stored planes (sum of planes).
template <typename T>
__device__ __forceinline__ void store32_planes(T* ptr, int planes, int plane_stride, int ncols, float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float r[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[i] = v[g * 8 + i]; acc[i] = 0.f; }
    for (int pl = 0; pl < planes; ++pl) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        w[j] = Cvt<T>::pack2(r[2 * j], r[2 * j + 1]);
        const float2 f = Cvt<T>::unpack2(w[j]);
        r[2 * j] -= f.x; r[2 * j + 1] -= f.y;
        acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
      }
      if (g * 8 < ncols)
        reinterpret_cast<uint4*>(ptr + (size_t)pl * plane_stride)[g] = make_uint4(w[0], w[1], w[2], w[3]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[g * 8 + i] = acc[i];
  }
}

__device__ __forceinline__ void store32_f32(float* ptr, int ncols, const float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    if (g * 4 < ncols)
      reinterpret_cast<float4*>(ptr)[g] = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
}

// One 32-column slice of one output row: everything after the accumulator is in registers.
template <int MODE, typename T>
__device__ __forceinline__ void epilogue_chunk(const bcosk_igemm_params& p, const RowInfo& ri, int64_t yrow, float inv_norm,
                                               int64_t add_row, const float* s_alpha, const float* s_beta, int j, int c0,
                                               int ncols, float (&v)[32], float& sq_acc) {
  T* y16 = reinterpret_cast<T*>(p.y);
  float* y32 = reinterpret_cast<float*>(p.y);
  if (MODE == BCOSK_MODE_FWD) {
    float t[32];
    if (p.scale_mode == BCOSK_SCALE_B2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = fabsf(v[i]) * inv_norm * s_alpha[j * 32 + i];
    } else if (p.scale_mode == BCOSK_SCALE_POW) {
      const float e = p.b_exp - 1.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = __powf(fabsf(v[i]) * inv_norm + 1e-6f, e) * s_alpha[j * 32 + i];
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = s_alpha[j * 32 + i];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], t[i], s_beta[j * 32 + i]);
    if (p.res != nullptr) {
      float r[32];
      load32_planes<T>(reinterpret_cast<const T*>(p.res) + (size_t)ri.m * p.res_ld + c0, p.res_planes,
                       p.res_plane_stride, ncols, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += r[i];
    }
    uint32_t mbits = 0xffffffffu;
    if (p.relu) {
      mbits = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool pos = v[i] > 0.f;
        mbits |= (pos ? 1u : 0u) << i;
        v[i] = pos ? v[i] : 0.f;
        t[i] = pos ? t[i] : 0.f;
      }
    }
    if (p.maskbits != nullptr) p.maskbits[(size_t)ri.m * p.mask_ld + (c0 >> 5)] = mbits;
    if (p.gain != nullptr) {
      if (p.gain_f32) {
        store32_f32(reinterpret_cast<float*>(p.gain) + (size_t)ri.m * p.gain_ld + c0, ncols, t);
      } else {
        store32_planes<T>(reinterpret_cast<T*>(p.gain) + (size_t)ri.m * p.gain_ld + c0, 1, 0, ncols, t);
      }
    }
    if (p.y_f32) {
      store32_f32(y32 + (size_t)yrow * p.y_ld + c0, ncols, v);
    } else {
      store32_planes<T>(y16 + (size_t)yrow * p.y_ld + c0, p.y_planes, p.y_plane_stride, ncols, v);
    }
    if (p.sq_out != nullptr) {
#pragma unroll
      for (int i = 0; i < 32; ++i) sq_acc = fmaf(v[i], v[i], sq_acc);  // columns >= n are exact zeros
    }
  } else {
    // ---------------- explain-dgrad epilogue ----------------
    if (add_row >= 0) {
      float a[32];
      load32_planes<T>(reinterpret_cast<const T*>(p.add) + (size_t)add_row * p.add_ld + c0, p.add_planes,
                       p.add_plane_stride, ncols, a);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += a[i];
    }
    if (p.out2 != nullptr) {
      float o[32];
      if (p.mul2 != nullptr) {
        if (p.mul2_f32) load32_f32(reinterpret_cast<const float*>(p.mul2) + (size_t)ri.m * p.mul2_ld + c0, ncols, o);
        else load32_planes<T>(reinterpret_cast<const T*>(p.mul2) + (size_t)ri.m * p.mul2_ld + c0, 1, 0, ncols, o);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= v[i];
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = v[i];
      }
      if (p.mask2 != nullptr) {
        const uint32_t mb = __ldg(p.mask2 + (size_t)ri.m * p.mask2_ld + (c0 >> 5));
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = ((mb >> i) & 1u) ? o[i] : 0.f;
      }
      store32_planes<T>(reinterpret_cast<T*>(p.out2) + (size_t)ri.m * p.out2_ld + c0, p.out2_planes,
                        p.out2_plane_stride, ncols, o);
    }
    if (p.mul1 != nullptr) {
      float g[32];
      if (p.mul1_f32) load32_f32(reinterpret_cast<const float*>(p.mul1) + (size_t)ri.m * p.mul1_ld + c0, ncols, g);
      else load32_planes<T>(reinterpret_cast<const T*>(p.mul1) + (size_t)ri.m * p.mul1_ld + c0, 1, 0, ncols, g);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= g[i];
    }
    if (p.y_f32) {
      store32_f32(y32 + (size_t)yrow * p.y_ld + c0, ncols, v);
    } else {
      store32_planes<T>(y16 + (size_t)yrow * p.y_ld + c0, p.y_planes, p.y_plane_stride, ncols, v);
    }
  }
}

template <int BN, int MODE, typename T, bool HP>
__global__ void __launch_bounds__(NUM_THREADS, TileCfg<BN, HP>::kMinBlocks)
bcosk_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ bcosk_igemm_params p) {
  using Cfg = TileCfg<BN, HP>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * Cfg::kBStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;   // non-HP: accumulator complete
  uint64_t* acc_full_bar = tmem_full_bar + 1;      // HP: [2] partial accumulator ready
  uint64_t* acc_empty_bar = acc_full_bar + 2;      // HP: [2] partial accumulator drained
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty_bar + 2);
  float* s_alpha = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);
  float* s_beta = s_alpha + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.n + BN - 1) / BN;
  const int tile_n = blockIdx.x % n_tiles;
  const int tile_m = blockIdx.x / n_tiles;
  const int M = p.a_nb * p.op * p.oq;
  const int m0 = tile_m * BM;
  const int n0 = tile_n * BN;

  const int chunks_per_stage = STAGE_K / p.kch;  // 1 (kch = 64) or 2 (kch = 32)
  const int total_chunks = p.num_segs * p.num_taps * p.chunks_per_tap;
  const int num_iters = total_chunks / chunks_per_stage;  // host guarantees divisibility

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full_bar[s], 1);
      mbar_init(&acc_empty_bar[s], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // first output pixel of this tile -> im2col base coordinates
      const int opq = p.op * p.oq;
      const int img = m0 / opq;
      const int rem = m0 - img * opq;
      const int pp = rem / p.oq;
      const int qq = rem - pp * p.oq;
      const int base_w = p.lo_w + qq * p.stride_w;
      const int base_h = p.lo_h + pp * p.stride_h;
      const uint32_t a_chunk_bytes = BM * p.kch * 2;
      const uint32_t b_chunk_bytes = BN * p.kch * 2;
      const uint32_t stage_bytes = A_STAGE_BYTES + Cfg::kBStageBytes;
      int stage = 0;
      uint32_t phase = 0;
      int seg = 0, tap = 0, kc = 0;
      for (int it = 0; it < num_iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        for (int j = 0; j < chunks_per_stage; ++j) {
          const int ci = it * chunks_per_stage + j;
          tma_load_im2col_4d(smem_a + stage * A_STAGE_BYTES + j * a_chunk_bytes, &tmap_a, &full_bar[stage],
                             p.seg_a_choff[seg] + kc * p.kch, base_w, base_h, img, p.tap_off_w[tap], p.tap_off_h[tap]);
          tma_load_2d(smem_b + stage * Cfg::kBStageBytes + j * b_chunk_bytes, &tmap_b, &full_bar[stage], ci * p.kch, n0);
          if (++kc == p.chunks_per_tap) {
            kc = 0;
            if (++tap == p.num_taps) { tap = 0; ++seg; }
          }
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16((uint32_t)p.dtype, BM, BN);
      const uint32_t row_bytes = p.kch * 2;
      const uint32_t a_chunk_bytes = BM * p.kch * 2;
      const uint32_t b_chunk_bytes = BN * p.kch * 2;
      const int mma_per_chunk = p.kch / 16;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      for (int it = 0; it < num_iters; ++it) {
        uint32_t tmem_d = tmem_base;
        if (HP) {
          // fresh accumulator per stage: the tensor core truncates when it aligns addends to a large running sum,
          // so long dot products are summed in registers (round-to-nearest fp32) by the epilogue warps instead
          const int buf = it & 1;
          mbar_wait(&acc_empty_bar[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
          tmem_d = tmem_base + (uint32_t)(buf * BN);
          accumulate = 0;
        }
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        for (int j = 0; j < chunks_per_stage; ++j) {
          const uint64_t da = umma_smem_desc_kmajor(smem_u32(smem_a + stage * A_STAGE_BYTES + j * a_chunk_bytes), row_bytes);
          const uint64_t db = umma_smem_desc_kmajor(smem_u32(smem_b + stage * Cfg::kBStageBytes + j * b_chunk_bytes), row_bytes);
          for (int k = 0; k < mma_per_chunk; ++k) {
            // advance 16 K-elements = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
            umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, accumulate);
            accumulate = 1;
          }
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (HP) umma_commit(&acc_full_bar[it & 1]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (!HP) umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const int et = (warp - 2) * 32 + lane;  // 0..127
    // stage per-channel vectors while the main loop runs
    if (MODE == BCOSK_MODE_FWD) {
      for (int i = et; i < BN; i += 128) {
        const int c = n0 + i;
        s_alpha[i] = (p.alpha != nullptr && c < p.n) ? __ldg(p.alpha + c) : 1.f;
        s_beta[i] = (p.beta != nullptr && c < p.n) ? __ldg(p.beta + c) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    RowInfo ri;
    ri.m = m0 + row;
    ri.valid = ri.m < M;
    {
      const int opq = p.op * p.oq;
      const int mm = ri.valid ? ri.m : 0;
      ri.img = mm / opq;
      const int rem = mm - ri.img * opq;
      ri.p = rem / p.oq;
      ri.q = rem - ri.p * p.oq;
    }
    const int64_t yrow = p.os_0 + (int64_t)ri.img * p.os_n + (int64_t)ri.p * p.os_p + (int64_t)ri.q * p.os_q;
    float inv_norm = 1.f;
    int64_t add_row = -1;
    if (MODE == BCOSK_MODE_FWD) {
      if (p.scale_mode != BCOSK_SCALE_NONE && ri.valid) inv_norm = __ldg(p.inv_norm + ri.m);
    } else {
      if (p.add != nullptr && ri.valid) {
        const int s = p.add_stride;
        if (ri.p % s == 0 && ri.q % s == 0 && ri.p / s < p.add_p && ri.q / s < p.add_q)
          add_row = ((int64_t)ri.img * p.add_p + ri.p / s) * p.add_q + ri.q / s;
      }
    }
    float sq_acc = 0.f;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;

    if constexpr (HP) {
      float acc[BN];
#pragma unroll
      for (int i = 0; i < BN; ++i) acc[i] = 0.f;
      for (int it = 0; it < num_iters; ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_full_bar[buf], ((uint32_t)it >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + lane_base + (uint32_t)(buf * BN + j * 32), raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[j * 32 + i] += __uint_as_float(raw[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty_bar[buf]);
      }
#pragma unroll
      for (int j = 0; j < BN / 32; ++j) {
        const int c0 = n0 + j * 32;
        if (c0 < p.n && ri.valid) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = acc[j * 32 + i];
          epilogue_chunk<MODE, T>(p, ri, yrow, inv_norm, add_row, s_alpha, s_beta, j, c0, min(32, p.n - c0), v, sq_acc);
        }
      }
    } else {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        const int c0 = n0 + j * 32;
        if (c0 >= p.n) break;  // warp-uniform
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + lane_base + (uint32_t)(j * 32), raw);
        tmem_ld_wait();
        if (!ri.valid) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        epilogue_chunk<MODE, T>(p, ri, yrow, inv_norm, add_row, s_alpha, s_beta, j, c0, min(32, p.n - c0), v, sq_acc);
      }
    }
    if (MODE == BCOSK_MODE_FWD) {
      if (p.sq_out != nullptr && ri.valid) p.sq_out[(size_t)tile_n * M + ri.m] = sq_acc;
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// debug: land one A chunk in smem via the im2col tensor map and copy it out de-swizzled
// ---------------------------------------------------------------------------------------------
__global__ void bcosk_debug_a_tile_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                          const __grid_constant__ bcosk_igemm_params p, int tile_m, int chunk,
                                          uint16_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int m0 = tile_m * BM;
    const int opq = p.op * p.oq;
    const int img = m0 / opq, rem = m0 % opq;
    const int pp = rem / p.oq, qq = rem % p.oq;
    const int kc = chunk % p.chunks_per_tap;
    const int t = chunk / p.chunks_per_tap;
    const int tap = t % p.num_taps, seg = t / p.num_taps;
    mbar_arrive_expect_tx(&bar, BM * p.kch * 2);
    tma_load_im2col_4d(smem, &tmap_a, &bar, p.seg_a_choff[seg] + kc * p.kch, p.lo_w + qq * p.stride_w,
                       p.lo_h + pp * p.stride_h, img, p.tap_off_w[tap], p.tap_off_h[tap]);
  }
  mbar_wait(&bar, 0);
  const int row_bytes = p.kch * 2;
  for (int i = threadIdx.x; i < BM * p.kch; i += blockDim.x) {
    const int r = i / p.kch, c = i % p.kch;
    const int u = (c * 2) / 16;  // logical 16-byte unit inside the row
    // Swizzle<3,4,3> (128B rows): unit ^= row % 8;  Swizzle<2,4,3> (64B rows): unit ^= (row / 2) % 4
    const int pu = (row_bytes == 128) ? (u ^ (r & 7)) : (u ^ ((r >> 1) & 3));
    const int byte = r * row_bytes + pu * 16 + (c * 2) % 16;
    out[i] = *reinterpret_cast<uint16_t*>(smem + byte);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN, int MODE, bool HP>
static int launch_igemm(const CUtensorMap& ta, const CUtensorMap& tb, const bcosk_igemm_params& p, cudaStream_t st) {
  using Cfg = TileCfg<BN, HP>;
  auto kern = bcosk_igemm_kernel<BN, MODE, __nv_bfloat16, HP>;
  auto kern_h = bcosk_igemm_kernel<BN, MODE, __half, HP>;
  const void* fn = (p.dtype == BCOSK_DTYPE_BF16) ? (const void*)kern : (const void*)kern_h;
  static bool attr_done[2] = {false, false};
  if (!attr_done[p.dtype]) {
    BCOSK_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done[p.dtype] = true;
  }
  const long long M = (long long)p.a_nb * p.op * p.oq;
  const long long m_tiles = (M + BM - 1) / BM;
  const long long n_tiles = (p.n + BN - 1) / BN;
  if (m_tiles * n_tiles > 0x7fffffffLL) return set_error(BCOSK_EUNSUPPORTED, "igemm: grid too large");
  dim3 grid((unsigned)(m_tiles * n_tiles));
  if (p.dtype == BCOSK_DTYPE_BF16)
    kern<<<grid, NUM_THREADS, Cfg::kSmemBytes, st>>>(ta, tb, p);
  else
    kern_h<<<grid, NUM_THREADS, Cfg::kSmemBytes, st>>>(ta, tb, p);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}

static int validate(const bcosk_igemm_params& p) {
  if (!p.a || !p.b || !p.y) return set_error(BCOSK_EINVAL, "igemm: null a/b/y");
  if (p.kch != 64 && p.kch != 32) return set_error(BCOSK_EINVAL, "igemm: kch must be 32 or 64");
  if (p.num_taps < 1 || p.num_taps > BCOSK_MAX_TAPS) return set_error(BCOSK_EINVAL, "igemm: num_taps out of range");
  if (p.num_segs < 1 || p.num_segs > BCOSK_MAX_SEGS) return set_error(BCOSK_EINVAL, "igemm: num_segs out of range");
  if (p.chunks_per_tap < 1) return set_error(BCOSK_EINVAL, "igemm: chunks_per_tap < 1");
  const int total = p.num_segs * p.num_taps * p.chunks_per_tap;
  if (total % (STAGE_K / p.kch) != 0) return set_error(BCOSK_EINVAL, "igemm: chunk count not a multiple of the stage");
  if (p.a_c % 8 != 0) return set_error(BCOSK_EINVAL, "igemm: a_c must be a multiple of 8 (16-byte pixels)");
  if (p.n < 8 || p.n % 8 != 0) return set_error(BCOSK_EINVAL, "igemm: n must be a positive multiple of 8");
  if (p.dtype != BCOSK_DTYPE_BF16 && p.dtype != BCOSK_DTYPE_F16) return set_error(BCOSK_EINVAL, "igemm: dtype");
  if (p.mode == BCOSK_MODE_FWD && p.scale_mode != BCOSK_SCALE_NONE && !p.inv_norm)
    return set_error(BCOSK_EINVAL, "igemm: inv_norm required for a B-cos scale");
  if (p.y_ld % (p.y_f32 ? 4 : 8) != 0) return set_error(BCOSK_EINVAL, "igemm: y_ld alignment");
  if (p.a_nb < 1 || p.op < 1 || p.oq < 1) return set_error(BCOSK_EINVAL, "igemm: empty problem");
  return BCOSK_OK;
}

static int make_maps(const bcosk_igemm_params& p, int bn, CUtensorMap* ta, CUtensorMap* tb) {
  int rc = make_im2col_map_nhwc(ta, p.a, p.a_nb, p.a_h, p.a_w, p.a_c, p.lo_w, p.lo_h, p.up_w, p.up_h, p.stride_w,
                                p.stride_h, p.kch, BM, p.kch == 64 ? 128 : 64);
  if (rc) return rc;
  const long long ktot = (long long)p.num_segs * p.num_taps * p.chunks_per_tap * p.kch;
  return make_tiled_map_2d(tb, p.b, ktot, p.n, p.kch, bn, p.kch == 64 ? 128 : 64);
}

}  // namespace bcosk

using namespace bcosk;

extern "C" int bcosk_igemm(const bcosk_igemm_params* pp, void* stream) {
  if (!pp) return set_error(BCOSK_EINVAL, "igemm: null params");
  bcosk_igemm_params p = *pp;
  int rc = validate(p);
  if (rc) return rc;
  int bn = p.block_n;
  if (bn == 0) bn = p.n <= 32 ? 32 : ((p.n <= 64 || p.hp_accum) ? 64 : 128);
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) return set_error(BCOSK_EINVAL, "igemm: block_n");
  if (p.hp_accum && bn > 64) return set_error(BCOSK_EINVAL, "igemm: hp_accum needs block_n <= 64");
  p.block_n = bn;
  CUtensorMap ta, tb;
  rc = make_maps(p, bn, &ta, &tb);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define BCOSK_DISPATCH(BN_)                                                                        \
  case BN_:                                                                                        \
    return p.mode == BCOSK_MODE_FWD ? launch_igemm<BN_, BCOSK_MODE_FWD, false>(ta, tb, p, st)      \
                                    : launch_igemm<BN_, BCOSK_MODE_EXPLAIN, false>(ta, tb, p, st);
  if (p.hp_accum) {
    if (bn == 32)
      return p.mode == BCOSK_MODE_FWD ? launch_igemm<32, BCOSK_MODE_FWD, true>(ta, tb, p, st)
                                      : launch_igemm<32, BCOSK_MODE_EXPLAIN, true>(ta, tb, p, st);
    return p.mode == BCOSK_MODE_FWD ? launch_igemm<64, BCOSK_MODE_FWD, true>(ta, tb, p, st)
                                    : launch_igemm<64, BCOSK_MODE_EXPLAIN, true>(ta, tb, p, st);
  }
  switch (bn) {
    BCOSK_DISPATCH(32)
    BCOSK_DISPATCH(64)
    BCOSK_DISPATCH(128)
    BCOSK_DISPATCH(256)
  }
#undef BCOSK_DISPATCH
  return set_error(BCOSK_EINVAL, "igemm: unreachable");
}

extern "C" int bcosk_debug_a_tile(const bcosk_igemm_params* pp, int32_t tile_m, int32_t chunk, void* out, void* stream) {
  if (!pp || !out) return set_error(BCOSK_EINVAL, "debug_a_tile: null");
  bcosk_igemm_params p = *pp;
  CUtensorMap ta;
  int rc = make_im2col_map_nhwc(&ta, p.a, p.a_nb, p.a_h, p.a_w, p.a_c, p.lo_w, p.lo_h, p.up_w, p.up_h, p.stride_w,
                                p.stride_h, p.kch, BM, p.kch == 64 ? 128 : 64);
  if (rc) return rc;
  bcosk_debug_a_tile_kernel<<<1, 128, BM * 128 + 1024, reinterpret_cast<cudaStream_t>(stream)>>>(
      ta, p, tile_m, chunk, reinterpret_cast<uint16_t*>(out));
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}
