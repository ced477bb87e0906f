// bcosk_wgrad.cu -- weight gradient of a (B-cos) convolution / linear map as a tcgen05 implicit GEMM, sm_100a.
//
//   dW[o, (tap, c)] += sum_m  g_lin[m, o] * x_patch[m, (tap, c)]          (torch: ConvolutionBackward's grad_weight,
//                                                                           reached from the reference's training_step,
//                                                                           bcos/training/trainer.py:666-784)
//   m = output pixel of the forward convolution; x_patch gathered by the SAME TMA im2col traversal the forward launch uses.
//
// Both operands are stored pixel-major (NHWC: a row per pixel), i.e. the contraction index is the SLOW axis of both:
// they are fed to the tensor core as MN-major operands (instruction descriptor a_major = b_major = 1) straight from the
// [64 pixels][64 channels] SWIZZLE_128B boxes the TMA unit lands - no transpose pass exists.
//   CTA tile: 128 output units x BN gradient columns (BN = 128: two 64-channel chunks of one tap, 64 or 32: one chunk),
//   K stage = 64 pixels (four K = 16 MMAs), 3-4 stage mbarrier ring, fp32 accumulator in TMEM.
//   grid = unit tiles x column tiles x split-K over the pixel blocks; every CTA adds its partial tile into dW with
//   16-byte vector reductions (red.global.add.v4.f32).  The caller zeroes dW.
//   warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include <cuda.h>
#include <cstring>

#include "../../include/bcosk.h"
#include "bcosk_common.cuh"
#include "bcosk_host.h"

namespace bcosk {
namespace wg {
constexpr int BM = 128;          // output units per tile
constexpr int KP = 64;           // pixels per K stage
constexpr int A_BYTES = BM * KP * 2;     // 16 KB: two [64 pixels][64 units] boxes
constexpr int THREADS = 192;
}  // namespace wg

struct WgradAux {
  int bn;            // gradient columns per tile: 128 / 64 / 32
  int stages;
  int split;         // CTAs along the pixel axis
  int blocks_total;  // 64-pixel blocks
  int blocks_per;    // per split
  int col_tiles;
};

// MN-major shared-memory descriptor: rows of `row_bytes` (= swizzle span) hold 64 (32) consecutive MN elements of one
// K index; 8 K-rows form a swizzle atom (SBO = 8 * row_bytes apart), the next group of MN elements is LBO bytes away.
__device__ __forceinline__ uint64_t umma_smem_desc_mnmajor(uint32_t smem_addr, uint32_t row_bytes, uint32_t lbo_bytes) {
  const uint64_t layout = (row_bytes == 128) ? 2ull : 4ull;
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(((8u * row_bytes) >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(wg::THREADS)
bcosk_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                   const __grid_constant__ bcosk_wgrad_params p, const WgradAux aux) {
  using namespace wg;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int BN = aux.bn;
  const int row_bytes = p.kch * 2;                       // bytes per pixel row of an X box (128 or 64)
  const int b_bytes = (BN / p.kch) * KP * row_bytes;     // X boxes of one stage
  const int slot = A_BYTES + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + aux.stages * slot);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* done_bar = empty_bar + 8;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col_tile = blockIdx.x % aux.col_tiles;
  const int rest = blockIdx.x / aux.col_tiles;
  const int o_tiles = (p.n + BM - 1) / BM;
  const int o_tile = rest % o_tiles;
  const int sp = rest / o_tiles;
  const int o0 = o_tile * BM;
  const int chunks_per_tile = BN / p.kch;                // 1 or 2
  const int chunk0 = col_tile * chunks_per_tile;         // global chunk index = tap * chunks_per_tap + kc
  const int tap = chunk0 / p.chunks_per_tap;
  const int kc0 = chunk0 - tap * p.chunks_per_tap;
  const int blk_begin = sp * aux.blocks_per;
  const int blk_end = min(aux.blocks_total, blk_begin + aux.blocks_per);
  const int iters = blk_end - blk_begin;
  const uint32_t tmem_cols = BN < 32 ? 32u : (uint32_t)BN;
  const int a_boxes = (p.n - o0) > 64 ? 2 : 1;           // 64-unit boxes of g_lin that lie inside the tensor

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_g);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < aux.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const int opq = p.op * p.oq;
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
          const int m0 = (blk_begin + it) * KP;
          const int img = m0 / opq;
          const int rem = m0 - img * opq;
          const int pp = rem / p.oq, qq = rem - pp * p.oq;
          uint8_t* sl = smem + stage * slot;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(a_boxes * (A_BYTES / 2) + b_bytes));
          for (int b = 0; b < a_boxes; ++b) tma_load_2d(sl + b * (A_BYTES / 2), &tmap_g, &full_bar[stage], o0 + b * 64, m0);
          for (int j = 0; j < chunks_per_tile; ++j)
            tma_load_im2col_4d(sl + A_BYTES + j * KP * row_bytes, &tmap_x, &full_bar[stage], (kc0 + j) * p.kch,
                               p.lo_w + qq * p.stride_w, p.lo_h + pp * p.stride_h, img, p.tap_off_w[tap], p.tap_off_h[tap]);
          if (++stage == aux.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // D[M = units, N = gradient columns] += A[units, pixels] * B[columns, pixels], both MN-major
        const uint32_t fmt = (uint32_t)p.dtype;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | (((uint32_t)BN >> 3) << 17) |
                               (((uint32_t)BM >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t accumulate = 0;
        for (int it = 0; it < iters; ++it) {
          uint8_t* sl = smem + stage * slot;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_smem_desc_mnmajor(smem_u32(sl), 128, A_BYTES / 2);
          const uint64_t db = umma_smem_desc_mnmajor(smem_u32(sl + A_BYTES), (uint32_t)row_bytes, (uint32_t)(KP * row_bytes));
          const uint64_t a_step = (16u * 128u) >> 4, b_step = (16u * (uint32_t)row_bytes) >> 4;   // 16 pixel rows
#pragma unroll
          for (int k = 0; k < KP / 16; ++k) {
            umma_f16(tmem_base, da + k * a_step, db + k * b_step, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == aux.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(done_bar);
      }
    } else {
      const int quad = warp & 3;
      const int row = quad * 32 + lane;
      const int o = o0 + row;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int ktot = p.num_taps * p.chunks_per_tap * p.kch;
      float* dst = p.dw + (size_t)o * ktot + (size_t)chunk0 * p.kch;
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * 32), raw);
        tmem_ld_wait();
        if (o < p.n) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            red_add_v4(dst + j * 32 + q * 4, __uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                       __uint_as_float(raw[4 * q + 3]));
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

static int g_wgrad_sms = 0;

}  // namespace bcosk

using namespace bcosk;

extern "C" int bcosk_wgrad(const bcosk_wgrad_params* pp, void* stream) {
  using namespace wg;
  if (!pp) return set_error(BCOSK_EINVAL, "wgrad: null params");
  bcosk_wgrad_params p = *pp;
  if (!p.x || !p.g || !p.dw) return set_error(BCOSK_EINVAL, "wgrad: null x / g / dw");
  if (p.kch != 64 && p.kch != 32) return set_error(BCOSK_EINVAL, "wgrad: kch must be 32 or 64");
  if (p.num_taps < 1 || p.num_taps > BCOSK_MAX_TAPS || p.chunks_per_tap < 1) return set_error(BCOSK_EINVAL, "wgrad: bad tap geometry");
  if (p.n < 8 || p.n % 8 != 0 || p.g_ld % 8 != 0 || p.a_c % 8 != 0) return set_error(BCOSK_EINVAL, "wgrad: channel counts must be multiples of 8");
  if (p.dtype != BCOSK_DTYPE_BF16 && p.dtype != BCOSK_DTYPE_F16) return set_error(BCOSK_EINVAL, "wgrad: dtype");
  if (p.a_nb < 1 || p.op < 1 || p.oq < 1) return set_error(BCOSK_EINVAL, "wgrad: empty problem");
  if ((reinterpret_cast<uintptr_t>(p.dw) & 15) != 0) return set_error(BCOSK_EINVAL, "wgrad: dw must be 16-byte aligned");
  const long long M = (long long)p.a_nb * p.op * p.oq;
  WgradAux aux;
  memset(&aux, 0, sizeof(aux));
  aux.bn = (p.kch == 64 && p.chunks_per_tap % 2 == 0) ? 128 : p.kch;
  const int chunks_per_tile = aux.bn / p.kch;
  aux.col_tiles = p.num_taps * p.chunks_per_tap / chunks_per_tile;
  aux.blocks_total = (int)((M + KP - 1) / KP);
  const int o_tiles = (p.n + BM - 1) / BM;
  if (g_wgrad_sms == 0) {
    int dev = 0;
    BCOSK_CUDA_CHECK(cudaGetDevice(&dev));
    BCOSK_CUDA_CHECK(cudaDeviceGetAttribute(&g_wgrad_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int split = p.split_k;
  if (split <= 0) {
    // ~4 waves of CTAs, at least 8 pixel blocks per CTA so that the reductions stay a small part of the work
    const long long tiles = (long long)o_tiles * aux.col_tiles;
    split = (int)((4LL * g_wgrad_sms + tiles - 1) / tiles);
    const int max_split = aux.blocks_total / 8 > 0 ? aux.blocks_total / 8 : 1;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
  }
  aux.blocks_per = (aux.blocks_total + split - 1) / split;
  aux.split = (aux.blocks_total + aux.blocks_per - 1) / aux.blocks_per;
  const int row_bytes = p.kch * 2;
  const int slot = A_BYTES + chunks_per_tile * KP * row_bytes;
  aux.stages = aux.bn == 128 ? 3 : 4;
  const int smem = aux.stages * slot + 256;

  CUtensorMap mg, mx;
  int rc = make_tiled_map_2d(&mg, p.g, p.g_ld, M, 64, KP, 128);
  if (rc) return rc;
  rc = make_im2col_map_nhwc(&mx, p.x, p.a_nb, p.a_h, p.a_w, p.a_c, p.lo_w, p.lo_h, p.up_w, p.up_h, p.stride_w, p.stride_h, p.kch, KP,
                            row_bytes);
  if (rc) return rc;
  BCOSK_CUDA_CHECK(cudaFuncSetAttribute(bcosk_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (A_BYTES + 16384) + 256));
  const long long grid = (long long)aux.col_tiles * o_tiles * aux.split;
  if (grid > 0x7fffffffLL) return set_error(BCOSK_EUNSUPPORTED, "wgrad: grid too large");
  bcosk_wgrad_kernel<<<(unsigned)grid, THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(mg, mx, p, aux);
  BCOSK_CUDA_CHECK(cudaGetLastError());
  return BCOSK_OK;
}
