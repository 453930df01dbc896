// Weight-gradient implicit GEMM for sm_100a (training path, SURVEY.md section 8 rows a3/a4/a10/a12, e):
//
//   dW[co][r][s][ci] += scale[co] * sum_m  g[m][co] * X_tap(r,s)[m][ci]        m = (n, p, q) output pixels
//
// GEMM view: D[M' = 128 output channels, N' = NB input channels] over K' = pixels.  Both operands
// are "MN-major" for tcgen05: the pixel (K') index is the slow dimension of the NHWC tensors, so
// the shared-memory tiles TMA deposits -- PIX pixel rows x 128 B (64 channels), SWIZZLE_128B -- are
// consumed with a_major = b_major = MN: 64 channels contiguous per row, 8-pixel groups 1024 B apart
// (SBO), the next 64 channels in the next slab (LBO = PIX*128 B).  X is gathered by the same kind of
// tensor map the forward A operand uses (tiled 2D for 1x1 stride 1, im2col 4D otherwise), so padding
// and stride semantics are the forward kernel's by construction.  g and X must share one 16-bit format
// (bf16, or fp16 with per-tensor exponents applied in the epilogue): tcgen05 kind::f16 traps on mixed
// fp16/bf16 operands (measured).
//
// One CTA = one (MT x 128-channel Cout tile, filter tap, NB-wide Cin group) and one slice of the pixel
// range (split-K over blockIdx.y); partial sums are combined with vectorised fp32 reductions
// (red.global.add.v4.f32) into the [Cout][kh][kw][Cin] accumulator, which the caller zeroes.
// MT = 2 keeps two accumulators (2 x NB TMEM columns) that share every X tile: 256 x 256 output tiles need 64 B of
// operands per tensor cycle instead of the 96 B of 128 x 256 (big layers: tensor pipe 85 % of elapsed cycles).  The
// split-K factor and, for layers with few pixels, the tile shape come from a cost model in build_wgrad (tdet_api.cu):
// such layers are bound by the reduction traffic (CTAs x tile bytes at ~1 TB/s), not by the main loop.
//   warp 0  TMA producer      warp 1  tcgen05.mma issuer      warps 2..5  epilogue (TMEM -> red.add)
#pragma once
#include "conv_gemm.cuh"

namespace tdet {

constexpr int kWgThreads = 192;

struct WgradParams {
  CUtensorMap tmap_g;  // 2D (Cout, M) box (64, PIX) over the output gradient
  CUtensorMap tmap_x;  // conv input: tiled 2D (Cin, M) box (64, PIX), or im2col 4D (64 ch x PIX pixels)
  int M;               // pixels: n * Ho * Wo
  int cout, cin, kh, kw, dil, stride, pad, Ho, Wo;
  int x_im2col;        // 0 = tiled 2D, 1 = im2col 4D
  int ci_groups;       // cin / NB
  int kblocks;         // ceil(M / PIX)
  int kb_per_cta;      // k-blocks per split
  int g_fp16, x_fp16;
  float* dw;           // fp32 [Cout][kh][kw][Cin], accumulated
  const float* scale;  // per-Cout multiplier (folded BN scale) or null
  const TensorMeta* x_meta;  // nullable: X is stored * 2^e
  const TensorMeta* g_meta;  // nullable
};

// MN-major SWIZZLE_128B operand: `lbo` = bytes between 64-element groups along M/N (slab stride),
// 8-row K groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int NB, int PIX, int STAGES, int MT>
struct WgradSmem {
  static constexpr int kSlab = PIX * 128;
  static constexpr int kStageBytes = (2 * MT + NB / 64) * kSlab;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kDynamic = kBarOffset + (2 * STAGES + 1) * 8 + 16;
  static_assert(kDynamic <= 232448, "exceeds the 227 KiB shared memory limit");
  static_assert(kSlab % 1024 == 0, "slabs must keep the 1024-byte swizzle alignment");
};

template <int NB, int PIX, int STAGES, int MT>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
  using L = WgradSmem<NB, PIX, STAGES, MT>;
  constexpr uint32_t kTmemCols = MT * NB < 32 ? 32 : MT * NB;
  static_assert(MT * NB <= 512 && (MT * NB & (MT * NB - 1)) == 0, "TMEM allocation: power of two <= 512 columns");
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t bar0 = base + L::kBarOffset;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * STAGES);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::kBarOffset + (2 * STAGES + 1) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile decode
  int t = blockIdx.x;
  const int cig = t % p.ci_groups;
  t /= p.ci_groups;
  const int taps = p.kh * p.kw;
  const int tap = t % taps;
  const int co_tile = t / taps;
  const int r = tap / p.kw, s = tap - r * p.kw;
  const int kb0 = blockIdx.y * p.kb_per_cta;
  const int kb1 = min(p.kblocks, kb0 + p.kb_per_cta);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_g);
    tma_prefetch_desc(&p.tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  grid_dependency_wait();

  if (nkb > 0) {
    if (warp == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (lane == 0) {
          const uint32_t fb = full_bar(stage);
          mbar_arrive_expect_tx(fb, L::kStageBytes);
          const uint32_t a = base + stage * L::kStageBytes;
          const uint32_t b = a + 2 * MT * L::kSlab;
          const int m0 = kb * PIX;
#pragma unroll
          for (int j = 0; j < 2 * MT; ++j)
            tma_load_2d(a + j * L::kSlab, &p.tmap_g, fb, co_tile * 128 * MT + j * 64, m0);
          if (p.x_im2col) {
            const int q0 = m0 % p.Wo;
            const int tt = m0 / p.Wo;
            const int p0 = tt % p.Ho;
            const int img = tt / p.Ho;
            for (int j = 0; j < NB / 64; ++j)
              tma_load_im2col_4d(b + j * L::kSlab, &p.tmap_x, fb, cig * NB + j * 64, q0 * p.stride - p.pad,
                                 p0 * p.stride - p.pad, img, static_cast<uint16_t>(s * p.dil),
                                 static_cast<uint16_t>(r * p.dil));
          } else {
            for (int j = 0; j < NB / 64; ++j)
              tma_load_2d(b + j * L::kSlab, &p.tmap_x, fb, cig * NB + j * 64, m0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    } else if (warp == 1) {
      // a_major (bit 15) = b_major (bit 16) = MN
      const uint32_t idesc = make_idesc_f16kind(128, NB, p.g_fp16 ? kFmtF16 : kFmtBF16,
                                                p.x_fp16 ? kFmtF16 : kFmtBF16) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a = base + stage * L::kStageBytes;
          const uint32_t b = a + 2 * MT * L::kSlab;
#pragma unroll
          for (int k = 0; k < PIX / kUmmaK; ++k) {
            // 16 pixels per MMA = two 8-row groups = 2048 B further down each slab
            const uint64_t db = make_smem_desc_mn_sw128(b + k * 2048, L::kSlab);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint64_t da = make_smem_desc_mn_sw128(a + mt * 2 * L::kSlab + k * 2048, L::kSlab);
              umma_bf16_ss(tmem_base + mt * NB, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          if (kb == nkb - 1) umma_commit(done_bar);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    } else {
      const int quad = warp & 3;
      int e = 0;
      if (p.x_meta) e += p.x_meta->e;
      if (p.g_meta) e += p.g_meta->e;
      mbar_wait(done_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int co = (co_tile * MT + mt) * 128 + quad * 32 + lane;
        float sc = (co < p.cout && p.scale) ? __ldg(p.scale + co) : 1.0f;
        sc = ldexpf(sc, e);
        float* dst_row = p.dw + (static_cast<long long>(co) * taps + tap) * p.cin + cig * NB;
#pragma unroll 1
        for (int c = 0; c < NB / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + mt * NB + c * 32, v);
          tmem_ld_wait();
          if (co < p.cout) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              red_add_v4(dst_row + c * 32 + j * 4, __uint_as_float(v[4 * j]) * sc, __uint_as_float(v[4 * j + 1]) * sc,
                         __uint_as_float(v[4 * j + 2]) * sc, __uint_as_float(v[4 * j + 3]) * sc);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tdet
