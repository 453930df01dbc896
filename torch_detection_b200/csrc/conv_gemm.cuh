// Implicit-GEMM convolution for sm_100a: TMA (tiled / im2col) -> shared memory ring ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> fused epilogue (BN scale/shift or bias, residual add,
// FPN nearest-x2 upsample-add, ReLU) -> NHWC bf16.
//
// GEMM view: D[M = N_img*Ho*Wo, Cout] = A[M, K = kh*kw*Cin] * W[Cout, K]^T.
// A rows are output pixels (NHWC), gathered by the TMA unit:
//   A_TILED  : 1x1 stride-1 conv -> A is the activation matrix itself (2D tiled tensor map)
//   A_IM2COL : kxk / strided conv -> 4D im2col tensor map, one load per (filter tap, 64-ch chunk)
//   A_STEM   : 7x7/2 stem on the padded NHWC4 staging -> 5D tiled map with overlapping windows
// W is pre-packed [Cout][kh][kw][Cin] bf16, i.e. K-major rows, loaded by a 2D tiled map.
//
// Persistent, warp-specialised CTA (192 threads):
//   warp 0      TMA producer            (full/empty mbarrier ring, kStages deep)
//   warp 1      tcgen05.mma issuer      (lane 0 issues; accumulators double-buffered in TMEM)
//   warps 2..5  epilogue                (tcgen05.ld 32x32b -> registers -> global), overlapped
//                                        with the next tile's main loop through tmem_full/empty.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx_sm100.cuh"

namespace tdet {

constexpr int kBM = 128;         // UMMA M (cta_group::1)
constexpr int kBK = 64;          // bf16 elements per 128-byte swizzle row
constexpr int kUmmaK = 16;       // K per tcgen05.mma for 16-bit operands
constexpr int kGemmThreads = 192;
constexpr int kABytes = kBM * kBK * 2;  // 16 KiB per stage

enum AMode : int { A_TILED = 0, A_IM2COL = 1, A_STEM = 2 };

// Operand formats.  Activations (A) are bf16 (8-bit exponent: the residual stream of a randomly
// initialised ResNet-101 reaches ~2.4e5).  Weights (B) use the same format: a mixed
// bf16 x fp16 tcgen05.mma (idesc a_format != b_format) raises "illegal instruction" on B200
// (measured), so TDET_WEIGHT_FP16=1 is only usable together with fp16 activations.
#ifndef TDET_WEIGHT_FP16
#define TDET_WEIGHT_FP16 0
#endif
#if TDET_WEIGHT_FP16
constexpr uint32_t kWeightFmt = kFmtF16;
using weight_t = __half;
__device__ __forceinline__ weight_t to_weight(float v) { return __float2half_rn(v); }
#else
constexpr uint32_t kWeightFmt = kFmtBF16;
using weight_t = __nv_bfloat16;
__device__ __forceinline__ weight_t to_weight(float v) { return __float2bfloat16_rn(v); }
#endif

struct ConvGemmParams {
  CUtensorMap tmap_a;
  CUtensorMap tmap_b;
  int M;            // valid output rows (pixels); for A_STEM rows are masked per pixel instead
  int N;            // Cout
  int num_m_tiles;
  int num_n_tiles;
  int k_chunks;     // Cin / 64 (A_STEM: 1)
  int kh, kw, dil;
  int cin;          // B column offset of tap (r,s) is (r*kw + s)*cin
  int a_mode;
  int Ho, Wo;       // output spatial size
  int stride, pad;
  int tiles_w, tiles_h;  // A_STEM: spatial tiles per image
  int tile_bw, tile_bh;  // A_STEM: tile shape in output pixels (tile_bw * tile_bh == 128)
  int Hc, Wc;       // coarse level size (upsample-add)
  int relu;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* coarse;
  __nv_bfloat16* out;
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kNumBars = 2 * STAGES + 4;
  static constexpr int kTmemPtrOffset = kBarOffset + kNumBars * 8;
  static constexpr int kParamOffset = (kTmemPtrOffset + 16 + 15) / 16 * 16;
  static constexpr int kTotal = kParamOffset + 2 * BN * 4;
  static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024-byte alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using L = GemmSmem<BN, STAGES>;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t kIdesc = make_idesc_f16kind(kBM, BN, kFmtBF16, kWeightFmt);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t smem_a = base;
  const uint32_t smem_b = base + STAGES * kABytes;
  const uint32_t bar0 = base + L::kBarOffset;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::kTmemPtrOffset);
  float* s_scale = reinterpret_cast<float*>(smem + L::kParamOffset);
  float* s_shift = s_scale + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
    }
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

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = p.kh * p.kw * p.k_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      const int n0 = n_tile * BN;
      // tile origin in the A coordinate space
      int cw = 0, ch = 0, cn = 0;
      if (p.a_mode == A_IM2COL) {
        const int m0 = m_tile * kBM;
        const int q0 = m0 % p.Wo;
        const int t = m0 / p.Wo;
        const int p0 = t % p.Ho;
        cn = t / p.Ho;
        cw = q0 * p.stride - p.pad;
        ch = p0 * p.stride - p.pad;
      } else if (p.a_mode == A_STEM) {
        const int tw = m_tile % p.tiles_w;
        const int t = m_tile / p.tiles_w;
        const int th = t % p.tiles_h;
        cn = t / p.tiles_h;
        cw = tw * p.tile_bw;
        ch = th * p.tile_bh;
      }
      int kb = 0;
      for (int r = 0; r < p.kh; ++r) {
        for (int s = 0; s < p.kw; ++s) {
          for (int kc = 0; kc < p.k_chunks; ++kc, ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (lane == 0) {
              const uint32_t fb = full_bar(stage);
              mbar_arrive_expect_tx(fb, L::kStageBytes);
              const uint32_t dst_a = smem_a + stage * kABytes;
              const uint32_t dst_b = smem_b + stage * L::kBBytes;
              if (p.a_mode == A_TILED) {
                tma_load_2d(dst_a, &p.tmap_a, fb, kc * kBK, m_tile * kBM);
              } else if (p.a_mode == A_IM2COL) {
                tma_load_im2col_4d(dst_a, &p.tmap_a, fb, kc * kBK, cw, ch, cn,
                                   static_cast<uint16_t>(s * p.dil),
                                   static_cast<uint16_t>(r * p.dil));
              } else {
                // filter row r of the 7x7 window = staged image row 2*(ho + r/2) + (r & 1)
                tma_load_5d(dst_a, &p.tmap_a, fb, 0, r & 1, cw, ch + (r >> 1), cn);
              }
              tma_load_2d(dst_b, &p.tmap_b, fb, (r * p.kw + s) * p.cin + kc * kBK, n0);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = make_smem_desc_sw128(smem_a + stage * kABytes);
          const uint64_t db = make_smem_desc_sw128(smem_b + stage * L::kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 32 bytes (16 bf16) along K inside the 128-byte swizzle row: +2 in 16-byte units
            umma_bf16_ss(d_tmem, da + 2u * k, db + 2u * k, kIdesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;      // accumulator row == TMEM lane
    const int epi_tid = threadIdx.x - 64;  // 0..127
    int acc = 0;
    uint32_t acc_phase = 0;
    int cur_n_tile = -1;
    const bool has_res = p.residual != nullptr;
    const bool has_coarse = p.coarse != nullptr;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      const int n0 = n_tile * BN;
      if (n_tile != cur_n_tile) {
        named_bar_sync(1, 128);
        for (int i = epi_tid; i < BN; i += 128) {
          s_scale[i] = p.scale ? __ldg(p.scale + n0 + i) : 1.0f;
          s_shift[i] = p.shift ? __ldg(p.shift + n0 + i) : 0.0f;
        }
        named_bar_sync(1, 128);
        cur_n_tile = n_tile;
      }
      // output row of this thread
      bool valid;
      long long pix;  // linear NHWC pixel index of the output row
      if (p.a_mode == A_STEM) {
        const int tw = m_tile % p.tiles_w;
        const int t = m_tile / p.tiles_w;
        const int th = t % p.tiles_h;
        const int img = t / p.tiles_h;
        const int hl = row / p.tile_bw;
        const int wl = row - hl * p.tile_bw;
        const int ho = th * p.tile_bh + hl;
        const int wo = tw * p.tile_bw + wl;
        valid = (ho < p.Ho) && (wo < p.Wo);
        pix = (static_cast<long long>(img) * p.Ho + ho) * p.Wo + wo;
      } else {
        const int m = m_tile * kBM + row;
        valid = m < p.M;
        pix = m;
      }
      const __nv_bfloat16* res_row = nullptr;
      const __nv_bfloat16* coarse_row = nullptr;
      if (valid && has_res) res_row = p.residual + pix * p.N + n0;
      if (valid && has_coarse) {
        const int m = static_cast<int>(pix);
        const int q = m % p.Wo;
        const int t = m / p.Wo;
        const int pp = t % p.Ho;
        const int img = t / p.Ho;
        coarse_row = p.coarse + ((static_cast<long long>(img) * p.Hc + (pp >> 1)) * p.Wc + (q >> 1)) * p.N + n0;
      }
      __nv_bfloat16* out_row = p.out + pix * p.N + n0;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                              static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr + chunk * 32, v);
        uint4 rres[4], rco[4];
        if (res_row) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rres[j] = ldg_nc_v4(res_row + chunk * 32 + j * 8);
        }
        if (coarse_row) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rco[j] = ldg_nc_v4(coarse_row + chunk * 32 + j * 8);
        }
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + chunk * 32 + j * 4);
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + chunk * 32 + j * 4);
          x[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, sh.x);
          x[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y);
          x[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z);
          x[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w);
        }
        if (res_row) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w4[4] = {rres[j].x, rres[j].y, rres[j].z, rres[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              x[8 * j + 2 * e] += bf16_lo(w4[e]);
              x[8 * j + 2 * e + 1] += bf16_hi(w4[e]);
            }
          }
        }
        if (coarse_row) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w4[4] = {rco[j].x, rco[j].y, rco[j].z, rco[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              x[8 * j + 2 * e] += bf16_lo(w4[e]);
              x[8 * j + 2 * e + 1] += bf16_hi(w4[e]);
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = fmaxf(x[i], 0.0f);
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(x[8 * j + 0], x[8 * j + 1]);
            o.y = pack_bf16x2(x[8 * j + 2], x[8 * j + 3]);
            o.z = pack_bf16x2(x[8 * j + 4], x[8 * j + 5]);
            o.w = pack_bf16x2(x[8 * j + 6], x[8 * j + 7]);
            stg_v4(out_row + chunk * 32 + j * 8, o);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Debug/test kernel: load ONE im2col A tile and write it out un-swizzled as [128][64] bf16.
__global__ void __launch_bounds__(128, 1)
im2col_tile_dump_kernel(const __grid_constant__ CUtensorMap tmap, int c, int w, int h, int n,
                        int off_w, int off_h, __nv_bfloat16* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar = base + kABytes;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, kABytes);
    tma_load_im2col_4d(base, &tmap, bar, c, w, h, n, static_cast<uint16_t>(off_w),
                       static_cast<uint16_t>(off_h));
  }
  mbar_wait(bar, 0);
  // row r, 16-byte chunk j lives at chunk (j ^ (r & 7)) of the 128-byte row
  const int r = threadIdx.x;
  for (int j = 0; j < 8; ++j) {
    const uint4 val = *reinterpret_cast<const uint4*>(smem + r * 128 + ((j ^ (r & 7)) << 4));
    *reinterpret_cast<uint4*>(out + r * 64 + j * 8) = val;
  }
}

}  // namespace tdet
