// Operand-swapped implicit-GEMM conv for narrow outputs (Cout = 64 or 128; SURVEY.md section 8 rows a3/a5: conv1 and
// conv2 of the bottlenecks of layer1/layer2, resnet.py:97-108).
//
// Measured on B200 (tools/probe_umma_m64.cu): a 128 x 128 x 16 MMA reads 8 KB of operands per
// 64 cycles -- all of the SM's shared-memory bandwidth -- so inside a conv kernel, where TMA fills and the epilogue
// share that bandwidth, 128-wide tiles reach ~45 % of the tensor rate (145 cycles per MMA), while the 256-wide shape
// (12 KB per 128 cycles) runs at its math rate.  Here the roles are swapped:
//
//     D^T[Cout (M = 128 lanes), 256 pixels (N)] = W[Cout, K] * A[256 pixels, K]^T
//
// so every MMA is the 128 x 256 x 16 shape (Cout = 64: the weight tile's upper 64 rows are TMA zero fill; an M = 64
// MMA takes the same 128 cycles, so narrower weight operands gain nothing).
// The accumulator arrives transposed -- TMEM lane = output channel, column = pixel -- which makes the epilogue's
// per-channel scale/shift two registers per thread; each thread converts its channel for 32 pixels at a time
// (tcgen05.ld 32x32b.x32) and scatters 16-bit values into the [pixel][channel] 128B-swizzled staging slab (a warp's 32
// lanes = 32 consecutive channels = 64 contiguous bytes of one pixel row: conflict-free), which leaves by TMA store.
//
//   warp 0      TMA producer: pixel tile (256 rows x 128 B; tiled 2D for 1x1 s1, im2col otherwise) + weight tile
//   warp 1      one thread issues tcgen05.mma (4 per stage), commits free the stage / publish the accumulator
//   warps 2..9  epilogue, two groups of four warps (a warp reads only TMEM lanes 32*(warp%4)..+31): group g converts
//               pixel columns [128 g, 128 g + 128) of every tile
// Epilogue: scale/shift (folded BN or bias), optional ReLU, 16-bit output with the optional per-tensor exponent
// (conv_gemm.cuh header).  No residual / coarse / mask operands: the convs that need them are 256 wide.
//
// PATCH (3x3, stride 1, pad 1): a tile is 8 x 32 output pixels; ONE (8+2) x (32+2)-pixel halo patch per 64-channel
// chunk is loaded (4D tiled TMA, zero-filled padding) and the nine taps are nine shifted pixel-operand descriptors
// over it (8-pixel row groups 10 rows apart: SBO = 1280 B), so the activations are read 1.3x instead of 9x; the
// weight tiles stream through their own ring.
#pragma once
#include <type_traits>
#include "conv_gemm.cuh"

namespace tdet {

constexpr int kSwapPix = 256;      // pixels per tile = UMMA N
constexpr int kSwapThreads = 352;  // 11 warps (warp 10: weight producer of the PATCH mode)
constexpr int kSwapPW = 8, kSwapPH = 32;                     // PATCH: output pixels per tile
constexpr int kSwapHaloW = kSwapPW + 2, kSwapHaloH = kSwapPH + 2;
constexpr int kSwapPatchBytes = kSwapHaloW * kSwapHaloH * 128;                  // 43 520
constexpr int kSwapPatchStage = (kSwapPatchBytes + 1023) / 1024 * 1024;         // 44 032
constexpr int kSwapPatchStages = 2;

// !PATCH: STAGES x (pixel tile + weight tile).  PATCH: two halo patches, then STAGES weight tiles.
template <int STAGES, bool PATCH>
struct SwapSmem {
  static constexpr int kPixBytes = kSwapPix * 128;  // 32 KiB
  static constexpr int kWBytes = 128 * 128;         // 16 KiB: 128 weight rows x 64 K elements
  static constexpr int kStageBytes = PATCH ? kWBytes : kPixBytes + kWBytes;
  static constexpr int kWOffset = PATCH ? kSwapPatchStages * kSwapPatchStage : kPixBytes;  // first weight tile
  static constexpr int kOutOffset = (PATCH ? kSwapPatchStages * kSwapPatchStage : 0) + STAGES * kStageBytes;
  static constexpr int kOutBytes = 2 * 2 * kSlabBytes;  // two groups x two 64-channel slabs of 128 pixels
  static constexpr int kBarOffset = kOutOffset + kOutBytes;
  static constexpr int kNumBars = 2 * STAGES + 4 + 2 * kSwapPatchStages;
  static constexpr int kTmemPtrOffset = kBarOffset + kNumBars * 8;
  static constexpr int kDynamic = kTmemPtrOffset + 16;
  static_assert(kDynamic <= 232448, "exceeds the 227 KiB shared memory limit");
};

template <int STAGES, bool PATCH = false>
__global__ void __launch_bounds__(kSwapThreads, 1)
conv_swap_kernel(const __grid_constant__ ConvGemmParams p) {
  using L = SwapSmem<STAGES, PATCH>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t smem_out = base + L::kOutOffset;
  const uint32_t bar0 = base + L::kBarOffset;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  auto afull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + s); };
  auto aempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + kSwapPatchStages + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::kTmemPtrOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
    tma_prefetch_desc(&p.tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);  // one arrive per epilogue warp
    }
    for (int s = 0; s < kSwapPatchStages; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  grid_dependency_wait();

  const int num_tiles = p.num_m_tiles;  // 256-pixel tiles
  const int num_kb = p.kh * p.kw * p.k_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    if (PATCH) {
      int as = 0;
      uint32_t aphase = 0;
      for (int it = blockIdx.x; it < num_tiles; it += gridDim.x) {
  const int tile = p.reverse ? num_tiles - 1 - it : it;  // TDET_FLAG_REVERSE
        const int tw = tile % p.tiles_w;
        const int t = tile / p.tiles_w;
        const int th = t % p.tiles_h;
        const int img = t / p.tiles_h;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(aempty_bar(as), aphase ^ 1u);
          if (lane == 0) {
            mbar_arrive_expect_tx(afull_bar(as), kSwapPatchBytes);
            tma_load_4d(base + as * kSwapPatchStage, &p.tmap_a, afull_bar(as), kc * kBK, tw * kSwapPW - 1,
                        th * kSwapPH - 1, img);
          }
          __syncwarp();
          if (++as == kSwapPatchStages) { as = 0; aphase ^= 1u; }
        }
      }
    } else
    for (int it = blockIdx.x; it < num_tiles; it += gridDim.x) {
  const int tile = p.reverse ? num_tiles - 1 - it : it;  // TDET_FLAG_REVERSE
      int cw = 0, ch = 0, cn = 0;
      if (p.a_mode == A_IM2COL) {
        const int m0 = tile * kSwapPix;
        const int q0 = m0 % p.Wo;
        const int t = m0 / p.Wo;
        cn = t / p.Ho;
        cw = q0 * p.stride - p.pad;
        ch = (t % p.Ho) * p.stride - p.pad;
      }
      for (int r = 0; r < p.kh; ++r) {
        for (int s = 0; s < p.kw; ++s) {
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (lane == 0) {
              const uint32_t fb = full_bar(stage);
              mbar_arrive_expect_tx(fb, L::kStageBytes);
              const uint32_t dst = base + stage * L::kStageBytes;
              if (p.a_mode == A_TILED)
                tma_load_2d(dst, &p.tmap_a, fb, kc * kBK, tile * kSwapPix);
              else
                tma_load_im2col_4d(dst, &p.tmap_a, fb, kc * kBK, cw, ch, cn, static_cast<uint16_t>(s * p.dil),
                                   static_cast<uint16_t>(r * p.dil));
              // rows >= Cout of the 128-row box lie outside the weight matrix: zero fill
              tma_load_2d(dst + L::kPixBytes, &p.tmap_b, fb, (r * p.kw + s) * p.b_tap_stride + kc * kBK, 0);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ TMA producer (weights, PATCH mode): its own
    // warp, so that the next tile's halo patch is requested a whole tile ahead instead of behind nine weight tiles
    if (PATCH) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < num_tiles; it += gridDim.x) {
  const int tile = p.reverse ? num_tiles - 1 - it : it;  // TDET_FLAG_REVERSE
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (lane == 0) {
              mbar_arrive_expect_tx(full_bar(stage), L::kWBytes);
              tma_load_2d(base + L::kWOffset + stage * L::kStageBytes, &p.tmap_b, full_bar(stage),
                          tap * p.b_tap_stride + kc * kBK, 0);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      // A operand = weights (M = 128 rows), B operand = pixels (N = 256 rows); both K-major SWIZZLE_128B tiles
      const uint32_t idesc = make_idesc_f16kind(128, kSwapPix, p.b_fp16 ? kFmtF16 : kFmtBF16,
                                                p.ab_fp16 ? kFmtF16 : kFmtBF16);
      const uint64_t dpix0 = make_smem_desc_sw128(base);
      const uint64_t dw0 = make_smem_desc_sw128(base + L::kPixBytes);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (PATCH) {
        const uint64_t dpatch0 = make_smem_desc_sw128_sbo(base, kSwapHaloW * 128);
        const uint64_t dwp0 = make_smem_desc_sw128(base + L::kWOffset);
        int as = 0;
        uint32_t aphase = 0;
        for (int it = blockIdx.x; it < num_tiles; it += gridDim.x) {
  const int tile = p.reverse ? num_tiles - 1 - it : it;  // TDET_FLAG_REVERSE
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kSwapPix);
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(afull_bar(as), aphase);
            tc_fence_after();
            const uint64_t dpatch = dpatch0 + static_cast<uint32_t>(as * (kSwapPatchStage >> 4));
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint64_t dpx = dpatch + static_cast<uint32_t>(((tap / 3) * kSwapHaloW + tap % 3) * 8);
              const uint64_t dw = dwp0 + static_cast<uint32_t>(stage * (L::kStageBytes >> 4));
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k)
                umma_bf16_ss(d_tmem, dw + 2u * k, dpx + 2u * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              umma_commit(empty_bar(stage));
              if (tap == 8) {
                umma_commit(aempty_bar(as));
                if (kc == p.k_chunks - 1) umma_commit(tfull_bar(acc));
              }
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (++as == kSwapPatchStages) { as = 0; aphase ^= 1u; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      } else
      for (int it = blockIdx.x; it < num_tiles; it += gridDim.x) {
  const int tile = p.reverse ? num_tiles - 1 - it : it;  // TDET_FLAG_REVERSE
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kSwapPix);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t off = static_cast<uint32_t>(stage * (L::kStageBytes >> 4));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_bf16_ss(d_tmem, dw0 + (off + 2u * k), dpix0 + (off + 2u * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int group = (warp - 2) >> 2;      // pixel columns [128 group, 128 group + 128) of every tile
    const int gtid = (threadIdx.x - 64) & (kEpiGroupThreads - 1);
    const uint32_t gbar = 1u + group;
    const bool issuer = gtid == 0;
    const int c = quad * 32 + lane;         // output channel of this thread
    const bool active = quad * 32 < p.N;    // warp-uniform: channels [32 quad, 32 quad + 32) exist
    const bool out_fp16 = p.out_fp16 != 0;
    const int slabs = p.N / 64;             // 1 or 2 staging slabs per group
    const uint32_t stage_g = smem_out + group * 2 * kSlabBytes;

    const int e_in = p.in_meta ? p.in_meta->e : 0;
    int e_out = 0;
    if (p.out_scaled) {
      float bound = p.bound_consts[1];
      const float a_in = p.in_meta ? __uint_as_float(p.in_meta->amax_bits) : 0.0f;
      bound += p.bound_consts[0] * a_in;
      if (bound > 0.0f && bound < 3.0e38f) e_out = ilogbf(bound) - 14;
      e_out = max(-100, min(100, e_out));
      if (blockIdx.x == 0 && threadIdx.x == 64) p.out_meta->e = e_out;
    }
    const float sc = active ? (p.scale ? __ldg(p.scale + c) : 1.0f) * ldexpf(1.0f, e_in - e_out) : 0.0f;
    const float sh = active ? (p.shift ? __ldg(p.shift + c) : 0.0f) * ldexpf(1.0f, -e_out) : 0.0f;
    // staging address of (pixel row 0, this channel): slab c/64, 16-byte chunk (c%64)/8 (XOR-swizzled per row)
    const uint32_t st_base = stage_g + (c >> 6) * kSlabBytes + (c & 7) * 2;
    const int chunk16 = (c & 63) >> 3;
    float amax_local = 0.0f;

    int seq = 0;
    for (int it = blockIdx.x; it < num_tiles; it += gridDim.x, ++seq) {
  const int tile = p.reverse ? num_tiles - 1 - it : it;  // TDET_FLAG_REVERSE
      const int acc = seq & 1;
      const uint32_t acc_phase = static_cast<uint32_t>(seq >> 1) & 1u;
      const int m0 = tile * kSwapPix + group * 128;  // first pixel of this group's half
      // PATCH: this group's half is rows [16 group, 16 group + 16) x 8 columns of the 8 x 32 spatial tile
      int ptw = 0, pth = 0, pimg = 0;
      if (PATCH) {
        ptw = tile % p.tiles_w;
        const int t = tile / p.tiles_w;
        pth = t % p.tiles_h;
        pimg = t / p.tiles_h;
      }
      const int h0 = pth * kSwapPH + group * 16, w0 = ptw * kSwapPW;
      // the staging slabs were last read by the TMA store of the previous tile
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      named_bar_sync(gbar, kEpiGroupThreads);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (active) {
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                static_cast<uint32_t>(acc * kSwapPix + group * 128);
        // every pixel row of this group's half lies inside the tensor (all but edge tiles): no per-element test
        const bool all_ok = PATCH ? (h0 + 16 <= p.Ho && w0 + 8 <= p.Wo) : (m0 + 128 <= p.M);
        // (format / activation / bounds fixed at compile time in the common variants: with run-time flags the compiler
        // computes both sides of every per-element branch of the unrolled loop)
        auto convert = [&](auto kOut16, auto kAllOk, auto kFast) {
#pragma unroll 1
          for (int ck = 0; ck < 4; ++ck) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_addr + ck * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int pr = ck * 32 + i;  // pixel row inside the group's 128
              float y = fmaf(__uint_as_float(v[i]), sc, sh);
              if (decltype(kFast)::value) y = fmaxf(y, 0.0f);
              else if (p.relu) y = p.relu == 2 ? fminf(fmaxf(y, 0.0f), 6.0f) : fmaxf(y, 0.0f);
              const bool ok = decltype(kAllOk)::value ||
                              (PATCH ? (h0 + (pr >> 3) < p.Ho && w0 + (pr & 7) < p.Wo) : (m0 + pr < p.M));
              if (ok) amax_local = fmaxf(amax_local, fabsf(y));
              uint16_t h;
              if (decltype(kFast)::value ? decltype(kOut16)::value : out_fp16) {
                const __half hh = __float2half_rn(y);
                h = *reinterpret_cast<const uint16_t*>(&hh);
              } else {
                const __nv_bfloat16 bb = __float2bfloat16_rn(y);
                h = *reinterpret_cast<const uint16_t*>(&bb);
              }
              const uint32_t a = st_base + pr * 128 + ((chunk16 ^ (pr & 7)) << 4);
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(h) : "memory");
            }
          }
        };
        using T = std::true_type;
        using F = std::false_type;
        // (not for the halo-patch 3x3 variant: it is bound by its MMA-issuing thread, and a denser epilogue instruction
        // stream on the same scheduler measured 10 % slower -- 87 -> 97 us on layer2's conv2)
        if (!PATCH && p.relu == 1 && all_ok && p.epi_fast) {
          if (out_fp16) convert(T{}, T{}, T{}); else convert(F{}, T{}, T{});
        } else {
          convert(F{}, F{}, F{});
        }
      }
      // all TMEM reads of this warp are complete: hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      fence_proxy_async_smem();
      named_bar_sync(gbar, kEpiGroupThreads);
      if (issuer && (PATCH ? h0 < p.Ho : m0 < p.M)) {
        for (int s = 0; s < slabs; ++s) {
          if (PATCH)
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)), "r"(stage_g + s * kSlabBytes), "r"(s * 64),
                         "r"(w0), "r"(h0), "r"(pimg)
                         : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)), "r"(stage_g + s * kSlabBytes), "r"(s * 64),
                         "r"(m0)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (p.out_meta) {
      float a = amax_local;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
      if (lane == 0) atomicMax(&p.out_meta->amax_bits, __float_as_uint(ldexpf(a, e_out)));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tdet
