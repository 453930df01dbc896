// Implicit-GEMM convolution for sm_100a: TMA (tiled / im2col) -> shared memory ring ->
// tcgen05.mma (16-bit x 16-bit -> fp32 in TMEM) -> fused epilogue (BN scale/shift or bias, residual
// add, FPN nearest-x2 upsample-add, ReLU) -> shared-memory staging -> TMA store, NHWC 16-bit.
//
// GEMM view: D[M = N_img*Ho*Wo, Cout] = A[M, K = kh*kw*Cin] * W[Cout, K]^T.
// A rows are output pixels (NHWC), gathered by the TMA unit:
//   A_TILED  : 1x1 stride-1 conv -> A is the activation matrix itself (2D tiled tensor map)
//   A_IM2COL : kxk / strided conv -> 4D im2col tensor map, one load per (filter tap, 64-ch chunk)
//   A_STEM   : 7x7/2 stem on the padded NHWC4 staging -> 5D tiled map with overlapping windows
//   A_STEM2  : the stem from linear staged rows (un-swizzled overlapping descriptors); POOL fuses the max-pool
//   A_PATCH  : 3x3 stride-1 conv -> one halo patch per tile, the nine taps are shifted descriptors over it
//   A_SPATIAL: 1x1 conv over 8x16-pixel tiles with the coarse FPN level staged by TMA (laterals)
// W is pre-packed [Cout][kh][kw][Cin], i.e. K-major rows, loaded by a 2D tiled map.
// Variants: PAIR (clusters of two CTAs, cta_group::2 MMAs of M = 256), SPLIT (split-precision fp32-I/O mode),
// MASKED (ReLU-backward mask operand: dgrad), POOL (stem + max-pool); narrow outputs run in conv_swap.cuh.
//
// Persistent, warp-specialised CTA (352 threads), every global<->shared transfer is asynchronous:
//   warp 0      TMA producer for A/B     (full/empty mbarrier ring, STAGES deep)
//   warp 1      tcgen05.mma issuer       (lane 0 issues; accumulators double-buffered in TMEM)
//   warp 2      TMA producer for the residual operand (64-column slabs, RES_SLABS-deep ring) so the
//                                         residual of tile i+1 streams in while tile i is finished
//   warps 3..10 epilogue                 (tcgen05.ld 32x32b -> registers -> fp32 math -> 128B-swizzled
//                                         staging slab -> TMA store), overlapped with the next tile's
//                                         main loop through tmem_full/empty.  A warp may only read
//                                         TMEM lanes 32*(warp%4)..+31, so the 8 warps form two groups
//                                         of four.  Group g converts every second 64-column slab of a
//                                         tile (every second tile when a tile is one slab wide) on its
//                                         own named barrier, staging slabs and TMA-store queue -- each
//                                         residual-ring slot then has a single consumer, which keeps
//                                         the mbarrier parity protocol sound -- so the two groups run
//                                         de-synchronised and
//                                         hide each other's latency chain (tcgen05.ld -> LDS -> convert ->
//                                         STS -> fence -> barrier -> store), which -- not issue rate and not
//                                         DRAM -- bounded the memory-bound 1x1 convs (ncu: 67 % of cycles
//                                         without an eligible warp at 55 % DRAM utilisation).
//
// Numerics: every stored activation tensor may carry a per-tensor power-of-two exponent
// (TensorMeta::e, value = stored * 2^e) and its true |max| (TensorMeta::amax_bits).  A kernel picks
// its output exponent from a rigorous bound |out| <= G*amax(in) + max|shift| + amax(res) +
// amax(coarse) (G = max_c |scale_c| * ||w_c||_1), which keeps fp16 storage (11-bit significand)
// overflow-free for any input without calibration.  With all metadata pointers null the exponents
// are 0 and the tensors are plain bf16/fp16.
#pragma once
#include <type_traits>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx_sm100.cuh"

namespace tdet {

constexpr int kBM = 128;         // UMMA M (cta_group::1)
constexpr int kBK = 64;          // 16-bit elements per 128-byte swizzle row
constexpr int kUmmaK = 16;       // K per tcgen05.mma for 16-bit operands
constexpr int kGemmThreads = 384;       // 12 warps; warp 11 is the B producer of the patch mode
constexpr int kGemmThreadsNoPatch = 352; // 11 warps otherwise: 186 instead of 170 registers per thread
constexpr int kGemmThreadsNG4 = 608;     // 3 role warps + 16 epilogue warps (four groups): 104 registers per thread
constexpr int kEpiThreads = 256;         // 8 epilogue warps: two warps per TMEM lane quadrant
constexpr int kABytes = kBM * kBK * 2;   // 16 KiB per stage
constexpr int kSlabBytes = kBM * 128;    // 128 rows x 64 columns x 2 B (one swizzle-128B slab)
constexpr int kEpiGroupThreads = 128;    // one epilogue group = 4 warps = all 128 TMEM lanes

enum AMode : int { A_TILED = 0, A_IM2COL = 1, A_STEM = 2, A_STEM2 = 3, A_PATCH = 4, A_SPATIAL = 5 };

// A_SPATIAL (1x1, stride 1, with a coarse operand and nothing else in the ring: the FPN laterals): an M tile is
// 8 x 16 output pixels (one 4D tiled TMA load per 64-channel chunk, a standard K-major tile), so the coarse
// pixels the tile's epilogue adds form ONE 4 x 8-pixel box that TMA stages in shared memory -- a quarter-size
// slot of the residual ring -- instead of 16-byte global loads whose latency the epilogue cannot hide.
constexpr int kCoarseSlotBytes = 4 * 8 * 128;  // 4096: 32 coarse pixels x 64 channels
constexpr int kRingSplit = 4;                  // coarse slots per 16 KiB residual slab

// A_PATCH (3x3, stride 1, pad 1, dil 1): an M tile is 8 x 16 output pixels; per 64-channel chunk ONE
// tiled TMA load deposits the (8+2) x (16+2) pixel halo patch (180 rows of 128 B, SWIZZLE_128B) and the
// nine filter taps are nine *shifted views* of it: the A descriptor of tap (r,s) starts at patch row
// r*10+s with 1280 B between 8-row groups.  tcgen05 applies the 128B swizzle as a function of the
// shared-memory address, so any 128-byte-row start and any group stride address the TMA-written
// pattern consistently (tools/probe_umma_shift.cu).  A traffic drops from 9x to 1.41x the input.
constexpr int kPatchBW = 8, kPatchBH = 16;
constexpr int kPatchPW = kPatchBW + 2, kPatchPH = kPatchBH + 2;
constexpr int kPatchBytes = kPatchPW * kPatchPH * 128;          // 23040
constexpr int kPatchStageBytes = (kPatchBytes + 1023) / 1024 * 1024;  // 23552
constexpr int kPatchStages = 3;

struct TensorMeta {
  int e;                  // stored value * 2^e = true value
  unsigned amax_bits;     // float bits of the true |max| (atomicMax on the bit pattern; values >= 0)
};

struct ConvGemmParams {
  CUtensorMap tmap_a;
  CUtensorMap tmap_b;
  CUtensorMap tmap_out;   // 2D (Cout, M) box (64,128); A_STEM: 4D (64, Wo, Ho, N) box (64,bw,bh,1)
  CUtensorMap tmap_res;   // 2D (Cout, M) box (64,128) over the residual tensor (if any)
  CUtensorMap tmap_mask;  // same geometry over the ReLU-mask tensor (if mask_tma)
  CUtensorMap tmap_coarse; // A_SPATIAL: 4D (C, Wc, Hc, N) box (64, 4, 8, 1) over the coarse operand
  int coarse_tma;
  int M;            // valid output rows (pixels); for A_STEM rows are masked per pixel instead
  int N;            // Cout
  int num_m_tiles;
  int num_n_tiles;
  int k_chunks;     // Cin / 64 (A_STEM: 1)
  int grouped;      // grouped conv as a 64-channel band: n-tile j (BN = 64) contracts only channel chunk j
  // Split precision (fp32-I/O mode): every tensor is a bf16 pair value = hi + lo stored as 2*C channels
  // [hi | lo]; the GEMM runs the three significant products hi*hi + lo*hi + hi*lo as a 3x longer K loop
  // (pure producer-side indexing), the epilogue splits its fp32 result into hi/lo again.
  int split;
  int b_tap_stride; // weight columns per filter tap (cin, or 2*cin in split mode; stem: 64)
  int b_lo_off;     // split mode: column offset of the lo weights inside a tap (cin; stem: 448)
  int a_lo_img;     // split mode, stem: image-index offset of the lo plane of the staged batch
  int kh, kw, dil;
  int cin;          // B column offset of tap (r,s) is (r*kw + s)*cin
  int a_mode;
  int Ho, Wo;       // output spatial size
  int stride, pad;
  int tiles_w, tiles_h;  // A_STEM: spatial tiles per image
  int tile_bw, tile_bh;  // A_STEM: tile shape in output pixels (tile_bw * tile_bh == 128)
  int a_stage_bytes;     // bytes one A load deposits per stage (kABytes except A_STEM2)
  int num_kb_b;          // number of 64-wide k-blocks of the weight matrix (resident-B kernels)
  int stem_row_bytes;    // A_STEM2: shared-memory pitch of one staged image row segment
  int pool_h, pool_w;    // POOL kernels: size of the 3x3/2 max-pooled output the stem writes instead of its own
  void* pool_out;        //   [n][pool_h][pool_w][64], 16-bit
  int Hc, Wc;       // coarse level size (upsample-add)
  int relu;
  int has_res;
  int ab_fp16;      // operand format of A: 1 = fp16, 0 = bf16
  int b_fp16;       // operand format of B (weights); the instruction descriptor has one field per operand, but a
                    // mixed F16 x BF16 tcgen05.mma traps on B200 ("illegal instruction", measured): keep them equal
  int out_fp16, res_fp16, coarse_fp16;  // storage formats
  int out_scaled;   // choose a power-of-two output exponent from the bound (else exponent 0)
  const float* scale;
  const float* shift;
  const void* coarse;
  int coarse_parity;              // 1: add coarse[p/2][q/2] only where p and q are even (dgrad of a
                                  //    stride-2 1x1 shortcut); 0: nearest-x2 upsample-add (FPN)
  const void* mask_src;           // nullable 16-bit [M][N] tensor: output is zeroed where it is <= 0
                                  //    (ReLU backward against the stored forward activation)
  int mask_tma;                   // 1: the mask streams through the residual ring by TMA (tmap_mask), one
                                  //    slab after the residual's; 0: read with global loads in the epilogue
  const TensorMeta* in_meta;      // nullable
  const TensorMeta* res_meta;     // nullable
  const TensorMeta* coarse_meta;  // nullable
  TensorMeta* out_meta;           // nullable; when set the true |max| of the output is recorded
  const float* bound_consts;      // {G, max|shift|}, required iff out_scaled
  // DUAL (1x1 / stride-1 convs): a SECOND A source -- a stage's projection shortcut, 1x1 with stride2 over x2 --
  // is contracted after the main K loop into the SAME accumulator: the weight matrix is the K-concatenation
  // [W (cin) | W2 (cin2)] with both BatchNorm scales folded into its rows, so the shortcut tensor never exists in
  // memory.  Both sources are plain tensors of one format (exponent 0).
  CUtensorMap tmap_a2;            // 2D tiled (stride2 == 1) or 4D im2col (stride2 > 1) view of x2
  // warp_stores: every epilogue warp stages and TMA-stores its OWN 32 rows of a slab (tmap_out then has 32-row boxes)
  // and tracks its own bulk groups, so a slab costs two warp-level syncs instead of two 128-thread named barriers and
  // the eight warps run fully de-synchronised (0: one 128-row store per slab, issued by the group's first thread)
  int warp_stores;
  int res_prefetch;               // residual / mask slabs are prefetched into L2 this many tiles ahead (0 = off)
  int dual;
  int k_chunks2;
  int stride2;
  const TensorMeta* in2_meta;     // nullable
  int epi_fast;                   // 1: compile-time epilogue variants where one matches (TDET_EPI_FAST, default 1)
  // reverse: the persistent tile loop walks the tiles from the LAST to the first (TDET_FLAG_REVERSE).  Results are
  // identical; consecutive launches of a plan alternate direction so that a kernel starts on the part of its input
  // the previous kernel wrote last -- the part that is still in the 126 MB L2 when the tensor is larger than L2.
  int reverse;
};

// BRES_KB > 0: the whole weight panel of the CTA's n-tile (up to BRES_KB k-blocks; one n-tile, or a grid that is a
// multiple of the number of n-tiles so that a CTA never changes its n-tile) is loaded once per CTA and stays resident
// in shared memory; only A tiles stream through the ring.
// PATCH: A_PATCH pipeline (A ring of kPatchStages halo patches; STAGES then counts B tiles).
// OSLABS: staging slabs per epilogue group (1: a group's stores serialise with its next slab; 2: double
// buffered).
// PAIR: CTA pair (cta_group::2): each CTA stages only its half of the weight tile's N rows.
// NG: epilogue groups (2, or 4 = one group per 64-column slab of a 256-wide tile).
template <int BN, int STAGES, int RES_SLABS, int BRES_KB, bool PATCH, int OSLABS, bool PAIR = false, int NG = 2>
struct GemmSmem {
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kBK * 2;
  static constexpr int kBSlots = BRES_KB > 0 ? BRES_KB : STAGES;
  static constexpr int kAStages = PATCH ? kPatchStages : STAGES;
  static constexpr int kAStageBytes = PATCH ? kPatchStageBytes : kABytes;
  static constexpr int kBOffset = kAStages * kAStageBytes;
  static constexpr int kResOffset = kBOffset + kBSlots * kBBytes;
  static constexpr int kOutOffset = kResOffset + RES_SLABS * kSlabBytes;
  static constexpr int kBarOffset = kOutOffset + NG * OSLABS * kSlabBytes;
  // the coarse-staging variants (BN 256, ring of 1 or 2 slabs) cut each slab into quarter-size slots
  static constexpr bool kCoarseRing = BN == 256 && !PATCH && !PAIR && OSLABS == 1 && (RES_SLABS == 1 || RES_SLABS == 2);
  static constexpr int kRingBars = kCoarseRing ? kRingSplit * RES_SLABS : (RES_SLABS > 0 ? RES_SLABS : 1);
  static constexpr int kNumBars = 2 * STAGES + 5 + 2 * kRingBars + 2 * kPatchStages;
  static constexpr int kTmemPtrOffset = kBarOffset + kNumBars * 8;  // + 8: ring progress of the two groups
  static constexpr int kParamOffset = (kTmemPtrOffset + 16 + 15) / 16 * 16;
  // scale/shift per epilogue group: a group only touches the columns of its own slabs (half of BN),
  // except for one-slab tiles where the groups alternate tiles
  static constexpr int kGroupCols = BN == 64 ? 64 : BN / NG;
  static constexpr int kDynamic = kParamOffset + NG * 2 * kGroupCols * 4;
  static_assert(kDynamic <= 232448, "exceeds the 227 KiB shared memory limit");
};

// Compile-time description of the epilogue's conversion step (conv_gemm_kernel): mode 0 = every flag at run time (the
// general code), 1 = no residual / coarse operand, 2 = fp16 residual, 3 = bf16 coarse operand (FPN laterals).
template <int MODE, bool OUT16, bool RELU>
struct EpiFast {
  static constexpr int mode = MODE;
  static constexpr bool out16 = OUT16;
  static constexpr bool relu = RELU;
};

__device__ __forceinline__ void unpack16x2(uint32_t w, bool fp16, float& lo, float& hi) {
  if (fp16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x;
    hi = f.y;
  } else {
    lo = bf16_lo(w);
    hi = bf16_hi(w);
  }
}
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, bool fp16) {
  if (fp16) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  return pack_bf16x2(lo, hi);
}

// MASKED: the kernel carries the ReLU-backward mask path (dgrad); forward instantiations compile it out.
// SPLIT: split-precision epilogue (hi/lo outputs, hi/lo residual and coarse operands); needs OSLABS == 2.
// PAIR: launched as clusters of two CTAs (one TPC); the pair computes two vertically adjacent 128-row tiles of one
// n-tile with M = 256 cta_group::2 MMAs issued by the leader.  Each CTA loads its own A tile and HALF of the weight
// tile (the MMA reads both halves across the pair), so the L2 -> SM operand traffic per MMA cycle drops by a third
// -- the streamed-weight kernels are bound by exactly that traffic -- and each weight tile is fetched once per 256
// output rows.  Accumulators, residual ring and epilogue stay per-CTA.
// POOL (stem, A_STEM2 only): the epilogue max-pools (3x3, stride 2, pad 1: resnet.py:218) before anything reaches
// HBM.  A CTA walks DOWN consecutive conv rows of one 128-pixel column strip; each epilogue group owns 32 of the 64
// channels of EVERY row (pooling is per channel, so the groups never talk), keeps the last three post-ReLU rows
// in shared memory and, after every odd row 2p+1, writes pooled row p = max over rows 2p-1..2p+1 x columns
// 2q-1..2q+1.  Strips start at conv column 112 j - 8 and produce pooled columns 56 j .. 56 j + 55 (12.5 % of the
// MMA rows overlap, as many tiles as before); out-of-image positions hold 0, which equals the -inf padding of the
// reference because every window has a valid element and post-ReLU values are >= 0.  This removes the write and
// the re-read of the 64 x H/2 x W/2 stem output (1.1 GB per batch of 16 at 800x1344) and the pooling launch.
constexpr int kPoolStep = 56;  // pooled columns per strip

#ifndef TDET_MMA_SINGLE_THREAD
#define TDET_MMA_SINGLE_THREAD 0
#endif

// NG = 4 (256-wide tiles of the short-K 1x1 convs: conv3 + residual, dual-source conv3, FPN laterals): FOUR epilogue
// groups, one per 64-column slab, and 16-column conversion steps so that 19 warps fit the register file.  Those
// kernels are bound by their epilogue -- ~12 000 warp instructions per 128 x 256 tile at 41 % issue utilisation with
// eight warps (ncu, layer3 conv3: 8 600 cycles per tile against ~1 000 of MMA) -- not by the tensor pipe or HBM.
template <int W>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[W]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_32x32b_x32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld_32x32b_x16(taddr, v); }

template <int BN, int STAGES, int RES_SLABS, int BRES_KB, bool PATCH, int OSLABS, bool MASKED, bool SPLIT = false,
          bool PAIR = false, bool POOL = false, int NG = 2>
__global__ void __launch_bounds__(NG == 4 ? kGemmThreadsNG4 : (PATCH ? kGemmThreads : kGemmThreadsNoPatch), 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  static_assert(NG == 2 || (NG == 4 && BN == 256 && !PATCH && !SPLIT && !POOL && !PAIR),
                "four epilogue groups: 256-wide tiles, tiled / spatial A, one CTA");
  constexpr int kEW = NG == 4 ? 16 : 32;   // accumulator columns converted per step
  constexpr int kEJ = kEW / 8;             // 16-byte chunks per step
  constexpr int kParts = 64 / kEW;         // steps per 64-column slab
  static_assert(!POOL || (BN == 64 && BRES_KB > 0 && !PATCH && !SPLIT && !PAIR && !MASKED && RES_SLABS == 0),
                "POOL: the stem kernel only");
  using L = GemmSmem<BN, STAGES, RES_SLABS, BRES_KB, PATCH, OSLABS, PAIR, NG>;
  static_assert(!PAIR || (BN >= 128 && BRES_KB == 0 && !SPLIT), "CTA pairs: streamed weight tiles, 128/256 wide");
  // Split precision keeps TWO accumulators per tile: hi*hi in one, the small cross terms lo*hi + hi*lo in the
  // other, summed in fp32 by the epilogue.  tcgen05's fp32 accumulation is not exact -- measured ~0.17 ulp of
  // systematic loss per MMA step, which over the 3x longer K loop and ~50 layers was the dominant error of
  // the fp32-I/O mode (1.0e-4 at P2 vs 1.7e-5 for the scheme with exact accumulation); keeping the main chain
  // a third as long and the cross terms (2^-8 of the magnitude) apart cuts it ~3x.
  constexpr int kAccStride = SPLIT ? 2 * BN : BN;            // TMEM columns per accumulator buffer
  constexpr int kAccBufs = (SPLIT && BN == 256) ? 1 : 2;      // 2 x 256 columns leave no room to double-buffer
  constexpr uint32_t kAccCols = kAccStride * kAccBufs;
  constexpr uint32_t kTmemCols = (kAccCols <= 32) ? 32 : (kAccCols <= 64) ? 64 : (kAccCols <= 128) ? 128
                                 : (kAccCols <= 256) ? 256 : 512;
  constexpr int kSlabsPerTile = BN / 64;
  constexpr int kRS = RES_SLABS > 0 ? RES_SLABS : 1;


  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0u) __trap();  // SWIZZLE_128B operands need 1024-byte alignment

  const uint32_t smem_a = base;
  const uint32_t smem_b = base + L::kBOffset;
  const uint32_t smem_res = base + L::kResOffset;
  const uint32_t smem_out = base + L::kOutOffset;
  const uint32_t bar0 = base + L::kBarOffset;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  auto rfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + s); };
  constexpr int kRB = L::kRingBars;
  auto rempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + kRB + s); };
  const uint32_t bres_bar = bar0 + 8u * (2 * STAGES + 4 + 2 * kRB);
  auto afull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 5 + 2 * kRB + s); };
  auto aempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 5 + 2 * kRB + kPatchStages + s); };
  // ring geometry: 16 KiB residual / mask slabs, or quarter-size coarse boxes (A_SPATIAL)
  const bool coarse_tma = L::kCoarseRing && !MASKED && !SPLIT && p.coarse_tma != 0;
  const int rdepth = coarse_tma ? kRingSplit * kRS : kRS;
  const uint32_t rslot = coarse_tma ? kCoarseSlotBytes : kSlabBytes;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::kTmemPtrOffset);
  // highest residual-ring index the producer has issued so far (see the epilogue)
  volatile int* ring_issued = reinterpret_cast<volatile int*>(smem + L::kTmemPtrOffset + 8);
  float* s_params = reinterpret_cast<float*>(smem + L::kParamOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
    if (!POOL) tma_prefetch_desc(&p.tmap_out);
    if (p.has_res) tma_prefetch_desc(&p.tmap_res);
    if (p.mask_tma) tma_prefetch_desc(&p.tmap_mask);
    if (p.coarse_tma) tma_prefetch_desc(&p.tmap_coarse);
    if (p.dual) tma_prefetch_desc(&p.tmap_a2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      // one arrive per warp draining buffer a (the leader's barrier collects both CTAs of a pair)
      mbar_init(tempty_bar(a), ((kSlabsPerTile == 1 && !POOL) ? 4 : 4 * NG) * (PAIR ? 2 : 1));
    }
    for (int s = 0; s < kRB; ++s) {
      mbar_init(rfull_bar(s), 1);
      mbar_init(rempty_bar(s), (p.warp_stores && !POOL) ? 4 : 1);  // per-warp consumers: one arrive per warp of the group
    }
    mbar_init(bres_bar, 1);
    *ring_issued = -1;
    for (int s = 0; s < kPatchStages; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    fence_mbar_init();
  }
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  if (warp == 3) {
    if (PAIR) {
      tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Launched with programmatic stream serialisation: everything above (barrier init, TMEM allocation,
  // descriptor prefetch) overlaps the tail of the previous kernel; its outputs are read only below.
  grid_dependency_wait();

  // tile schedule: CTA (or pair) t takes tiles t, t + step, ...; a pair's tile is two consecutive m-tiles, the
  // rank picks one (the last pair of an odd count re-loads the last valid tile and stores nothing)
  // POOL: this CTA's contiguous range [pool_g0, pool_g1) of (image, strip, pooled row) units; a run of pooled rows
  // [pa, pb) of one strip takes the conv rows 2 pa - 1 .. 2 pb - 1 in order (one tile each)
  int pool_g0 = 0, pool_g1 = 0, pool_tiles = 0;
  if (POOL) {
    const long long total = static_cast<long long>(p.num_m_tiles);  // = images * strips * pool_h
    pool_g0 = static_cast<int>(total * blockIdx.x / gridDim.x);
    pool_g1 = static_cast<int>(total * (blockIdx.x + 1) / gridDim.x);
    for (int g = pool_g0; g < pool_g1;) {
      const int pa = g % p.pool_h;
      const int len = min(p.pool_h - pa, pool_g1 - g);
      pool_tiles += 2 * len + 1;
      g += len;
    }
  }
  const int tile0 = POOL ? 0 : PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = POOL ? 1 : PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int num_tiles = POOL ? pool_tiles : (PAIR ? (p.num_m_tiles + 1) >> 1 : p.num_m_tiles) * p.num_n_tiles;
  const int rev_last = (!POOL && p.reverse) ? num_tiles - 1 : -1;
  auto phys = [&](int tile) { return rev_last >= 0 ? rev_last - tile : tile; };   // loop index -> tile
  auto m_tile_of = [&](int tile) {
    const int mt = phys(tile) / p.num_n_tiles;
    return PAIR ? 2 * mt + static_cast<int>(cta_rank) : mt;
  };
  // grouped convs (group width divides 64): output channels [64j, 64j+64) only see input channels of the
  // same range, so each 64-wide n-tile runs ONE channel chunk per filter tap over the densely packed,
  // block-diagonal weight matrix
  const int kcn = p.grouped ? 1 : p.k_chunks;
  const int nv = (SPLIT && p.split) ? 3 : 1;  // virtual K passes: hi*hi, lo*hi, hi*lo
  const int num_kb = p.kh * p.kw * kcn * nv + (p.dual ? p.k_chunks2 : 0);  // DUAL: + the second source's k-blocks

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (A, B)
    int stage = 0;
    uint32_t phase = 0;
    if (BRES_KB > 0 && lane == 0) {
      // (several n-tiles: the host sizes the grid to a multiple of their number, so that every tile of this CTA
      // belongs to n-tile tile0 % num_n_tiles and its weight panel can stay resident)
      const int n_res = (phys(tile0) % p.num_n_tiles) * BN;
      mbar_arrive_expect_tx(bres_bar, static_cast<uint32_t>(p.num_kb_b) * L::kBBytes);
      for (int kb = 0; kb < p.num_kb_b; ++kb)
        tma_load_2d(smem_b + kb * L::kBBytes, &p.tmap_b, bres_bar, kb * kBK, n_res);
    }
    const uint32_t stage_tx = static_cast<uint32_t>(p.a_stage_bytes) + (BRES_KB > 0 ? 0u : L::kBBytes);
    if (PATCH) {
      // one halo patch per (tile, 64-channel chunk); OOB pixels (the conv padding) are zero-filled
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m_tile = PAIR ? min(m_tile_of(tile), p.num_m_tiles - 1) : m_tile_of(tile);
        const int tw = m_tile % p.tiles_w;
        const int t = m_tile / p.tiles_w;
        const int th = t % p.tiles_h;
        const int img = t / p.tiles_h;
        const int kc_lo = p.grouped ? (phys(tile) % p.num_n_tiles) : 0;
        for (int kc = kc_lo; kc < kc_lo + kcn; ++kc) {
          mbar_wait(aempty_bar(stage), phase ^ 1u);
          if (lane == 0) {
            if (PAIR) {
              if (cta_rank == 0) mbar_arrive_expect_tx(afull_bar(stage), 2 * kPatchBytes);
              tma_load_4d_pair(smem_a + stage * kPatchStageBytes, &p.tmap_a, mapa_shared(afull_bar(stage), 0),
                               kc * kBK, tw * kPatchBW - 1, th * kPatchBH - 1, img);
            } else {
              mbar_arrive_expect_tx(afull_bar(stage), kPatchBytes);
              tma_load_4d(smem_a + stage * kPatchStageBytes, &p.tmap_a, afull_bar(stage), kc * kBK,
                          tw * kPatchBW - 1, th * kPatchBH - 1, img);
            }
          }
          __syncwarp();
          if (++stage == kPatchStages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (POOL) {
      for (int g = pool_g0; g < pool_g1;) {
        const int unit = g / p.pool_h;
        const int pa = g - unit * p.pool_h;
        const int len = min(p.pool_h - pa, pool_g1 - g);
        const int img = unit / p.tiles_w;
        const int strip = unit - img * p.tiles_w;
        for (int r = 2 * pa - 1; r <= 2 * (pa + len) - 1; ++r) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (lane == 0) {
            mbar_arrive_expect_tx(full_bar(stage), stage_tx);
            // 7 staged rows x 272 px from staged pixel 2 * (112 strip - 8), staged row 2 r (out of range: zeros)
            tma_load_4d(smem_a + stage * kABytes, &p.tmap_a, full_bar(stage), 0, 2 * kPoolStep / 8 * strip - 1, 2 * r,
                        img);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        g += len;
      }
    } else
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_tile = m_tile_of(tile);
      const int n_tile = phys(tile) % p.num_n_tiles;
      const int n0 = n_tile * BN;
      const int m_ld = PAIR ? min(m_tile, p.num_m_tiles - 1) : m_tile;
      // tile origin in the A coordinate space
      int cw = 0, ch = 0, cn = 0;
      if (p.a_mode == A_IM2COL) {
        const int m0 = m_ld * kBM;
        const int q0 = m0 % p.Wo;
        const int t = m0 / p.Wo;
        const int p0 = t % p.Ho;
        cn = t / p.Ho;
        cw = q0 * p.stride - p.pad;
        ch = p0 * p.stride - p.pad;
      } else if (p.a_mode >= A_STEM) {
        const int tw = m_tile % p.tiles_w;
        const int t = m_tile / p.tiles_w;
        const int th = t % p.tiles_h;
        cn = t / p.tiles_h;
        cw = tw * p.tile_bw;
        ch = th * p.tile_bh;
      }
      const int kc_lo = p.grouped ? n_tile : 0;
      for (int r = 0; r < p.kh; ++r) {
        for (int s = 0; s < p.kw; ++s) {
         for (int v = 0; v < nv; ++v) {
          for (int kc = kc_lo; kc < kc_lo + kcn; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (lane == 0) {
              const uint32_t dst_a = smem_a + stage * kABytes;
              const uint32_t dst_b = smem_b + stage * L::kBBytes;
              const int ka = (kc + (v == 1 ? p.k_chunks : 0)) * kBK;  // lo activations sit C channels further
              if (PAIR) {
                // both CTAs' bytes are counted on the leader's barrier, which the (leader's) MMA warp waits on
                if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * stage_tx);
                const uint32_t lb = mapa_shared(full_bar(stage), 0);
                if (p.a_mode == A_TILED) {
                  tma_load_2d_pair(dst_a, &p.tmap_a, lb, ka, m_ld * kBM);
                } else {
                  tma_load_im2col_4d_pair(dst_a, &p.tmap_a, lb, ka, cw, ch, cn, static_cast<uint16_t>(s * p.dil),
                                          static_cast<uint16_t>(r * p.dil));
                }
                tma_load_2d_pair(dst_b, &p.tmap_b, lb, (r * p.kw + s) * p.b_tap_stride + kc * kBK,
                                 n0 + static_cast<int>(cta_rank) * (BN / 2));
              }
              const uint32_t fb = full_bar(stage);
              if (!PAIR) mbar_arrive_expect_tx(fb, stage_tx);
              if (PAIR) {
              } else if (p.a_mode == A_TILED) {
                tma_load_2d(dst_a, &p.tmap_a, fb, ka, m_tile * kBM);
              } else if (p.a_mode == A_IM2COL) {
                tma_load_im2col_4d(dst_a, &p.tmap_a, fb, ka, cw, ch, cn,
                                   static_cast<uint16_t>(s * p.dil),
                                   static_cast<uint16_t>(r * p.dil));
              } else if (p.a_mode == A_SPATIAL) {
                tma_load_4d(dst_a, &p.tmap_a, fb, ka, cw, ch, cn);
              } else if (p.a_mode == A_STEM) {
                // filter row r of the 7x7 window = staged image row 2*(ho + r/2) + (r & 1)
                tma_load_5d(dst_a, &p.tmap_a, fb, 0, r & 1, cw, ch + (r >> 1), cn + (v == 1 ? p.a_lo_img : 0));
              } else {
                // A_STEM2: the 7 staged image rows x 272 pixels all windows of this tile live in,
                // copied linearly (no swizzle); coordinates (64-element chunk, chunk index, row, image)
                tma_load_4d(dst_a, &p.tmap_a, fb, 0, cw >> 3, 2 * ch, cn);
              }
              if (BRES_KB == 0 && !PAIR)
                tma_load_2d(dst_b, &p.tmap_b, fb,
                            (r * p.kw + s) * p.b_tap_stride + (v == 2 ? p.b_lo_off : 0) + kc * kBK, n0);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
         }
        }
      }
      if (!SPLIT && p.dual) {
        // second source: 1x1 / stride2 over x2 on this conv's output pixel grid; its weights are columns
        // [cin, cin + cin2) of the concatenated matrix
        const int m0 = m_ld * kBM;
        const int q0 = m0 % p.Wo;
        const int t2 = m0 / p.Wo;
        const int p0 = t2 % p.Ho;
        const int n2 = t2 / p.Ho;
        for (int kc = 0; kc < p.k_chunks2; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (lane == 0) {
            const uint32_t dst_a = smem_a + stage * kABytes;
            const uint32_t dst_b = smem_b + stage * L::kBBytes;
            const int kcol = p.cin + kc * kBK;
            if (PAIR) {
              if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * stage_tx);
              const uint32_t lb = mapa_shared(full_bar(stage), 0);
              if (p.stride2 == 1) tma_load_2d_pair(dst_a, &p.tmap_a2, lb, kc * kBK, m0);
              else tma_load_im2col_4d_pair(dst_a, &p.tmap_a2, lb, kc * kBK, q0 * p.stride2, p0 * p.stride2, n2, 0, 0);
              tma_load_2d_pair(dst_b, &p.tmap_b, lb, kcol, n0 + static_cast<int>(cta_rank) * (BN / 2));
            } else {
              const uint32_t fb = full_bar(stage);
              mbar_arrive_expect_tx(fb, stage_tx);
              if (p.stride2 == 1) tma_load_2d(dst_a, &p.tmap_a2, fb, kc * kBK, m0);
              else tma_load_im2col_4d(dst_a, &p.tmap_a2, fb, kc * kBK, q0 * p.stride2, p0 * p.stride2, n2, 0, 0);
              if (BRES_KB == 0) tma_load_2d(dst_b, &p.tmap_b, fb, kcol, n0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // kSingle: ONE thread runs the whole issue loop (no per-step warp reconvergence on the critical path)
    constexpr bool kSingle = TDET_MMA_SINGLE_THREAD != 0;
    if (!kSingle || lane == 0) {
    const uint32_t idesc = make_idesc_f16kind(PAIR ? 2 * kBM : kBM, BN, p.ab_fp16 ? kFmtF16 : kFmtBF16,
                                              p.b_fp16 ? kFmtF16 : kFmtBF16);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (BRES_KB > 0) mbar_wait(bres_bar, 0);
    // descriptor bases: an operand tile at byte offset x is base + (x >> 4)
    const uint64_t da0 = make_smem_desc_sw128(smem_a);
    const uint64_t db0 = make_smem_desc_sw128(smem_b);
    const uint64_t da_patch0 = make_smem_desc_sw128_sbo(smem_a, kPatchPW * 128);
    if (PATCH) {
      int as = 0;
      uint32_t aphase = 0;
      if (!PAIR || cta_rank == 0)
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kAccStride);
        const int kc_lo = p.grouped ? (phys(tile) % p.num_n_tiles) : 0;
        for (int kc = kc_lo; kc < kc_lo + kcn; ++kc) {
          mbar_wait(afull_bar(as), aphase);
          tc_fence_after();
          // (the issuing thread must stay ahead of 128-/64-wide MMAs of 64 / 32 tensor cycles: descriptors are
          // base + immediate, the tap loop is unrolled)
          const uint64_t da_patch = da_patch0 + static_cast<uint32_t>(as * (kPatchStageBytes >> 4));
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (BRES_KB == 0) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
            }
            if (kSingle || lane == 0) {
              const uint64_t da = da_patch + static_cast<uint32_t>(((tap / 3) * kPatchPW + tap % 3) * 8);
              const uint64_t db = db0 + static_cast<uint32_t>((BRES_KB > 0 ? tap * p.k_chunks + kc : stage) *
                                                              (L::kBBytes >> 4));
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k) {
                if (PAIR) umma_bf16_ss_pair(d_tmem, da + 2u * k, db + 2u * k, idesc, ((kc - kc_lo) | tap | k) != 0 ? 1u : 0u);
                else umma_bf16_ss(d_tmem, da + 2u * k, db + 2u * k, idesc, ((kc - kc_lo) | tap | k) != 0 ? 1u : 0u);
              }
              if (PAIR) {
                umma_commit_pair(empty_bar(stage));
                if (tap == 8) {
                  umma_commit_pair(aempty_bar(as));
                  if (kc == kc_lo + kcn - 1) umma_commit_pair(tfull_bar(acc));
                }
              } else {
                if (BRES_KB == 0) umma_commit(empty_bar(stage));
                if (tap == 8) {
                  umma_commit(aempty_bar(as));
                  if (kc == kc_lo + kcn - 1) umma_commit(tfull_bar(acc));
                }
              }
            }
            if (!kSingle) __syncwarp();
            if (BRES_KB == 0) {
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
          if (++as == kPatchStages) { as = 0; aphase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    } else if (!PAIR || cta_rank == 0)
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_main = tmem_base + static_cast<uint32_t>(acc * kAccStride);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        // split precision: pass v of this k-block (0 hi*hi -> main accumulator, 1/2 cross terms -> the second)
        const int v = (nv == 3) ? (kb / kcn) % 3 : 0;
        const uint32_t d_tmem = d_main + (v != 0 ? static_cast<uint32_t>(BN) : 0u);
        const int kb_first = (v != 0) ? kcn : 0;  // first k-block that writes this accumulator
        if (kSingle || lane == 0) {
          if (BRES_KB == 7 && p.a_mode == A_STEM2) {  // (the stem's seven filter rows are its resident k-blocks)
            // Row i of the A operand is the 8-pixel x 4-channel window starting at staged pixel 2*i:
            // an un-swizzled K-major view whose rows are 16 bytes apart and OVERLAP (the second
            // 8-element K chunk of row i is the first chunk of row i+1): LBO = 16 B, SBO = 128 B.
            // (one thread issues 14 MMAs of ~32 tensor cycles each: the descriptors are base + constant, so that
            // the issue loop stays shorter than the MMAs it feeds)
            const uint64_t da0 = make_smem_desc_nosw(smem_a + stage * kABytes, 16, 128);
            const uint64_t db0 = make_smem_desc_sw128(smem_b);
            const uint32_t row16 = static_cast<uint32_t>(p.stem_row_bytes) >> 4;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma_bf16_ss(d_tmem, da0 + (r * row16 + 2u * k), db0 + (r * (L::kBBytes >> 4) + 2u * k), idesc,
                             (r | k) != 0 ? 1u : 0u);
            }
          } else {
            const uint64_t da = da0 + static_cast<uint32_t>(stage * (kABytes >> 4));
            const uint64_t db = db0 + static_cast<uint32_t>((BRES_KB > 0 ? kb : stage) * (L::kBBytes >> 4));
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // advance 32 bytes (16 elements) along K inside the 128-byte swizzle row: +2 (16-byte units)
              if (PAIR) umma_bf16_ss_pair(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16_ss(d_tmem, da + 2u * k, db + 2u * k, idesc, ((kb - kb_first) | k) != 0 ? 1u : 0u);
            }
          }
          if (PAIR) {
            // frees the stage / publishes the accumulator in both CTAs
            umma_commit_pair(empty_bar(stage));
            if (kb == num_kb - 1) umma_commit_pair(tfull_bar(acc));
          } else {
            umma_commit(empty_bar(stage));
            if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
          }
        }
        if (!kSingle) __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (kAccBufs == 2) acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ TMA producer (residual, mask)
    // Per 64-column output slab the ring receives the residual slab (if any), then the mask slab (if it
    // is TMA-staged), in tile order; the two epilogue groups consume alternate slabs from this ONE ring
    // (sharing it lets the faster group borrow slots; two private rings of half the depth measured 15 %
    // slower on the residual convs).
    const bool rsplit = SPLIT && p.split && p.has_res;
    const int nload = coarse_tma ? 1 : rsplit ? 2 : (p.has_res ? 1 : 0) + ((MASKED && p.mask_tma) ? 1 : 0);
    if (RES_SLABS > 0 && nload > 0) {
      int rs = 0;
      int issued = 0;
      uint32_t rphase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m_tile = m_tile_of(tile);
        const int n_tile = phys(tile) % p.num_n_tiles;
        // The ring only looks ~one slab ahead of the epilogue, far less than the HBM latency under load: pull the
        // residual / mask tile of a later tile into L2 now (128 rows x BN columns; 2D operands only)
        if (p.res_prefetch > 0 && !coarse_tma && p.a_mode < A_STEM && lane == 0) {
          const int ptile = tile + p.res_prefetch * tile_step;
          if (ptile < num_tiles) {
            const int pm = PAIR ? min(m_tile_of(ptile), p.num_m_tiles - 1) : m_tile_of(ptile);
            const int pn = (phys(ptile) % p.num_n_tiles) * BN;
            for (int s = 0; s < kSlabsPerTile; ++s) {
              if (p.has_res) tma_prefetch_l2_2d(&p.tmap_res, pn + s * 64, pm * kBM);
              if (MASKED && p.mask_tma) tma_prefetch_l2_2d(&p.tmap_mask, pn + s * 64, pm * kBM);
            }
          }
        }
        for (int s = 0; s < kSlabsPerTile; ++s) {
          for (int j = 0; j < nload; ++j) {
            const CUtensorMap* tm = (rsplit || (j == 0 && p.has_res)) ? &p.tmap_res : &p.tmap_mask;
            const int cres = n_tile * BN + s * 64 + ((rsplit && j == 1) ? p.N : 0);  // lo half: N channels on
            mbar_wait(rempty_bar(rs), rphase ^ 1u);
            if (lane == 0) {
              mbar_arrive_expect_tx(rfull_bar(rs), rslot);
              if (coarse_tma) {
                // the 4 x 8 coarse pixels under this 8 x 16 tile
                const int tw = m_tile % p.tiles_w;
                const int t = m_tile / p.tiles_w;
                tma_load_4d(smem_res + rs * rslot, &p.tmap_coarse, rfull_bar(rs), cres, tw * (p.tile_bw >> 1),
                            (t % p.tiles_h) * (p.tile_bh >> 1), t / p.tiles_h);
              } else if (p.a_mode >= A_STEM) {
                const int tw = m_tile % p.tiles_w;
                const int t = m_tile / p.tiles_w;
                tma_load_4d(smem_res + rs * kSlabBytes, tm, rfull_bar(rs), cres,
                            tw * p.tile_bw, (t % p.tiles_h) * p.tile_bh, t / p.tiles_h);
              } else {
                tma_load_2d(smem_res + rs * kSlabBytes, tm, rfull_bar(rs), cres,
                            (PAIR ? min(m_tile, p.num_m_tiles - 1) : m_tile) * kBM);
              }
              *ring_issued = issued;  // this fill's predecessor in the slot has been consumed, hence completed
            }
            ++issued;
            __syncwarp();
            if (++rs == rdepth) { rs = 0; rphase ^= 1u; }
          }
        }
      }
    }
  } else if (PATCH && warp == 11) {
    // ------------------------------------------------------------------ TMA producer (B, patch mode)
    if (PATCH && BRES_KB == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int n0 = (phys(tile) % p.num_n_tiles) * BN;
        const int kc_lo = p.grouped ? (phys(tile) % p.num_n_tiles) : 0;
        for (int kc = kc_lo; kc < kc_lo + kcn; ++kc) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (lane == 0) {
              if (PAIR) {
                if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * L::kBBytes);
                tma_load_2d_pair(smem_b + stage * L::kBBytes, &p.tmap_b, mapa_shared(full_bar(stage), 0),
                                 tap * p.b_tap_stride + kc * kBK, n0 + static_cast<int>(cta_rank) * (BN / 2));
              } else {
                mbar_arrive_expect_tx(full_bar(stage), L::kBBytes);
                tma_load_2d(smem_b + stage * L::kBBytes, &p.tmap_b, full_bar(stage), tap * p.b_tap_stride + kc * kBK, n0);
              }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 3..10)
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;      // accumulator row == TMEM lane == staging row
    const int group = (warp - 3) >> 2;     // epilogue group: slabs (or tiles) with index % NG == group
    constexpr bool kByTile = kSlabsPerTile == 1;
    const int gtid = (threadIdx.x - 96) & (kEpiGroupThreads - 1);
    const uint32_t gbar = 1u + group;      // the group's named barrier
    const bool ws = p.warp_stores != 0 && !POOL;
    const bool issuer = ws ? (lane == 0) : (gtid == 0);  // issues the TMA stores (its warp's / its group's), frees residual slabs
    const bool has_res_rt = RES_SLABS > 0 && p.has_res;
    const bool has_res = has_res_rt;
    const bool mask_tma = MASKED && RES_SLABS > 0 && p.mask_tma != 0;
    const bool split = SPLIT && p.split != 0;
    static_assert(!SPLIT || OSLABS == 2, "split precision stages the hi and the lo slab side by side");
    // split mode: the residual's lo slab rides in the ring slot the mask would use
    const int nload = coarse_tma ? 1 : (split && has_res) ? 2 : (has_res ? 1 : 0) + (mask_tma ? 1 : 0);
    const bool has_coarse = p.coarse != nullptr;
    const bool out_fp16_rt = p.out_fp16 != 0, res_fp16_rt = p.res_fp16 != 0, co_fp16_rt = p.coarse_fp16 != 0;
    const bool out_fp16 = out_fp16_rt;
    float* s_scale = s_params + group * 2 * L::kGroupCols;
    float* s_shift = s_scale + L::kGroupCols;
    const uint32_t smem_out_g = smem_out + group * OSLABS * kSlabBytes;

    // per-tensor exponents (all zero when no metadata is attached)
    const int e_in = p.in_meta ? p.in_meta->e : 0;
    const int e_res = p.res_meta ? p.res_meta->e : 0;
    const int e_co = p.coarse_meta ? p.coarse_meta->e : 0;
    int e_out = 0;
    if (p.out_scaled) {
      float bound = p.bound_consts[1];
      float a_in = p.in_meta ? __uint_as_float(p.in_meta->amax_bits) : 0.0f;
      if (p.dual && p.in2_meta) a_in = fmaxf(a_in, __uint_as_float(p.in2_meta->amax_bits));  // (G of [W | W2])
      bound += p.bound_consts[0] * a_in;
      if (p.res_meta && p.has_res) bound += __uint_as_float(p.res_meta->amax_bits);
      if (p.coarse_meta && has_coarse) bound += __uint_as_float(p.coarse_meta->amax_bits);
      if (bound > 0.0f && bound < 3.0e38f) e_out = ilogbf(bound) - 14;  // bound * 2^-e_out < 2^15
      e_out = max(-100, min(100, e_out));
      if (blockIdx.x == 0 && threadIdx.x == 96) p.out_meta->e = e_out;
    }
    const float mul_in = ldexpf(1.0f, e_in - e_out);
    const float mul_shift = ldexpf(1.0f, -e_out);
    const float mul_res = ldexpf(1.0f, e_res - e_out);
    const float mul_co = ldexpf(1.0f, e_co - e_out);
    float amax_local = 0.0f;
    // conversion-step variant (EpiFast): the common forward epilogues run with their flags fixed at compile time
    int epi_variant = 0;
    if (!MASKED && !(SPLIT && split) && !POOL && p.relu != 2 && p.epi_fast) {
      const bool r1 = p.relu == 1;
      if (!has_res_rt && !has_coarse) epi_variant = out_fp16_rt ? (r1 ? 1 : 3) : (r1 ? 2 : 4);
      else if (has_res_rt && !has_coarse && res_fp16_rt && r1) epi_variant = out_fp16_rt ? 5 : 6;
      else if (!has_res_rt && has_coarse && !co_fp16_rt && !out_fp16_rt && !r1) epi_variant = 7;
    }

    if (POOL) {
      // ---- stem + 3x3/2 max-pool: this group owns channels [32 group, 32 group + 32) of every conv row.
      // Vertical max in registers (each thread keeps the packed previous two rows of its own pixel), horizontal
      // max through one shared-memory row (double buffered: one named barrier per pooled row).
      const uint32_t vbuf = smem_out_g;  // two 128 px x 64 B rows (the group's staging slabs are free)
      for (int i = gtid; i < 64; i += kEpiGroupThreads) {
        s_scale[i] = (p.scale ? __ldg(p.scale + i) : 1.0f) * mul_in;
        s_shift[i] = (p.shift ? __ldg(p.shift + i) : 0.0f) * mul_shift;
      }
      named_bar_sync(gbar, kEpiGroupThreads);
      uint8_t* const out = static_cast<uint8_t*>(p.pool_out);
      uint32_t amax_pk = 0;  // running max of the packed (non-negative) outputs, both 16-bit halves
      int pseq = 0;
      int vb = 0;
      // Accumulator rows are fetched one tile ahead (two register buffers): the TMEM read latency and the wait for
      // the MMA hide behind the arithmetic of the previous row, and the accumulator is handed back early.
      int ld_acc = 0;
      auto issue_ld = [&](uint32_t (&v)[32]) {
        ld_acc = pseq & 1;
        mbar_wait(tfull_bar(ld_acc), static_cast<uint32_t>(pseq >> 1) & 1u);
        ++pseq;
        tc_fence_after();
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(ld_acc * kAccStride + group * 32), v);
      };
      auto finish_ld = [&]() {
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(ld_acc));
      };
      // (the output format as a compile-time constant: a run-time flag costs a second, predicated pack per pair)
      auto pool_rows = [&](auto out16_c) {
        constexpr bool out_fp16 = decltype(out16_c)::value;
        // one conv row of this thread's pixel: BN + ReLU, packed to 16 x (2 x 16 bit); zeros outside the image
        auto conv_row = [&](const uint32_t (&v)[32], int r, int col, uint32_t (&o)[16]) {
          if (r >= 0 && r < p.Ho && col >= 0 && col < p.Wo) {
  #pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 sc = *reinterpret_cast<const float4*>(s_scale + group * 32 + 4 * j);
              const float4 sh = *reinterpret_cast<const float4*>(s_shift + group * 32 + 4 * j);
              // (PTX max: max(-0, +0) = +0, so the packed values order like unsigned integers)
              o[2 * j] = pack16x2(fmaxf(fmaf(__uint_as_float(v[4 * j]), sc.x, sh.x), 0.0f),
                                  fmaxf(fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y), 0.0f), out_fp16);
              o[2 * j + 1] = pack16x2(fmaxf(fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z), 0.0f),
                                      fmaxf(fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w), 0.0f), out_fp16);
            }
          } else {
  #pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = 0u;
          }
        };
        for (int g = pool_g0; g < pool_g1;) {
          const int unit = g / p.pool_h;
          const int pa = g - unit * p.pool_h;
          const int len = min(p.pool_h - pa, pool_g1 - g);
          const int img = unit / p.tiles_w;
          const int strip = unit - img * p.tiles_w;
          const int col = 2 * kPoolStep * strip - 8 + row;  // conv column of this thread's accumulator row
          uint32_t va[32], vb2[32];
          uint32_t prev[16], vm[16];
          issue_ld(va);                 // row 2 pa - 1
          finish_ld();
          issue_ld(vb2);                // row 2 pa
          conv_row(va, 2 * pa - 1, col, prev);
          for (int pr = pa; pr < pa + len; ++pr) {
            finish_ld();
            issue_ld(va);               // row 2 pr + 1
            conv_row(vb2, 2 * pr, col, vm);
  #pragma unroll
            for (int j = 0; j < 16; ++j) vm[j] = __vmaxu2(vm[j], prev[j]);
            finish_ld();
            if (pr + 1 < pa + len) issue_ld(vb2);  // row 2 pr + 2
            conv_row(va, 2 * pr + 1, col, prev);
            const uint32_t dst = vbuf + static_cast<uint32_t>(vb) * 8192u + row * 64;
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 m;
              m.x = __vmaxu2(vm[4 * j], prev[4 * j]);
              m.y = __vmaxu2(vm[4 * j + 1], prev[4 * j + 1]);
              m.z = __vmaxu2(vm[4 * j + 2], prev[4 * j + 2]);
              m.w = __vmaxu2(vm[4 * j + 3], prev[4 * j + 3]);
              amax_pk = __vmaxu2(amax_pk, __vmaxu2(__vmaxu2(m.x, m.y), __vmaxu2(m.z, m.w)));
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((j ^ ((row >> 1) & 3)) << 4)),
                           "r"(m.x), "r"(m.y), "r"(m.z), "r"(m.w)
                           : "memory");
            }
            named_bar_sync(gbar, kEpiGroupThreads);
            // pooled row pr: horizontal max over conv columns 2q-1 .. 2q+1 (this strip's rows 2 ql + 7 .. + 9)
            const uint32_t rb = vbuf + static_cast<uint32_t>(vb) * 8192u;
            for (int it = gtid; it < kPoolStep * 4; it += kEpiGroupThreads) {
              const int ql = it >> 2, c = it & 3;
              const int q = kPoolStep * strip + ql;
              if (q < p.pool_w) {
                uint4 t[3];
  #pragma unroll
                for (int dc = 0; dc < 3; ++dc) {
                  const int i = 2 * ql + 7 + dc;
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(t[dc].x), "=r"(t[dc].y), "=r"(t[dc].z), "=r"(t[dc].w)
                               : "r"(rb + i * 64 + ((c ^ ((i >> 1) & 3)) << 4)));
                }
                uint4 m;
                m.x = __vmaxu2(__vmaxu2(t[0].x, t[1].x), t[2].x);
                m.y = __vmaxu2(__vmaxu2(t[0].y, t[1].y), t[2].y);
                m.z = __vmaxu2(__vmaxu2(t[0].z, t[1].z), t[2].z);
                m.w = __vmaxu2(__vmaxu2(t[0].w, t[1].w), t[2].w);
                stg_v4(out + ((static_cast<long long>(img) * p.pool_h + pr) * p.pool_w + q) * 128 + group * 64 + c * 16, m);
              }
            }
            vb ^= 1;  // (the other buffer was last read before this row's barrier)
          }
          g += len;
        }
      };
      if (out_fp16_rt) pool_rows(std::true_type{}); else pool_rows(std::false_type{});
      {
        float lo, hi;
        unpack16x2(amax_pk, out_fp16, lo, hi);
        amax_local = fmaxf(lo, hi);
      }
    }
    int cur_n_tile = -1;
    int ob = 0;  // staging buffer of the next slab
    int seq = 0; // CTA-local tile counter: tile seq accumulates in TMEM buffer seq & 1
    for (int tile = POOL ? num_tiles : tile0; tile < num_tiles; tile += tile_step, ++seq) {
      if (kByTile && (seq & 1) != group) continue;
      const int acc = kAccBufs == 2 ? (seq & 1) : 0;
      const uint32_t acc_phase = static_cast<uint32_t>(kAccBufs == 2 ? (seq >> 1) : seq) & 1u;
      const int m_tile = m_tile_of(tile);
      const int n_tile = phys(tile) % p.num_n_tiles;
      const int n0 = n_tile * BN;
      if (n_tile != cur_n_tile) {
        named_bar_sync(gbar, kEpiGroupThreads);
        for (int i = gtid; i < L::kGroupCols; i += kEpiGroupThreads) {
          // local column i of this group = column (i & 63) of its (i >> 6)-th slab
          const int col = kByTile ? i : (((i >> 6) * NG + group) * 64 + (i & 63));
          s_scale[i] = (p.scale ? __ldg(p.scale + n0 + col) : 1.0f) * mul_in;
          s_shift[i] = (p.shift ? __ldg(p.shift + n0 + col) : 0.0f) * mul_shift;
        }
        named_bar_sync(gbar, kEpiGroupThreads);
        cur_n_tile = n_tile;
      }
      // output pixel of this thread's row
      bool valid;
      long long pix;
      int st_c1 = 0, st_c2 = 0, st_c3 = 0;  // TMA store coordinates beyond the channel
      if (p.a_mode >= A_STEM) {
        const int tw = m_tile % p.tiles_w;
        const int t = m_tile / p.tiles_w;
        const int th = t % p.tiles_h;
        const int img = t / p.tiles_h;
        const int hl = row / p.tile_bw;
        const int wl = row - hl * p.tile_bw;
        const int ho = th * p.tile_bh + hl;
        const int wo = tw * p.tile_bw + wl;
        valid = (ho < p.Ho) && (wo < p.Wo) && !(PAIR && m_tile >= p.num_m_tiles);
        pix = (static_cast<long long>(img) * p.Ho + ho) * p.Wo + wo;
        st_c1 = tw * p.tile_bw;
        st_c2 = th * p.tile_bh;
        st_c3 = img;
      } else {
        const int m = m_tile * kBM + row;
        valid = m < p.M;
        pix = m;
        st_c1 = m_tile * kBM;
      }
      const uint8_t* coarse_row_rt = nullptr;
      // TMA-staged coarse operand: this thread's pixel (hl, wl) of the 8 x 16 tile reads coarse box row
      // (hl/2)*4 + wl/2; parity mode only adds at even pixels
      int crow_rt = -1;
      if (coarse_tma && valid) {
        const int hl = row / p.tile_bw, wl = row - hl * p.tile_bw;
        if (!(p.coarse_parity && ((hl | wl) & 1))) crow_rt = (hl >> 1) * (p.tile_bw >> 1) + (wl >> 1);
      }
      if (valid && has_coarse && !coarse_tma) {
        const int m = static_cast<int>(pix);
        const int q = m % p.Wo;
        const int t = m / p.Wo;
        const int pp = t % p.Ho;
        const int img = t / p.Ho;
        if (!(p.coarse_parity && ((pp | q) & 1)))
          coarse_row_rt = static_cast<const uint8_t*>(p.coarse) +
                       (((static_cast<long long>(img) * p.Hc + (pp >> 1)) * p.Wc + (q >> 1)) *
                            (split ? 2 * p.N : p.N) + n0) * 2;
      }
      const uint8_t* mask_row = nullptr;
      if (MASKED && valid && p.mask_src && !mask_tma)
        mask_row = static_cast<const uint8_t*>(p.mask_src) + (pix * p.N + n0) * 2;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                              static_cast<uint32_t>(acc * kAccStride);
#pragma unroll 1
      for (int slab = kByTile ? 0 : group; slab < kSlabsPerTile; slab += kByTile ? 1 : NG) {
        // residual slabs are produced in tile order into one ring shared by both groups; consecutive
        // slabs of a group are two ring positions apart, so with any ring depth >= 2 the previous fill of
        // a slot has completed before the group waits for the next one (no parity aliasing)
        const int ridx = (seq * kSlabsPerTile + slab) * nload;
        const int rs = ridx % rdepth;
        const uint32_t rphase = static_cast<uint32_t>(ridx / rdepth) & 1u;
        const int midx = ridx + (has_res ? 1 : 0);
        const int ms = midx % rdepth;
        const uint32_t mphase = static_cast<uint32_t>(midx / rdepth) & 1u;
        // the staging buffer `ob` was last read by the TMA store this group (warp) issued OSLABS slabs ago
        if (issuer) {
          if (OSLABS == 1 || split) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        if (ws) __syncwarp();
        // An mbarrier parity wait is only sound if the PREVIOUS fill of the slot (ring index - kRS) has
        // completed before the consumer waits for this one -- otherwise the barrier is still one phase
        // behind and the wait falls through.  A single in-order consumer has that for free; with two groups
        // the previous fill may belong to the other group.  The producer issues a fill only after its
        // predecessor was consumed, so "the producer has issued ring index r" implies it: wait for that
        // first (no extra coupling: the data cannot arrive before it is requested).
        if (nload > 0) {
          uint32_t spins = 0;
          while (*ring_issued < ridx + nload - 1) {
            if (++spins > (1u << 24)) __trap();
          }
        }
        if (has_res || coarse_tma) mbar_wait(rfull_bar(rs), rphase);
        if (mask_tma || (split && has_res)) mbar_wait(rfull_bar(ms), mphase);
        if (!ws) named_bar_sync(gbar, kEpiGroupThreads);
        if (split) ob = 0;
        const uint32_t out_row = smem_out_g + ob * kSlabBytes + row * 128;
        const uint32_t res_row = smem_res + rs * kSlabBytes + row * 128;
        const uint32_t mk_row = smem_res + ms * kSlabBytes + row * 128;
        auto convert_slab = [&](auto fast) {
#pragma unroll 1
        for (int half = 0; half < kParts; ++half) {   // (a 32- or 16-column step of the slab)
          // Fast variants fix the operand set and the storage formats at compile time.  With run-time flags the
          // compiler if-converts every per-element format / operand branch of the unrolled loops and computes BOTH
          // sides (SASS of the residual kernel: 560 instructions per 32 columns, 17 per element, of them 160 FFMA and
          // 64 fp16 + 57 bf16 unpack operations) -- and these kernels are bound by exactly this instruction stream.
          constexpr int kMode = decltype(fast)::mode;   // 0 generic, 1 no operand, 2 fp16 residual, 3 bf16 coarse
          constexpr bool kGen = kMode == 0;
          const bool has_res = kGen ? has_res_rt : (kMode == 2);
          const bool res_fp16 = kGen ? res_fp16_rt : true;
          const bool co_fp16 = kGen ? co_fp16_rt : false;
          const bool out_fp16 = kGen ? out_fp16_rt : decltype(fast)::out16;
          const int relu = kGen ? p.relu : (decltype(fast)::relu ? 1 : 0);
          const uint8_t* const coarse_row = (kGen || kMode == 3) ? coarse_row_rt : nullptr;
          const int crow = (kGen || kMode == 3) ? crow_rt : -1;
          uint32_t v[kEW];
          tmem_ld_cols<kEW>(t_addr + slab * 64 + half * kEW, v);
          uint4 rco[kEJ];
          if (coarse_row) {
#pragma unroll
            for (int j = 0; j < kEJ; ++j) rco[j] = ldg_nc_v4(coarse_row + (slab * 64 + half * kEW + j * 8) * 2);
          }
          uint4 rmk[kEJ];
          if (!MASKED) {
          } else if (mask_tma) {
#pragma unroll
            for (int j = 0; j < kEJ; ++j) {
              const uint32_t a = mk_row + (((half * kEJ + j) ^ (row & 7)) << 4);
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(rmk[j].x), "=r"(rmk[j].y), "=r"(rmk[j].z), "=r"(rmk[j].w)
                           : "r"(a));
            }
          } else if (mask_row) {
#pragma unroll
            for (int j = 0; j < kEJ; ++j) rmk[j] = ldg_nc_v4(mask_row + (slab * 64 + half * kEW + j * 8) * 2);
          }
          uint4 rres[kEJ];
          if (has_res) {
#pragma unroll
            for (int j = 0; j < kEJ; ++j) {
              const uint32_t a = res_row + (((half * kEJ + j) ^ (row & 7)) << 4);
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(rres[j].x), "=r"(rres[j].y), "=r"(rres[j].z), "=r"(rres[j].w)
                           : "r"(a));
            }
          }
          tmem_ld_wait();
          float x[kEW];
          if (SPLIT && split) {
            // main (hi*hi) + cross (lo*hi + hi*lo) accumulators, summed in fp32
#pragma unroll
            for (int i = 0; i < kEW; ++i) x[i] = __uint_as_float(v[i]);
            tmem_ld_cols<kEW>(t_addr + BN + slab * 64 + half * kEW, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < kEW; ++i) v[i] = __float_as_uint(x[i] + __uint_as_float(v[i]));
          }
          const int cb = (kByTile ? slab : (slab / NG)) * 64 + half * kEW;  // group-local column
#pragma unroll
          for (int j = 0; j < kEW / 4; ++j) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + cb + j * 4);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + cb + j * 4);
            x[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, sh.x);
            x[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y);
            x[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z);
            x[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w);
          }
          if (has_res) {
#pragma unroll
            for (int j = 0; j < kEJ; ++j) {
              const uint32_t w4[4] = {rres[j].x, rres[j].y, rres[j].z, rres[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float lo, hi;
                unpack16x2(w4[e], res_fp16, lo, hi);
                x[8 * j + 2 * e] = fmaf(lo, mul_res, x[8 * j + 2 * e]);
                x[8 * j + 2 * e + 1] = fmaf(hi, mul_res, x[8 * j + 2 * e + 1]);
              }
            }
            if (SPLIT && split) {
              // lo half of the residual pair (next ring slot)
#pragma unroll
              for (int j = 0; j < kEJ; ++j) {
                const uint32_t a = mk_row + (((half * kEJ + j) ^ (row & 7)) << 4);
                uint4 r2;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r2.x), "=r"(r2.y), "=r"(r2.z), "=r"(r2.w)
                             : "r"(a));
                const uint32_t w4[4] = {r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  x[8 * j + 2 * e] += bf16_lo(w4[e]);
                  x[8 * j + 2 * e + 1] += bf16_hi(w4[e]);
                }
              }
            }
          }
          if (coarse_row) {
#pragma unroll
            for (int j = 0; j < kEJ; ++j) {
              const uint32_t w4[4] = {rco[j].x, rco[j].y, rco[j].z, rco[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float lo, hi;
                unpack16x2(w4[e], co_fp16, lo, hi);
                x[8 * j + 2 * e] = fmaf(lo, mul_co, x[8 * j + 2 * e]);
                x[8 * j + 2 * e + 1] = fmaf(hi, mul_co, x[8 * j + 2 * e + 1]);
              }
            }
            if (SPLIT && split) {
#pragma unroll
              for (int j = 0; j < kEJ; ++j) {
                const uint4 c2 = ldg_nc_v4(coarse_row + (p.N + slab * 64 + half * kEW + j * 8) * 2);
                const uint32_t w4[4] = {c2.x, c2.y, c2.z, c2.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  x[8 * j + 2 * e] += bf16_lo(w4[e]);
                  x[8 * j + 2 * e + 1] += bf16_hi(w4[e]);
                }
              }
            }
          }
          if (crow >= 0) {
            const uint32_t cbase = smem_res + rs * rslot + crow * 128;
#pragma unroll
            for (int j = 0; j < kEJ; ++j) {
              const uint32_t a = cbase + (((half * kEJ + j) ^ (crow & 7)) << 4);
              uint4 c4;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(c4.x), "=r"(c4.y), "=r"(c4.z), "=r"(c4.w)
                           : "r"(a));
              const uint32_t w4[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float lo, hi;
                unpack16x2(w4[e], co_fp16, lo, hi);
                x[8 * j + 2 * e] = fmaf(lo, mul_co, x[8 * j + 2 * e]);
                x[8 * j + 2 * e + 1] = fmaf(hi, mul_co, x[8 * j + 2 * e + 1]);
              }
            }
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < kEW; ++i) x[i] = fmaxf(x[i], 0.0f);
            if (relu == 2) {   // ReLU6 (plain outputs only: x is the true value)
#pragma unroll
              for (int i = 0; i < kEW; ++i) x[i] = fminf(x[i], 6.0f);
            }
          }
          if (MASKED && (mask_tma ? valid : (mask_row != nullptr))) {
#pragma unroll
            for (int j = 0; j < kEJ; ++j) {
              const uint32_t w4[4] = {rmk[j].x, rmk[j].y, rmk[j].z, rmk[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // forward activations are post-ReLU (>= 0): positive <=> non-zero magnitude bits
                if ((w4[e] & 0x00007FFFu) == 0u) x[8 * j + 2 * e] = 0.0f;
                if ((w4[e] & 0x7FFF0000u) == 0u) x[8 * j + 2 * e + 1] = 0.0f;
              }
            }
          }
          if (valid && p.out_meta) {
#pragma unroll
            for (int i = 0; i < kEW; ++i) amax_local = fmaxf(amax_local, fabsf(x[i]));
          }
#pragma unroll
          for (int j = 0; j < kEJ; ++j) {
            uint4 o;
            o.x = pack16x2(x[8 * j + 0], x[8 * j + 1], out_fp16);
            o.y = pack16x2(x[8 * j + 2], x[8 * j + 3], out_fp16);
            o.z = pack16x2(x[8 * j + 4], x[8 * j + 5], out_fp16);
            o.w = pack16x2(x[8 * j + 6], x[8 * j + 7], out_fp16);
            const uint32_t a = out_row + (((half * kEJ + j) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(o.x), "r"(o.y),
                         "r"(o.z), "r"(o.w)
                         : "memory");
            if (SPLIT && split) {
              // lo = bf16(x - hi): the second staging slab of the group
              uint4 l;
              l.x = pack_bf16x2(x[8 * j + 0] - bf16_lo(o.x), x[8 * j + 1] - bf16_hi(o.x));
              l.y = pack_bf16x2(x[8 * j + 2] - bf16_lo(o.y), x[8 * j + 3] - bf16_hi(o.y));
              l.z = pack_bf16x2(x[8 * j + 4] - bf16_lo(o.z), x[8 * j + 5] - bf16_hi(o.z));
              l.w = pack_bf16x2(x[8 * j + 6] - bf16_lo(o.w), x[8 * j + 7] - bf16_hi(o.w));
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a + kSlabBytes), "r"(l.x), "r"(l.y),
                           "r"(l.z), "r"(l.w)
                           : "memory");
            }
          }
        }
        };
        switch (epi_variant) {
          case 1: convert_slab(EpiFast<1, true, true>{}); break;
          case 2: convert_slab(EpiFast<1, false, true>{}); break;
          case 3: convert_slab(EpiFast<1, true, false>{}); break;
          case 4: convert_slab(EpiFast<1, false, false>{}); break;
          case 5: if constexpr (RES_SLABS > 0) { convert_slab(EpiFast<2, true, true>{}); } break;
          case 6: if constexpr (RES_SLABS > 0) { convert_slab(EpiFast<2, false, true>{}); } break;
          case 7: convert_slab(EpiFast<3, false, false>{}); break;
          default: convert_slab(EpiFast<0, false, false>{}); break;
        }
        if (slab + (kByTile ? 1 : NG) >= kSlabsPerTile) {
          // this warp's TMEM reads of the accumulator are complete: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
            else mbar_arrive(tempty_bar(acc));
          }
        }
        fence_proxy_async_smem();  // staging writes -> visible to the TMA (async proxy)
        if (ws) __syncwarp(); else named_bar_sync(gbar, kEpiGroupThreads);
        if (issuer) {
          // warp_stores: rows [32 quad, 32 quad + 32) of the slab = a (64, 32)-row box, or the sub-box of the spatial
          // tile those rows cover
          const uint32_t src = smem_out_g + ob * kSlabBytes + (ws ? quad * 4096 : 0);
          int c1 = st_c1, c2 = st_c2;
          if (ws) {
            if (p.a_mode >= A_STEM) {
              c1 += (quad * 32) % p.tile_bw;
              c2 += (quad * 32) / p.tile_bw;
            } else {
              c1 += quad * 32;
            }
          }
          if (PAIR && m_tile >= p.num_m_tiles) {
            // the odd pair's padding tile: nothing to store
          } else if (p.a_mode >= A_STEM) {
            asm volatile(
                "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)), "r"(src), "r"(n0 + slab * 64),
                "r"(c1), "r"(c2), "r"(st_c3)
                : "memory");
          } else {
            asm volatile(
                "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)), "r"(src), "r"(n0 + slab * 64),
                "r"(c1)
                : "memory");
          }
          if (SPLIT && split) {
            if (p.a_mode >= A_STEM) {
              asm volatile(
                  "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                  ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)), "r"(src + kSlabBytes), "r"(p.N + n0 + slab * 64),
                  "r"(c1), "r"(c2), "r"(st_c3)
                  : "memory");
            } else {
              asm volatile(
                  "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                  ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)), "r"(src + kSlabBytes), "r"(p.N + n0 + slab * 64),
                  "r"(c1)
                  : "memory");
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (has_res || coarse_tma) mbar_arrive(rempty_bar(rs));  // every thread passed the barrier: slab consumed
          if (mask_tma || (split && has_res)) mbar_arrive(rempty_bar(ms));
        }
        if (OSLABS > 1 && !split) ob ^= 1;
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
  if (PAIR) cluster_sync_all(); else __syncthreads();  // (pair: the peer may still signal this CTA's barriers)
  if (warp == 3) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Debug/test kernel: load ONE im2col A tile and write it out un-swizzled as [128][64] 16-bit.
__global__ void __launch_bounds__(128, 1)
im2col_tile_dump_kernel(const __grid_constant__ CUtensorMap tmap, int c, int w, int h, int n,
                        int off_w, int off_h, __nv_bfloat16* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0u) __trap();
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
