// Fused bottleneck tail for planes = 128 (layer2 of the bottleneck ResNets, stride 1, identity shortcut):
//
//     z2   = relu(bn2(conv3x3(z1)))                          resnet.py:105-108   (z1 = this block's conv1 output, 128 ch)
//     out  = relu(bn3(conv1x1(z2)) + x)                      resnet.py:110-118   (x = block input, 512 ch)
//
// in ONE persistent kernel, the planes = 128 sibling of bottleneck_tail_kernel<0> (bottleneck_fused.cuh): z2 goes
// TMEM -> registers -> two swizzled shared-memory slabs that ARE the A operand of the second GEMM and never reaches
// HBM.  Per block (batch 16 @800x1344: 100 x 168 pixels) the unfused pair moves 69 + 69 MB (conv2) and 69 + 275 + 275
// MB (conv3 + residual); fused: 69 + 275 + 275 MB and one launch less.
//
// What differs from the 64-plane kernel is what no longer fits: W2 is 288 KiB and W3 128 KiB, so BOTH stream through
// one ring of four 16 KiB slots in exactly the order the tensor pipe consumes them (L2 hits: 416 KiB per 128-pixel
// tile), and the 512 output channels are produced in two passes of 256 columns through the same four output slabs.
//
// One CTA per SM, tiles of 8 x 16 output pixels (128 GEMM rows), per tile:
//   G1  D1[128 x 128] = patch(z1)[128 x 2*9*64] * W2^T   two 64-channel halo patches ((8+2) x (16+2) TMA boxes), nine
//                                                         shifted descriptors each: 18 k-blocks of four N = 128 MMAs
//   G2  D2[128 x 256] = z2[128 x 128] * W3[h]^T, h = 0, 1 two k-blocks of four N = 256 MMAs per pass
// TMEM: D1 double-buffered (2 x 128 columns) + D2 (256) = 512 columns.
//
// Shared memory (227 KiB): 2 patch stages x 2 chunks (92 KiB; z2(k) is written over the patch of tile k once G1(k) has
// read it), W ring 64 KiB, four output slabs 64 KiB (the residual lands in them by TMA, the epilogue adds in place).
//
// Issue order of the one tcgen05 thread (k-blocks b = chunk * 9 + tap of G1):
//     G1(0) | b 0-8 of G1(1) | { G2(k,0) | b 9-14 of G1(k+1) | G2(k,1) | b 15-17 of G1(k+1) | b 0-8 of G1(k+2) } ...
// Six k-blocks (~3500 cycles) separate the two passes of a tile -- E2(k,0) drains D2 meanwhile -- and three more
// cover the load of patch k+2, whose stage (it held z2(k)) is only free after G2(k,1).  A W3 tile is two consecutive
// slots that must not wrap: every iteration consumes an even number of slots (a dummy slot pads the odd cases), so
// W3 tiles always start on an even slot.
//
// Warp roles (20 warps): warp 0 lane 0 patches, lane 1 the W ring; warp 1 MMA issuer; warp 2 residual producer (into
// the output slabs; the NEXT pass's residual is prefetched into L2); warp 3 store issuer; warps 4-19 epilogue in four
// groups: E1 (D1 -> z2, 32 columns per group), E2 (D2 + residual -> out slab g of the pass, in place).
// Numerics are those of the unfused launches (fp32 accumulation, k-blocks in chunk-major / tap-minor order like the
// halo-patch conv kernels, fp32 scale/shift/residual/ReLU, one rounding per stored tensor); exponents as in
// bottleneck_fused.cuh.
#pragma once
#include "bottleneck_fused.cuh"

#ifndef TDET_T2_CONVERGENT
#define TDET_T2_CONVERGENT 1
#endif

namespace tdet {

// W ring: 64 KiB.  Single CTA: four 16 KiB slots (one W2 (tap, chunk) block [128][64], or half a W3 tile).  CTA pair:
// each CTA stages only ITS half of every block's N rows (cta_group::2 MMAs read both halves), so eight 8 KiB slots.
template <bool PAIR> constexpr int kT2WSlotsOf = PAIR ? 8 : 4;       // power of two
template <bool PAIR> constexpr int kT2SlotBytesOf = PAIR ? 64 * 128 : 128 * 128;
constexpr int kT2StageBytes = 2 * kPatchStageBytes;   // two 64-channel halo patches of one tile

struct T2Params {
  CUtensorMap tmap_z1;     // 4D (128, W, H, N) box (64, 10, 18, 1)
  CUtensorMap tmap_w2;     // 2D [128][9*128] box (64, 128 | pair: 64): one (tap, chunk) block (this CTA's half)
  CUtensorMap tmap_w3;     // 2D [512][128] box (64, 128 | pair: 64): one slot of a (pass, chunk) tile
  CUtensorMap tmap_res;    // 4D (512, W, H, N) box (64, 8, 16, 1)
  CUtensorMap tmap_out;    // same geometry over the block output
  int H, W, N;             // N: images
  int tiles_w, tiles_h, num_tiles;
  int x_fp16, out_fp16, res_fp16;
  int z2_scaled, out_scaled;
  const float* scale2; const float* shift2;   // bn2 (128)
  const float* scale3; const float* shift3;   // bn3 (512)
  const float* consts2; const float* consts3;
  const TensorMeta* z1_meta;
  const TensorMeta* res_meta;
  TensorMeta* out_meta;
  int res_prefetch;
  int dbg;                     // timing experiments (WRONG results): 1 = W2 blocks are not loaded, 2 = one MMA per k-block
  unsigned long long* trace;   // debugging aid (tools/trace_bottleneck_tail.py --planes 128), as FbParams::trace
};

struct T2Smem {
  static constexpr int kPatchOff = 0;
  static constexpr int kWOff = kPatchOff + kFbPatchStages * kT2StageBytes;
  static constexpr int kOutOff = kWOff + 65536;
  static constexpr int kBarOff = kOutOff + 4 * kSlabBytes;
  static constexpr int kNumBars = 44;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kParamOff = kTmemPtrOff + 16;
  static constexpr int kDynamic = kParamOff + (2 * 128 + 2 * 512) * 4;
  static_assert(kWOff % 1024 == 0 && kOutOff % 1024 == 0 && kT2StageBytes % 1024 == 0, "swizzled operands: 1024-byte alignment");
  static_assert(2 * kSlabBytes <= kT2StageBytes, "z2 (two slabs) must fit the patch stage it overwrites");
  static_assert(kDynamic <= 232448, "exceeds the 227 KiB shared memory limit");
};

// PAIR: clusters of two CTAs; the pair computes two consecutive tiles with M = 256 cta_group::2 MMAs issued by the
// leader (rank 0).  Per CTA: half the W bytes through shared memory, half the MMA instructions per tile -- the
// single-CTA kernel is bound by shared-memory bandwidth (MMA operand reads + TMA fills + epilogue ~ 1.8 MB per tile at
// 128 B/clk) and by its one issuing thread.  Patches, z2, accumulators, residual / output slabs and the epilogue stay
// per CTA; the leader's barriers collect both CTAs (TMA bytes, epilogue arrivals), its commits are multicast.
template <bool PAIR>
__global__ void __launch_bounds__(kFbThreads, 1)
bottleneck_tail2_kernel(const __grid_constant__ T2Params p) {
  using L = T2Smem;
  constexpr int kT2WSlots = kT2WSlotsOf<PAIR>;
  constexpr int kT2SlotBytes = kT2SlotBytesOf<PAIR>;
  constexpr int kNC = PAIR ? 2 : 1;            // CTAs whose arrivals / bytes the leader's barriers collect
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t s_patch = base + L::kPatchOff;
  const uint32_t s_w = base + L::kWOff;
  const uint32_t s_out = base + L::kOutOff;
  const uint32_t bar0 = base + L::kBarOff;
  auto p_full = [&](int s) { return bar0 + 8u * (0 + s); };
  auto p_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (12 + s); };
  auto d1_full = [&](int a) { return bar0 + 8u * (20 + a); };
  auto d1_empty = [&](int a) { return bar0 + 8u * (22 + a); };
  const uint32_t z2_full = bar0 + 8u * 24;
  const uint32_t d2_full = bar0 + 8u * 25;
  const uint32_t d2_empty = bar0 + 8u * 26;
  auto r_full = [&](int j) { return bar0 + 8u * (28 + j); };
  auto r_free = [&](int j) { return bar0 + 8u * (32 + j); };
  auto o_written = [&](int j) { return bar0 + 8u * (36 + j); };
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  // the barriers the MMA thread waits on live in the leader
  auto arrive_leader = [&](uint32_t bar) {
    if (PAIR) mbar_arrive_cluster(mapa_shared(bar, 0)); else mbar_arrive(bar);
  };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::kTmemPtrOff);
  float* s_par = reinterpret_cast<float*>(smem + L::kParamOff);  // sc2[128] sh2[128] sc3[512] sh3[512]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kD1 = 0, kD2 = 256;
  int trace_n = 0;
  auto trace = [&](int code) {
    // rows 0-4: lane 0 of warps 0-4 (patch producer, MMA issuer, residual producer, store issuer, first epilogue warp);
    // row 5: the W-ring producer (warp 0, lane 1)
    if (p.trace != nullptr && blockIdx.x == 0 && ((warp < 5 && lane == 0) || (warp == 0 && lane == 1)) && trace_n < kFbTraceLen)
      p.trace[(warp == 0 && lane == 1 ? 5 : warp) * kFbTraceLen + trace_n++] =
          (static_cast<unsigned long long>(clock64()) << 8) | static_cast<unsigned>(code);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_z1);
    tma_prefetch_desc(&p.tmap_w2);
    tma_prefetch_desc(&p.tmap_w3);
    tma_prefetch_desc(&p.tmap_res);
    tma_prefetch_desc(&p.tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(p_full(s), 1);
      mbar_init(p_empty(s), 1);
      mbar_init(d1_full(s), 1);
      mbar_init(d1_empty(s), kFbEpiWarps * kNC);
    }
    for (int s = 0; s < kT2WSlots; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    mbar_init(z2_full, kFbEpiWarps * kNC);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, kFbEpiWarps * kNC);
    for (int j = 0; j < 4; ++j) {
      mbar_init(r_full(j), 1);
      mbar_init(r_free(j), 1);
      mbar_init(o_written(j), 4);
    }
    fence_mbar_init();
  }
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
  grid_dependency_wait();

  // schedule: CTA (pair) u takes tiles (pair-tiles) u, u + step, ...; rank r of a pair computes tile 2 * pair-tile + r
  // (an odd tile count leaves the last pair's second CTA a tile beyond the batch: its loads are zero-filled, its
  // stores clipped away by the tensor maps)
  const int units = PAIR ? (p.num_tiles + 1) >> 1 : p.num_tiles;
  const int tile0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int nt = tile0 < units ? (units - tile0 + step - 1) / step : 0;  // tiles of this CTA
  auto tile_origin = [&](int k, int& w0, int& h0, int& img) {
    const int t = PAIR ? 2 * (tile0 + k * step) + static_cast<int>(cta_rank) : tile0 + k * step;
    const int tw = t % p.tiles_w;
    const int r = t / p.tiles_w;
    w0 = tw * kPatchBW;
    h0 = (r % p.tiles_h) * kPatchBH;
    img = r / p.tiles_h;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- halo patches (two 64-channel chunks per tile)
      for (int k = 0; k < nt; ++k) {
        int w0, h0, img;
        tile_origin(k, w0, h0, img);
        const int ps = k & 1;
        if (k + 1 < nt) {   // the next tile's patch -> L2 (its stage frees late: after G2(k-1, 1))
          int w1, h1, img1;
          tile_origin(k + 1, w1, h1, img1);
          tma_prefetch_l2_4d(&p.tmap_z1, 0, w1 - 1, h1 - 1, img1);
          tma_prefetch_l2_4d(&p.tmap_z1, 64, w1 - 1, h1 - 1, img1);
        }
        mbar_wait(p_empty(ps), (static_cast<uint32_t>(k >> 1) & 1u) ^ 1u);
        trace(1);   // patch load issued
        if (PAIR) {
          // both CTAs' bytes are counted on the leader's barrier (the MMA thread's)
          if (cta_rank == 0) mbar_arrive_expect_tx(p_full(ps), 4 * kPatchBytes);
          const uint32_t lb = mapa_shared(p_full(ps), 0);
          tma_load_4d_pair(s_patch + ps * kT2StageBytes, &p.tmap_z1, lb, 0, w0 - 1, h0 - 1, img);
          tma_load_4d_pair(s_patch + ps * kT2StageBytes + kPatchStageBytes, &p.tmap_z1, lb, 64, w0 - 1, h0 - 1, img);
        } else {
          mbar_arrive_expect_tx(p_full(ps), 2 * kPatchBytes);
          tma_load_4d(s_patch + ps * kT2StageBytes, &p.tmap_z1, p_full(ps), 0, w0 - 1, h0 - 1, img);
          tma_load_4d(s_patch + ps * kT2StageBytes + kPatchStageBytes, &p.tmap_z1, p_full(ps), 64, w0 - 1, h0 - 1, img);
        }
      }
    } else if (lane == 1) {
      // ---------------------------------------------------------------- W ring, in the MMA thread's order
      uint32_t c = 0;
      auto put = [&](const CUtensorMap* tm, int col, int rowc) {
        const uint32_t s = c & (kT2WSlots - 1);
        mbar_wait(w_empty(s), ((c / kT2WSlots) & 1u) ^ 1u);
        trace(2);   // W slot requested
        if (!PAIR && (p.dbg & 1) && tm == &p.tmap_w2) {
          mbar_arrive(w_full(s));
        } else if (PAIR) {
          if (cta_rank == 0) mbar_arrive_expect_tx(w_full(s), 2 * kT2SlotBytes);
          tma_load_2d_pair(s_w + s * kT2SlotBytes, tm, mapa_shared(w_full(s), 0), col, rowc);
        } else {
          mbar_arrive_expect_tx(w_full(s), kT2SlotBytes);
          tma_load_2d(s_w + s * kT2SlotBytes, tm, w_full(s), col, rowc);
        }
        ++c;
      };
      // a slot holds kSlotRows weight rows; this CTA stages rows [rank * N / 2, + N / 2) of every block / tile
      constexpr int kSlotRows = kT2SlotBytes / 128;
      auto put_w2 = [&](int b0, int b1) {
        for (int b = b0; b < b1; ++b) put(&p.tmap_w2, (b % 9) * 128 + (b / 9) * 64, static_cast<int>(cta_rank) * 64);
      };
      auto put_w3 = [&](int h) {
        for (int kc = 0; kc < 2; ++kc)
          for (int half = 0; half < 2; ++half)
            put(&p.tmap_w3, kc * 64, h * 256 + static_cast<int>(cta_rank) * 128 + half * kSlotRows);
      };
      auto pad_even = [&]() { if (c & 1u) put(&p.tmap_w2, 0, 0); };
      if (nt > 0) put_w2(0, 18);
      if (nt > 1) put_w2(0, 9);
      pad_even();
      for (int k = 0; k < nt; ++k) {
        put_w3(0);
        if (k + 1 < nt) put_w2(9, 15);
        put_w3(1);
        if (k + 1 < nt) put_w2(15, 18);
        if (k + 2 < nt) put_w2(0, 9);
        pad_even();
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair: the leader's, for both CTAs)
    // (one thread; a warp-convergent loop with the descriptors in uniform registers measured 10 % SLOWER)
    // kConv: the whole warp runs the loop convergently, the tcgen05 instructions are predicated on an elected lane
    constexpr bool kConv = !PAIR && TDET_T2_CONVERGENT != 0;
    if ((kConv || lane == 0) && cta_rank == 0) {
      const uint32_t fx = p.x_fp16 ? kFmtF16 : kFmtBF16;
      const uint32_t idesc1 = make_idesc_f16kind(PAIR ? 2 * kBM : kBM, 128, fx, fx);
      const uint32_t idesc2 = make_idesc_f16kind(PAIR ? 2 * kBM : kBM, 256, fx, fx);
      const uint64_t d_patch0 = make_smem_desc_sw128_sbo(s_patch, kPatchPW * 128);
      const uint64_t d_z2_0 = make_smem_desc_sw128(s_patch);
      const uint64_t d_w0 = make_smem_desc_sw128(s_w);
      uint32_t c = 0;        // ring counter (slot c % slots, fill number c / slots)
      uint64_t dp = 0;       // A descriptor base of the tile whose k-blocks are being issued
      uint32_t d1 = 0;       // its accumulator
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
        if (PAIR) umma_bf16_ss_pair(d, da, db, idesc, acc);
        else if (kConv) umma_bf16_ss_elect(d, da, db, idesc, acc);
        else umma_bf16_ss(d, da, db, idesc, acc);
      };
      auto commit = [&](uint32_t bar) {   // (pair: arrives on the barrier of BOTH CTAs)
        if (PAIR) umma_commit_pair(bar);
        else if (kConv) umma_commit_elect(bar);
        else umma_commit(bar);
      };
      auto g1_begin = [&](int k) {
        const int a = k & 1;
        const uint32_t ph = static_cast<uint32_t>(k >> 1) & 1u;
        trace(9);
        mbar_wait(p_full(a), ph);
        trace(10);  // patch landed
        if (PAIR) mbar_wait_cluster(d1_empty(a), ph ^ 1u); else mbar_wait(d1_empty(a), ph ^ 1u);
        trace(11);  // D1 free
        tc_fence_after();
      };
      auto g1_select = [&](int k) {
        const int a = k & 1;
        dp = d_patch0 + static_cast<uint32_t>(a * (kT2StageBytes >> 4));
        d1 = tmem_base + kD1 + static_cast<uint32_t>(a * 128);
      };
      auto g1_block = [&](int b) {   // `b` is a compile-time constant in the unrolled callers
        const int kc = b / 9, tap = b % 9;
        const uint64_t da = dp + static_cast<uint32_t>(kc * (kPatchStageBytes >> 4) + ((tap / 3) * kPatchPW + tap % 3) * 8);
        const uint32_t s = c & (kT2WSlots - 1);
        trace(13);  // k-block: waiting for its W slot
        mbar_wait(w_full(s), (c / kT2WSlots) & 1u);
        trace(12);  // k-block issue
        tc_fence_after();
        const uint64_t db = d_w0 + s * (kT2SlotBytes >> 4);
        if (p.dbg & 2) {
          mma(d1, da, db, idesc1, b != 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) mma(d1, da + 2u * q, db + 2u * q, idesc1, (b | q) != 0 ? 1u : 0u);
        }
        commit(w_empty(s));
        ++c;
      };
      auto g1_end = [&](int k) { commit(d1_full(k & 1)); };
      auto pad_even = [&]() {
        if (c & 1u) {
          const uint32_t s = c & (kT2WSlots - 1);
          mbar_wait(w_full(s), (c / kT2WSlots) & 1u);
          commit(w_empty(s));
          ++c;
        }
      };
      auto g2 = [&](int k, int h) {
        const uint32_t q2 = static_cast<uint32_t>(2 * k + h);
        trace(20);  // G2: start waiting
        if (h == 0) {
          if (PAIR) mbar_wait_cluster(z2_full, static_cast<uint32_t>(k) & 1u); else mbar_wait(z2_full, static_cast<uint32_t>(k) & 1u);
        }
        trace(21);  // z2 ready
        if (PAIR) mbar_wait_cluster(d2_empty, (q2 & 1u) ^ 1u); else mbar_wait(d2_empty, (q2 & 1u) ^ 1u);
        trace(22);  // D2 free
        tc_fence_after();
        const uint64_t dz = d_z2_0 + static_cast<uint32_t>((k & 1) * (kT2StageBytes >> 4));
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          const uint32_t s = c & (kT2WSlots - 1);   // even: the tile (this CTA's half of it) is slots s, s + 1
          mbar_wait(w_full(s), (c / kT2WSlots) & 1u);
          mbar_wait(w_full(s + 1), (c / kT2WSlots) & 1u);
          trace(23);  // W3 tile landed -> issue
          tc_fence_after();
          const uint64_t da = dz + static_cast<uint32_t>(kc * (kSlabBytes >> 4));
          const uint64_t db = d_w0 + s * (kT2SlotBytes >> 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) mma(tmem_base + kD2, da + 2u * q, db + 2u * q, idesc2, (kc | q) != 0 ? 1u : 0u);
          commit(w_empty(s));
          commit(w_empty(s + 1));
          c += 2;
        }
        commit(d2_full);
        if (h == 1) commit(p_empty(k & 1));   // z2(k) consumed: the patch stage may be refilled
      };
      if (nt > 0) {
        g1_begin(0);
        g1_select(0);
#pragma unroll
        for (int b = 0; b < 18; ++b) g1_block(b);
        g1_end(0);
      }
      if (nt > 1) {
        g1_begin(1);
        g1_select(1);
#pragma unroll
        for (int b = 0; b < 9; ++b) g1_block(b);
      }
      pad_even();
      for (int k = 0; k < nt; ++k) {
        g2(k, 0);
        if (k + 1 < nt) {
          g1_select(k + 1);
#pragma unroll
          for (int b = 9; b < 15; ++b) g1_block(b);
        }
        g2(k, 1);
        if (k + 1 < nt) {
#pragma unroll
          for (int b = 15; b < 18; ++b) g1_block(b);
          g1_end(k + 1);
        }
        if (k + 2 < nt) {
          g1_begin(k + 2);
          g1_select(k + 2);
#pragma unroll
          for (int b = 0; b < 9; ++b) g1_block(b);
        }
        pad_even();
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual producer: pass q = 2 k + h of the
    // CTA loads channels [256 h + 64 j, + 64) of tile k into output slab j
    if (lane == 0) {
      for (int q = 0; q < 2 * nt; ++q) {
        int w0, h0, img;
        tile_origin(q >> 1, w0, h0, img);
        const int hc = (q & 1) * 256;
        const uint32_t ph = static_cast<uint32_t>(q) & 1u;
        if (p.res_prefetch && q + 1 < 2 * nt) {
          int w1, h1, img1;
          tile_origin((q + 1) >> 1, w1, h1, img1);
          const int hc1 = ((q + 1) & 1) * 256;
          for (int j = 0; j < 4; ++j) tma_prefetch_l2_4d(&p.tmap_res, hc1 + j * 64, w1, h1, img1);
        }
        for (int j = 0; j < 4; ++j) {
          mbar_wait(r_free(j), ph ^ 1u);
          trace(40 + j);  // residual slab j requested
          mbar_arrive_expect_tx(r_full(j), kSlabBytes);
          tma_load_4d(s_out + j * kSlabBytes, &p.tmap_res, r_full(j), hc + j * 64, w0, h0, img);
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------------------------------------------ store issuer
    if (lane == 0) {
      for (int q = 0; q < 2 * nt; ++q) {
        int w0, h0, img;
        tile_origin(q >> 1, w0, h0, img);
        const int hc = (q & 1) * 256;
        const uint32_t ph = static_cast<uint32_t>(q) & 1u;
        for (int j = 0; j < 4; ++j) {
          mbar_wait(o_written(j), ph);
          trace(50 + j);  // output slab j written -> store
          asm volatile(
              "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
              ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)), "r"(s_out + j * kSlabBytes), "r"(hc + j * 64), "r"(w0), "r"(h0),
              "r"(img)
              : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        mbar_arrive(r_free(0));
        asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        mbar_arrive(r_free(1));
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        mbar_arrive(r_free(2));
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        trace(57);  // all four stores have read their slabs
        mbar_arrive(r_free(3));
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 4..19)
    const int quad = warp & 3;
    const int grp = (warp - 4) >> 2;  // 0..3
    const int row = quad * 32 + lane;
    const int etid = threadIdx.x - 128;  // 0..511
    const bool x_fp16 = p.x_fp16 != 0, out_fp16 = p.out_fp16 != 0, res_fp16 = p.res_fp16 != 0;
    const int e_z1 = p.z1_meta ? p.z1_meta->e : 0;
    const int e_res = p.res_meta ? p.res_meta->e : 0;
    const float a_z1 = p.z1_meta ? __uint_as_float(p.z1_meta->amax_bits) : 0.0f;
    const float a_res = p.res_meta ? __uint_as_float(p.res_meta->amax_bits) : 0.0f;
    auto exp_of = [](float bound) {
      int e = 0;
      if (bound > 0.0f && bound < 3.0e38f) e = ilogbf(bound) - 14;
      return max(-100, min(100, e));
    };
    const float b2 = p.consts2 ? p.consts2[0] * a_z1 + p.consts2[1] : 0.0f;
    const float b3 = p.consts3 ? p.consts3[0] * b2 + p.consts3[1] + a_res : 0.0f;
    const int e_z2 = p.z2_scaled ? exp_of(b2) : 0;
    const int e_out = p.out_scaled ? exp_of(b3) : 0;
    if (blockIdx.x == 0 && etid == 0 && p.out_meta && p.out_scaled) p.out_meta->e = e_out;
    {
      const float m2 = ldexpf(1.0f, e_z1 - e_z2), a2 = ldexpf(1.0f, -e_z2);
      const float m3 = ldexpf(1.0f, e_z2 - e_out), a3 = ldexpf(1.0f, -e_out);
      for (int i = etid; i < 128; i += kFbEpiThreads) {
        s_par[i] = (p.scale2 ? __ldg(p.scale2 + i) : 1.0f) * m2;
        s_par[128 + i] = (p.shift2 ? __ldg(p.shift2 + i) : 0.0f) * a2;
      }
      for (int i = etid; i < 512; i += kFbEpiThreads) {
        s_par[256 + i] = (p.scale3 ? __ldg(p.scale3 + i) : 1.0f) * m3;
        s_par[768 + i] = (p.shift3 ? __ldg(p.shift3 + i) : 0.0f) * a3;
      }
      named_bar_sync(1, kFbEpiThreads);
    }
    const float mul_res = ldexpf(1.0f, e_res - e_out);
    const float* sc2 = s_par;
    const float* sh2 = s_par + 128;
    const float* sc3 = s_par + 256;
    const float* sh3 = s_par + 768;
    float amax_out = 0.0f;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);

    // E1: columns [32 grp, 32 grp + 32) of D1 -> bn2 + ReLU -> 16-bit -> z2 slab (grp >> 1), chunks 4 (grp & 1) ..
    auto e1 = [&](int k) {
      const int a = k & 1;
      trace(60);  // E1: waiting for D1
      mbar_wait(d1_full(a), static_cast<uint32_t>(k >> 1) & 1u);
      trace(61);  // D1 ready
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(lane_base + kD1 + static_cast<uint32_t>(a * 128 + grp * 32), v);
      tmem_ld_wait();
      const uint32_t rbase = s_patch + static_cast<uint32_t>(a * kT2StageBytes + (grp >> 1) * kSlabBytes) + row * 128;
      const int c0 = grp * 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; e += 4) {
          const float4 s4 = *reinterpret_cast<const float4*>(sc2 + c0 + 8 * j + e);
          const float4 h4 = *reinterpret_cast<const float4*>(sh2 + c0 + 8 * j + e);
          x[e + 0] = fmaxf(fmaf(__uint_as_float(v[8 * j + e + 0]), s4.x, h4.x), 0.0f);
          x[e + 1] = fmaxf(fmaf(__uint_as_float(v[8 * j + e + 1]), s4.y, h4.y), 0.0f);
          x[e + 2] = fmaxf(fmaf(__uint_as_float(v[8 * j + e + 2]), s4.z, h4.z), 0.0f);
          x[e + 3] = fmaxf(fmaf(__uint_as_float(v[8 * j + e + 3]), s4.w, h4.w), 0.0f);
        }
        uint4 o;
        o.x = pack16x2(x[0], x[1], x_fp16);
        o.y = pack16x2(x[2], x[3], x_fp16);
        o.z = pack16x2(x[4], x[5], x_fp16);
        o.w = pack16x2(x[6], x[7], x_fp16);
        const uint32_t ad = rbase + (((((grp & 1) << 2) + j) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                     : "memory");
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        arrive_leader(d1_empty(a));
        arrive_leader(z2_full);
      }
      trace(62);  // z2 published
    };

    // E2 of pass q: out = relu(D2 * scale3 + shift3 + residual), in place over residual slab `grp`
    auto e2 = [&](int q, bool valid) {
      const uint32_t ph = static_cast<uint32_t>(q) & 1u;
      const int j = grp;
      trace(70);  // E2: waiting for D2
      mbar_wait(d2_full, ph);
      trace(71);  // D2 ready
      tc_fence_after();
      mbar_wait(r_full(j), ph);
      trace(72);  // residual slab landed
      const uint32_t rbase = s_out + j * kSlabBytes + row * 128;
      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(lane_base + kD2 + static_cast<uint32_t>(j * 64), va);
      tmem_ld_32x32b_x32(lane_base + kD2 + static_cast<uint32_t>(j * 64 + 32), vb);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t (&v)[32] = half ? vb : va;
        uint4 rr[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t ad = rbase + ((((half << 2) | c) ^ (row & 7)) << 4);
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(rr[c].x), "=r"(rr[c].y), "=r"(rr[c].z), "=r"(rr[c].w)
                       : "r"(ad));
        }
        if (half == 0) tmem_ld_wait();
        const int cb = (q & 1) * 256 + j * 64 + half * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t w4[4] = {rr[c].x, rr[c].y, rr[c].z, rr[c].w};
          float x[8];
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            float r0, r1, r2, r3;
            unpack16x2(w4[e], res_fp16, r0, r1);
            unpack16x2(w4[e + 1], res_fp16, r2, r3);
            const float4 s4 = *reinterpret_cast<const float4*>(sc3 + cb + 8 * c + 2 * e);
            const float4 h4 = *reinterpret_cast<const float4*>(sh3 + cb + 8 * c + 2 * e);
            x[2 * e + 0] = fmaxf(fmaf(r0, mul_res, fmaf(__uint_as_float(v[8 * c + 2 * e + 0]), s4.x, h4.x)), 0.0f);
            x[2 * e + 1] = fmaxf(fmaf(r1, mul_res, fmaf(__uint_as_float(v[8 * c + 2 * e + 1]), s4.y, h4.y)), 0.0f);
            x[2 * e + 2] = fmaxf(fmaf(r2, mul_res, fmaf(__uint_as_float(v[8 * c + 2 * e + 2]), s4.z, h4.z)), 0.0f);
            x[2 * e + 3] = fmaxf(fmaf(r3, mul_res, fmaf(__uint_as_float(v[8 * c + 2 * e + 3]), s4.w, h4.w)), 0.0f);
          }
          if (valid) {
#pragma unroll
            for (int e = 0; e < 8; ++e) amax_out = fmaxf(amax_out, x[e]);
          }
          uint4 o;
          o.x = pack16x2(x[0], x[1], out_fp16);
          o.y = pack16x2(x[2], x[3], out_fp16);
          o.z = pack16x2(x[4], x[5], out_fp16);
          o.w = pack16x2(x[6], x[7], out_fp16);
          const uint32_t ad = rbase + ((((half << 2) | c) ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                       : "memory");
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(o_written(j));
        arrive_leader(d2_empty);
      }
      trace(74);  // output slab published
    };

    if (nt > 0) e1(0);
    for (int k = 0; k < nt; ++k) {
      int w0, h0, img;
      tile_origin(k, w0, h0, img);
      const bool valid = (h0 + (row >> 3) < p.H) && (w0 + (row & 7) < p.W) && img < p.N;
      e2(2 * k, valid);
      e2(2 * k + 1, valid);
      if (k + 1 < nt) e1(k + 1);
    }
    if (p.out_meta) {
      float a = amax_out;
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

}  // namespace tdet
