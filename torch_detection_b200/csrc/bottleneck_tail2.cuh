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
// What no longer fits at 128 planes: W2 is 288 KiB and W3 128 KiB, so BOTH stream through one ring of shared-memory
// slots in exactly the order the tensor pipe consumes them (L2 hits), and the 512 output channels are produced in FOUR
// passes of 128 columns.  The kernel runs as CTA PAIRS (clusters of two, cta_group::2): the pair computes two
// consecutive 128-pixel tiles with M = 256 MMAs issued by the leader, each CTA stages only ITS half of every weight
// block (8 KiB slots) -- that is what makes room for a z2 buffer of its own.
//
// Per CTA and tile of 8 x 16 output pixels (128 GEMM rows):
//   G1  D1[128 x 128] = patch(z1)[128 x 2*9*64] * W2^T    two 64-channel halo patches ((8+2) x (16+2) TMA boxes), nine
//                                                          shifted descriptors each: 18 k-blocks of four N = 128 MMAs
//   G2  D2[128 x 128] = z2[128 x 128] * W3[h]^T, h = 0..3  two k-blocks of four N = 128 MMAs per pass
// TMEM: D1 double-buffered (2 x 128 columns) + D2 double-buffered (2 x 128) = 512 columns.
//
// How the first version (z2 written over the consumed patch stage, one 256-column D2, one set of output slabs; 205-240
// us per block against 217-225 us unfused) was bound, from its cycle trace: NOT by the W ring and NOT by the G1 MMAs
// (skipping the W2 loads and three of every four MMAs changed the tile period by 10 %), but by a dependency chain --
// the patch of tile k+2 could only be requested after G2(k) had read z2(k) out of the same stage (3400 cycles to land),
// and each G2 pass had to wait for the previous pass's epilogue to drain the one D2 buffer (2 x (2300 + 3700) cycles
// per tile).  Hence: z2 in its own buffer (a patch stage is free as soon as G1 has read it: the next-but-one patch is
// requested a whole tile ahead), D2 and the output slabs double-buffered (pass q+1 is contracted, and its residual
// lands, while pass q is converted and stored).
//
// Issue order of the one tcgen05 thread (k-blocks b = chunk * 9 + tap of G1):
//   G1(0) | b 0-8 of G1(1) | { G2(k,0) | b 9-11 of G1(k+1) | G2(k,1) | b 12-14 | G2(k,2) | b 15-17 | G2(k,3) | b 0-8 of G1(k+2) }
//
// Warp roles (20 warps): warp 0 lane 0 patches, lane 1 the W ring; warp 1 MMA issuer (leader only); warp 2 residual producer (into
// the output slabs; the residual two passes ahead is prefetched into L2); warp 3 store issuer; warps 4-19 epilogue in
// four groups: E1 (D1 -> z2, 32 columns per group), E2 (32 columns of D2 + residual -> half an output slab, in place).
// Status (round 2, final build): correct and bit-identical, but 265 us per block in situ against 88 + 127 us for the
// unfused pair, so the plan compiler leaves it off (TDET_FUSE_TAIL2=1 enables).  TDET_T2_DBG and a seven-slot ring
// variant showed that neither the weight ring nor the G1 MMAs bound it: each epilogue warp runs five serial 32-column
// conversion steps per tile (E2 x 4 + E1) of 2 400 - 4 000 cycles each.
// Numerics are those of the unfused launches (fp32 accumulation, k-blocks in chunk-major / tap-minor order like the
// halo-patch conv kernels, fp32 scale/shift/residual/ReLU, one rounding per stored tensor); exponents as in
// bottleneck_fused.cuh.
#pragma once
#include "bottleneck_fused.cuh"

namespace tdet {

constexpr int kT2WSlots = 4;                          // power of two
constexpr int kT2SlotBytes = 64 * 128;                // this CTA's half (64 rows) of a W2 (tap, chunk) block / W3 tile
constexpr int kT2StageBytes = 2 * kPatchStageBytes;   // two 64-channel halo patches of one tile
// 21 warps: the W-ring producer has a warp of its own (warp 20).  As lane 1 of the patch producer's warp (the layout of
// bottleneck_tail_kernel) it cost ~800 cycles per ring slot even with the loads and MMAs switched off: two lanes of
// one warp spinning on different mbarriers take turns, and the patch lane's try_wait -- it waits a whole tile --
// suspends the warp for the hardware's time limit again and again.
constexpr int kT2Threads = 672;

struct T2Params {
  CUtensorMap tmap_z1;     // 4D (128, W, H, N) box (64, 10, 18, 1)
  CUtensorMap tmap_w2;     // 2D [128][9*128] box (64, 64): this CTA's half of one (tap, chunk) block
  CUtensorMap tmap_w3;     // 2D [512][128] box (64, 64): this CTA's half of one (pass, chunk) tile
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
  int dbg;                     // timing experiments (WRONG results): 1 = W2 blocks are not loaded, 2 = one MMA per k-block,
                               // 4 = G1 without the W ring (no waits, no commits per k-block)
  unsigned long long* trace;   // debugging aid (tools/trace_bottleneck_tail.py --planes 128), as FbParams::trace
};

struct T2Smem {
  static constexpr int kPatchOff = 0;
  static constexpr int kZ2Off = kPatchOff + kFbPatchStages * kT2StageBytes;
  static constexpr int kWOff = kZ2Off + 2 * kSlabBytes;
  static constexpr int kOutOff = kWOff + kT2WSlots * kT2SlotBytes;
  static constexpr int kBarOff = kOutOff + 4 * kSlabBytes;
  static constexpr int kNumBars = 36;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kParamOff = kTmemPtrOff + 16;
  static constexpr int kDynamic = kParamOff + (2 * 128 + 2 * 512) * 4;
  static_assert(kZ2Off % 1024 == 0 && kWOff % 1024 == 0 && kOutOff % 1024 == 0 && kT2StageBytes % 1024 == 0,
                "swizzled operands: 1024-byte alignment");
  static_assert(kDynamic <= 232448, "exceeds the 227 KiB shared memory limit");
};

__global__ void __launch_bounds__(kT2Threads, 1)
bottleneck_tail2_kernel(const __grid_constant__ T2Params p) {
  using L = T2Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t s_patch = base + L::kPatchOff;
  const uint32_t s_z2 = base + L::kZ2Off;
  const uint32_t s_w = base + L::kWOff;
  const uint32_t s_out = base + L::kOutOff;
  const uint32_t bar0 = base + L::kBarOff;
  auto p_full = [&](int s) { return bar0 + 8u * (0 + s); };
  auto p_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (8 + s); };
  auto d1_full = [&](int a) { return bar0 + 8u * (12 + a); };
  auto d1_empty = [&](int a) { return bar0 + 8u * (14 + a); };
  const uint32_t z2_full = bar0 + 8u * 16;
  const uint32_t z2_free = bar0 + 8u * 17;
  auto d2_full = [&](int a) { return bar0 + 8u * (18 + a); };
  auto d2_empty = [&](int a) { return bar0 + 8u * (20 + a); };
  auto r_full = [&](int j) { return bar0 + 8u * (22 + j); };
  auto r_free = [&](int j) { return bar0 + 8u * (26 + j); };
  auto o_written = [&](int j) { return bar0 + 8u * (30 + j); };
  const uint32_t cta_rank = cluster_ctarank();
  // the barriers the MMA thread waits on live in the leader (rank 0): both CTAs' epilogue warps arrive there
  auto arrive_leader = [&](uint32_t bar) { mbar_arrive_cluster(mapa_shared(bar, 0)); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::kTmemPtrOff);
  float* s_par = reinterpret_cast<float*>(smem + L::kParamOff);  // sc2[128] sh2[128] sc3[512] sh3[512]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kD1 = 0, kD2 = 256;
  int trace_n = 0;
  auto trace = [&](int code) {
    // rows 0-4: lane 0 of warps 0-4 (patch producer, MMA issuer, residual producer, store issuer, first epilogue warp);
    // row 5: the W-ring producer (warp 20)
    if (p.trace != nullptr && blockIdx.x == 0 && (warp < 5 || warp == 20) && lane == 0 && trace_n < kFbTraceLen)
      p.trace[(warp == 20 ? 5 : warp) * kFbTraceLen + trace_n++] =
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
      mbar_init(d1_empty(s), 2 * kFbEpiWarps);   // both CTAs' epilogue warps
      mbar_init(d2_full(s), 1);
      mbar_init(d2_empty(s), 2 * kFbEpiWarps);
    }
    for (int s = 0; s < kT2WSlots; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    mbar_init(z2_full, 2 * kFbEpiWarps);
    mbar_init(z2_free, 1);
    for (int j = 0; j < 4; ++j) {
      mbar_init(r_full(j), 1);
      mbar_init(r_free(j), 1);
      mbar_init(o_written(j), 8);   // the two groups (eight warps) that convert the halves of slab j
    }
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // schedule: pair u takes pair-tiles u, u + step, ...; rank r computes tile 2 * pair-tile + r (an odd tile count leaves
  // the last pair's second CTA a tile beyond the batch: its loads are zero-filled, its stores clipped by the tensor maps)
  const int units = (p.num_tiles + 1) >> 1;
  const int tile0 = static_cast<int>(blockIdx.x >> 1);
  const int step = static_cast<int>(gridDim.x >> 1);
  const int nt = tile0 < units ? (units - tile0 + step - 1) / step : 0;  // tiles of this CTA
  auto tile_origin = [&](int k, int& w0, int& h0, int& img) {
    const int t = 2 * (tile0 + k * step) + static_cast<int>(cta_rank);
    const int tw = t % p.tiles_w;
    const int r = t / p.tiles_w;
    w0 = tw * kPatchBW;
    h0 = (r % p.tiles_h) * kPatchBH;
    img = r / p.tiles_h;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- halo patches (two 64-channel chunks per tile);
      // stage k & 1 is free as soon as G1(k - 2) has read it
      grid_dependency_wait();
      for (int k = 0; k < nt; ++k) {
        int w0, h0, img;
        tile_origin(k, w0, h0, img);
        const int ps = k & 1;
        mbar_wait(p_empty(ps), (static_cast<uint32_t>(k >> 1) & 1u) ^ 1u);
        trace(1);   // patch load issued
        // both CTAs' bytes are counted on the leader's barrier (the MMA thread's)
        if (cta_rank == 0) mbar_arrive_expect_tx(p_full(ps), 4 * kPatchBytes);
        const uint32_t lb = mapa_shared(p_full(ps), 0);
        tma_load_4d_pair(s_patch + ps * kT2StageBytes, &p.tmap_z1, lb, 0, w0 - 1, h0 - 1, img);
        tma_load_4d_pair(s_patch + ps * kT2StageBytes + kPatchStageBytes, &p.tmap_z1, lb, 64, w0 - 1, h0 - 1, img);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (the leader's, for both CTAs)
    // (one thread; a warp-convergent loop with the descriptors in uniform registers measured no faster)
    if (lane == 0 && cta_rank == 0) {
      const uint32_t fx = p.x_fp16 ? kFmtF16 : kFmtBF16;
      const uint32_t idesc = make_idesc_f16kind(2 * kBM, 128, fx, fx);
      const uint64_t d_patch0 = make_smem_desc_sw128_sbo(s_patch, kPatchPW * 128);
      const uint64_t d_z2 = make_smem_desc_sw128(s_z2);
      const uint64_t d_w0 = make_smem_desc_sw128(s_w);
      uint32_t c = 0;        // ring counter (slot c % slots, fill number c / slots)
      uint64_t dp = 0;       // A descriptor base of the tile whose k-blocks are being issued
      uint32_t d1 = 0;       // its accumulator
      auto g1_begin = [&](int k) {
        const int a = k & 1;
        const uint32_t ph = static_cast<uint32_t>(k >> 1) & 1u;
        trace(9);
        mbar_wait(p_full(a), ph);
        trace(10);  // patch landed
        mbar_wait_cluster(d1_empty(a), ph ^ 1u);
        trace(11);  // D1 free
        tc_fence_after();
      };
      auto g1_select = [&](int k) {
        const int a = k & 1;
        dp = d_patch0 + static_cast<uint32_t>(a * (kT2StageBytes >> 4));
        d1 = tmem_base + kD1 + static_cast<uint32_t>(a * 128);
      };
      // next W slot: waits for it, returns its B descriptor
      auto w_take = [&]() -> uint64_t {
        const uint32_t s = c & (kT2WSlots - 1);
        mbar_wait(w_full(s), (c / kT2WSlots) & 1u);
        tc_fence_after();
        return d_w0 + s * (kT2SlotBytes >> 4);
      };
      auto w_done = [&]() {   // the slot's MMAs are issued: free it (in both CTAs) when they complete
        umma_commit_pair(w_empty(c & (kT2WSlots - 1)));
        ++c;
      };
      auto g1_block = [&](int b) {   // `b` is a compile-time constant in the unrolled callers
        const int kc = b / 9, tap = b % 9;
        const uint64_t da = dp + static_cast<uint32_t>(kc * (kPatchStageBytes >> 4) + ((tap / 3) * kPatchPW + tap % 3) * 8);
        trace(13);  // k-block: waiting for its W slot
        const uint64_t db = (p.dbg & 4) ? d_w0 : w_take();
        trace(12);  // k-block issue
        if (p.dbg & 2) {
          umma_bf16_ss_pair(d1, da, db, idesc, b != 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) umma_bf16_ss_pair(d1, da + 2u * q, db + 2u * q, idesc, (b | q) != 0 ? 1u : 0u);
        }
        if (!(p.dbg & 4)) w_done();
      };
      auto g1_end = [&](int k) {
        umma_commit_pair(d1_full(k & 1));
        umma_commit_pair(p_empty(k & 1));   // the patch stage has been read: request the next-but-one tile's
      };
      auto g2 = [&](int k, int h) {
        const uint32_t q2 = static_cast<uint32_t>(4 * k + h);
        const uint32_t buf = q2 & 1u;
        trace(20);  // G2: start waiting
        if (h == 0) mbar_wait_cluster(z2_full, static_cast<uint32_t>(k) & 1u);
        trace(21);  // z2 ready
        mbar_wait_cluster(d2_empty(buf), ((q2 >> 1) & 1u) ^ 1u);
        trace(22);  // D2 free
        tc_fence_after();
        const uint32_t d2 = tmem_base + kD2 + buf * 128u;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          const uint64_t db = w_take();
          const uint64_t da = d_z2 + static_cast<uint32_t>(kc * (kSlabBytes >> 4));
#pragma unroll
          for (int q = 0; q < 4; ++q) umma_bf16_ss_pair(d2, da + 2u * q, db + 2u * q, idesc, (kc | q) != 0 ? 1u : 0u);
          w_done();
        }
        umma_commit_pair(d2_full(buf));
        if (h == 3) umma_commit_pair(z2_free);   // z2(k) consumed: E1(k+1) may overwrite it
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
      for (int k = 0; k < nt; ++k) {
        const bool nx = k + 1 < nt;
        g2(k, 0);
        if (nx) {
          g1_select(k + 1);
#pragma unroll
          for (int b = 9; b < 12; ++b) g1_block(b);
        }
        g2(k, 1);
        if (nx) {
#pragma unroll
          for (int b = 12; b < 15; ++b) g1_block(b);
        }
        g2(k, 2);
        if (nx) {
#pragma unroll
          for (int b = 15; b < 18; ++b) g1_block(b);
          g1_end(k + 1);
        }
        g2(k, 3);
        if (k + 2 < nt) {
          g1_begin(k + 2);
          g1_select(k + 2);
#pragma unroll
          for (int b = 0; b < 9; ++b) g1_block(b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual producer: pass q = 4 k + h of the
    // CTA loads channels [128 h + 64 jj, + 64) of tile k into output slab 2 (q & 1) + jj
    if (lane == 0) {
      grid_dependency_wait();
      const int np = 4 * nt;
      for (int q = 0; q < np; ++q) {
        int w0, h0, img;
        tile_origin(q >> 2, w0, h0, img);
        const int hc = (q & 3) * 128;
        const int set = q & 1;
        if (p.res_prefetch && q + 2 < np) {
          int w1, h1, img1;
          tile_origin((q + 2) >> 2, w1, h1, img1);
          const int hc1 = ((q + 2) & 3) * 128;
          for (int jj = 0; jj < 2; ++jj) tma_prefetch_l2_4d(&p.tmap_res, hc1 + jj * 64, w1, h1, img1);
        }
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * set + jj;
          mbar_wait(r_free(j), (static_cast<uint32_t>(q >> 1) & 1u) ^ 1u);
          trace(40 + j);  // residual slab j requested
          mbar_arrive_expect_tx(r_full(j), kSlabBytes);
          tma_load_4d(s_out + j * kSlabBytes, &p.tmap_res, r_full(j), hc + jj * 64, w0, h0, img);
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------------------------------------------ store issuer: the two slab stores of a pass go
    // out as their slabs get published; the slabs of the PREVIOUS pass are handed back once their stores have been read
    if (lane == 0) {
      const int np = 4 * nt;
      for (int q = 0; q < np; ++q) {
        int w0, h0, img;
        tile_origin(q >> 2, w0, h0, img);
        const int hc = (q & 3) * 128;
        const int set = q & 1;
        const uint32_t ph = static_cast<uint32_t>(q >> 1) & 1u;
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * set + jj;
          mbar_wait(o_written(j), ph);
          trace(50 + j);  // output slab j written -> store
          asm volatile(
              "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
              ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)), "r"(s_out + j * kSlabBytes), "r"(hc + jj * 64), "r"(w0), "r"(h0),
              "r"(img)
              : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (q > 0) {
          asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          trace(57);  // the previous pass's stores have read their slabs
          mbar_arrive(r_free(2 * (set ^ 1)));
          mbar_arrive(r_free(2 * (set ^ 1) + 1));
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else if (warp == 20) {
    if (lane == 0) {
      // ---------------------------------------------------------------- W ring, in the MMA thread's order
      grid_dependency_wait();   // (the weights may have been re-packed by the previous launch)
      uint32_t c = 0;
      auto put = [&](const CUtensorMap* tm, int col, int rowc) {
        const uint32_t s = c & (kT2WSlots - 1);
        mbar_wait(w_empty(s), ((c / kT2WSlots) & 1u) ^ 1u);
        trace(2);   // W slot requested
        if ((p.dbg & 1) && tm == &p.tmap_w2) {
          if (cta_rank == 0) mbar_arrive(w_full(s));
        } else {
          if (cta_rank == 0) mbar_arrive_expect_tx(w_full(s), 2 * kT2SlotBytes);
          tma_load_2d_pair(s_w + s * kT2SlotBytes, tm, mapa_shared(w_full(s), 0), col, rowc);
        }
        ++c;
      };
      // this CTA stages rows [64 rank, + 64) of every 128-row block / tile
      auto put_w2 = [&](int b0, int b1) {
        if (p.dbg & 4) return;   // (timing experiment: G1 runs without ring synchronisation)
        for (int b = b0; b < b1; ++b) put(&p.tmap_w2, (b % 9) * 128 + (b / 9) * 64, static_cast<int>(cta_rank) * 64);
      };
      auto put_w3 = [&](int h) {
        for (int kc = 0; kc < 2; ++kc) put(&p.tmap_w3, kc * 64, h * 128 + static_cast<int>(cta_rank) * 64);
      };
      if (nt > 0) put_w2(0, 18);
      if (nt > 1) put_w2(0, 9);
      for (int k = 0; k < nt; ++k) {
        put_w3(0);
        if (k + 1 < nt) put_w2(9, 12);
        put_w3(1);
        if (k + 1 < nt) put_w2(12, 15);
        put_w3(2);
        if (k + 1 < nt) put_w2(15, 18);
        put_w3(3);
        if (k + 2 < nt) put_w2(0, 9);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 4..19)
    grid_dependency_wait();
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
    // this group's quarter of a 128-column accumulator: slab (grp >> 1), 16-byte chunks 4 (grp & 1) .. + 3 of a row
    const int slab_of = grp >> 1;
    const int chunk0 = (grp & 1) << 2;

    // E1: columns [32 grp, 32 grp + 32) of D1 -> bn2 + ReLU -> 16-bit -> z2 slab (grp >> 1)
    auto e1 = [&](int k) {
      const int a = k & 1;
      trace(60);  // E1: waiting for D1
      mbar_wait(d1_full(a), static_cast<uint32_t>(k >> 1) & 1u);
      trace(61);  // D1 ready
      if (k > 0) mbar_wait(z2_free, static_cast<uint32_t>(k - 1) & 1u);   // G2(k-1, *) has read z2(k-1)
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(lane_base + kD1 + static_cast<uint32_t>(a * 128 + grp * 32), v);
      tmem_ld_wait();
      const uint32_t rbase = s_z2 + static_cast<uint32_t>(slab_of * kSlabBytes) + row * 128;
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
        const uint32_t ad = rbase + (((chunk0 + j) ^ (row & 7)) << 4);
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

    // E2 of pass q = 4 k + h: out = relu(D2 * scale3 + shift3 + residual) for this group's 32 columns, in place over
    // its half of residual slab 2 (q & 1) + (grp >> 1)
    auto e2 = [&](int q, bool valid) {
      const int buf = q & 1;
      const uint32_t ph = static_cast<uint32_t>(q >> 1) & 1u;
      const int j = 2 * buf + slab_of;
      trace(70);  // E2: waiting for D2
      mbar_wait(d2_full(buf), ph);
      trace(71);  // D2 ready
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(lane_base + kD2 + static_cast<uint32_t>(buf * 128 + grp * 32), v);
      mbar_wait(r_full(j), ph);
      trace(72);  // residual slab landed
      const uint32_t rbase = s_out + j * kSlabBytes + row * 128;
      auto convert = [&](auto res16_c, auto out16_c) {   // (formats fixed at compile time, see bottleneck_fused.cuh)
        constexpr bool kR16 = decltype(res16_c)::value, kO16 = decltype(out16_c)::value;
        uint4 rr[4];
  #pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t ad = rbase + (((chunk0 + c) ^ (row & 7)) << 4);
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(rr[c].x), "=r"(rr[c].y), "=r"(rr[c].z), "=r"(rr[c].w)
                       : "r"(ad));
        }
        tmem_ld_wait();
        const int cb = (q & 3) * 128 + grp * 32;
  #pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t w4[4] = {rr[c].x, rr[c].y, rr[c].z, rr[c].w};
          float x[8];
  #pragma unroll
          for (int e = 0; e < 4; e += 2) {
            float r0, r1, r2, r3;
            unpack16x2(w4[e], kR16, r0, r1);
            unpack16x2(w4[e + 1], kR16, r2, r3);
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
          o.x = pack16x2(x[0], x[1], kO16);
          o.y = pack16x2(x[2], x[3], kO16);
          o.z = pack16x2(x[4], x[5], kO16);
          o.w = pack16x2(x[6], x[7], kO16);
          const uint32_t ad = rbase + (((chunk0 + c) ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                       : "memory");
        }
      };
      using T = std::true_type;
      using F = std::false_type;
      if (res_fp16) { if (out_fp16) convert(T{}, T{}); else convert(T{}, F{}); }
      else { if (out_fp16) convert(F{}, T{}); else convert(F{}, F{}); }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(o_written(j));
        arrive_leader(d2_empty(buf));
      }
      trace(74);  // output half-slab published
    };

    if (nt > 0) e1(0);
    for (int k = 0; k < nt; ++k) {
      int w0, h0, img;
      tile_origin(k, w0, h0, img);
      const bool valid = (h0 + (row >> 3) < p.H) && (w0 + (row & 7) < p.W) && img < p.N;
      e2(4 * k, valid);
      e2(4 * k + 1, valid);
      e2(4 * k + 2, valid);
      e2(4 * k + 3, valid);
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
  cluster_sync_all();  // (the peer may still signal this CTA's barriers)
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace tdet
