// Fused bottleneck tail for sm_100a (layer1 of the bottleneck ResNets: planes = 64, stride 1, identity shortcut):
//
//     z2   = relu(bn2(conv3x3(z1)))                          resnet.py:105-108   (z1 = this block's conv1 output)
//     out  = relu(bn3(conv1x1(z2)) + x)                      resnet.py:110-118   (x = block input, the residual)
//     z1'  = relu(bn1'(conv1x1'(out)))                       resnet.py:101-103 of the NEXT block (N3 = 64 only)
//
// in ONE persistent kernel: z2 never leaves the SM (it goes TMEM -> registers -> a swizzled shared-memory slab that IS
// the A operand of the second GEMM), `out` is written once and -- still in shared memory -- is the A operand of the
// next block's conv1, so it is not re-read either.  Per block that removes the write + read of z2 (2 x 137 MB at
// batch 16 @800x1344), the re-read of `out` (550 MB) and two launches: 2.2 GB -> 1.43 GB of HBM traffic.  Every
// kernel of layer1 runs at 75-90 % of its own HBM floor when launched alone, so traffic is what is left to remove.
//
// One CTA per SM, tiles of 8 x 16 output pixels (128 GEMM rows), three chained GEMMs per tile:
//   G1  D1[128 x 64]  = patch(z1)[128 x 9*64] * W2^T     halo-patch A operand (one (8+2) x (16+2) TMA box, nine
//                                                         shifted descriptors), W2 taps streamed through a ring
//   G2  D2[128 x 256] = z2[128 x 64] * W3^T              W3 resident
//   G3  D3[128 x 64]  = out[128 x 256] * W1'^T           W1' chunks streamed (N3 = 64)
// TMEM: D1 double-buffered (2 x 64 columns), D2 256, D3 64 = 448 of 512 columns.
//
// Warp roles (20 warps):
//   warp 0      two TMA producers on two lanes (independent thread scheduling; both mostly sleep in mbarrier waits):
//               lane 0  z1 halo patches (ring of 2), requested the moment their slot frees (an HBM load of ~3000
//                       cycles: it must not queue behind the weight taps); loads W3 once
//               lane 1  W2 taps (ring of 6 / 4, L2 hits)
//   warp 1      tcgen05.mma issuer of all three GEMMs, software-pipelined across tiles:
//                   G2(t) | taps 5-8 of G1(t+1) | taps 0-4 of G1(t+2), with G3(t) chunks slipped in as they become ready
//               so the tensor pipe runs the following tiles' 3x3 while the epilogue warps convert this tile
//   warp 2      TMA producer: the residual tile, loaded straight INTO the four output slabs (the epilogue adds in
//               place), and the W1' chunks (ring of 2)
//   warp 3      TMA store issuer: the four slab stores of a tile go out back to back (own bulk groups, overlapping
//               read-waits) and the z1' slab; a slab is freed for the next residual load once its store has been read
//               (and, N3 > 0, once G3 has consumed it: tcgen05.commit on the same barrier)
//   warps 4-19  epilogue: E1 (D1 -> z2 slab), E2 (D2 + residual -> out slabs, in place), E3 (D3 -> z1' slab); FOUR
//               groups of four warps (a warp reads its own TMEM lane quadrant): group g converts output slab g in E2
//               and columns [16 g, 16 g + 16) in E1 / E3.  The conversion is latency-bound per warp (~2400 cycles per
//               32 x 64 slab piece with two warps per scheduler), so the phase times halve against two groups
//
// Numerics are those of the unfused kernels: fp32 accumulation, fp32 scale/shift/residual/ReLU, one rounding per
// stored tensor; tensors may carry a per-tensor power-of-two exponent (TensorMeta).  The exponents of z2 / out / z1'
// come from the same rigorous bounds, chained (the true |max| of an intermediate is not known before the kernel
// ends): |z2| <= G2*amax(z1) + S2, |out| <= G3*|z2| + S3 + amax(x), |z1'| <= G1'*|out| + S1'.
#pragma once
#include <type_traits>
#include "conv_gemm.cuh"

namespace tdet {

constexpr int kFbThreads = 640;            // 20 warps: 4 role warps + 16 epilogue warps (96 registers per thread)
constexpr int kFbEpiWarps = 16;
constexpr int kFbEpiThreads = kFbEpiWarps * 32;
constexpr int kFbPatchStages = 2;
// W2 tap ring: a slot turns around in (its four MMAs) + (commit -> refill request -> ~700 cycles of L2 latency), i.e.
// ~2.5 taps of MMA time: three slots measured 980 cycles per tap instead of 560; six (four where the next conv1's
// buffers take the room) keep the tensor pipe fed
template <int N3> constexpr int kFbW2SlotsOf = N3 > 0 ? 4 : 6;
constexpr int kFbW1Slots = 2;
constexpr int kFbTapBytes = 64 * 128;      // one W2 tap / one W1' chunk: 64 rows x 64 channels x 2 B

struct FbParams {
  CUtensorMap tmap_z1;     // 4D (64, W, H, N) box (64, 10, 18, 1): halo patch of the conv2 input
  CUtensorMap tmap_w2;     // 2D [64][9*64] box (64, 64): one filter tap
  CUtensorMap tmap_w3;     // 2D [256][64] box (64, 256)
  CUtensorMap tmap_w1n;    // 2D [64][256] box (64, 64): one 64-channel chunk of the next conv1 (N3 = 64)
  CUtensorMap tmap_res;    // 4D (256, W, H, N) box (64, 8, 16, 1) over the residual (block input)
  CUtensorMap tmap_out;    // same geometry over the block output
  CUtensorMap tmap_z1o;    // 4D (64, W, H, N) box (64, 8, 16, 1) over the next block's conv1 output
  int H, W;
  int tiles_w, tiles_h, num_tiles;
  int x_fp16;              // format of z1, z2 and W2 / W3
  int out_fp16;            // format of out and W1'
  int res_fp16;
  int z1o_fp16;
  int z2_scaled, out_scaled, z1o_scaled;   // device-chosen power-of-two exponents (else exponent 0)
  const float* scale2; const float* shift2;   // bn2 (64)
  const float* scale3; const float* shift3;   // bn3 (256)
  const float* scale1n; const float* shift1n; // next block's bn1 (64)
  const float* consts2; const float* consts3; const float* consts1n;  // {G, max|shift|} of the three convs
  const TensorMeta* z1_meta;
  const TensorMeta* res_meta;
  TensorMeta* out_meta;
  TensorMeta* z1o_meta;
  // debugging aid (tools/trace_bottleneck_tail.py): lane 0 of every warp of CTA 0 appends (clock64 << 8 | event code)
  // to trace[warp * kFbTraceLen ...]; null in normal runs
  unsigned long long* trace;
  int res_prefetch;        // L2 prefetch of the next tile's residual (TDET_TAIL_PREFETCH, default 1)
};
constexpr int kFbTraceLen = 2048;  // per warp (13 warps)

// Tail-only variant (N3 == 0): W2 is RESIDENT (nine taps, 72 KiB, loaded once per CTA) -- no barrier wait and no commit
// per tap in the issue loop -- and z2 has no slab of its own: E1(k) writes it into the halo-patch stage that G1(k) has
// just finished reading (the stage is handed back to the patch producer after G2(k) instead of after G1(k)).
template <int N3>
struct FbSmem {
  static constexpr bool kW2Resident = N3 == 0;
  static constexpr int kPatchOff = 0;
  static constexpr int kW2Off = kPatchOff + kFbPatchStages * kPatchStageBytes;
  static constexpr int kW3Off = kW2Off + (kW2Resident ? 9 : kFbW2SlotsOf<N3>) * kFbTapBytes;
  static constexpr int kW1Off = kW3Off + 256 * 128;
  static constexpr int kZ2Off = kW1Off + (N3 > 0 ? kFbW1Slots * kFbTapBytes : 0);
  static constexpr int kZ1oOff = kZ2Off + (kW2Resident ? 0 : kSlabBytes);
  static constexpr int kOutOff = kZ1oOff + (N3 > 0 ? kSlabBytes : 0);
  static constexpr int kBarOff = kOutOff + 4 * kSlabBytes;
  static constexpr int kNumBars = 44;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kParamOff = kTmemPtrOff + 16;
  static constexpr int kDynamic = kParamOff + (2 * 64 + 2 * 256 + 2 * 64) * 4;
  static_assert(kW2Off % 1024 == 0 && kW3Off % 1024 == 0 && kW1Off % 1024 == 0 && kZ2Off % 1024 == 0 &&
                kZ1oOff % 1024 == 0 && kOutOff % 1024 == 0, "swizzled operands need 1024-byte alignment");
  static_assert(kDynamic <= 232448, "exceeds the 227 KiB shared memory limit");
};

__device__ __forceinline__ bool fb_try(uint32_t bar, uint32_t parity) { return mbar_try_wait(bar, parity); }

template <int N3>
__global__ void __launch_bounds__(kFbThreads, 1)
bottleneck_tail_kernel(const __grid_constant__ FbParams p) {
  static_assert(N3 == 0 || N3 == 64, "next-conv1 fusion: 64 output channels (layer1) or none");
  using L = FbSmem<N3>;
  constexpr int kW2Slots = kFbW2SlotsOf<N3>;
  constexpr bool kW2Res = L::kW2Resident;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t s_patch = base + L::kPatchOff;
  const uint32_t s_w2 = base + L::kW2Off;
  const uint32_t s_w3 = base + L::kW3Off;
  const uint32_t s_w1 = base + L::kW1Off;
  const uint32_t s_z2 = base + L::kZ2Off;
  // z2 of this CTA's k-th tile (see FbSmem)
  auto z2_of = [&](int k) { return kW2Res ? s_patch + static_cast<uint32_t>((k & 1) * kPatchStageBytes) : s_z2; };
  const uint32_t s_z1o = base + L::kZ1oOff;
  const uint32_t s_out = base + L::kOutOff;
  const uint32_t bar0 = base + L::kBarOff;
  // barrier map
  auto p_full = [&](int s) { return bar0 + 8u * (0 + s); };
  auto p_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto w2_full = [&](int s) { return bar0 + 8u * (4 + s); };    // up to 6 slots
  auto w2_empty = [&](int s) { return bar0 + 8u * (10 + s); };
  auto w1_full = [&](int s) { return bar0 + 8u * (16 + s); };
  auto w1_empty = [&](int s) { return bar0 + 8u * (18 + s); };
  const uint32_t w3_full = bar0 + 8u * 20;
  auto d1_full = [&](int a) { return bar0 + 8u * (21 + a); };
  auto d1_empty = [&](int a) { return bar0 + 8u * (23 + a); };
  const uint32_t z2_full = bar0 + 8u * 25;
  const uint32_t d2_full = bar0 + 8u * 26;
  const uint32_t d2_empty = bar0 + 8u * 27;
  auto r_full = [&](int j) { return bar0 + 8u * (28 + j); };
  auto r_free = [&](int j) { return bar0 + 8u * (32 + j); };
  auto o_written = [&](int j) { return bar0 + 8u * (36 + j); };
  const uint32_t d3_full = bar0 + 8u * 40;
  const uint32_t d3_empty = bar0 + 8u * 41;
  const uint32_t z1o_written = bar0 + 8u * 42;
  const uint32_t z1o_free = bar0 + 8u * 43;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::kTmemPtrOff);
  float* s_par = reinterpret_cast<float*>(smem + L::kParamOff);  // sc2[64] sh2[64] sc3[256] sh3[256] sc1n[64] sh1n[64]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int trace_n = 0;
  auto trace = [&](int code) {
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < kFbTraceLen)
      p.trace[warp * kFbTraceLen + trace_n++] = (static_cast<unsigned long long>(clock64()) << 8) | static_cast<unsigned>(code);
  };
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kD1 = 0, kD2 = 128, kD3 = 384;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_z1);
    tma_prefetch_desc(&p.tmap_w2);
    tma_prefetch_desc(&p.tmap_w3);
    tma_prefetch_desc(&p.tmap_res);
    tma_prefetch_desc(&p.tmap_out);
    if (N3 > 0) {
      tma_prefetch_desc(&p.tmap_w1n);
      tma_prefetch_desc(&p.tmap_z1o);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(p_full(s), 1);
      mbar_init(p_empty(s), 1);
      mbar_init(w1_full(s), 1);
      mbar_init(w1_empty(s), 1);
      mbar_init(d1_full(s), 1);
      mbar_init(d1_empty(s), kFbEpiWarps);       // one arrive per epilogue warp
    }
    for (int s = 0; s < kW2Slots; ++s) {
      mbar_init(w2_full(s), 1);
      mbar_init(w2_empty(s), 1);
    }
    mbar_init(w3_full, 1);
    mbar_init(z2_full, kFbEpiWarps);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, kFbEpiWarps);
    for (int j = 0; j < 4; ++j) {
      mbar_init(r_full(j), 1);
      mbar_init(r_free(j), N3 > 0 ? 2 : 1);  // the slab's store has been read (+ G3 has consumed it)
      mbar_init(o_written(j), 4);            // the four warps of the group that owns slab j
    }
    mbar_init(d3_full, 1);
    mbar_init(d3_empty, kFbEpiWarps);
    mbar_init(z1o_written, kFbEpiWarps);
    mbar_init(z1o_free, 1);
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  grid_dependency_wait();

  const int tile0 = static_cast<int>(blockIdx.x);
  const int step = static_cast<int>(gridDim.x);
  const int nt = tile0 < p.num_tiles ? (p.num_tiles - tile0 + step - 1) / step : 0;  // tiles of this CTA
  auto tile_origin = [&](int k, int& w0, int& h0, int& img) {
    const int t = tile0 + k * step;
    const int tw = t % p.tiles_w;
    const int r = t / p.tiles_w;
    w0 = tw * kPatchBW;
    h0 = (r % p.tiles_h) * kPatchBH;
    img = r / p.tiles_h;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ producers: lane 0 halo patches (+ W3 once),
    // lane 1 W2 taps.  The two lanes run their own loops (divergent; each waits on its own mbarriers).
    if (lane == 0) {
      mbar_arrive_expect_tx(w3_full, 256 * 128);
      tma_load_2d(s_w3, &p.tmap_w3, w3_full, 0, 0);
      for (int k = 0; k < nt; ++k) {
        int w0, h0, img;
        tile_origin(k, w0, h0, img);
        const int ps = k & 1;
        mbar_wait(p_empty(ps), (static_cast<uint32_t>(k >> 1) & 1u) ^ 1u);
        trace(1);   // patch load issued
        mbar_arrive_expect_tx(p_full(ps), kPatchBytes);
        tma_load_4d(s_patch + ps * kPatchStageBytes, &p.tmap_z1, p_full(ps), 0, w0 - 1, h0 - 1, img);
      }
    } else if (lane == 1 && kW2Res) {
      mbar_arrive_expect_tx(w2_full(0), 9 * kFbTapBytes);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(s_w2 + tap * kFbTapBytes, &p.tmap_w2, w2_full(0), tap * 64, 0);
    } else if (lane == 1) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int k = 0; k < nt; ++k) {
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(w2_empty(ws), wphase ^ 1u);
          mbar_arrive_expect_tx(w2_full(ws), kFbTapBytes);
          tma_load_2d(s_w2 + ws * kFbTapBytes, &p.tmap_w2, w2_full(ws), tap * 64, 0);
          if (++ws == kW2Slots) { ws = 0; wphase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t fx = p.x_fp16 ? kFmtF16 : kFmtBF16;
      const uint32_t fo = p.out_fp16 ? kFmtF16 : kFmtBF16;
      const uint32_t idesc1 = make_idesc_f16kind(kBM, 64, fx, fx);
      const uint32_t idesc2 = make_idesc_f16kind(kBM, 256, fx, fx);
      const uint32_t idesc3 = make_idesc_f16kind(kBM, 64, fo, fo);
      const uint64_t d_patch0 = make_smem_desc_sw128_sbo(s_patch, kPatchPW * 128);
      const uint64_t d_w2 = make_smem_desc_sw128(s_w2);
      const uint64_t d_w3 = make_smem_desc_sw128(s_w3);
      const uint64_t d_w1 = make_smem_desc_sw128(s_w1);
      const uint64_t d_z2 = make_smem_desc_sw128(s_z2);
      const uint64_t d_patch_z2 = make_smem_desc_sw128(s_patch);   // z2 inside a patch stage: a plain 128-row slab
      const uint64_t d_out = make_smem_desc_sw128(s_out);
      int ws = 0;
      uint32_t wphase = 0;
      int w1s = 0;
      uint32_t w1phase = 0;
      // The issuing thread is the scarce resource of this kernel (56 MMAs per tile from ONE thread that shares its
      // scheduler with two busy epilogue warps; a first version with run-time tap / tile arithmetic in the loop
      // measured ~1000 cycles per tap instead of the ~560 its four MMAs take): the tap loops below are fully
      // unrolled, every A descriptor is (tile base + compile-time tap offset), the B descriptor advances by a constant.
      uint64_t dp = 0;       // A descriptor base of the tile whose taps are being issued
      uint32_t d1 = 0;       // its accumulator
      uint64_t db = d_w2;    // B descriptor of the next W2 ring slot
      auto g1_begin = [&](int k) {   // first tap of tile k: the patch has landed, the accumulator is free
        const int a = k & 1;
        const uint32_t ph = static_cast<uint32_t>(k >> 1) & 1u;
        mbar_wait(p_full(a), ph);
        trace(10);  // patch landed
        mbar_wait(d1_empty(a), ph ^ 1u);
        trace(11);  // D1 buffer free
        if (kW2Res && k == 0) mbar_wait(w2_full(0), 0);
        tc_fence_after();
      };
      auto g1_select = [&](int k) {  // descriptors of tile k
        const int a = k & 1;
        dp = d_patch0 + static_cast<uint32_t>(a * (kPatchStageBytes >> 4));
        d1 = tmem_base + kD1 + static_cast<uint32_t>(a * 64);
      };
      auto g1_tap = [&](int tap) {   // `tap` is a compile-time constant wherever this is called from an unrolled loop
        const uint64_t da = dp + static_cast<uint32_t>(((tap / 3) * kPatchPW + tap % 3) * 8);
        if (kW2Res) {
          const uint64_t dbr = d_w2 + static_cast<uint32_t>(tap * (kFbTapBytes >> 4));
#pragma unroll
          for (int q = 0; q < 4; ++q) umma_bf16_ss(d1, da + 2u * q, dbr + 2u * q, idesc1, (tap | q) != 0 ? 1u : 0u);
          return;
        }
        mbar_wait(w2_full(ws), wphase);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) umma_bf16_ss(d1, da + 2u * q, db + 2u * q, idesc1, (tap | q) != 0 ? 1u : 0u);
        umma_commit(w2_empty(ws));
        db += static_cast<uint32_t>(kFbTapBytes >> 4);
        if (++ws == kW2Slots) { ws = 0; wphase ^= 1u; db = d_w2; }
      };
      auto g1_end = [&](int k) {
        if (!kW2Res) umma_commit(p_empty(k & 1));  // (resident W2: the stage also holds z2(k) -- freed after G2(k))
        umma_commit(d1_full(k & 1));
      };
      auto g2 = [&](int k) {
        const uint32_t ph = static_cast<uint32_t>(k) & 1u;
        trace(20);  // G2: start waiting for z2
        mbar_wait(z2_full, ph);
        trace(21);  // z2 ready
        mbar_wait(d2_empty, ph ^ 1u);
        trace(22);  // D2 free -> issue
        if (k == 0) mbar_wait(w3_full, 0);
        tc_fence_after();
        const uint64_t dz = kW2Res ? d_patch_z2 + static_cast<uint32_t>((k & 1) * (kPatchStageBytes >> 4)) : d_z2;
#pragma unroll
        for (int q = 0; q < 4; ++q) umma_bf16_ss(tmem_base + kD2, dz + 2u * q, d_w3 + 2u * q, idesc2, q != 0 ? 1u : 0u);
        umma_commit(d2_full);
        if (kW2Res) umma_commit(p_empty(k & 1));  // z2(k) consumed: the patch stage may be refilled
      };
      // next G3 chunk of tile k (chunk j = K range [64 j, 64 j + 64) = output slab j); returns false if `blocking` is
      // false and the chunk's operands are not ready yet
      int g3_next = 0;
      auto g3_chunk = [&](int k, bool blocking) -> bool {
        const uint32_t ph = static_cast<uint32_t>(k) & 1u;
        const int j = g3_next;
        if (blocking) {
          mbar_wait(o_written(j), ph);
          mbar_wait(w1_full(w1s), w1phase);
          if (j == 0) mbar_wait(d3_empty, ph ^ 1u);
        } else {
          if (!fb_try(o_written(j), ph) || !fb_try(w1_full(w1s), w1phase)) return false;
          if (j == 0 && !fb_try(d3_empty, ph ^ 1u)) return false;
        }
        trace(30 + j);  // G3 chunk j issued
        tc_fence_after();
        const uint64_t da = d_out + static_cast<uint32_t>(j * (kSlabBytes >> 4));
        const uint64_t db = d_w1 + static_cast<uint32_t>(w1s * (kFbTapBytes >> 4));
#pragma unroll
        for (int q = 0; q < 4; ++q) umma_bf16_ss(tmem_base + kD3, da + 2u * q, db + 2u * q, idesc3, (j | q) != 0 ? 1u : 0u);
        umma_commit(w1_empty(w1s));
        umma_commit(r_free(j));
        if (j == 3) umma_commit(d3_full);
        if (++w1s == kFbW1Slots) { w1s = 0; w1phase ^= 1u; }
        ++g3_next;
        return true;
      };
      // Issue order (the tensor pipe executes in order, the issuing thread runs at most a W2 ring ahead of it):
      //     G1(0) | taps 0-4 of G1(1) | { G2(k) | taps 5-8 of G1(k+1) | taps 0-4 of G1(k+2) }, k = 0, 1, ...
      // with the chunks of G3(k) slipped in between taps as soon as E2(k) has written their slab.  Nine taps of 3x3
      // work separate G2(k) from G2(k+1): while they run, the epilogue warps convert tile k (E2), then E1(k+1), whose
      // accumulator is complete after the first four of them.  D1[(k+2) & 1] is free since E1(k), i.e. before G2(k).
      if (nt > 0) {
        g1_begin(0);
        g1_select(0);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) g1_tap(tap);
        g1_end(0);
      }
      if (nt > 1) {
        g1_begin(1);
        g1_select(1);
#pragma unroll
        for (int tap = 0; tap < 5; ++tap) g1_tap(tap);
      }
      for (int k = 0; k < nt; ++k) {
        g2(k);
        g3_next = 0;
        if (k + 1 < nt) {
          g1_select(k + 1);
#pragma unroll
          for (int tap = 5; tap < 9; ++tap) {
            g1_tap(tap);
            if (N3 > 0 && g3_next < 4) g3_chunk(k, false);
          }
          g1_end(k + 1);
        }
        if (k + 2 < nt) {
          g1_begin(k + 2);
          g1_select(k + 2);
#pragma unroll
          for (int tap = 0; tap < 5; ++tap) {
            g1_tap(tap);
            if (N3 > 0 && g3_next < 4) g3_chunk(k, false);
          }
        }
        if (N3 > 0)
          while (g3_next < 4) g3_chunk(k, true);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ producer: residual tile (into the output slabs)
    // and the next conv1's weight chunks.  (One lane: lanes of a warp that spin on different mbarriers serialise
    // badly -- a lane-per-slab version measured 30 % slower.)
    if (lane == 0) {
      int w1s = 0;
      uint32_t w1phase = 0;
      for (int k = 0; k < nt; ++k) {
        int w0, h0, img;
        tile_origin(k, w0, h0, img);
        const uint32_t ph = static_cast<uint32_t>(k) & 1u;
        // The residual of tile k can only be requested once tile k-1's stores have been read (it lands in the same
        // slabs), which leaves its ~3000-cycle HBM latency exposed on the chain request -> E2 -> store -> request.  Its
        // L2 prefetch has no such dependence: issued a whole tile ahead, it turns the load below into an L2 hit.
        if (p.res_prefetch && k + 1 < nt) {
          int w1p, h1p, imgp;
          tile_origin(k + 1, w1p, h1p, imgp);
          for (int j = 0; j < 4; ++j) tma_prefetch_l2_4d(&p.tmap_res, j * 64, w1p, h1p, imgp);
        }
        for (int j = 0; j < 4; ++j) {
          mbar_wait(r_free(j), ph ^ 1u);
          trace(40 + j);  // residual slab j requested
          mbar_arrive_expect_tx(r_full(j), kSlabBytes);
          tma_load_4d(s_out + j * kSlabBytes, &p.tmap_res, r_full(j), j * 64, w0, h0, img);
        }
        if (N3 > 0) {
          for (int j = 0; j < 4; ++j) {
            mbar_wait(w1_empty(w1s), w1phase ^ 1u);
            mbar_arrive_expect_tx(w1_full(w1s), kFbTapBytes);
            tma_load_2d(s_w1 + w1s * kFbTapBytes, &p.tmap_w1n, w1_full(w1s), j * 64, 0);
            if (++w1s == kFbW1Slots) { w1s = 0; w1phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------------------------------------------ store issuer: the four slab stores of a tile are
    // issued as their slabs get published, each in its own bulk group; the read-waits (~700 cycles each when taken
    // one after the other) then overlap: wait_group.read 3, 2, 1, 0 frees the slabs in order
    if (lane == 0) {
      for (int k = 0; k < nt; ++k) {
        int w0, h0, img;
        tile_origin(k, w0, h0, img);
        const uint32_t ph = static_cast<uint32_t>(k) & 1u;
        for (int j = 0; j < 4; ++j) {
          mbar_wait(o_written(j), ph);
          trace(50 + j);  // output slab j written -> store
          asm volatile(
              "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
              ::"l"(reinterpret_cast<uint64_t>(&p.tmap_out)), "r"(s_out + j * kSlabBytes), "r"(j * 64), "r"(w0), "r"(h0),
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
        if (N3 > 0) {
          mbar_wait(z1o_written, ph);
          asm volatile(
              "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
              ::"l"(reinterpret_cast<uint64_t>(&p.tmap_z1o)), "r"(s_z1o), "r"(0), "r"(w0), "r"(h0), "r"(img)
              : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(z1o_free);
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 4..11)
    const int quad = warp & 3;
    const int grp = (warp - 4) >> 2;  // 0..3
    const int row = quad * 32 + lane;
    const int etid = threadIdx.x - 128;  // 0..511
    const bool x_fp16 = p.x_fp16 != 0, out_fp16 = p.out_fp16 != 0, res_fp16 = p.res_fp16 != 0, z1o_fp16 = p.z1o_fp16 != 0;
    // exponents from the chained bounds
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
    const float b1n = (N3 > 0 && p.consts1n) ? p.consts1n[0] * b3 + p.consts1n[1] : 0.0f;
    const int e_z2 = p.z2_scaled ? exp_of(b2) : 0;
    const int e_out = p.out_scaled ? exp_of(b3) : 0;
    const int e_z1o = (N3 > 0 && p.z1o_scaled) ? exp_of(b1n) : 0;
    if (blockIdx.x == 0 && etid == 0) {
      if (p.out_meta && p.out_scaled) p.out_meta->e = e_out;
      if (N3 > 0 && p.z1o_meta && p.z1o_scaled) p.z1o_meta->e = e_z1o;
    }
    {
      const float m2 = ldexpf(1.0f, e_z1 - e_z2), a2 = ldexpf(1.0f, -e_z2);
      const float m3 = ldexpf(1.0f, e_z2 - e_out), a3 = ldexpf(1.0f, -e_out);
      const float m1 = ldexpf(1.0f, e_out - e_z1o), a1 = ldexpf(1.0f, -e_z1o);
      for (int i = etid; i < 64; i += kFbEpiThreads) {
        s_par[i] = (p.scale2 ? __ldg(p.scale2 + i) : 1.0f) * m2;
        s_par[64 + i] = (p.shift2 ? __ldg(p.shift2 + i) : 0.0f) * a2;
        if (N3 > 0) {
          s_par[640 + i] = (p.scale1n ? __ldg(p.scale1n + i) : 1.0f) * m1;
          s_par[704 + i] = (p.shift1n ? __ldg(p.shift1n + i) : 0.0f) * a1;
        }
      }
      for (int i = etid; i < 256; i += kFbEpiThreads) {
        s_par[128 + i] = (p.scale3 ? __ldg(p.scale3 + i) : 1.0f) * m3;
        s_par[384 + i] = (p.shift3 ? __ldg(p.shift3 + i) : 0.0f) * a3;
      }
      named_bar_sync(1, kFbEpiThreads);
    }
    const float mul_res = ldexpf(1.0f, e_res - e_out);
    const float* sc2 = s_par;
    const float* sh2 = s_par + 64;
    const float* sc3 = s_par + 128;
    const float* sh3 = s_par + 384;
    const float* sc1n = s_par + 640;
    const float* sh1n = s_par + 704;
    float amax_out = 0.0f, amax_z1o = 0.0f;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);

    // 16 accumulator columns [c0, c0 + 16) of `acc` -> scale/shift + ReLU -> 16-bit -> row `row` of a 64-column slab
    auto convert16 = [&](uint32_t acc_col, const float* sc, const float* sh, int c0, uint32_t slab, bool fp16,
                         bool track, float& amax) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(lane_base + acc_col + static_cast<uint32_t>(c0), v);
      tmem_ld_wait();
      const uint32_t rbase = slab + row * 128;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; e += 4) {
          const float4 s4 = *reinterpret_cast<const float4*>(sc + c0 + 8 * j + e);
          const float4 h4 = *reinterpret_cast<const float4*>(sh + c0 + 8 * j + e);
          x[e + 0] = fmaxf(fmaf(__uint_as_float(v[8 * j + e + 0]), s4.x, h4.x), 0.0f);
          x[e + 1] = fmaxf(fmaf(__uint_as_float(v[8 * j + e + 1]), s4.y, h4.y), 0.0f);
          x[e + 2] = fmaxf(fmaf(__uint_as_float(v[8 * j + e + 2]), s4.z, h4.z), 0.0f);
          x[e + 3] = fmaxf(fmaf(__uint_as_float(v[8 * j + e + 3]), s4.w, h4.w), 0.0f);
        }
        if (track) {
#pragma unroll
          for (int e = 0; e < 8; ++e) amax = fmaxf(amax, x[e]);
        }
        uint4 o;
        o.x = pack16x2(x[0], x[1], fp16);
        o.y = pack16x2(x[2], x[3], fp16);
        o.z = pack16x2(x[4], x[5], fp16);
        o.w = pack16x2(x[6], x[7], fp16);
        const uint32_t a = rbase + ((((c0 >> 3) + j) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                     : "memory");
      }
    };

    auto e1 = [&](int k) {
      const int a = k & 1;
      trace(60);  // E1: waiting for D1
      mbar_wait(d1_full(a), static_cast<uint32_t>(k >> 1) & 1u);
      trace(61);  // D1 ready
      tc_fence_after();
      float dummy = 0.0f;
      convert16(kD1 + static_cast<uint32_t>(a * 64), sc2, sh2, grp * 16, z2_of(k), x_fp16, false, dummy);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(d1_empty(a));
        mbar_arrive(z2_full);
      }
      trace(62);  // z2 published
    };

    if (nt > 0) e1(0);
    for (int k = 0; k < nt; ++k) {
      int w0, h0, img;
      tile_origin(k, w0, h0, img);
      const bool valid = (h0 + (row >> 3) < p.H) && (w0 + (row & 7) < p.W);
      const uint32_t ph = static_cast<uint32_t>(k) & 1u;
      // ---- E2: out = relu(D2 * scale3 + shift3 + residual), in place over the residual slabs of this group
      trace(70);  // E2: waiting for D2
      mbar_wait(d2_full, ph);
      trace(71);  // D2 ready
      tc_fence_after();
      {
        const int j = grp;      // this group's output slab
        constexpr int jj = 1;   // (its only one: the accumulator is handed back after it)
        mbar_wait(r_full(j), ph);
        trace(72);  // residual slab landed
        const uint32_t rbase = s_out + j * kSlabBytes + row * 128;
        // (storage formats as compile-time constants: with run-time flags the compiler computes both the fp16 and the
        // bf16 side of every unpack / pack in the unrolled loops)
        auto convert = [&](auto res16_c, auto out16_c) {
          constexpr bool kR16 = decltype(res16_c)::value, kO16 = decltype(out16_c)::value;
          // both halves of the slab's accumulator are requested before any arithmetic (two tcgen05.ld in flight)
          uint32_t va[32], vb[32];
          tmem_ld_32x32b_x32(lane_base + kD2 + static_cast<uint32_t>(j * 64), va);
          tmem_ld_32x32b_x32(lane_base + kD2 + static_cast<uint32_t>(j * 64 + 32), vb);
  #pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t (&v)[32] = half ? vb : va;
            uint4 rr[4];
  #pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t a = rbase + ((((half << 2) | c) ^ (row & 7)) << 4);
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(rr[c].x), "=r"(rr[c].y), "=r"(rr[c].z), "=r"(rr[c].w)
                           : "r"(a));
            }
            if (half == 0) tmem_ld_wait();
            const int cb = j * 64 + half * 32;
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
              const uint32_t a = rbase + ((((half << 2) | c) ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                           : "memory");
            }
          }
        };
        using T = std::true_type;
        using F = std::false_type;
        if (res_fp16) { if (out_fp16) convert(T{}, T{}); else convert(T{}, F{}); }
        else { if (out_fp16) convert(F{}, T{}); else convert(F{}, F{}); }
        if (jj == 1) tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(o_written(j));
          if (jj == 1) mbar_arrive(d2_empty);
        }
        trace(74);  // output slab published
      }
      // ---- E1 of the next tile (its 3x3 ran on the tensor pipe meanwhile), then E3 of this one
      if (k + 1 < nt) e1(k + 1);
      if (N3 > 0) {
        trace(80);  // E3: waiting for D3
        mbar_wait(d3_full, ph);
        trace(81);
        mbar_wait(z1o_free, ph ^ 1u);
        tc_fence_after();
        convert16(kD3, sc1n, sh1n, grp * 16, s_z1o, z1o_fp16, valid, amax_z1o);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(d3_empty);
          mbar_arrive(z1o_written);
        }
        trace(82);  // z1' published
      }
    }
    if (p.out_meta) {
      float a = amax_out;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
      if (lane == 0) atomicMax(&p.out_meta->amax_bits, __float_as_uint(ldexpf(a, e_out)));
    }
    if (N3 > 0 && p.z1o_meta) {
      float a = amax_z1o;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
      if (lane == 0) atomicMax(&p.z1o_meta->amax_bits, __float_as_uint(ldexpf(a, e_z1o)));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tdet
