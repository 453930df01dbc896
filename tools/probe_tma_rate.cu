// Probe: per-SM throughput of TMA loads that deposit a 128-row x 128-byte A tile (16 KiB), for the
// load shapes the conv kernels could use.  One producer thread per CTA issues `iters` loads into a
// 4-deep ring; a consumer thread waits for each and frees the slot.  Source tensor 2 x 200 x 336 x C
// (L2 resident after the first pass).  Development probe; not part of the library.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../torch_detection_b200/csrc/ptx_sm100.cuh"
using namespace tdet;

constexpr int kStages = 6;
constexpr int kTile = 16384;

// mode 0: im2col 4D (c,w,h,n) 128 pixels; 1: tiled 4D box (64,bw,bh,1); 2: tiled 2D (64,128 rows)
__global__ void __launch_bounds__(64, 1)
rate(const __grid_constant__ CUtensorMap tm, int mode, int iters, int W, int H, int N, int bw, int bh,
     int kchunks, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar0 = base + kStages * kTile;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * (kStages + s), 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int tiles_w = W / bw, tiles_h = H / bh;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    unsigned tile = blockIdx.x * 7919u;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(bar0 + 8 * (kStages + stage), phase ^ 1u);
      const uint32_t fb = bar0 + 8 * stage;
      mbar_arrive_expect_tx(fb, kTile);
      const uint32_t dst = base + stage * kTile;
      const int kc = (i % kchunks) * 64;
      const int tap = (i / kchunks) % 9;
      if (i % (9 * kchunks) == 0) tile = tile * 1664525u + 1013904223u;
      if (mode == 0) {
        const int m0 = (tile % (unsigned)((W * H * N) / 128)) * 128;
        const int q = m0 % W, t = m0 / W, p = t % H, n = t / H;
        tma_load_im2col_4d(dst, &tm, fb, kc, q - 1, p - 1, n, (uint16_t)(tap % 3), (uint16_t)(tap / 3));
      } else if (mode == 1) {
        const unsigned tt = tile % (unsigned)(tiles_w * tiles_h * N);
        const int tw = tt % tiles_w, th = (tt / tiles_w) % tiles_h, n = tt / (tiles_w * tiles_h);
        tma_load_4d(dst, &tm, fb, kc, tw * bw + tap % 3 - 1, th * bh + tap / 3 - 1, n);
      } else {
        const int m0 = (tile % (unsigned)((W * H * N) / 128)) * 128;
        tma_load_2d(dst, &tm, fb, kc, m0);
      }
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(bar0 + 8 * stage, phase);
      mbar_arrive(bar0 + 8 * (kStages + stage));
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int N = 2, H = 200, W = 336;
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)f;
  cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q);
  EncodeIm2colFn enci = (EncodeIm2colFn)f;
  const int smem = kStages * kTile + 256;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* dcy; cudaMalloc(&dcy, 148 * 8);
  for (int C : {64, 256}) {
    const size_t elems = (size_t)N * H * W * C;
    uint16_t* d; cudaMalloc(&d, elems * 2); cudaMemset(d, 0, elems * 2);
    struct V { const char* name; int mode, bw, bh; } vs[] = {
      {"im2col_128px", 0, 1, 1}, {"tiled4d_16x8", 1, 16, 8}, {"tiled4d_32x4", 1, 32, 4}, {"tiled4d_8x16", 1, 8, 16},
      {"tiled2d_flat", 2, 1, 1}};
    for (auto& v : vs) {
      CUtensorMap tm; CUresult r;
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t st[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t es[4] = {1, 1, 1, 1};
      if (v.mode == 0) {
        int lo[2] = {-1, -1}, up[2] = {-1, -1};
        r = enci(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, st, lo, up, 64, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else if (v.mode == 1) {
        cuuint32_t box[4] = {64, (cuuint32_t)v.bw, (cuuint32_t)v.bh, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        cuuint64_t d2[2] = {(cuuint64_t)C, (cuuint64_t)N * H * W};
        cuuint64_t s2[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, 128}; cuuint32_t e2[2] = {1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, d2, s2, box, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (r != CUDA_SUCCESS) { printf("%s C=%d encode failed %d\n", v.name, C, (int)r); continue; }
      const int iters = 4000;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        rate<<<148, 64, smem>>>(tm, v.mode, iters, W, H, N, v.bw, v.bh, C / 64, dcy);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long cy[148]; cudaMemcpy(cy, dcy, sizeof(cy), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += cy[i]; avg /= 148;
        if (rep == 1)
          printf("C=%3d %-14s: %.3f ms, %.0f cycles/load (%.1f cycles per 128B row), %.2f TB/s aggregate\n", C, v.name, ms,
                 avg / iters, avg / iters / 128, 148.0 * iters * kTile / (ms * 1e-3) / 1e12);
      }
    }
    cudaFree(d);
  }
  return 0;
}
