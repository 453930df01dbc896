// Probe: can a tcgen05.mma A descriptor (K-major, SWIZZLE_128B) start at an arbitrary 128-byte row of a
// shared-memory patch that TMA-style swizzling laid out relative to a 1024-byte aligned base, and use a
// stride between 8-row groups (SBO) that is not a multiple of 1024?  This decides whether the 9 taps
// of a 3x3 convolution can be formed as shifted views of ONE halo patch instead of 9 im2col loads.
// Development probe; not part of the library.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../torch_detection_b200/csrc/ptx_sm100.cuh"
using namespace tdet;

constexpr int PW = 10, PH = 18;  // patch, pixels (for the 8x16 tile); also used as 1 x 180 strip
constexpr int N = 64;

__global__ void __launch_bounds__(128, 1)
probe(const __nv_bfloat16* patch /*[PH*PW][64]*/, const __nv_bfloat16* bmat /*[64][64]*/, float* out /*[128][64]*/,
      int row0, int sbo_bytes, int base_offset) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  uint8_t* a_s = smem;                      // PH*PW rows x 128 B, swizzled like TMA SWIZZLE_128B
  uint8_t* b_s = smem + 24576;              // 64 rows x 128 B
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 24576 + 8192);
  const uint32_t bar = base + 24576 + 8192 + 16;
  for (int i = threadIdx.x; i < PH * PW * 8; i += blockDim.x) {
    const int row = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(a_s + row * 128 + ((c ^ (row & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(patch + row * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int row = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(b_s + row * 128 + ((c ^ (row & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(bmat + row * 64 + c * 8);
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(tptr), 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16kind(128, N, kFmtBF16, kFmtBF16);
    uint64_t da = 0;
    const uint32_t a_addr = base + row0 * 128;
    da |= static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4);
    da |= static_cast<uint64_t>(1) << 16;
    da |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    da |= static_cast<uint64_t>(1) << 46;
    da |= static_cast<uint64_t>(base_offset & 7) << 49;
    da |= static_cast<uint64_t>(2) << 61;
    const uint64_t db = make_smem_desc_sw128(base + 24576);
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, da + 2u * k, db + 2u * k, idesc, k != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + half * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + half * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__nv_bfloat16> hp(PH * PW * 64), hb(64 * 64);
  std::vector<float> fp(PH * PW * 64), fb(64 * 64);
  srand(1);
  for (size_t i = 0; i < hp.size(); ++i) { float v = (rand() % 17 - 8) / 8.0f; hp[i] = __float2bfloat16(v); fp[i] = v; }
  for (size_t i = 0; i < hb.size(); ++i) { float v = (rand() % 13 - 6) / 4.0f; hb[i] = __float2bfloat16(v); fb[i] = v; }
  __nv_bfloat16 *dp, *db; float* dout;
  cudaMalloc(&dp, hp.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 24576 + 8192 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Case { const char* name; int bw, bh, r, s; };
  Case cases[] = {{"strip 128x1 shift 0", 128, 1, 0, 0}, {"strip 128x1 shift 1", 128, 1, 0, 1},
                  {"strip 128x1 shift 2", 128, 1, 0, 2}, {"strip 128x1 shift 11", 128, 1, 0, 11},
                  {"tile 8x16 tap(0,0)", 8, 16, 0, 0},   {"tile 8x16 tap(0,1)", 8, 16, 0, 1},
                  {"tile 8x16 tap(1,0)", 8, 16, 1, 0},   {"tile 8x16 tap(1,2)", 8, 16, 1, 2},
                  {"tile 8x16 tap(2,2)", 8, 16, 2, 2}};
  for (auto& c : cases) {
    // expected: output row m=(j,i) <- patch pixel rho = (j+r)*PW + (i+s) for the 8x16 tile (PW=10);
    // for the strip the patch is one line of 180 pixels: rho = m + s
    const int row0 = c.bh == 1 ? c.s : c.r * PW + c.s;
    const int sbo = c.bh == 1 ? 1024 : PW * 128;
    for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
      const int bo = bo_mode == 0 ? 0 : (row0 & 7);
      cudaMemset(dout, 0, 128 * 64 * 4);
      probe<<<1, 128, smem>>>(dp, db, dout, row0, sbo, bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      std::vector<float> o(128 * 64);
      cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0; int firstbad = -1;
      for (int m = 0; m < 128; ++m) {
        const int rho = c.bh == 1 ? m + c.s : ((m / 8) + c.r) * PW + (m % 8) + c.s;
        for (int n = 0; n < 64; ++n) {
          float ref = 0; for (int k = 0; k < 64; ++k) ref += fp[rho * 64 + k] * fb[n * 64 + k];
          if (fabsf(ref - o[m * 64 + n]) > 1e-3f) { ++bad; if (firstbad < 0) firstbad = m; }
        }
      }
      printf("%-22s row0=%3d sbo=%4d base_offset=%d : %s (bad=%d, first bad row %d)\n", c.name, row0, sbo, bo,
             bad == 0 ? "OK" : "MISMATCH", bad, firstbad);
    }
  }
  return 0;
}
