// Probe: where does a 5D tiled TMA box with a 64-byte inner dimension land in shared memory under
// the different swizzle modes?  (Development probe for the stem loader; not part of the library.)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../torch_detection_b200/csrc/ptx_sm100.cuh"
using namespace tdet;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int rank, int c2, int c3, uint16_t* out, int bytes, int nbytes_expect) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xFFFF;
  const uint32_t bar = base + bytes;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, nbytes_expect);
    if (rank == 5) tma_load_5d(base, &tm, bar, 0, 0, c2, c3, 0);
    else asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(base), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bar), "r"(0), "r"(c2), "r"(c3), "r"(0) : "memory");
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int n = 1, ho = 16, wo = 40, hp = 2 * ho + 6, wp = 2 * wo + 6;
  const int total = n * hp * wp * 4;
  std::vector<uint16_t> h(total);
  for (int i = 0; i < total; ++i) h[i] = (uint16_t)i;  // value == element index
  uint16_t* d; cudaMalloc(&d, total * 2); cudaMemcpy(d, h.data(), total * 2, cudaMemcpyHostToDevice);
  const int bytes = 32768;
  uint16_t* dout; cudaMalloc(&dout, bytes);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)f;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 2048);
  struct V { const char* name; int rank; CUtensorMapSwizzle sw; } vs[] = {
    {"5D_SW128", 5, CU_TENSOR_MAP_SWIZZLE_128B}, {"5D_NONE", 5, CU_TENSOR_MAP_SWIZZLE_NONE},
    {"5D_SW64", 5, CU_TENSOR_MAP_SWIZZLE_64B}, {"4D_SW64", 4, CU_TENSOR_MAP_SWIZZLE_64B}, {"4D_NONE", 4, CU_TENSOR_MAP_SWIZZLE_NONE}};
  for (auto& v : vs) {
    CUtensorMap tm;
    CUresult r;
    if (v.rank == 5) {
      cuuint64_t dims[5] = {32, 2, (cuuint64_t)wo, (cuuint64_t)(ho + 3), (cuuint64_t)n};
      cuuint64_t st[4] = {(cuuint64_t)wp * 8, 16, (cuuint64_t)wp * 16, (cuuint64_t)hp * wp * 8};
      cuuint32_t box[5] = {32, 2, 32, 4, 1}; cuuint32_t es[5] = {1, 1, 1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[4] = {32, (cuuint64_t)wo, (cuuint64_t)(2 * ho + 6), (cuuint64_t)n};
      cuuint64_t st[3] = {16, (cuuint64_t)wp * 8, (cuuint64_t)hp * wp * 8};
      cuuint32_t box[4] = {32, 32, 4, 1}; cuuint32_t es[4] = {1, 1, 2, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    printf("== %s encode=%d\n", v.name, (int)r);
    if (r != CUDA_SUCCESS) continue;
    const int expect = v.rank == 5 ? 16384 : (32 * 32 * 4 * 2);
    probe<<<1, 128, bytes + 2048>>>(tm, v.rank, 0, 1, dout, bytes, expect);
    cudaError_t e = cudaDeviceSynchronize();
    printf("   run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> o(bytes / 2);
    cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
    // print, per 16-byte chunk of the first 2 KiB, the source element index of its first element
    int written = 0;
    for (int i = 0; i < bytes / 2; ++i) if (o[i] != 0xFFFF) ++written;
    printf("   elements written (non-sentinel): %d of %d\n", written, bytes / 2);
    for (int row = 0; row < 20; ++row) {
      printf("   smem+%5d:", row * 128);
      for (int c = 0; c < 8; ++c) {
        uint16_t val = o[(row * 128 + c * 16) / 2];
        if (val == 0xFFFF) printf("   ----"); else printf(" %6d", (int)val);
      }
      printf("\n");
    }
    // find last written chunk
    int last = -1;
    for (int i = 0; i < bytes / 2; ++i) if (o[i] != 0xFFFF) last = i;
    printf("   last written byte offset: %d\n", last * 2);
  }
  printf("ref: wp=%d row stride elems=%d ; window (wo=0,row r) starts at elem r*%d ; (wo=1) +8\n", wp, wp * 4, wp * 4);
  return 0;
}
