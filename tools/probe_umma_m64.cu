// Probe: where do the 64 accumulator rows of a tcgen05.mma with M = 64 (cta_group::1) land in TMEM, and how long
// does such an MMA take next to M = 128?  Decides whether the operand-swapped conv kernel can run Cout = 64 convs
// with an M = 64 weight operand instead of a zero-padded M = 128 one.  Development probe; not part of the library.
//   A[r][k] = (k == 0 ? r + 1 : 0)  (M rows), B[n][k] = (k == 0 ? 1 : 0)  ->  D[r][n] = r + 1.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../torch_detection_b200/csrc/ptx_sm100.cuh"
using namespace tdet;

template <int M, int N>
__global__ void __launch_bounds__(128, 1) probe(float* out /*[128][2]*/, long long* cycles, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  uint8_t* a_s = smem;              // 128 rows x 128 B (SW128 K-major)
  uint8_t* b_s = smem + 16384;      // 256 rows x 128 B
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 16384 + 32768);
  const uint32_t bar = base + 16384 + 32768 + 16;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  for (int r = threadIdx.x; r < 128; r += blockDim.x) {
    const __nv_bfloat16 v = __float2bfloat16(static_cast<float>(r + 1));
    *reinterpret_cast<__nv_bfloat16*>(a_s + r * 128 + ((0 ^ (r & 7)) << 4)) = v;
  }
  for (int n = threadIdx.x; n < N; n += blockDim.x)
    *reinterpret_cast<__nv_bfloat16*>(b_s + n * 128 + ((0 ^ (n & 7)) << 4)) = __float2bfloat16(1.0f);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(tptr), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16kind(M, N, kFmtBF16, kFmtBF16);
    const uint64_t da = make_smem_desc_sw128(base);
    const uint64_t db = make_smem_desc_sw128(base + 16384);
    const long long t0 = clock64();
    for (int it = 0; it < reps; ++it)
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, da + 2u * k, db + 2u * k, idesc, (it | k) != 0);
    umma_commit(bar);
    mbar_wait(bar, 0);
    cycles[0] = clock64() - t0;
  }
  __syncthreads();
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t v[32];
  const int warp = threadIdx.x >> 5;
  tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  out[threadIdx.x * 2 + 0] = __uint_as_float(v[0]);
  out[threadIdx.x * 2 + 1] = __uint_as_float(v[31]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

template <int M, int N>
void run(const char* name) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 128 * 2 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  const int smem = 16384 + 32768 + 64;
  cudaFuncSetAttribute(probe<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int reps : {1, 256}) {
    cudaMemset(out, 0xFF, 128 * 2 * sizeof(float));
    probe<M, N><<<1, 128, smem>>>(out, cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<float> h(256);
    long long c;
    cudaMemcpy(h.data(), out, 256 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%s reps %d: %lld cycles (%.1f per MMA)\n", name, reps, c, static_cast<double>(c) / (4.0 * reps));
    if (reps == 1 && N == 256) {
      printf("  lane -> D row (value/1 of column 0; column 31 in brackets)\n  ");
      for (int l = 0; l < 128; ++l) {
        printf("%3d:%g[%g] ", l, h[2 * l], h[2 * l + 1]);
        if ((l & 7) == 7) printf("\n  ");
      }
      printf("\n");
    }
  }
}

int main() {
  run<128, 256>("M=128 N=256");
  run<64, 256>("M=64  N=256");
  run<128, 128>("M=128 N=128");
  run<128, 64>("M=128 N=64");
  run<64, 64>("M=64  N=64");
  return 0;
}
