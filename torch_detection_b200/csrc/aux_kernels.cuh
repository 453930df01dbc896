// Memory-bound tails of the path (HBM roofline, not tensor): image staging, max-pool, stride-2
// subsample, and the one-off operand preparation kernels (weight packing, BN folding, bound
// constants).  All activation kernels use 8- or 16-byte vector accesses on dense NHWC 16-bit data.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "conv_gemm.cuh"

namespace tdet {

// (n,3,h,w) fp32/bf16 with arbitrary element strides -> [n][hp][wp][4] bf16, image at (3,3), zero
// border, zero 4th channel.  One thread per staged pixel (8-byte store, coalesced along wp).
// Optionally records the image's |max| (of the bf16-rounded values) in `meta`.
template <typename T>
__global__ void __launch_bounds__(256)
prep_image_kernel(const T* __restrict__ x, long long sn, long long sc, long long sh, long long sw,
                  int n, int h, int w, int hp, int wp, uint2* __restrict__ y, TensorMeta* meta) {
  const long long total = static_cast<long long>(n) * hp * wp;
  float amax = 0.0f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xw = static_cast<int>(i % wp);
    const long long t = i / wp;
    const int yh = static_cast<int>(t % hp);
    const int img = static_cast<int>(t / hp);
    const int iw = xw - 3, ih = yh - 3;
    uint2 o = make_uint2(0u, 0u);
    if (iw >= 0 && iw < w && ih >= 0 && ih < h) {
      const T* px = x + img * sn + ih * sh + iw * sw;
      const float c0 = static_cast<float>(px[0]);
      const float c1 = static_cast<float>(px[sc]);
      const float c2 = static_cast<float>(px[2 * sc]);
      o.x = pack_bf16x2(c0, c1);
      o.y = pack_bf16x2(c2, 0.0f);
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(bf16_lo(o.x)), fabsf(bf16_hi(o.x))), fabsf(bf16_lo(o.y))));
    }
    y[i] = o;
  }
  if (meta) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0 && amax > 0.0f) atomicMax(&meta->amax_bits, __float_as_uint(amax));
  }
}

template <bool FP16>
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b) {
  if (FP16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a),
                                     *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}

// 3x3 / stride 2 / pad 1 max-pool on NHWC 16-bit; one thread per (output pixel, 8 channels).
template <bool FP16>
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int c8,
                    int ho, int wo) {
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  // -inf: fp16 0xFC00, bf16 0xFF80
  const uint32_t ninf = FP16 ? 0xFC00FC00u : 0xFF80FF80u;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    long long t = i / c8;
    const int ow = static_cast<int>(t % wo);
    t /= wo;
    const int oh = static_cast<int>(t % ho);
    const int img = static_cast<int>(t / ho);
    uint4 m = make_uint4(ninf, ninf, ninf, ninf);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ih = 2 * oh - 1 + dy;
      if (ih < 0 || ih >= h) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int iw = 2 * ow - 1 + dx;
        if (iw < 0 || iw >= w) continue;
        const uint4 v = __ldg(x + ((static_cast<long long>(img) * h + ih) * w + iw) * c8 + cg);
        m.x = max16x2<FP16>(m.x, v.x);
        m.y = max16x2<FP16>(m.y, v.y);
        m.z = max16x2<FP16>(m.z, v.z);
        m.w = max16x2<FP16>(m.w, v.w);
      }
    }
    y[i] = m;
  }
}

// y[n][i][j][:] = x[n][2i][2j][:]  (F.max_pool2d(x, 1, stride=2))
__global__ void __launch_bounds__(256)
subsample2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int c8,
                  int ho, int wo) {
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    long long t = i / c8;
    const int ow = static_cast<int>(t % wo);
    t /= wo;
    const int oh = static_cast<int>(t % ho);
    const int img = static_cast<int>(t / ho);
    y[i] = __ldg(x + ((static_cast<long long>(img) * h + 2 * oh) * w + 2 * ow) * c8 + cg);
  }
}

template <typename W>
__device__ __forceinline__ W to_w16(float v);
template <>
__device__ __forceinline__ __half to_w16<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 to_w16<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// fp32 OIHW -> 16-bit [O][kh][kw][I]
template <typename W>
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, W* __restrict__ out, int cout, int cin, int kh, int kw) {
  const long long total = static_cast<long long>(cout) * cin * kh * kw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin);
    long long t = i / cin;
    const int s = static_cast<int>(t % kw);
    t /= kw;
    const int r = static_cast<int>(t % kh);
    const int co = static_cast<int>(t / kh);
    out[i] = to_w16<W>(w[((static_cast<long long>(co) * cin + ci) * kh + r) * kw + s]);
  }
}

// fp32 [64][3][7][7] -> bf16 [64][448], k = r*64 + s*4 + c for s < 7, c < 3 (zero elsewhere): one
// 64-wide k-block per filter row, matching the 16-pixel x 4-channel window rows the stem loads.
__global__ void __launch_bounds__(256)
pack_stem_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 448) return;
  const int co = i / 448;
  const int k = i - co * 448;
  const int r = k >> 6;
  const int s = (k & 63) >> 2;
  const int c = k & 3;
  float v = 0.0f;
  if (s < 7 && c < 3) v = w[((co * 3 + c) * 7 + r) * 7 + s];
  out[i] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
fold_bn_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ mean, const float* __restrict__ var, float eps,
               float* __restrict__ scale, float* __restrict__ shift, int ch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ch) return;
  // same operation order as ATen's eval batch_norm: invstd = 1/sqrt(var+eps); w*invstd; b - mean*scale
  const float invstd = 1.0f / sqrtf(var[i] + eps);
  const float sc = gamma[i] * invstd;
  scale[i] = sc;
  shift[i] = beta[i] - mean[i] * sc;
}

// consts[0] = max_c |scale_c| * sum_k |w[c][k]| ; consts[1] = max_c |shift_c|.  One block per output
// channel; consts must be zeroed by the caller (float bit patterns of non-negative values order like
// unsigned integers, so atomicMax works).  A 1+2^-10 safety factor covers the fp32 summation error.
template <typename W>
__global__ void __launch_bounds__(256)
bound_consts_kernel(const W* __restrict__ w, const float* __restrict__ scale,
                    const float* __restrict__ shift, int k, unsigned* __restrict__ consts) {
  const int c = blockIdx.x;
  float s = 0.0f;
  for (int i = threadIdx.x; i < k; i += blockDim.x) s += fabsf(static_cast<float>(w[static_cast<long long>(c) * k + i]));
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float g = red[0] * fabsf(scale ? scale[c] : 1.0f) * 1.001f;
    atomicMax(&consts[0], __float_as_uint(g));
    atomicMax(&consts[1], __float_as_uint(fabsf(shift ? shift[c] : 0.0f)));
  }
}

}  // namespace tdet
