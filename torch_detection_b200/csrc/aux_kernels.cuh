// Memory-bound tails of the path (HBM roofline, not tensor): image staging, max-pool, stride-2
// subsample, and the one-off operand preparation kernels (weight packing, BN folding, bound
// constants).  All activation kernels use 8- or 16-byte vector accesses on dense NHWC 16-bit data.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "conv_gemm.cuh"

namespace tdet {

// (n,3,h,w) uint8/fp32/bf16 with arbitrary element strides (NCHW, channels_last or an HWC image batch
// viewed as NCHW) -> [n][hp][wp][4] bf16, image at (3,3), zero border, zero 4th channel.  One thread per
// staged pixel (8-byte store, coalesced along wp).  Optional per-channel affine `v * scale[c] + shift[c]`
// (the data layer's normalisation, scale = 1/std, shift = -mean/std) and zero padding beyond the valid
// (hv, wv) extent (pad-to-size-divisor): the reference's ImageTransforms steps 2 and 5
// (datasets/dataset_transforms.py:29-44) folded into the stem's loader.  Optionally records the staged
// image's |max| (of the bf16-rounded values) in `meta`.
template <typename T>
__device__ __forceinline__ float prep_load(const T* p) { return static_cast<float>(*p); }

// One thread stages TWO horizontally adjacent pixels (the pitch is a multiple of 16 pixels): one 16-byte store, and
// the (image, row, column) decode runs in 32-bit arithmetic per pair -- the staged batch has < 2^31 pixels, and the
// 64-bit div/mod per pixel of the first version cost more than the memory traffic (0.10 ms -> see DESIGN 3.3).
template <typename T>
__global__ void __launch_bounds__(256)
prep_image_kernel(const T* __restrict__ x, long long sn, long long sc, long long sh, long long sw,
                  int n, int hv, int wv, int hp, int wp, const float* __restrict__ scale,
                  const float* __restrict__ shift, uint2* __restrict__ y, TensorMeta* meta, int split, int y_fp16) {
  const long long total = static_cast<long long>(n) * hp * wp;
  const bool yf = y_fp16 != 0;   // fp16 staging (three more significand bits; |v| saturates at 65504); never with split
  const unsigned pairs = static_cast<unsigned>(total >> 1);
  const unsigned wp2 = static_cast<unsigned>(wp) >> 1;
  float amax = 0.0f;
  float s0 = 1.0f, s1 = 1.0f, s2 = 1.0f, b0 = 0.0f, b1 = 0.0f, b2 = 0.0f;
  if (scale) { s0 = scale[0]; s1 = scale[1]; s2 = scale[2]; }
  if (shift) { b0 = shift[0]; b1 = shift[1]; b2 = shift[2]; }
  uint4* __restrict__ y4 = reinterpret_cast<uint4*>(y);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += gridDim.x * blockDim.x) {
    const unsigned t = i / wp2;
    const int xw = static_cast<int>(i - t * wp2) * 2;
    const unsigned img = t / static_cast<unsigned>(hp);
    const int ih = static_cast<int>(t - img * static_cast<unsigned>(hp)) - 3;
    uint2 o[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
    uint2 lo[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
    if (ih >= 0 && ih < hv) {
      const T* row = x + img * sn + ih * sh;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int iw = xw + k - 3;
        if (iw >= 0 && iw < wv) {
          const T* px = row + iw * sw;
          float c0 = fmaf(prep_load(px), s0, b0);
          float c1 = fmaf(prep_load(px + sc), s1, b1);
          float c2 = fmaf(prep_load(px + 2 * sc), s2, b2);
          if (yf) {
            c0 = fminf(fmaxf(c0, -65504.0f), 65504.0f);
            c1 = fminf(fmaxf(c1, -65504.0f), 65504.0f);
            c2 = fminf(fmaxf(c2, -65504.0f), 65504.0f);
            o[k].x = pack16x2(c0, c1, true);
            o[k].y = pack16x2(c2, 0.0f, true);
            float v0, v1, v2, vz;
            unpack16x2(o[k].x, true, v0, v1);
            unpack16x2(o[k].y, true, v2, vz);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v0), fabsf(v1)), fabsf(v2)));
            continue;
          }
          o[k].x = pack_bf16x2(c0, c1);
          o[k].y = pack_bf16x2(c2, 0.0f);
          amax = fmaxf(amax, fmaxf(fmaxf(fabsf(bf16_lo(o[k].x)), fabsf(bf16_hi(o[k].x))), fabsf(bf16_lo(o[k].y))));
          if (split) {
            lo[k].x = pack_bf16x2(c0 - bf16_lo(o[k].x), c1 - bf16_hi(o[k].x));
            lo[k].y = pack_bf16x2(c2 - bf16_lo(o[k].y), 0.0f);
          }
        }
      }
    }
    y4[i] = make_uint4(o[0].x, o[0].y, o[1].x, o[1].y);
    // lo plane of the split-precision staging: the batch's second half [n .. 2n)
    if (split) y4[pairs + i] = make_uint4(lo[0].x, lo[0].y, lo[1].x, lo[1].y);
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

// fp32 OIHW -> 16-bit out[co * ld + (r*kw + s)*cin + ci] = round16(scale[co] * w): a folded BatchNorm scale multiplied
// into the rows, written with a row pitch so that two matrices can sit side by side ([W | W2], TDET_FLAG_DUAL)
template <typename W>
__global__ void __launch_bounds__(256)
pack_weight_scaled_kernel(const float* __restrict__ w, const float* __restrict__ scale, W* __restrict__ out, int cout,
                          int cin, int kh, int kw, int ld) {
  const long long total = static_cast<long long>(cout) * cin * kh * kw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin);
    long long t = i / cin;
    const int s = static_cast<int>(t % kw);
    t /= kw;
    const int r = static_cast<int>(t % kh);
    const int co = static_cast<int>(t / kh);
    const float v = w[((static_cast<long long>(co) * cin + ci) * kh + r) * kw + s] * (scale ? scale[co] : 1.0f);
    out[static_cast<long long>(co) * ld + (static_cast<long long>(r) * kw + s) * cin + ci] = to_w16<W>(v);
  }
}

// split precision: fp32 OIHW -> bf16 [O][kh][kw][2*I]: per tap I hi values then I lo values
__global__ void __launch_bounds__(256)
pack_weight_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin, int kh, int kw) {
  const long long total = static_cast<long long>(cout) * cin * kh * kw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin);
    long long t = i / cin;
    const int s = static_cast<int>(t % kw);
    t /= kw;
    const int r = static_cast<int>(t % kh);
    const int co = static_cast<int>(t / kh);
    const float v = w[((static_cast<long long>(co) * cin + ci) * kh + r) * kw + s];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const long long base = ((static_cast<long long>(co) * kh + r) * kw + s) * (2 * cin);
    out[base + ci] = hi;
    out[base + cin + ci] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// split-precision stem weights: [64][896] = {448 hi | 448 lo}, each in the pack_stem_weight layout
__global__ void __launch_bounds__(256)
pack_stem_weight_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 448) return;
  const int co = i / 448;
  const int k = i - co * 448;
  const int r = k >> 6;
  const int s = (k & 63) >> 2;
  const int c = k & 3;
  float v = 0.0f;
  if (s < 7 && c < 3) v = w[((co * 3 + c) * 7 + r) * 7 + s];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  out[co * 896 + k] = hi;
  out[co * 896 + 448 + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// split-precision 3x3/2 max-pool on [n][h][w][2*C] (hi | lo): the max is taken on hi + lo, the winning pair kept
__global__ void __launch_bounds__(256)
maxpool3x3s2_split_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int c8,
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
    float best[8];
    uint32_t bh[4] = {0, 0, 0, 0}, bl[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -3.0e38f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ih = 2 * oh - 1 + dy;
      if (ih < 0 || ih >= h) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int iw = 2 * ow - 1 + dx;
        if (iw < 0 || iw >= w) continue;
        const uint4* px = x + ((static_cast<long long>(img) * h + ih) * w + iw) * (2 * c8);
        const uint4 vh = __ldg(px + cg), vl = __ldg(px + c8 + cg);
        const uint32_t wh[4] = {vh.x, vh.y, vh.z, vh.w}, wl[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = bf16_lo(wh[j]) + bf16_lo(wl[j]), v1 = bf16_hi(wh[j]) + bf16_hi(wl[j]);
          if (v0 > best[2 * j]) {
            best[2 * j] = v0;
            bh[j] = (bh[j] & 0xFFFF0000u) | (wh[j] & 0xFFFFu);
            bl[j] = (bl[j] & 0xFFFF0000u) | (wl[j] & 0xFFFFu);
          }
          if (v1 > best[2 * j + 1]) {
            best[2 * j + 1] = v1;
            bh[j] = (bh[j] & 0xFFFFu) | (wh[j] & 0xFFFF0000u);
            bl[j] = (bl[j] & 0xFFFFu) | (wl[j] & 0xFFFF0000u);
          }
        }
      }
    }
    uint4* py = y + ((static_cast<long long>(img) * ho + oh) * wo + ow) * (2 * c8);
    py[cg] = make_uint4(bh[0], bh[1], bh[2], bh[3]);
    py[c8 + cg] = make_uint4(bl[0], bl[1], bl[2], bl[3]);
  }
}

// fp32 y[m][c] = hi + lo of a split-precision tensor x[m][2c]
__global__ void __launch_bounds__(256)
split_combine_kernel(const uint4* __restrict__ x, float4* __restrict__ y, long long rows, int c8) {
  const long long total = rows * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / c8;
    const int cg = static_cast<int>(i - m * c8);
    const uint4 vh = __ldg(x + m * (2 * c8) + cg), vl = __ldg(x + m * (2 * c8) + c8 + cg);
    float4* dst = y + i * 2;
    dst[0] = make_float4(bf16_lo(vh.x) + bf16_lo(vl.x), bf16_hi(vh.x) + bf16_hi(vl.x),
                         bf16_lo(vh.y) + bf16_lo(vl.y), bf16_hi(vh.y) + bf16_hi(vl.y));
    dst[1] = make_float4(bf16_lo(vh.z) + bf16_lo(vl.z), bf16_hi(vh.z) + bf16_hi(vl.z),
                         bf16_lo(vh.w) + bf16_lo(vl.w), bf16_hi(vh.w) + bf16_hi(vl.w));
  }
}

// grouped fp32 [O][I/g][kh][kw] -> dense 16-bit [O][kh][kw][I], zero outside the output channel's group
template <typename W>
__global__ void __launch_bounds__(256)
pack_grouped_weight_kernel(const float* __restrict__ w, W* __restrict__ out, int cout, int cin, int kh, int kw,
                           int groups) {
  const long long total = static_cast<long long>(cout) * cin * kh * kw;
  const int cig = cin / groups, cog = cout / groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin);
    long long t = i / cin;
    const int s = static_cast<int>(t % kw);
    t /= kw;
    const int r = static_cast<int>(t % kh);
    const int co = static_cast<int>(t / kh);
    const int g = co / cog;
    float v = 0.0f;
    if (ci / cig == g) v = w[((static_cast<long long>(co) * cig + (ci - g * cig)) * kh + r) * kw + s];
    out[i] = to_w16<W>(v);
  }
}

// fp32 [64][3][7][7] -> 16-bit [64][448], k = r*64 + s*4 + c for s < 7, c < 3 (zero elsewhere): one
// 64-wide k-block per filter row, matching the 16-pixel x 4-channel window rows the stem loads.
template <typename W>
__global__ void __launch_bounds__(256)
pack_stem_weight_kernel(const float* __restrict__ w, W* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 448) return;
  const int co = i / 448;
  const int k = i - co * 448;
  const int r = k >> 6;
  const int s = (k & 63) >> 2;
  const int c = k & 3;
  float v = 0.0f;
  if (s < 7 && c < 3) v = w[((co * 3 + c) * 7 + r) * 7 + s];
  out[i] = to_w16<W>(v);
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

// ---------------------------------------------------------------------------------------------------
// training path: small memory-bound helpers of the backward pass (all dense NHWC bf16, 16-byte accesses)
// ---------------------------------------------------------------------------------------------------

// Operand of the data-gradient conv: out[ci][kh-1-r][kw-1-s][co] = scale[co] * w[co][ci][r][s].
template <typename W>
__global__ void __launch_bounds__(256)
pack_dgrad_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale, W* __restrict__ out,
                         int cout, int cin, int kh, int kw) {
  const long long total = static_cast<long long>(cout) * cin * kh * kw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % cout);
    long long t = i / cout;
    const int s = static_cast<int>(t % kw);
    t /= kw;
    const int r = static_cast<int>(t % kh);
    const int ci = static_cast<int>(t / kh);
    const float v = w[((static_cast<long long>(co) * cin + ci) * kh + (kh - 1 - r)) * kw + (kw - 1 - s)];
    out[i] = to_w16<W>(v * (scale ? scale[co] : 1.0f));
  }
}

// fp32 [cout][kh][kw][cin] -> fp32 [cout][cin / groups][kh][kw]: the OIHW gradient of a (grouped) conv.  With
// groups > 1 the input is the DENSE weight gradient and only each output channel's own group of input channels is
// kept (the block diagonal; models/backbone/resnext.py:84-87).
__global__ void __launch_bounds__(256)
dw_unpack_kernel(const float* __restrict__ in, float* __restrict__ out, int cout, int cin, int kh, int kw, int groups) {
  const int cig = cin / groups, cog = cout / groups;
  const long long total = static_cast<long long>(cout) * cig * kh * kw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(i % kw);
    long long t = i / kw;
    const int r = static_cast<int>(t % kh);
    t /= kh;
    const int cl = static_cast<int>(t % cig);
    const int co = static_cast<int>(t / cig);
    const int ci = (co / cog) * cig + cl;
    out[i] = in[((static_cast<long long>(co) * kh + r) * kw + s) * cin + ci];
  }
}

// dw[c] += sum_m x[m][c]   (bf16 x, fp32 sums).  blockIdx.y = 64-channel chunk, blockIdx.x = row strip;
// thread = (8-channel group, row lane).
__global__ void __launch_bounds__(256)
colsum_kernel(const uint4* __restrict__ x, float* __restrict__ dw, long long rows, int c8, int rows_per_block) {
  const int cg = threadIdx.x & 7;
  const int rl = threadIdx.x >> 3;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = r0 + rl; r < r1; r += 32) {
    const uint4 v = __ldg(x + r * c8 + blockIdx.y * 8 + cg);
    acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x);
    acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
    acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z);
    acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
  }
  __shared__ float red[32][65];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.0f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    atomicAdd(dw + blockIdx.y * 64 + threadIdx.x, s);
  }
}

// Frozen-statistics BatchNorm affine gradients from stored tensors (training with bn_frozen=False):
//   y = gamma * xhat + beta  =>  dbeta[c] += sum_m g[m][c],  dgamma[c] += sum_m g[m][c] * (y[m][c] - beta[c]) / gamma[c]
// where g is the (ReLU-masked) gradient w.r.t. y.  y itself is not stored, but wherever g != 0 it equals the
// stored activation `a` (ReLU passed) minus, for the block's last conv, the residual operand `b`:
// y = a - b.  One pass over g, a (and b); per-tensor exponents applied.  Channels with gamma == 0 get
// dgamma = 0 (xhat cannot be recovered from y there).
struct BnAffineParams {
  const uint4* g;
  const uint4* a;
  const uint4* b;   // nullable
  long long rows;
  int c8, rows_per_block;
  int g_fp16, a_fp16, b_fp16;
  const TensorMeta* g_meta;
  const TensorMeta* a_meta;
  const TensorMeta* b_meta;
  const float* gamma;
  const float* beta;
  float* dgamma;
  float* dbeta;
};

__global__ void __launch_bounds__(256)
bn_affine_grad_kernel(const BnAffineParams p) {
  const int cg = threadIdx.x & 7;
  const int rl = threadIdx.x >> 3;
  const long long r0 = static_cast<long long>(blockIdx.x) * p.rows_per_block;
  const long long r1 = min(p.rows, r0 + p.rows_per_block);
  const float mg = ldexpf(1.0f, p.g_meta ? p.g_meta->e : 0);
  const float ma = ldexpf(1.0f, p.a_meta ? p.a_meta->e : 0);
  const float mb = ldexpf(1.0f, (p.b && p.b_meta) ? p.b_meta->e : 0);
  const bool gf = p.g_fp16 != 0, af = p.a_fp16 != 0, bf = p.b_fp16 != 0;
  float dot[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = r0 + rl; r < r1; r += 32) {
    const long long idx = r * p.c8 + blockIdx.y * 8 + cg;
    const uint4 vg = __ldg(p.g + idx);
    const uint4 va = __ldg(p.a + idx);
    const uint4 vb = p.b ? __ldg(p.b + idx) : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t wg[4] = {vg.x, vg.y, vg.z, vg.w}, wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float g0, g1, a0, a1, b0, b1;
      unpack16x2(wg[j], gf, g0, g1);
      unpack16x2(wa[j], af, a0, a1);
      unpack16x2(wb[j], bf, b0, b1);
      const float y0 = fmaf(a0, ma, -b0 * mb), y1 = fmaf(a1, ma, -b1 * mb);
      dot[2 * j] = fmaf(g0, y0, dot[2 * j]);
      dot[2 * j + 1] = fmaf(g1, y1, dot[2 * j + 1]);
      sum[2 * j] += g0;
      sum[2 * j + 1] += g1;
    }
  }
  __shared__ float red[2][32][65];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][rl][cg * 8 + j] = dot[j];
    red[1][rl][cg * 8 + j] = sum[j];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float d = 0.0f, s = 0.0f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      d += red[0][i][threadIdx.x];
      s += red[1][i][threadIdx.x];
    }
    const int c = blockIdx.y * 64 + threadIdx.x;
    d *= mg;
    s *= mg;
    const float gm = __ldg(p.gamma + c);
    atomicAdd(p.dbeta + c, s);
    if (gm != 0.0f) atomicAdd(p.dgamma + c, (d - __ldg(p.beta + c) * s) / gm);
  }
}

// Backward of the stem's max-pool (3x3 / stride 2 / pad 1) fused with the ReLU backward of the stem output:
//   dS[n][ih][iw][c] = (S > 0) * sum over the (up to four) pooling windows whose FIRST maximum (row-major
//   scan, the tie rule of aten::max_pool2d_with_indices) is (ih, iw) of g[window]
// S: stem output [n][h][w][C] (16-bit, any common exponent: only compared), g: gradient w.r.t. the pooled tensor
// [n][ho][wo][C] (16-bit, exponent from g_meta), dS: bf16 true values.  One thread per (input pixel, 8 channels).
struct MaxpoolBwdParams {
  const uint4* s;
  const uint4* g;
  uint4* ds;
  int n, h, w, c8, ho, wo;
  int s_fp16, g_fp16;
  const TensorMeta* g_meta;
};

__global__ void __launch_bounds__(256)
maxpool3x3s2_bwd_kernel(const MaxpoolBwdParams p) {
  const long long total = static_cast<long long>(p.n) * p.h * p.w * p.c8;
  const float mg = ldexpf(1.0f, p.g_meta ? p.g_meta->e : 0);
  const bool sf = p.s_fp16 != 0, gf = p.g_fp16 != 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % p.c8);
    long long t = i / p.c8;
    const int iw = static_cast<int>(t % p.w);
    t /= p.w;
    const int ih = static_cast<int>(t % p.h);
    const int img = static_cast<int>(t / p.h);
    const uint4 self = __ldg(p.s + i);
    const uint32_t sw[4] = {self.x, self.y, self.z, self.w};
    float me[8], acc[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unpack16x2(sw[j], sf, me[2 * j], me[2 * j + 1]);
      acc[2 * j] = acc[2 * j + 1] = 0.0f;
    }
    // windows (oh, ow) that contain (ih, iw): 2*oh - 1 <= ih <= 2*oh + 1
    for (int oh = (ih >> 1); oh <= ((ih + 1) >> 1); ++oh) {
      if (oh < 0 || oh >= p.ho) continue;
      for (int ow = (iw >> 1); ow <= ((iw + 1) >> 1); ++ow) {
        if (ow < 0 || ow >= p.wo) continue;
        // is (ih, iw) the first maximum of this window, per channel?
        bool win[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) win[j] = true;
        for (int dy = 0; dy < 3; ++dy) {
          const int yh = 2 * oh - 1 + dy;
          if (yh < 0 || yh >= p.h) continue;
          for (int dx = 0; dx < 3; ++dx) {
            const int xw = 2 * ow - 1 + dx;
            if (xw < 0 || xw >= p.w || (yh == ih && xw == iw)) continue;
            const bool before = (yh < ih) || (yh == ih && xw < iw);  // scanned earlier: wins ties
            const uint4 o = __ldg(p.s + ((static_cast<long long>(img) * p.h + yh) * p.w + xw) * p.c8 + cg);
            const uint32_t ow4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a, b;
              unpack16x2(ow4[j], sf, a, b);
              if (before ? (a >= me[2 * j]) : (a > me[2 * j])) win[2 * j] = false;
              if (before ? (b >= me[2 * j + 1]) : (b > me[2 * j + 1])) win[2 * j + 1] = false;
            }
          }
        }
        const uint4 gv = __ldg(p.g + ((static_cast<long long>(img) * p.ho + oh) * p.wo + ow) * p.c8 + cg);
        const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a, b;
          unpack16x2(gw[j], gf, a, b);
          if (win[2 * j]) acc[2 * j] += a;
          if (win[2 * j + 1]) acc[2 * j + 1] += b;
        }
      }
    }
    uint32_t o4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o4[j] = pack_bf16x2(me[2 * j] > 0.0f ? acc[2 * j] * mg : 0.0f, me[2 * j + 1] > 0.0f ? acc[2 * j + 1] * mg : 0.0f);
    p.ds[i] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
  }
}

// Weight gradient of the 7x7 / stride 2 / pad 3 stem conv (3 -> 64 channels; CUDA cores: 40 GFLOP per batch of 8):
//   dW[co][c][r][s] += scale[co] * sum_{n,p,q} g[n][p][q][co] * img[n][2p-3+r][2q-3+s][c]
// img = the padded NHWC4 bf16 staging TDET_OP_PREP wrote (pixel (ih, iw) at (ih+3, iw+3)), g = bf16 [n][ho][wo][64].
// A block walks tiles of 4 x 64 output pixels, staging the g tile and the (13 x 133)-pixel image patch in shared
// memory as fp32; thread (co = t & 63, grp = t >> 6) accumulates the 37 filter taps k = grp + 4j of channel co.
constexpr int kSwTileH = 4, kSwTileW = 64;
constexpr int kSwPatchH = 2 * kSwTileH + 5, kSwPatchW = 2 * kSwTileW + 5;
constexpr int kSwSmemBytes = (kSwTileH * kSwTileW * 64 + kSwPatchH * kSwPatchW * 4) * 4;

__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const uint2* __restrict__ img, const __nv_bfloat16* __restrict__ g, const float* __restrict__ scale,
                  float* __restrict__ dw, int n, int ho, int wo, int hp, int wp) {
  extern __shared__ float sw_smem[];
  float* sg = sw_smem;                               // [256 pixels][64 channels]
  float* si = sw_smem + kSwTileH * kSwTileW * 64;    // [13][133][4]
  const int co = threadIdx.x & 63, grp = threadIdx.x >> 6;
  int koff[37];
  float acc[37];
#pragma unroll
  for (int j = 0; j < 37; ++j) {
    const int k = grp + 4 * j;  // index into [c][r][s] (147 valid)
    const int c = k / 49, r = (k % 49) / 7, s = k % 7;
    koff[j] = k < 147 ? (r * kSwPatchW + s) * 4 + c : 0;
    acc[j] = 0.0f;
  }
  const int tiles_w = (wo + kSwTileW - 1) / kSwTileW, tiles_h = (ho + kSwTileH - 1) / kSwTileH;
  const long long tiles = static_cast<long long>(n) * tiles_h * tiles_w;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tw = static_cast<int>(tile % tiles_w);
    long long t = tile / tiles_w;
    const int th = static_cast<int>(t % tiles_h);
    const int im = static_cast<int>(t / tiles_h);
    const int p0 = th * kSwTileH, q0 = tw * kSwTileW;
    __syncthreads();
    // g tile (zero beyond the image)
    for (int i = threadIdx.x; i < kSwTileH * kSwTileW * 64; i += 256) {
      const int c = i & 63, pix = i >> 6;
      const int pl = pix / kSwTileW, ql = pix - pl * kSwTileW;
      const int pp = p0 + pl, qq = q0 + ql;
      sg[i] = (pp < ho && qq < wo)
                  ? __bfloat162float(g[((static_cast<long long>(im) * ho + pp) * wo + qq) * 64 + c]) : 0.0f;
    }
    // image patch: staged rows 2*p0 .. 2*p0 + 12, columns 2*q0 .. 2*q0 + 132 (zero beyond the staging)
    for (int i = threadIdx.x; i < kSwPatchH * kSwPatchW; i += 256) {
      const int yl = i / kSwPatchW, xl = i - yl * kSwPatchW;
      const int y = 2 * p0 + yl, x = 2 * q0 + xl;
      uint2 v = make_uint2(0u, 0u);
      if (y < hp && x < wp) v = __ldg(img + (static_cast<long long>(im) * hp + y) * wp + x);
      si[4 * i + 0] = bf16_lo(v.x);
      si[4 * i + 1] = bf16_hi(v.x);
      si[4 * i + 2] = bf16_lo(v.y);
      si[4 * i + 3] = 0.0f;
    }
    __syncthreads();
    for (int pix = 0; pix < kSwTileH * kSwTileW; ++pix) {
      const int pl = pix / kSwTileW, ql = pix - pl * kSwTileW;
      const float gv = sg[pix * 64 + co];
      const float* base = si + (2 * pl * kSwPatchW + 2 * ql) * 4;
#pragma unroll
      for (int j = 0; j < 37; ++j) acc[j] = fmaf(gv, base[koff[j]], acc[j]);
    }
  }
  const float sc = scale ? scale[co] : 1.0f;
#pragma unroll
  for (int j = 0; j < 37; ++j) {
    const int k = grp + 4 * j;
    if (k < 147) atomicAdd(dw + co * 147 + k, acc[j] * sc);
  }
}

__device__ __forceinline__ uint32_t add4_bf16x2(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return pack_bf16x2(bf16_lo(a) + bf16_lo(b) + bf16_lo(c) + bf16_lo(d),
                     bf16_hi(a) + bf16_hi(b) + bf16_hi(c) + bf16_hi(d));
}

// y[n][i][j][:] = sum over the 2x2 block of x (adjoint of nearest-x2 upsample)
__global__ void __launch_bounds__(256)
sumpool2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int c8, int ho, int wo) {
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    long long t = i / c8;
    const int ow = static_cast<int>(t % wo);
    t /= wo;
    const int oh = static_cast<int>(t % ho);
    const int img = static_cast<int>(t / ho);
    const uint4* px = x + ((static_cast<long long>(img) * h + 2 * oh) * w + 2 * ow) * c8 + cg;
    const uint4 a = __ldg(px), b = __ldg(px + c8), c = __ldg(px + static_cast<long long>(w) * c8),
                d = __ldg(px + static_cast<long long>(w) * c8 + c8);
    uint4 o;
    o.x = add4_bf16x2(a.x, b.x, c.x, d.x);
    o.y = add4_bf16x2(a.y, b.y, c.y, d.y);
    o.z = add4_bf16x2(a.z, b.z, c.z, d.z);
    o.w = add4_bf16x2(a.w, b.w, c.w, d.w);
    y[i] = o;
  }
}

// y[n][u][v][:] = (u, v even) ? x[n][u/2][v/2][:] : 0   (adjoint of a stride-2 subsample)
__global__ void __launch_bounds__(256)
dilate2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int c8, int ho, int wo) {
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    long long t = i / c8;
    const int ow = static_cast<int>(t % wo);
    t /= wo;
    const int oh = static_cast<int>(t % ho);
    const int img = static_cast<int>(t / ho);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (((oh | ow) & 1) == 0)
      o = __ldg(x + ((static_cast<long long>(img) * h + (oh >> 1)) * w + (ow >> 1)) * c8 + cg);
    y[i] = o;
  }
}

// y = (x * 2^ex + res * 2^er) * (mask != 0), stored as y_dtype * 2^ey.  Formats are per tensor (bf16 or fp16);
// ey is either 0 or, when y_meta/scaled is set, chosen so that amax(x) + amax(res) < 2^15 (fp16-safe), in which
// case that bound is also recorded as y's amax (an upper bound is all downstream exponent choices need).
struct AddMaskParams {
  const uint4* x;
  const uint4* res;
  const uint4* mask;
  uint4* y;
  long long total;  // uint4 groups
  int x_fp16, res_fp16, y_fp16, scaled;
  const TensorMeta* x_meta;
  const TensorMeta* res_meta;
  TensorMeta* y_meta;
};

__global__ void __launch_bounds__(256)
add_mask_kernel(const AddMaskParams p) {
  const int ex = p.x_meta ? p.x_meta->e : 0;
  const int er = (p.res && p.res_meta) ? p.res_meta->e : 0;
  int ey = 0;
  if (p.scaled) {
    float bound = p.x_meta ? __uint_as_float(p.x_meta->amax_bits) : 0.0f;
    if (p.res && p.res_meta) bound += __uint_as_float(p.res_meta->amax_bits);
    if (bound > 0.0f && bound < 3.0e38f) ey = ilogbf(bound) - 14;
    ey = max(-100, min(100, ey));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      p.y_meta->e = ey;
      p.y_meta->amax_bits = __float_as_uint(bound);
    }
  }
  const float mx = ldexpf(1.0f, ex - ey), mr = ldexpf(1.0f, er - ey);
  const bool xf = p.x_fp16 != 0, rf = p.res_fp16 != 0, yf = p.y_fp16 != 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < p.total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 a = __ldg(p.x + i);
    const uint4 b = p.res ? __ldg(p.res + i) : make_uint4(0u, 0u, 0u, 0u);
    const uint4 m = p.mask ? __ldg(p.mask + i) : make_uint4(0x7FFF7FFFu, 0x7FFF7FFFu, 0x7FFF7FFFu, 0x7FFF7FFFu);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, mw[4] = {m.x, m.y, m.z, m.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float alo, ahi, blo, bhi;
      unpack16x2(aw[j], xf, alo, ahi);
      unpack16x2(bw[j], rf, blo, bhi);
      float lo = fmaf(blo, mr, alo * mx), hi = fmaf(bhi, mr, ahi * mx);
      if ((mw[j] & 0x00007FFFu) == 0u) lo = 0.0f;
      if ((mw[j] & 0x7FFF0000u) == 0u) hi = 0.0f;
      ow[j] = pack16x2(lo, hi, yf);
    }
    p.y[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// ---- GroupNorm (use_gn=True backbones, necks with normalize=GN; models/utils/layers.py:50-54,138-154) --------------
// GroupNorm statistics depend on the data, so it cannot be folded into a GEMM epilogue: the conv writes its raw output
// (16-bit, per-tensor exponent), gn_stats_kernel reduces sum / sum of squares of the TRUE values per (image, group) --
// fp32 partial sums per thread, shared-memory bins per block, one atomicAdd per (block, group, moment) -- and
// gn_apply_kernel writes act((x - mean) * rstd * gamma + beta [+ residual] [+ nearest-x2 upsampled coarse level]).
// Both are HBM-bound passes with 16-byte accesses.  Groups are channel-contiguous in NHWC; a thread keeps ONE
// 8-channel column of the tensor for its whole loop (the grid stride is a multiple of c / 8), so its eight channel
// sums stay in registers.
constexpr int kGnMaxGroups = 64;

constexpr int kGnStatBlocks = TDET_GN_STAT_BLOCKS;   // partial-sum rows per image (at most; include/tdet_b200.h)

// Deterministic: no atomics anywhere.  A thread owns eight fixed channels; the block's 256 x 16 partial sums are
// combined per group by one thread in a fixed order and written to the block's own row of `stats`
// ([n][kGnStatBlocks][groups][2]); gn_apply_kernel adds the rows of its image in order.
__global__ void __launch_bounds__(256)
gn_stats_kernel(const uint4* __restrict__ x, float* __restrict__ stats, int hw, int c8, int groups, int x_fp16,
                const TensorMeta* __restrict__ x_meta) {
  // grid: (blocks per image <= kGnStatBlocks, n); blockDim.x = 256, a multiple of c8 (c8 in {8, 16, ..., 256})
  __shared__ float part[16][257];   // [2 * channel-in-octet + {sum, sumsq}][thread], padded against bank conflicts
  const int img = blockIdx.y;
  const long long total = static_cast<long long>(hw) * c8;
  const uint4* xi = x + static_cast<long long>(img) * total;
  const float mul = ldexpf(1.0f, x_meta ? x_meta->e : 0);
  const bool xf = x_fp16 != 0;
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.0f;
  const long long start = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;   // multiple of c8
  for (long long i = start; i < total; i += stride) {
    const uint4 v = __ldg(xi + i);
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo, hi;
      unpack16x2(w4[j], xf, lo, hi);
      lo *= mul;
      hi *= mul;
      s1[2 * j] += lo;
      s2[2 * j] = fmaf(lo, lo, s2[2 * j]);
      s1[2 * j + 1] += hi;
      s2[2 * j + 1] = fmaf(hi, hi, s2[2 * j + 1]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    part[2 * e][threadIdx.x] = s1[e];
    part[2 * e + 1][threadIdx.x] = s2[e];
  }
  __syncthreads();
  // thread (g, which) adds, in a fixed order, every partial that belongs to group g: thread t holds channels
  // [(t % c8) * 8, +8) (blockDim.x is a multiple of c8, so is the grid stride)
  if (threadIdx.x < 2 * groups) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    const int cg = (c8 * 8) / groups;                   // channels per group
    const int ch_lo = g * cg, ch_hi = ch_lo + cg;
    float acc = 0.0f;
    for (int oct = ch_lo / 8; oct * 8 < ch_hi; ++oct) {
      const int e_lo = max(ch_lo - oct * 8, 0), e_hi = min(ch_hi - oct * 8, 8);
      for (int t = oct; t < 256; t += c8)
        for (int e = e_lo; e < e_hi; ++e) acc += part[2 * e + which][t];
    }
    stats[((static_cast<long long>(img) * kGnStatBlocks + blockIdx.x) * groups + g) * 2 + which] = acc;
  }
}

struct GnApplyParams {
  const uint4* x;          // raw conv output [n][h][w][c]
  const uint4* res;        // nullable, same shape
  const uint4* coarse;     // nullable, [n][h/2][w/2][c]: nearest-x2 upsample-add (FPN top-down path after the norm)
  uint4* y;
  const float* stats;      // [n][kGnStatBlocks][groups][2] partial sum, sum of squares of the true values
  int stat_blocks;         // rows gn_stats_kernel wrote per image
  const float* gamma;
  const float* beta;
  int h, w, c8, groups;
  float eps;
  int relu;
  int x_fp16, res_fp16, coarse_fp16, y_fp16;
  const TensorMeta* x_meta;
  const TensorMeta* res_meta;
  const TensorMeta* coarse_meta;
  TensorMeta* y_meta;      // nullable: receives max |y| (exponent 0)
};

__global__ void __launch_bounds__(256)
gn_apply_kernel(const GnApplyParams p) {
  __shared__ float s_mean[kGnMaxGroups], s_rstd[kGnMaxGroups];
  const int img = blockIdx.y;
  const int cg = (p.c8 * 8) / p.groups;
  const float cnt = static_cast<float>(p.h) * p.w * cg;
  for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {
    float s1 = 0.0f, s2 = 0.0f;
    for (int b = 0; b < p.stat_blocks; ++b) {   // fixed order: every block of the image computes the same statistics
      const float2 v = __ldg(reinterpret_cast<const float2*>(
          p.stats + ((static_cast<long long>(img) * kGnStatBlocks + b) * p.groups + g) * 2));
      s1 += v.x;
      s2 += v.y;
    }
    const float mean = s1 / cnt;
    const float var = fmaxf(s2 / cnt - mean * mean, 0.0f);   // biased variance, as nn.GroupNorm
    s_mean[g] = mean;
    s_rstd[g] = rsqrtf(var + p.eps);
  }
  __syncthreads();
  const long long total = static_cast<long long>(p.h) * p.w * p.c8;
  const long long base = static_cast<long long>(img) * total;
  const float mx = ldexpf(1.0f, p.x_meta ? p.x_meta->e : 0);
  const float mr = ldexpf(1.0f, (p.res && p.res_meta) ? p.res_meta->e : 0);
  const float mc = ldexpf(1.0f, (p.coarse && p.coarse_meta) ? p.coarse_meta->e : 0);
  const bool xf = p.x_fp16 != 0, rf = p.res_fp16 != 0, cf = p.coarse_fp16 != 0, yf = p.y_fp16 != 0;
  const long long start = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;   // multiple of c8
  const int c0 = static_cast<int>(start % p.c8) * 8;
  float a[8], b[8];   // y = x * a + b for this thread's eight channels
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int g = (c0 + e) / cg;
    const float ga = __ldg(p.gamma + c0 + e) * s_rstd[g];
    a[e] = ga * mx;
    b[e] = __ldg(p.beta + c0 + e) - s_mean[g] * ga;
  }
  float amax = 0.0f;
  const int hc = p.h >> 1, wc = p.w >> 1;
  for (long long i = start; i < total; i += stride) {
    const uint4 v = __ldg(p.x + base + i);
    uint4 r = make_uint4(0u, 0u, 0u, 0u), cz = make_uint4(0u, 0u, 0u, 0u);
    if (p.res) r = __ldg(p.res + base + i);
    if (p.coarse) {
      const long long pix = i / p.c8;
      const int px = static_cast<int>(pix % p.w), py = static_cast<int>(pix / p.w);
      cz = __ldg(p.coarse + ((static_cast<long long>(img) * hc + (py >> 1)) * wc + (px >> 1)) * p.c8 + (i % p.c8));
    }
    const uint32_t vw[4] = {v.x, v.y, v.z, v.w}, rw[4] = {r.x, r.y, r.z, r.w}, cw[4] = {cz.x, cz.y, cz.z, cz.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo, hi, rlo, rhi, clo, chi;
      unpack16x2(vw[j], xf, lo, hi);
      unpack16x2(rw[j], rf, rlo, rhi);
      unpack16x2(cw[j], cf, clo, chi);
      lo = fmaf(lo, a[2 * j], b[2 * j]);
      hi = fmaf(hi, a[2 * j + 1], b[2 * j + 1]);
      lo = fmaf(rlo, mr, fmaf(clo, mc, lo));
      hi = fmaf(rhi, mr, fmaf(chi, mc, hi));
      if (p.relu) {
        lo = fmaxf(lo, 0.0f);
        hi = fmaxf(hi, 0.0f);
        if (p.relu == 2) {
          lo = fminf(lo, 6.0f);
          hi = fminf(hi, 6.0f);
        }
      }
      if (yf) {   // plain fp16 (exponent 0): saturate instead of overflowing to infinity
        lo = fminf(fmaxf(lo, -65504.0f), 65504.0f);
        hi = fminf(fmaxf(hi, -65504.0f), 65504.0f);
      }
      amax = fmaxf(amax, fmaxf(fabsf(lo), fabsf(hi)));
      ow[j] = pack16x2(lo, hi, yf);
    }
    p.y[base + i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
  if (p.y_meta) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&p.y_meta->amax_bits, __float_as_uint(amax));
  }
}

// Stride-2 3x3 dgrad as four parity classes (TDET_OP_PARITY_MERGE): dx[2i+a][2j+b] only receives the filter taps of
// matching parity, so the input gradient is four small stride-1 convs over the COARSE gradient g --
//   (0,0): 1x1, tap (1,1)      (0,1): 1x2, taps (1,0),(1,2)      (1,0): 2x1, taps (0,1),(2,1)      (1,1): 2x2, the corners
// of the rotated kernel -- a quarter of the MMAs of a 3x3 conv over the zero-inserted gradient.  The multi-tap classes
// run with pad 1 (symmetric padding is all the conv op has), which shifts their outputs by one pixel and adds a
// border: p01 is [n][hc+2][wc+1], p10 [n][hc+1][wc+2], p11 [n][hc+1][wc+1], p00 [n][hc][wc].  This kernel interleaves
// the four results into dx [n][h][w][c], applies the ReLU-backward mask and converts formats / exponents.
struct ParityMergeParams {
  const uint4* p[4];         // p00, p01, p10, p11
  const TensorMeta* pm[4];   // nullable
  const uint4* mask;         // nullable, [n][h][w][c]
  uint4* y;
  TensorMeta* y_meta;
  int n, h, w, c8, hc, wc;
  int p_fp16, y_fp16, scaled;
};

__global__ void __launch_bounds__(256)
parity_merge_kernel(const ParityMergeParams q) {
  int e[4];
  float bound = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    e[k] = q.pm[k] ? q.pm[k]->e : 0;
    if (q.pm[k]) bound = fmaxf(bound, __uint_as_float(q.pm[k]->amax_bits));
  }
  int ey = 0;
  if (q.scaled) {
    if (bound > 0.0f && bound < 3.0e38f) ey = ilogbf(bound) - 14;
    ey = max(-100, min(100, ey));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      q.y_meta->e = ey;
      q.y_meta->amax_bits = __float_as_uint(bound);
    }
  }
  float mul[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) mul[k] = ldexpf(1.0f, e[k] - ey);
  const bool pf = q.p_fp16 != 0, yf = q.y_fp16 != 0;
  const unsigned total = static_cast<unsigned>(q.n) * q.h * q.w * q.c8;
#pragma unroll 4
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned cg = i % q.c8;
    unsigned t = i / q.c8;
    const int v = static_cast<int>(t % q.w);
    t /= q.w;
    const int u = static_cast<int>(t % q.h);
    const int img = static_cast<int>(t / q.h);
    const int a = u & 1, b = v & 1, k = 2 * a + b;
    const int off = k ? 1 : 0;                                       // the padded classes are shifted by one pixel
    const int ph = q.hc + (k == 1 ? 2 : k ? 1 : 0), pw = q.wc + (k == 2 ? 2 : k ? 1 : 0);
    const long long src = ((static_cast<long long>(img) * ph + (u >> 1) + off) * pw + (v >> 1) + off) * q.c8 + cg;
    const uint4 x = __ldg(q.p[k] + src);
    const uint4 m = q.mask ? __ldg(q.mask + i) : make_uint4(0x7FFF7FFFu, 0x7FFF7FFFu, 0x7FFF7FFFu, 0x7FFF7FFFu);
    const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, mw[4] = {m.x, m.y, m.z, m.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo, hi;
      unpack16x2(xw[j], pf, lo, hi);
      lo *= mul[k];
      hi *= mul[k];
      if ((mw[j] & 0x00007FFFu) == 0u) lo = 0.0f;
      if ((mw[j] & 0x7FFF0000u) == 0u) hi = 0.0f;
      ow[j] = pack16x2(lo, hi, yf);
    }
    q.y[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// meta->amax_bits = max |x| * 2^e_x  (atomicMax on the bit pattern; the plan zeroes the arena first)
__global__ void __launch_bounds__(256)
amax_kernel(const uint4* __restrict__ x, long long total, int x_fp16, const TensorMeta* x_meta, TensorMeta* meta) {
  float amax = 0.0f;
  const bool xf = x_fp16 != 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 a = __ldg(x + i);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo, hi;
      unpack16x2(aw[j], xf, lo, hi);
      amax = fmaxf(amax, fmaxf(fabsf(lo), fabsf(hi)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.0f)
    atomicMax(&meta->amax_bits, __float_as_uint(ldexpf(amax, x_meta ? x_meta->e : 0)));
}

}  // namespace tdet
