/*
 * tdet_b200.h -- C ABI of libtdet_b200.so: the B200 (sm_100a) implementation of the
 * ResNet + FPN feature-extraction hot path of TCGGroup/Torch_Detection.
 *
 * The reference implements this path entirely as torch.nn modules (there is no native code and no
 * FFI in the reference), so "what the reference's FFI for this path would bind" is the set of
 * ATen ops its modules dispatch to.  Each entry point / op kind below cites the reference call site
 * it replaces (paths relative to the reference root):
 *
 *   TDET_OP_PREP       image hand-off: NCHW fp32/bf16 batch from datasets/loader/collate.py:42-63
 *                      -> padded NHWC4 bf16 staging for the stem (no reference op; layout change)
 *   TDET_OP_STEM       conv1 7x7/2 p3 + bn1 + relu        models/backbone/resnet.py:254-257
 *                      (conv7x7_group models/utils/layers.py:35-47, norm_layer :50-54)
 *   TDET_OP_MAXPOOL    nn.MaxPool2d(3, 2, 1)              models/backbone/resnet.py:218,258
 *   TDET_OP_CONV       every other conv with its fused epilogue:
 *                        Bottleneck / BasicBlock conv+BN(+ReLU)(+residual add+ReLU)
 *                                                         models/backbone/resnet.py:42-59,97-119
 *                        downsample 1x1 stride s + BN     models/backbone/resnet.py:129-136
 *                        FPN lateral 1x1 + bias (+ nearest-x2 upsample-add of the coarser level)
 *                                                         models/necks/fpn.py:92-101
 *                        FPN output 3x3 p1 + bias         models/necks/fpn.py:106-108
 *                        FPN extra stride-2 3x3 (+ReLU on input done by producer) fpn.py:118-124
 *   TDET_OP_SUBSAMPLE  F.max_pool2d(x, 1, stride=2)       models/necks/fpn.py:114-116
 *   tdet_pack_conv_weight / tdet_pack_stem_weight / tdet_fold_bn / tdet_conv_bound_consts
 *                      derive the kernels' operand formats from the reference's fp32 OIHW
 *                      parameters and BatchNorm2d buffers (state_dict layout of resnet.py/fpn.py)
 *
 * Conventions
 *   - extern "C", plain C types only; no torch types cross this boundary.
 *   - Every function returns 0 (TDET_OK) or a negative tdet_status; nothing throws.
 *     tdet_last_error() returns a thread-local, human-readable message for the last failure.
 *   - The library never allocates or frees caller tensors.  All device buffers (activations,
 *     packed weights, outputs) are owned by the caller (torch's caching allocator on the Python side).
 *   - All GPU work is enqueued asynchronously on the caller's stream (cudaStream_t passed as void*);
 *     no hidden synchronisation (tdet_plan_run_timed is the one documented exception).
 *   - Activations are dense NHWC 16-bit.  At the module boundary (network input, returned feature
 *     maps) they are plain bf16.  Internal tensors may instead be fp16 significands with ONE
 *     power-of-two exponent per tensor (tdet_tensor_meta): 3 more mantissa bits than bf16 at the same
 *     tensor-core rate, with the exponent chosen on the device from a rigorous bound so that no
 *     input can overflow (see DESIGN.md "Numerics").  tcgen05 requires A and B of one MMA to share a
 *     format, so a conv's weights are packed in the format of its input tensor.
 *   - Conv weights are packed [Cout][kh][kw][Cin] (K-major rows, the tcgen05 B operand);
 *     per-channel epilogue parameters are fp32.
 *   - A plan is bound to one device, is not re-entrant (one in-flight run per plan); distinct plans
 *     are independent.
 *   - There is no CPU fallback: on a device that is not sm_100 every compute entry point returns
 *     TDET_ERR_UNSUPPORTED_DEVICE.
 */
#ifndef TDET_B200_H_
#define TDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDET_ABI_VERSION 11
#define TDET_GN_STAT_BLOCKS 128 /* rows per image of the TDET_OP_GN_STATS output */

typedef enum tdet_status {
  TDET_OK = 0,
  TDET_ERR_INVALID_ARGUMENT = -1,
  TDET_ERR_UNSUPPORTED_SHAPE = -2,
  TDET_ERR_UNSUPPORTED_DEVICE = -3,
  TDET_ERR_CUDA = -4,
  TDET_ERR_DRIVER = -5,
  TDET_ERR_OUT_OF_MEMORY = -6
} tdet_status;

typedef enum tdet_op_kind {
  TDET_OP_PREP = 0,
  TDET_OP_STEM = 1,
  TDET_OP_MAXPOOL = 2,
  TDET_OP_CONV = 3,
  TDET_OP_SUBSAMPLE = 4,
  /* ---- training path (config 4: frozen BN, hand-written dgrad / wgrad) ---- */
  TDET_OP_WGRAD = 5,     /* weight gradient of a conv (implicit GEMM over pixels) */
  TDET_OP_DW_UNPACK = 6, /* fp32 [cout][kh][kw][cin] wgrad accumulator -> fp32 OIHW parameter gradient */
  TDET_OP_COLSUM = 7,    /* bias gradient: per-channel sum over all pixels */
  TDET_OP_SUMPOOL2 = 8,  /* adjoint of the nearest-x2 upsample (fpn.py:100-101): 2x2 sum pool */
  TDET_OP_DILATE2 = 9,   /* adjoint of a stride-2 subsample: zero-insertion upsample to (ho, wo) */
  TDET_OP_ADD_MASK = 10, /* y = (x + residual) * (mask > 0): gradient merge / ReLU backward */
  TDET_OP_ZERO = 11,     /* cudaMemsetAsync(y, 0, x_stride[0] bytes): gradient accumulators */
  TDET_OP_AMAX = 12,     /* y_meta->amax_bits = max |x| (true values): bound input for tensors produced elsewhere */
  TDET_OP_BN_AFFINE_GRAD = 13, /* gamma / beta gradients of a frozen-statistics BatchNorm from stored tensors */
  TDET_OP_SPLIT_COMBINE = 14, /* y (fp32 [n][h][w][cin]) = hi + lo of a split-precision tensor x ([n][h][w][2*cin] bf16) */
  TDET_OP_MAXPOOL_BWD = 15,  /* backward of MaxPool2d(3, 2, 1) fused with the ReLU backward of its input (stem) */
  TDET_OP_STEM_WGRAD = 16,   /* weight gradient of the 7x7/2 stem conv from the staged image */
  TDET_OP_PARITY_MERGE = 17, /* interleave the four parity-class results of a stride-2 3x3 dgrad (+ ReLU mask) */
  TDET_OP_BOTTLENECK_TAIL = 18, /* fused conv2 (3x3) -> conv3 (1x1) + residual + ReLU [-> the next block's conv1] */
  TDET_OP_GN_STATS = 19,     /* GroupNorm statistics: per (image, group) sum and sum of squares */
  TDET_OP_GN_APPLY = 20      /* GroupNorm normalise + affine (+ residual) (+ upsampled coarse level) (+ ReLU) */
} tdet_op_kind;

typedef enum tdet_dtype { TDET_BF16 = 0, TDET_F32 = 1, TDET_F16 = 2, TDET_U8 = 3 } tdet_dtype;

enum {
  TDET_FLAG_RELU = 1,       /* ReLU after scale/shift/residual (resnet.py:48,58,103,108,118) */
  TDET_FLAG_SCALED_OUT = 2, /* y is stored with a device-chosen power-of-two exponent (needs y_meta,
                               x_meta and bound_consts) */
  TDET_FLAG_COARSE_PARITY = 4, /* coarse[i][j] is added at y[2i][2j] only (adjoint of a stride-2 1x1 conv,
                                  resnet.py:129-136) instead of nearest-x2 upsampled; hc = (ho+1)/2 */
  TDET_FLAG_SPLIT = 8,      /* split precision (the fp32-I/O mode, <= 1e-4 vs the fp32 reference): every
                               activation tensor is a bf16 pair value = hi + lo stored as 2*C channels
                               [hi | lo] (PREP / STEM staging: two image planes), conv weights come from
                               tdet_pack_conv_weight_split; the GEMM accumulates hi*hi + lo*hi + hi*lo in
                               fp32 and the epilogue splits its fp32 result again.  cin / cout stay the
                               LOGICAL channel counts.  Valid on PREP, STEM, MAXPOOL, CONV. */
  TDET_FLAG_POOL = 16,      /* TDET_OP_STEM only: the kernel also applies the 3x3/2 p1 max-pool that follows the
                               stem (resnet.py:218,258) and y is the POOLED tensor [n][hp][wp][64], hp =
                               (ho - 1) / 2 + 1; the stem's own output never reaches memory.  Needs
                               TDET_FLAG_RELU (out-of-image window positions are taken as 0). */
  TDET_FLAG_SCALED_OUT2 = 64, /* TDET_OP_BOTTLENECK_TAIL: y2 is stored with a device-chosen exponent (needs y2_meta) */
  TDET_FLAG_RELU6 = 128,    /* TDET_OP_CONV / TDET_OP_GN_APPLY: y = min(max(v, 0), 6) -- ConvModule(activation='relu6'),
                               models/utils/layers.py:114-119.  Plain outputs only (not with TDET_FLAG_SCALED_OUT). */
  TDET_FLAG_REVERSE = 256,  /* TDET_OP_CONV: the persistent kernel walks its output tiles from the last to the first.
                               Results are identical; tdet_plan_create sets it on every second conv of a plan
                               (TDET_SERPENTINE, default on) so that a launch starts on the rows its producer wrote
                               last, i.e. the ones still resident in L2.  No reference counterpart (scheduling only). */
  TDET_FLAG_DUAL = 32       /* TDET_OP_CONV, 1x1 / stride 1 / cout % 256 == 0 only: a SECOND input,
                               y = act( ([x | x2'] * wgt^T) * scale + shift ),  x2' = x2 sampled with stride2,
                               i.e. wgt is the K-concatenation [cout][cin + cin2] of two 1x1 weight matrices and both
                               GEMMs run in one launch into one accumulator.  This is a stage's first Bottleneck:
                               conv3 + bn3 and the projection shortcut downsample(x) + its BN + the residual add +
                               ReLU (resnet.py:110-118 with :129-136) without the shortcut tensor ever reaching memory
                               (tdet_pack_conv_weight_scaled folds each BatchNorm's scale into its half of wgt, `shift`
                               is the sum of the two BN shifts, scale = NULL).  x and x2 are plain tensors of ONE
                               16-bit format (exponent 0; their metas only carry |max|); no residual / coarse / mask. */
};

/* Per-tensor metadata living in device memory (8 bytes): true value = stored * 2^e; amax_bits is
 * the IEEE-754 bit pattern of the true max |value| (accumulated with atomicMax, so the caller must
 * zero it before the producing op runs; tdet_plan_run does that for the arena it is told about). */
typedef struct tdet_tensor_meta {
  int32_t e;
  uint32_t amax_bits;
} tdet_tensor_meta;

/*
 * One step of the path.  Unused fields must be zero / NULL.
 *
 * TDET_OP_PREP      x: logical (n, 3, hc, wc) image batch of x_dtype (F32/BF16/U8) with element strides
 *                   x_stride[] = {n, c, h, w} (NCHW-contiguous, channels_last, or an HWC uint8 batch viewed
 *                   as NCHW all work); hc/wc = valid extent (0 = h/w), zero-padded to the (h, w) the stem
 *                   sees (pad-to-size-divisor); optional scale/shift[3]: v*scale[c] + shift[c], the data
 *                   layer's (v - mean)/std (datasets/dataset_transforms.py:29-44 steps 2 and 5).
 *                   y: 16-bit [n][hp][wp][4] with (hp, wp) = tdet_stem_staging_dims(ho, wo) (ho/wo = stem output
 *                   size), the image at offset (3,3), zero elsewhere, channel 3 zero.  y_dtype = TDET_BF16 (default),
 *                   or TDET_F16: three more significand bits, |v| saturates at 65504 (used ahead of GroupNorm chains,
 *                   which amplify the staging's rounding); not with TDET_FLAG_SPLIT.
 *                   y_meta (optional): receives the image's |max| (exponent 0).
 * TDET_OP_STEM      x: the PREP output (x_dtype = its format); wgt: tdet_pack_stem_weight output of the same format;
 *                   scale/shift: folded bn1; y: [n][ho][wo][64] of y_dtype (with TDET_FLAG_POOL: the max-pooled
 *                   [n][(ho-1)/2+1][(wo-1)/2+1][64]).  h,w = image size.
 * TDET_OP_MAXPOOL   x: [n][h][w][cin] of x_dtype; y: [n][ho][wo][cin] same dtype; 3x3, stride 2,
 *                   pad 1.  (Metadata passes through unchanged: the caller aliases it.)
 * TDET_OP_CONV      x: [n][h][w][cin] of x_dtype; wgt: [cout][kh][kw][cin] of the SAME dtype;
 *                   y: [n][ho][wo][cout] of y_dtype
 *                   y = act( conv(x) * scale + shift + residual + up2(coarse) )   (true values)
 *                   scale == NULL means 1; shift == NULL means 0 (shift is the conv bias for FPN);
 *                   residual: [n][ho][wo][cout] of residual_dtype or NULL; coarse: [n][hc][wc][cout]
 *                   of coarse_dtype or NULL, with ho == 2*hc and wo == 2*wc (nearest x2,
 *                   fpn.py:100-101).  cin and cout must be multiples of 64.
 *                   *_meta: optional tdet_tensor_meta of each tensor (NULL = exponent 0, unknown max).
 *                   mask (optional): 16-bit [n][ho][wo][cout] forward activation (post-ReLU, so
 *                   >= 0); y is zeroed wherever mask == 0.  This is how a dgrad conv applies the ReLU
 *                   backward of the layer it feeds (resnet.py:48,58,103,108,118).
 *                   A data-gradient (dgrad) of a stride-1 conv IS a TDET_OP_CONV over the output
 *                   gradient with weights from tdet_pack_dgrad_weight and pad' = dil*(k-1) - pad.
 * TDET_OP_SUBSAMPLE y[n][i][j][:] = x[n][2i][2j][:]   (ho = (h-1)/2+1), any 16-bit dtype
 * TDET_OP_WGRAD     x: the conv's forward input [n][h][w][cin] (x_dtype, x_meta); gy: the gradient
 *                   w.r.t. the BN output of that conv (ReLU mask already applied)
 *                   [n][ho][wo][cout] of gy_dtype, which must equal x_dtype (tcgen05 kind::f16 rejects
 *                   mixed fp16/bf16 operands: measured, illegal instruction);
 *                   dw: fp32 [cout][kh][kw][cin], ACCUMULATED (caller zeroes it):
 *                   dw[co][r][s][ci] += scale[co] * sum_{n,p,q} gy[n][p][q][co] * x[n][p*stride-pad+r*dil][..][ci]
 *                   (scale = folded BN scale or NULL = 1).  geometry fields as for the forward conv.
 * TDET_OP_DW_UNPACK x: fp32 [cout][kh][kw][cin]; y: fp32 [cout][cin / groups][kh][kw] (groups > 1: x is the DENSE weight
 *                   gradient of a grouped conv, y keeps each output channel's own group: resnext.py:84-87)
 * TDET_OP_COLSUM    x: 16-bit [n*h*w][cin]; dw: fp32 [cin] += column sums (caller zeroes it)
 * TDET_OP_SUMPOOL2  y[n][i][j][:] = sum_{a,b<2} x[n][2i+a][2j+b][:]   (h == 2*ho, w == 2*wo), bf16,
 *                   fp32 accumulation
 * TDET_OP_DILATE2   y[n][u][v][:] = (u,v even) ? x[n][u/2][v/2][:] : 0; y is (ho, wo) with
 *                   h == (ho+1)/2, w == (wo+1)/2
 * TDET_OP_ADD_MASK  y = (x + residual) * (mask != 0); residual and mask optional; all [n][h][w][cin] 16-bit of
 *                   x_dtype / residual_dtype / y_dtype with optional x_meta / residual_meta exponents.  With
 *                   TDET_FLAG_SCALED_OUT (y_dtype F16, needs y_meta and the inputs' metas with valid amax) the
 *                   output exponent is chosen from amax(x) + amax(residual) and that bound is recorded as
 *                   y's amax.  Without inputs to add or mask it is the format conversion fp16*2^e -> bf16.
 * TDET_OP_BN_AFFINE_GRAD  gy: masked gradient w.r.t. the BN output y [n][h][w][cin] (gy_dtype, gy_meta); x: the
 *                   stored activation that equals y wherever gy != 0 (x_dtype, x_meta); residual (optional):
 *                   tensor to subtract (the block's residual operand, y = x - residual); scale = gamma,
 *                   shift = beta (fp32 [cin]); dw: fp32 [2*cin] accumulator {dgamma[cin], dbeta[cin]}
 *                   (caller zeroes it).  norm_layer / nn.BatchNorm2d in eval mode with trainable affine
 *                   parameters (bn_frozen=False, resnet.py:272-281).
 * TDET_OP_MAXPOOL_BWD  x: the max-pool's input S [n][h][w][cin] (post-ReLU stem output, x_dtype); gy: gradient
 *                   w.r.t. the pooled tensor [n][ho][wo][cin] (gy_dtype, gy_meta); y: bf16 [n][h][w][cin] =
 *                   (S > 0) * scatter of gy to each window's first maximum (aten max_pool2d tie rule)
 *                   (resnet.py:257-258 backward).
 * TDET_OP_PARITY_MERGE  input gradient of a 3x3 / stride 2 / pad 1 conv (resnet.py:75,99) from four stride-1 convs
 *                   over the coarse gradient g [n][hc][wc][.] with the parity classes of the rotated kernel:
 *                   x = p00 [n][hc][wc][cin] (1x1, tap (1,1)), residual = p01 [n][hc+2][wc+1][cin] (1x2, pad 1, taps
 *                   (1,0),(1,2)), coarse = p10 [n][hc+1][wc+2][cin] (2x1, pad 1), gy = p11 [n][hc+1][wc+1][cin] (2x2,
 *                   pad 1), all of x_dtype with their metas; y [n][h][w][cin]: y[2i+a][2j+b] = p_ab[i+o][j+o] (o = 1
 *                   for the padded classes), zeroed where mask (nullable, [n][h][w][cin]) is not positive.
 * TDET_OP_STEM_WGRAD  x: the TDET_OP_PREP staging of the batch (bf16 NHWC4); gy: bf16 [n][ho][wo][64] gradient
 *                   w.r.t. bn1's output; scale: folded bn1 scale (or NULL); dw: fp32 [64][3][7][7], accumulated.
 *                   h, w = image size as for TDET_OP_STEM.
 * TDET_OP_BOTTLENECK_TAIL  the tail of an identity-shortcut Bottleneck with planes = 64 (layer1; resnet.py:105-118) and,
 *                   optionally, the head of the next one (resnet.py:101-103), in ONE kernel:
 *                     z2 = relu(conv3x3(x) * scale + shift)                 x = this block's conv1 output [n][h][w][64],
 *                                                                           wgt [64][3][3][64], 3x3 / stride 1 / pad 1
 *                     y  = relu(conv1x1(z2) * scale2 + shift2 + residual)   wgt2 [256][64]; residual, y [n][h][w][256]
 *                     y2 = relu(conv1x1(y) * scale3 + shift3)               wgt3 [64][256] or NULL; y2 [n][h][w][64]
 *                   z2 never reaches memory and y is not re-read.  x, wgt, wgt2 share x_dtype; wgt3 has y_dtype.
 *                   cin == cout2 == 64 (cout2 = channels of z2), cout == 256, cout3 == 64 (or 0 without wgt3).
 *                   TDET_FLAG_SCALED_OUT: y (F16) gets a device-chosen exponent; TDET_FLAG_SCALED_OUT2: y2 likewise.  The
 *                   exponents come from the chained bounds of bound_consts (conv2), bound_consts2 (conv3) and
 *                   bound_consts3 (next conv1); z2 is scaled like x (an F16 x with x_meta).
 * TDET_OP_GN_STATS  nn.GroupNorm(groups, cin) (models/utils/layers.py:50-54,138-154: use_gn=True backbones, necks with
 *                   normalize + use_gn), first pass.  x: raw conv output [n][h][w][cin] (x_dtype, x_meta exponent);
 *                   dw: fp32 [n][TDET_GN_STAT_BLOCKS][groups][2], partial {sum, sum of squares} of the true values,
 *                   one row per thread block (no atomics: results are bit-reproducible; rows beyond the launched
 *                   blocks are not touched).  cin in {64, 128, ..., 2048} (a power of two), groups <= 64 dividing it.
 * TDET_OP_GN_APPLY  second pass: y = act( (x - mean) * rstd * scale[c] + shift[c] + residual + up2(coarse) ) with
 *                   mean / rstd = 1/sqrt(var + eps) of x's (image, group) from dw (the TDET_OP_GN_STATS output of the
 *                   same x: the rows are added in order; biased variance); scale = gamma, shift = beta (fp32 [cin]); residual [n][h][w][cin] and coarse
 *                   [n][h/2][w/2][cin] optional (16-bit, with metas); TDET_FLAG_RELU; y: 16-bit of y_dtype with
 *                   exponent 0 (F16 saturates at +-65504); y_meta (optional) receives max |y|.
 * TDET_OP_AMAX      x: 16-bit [n][h][w][cin] (x_dtype, x_meta exponent); y_meta: receives max |x| (true values;
 *                   its exponent field is left untouched)
 * TDET_OP_ZERO      y: buffer of x_stride[0] bytes, zero-filled
 */
typedef struct tdet_op {
  int32_t kind;  /* tdet_op_kind */
  int32_t flags; /* TDET_FLAG_* */
  int32_t n, h, w, cin;
  int32_t cout, kh, kw;
  int32_t stride, pad, dil;
  int32_t ho, wo;
  int32_t hc, wc;     /* coarse level size for the upsample-add epilogue */
  int32_t x_dtype;    /* tdet_dtype of x (and of wgt for conv ops) */
  int32_t y_dtype;    /* tdet_dtype of y (BF16 or F16) */
  int32_t residual_dtype;
  int32_t coarse_dtype;
  int64_t x_stride[4]; /* PREP only: element strides of x for (n, c, h, w) */
  const void* x;
  const void* wgt;
  void* y;
  const float* scale;
  const float* shift;
  const void* residual;
  const void* coarse;
  const tdet_tensor_meta* x_meta;
  const tdet_tensor_meta* residual_meta;
  const tdet_tensor_meta* coarse_meta;
  tdet_tensor_meta* y_meta;
  const float* bound_consts; /* tdet_conv_bound_consts output; needed with TDET_FLAG_SCALED_OUT */
  /* ---- training path ---- */
  const void* mask;          /* CONV / ADD_MASK: forward activation to mask the result against */
  const void* gy;            /* WGRAD: output gradient */
  int32_t gy_dtype;
  int32_t groups;            /* CONV: 0/1 = dense; > 1 = grouped conv (cin == cout, 64 % (cin/groups) == 0) over
                                weights from tdet_pack_grouped_conv_weight (ResNeXt, resnext.py:84-87) */
  float* dw;                 /* WGRAD / COLSUM: fp32 accumulator */
  const tdet_tensor_meta* gy_meta; /* WGRAD: exponent of gy (NULL = 0) */
  /* ---- second input (TDET_FLAG_DUAL) ---- */
  const void* x2;            /* [n][h2][w2][cin2] of x2_dtype == x_dtype; ho == (h2-1)/stride2+1, wo likewise */
  const tdet_tensor_meta* x2_meta;
  int32_t cin2, stride2, h2, w2;
  int32_t x2_dtype;
  float eps;                  /* TDET_OP_GN_APPLY: GroupNorm eps */
  /* ---- TDET_OP_BOTTLENECK_TAIL: conv3 and the next block's conv1 ---- */
  const void* wgt2;           /* conv3 weights [cout][cout2] of x_dtype */
  const float* scale2;
  const float* shift2;
  const float* bound_consts2;
  const void* wgt3;           /* next conv1 weights [cout3][cout] of y_dtype, or NULL */
  const float* scale3;
  const float* shift3;
  const float* bound_consts3;
  void* y2;                   /* [n][h][w][cout3] of y2_dtype */
  tdet_tensor_meta* y2_meta;
  int32_t cout2, cout3;
  int32_t y2_dtype;
  int32_t reserved1;
} tdet_op;

typedef struct tdet_plan tdet_plan; /* opaque */

/* ---- library / device ------------------------------------------------------------------- */
int tdet_abi_version(void);
const char* tdet_last_error(void);
/* 0 if `device` is an sm_100 part this build can run on, else TDET_ERR_UNSUPPORTED_DEVICE. */
int tdet_device_supported(int device);

/* The GEMM kernels are persistent, one CTA per SM.  While another stream needs SMs of its own -- NCCL
 * all-reducing gradient buckets under the backward pass -- a full-width grid would wait for them and
 * run a second wave; plans built after this call size their grids to (SM count - sms). */
int tdet_set_sm_reserve(int device, int sms);

/* Geometry of the TDET_OP_PREP output for a stem output of ho x wo pixels: hp = 2*ho + 6 rows,
 * wp = 2*wo + 16 rounded up to a multiple of 16 pixels (the image sits at offset (3,3)). */
int tdet_stem_staging_dims(int ho, int wo, int* hp, int* wp);

/* ---- operand preparation (run once per weight version) ----------------------------------- */
/* fp32 OIHW [cout][cin][kh][kw] -> 16-bit [cout][kh][kw][cin] (round-to-nearest-even);
 * dtype = TDET_BF16 or TDET_F16. */
int tdet_pack_conv_weight(const float* w_oihw, void* w_packed, int cout, int cin, int kh, int kw,
                          int dtype, void* stream);
/* Same with a per-output-channel factor (a folded BatchNorm scale) multiplied in before rounding, written into a
 * row-pitched destination: w_packed[co * ld + (r*kw + s)*cin + ci] = round16(scale[co] * w[co][ci][r][s]).
 * ld >= kh*kw*cin lets two matrices be packed side by side ([W | W2], TDET_FLAG_DUAL). */
int tdet_pack_conv_weight_scaled(const float* w_oihw, const float* scale, void* w_packed, int cout, int cin, int kh,
                                 int kw, int ld, int dtype, void* stream);
/* fp32 [64][3][7][7] -> 16-bit [64][448]: k = r*64 + s*4 + c (s < 7, c < 3), zero padded.  dtype = TDET_BF16, or
 * TDET_F16 for a stem over an fp16 staging (TDET_OP_PREP with y_dtype = TDET_F16). */
int tdet_pack_stem_weight(const float* w_oihw, void* w_packed, int32_t dtype, void* stream);
/* eval-mode BatchNorm2d -> per-channel fp32 scale = gamma/sqrt(var+eps), shift = beta-mean*scale
 * (models/utils/layers.py:50-54 builds nn.BatchNorm2d; eps is its default 1e-5). */
int tdet_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var,
                 float eps, float* scale, float* shift, int channels, void* stream);
/* Split-precision operands: fp32 OIHW -> bf16 [cout][kh][kw][2*cin], per filter tap the cin hi values
 * bf16(w) followed by the cin lo values bf16(w - hi).  Stem: fp32 [64][3][7][7] -> bf16 [64][896] (the 448-column
 * layout of tdet_pack_stem_weight for hi, then for lo). */
int tdet_pack_conv_weight_split(const float* w_oihw, void* w_packed, int cout, int cin, int kh, int kw, void* stream);
int tdet_pack_stem_weight_split(const float* w_oihw, void* w_packed, void* stream);
/* Grouped conv weights: fp32 [cout][cin/groups][kh][kw] -> DENSE 16-bit [cout][kh][kw][cin], zero outside each
 * output channel's group (block-diagonal): the GEMM kernel then contracts, per 64-wide tile of output
 * channels, only the 64 input channels of the same range. */
int tdet_pack_grouped_conv_weight(const float* w, void* w_packed, int cout, int cin, int kh, int kw, int groups,
                                  int dtype, void* stream);
/* Operand of the data-gradient conv: out[ci][kh-1-r][kw-1-s][co] = scale[co] * w[co][ci][r][s]
 * (scale = folded BN scale of the forward conv, NULL = 1), 16-bit, round-to-nearest-even. */
int tdet_pack_dgrad_weight(const float* w_oihw, const float* scale, void* w_packed, int cout, int cin,
                           int kh, int kw, int dtype, void* stream);
/* consts[0] = max_c |scale_c| * sum_k |w_packed[c][k]|, consts[1] = max_c |shift_c|  (scale NULL = 1,
 * shift NULL = 0): |conv(x)*scale + shift| <= consts[0] * max|x| + consts[1] for any x. */
int tdet_conv_bound_consts(const void* w_packed, int dtype, const float* scale, const float* shift,
                           int cout, int k, float* consts, void* stream);

/* ---- execution ---------------------------------------------------------------------------- */
/* Validates and runs one op immediately (descriptors are built on the fly). */
int tdet_op_run(const tdet_op* op, int device, void* stream);

/*
 * A plan is a validated op sequence with pre-built TMA descriptors and launch configurations.
 * ext_ptrs/ext_bytes list the n_ext caller tensors (base pointer, size in bytes) that may move
 * between runs (network input, returned outputs); every op field pointing into
 * [ext_ptrs[i], ext_ptrs[i] + ext_bytes[i]) is re-bound, keeping its byte offset, to the i-th pointer
 * given to tdet_plan_run.  Pass n_ext = 0 for a fully static plan.
 * meta_arena/meta_count (optional): the contiguous tdet_tensor_meta array the ops' *_meta pointers
 * live in; it is zeroed (one cudaMemsetAsync) at the start of every run.
 */
int tdet_plan_create(tdet_plan** out, const tdet_op* ops, int n_ops, const void* const* ext_ptrs,
                     const size_t* ext_bytes, int n_ext, tdet_tensor_meta* meta_arena,
                     int meta_count, int device);
int tdet_plan_run(tdet_plan* plan, const void* const* ext_ptrs, int n_ext, void* stream);
/* Runs ops [first, last) only (no metadata reset unless first == 0): lets the caller interleave its
 * own stream work -- the bucketed gradient all-reduce of the training path -- between segments. */
int tdet_plan_run_range(tdet_plan* plan, const void* const* ext_ptrs, int n_ext, int first, int last,
                        void* stream);
/* Same as tdet_plan_run with unchanged external pointers, but brackets every kernel launch with
 * CUDA events on `stream` and returns the per-launch device time in milliseconds
 * (ms_per_launch[tdet_plan_num_launches]).  Synchronises the stream; measurement only. */
int tdet_plan_run_timed(tdet_plan* plan, const void* const* ext_ptrs, int n_ext, void* stream,
                        float* ms_per_launch);
int tdet_plan_num_launches(const tdet_plan* plan);
/* Per-launch accounting for roofline reports: GEMM view, algorithmic FLOPs (2*M*N*K, real dims)
 * and algorithmic HBM bytes (each operand read once, the output written once). */
typedef struct tdet_launch_info {
  int32_t kind;   /* tdet_op_kind */
  int32_t tile_n; /* GEMM tile width (0 for non-GEMM kernels) */
  int32_t grid;   /* CTAs launched (GEMM kernels) */
  int32_t a_mode; /* 0 tiled, 1 im2col, 2 stem; -1 for non-GEMM kernels */
  int32_t m, n, k;
  int32_t variant; /* GEMM kernel instantiation: 4096 * patch-mode + stages * 256 + residual slabs * 16 + resident-B k-blocks */
  double flops;
  double bytes;
} tdet_launch_info;
int tdet_plan_launch_info(const tdet_plan* plan, int index, tdet_launch_info* out);
/* Sum over TDET_OP_CONV/STEM ops of 2*M*N*K (un-padded dims), for roofline accounting. */
double tdet_plan_flops(const tdet_plan* plan);
int tdet_plan_destroy(tdet_plan* plan);

/* Test hook: runs the TMA im2col loader alone and dumps one 128x64 A tile (un-swizzled, 16-bit
 * [128][64]) for output rows [m0, m0+128), filter tap (r, s), channel chunk kc.  Used by the
 * GPU tests to pin the descriptor semantics independently of the MMA. */
int tdet_debug_im2col_tile(const tdet_op* op, int m0, int r, int s, int kc, void* tile_out,
                           int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TDET_B200_H_ */
