/*
 * lce_b200_types.h -- plain-C descriptors shared by the C-ABI CUDA layer
 * (include/lce_b200.h), the CPU oracle (oracle/lce_oracle.c) and the compiled
 * reference shim (oracle/ref_shim.cc). No torch / TFLite types here.
 *
 * Field meanings follow the reference's BConv2DParams
 * (larq_compute_engine/core/bconv2d/params.h:12-32) and BMaxPoolParams
 * (larq_compute_engine/core/bmaxpool.h:14-21); enum integers follow the TFLite
 * schema the converter writes into the op's flexbuffer
 * (larq_compute_engine/mlir/ir/lce_ops.cc:36-64).
 */
#ifndef LCE_B200_TYPES_H_
#define LCE_B200_TYPES_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TFLite schema `Padding` (schema.fbs:792) */
enum { LCE_PADDING_SAME = 0, LCE_PADDING_VALID = 1 };
/* TFLite schema `ActivationFunctionType` (schema.fbs:796) */
enum {
  LCE_ACT_NONE = 0,
  LCE_ACT_RELU = 1,
  LCE_ACT_RELU_N1_TO_1 = 2,
  LCE_ACT_RELU6 = 3
};
/* Output element type of LceBconv2d: dispatch of bconv2d.cc:551-564 */
enum { LCE_OUT_FLOAT = 0, LCE_OUT_INT8 = 1, LCE_OUT_BITPACKED = 2 };
/* SAME padding with pad_values == 0 ("zero padding"), the two results the reference has:
 *  REFERENCE  : the reference kernel's integers (LCE/core/bconv2d/reference.h:76-77,100-103: an
 *               out-of-bounds tap adds channels_in_per_group / 2), any output type, needs an
 *               even channels_in (bconv2d.cc:188-200) -- Register_BCONV_2D_REF.
 *  CORRECTION : the optimised kernels' result -- one-padding, OutputTransform, then a FLOAT
 *               correction on the edge outputs (optimized_bgemm.h:153-177,
 *               zero_padding_correction.h:39-299); float output without fused activation only --
 *               Register_BCONV_2D (the reference's default) and ..._OPT_INDIRECT_BGEMM.
 * The two differ by float rounding (up to ~1e5 ULP near cancellation). */
enum { LCE_ZERO_PADDING_REFERENCE = 0, LCE_ZERO_PADDING_CORRECTION = 1 };
/* Input element type of LceQuantize / output of LceDequantize
 * (quantization.cc:76-147) */
enum { LCE_T_FLOAT = 0, LCE_T_INT8 = 1, LCE_T_BOOL = 2 };

/* One LceBconv2d invocation. Activations are NHWC with the channel axis
 * bitpacked into int32 words (bit i of word w = channel 32w+i, bit 1 <=> value
 * < 0; types.h:41-47, bitpack.h:159); filters are OHWI-packed
 * [channels_out, filter_h, filter_w, ceil(channels_in/groups/32)]. */
typedef struct lce_bconv2d_desc {
  int32_t batch, in_h, in_w;
  int32_t channels_in;  /* unpacked, whole tensor (attribute `channels_in`) */
  int32_t filter_h, filter_w;
  int32_t channels_out;
  int32_t groups;
  int32_t stride_h, stride_w;
  int32_t dilation_h, dilation_w;
  int32_t padding;    /* LCE_PADDING_* */
  int32_t pad_value;  /* 0 or 1 (attribute `pad_values`) */
  int32_t activation; /* LCE_ACT_* */
  int32_t out_type;   /* LCE_OUT_* */
  float out_scale;    /* int8 output only (TfLiteTensor::params.scale) */
  int32_t out_zero_point;
} lce_bconv2d_desc;

typedef struct lce_bmaxpool_desc {
  int32_t batch, in_h, in_w;
  int32_t channels_packed; /* int32 words per pixel */
  int32_t filter_h, filter_w;
  int32_t stride_h, stride_w;
  int32_t padding; /* LCE_PADDING_* */
} lce_bmaxpool_desc;

/* Epilogue of a plain BGEMM (config 5 sweep): what OutputTransform<Dst>
 * (output_transform.h:94-168) needs, already folded. */
typedef struct lce_bgemm_epilogue {
  int32_t out_type;  /* LCE_OUT_*; 3 = raw int32 accumulators */
  int32_t clamp_min, clamp_max;
  const float* multiplier; /* [N], folded (mul') */
  const float* bias;       /* [N], folded (bias') */
  const int32_t* thresholds; /* [N] */
} lce_bgemm_epilogue;
#define LCE_OUT_RAW_ACC 3

#ifdef __cplusplus
}
#endif
#endif /* LCE_B200_TYPES_H_ */
