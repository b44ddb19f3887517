/*
 * lce_b200.h -- the drop-in boundary: a C-ABI CUDA layer (sm_100a) for the one
 * hot path of larq/compute-engine: LceQuantize -> LceBconv2d (-> LceBMaxPool2d,
 * LceDequantize). Plain pointers and sizes only; no torch / TFLite types.
 *
 * Every entry point cites the reference interface it replaces (LCE =
 * /root/reference/larq_compute_engine). The TFLite custom-op shell that sits on
 * top of these calls (TfLiteRegistration {init, free, prepare, invoke}) is
 * declared in include/lce_b200_tflite.h; the binding a reference maintainer
 * would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - `*_dev` pointers are CUDA device pointers on the current device; `stream`
 *    is a cudaStream_t passed as void* (NULL = default stream). Calls are
 *    asynchronous on that stream unless the name ends in `_host`.
 *  - Return value: 0 = ok (kTfLiteOk), non-zero = error (kTfLiteError); the
 *    message is available from lce_b200_last_error() (thread local), mirroring
 *    context->ReportError in the reference (bconv2d.cc:76-83).
 *  - There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef LCE_B200_H_
#define LCE_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "lce_b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LCE_B200_ABI_VERSION 1

int lce_b200_abi_version(void);
const char* lce_b200_last_error(void);
/* Number of visible CUDA devices (0 if none / driver missing). */
int lce_b200_device_count(void);

/* ---- LceQuantize --------------------------------------------------------
 * Replaces QuantizeEval (LCE/tflite/kernels/quantization.cc:76-114) ->
 * bitpack_tensor (LCE/core/bitpacking/utils.h:24-33) -> bitpack_matrix
 * (LCE/core/bitpacking/bitpack.h:249-308). in: [rows, cols] of in_type
 * (LCE_T_*); out: [rows, ceil(cols/32)] int32. */
int lce_b200_quantize(int in_type, const void* in_dev, int64_t rows,
                      int64_t cols, int32_t zero_point, int32_t* out_dev,
                      void* stream);

/* ---- LceDequantize ------------------------------------------------------
 * Replaces DequantizeEval (quantization.cc:116-147) -> unpack_matrix
 * (bitpack.h:325-346). `cols` is the unpacked channel count. */
int lce_b200_dequantize(int out_type, const int32_t* in_dev, int64_t rows,
                        int64_t cols, float scale, int32_t zero_point,
                        void* out_dev, void* stream);

/* ---- LceBMaxPool2d ------------------------------------------------------
 * Replaces bmaxpool::Prepare/Eval (LCE/tflite/kernels/bmaxpool.cc:40-88) ->
 * BMaxPool (LCE/core/bmaxpool.h:24-88). */
int lce_b200_bmaxpool_out_shape(const lce_bmaxpool_desc* d, int* out_h,
                                int* out_w);
int lce_b200_bmaxpool(const lce_bmaxpool_desc* d, const int32_t* in_dev,
                      int32_t* out_dev, void* stream);

/* ---- LceBconv2d ---------------------------------------------------------
 * A plan object carries what bconv2d::OpData carries (bconv2d.cc:44-74): the
 * parameters, the folded output transform, and device copies of the constant
 * inputs, re-laid-out once for the kernel ("LHS cached", optimized_bgemm.h:135).
 *
 * out_shape  : shape inference of Prepare (bconv2d.cc:169-248).
 * create     : Init + Prepare + OneTimeSetup (bconv2d.cc:85-131,138-300,324-392).
 *              `filter` is OHWI-packed [channels_out, filter_h, filter_w,
 *              ceil(channels_in/groups/32)]; `post_mul`/`post_bias` are the op's
 *              inputs 2/3 (float/int8 output), `thresholds` input 4 (bitpacked
 *              output); absent optional inputs are NULL (bconv2d.cc:145-152).
 *              These four may be host or device pointers. desc->batch/in_h/in_w
 *              give the initial input shape. Refuses what the reference refuses
 *              (bconv2d.cc:113-116,169-200), with the reference-kernel rule for
 *              zero padding (even channels_in).
 * set_input_shape : what a second Prepare after ResizeInputTensor does
 *              (bconv2d.cc:295-297).
 * run        : Eval (bconv2d.cc:551-564) on bitpacked NHWC input
 *              [batch, in_h, in_w, ceil(channels_in/32)] -> output NHWC float /
 *              int8 [.., channels_out] or int32 [.., ceil(channels_out/32)].
 * run_f32    : the same with LceQuantize fused into the prologue: input is the
 *              float NHWC tensor [batch, in_h, in_w, channels_in].
 * run_host   : run with HOST buffers (H2D, kernel, D2H, synchronised) -- the
 *              call a stock TFLite interpreter with a host arena makes. */
typedef struct lce_b200_bconv2d lce_b200_bconv2d;

int lce_b200_bconv2d_out_shape(const lce_bconv2d_desc* d, int* out_h,
                               int* out_w, int* pad_h, int* pad_w);
int lce_b200_bconv2d_create(const lce_bconv2d_desc* d, const int32_t* filter,
                            const float* post_mul, const float* post_bias,
                            const int32_t* thresholds, lce_b200_bconv2d** plan);
int lce_b200_bconv2d_set_input_shape(lce_b200_bconv2d* plan, int batch,
                                     int in_h, int in_w);
/* Which of the reference's two zero-padding results a plan reproduces (LCE_ZERO_PADDING_*,
 * lce_b200_types.h). create() picks REFERENCE when channels_in is even (what
 * Register_BCONV_2D_REF computes) and CORRECTION otherwise; the op shell sets it from the
 * registration it was built through. No effect unless padding == SAME and pad_value == 0. */
int lce_b200_bconv2d_set_zero_padding_mode(lce_b200_bconv2d* plan, int mode);
int lce_b200_bconv2d_get_desc(const lce_b200_bconv2d* plan,
                              lce_bconv2d_desc* d, int* out_h, int* out_w);
int lce_b200_bconv2d_run(lce_b200_bconv2d* plan, const int32_t* in_dev,
                         void* out_dev, void* stream);
int lce_b200_bconv2d_run_f32(lce_b200_bconv2d* plan, const float* in_dev,
                             void* out_dev, void* stream);
/* Fused residual-block tail (graph-level fusion of the converter pattern
 *   LceBconv2d(float out) -> ADD(shortcut) [-> LceQuantize of the sum]):
 *   out = act_add(OutputTransform(acc) + residual);  packed_out = bitpack(out) (optional).
 * Bit-identical to running the three ops one after the other. Float output plans only;
 * packed_out additionally needs groups == 1. residual / packed_out may be NULL. */
int lce_b200_bconv2d_run_fused(lce_b200_bconv2d* plan, const int32_t* in_dev,
                               const float* residual_dev, int add_activation,
                               float* out_dev, int32_t* packed_out_dev, void* stream);
int lce_b200_bconv2d_run_host(lce_b200_bconv2d* plan, const int32_t* in_host,
                              void* out_host);
void lce_b200_bconv2d_destroy(lce_b200_bconv2d* plan);

/* ---- BGEMM --------------------------------------------------------------
 * Replaces bgemm::BGemm (LCE/core/bgemm/bgemm.h:25-84) with the orientation of
 * optimized_bgemm.h:126-151: A = activations [M, Kw] (row per output pixel),
 * W = filters [N, Kw]; out [M, N] row-major (raw int32 accumulators, float or
 * int8) or [M, ceil(N/32)] bitpacked. The epilogue is OutputTransform<Dst>
 * (LCE/core/bconv2d/output_transform.h:94-168), already folded. */
typedef struct lce_b200_bgemm lce_b200_bgemm;

int lce_b200_bgemm_create(int N, int Kw, const int32_t* W,
                          const lce_bgemm_epilogue* ep, lce_b200_bgemm** plan);
int lce_b200_bgemm_run(lce_b200_bgemm* plan, int64_t M, const int32_t* A_dev,
                       void* out_dev, void* stream);
void lce_b200_bgemm_destroy(lce_b200_bgemm* plan);

/* Number of kernels this library has launched in this process (bench.py's
 * `gpu_launches`). */
uint64_t lce_b200_launch_count(void);
/* Diagnostics. path_counts: inner-product launches so far by kernel family
 * {tcgen05 (lce_b200_tc.cuh), mma.sync int8 (lce_b200_imma.cuh), XOR + POPC
 * (lce_b200_kernels.cuh)} -- the tests use it to prove which kernel ran.
 * tc_debug: the tcgen05 kernel's deadlock-watchdog record; out[7] != 0 means a
 * wait timed out ({barrier tag, block, thread, parity, count} in out[0..4]);
 * reading clears the flag. */
void lce_b200_path_counts(uint64_t out[3]);
int lce_b200_tc_debug(int32_t out[8]);

#ifdef __cplusplus
}
#endif
#endif /* LCE_B200_H_ */
