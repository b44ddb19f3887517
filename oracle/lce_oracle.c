/*
 * oracle/lce_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the reference's binary-convolution hot path,
 * used as the checker for the CUDA path (tests/, __graft_entry__.smoke()) and
 * as the "port" CPU baseline in bench.py. It is NOT part of the product: nothing
 * under compute_engine_b200/ links, imports or calls it.
 *
 * Parity pinning: this file is checked (tests/test_oracle.py) against
 *   (1) the known-answer vectors the reference's own tests hold,
 *   (2) golden vectors under tests/golden/ minted with fixed seeds from the
 *       reference's own headers compiled here (oracle/ref_shim.cc ->
 *       oracle/_ref/liblce_ref.so; generator: tests/golden/make_golden.py),
 *   (3) live against oracle/_ref/liblce_ref.so when that library is present.
 *
 * Each function cites the reference file:line it follows. LCE = larq_compute_engine.
 */
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lce_b200_types.h"

/* LCE/core/types.h:45-47 */
static inline int xor_popcount(int32_t a, int32_t b) {
  return __builtin_popcount((uint32_t)(a ^ b));
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

/* tensorflow/lite/kernels/padding.h:42-60 (ComputeOutSize) */
static int out_size(int padding, int image, int filter, int stride, int dil) {
  const int eff = (filter - 1) * dil + 1;
  if (stride == 0) return 0;
  if (padding == LCE_PADDING_SAME) return (image + stride - 1) / stride;
  if (padding == LCE_PADDING_VALID) return (image + stride - eff) / stride;
  return 0;
}

/* tensorflow/lite/kernels/padding.h:32-40 (ComputePaddingWithOffset; the
 * odd pixel goes after). */
static int pad_before(int stride, int dil, int in, int filter, int out) {
  const int eff = (filter - 1) * dil + 1;
  int total = (out - 1) * stride + eff - in;
  if (total < 0) total = 0;
  return total / 2;
}

/* Shape inference of Prepare: LCE/tflite/kernels/bconv2d.cc:169-248. */
int lce_oracle_bconv2d_out_shape(const lce_bconv2d_desc* d, int* out_h,
                                 int* out_w, int* pad_h, int* pad_w) {
  if (d->groups < 1 || d->channels_in % d->groups ||
      d->channels_out % d->groups)
    return 1;
  if (d->groups > 1 && (d->channels_in / d->groups) % 32) return 1;
  if (d->pad_value != 0 && d->pad_value != 1) return 1; /* bconv2d.cc:113-116 */
  *out_h = out_size(d->padding, d->in_h, d->filter_h, d->stride_h, d->dilation_h);
  *out_w = out_size(d->padding, d->in_w, d->filter_w, d->stride_w, d->dilation_w);
  *pad_h = pad_before(d->stride_h, d->dilation_h, d->in_h, d->filter_h, *out_h);
  *pad_w = pad_before(d->stride_w, d->dilation_w, d->in_w, d->filter_w, *out_w);
  return 0;
}

/* OneTimeSetup: LCE/tflite/kernels/bconv2d.cc:353-389. Multiplier and bias are
 * folded in double and rounded to float once; clamp bounds follow
 * CalculateActivationRange<int32> (tensorflow/lite/kernels/kernel_util.h:285-300). */
void lce_oracle_fold_output_transform(const lce_bconv2d_desc* d,
                                      const float* post_mul,
                                      const float* post_bias, float* mul_out,
                                      float* bias_out, int32_t* clamp_min,
                                      int32_t* clamp_max) {
  const int32_t backtransform_add =
      d->filter_h * d->filter_w * (d->channels_in / d->groups);
  const double scale = d->out_type == LCE_OUT_INT8 ? (double)d->out_scale : 1.0;
  const double zp = d->out_type == LCE_OUT_INT8 ? (double)d->out_zero_point : 0.0;
  for (int i = 0; i < d->channels_out; ++i) {
    const double m = post_mul[i];
    const double b = post_bias[i];
    mul_out[i] = (float)(-1 * m / scale);
    bias_out[i] = (float)((b + (double)backtransform_add * m) / scale + zp);
  }
  int32_t nmin, nmax;
  switch (d->activation) {
    case LCE_ACT_RELU: nmin = 0; nmax = INT32_MAX; break;
    case LCE_ACT_RELU6: nmin = 0; nmax = 6; break;
    case LCE_ACT_RELU_N1_TO_1: nmin = -1; nmax = 1; break;
    default: nmin = INT32_MIN; nmax = INT32_MAX; break;
  }
  if (nmin < -backtransform_add) nmin = -backtransform_add;
  if (nmax > backtransform_add) nmax = backtransform_add;
  *clamp_min = -nmax + backtransform_add;
  *clamp_max = -nmin + backtransform_add;
}

/* OutputTransform<float>::Run, LCE/core/bconv2d/output_transform.h:100-106.
 * Two roundings (mul then add): compile with -ffp-contract=off. */
static inline float transform_float(int32_t accum, int32_t cmin, int32_t cmax,
                                    float mul, float bias) {
  int32_t x = (int32_t)((uint32_t)accum << 1);
  x = x < cmax ? x : cmax;
  x = x > cmin ? x : cmin;
  volatile float prod = (float)x * mul; /* forbid contraction */
  return prod + bias;
}

/* core::round + saturate, LCE/core/types.h:50-59,81-94 (x86: TfLiteRound =
 * std::round, ties away from zero; the float->int32 conversion is the x86
 * cvttss2si, which yields INT32_MIN for out-of-range / NaN). */
static inline int8_t round_saturate(float y) {
  const float r = roundf(y);
  int32_t q;
  if (r >= -2147483648.0f && r < 2147483648.0f) q = (int32_t)r;
  else q = INT32_MIN;
  if (q > 127) q = 127;
  if (q < -128) q = -128;
  return (int8_t)q;
}

/* BConv2DReference, LCE/core/bconv2d/reference.h:35-148, for batch images
 * [b0, b1). Out-of-bounds taps: pad_value 1 -> input word 0 (:106); SAME +
 * pad_value 0 -> accum += channels_in_per_group/2 (:76-77,100-103). */
static void bconv2d_range(const lce_bconv2d_desc* d, int out_h, int out_w,
                          int pad_h, int pad_w, const int32_t* input,
                          const int32_t* filter, const float* mul,
                          const float* bias, int32_t cmin, int32_t cmax,
                          const int32_t* thresholds, void* output, int b0,
                          int b1) {
  const int cw_in = ceil_div(d->channels_in, 32);
  const int cin_pg = d->channels_in / d->groups;
  const int cw_pg = ceil_div(cin_pg, 32);
  const int cout_pg = d->channels_out / d->groups;
  const int cw_out = ceil_div(d->channels_out, 32);
  const int zero_padding = d->padding == LCE_PADDING_SAME && d->pad_value == 0;
  const int binary_zero_point = cin_pg / 2;
  for (int b = b0; b < b1; ++b)
    for (int oy = 0; oy < out_h; ++oy)
      for (int ox = 0; ox < out_w; ++ox) {
        int32_t column = 0;
        const size_t pix = ((size_t)b * out_h + oy) * out_w + ox;
        for (int oc = 0; oc < d->channels_out; ++oc) {
          const int g = oc / cout_pg;
          int32_t accum = 0;
          for (int fy = 0; fy < d->filter_h; ++fy)
            for (int fx = 0; fx < d->filter_w; ++fx) {
              const int ix = ox * d->stride_w - pad_w + d->dilation_w * fx;
              const int iy = oy * d->stride_h - pad_h + d->dilation_h * fy;
              const int inside =
                  ix >= 0 && ix < d->in_w && iy >= 0 && iy < d->in_h;
              if (zero_padding && !inside) {
                accum += binary_zero_point;
                continue;
              }
              const int32_t* f =
                  filter +
                  (((size_t)oc * d->filter_h + fy) * d->filter_w + fx) * cw_pg;
              if (inside) {
                const int32_t* a =
                    input + (((size_t)b * d->in_h + iy) * d->in_w + ix) * cw_in +
                    (size_t)g * cw_pg;
                for (int w = 0; w < cw_pg; ++w) accum += xor_popcount(a[w], f[w]);
              } else {
                for (int w = 0; w < cw_pg; ++w) accum += xor_popcount(0, f[w]);
              }
            }
          if (d->out_type == LCE_OUT_BITPACKED) {
            /* OutputTransform<TBitpacked>::Run, output_transform.h:164-167 */
            if (accum > thresholds[oc]) column |= (int32_t)(1u << (oc % 32));
            if ((oc + 1) % 32 == 0 || oc + 1 == d->channels_out) {
              ((int32_t*)output)[pix * cw_out + oc / 32] = column;
              column = 0;
            }
          } else if (d->out_type == LCE_OUT_FLOAT) {
            ((float*)output)[pix * d->channels_out + oc] =
                transform_float(accum, cmin, cmax, mul[oc], bias[oc]);
          } else { /* OutputTransform<int8>::Run, output_transform.h:132-143 */
            ((int8_t*)output)[pix * d->channels_out + oc] = round_saturate(
                transform_float(accum, cmin, cmax, mul[oc], bias[oc]));
          }
        }
      }
}

typedef struct {
  const lce_bconv2d_desc* d;
  int out_h, out_w, pad_h, pad_w;
  const int32_t *input, *filter, *thresholds;
  const float *mul, *bias;
  int32_t cmin, cmax;
  void* output;
  int next; /* protected by mu */
  pthread_mutex_t mu;
} conv_job;

static void* conv_worker(void* arg) {
  conv_job* j = (conv_job*)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    const int b = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (b >= j->d->batch) break;
    bconv2d_range(j->d, j->out_h, j->out_w, j->pad_h, j->pad_w, j->input,
                  j->filter, j->mul, j->bias, j->cmin, j->cmax, j->thresholds,
                  j->output, b, b + 1);
  }
  return NULL;
}

/* LceBconv2d with the reference kernel (Register_BCONV_2D_REF): post_mul /
 * post_bias are the op's raw inputs 2/3; thresholds is input 4. `threads`
 * images are processed concurrently (the reference itself is single-threaded;
 * images are independent). Returns 0, or 1 on invalid parameters, 2 when the
 * reference would refuse in Prepare (bconv2d.cc:188-200). */
int lce_oracle_bconv2d_mt(const lce_bconv2d_desc* d, int threads,
                          const int32_t* input, const int32_t* filter,
                          const float* post_mul, const float* post_bias,
                          const int32_t* thresholds, void* output) {
  conv_job j;
  memset(&j, 0, sizeof(j));
  if (lce_oracle_bconv2d_out_shape(d, &j.out_h, &j.out_w, &j.pad_h, &j.pad_w))
    return 1;
  if (d->padding == LCE_PADDING_SAME && d->pad_value == 0 &&
      d->channels_in % 2 != 0)
    return 2;
  float* mul = NULL;
  float* bias = NULL;
  if (d->out_type != LCE_OUT_BITPACKED) {
    if (!post_mul || !post_bias) return 1;
    mul = (float*)malloc(sizeof(float) * d->channels_out);
    bias = (float*)malloc(sizeof(float) * d->channels_out);
    lce_oracle_fold_output_transform(d, post_mul, post_bias, mul, bias, &j.cmin,
                                     &j.cmax);
  } else if (!thresholds) {
    return 1;
  }
  j.d = d; j.input = input; j.filter = filter; j.thresholds = thresholds;
  j.mul = mul; j.bias = bias; j.output = output; j.next = 0;
  pthread_mutex_init(&j.mu, NULL);
  if (threads > d->batch) threads = d->batch;
  if (threads < 1) threads = 1;
  pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  for (int i = 1; i < threads; ++i) pthread_create(&tid[i], NULL, conv_worker, &j);
  conv_worker(&j);
  for (int i = 1; i < threads; ++i) pthread_join(tid[i], NULL);
  free(tid);
  pthread_mutex_destroy(&j.mu);
  free(mul);
  free(bias);
  return 0;
}

int lce_oracle_bconv2d(const lce_bconv2d_desc* d, const int32_t* input,
                       const int32_t* filter, const float* post_mul,
                       const float* post_bias, const int32_t* thresholds,
                       void* output) {
  return lce_oracle_bconv2d_mt(d, 1, input, filter, post_mul, post_bias,
                               thresholds, output);
}

/* --------------------------------------------------------------------------
 * LceBconv2d with the OPTIMISED kernels' semantics (Register_BCONV_2D -- the
 * reference's default registration -- and Register_BCONV_2D_OPT_INDIRECT_BGEMM).
 * They produce the reference kernel's integers except under SAME padding with
 * pad_values 0 ("zero padding"), float output: there the convolution runs with
 * one-padding (out-of-bounds taps read as bit 0 = +1, optimized_bgemm.h:30-31,
 * indirect_bgemm/kernel.h:101-174), the OutputTransform is applied, and a FLOAT
 * correction is added to the edge outputs afterwards
 * (optimized_bgemm.h:153-177 -> zero_padding_correction.h:178-299). Prepare only
 * admits float output without a fused activation there (bconv2d.cc:188-200).
 * -------------------------------------------------------------------------- */

/* CacheCorrectionValues, zero_padding_correction.h:39-176, for one (case, y, x, out_c):
 * -post_mul[c] * sum over the taps the case counts of (channels_in_pg - 2 popc(filter tap)). */
static float zpc_cache_value(const lce_bconv2d_desc* d, const int32_t* filter,
                             const float* post_mul, int direction, int y, int x,
                             int out_c) {
  const int cin_pg = d->channels_in / d->groups;
  const int cw_pg = ceil_div(cin_pg, 32);
  const int eff_w = (d->filter_w - 1) * d->dilation_w + 1;
  const int eff_h = (d->filter_h - 1) * d->dilation_h + 1;
  float correction = 0.0f;
  for (int fy = 0; fy < d->filter_h; ++fy)
    for (int fx = 0; fx < d->filter_w; ++fx) {
      int popcount = 0;
      const int32_t* f =
          filter + (((size_t)out_c * d->filter_h + fy) * d->filter_w + fx) * cw_pg;
      for (int w = 0; w < cw_pg; ++w) popcount += xor_popcount(f[w], 0);
      const float cur = (float)(cin_pg - 2 * popcount);
      const int efx = d->dilation_w * fx, efy = d->dilation_h * fy;
      int counted;
      switch (direction) {
        case 0: counted = efy < y || efx < x; break;
        case 1: counted = efy < y || (eff_w - efx) <= x; break;
        case 2: counted = (eff_h - efy) <= y || efx < x; break;
        default: counted = (eff_h - efy) <= y || (eff_w - efx) <= x; break;
      }
      if (counted) correction += cur;
    }
  const float mul = -1.0f * post_mul[out_c];
  volatile float r = mul * correction;
  return r;
}

/* ApplyCorrection, zero_padding_correction.h:178-299, restated literally (including its
 * "cannot happen" fall-through for images smaller than the filter). */
static void zpc_apply(const lce_bconv2d_desc* d, int out_h, int out_w,
                      const int32_t* filter, const float* post_mul, float* output) {
  const int eff_w = (d->filter_w - 1) * d->dilation_w + 1;
  const int eff_h = (d->filter_h - 1) * d->dilation_h + 1;
  const int left_off = ((out_w - 1) * d->stride_w + eff_w - d->in_w) / 2;
  const int top_off = ((out_h - 1) * d->stride_h + eff_h - d->in_h) / 2;
  const int C = d->channels_out;
  for (int b = 0; b < d->batch; ++b)
    for (int oy = 0; oy < out_h; ++oy) {
      const int o_top = top_off - oy * d->stride_h;
      const int o_bot = -o_top - d->in_h + eff_h;
      for (int ox = 0; ox < out_w; ++ox) {
        const int o_left = left_off - ox * d->stride_w;
        const int o_right = -o_left - d->in_w + eff_w;
        if (o_left <= 0 && o_right <= 0 && o_top <= 0 && o_bot <= 0) continue;
        int cs, X, Y;
        if (o_right <= 0 && o_top > 0 && o_bot < 0) {
          cs = 0; X = o_left >= 0 ? o_left : 0; Y = o_top;
        } else if (o_left < 0 && o_right > 0 && o_bot <= 0) {
          cs = 1; X = o_right; Y = o_top >= 0 ? o_top : 0;
        } else if (o_left > 0 && o_right < 0 && o_top <= 0) {
          cs = 2; X = o_left; Y = o_bot >= 0 ? o_bot : 0;
        } else if (o_left <= 0 && o_top < 0 && o_bot > 0) {
          cs = 3; X = o_right >= 0 ? o_right : 0; Y = o_bot;
        } else {
          continue;
        }
        float* o = output + (((size_t)b * out_h + oy) * out_w + ox) * C;
        for (int c = 0; c < C; ++c) {
          volatile float sum = o[c] + zpc_cache_value(d, filter, post_mul, cs, Y, X, c);
          o[c] = sum;
        }
      }
    }
}

/* Returns 0, 1 on invalid parameters, 2 when Prepare would refuse (bconv2d.cc:188-200,
 * optimised branch: zero padding needs float output and no fused activation). */
int lce_oracle_bconv2d_opt_mt(const lce_bconv2d_desc* d, int threads,
                              const int32_t* input, const int32_t* filter,
                              const float* post_mul, const float* post_bias,
                              const int32_t* thresholds, void* output) {
  const int zero_padding = d->padding == LCE_PADDING_SAME && d->pad_value == 0;
  if (!zero_padding)
    return lce_oracle_bconv2d_mt(d, threads, input, filter, post_mul, post_bias,
                                 thresholds, output);
  if (!(d->out_type == LCE_OUT_FLOAT && d->activation == LCE_ACT_NONE)) return 2;
  lce_bconv2d_desc one = *d;
  one.pad_value = 1;
  const int rc = lce_oracle_bconv2d_mt(&one, threads, input, filter, post_mul,
                                       post_bias, thresholds, output);
  if (rc) return rc;
  int out_h, out_w, pad_h, pad_w;
  if (lce_oracle_bconv2d_out_shape(d, &out_h, &out_w, &pad_h, &pad_w)) return 1;
  zpc_apply(d, out_h, out_w, filter, post_mul, (float*)output);
  return 0;
}

/* Portable BGEMM semantics: BGemmKernel<kStandardCpp>::Run,
 * LCE/core/bgemm/kernels.h:48-59 (float/int8) and :111-131 (bitpacked), with
 * the matrix orientation of optimized_bgemm.h:126-151: A = activations
 * [M, Kw] (one row per output pixel), W = filters [N, Kw]; dst is [M, N]
 * row-major (= NHWC), bitpacked dst is [M, ceil(N/32)]. */
typedef struct {
  int64_t M;
  int N, Kw;
  const int32_t *A, *W;
  const lce_bgemm_epilogue* ep;
  void* out;
  int64_t next;
  pthread_mutex_t mu;
} gemm_job;

static void bgemm_rows(const gemm_job* j, int64_t m0, int64_t m1) {
  const int N = j->N, Kw = j->Kw;
  const lce_bgemm_epilogue* ep = j->ep;
  const int nw = ceil_div(N, 32);
  for (int64_t m = m0; m < m1; ++m) {
    const int32_t* a = j->A + m * Kw;
    int32_t column = 0;
    for (int n = 0; n < N; ++n) {
      const int32_t* w = j->W + (size_t)n * Kw;
      int32_t acc = 0;
      for (int k = 0; k < Kw; ++k) acc += xor_popcount(a[k], w[k]);
      switch (ep->out_type) {
        case LCE_OUT_RAW_ACC:
          ((int32_t*)j->out)[m * N + n] = acc;
          break;
        case LCE_OUT_FLOAT:
          ((float*)j->out)[m * N + n] = transform_float(
              acc, ep->clamp_min, ep->clamp_max, ep->multiplier[n], ep->bias[n]);
          break;
        case LCE_OUT_INT8:
          ((int8_t*)j->out)[m * N + n] = round_saturate(transform_float(
              acc, ep->clamp_min, ep->clamp_max, ep->multiplier[n], ep->bias[n]));
          break;
        default:
          if (acc > ep->thresholds[n]) column |= (int32_t)(1u << (n % 32));
          if ((n + 1) % 32 == 0 || n + 1 == N) {
            ((int32_t*)j->out)[m * nw + n / 32] = column;
            column = 0;
          }
      }
    }
  }
}

static void* gemm_worker(void* arg) {
  gemm_job* j = (gemm_job*)arg;
  const int64_t chunk = 16;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    const int64_t m0 = j->next;
    j->next += chunk;
    pthread_mutex_unlock(&j->mu);
    if (m0 >= j->M) break;
    bgemm_rows(j, m0, m0 + chunk < j->M ? m0 + chunk : j->M);
  }
  return NULL;
}

int lce_oracle_bgemm_mt(int threads, int64_t M, int N, int Kw, const int32_t* A,
                        const int32_t* W, const lce_bgemm_epilogue* ep,
                        void* out) {
  gemm_job j;
  memset(&j, 0, sizeof(j));
  j.M = M; j.N = N; j.Kw = Kw; j.A = A; j.W = W; j.ep = ep; j.out = out;
  pthread_mutex_init(&j.mu, NULL);
  if (threads < 1) threads = 1;
  pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  for (int i = 1; i < threads; ++i) pthread_create(&tid[i], NULL, gemm_worker, &j);
  gemm_worker(&j);
  for (int i = 1; i < threads; ++i) pthread_join(tid[i], NULL);
  free(tid);
  pthread_mutex_destroy(&j.mu);
  return 0;
}

int lce_oracle_bgemm(int64_t M, int N, int Kw, const int32_t* A,
                     const int32_t* W, const lce_bgemm_epilogue* ep, void* out) {
  return lce_oracle_bgemm_mt(1, M, N, Kw, A, W, ep, out);
}

/* LceQuantize: QuantizeEval LCE/tflite/kernels/quantization.cc:76-114 ->
 * bitpack_matrix LCE/core/bitpacking/bitpack.h:249-308. Bit = value < zero
 * point (float: value < 0, so -0.0 and NaN pack as 0); tail bits of the last
 * word are 0 (bitpack.h:238-244); zero points outside the input type's range
 * short-circuit (bitpack.h:259-288). Bool is read as uint8 with zero point 1
 * (quantization.cc:88-108). */
int lce_oracle_quantize(int in_type, const void* in, int64_t rows, int64_t cols,
                        int32_t zero_point, int32_t* out) {
  const int64_t cw = (cols + 31) / 32;
  if (in_type != LCE_T_FLOAT && in_type != LCE_T_INT8 && in_type != LCE_T_BOOL)
    return 1;
  if (in_type == LCE_T_BOOL) zero_point = 1;
  if (in_type == LCE_T_FLOAT) zero_point = 0;
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t w = 0; w < cw; ++w) {
      uint32_t word = 0;
      const int64_t c0 = w * 32;
      const int nbits = cols - c0 < 32 ? (int)(cols - c0) : 32;
      for (int i = 0; i < nbits; ++i) {
        const int64_t idx = r * cols + c0 + i;
        int bit;
        if (in_type == LCE_T_FLOAT) bit = ((const float*)in)[idx] < 0.0f;
        else if (in_type == LCE_T_INT8)
          bit = (int32_t)((const int8_t*)in)[idx] < zero_point;
        else bit = (int32_t)((const uint8_t*)in)[idx] < zero_point;
        word |= (uint32_t)bit << i;
      }
      out[r * cw + w] = (int32_t)word;
    }
  return 0;
}

/* LceDequantize: DequantizeEval quantization.cc:116-147 -> unpack_matrix
 * bitpack.h:325-346. */
int lce_oracle_dequantize(int out_type, const int32_t* in, int64_t rows,
                          int64_t cols, float scale, int32_t zero_point,
                          void* out) {
  const int64_t cw = (cols + 31) / 32;
  int8_t zero_bit_i8 = 0, one_bit_i8 = 0;
  if (out_type == LCE_T_INT8) {
    const int offset = (int)roundf(1.0f / scale);
    int z = zero_point + offset, o = zero_point - offset;
    zero_bit_i8 = (int8_t)(z < 127 ? z : 127);
    one_bit_i8 = (int8_t)(o > -128 ? o : -128);
  } else if (out_type != LCE_T_FLOAT && out_type != LCE_T_BOOL) {
    return 1;
  }
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < cols; ++c) {
      const int bit = ((uint32_t)in[r * cw + c / 32] >> (c % 32)) & 1;
      const int64_t idx = r * cols + c;
      if (out_type == LCE_T_FLOAT) ((float*)out)[idx] = bit ? -1.0f : 1.0f;
      else if (out_type == LCE_T_INT8)
        ((int8_t*)out)[idx] = bit ? one_bit_i8 : zero_bit_i8;
      else ((uint8_t*)out)[idx] = bit ? 0 : 1;
    }
  return 0;
}

int lce_oracle_bmaxpool_out_shape(const lce_bmaxpool_desc* d, int* out_h,
                                  int* out_w) {
  if (!d->stride_h || !d->stride_w || !d->filter_h || !d->filter_w) return 1;
  *out_h = out_size(d->padding, d->in_h, d->filter_h, d->stride_h, 1);
  *out_w = out_size(d->padding, d->in_w, d->filter_w, d->stride_w, 1);
  return 0;
}

/* LceBMaxPool2d: BMaxPool LCE/core/bmaxpool.h:24-88 -- bitwise AND over the
 * in-bounds part of the window (max of +-1 values = AND of sign bits);
 * out-of-bounds positions are skipped, not padded (:45-56). */
int lce_oracle_bmaxpool(const lce_bmaxpool_desc* d, const int32_t* in,
                        int32_t* out) {
  int out_h, out_w;
  if (lce_oracle_bmaxpool_out_shape(d, &out_h, &out_w)) return 1;
  const int pad_h = pad_before(d->stride_h, 1, d->in_h, d->filter_h, out_h);
  const int pad_w = pad_before(d->stride_w, 1, d->in_w, d->filter_w, out_w);
  const int C = d->channels_packed;
  for (int b = 0; b < d->batch; ++b)
    for (int oy = 0; oy < out_h; ++oy)
      for (int ox = 0; ox < out_w; ++ox) {
        const int x0 = ox * d->stride_w - pad_w, y0 = oy * d->stride_h - pad_h;
        const int fx0 = x0 < 0 ? -x0 : 0, fy0 = y0 < 0 ? -y0 : 0;
        int fx1 = d->filter_w, fy1 = d->filter_h;
        if (x0 + fx1 > d->in_w) fx1 = d->in_w - x0;
        if (y0 + fy1 > d->in_h) fy1 = d->in_h - y0;
        int32_t* o = out + (((size_t)b * out_h + oy) * out_w + ox) * C;
        for (int c = 0; c < C; ++c) {
          int32_t m = ~(int32_t)0;
          for (int fy = fy0; fy < fy1; ++fy)
            for (int fx = fx0; fx < fx1; ++fx)
              m &= in[(((size_t)b * d->in_h + y0 + fy) * d->in_w + x0 + fx) * C + c];
          o[c] = m;
        }
      }
  return 0;
}

/* Threshold construction used by the reference's op tests for bitpacked
 * output: ComputeThresholds, LCE/tflite/tests/bconv2d_test.cc:327-368 (all
 * intermediates in double, truncating cast). Only NONE and RELU are
 * distinguished there. */
void lce_oracle_compute_thresholds(int cin_per_group, int filter_h, int filter_w,
                                   int n, const float* post_mul,
                                   const float* post_bias, int activation,
                                   int32_t* thresholds) {
  double act_min, act_max;
  if (activation == LCE_ACT_RELU) { act_min = 0; act_max = (double)INT32_MAX; }
  else { act_min = (double)INT32_MIN; act_max = (double)INT32_MAX; }
  const double backtransform_add = (double)filter_h * filter_w * cin_per_group;
  for (int i = 0; i < n; ++i) {
    const double m = post_mul[i], b = post_bias[i];
    const double t1 = -b / m;
    const double t2 = 0.5 * (backtransform_add + b / m);
    int32_t t;
    if (t2 >= 2147483648.0 || t2 < -2147483648.0 || t2 != t2) t = INT32_MIN;
    else t = (int32_t)t2;
    if (t2 >= 2 * backtransform_add || t1 <= act_min) t = INT32_MAX;
    else if (t2 <= 0.0 || t1 >= act_max) t = INT32_MIN;
    thresholds[i] = t;
  }
}

/* Threshold construction of the converter: ComputeWriteBitpackedOutputThresholds
 * + GetBitpackedOutputThresholds, LCE/mlir/transforms/optimize.cc:128-243 (float
 * intermediates, floor). The caller must also flip the filter signs of channels
 * with a negative multiplier (bitpack_activations_patterns.td:19-60). Pinned by
 * the known answer in LCE/mlir/tests/optimize.mlir:217-241. */
void lce_oracle_converter_thresholds(int cin_per_group, int filter_h,
                                     int filter_w, int n, const float* post_mul,
                                     const float* post_bias, int activation,
                                     int32_t* thresholds) {
  const int32_t bt = filter_h * filter_w * cin_per_group;
  float clamp_min, clamp_max;
  switch (activation) {
    case LCE_ACT_RELU: clamp_min = 0; clamp_max = (float)bt; break;
    case LCE_ACT_RELU_N1_TO_1: clamp_min = -1; clamp_max = 1; break;
    case LCE_ACT_RELU6: clamp_min = 0; clamp_max = 6; break;
    default: clamp_min = (float)-bt; clamp_max = (float)bt; break;
  }
  const float backtransform_add = (float)bt;
  for (int i = 0; i < n; ++i) {
    const float mult = post_mul[i], bias = post_bias[i];
    if (mult == 0.0f) {
      thresholds[i] = bias < 0.0f ? INT32_MIN : INT32_MAX;
      continue;
    }
    float emin, emax;
    if (mult > 0.0f) { emin = clamp_min; emax = clamp_max; }
    else { emin = -1 * clamp_max; emax = -1 * clamp_min; }
    const float start = emin * fabsf(mult) + bias;
    const float end = emax * fabsf(mult) + bias;
    if (start < 0 && end < 0) { thresholds[i] = INT32_MIN; continue; }
    if (start >= 0 && end >= 0) { thresholds[i] = INT32_MAX; continue; }
    thresholds[i] =
        (int32_t)floor(0.5 * (double)(bias / fabsf(mult) + backtransform_add));
  }
}

const char* lce_oracle_version(void) { return "lce_oracle port of LCE e6860fcf"; }
