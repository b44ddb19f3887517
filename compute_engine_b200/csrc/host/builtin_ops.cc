// builtin_ops.cc -- TfLiteRegistrations for the float builtins QuickNet /
// QuickNetLarge / Bi-RealNet-18 use around the binary path, backed by
// the fp32 CUDA kernels of lce_b200_builtins.cu. Shape inference follows TFLite's
// builtin kernels (tensorflow/lite/kernels/{conv,depthwise_conv,pooling,add,mul,
// fully_connected,softmax,reduce,reshape,activations}.cc). Device arena only.
#include <cuda_runtime.h>

#include <cstring>

#include "builtin_params.h"
#include "host_graph.h"
#include "lce_b200.h"
#include "lce_b200_builtins.h"

namespace lce_b200 {
namespace {

#define B_ENSURE(ctx, cond, msg)                 \
  do {                                           \
    if (!(cond)) {                               \
      (ctx)->ReportError((ctx), "%s", (msg));    \
      return kTfLiteError;                       \
    }                                            \
  } while (0)
#define B_CAPI(ctx, call)                                     \
  do {                                                        \
    if ((call) != 0) {                                        \
      (ctx)->ReportError((ctx), "%s", lce_b200_last_error()); \
      return kTfLiteError;                                    \
    }                                                         \
  } while (0)

TfLiteTensor* T(TfLiteContext* c, const TfLiteIntArray* a, int i) {
  if (i >= a->size || a->data[i] < 0) return nullptr;
  return &c->tensors[a->data[i]];
}
int64_t Count(const TfLiteTensor* t) {
  int64_t n = 1;
  for (int i = 0; i < t->dims->size; ++i) n *= t->dims->data[i];
  return n;
}
TfLiteStatus Resize(TfLiteContext* c, TfLiteTensor* t, std::initializer_list<int> dims) {
  TfLiteIntArray* a = LceB200IntArrayCreate(static_cast<int>(dims.size()));
  int i = 0;
  for (int d : dims) a->data[i++] = d;
  return c->ResizeTensor(c, t, a);
}
bool OnDevice(const void* p) {
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

void* Init(TfLiteContext*, const char* buffer, size_t length) {
  auto* p = new BuiltinParams();
  memset(p, 0, sizeof(*p));
  if (buffer && length >= sizeof(BuiltinParams)) memcpy(p, buffer, sizeof(BuiltinParams));
  return p;
}
void Free(TfLiteContext*, void* p) { delete static_cast<BuiltinParams*>(p); }
const BuiltinParams& P(TfLiteNode* n) { return *static_cast<BuiltinParams*>(n->user_data); }

// ---------------------------- CONV_2D / DEPTHWISE ---------------------------- //
bool ConvDesc(TfLiteContext* c, TfLiteNode* n, bool depthwise, lce_f32_conv_desc* d) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const TfLiteTensor* f = T(c, n->inputs, 1);
  if (!in || !f || in->dims->size != 4 || f->dims->size != 4) return false;
  const BuiltinParams& p = P(n);
  d->batch = in->dims->data[0];
  d->in_h = in->dims->data[1];
  d->in_w = in->dims->data[2];
  d->in_c = in->dims->data[3];
  d->filter_h = f->dims->data[1];
  d->filter_w = f->dims->data[2];
  d->out_c = depthwise ? f->dims->data[3] : f->dims->data[0];
  d->stride_h = p.stride_h;
  d->stride_w = p.stride_w;
  d->dilation_h = p.dilation_h;
  d->dilation_w = p.dilation_w;
  d->padding = p.padding;
  d->activation = p.activation;
  if (!depthwise && f->dims->data[3] != d->in_c) return false;
  return true;
}
template <bool DW>
TfLiteStatus ConvPrepare(TfLiteContext* c, TfLiteNode* n) {
  lce_f32_conv_desc d;
  B_ENSURE(c, ConvDesc(c, n, DW, &d), "CONV_2D: bad input / filter shapes");
  B_ENSURE(c, T(c, n->inputs, 0)->type == kTfLiteFloat32, "CONV_2D: only float32 is supported");
  if (DW) B_ENSURE(c, P(n).depth_multiplier == 1 && d.out_c == d.in_c,
                   "DEPTHWISE_CONV_2D: depth_multiplier must be 1");
  int oh, ow;
  B_CAPI(c, lce_b200_f32_conv_out_shape(&d, &oh, &ow));
  if (!DW && n->outputs->size == 2) {
    // fused LceQuantize (Graph::FuseFloatGlue): second output = bitpacked signs of the first
    TfLiteTensor* pk = T(c, n->outputs, 1);
    B_ENSURE(c, pk && pk->type == kTfLiteInt32, "CONV_2D+LceQuantize: packed output must be int32");
    if (Resize(c, pk, {d.batch, oh, ow, (d.out_c + 31) / 32}) != kTfLiteOk) return kTfLiteError;
  }
  return Resize(c, T(c, n->outputs, 0), {d.batch, oh, ow, d.out_c});
}
template <bool DW>
TfLiteStatus ConvInvoke(TfLiteContext* c, TfLiteNode* n) {
  lce_f32_conv_desc d;
  ConvDesc(c, n, DW, &d);
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const TfLiteTensor* f = T(c, n->inputs, 1);
  const TfLiteTensor* bias = T(c, n->inputs, 2);
  TfLiteTensor* out = T(c, n->outputs, 0);
  B_ENSURE(c, OnDevice(in->data.raw) && OnDevice(f->data.raw) && OnDevice(out->data.raw),
           "float builtins need a device arena (no CPU path)");
  if (DW)
    B_CAPI(c, lce_b200_f32_depthwise_conv2d(&d, in->data.f, f->data.f, bias ? bias->data.f : nullptr,
                                             out->data.f, lce_b200_get_stream()));
  else if (n->outputs->size == 2)
    B_CAPI(c, lce_b200_f32_conv2d_packed(&d, in->data.f, f->data.f, bias ? bias->data.f : nullptr,
                                         out->data.f, T(c, n->outputs, 1)->data.i32,
                                         lce_b200_get_stream()));
  else
    B_CAPI(c, lce_b200_f32_conv2d(&d, in->data.f, f->data.f, bias ? bias->data.f : nullptr,
                                  out->data.f, lce_b200_get_stream()));
  return kTfLiteOk;
}

// ------------------------------ FULLY_CONNECTED ------------------------------ //
TfLiteStatus FcPrepare(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const TfLiteTensor* w = T(c, n->inputs, 1);
  B_ENSURE(c, in && w && w->dims->size == 2, "FULLY_CONNECTED: weights must be [out, in]");
  const int k = w->dims->data[1];
  B_ENSURE(c, k > 0 && Count(in) % k == 0, "FULLY_CONNECTED: input size mismatch");
  return Resize(c, T(c, n->outputs, 0), {static_cast<int>(Count(in) / k), w->dims->data[0]});
}
TfLiteStatus FcInvoke(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const TfLiteTensor* w = T(c, n->inputs, 1);
  const TfLiteTensor* bias = T(c, n->inputs, 2);
  TfLiteTensor* out = T(c, n->outputs, 0);
  lce_f32_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.batch = static_cast<int>(Count(in) / w->dims->data[1]);
  d.in_h = d.in_w = d.filter_h = d.filter_w = d.stride_h = d.stride_w = d.dilation_h =
      d.dilation_w = 1;
  d.in_c = w->dims->data[1];
  d.out_c = w->dims->data[0];
  d.padding = LCE_PADDING_VALID;
  d.activation = P(n).activation;
  B_CAPI(c, lce_b200_f32_conv2d(&d, in->data.f, w->data.f, bias ? bias->data.f : nullptr,
                                out->data.f, lce_b200_get_stream()));
  return kTfLiteOk;
}

// ---------------------------------- pooling ---------------------------------- //
bool PoolDesc(TfLiteContext* c, TfLiteNode* n, lce_f32_pool_desc* d) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  if (!in || in->dims->size != 4) return false;
  const BuiltinParams& p = P(n);
  d->batch = in->dims->data[0]; d->in_h = in->dims->data[1]; d->in_w = in->dims->data[2];
  d->channels = in->dims->data[3];
  d->filter_h = p.filter_h; d->filter_w = p.filter_w;
  d->stride_h = p.stride_h; d->stride_w = p.stride_w;
  d->padding = p.padding; d->activation = p.activation;
  return true;
}
TfLiteStatus PoolPrepare(TfLiteContext* c, TfLiteNode* n) {
  lce_f32_pool_desc d;
  B_ENSURE(c, PoolDesc(c, n, &d), "POOL_2D: input must be 4-D");
  int oh, ow;
  B_CAPI(c, lce_b200_f32_pool_out_shape(&d, &oh, &ow));
  return Resize(c, T(c, n->outputs, 0), {d.batch, oh, ow, d.channels});
}
template <bool MAX>
TfLiteStatus PoolInvoke(TfLiteContext* c, TfLiteNode* n) {
  lce_f32_pool_desc d;
  PoolDesc(c, n, &d);
  const float* in = T(c, n->inputs, 0)->data.f;
  float* out = T(c, n->outputs, 0)->data.f;
  if (MAX) B_CAPI(c, lce_b200_f32_max_pool(&d, in, out, lce_b200_get_stream()));
  else B_CAPI(c, lce_b200_f32_avg_pool(&d, in, out, lce_b200_get_stream()));
  return kTfLiteOk;
}

// ------------- fused MAX_POOL_2D(2x2 s1 VALID) + DEPTHWISE_CONV_2D(3x3) ------------- //
// Created by Graph::FuseFloatGlue; builtin_data = { BuiltinParams pool, BuiltinParams dw }.
struct PoolDwParams {
  BuiltinParams pool, dw;
};
void* PoolDwInit(TfLiteContext*, const char* buffer, size_t length) {
  auto* p = new PoolDwParams();
  memset(p, 0, sizeof(*p));
  if (buffer && length >= sizeof(PoolDwParams)) memcpy(p, buffer, sizeof(PoolDwParams));
  return p;
}
void PoolDwFree(TfLiteContext*, void* p) { delete static_cast<PoolDwParams*>(p); }
bool PoolDwDescs(TfLiteContext* c, TfLiteNode* n, lce_f32_pool_desc* pd, lce_f32_conv_desc* dd) {
  const auto& pp = *static_cast<PoolDwParams*>(n->user_data);
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const TfLiteTensor* f = T(c, n->inputs, 1);
  if (!in || !f || in->dims->size != 4 || f->dims->size != 4) return false;
  pd->batch = in->dims->data[0]; pd->in_h = in->dims->data[1]; pd->in_w = in->dims->data[2];
  pd->channels = in->dims->data[3];
  pd->filter_h = pp.pool.filter_h; pd->filter_w = pp.pool.filter_w;
  pd->stride_h = pp.pool.stride_h; pd->stride_w = pp.pool.stride_w;
  pd->padding = pp.pool.padding; pd->activation = pp.pool.activation;
  int ph, pw;
  if (lce_b200_f32_pool_out_shape(pd, &ph, &pw)) return false;
  dd->batch = pd->batch; dd->in_h = ph; dd->in_w = pw; dd->in_c = pd->channels;
  dd->filter_h = f->dims->data[1]; dd->filter_w = f->dims->data[2]; dd->out_c = f->dims->data[3];
  dd->stride_h = pp.dw.stride_h; dd->stride_w = pp.dw.stride_w;
  dd->dilation_h = pp.dw.dilation_h; dd->dilation_w = pp.dw.dilation_w;
  dd->padding = pp.dw.padding; dd->activation = pp.dw.activation;
  return true;
}
TfLiteStatus PoolDwPrepare(TfLiteContext* c, TfLiteNode* n) {
  lce_f32_pool_desc pd;
  lce_f32_conv_desc dd;
  B_ENSURE(c, PoolDwDescs(c, n, &pd, &dd), "MAX_POOL_2D+DEPTHWISE_CONV_2D: bad shapes");
  int oh, ow;
  B_CAPI(c, lce_b200_f32_conv_out_shape(&dd, &oh, &ow));
  return Resize(c, T(c, n->outputs, 0), {dd.batch, oh, ow, dd.out_c});
}
TfLiteStatus PoolDwInvoke(TfLiteContext* c, TfLiteNode* n) {
  lce_f32_pool_desc pd;
  lce_f32_conv_desc dd;
  PoolDwDescs(c, n, &pd, &dd);
  const TfLiteTensor* bias = T(c, n->inputs, 2);
  B_CAPI(c, lce_b200_f32_maxpool2x2_depthwise3x3(&pd, &dd, T(c, n->inputs, 0)->data.f,
                                                   T(c, n->inputs, 1)->data.f,
                                                   bias ? bias->data.f : nullptr,
                                                   T(c, n->outputs, 0)->data.f,
                                                   lce_b200_get_stream()));
  return kTfLiteOk;
}

// ---- fused stem [DEQUANTIZE +] CONV_2D(3x3 s2, 3 -> 16) + DEPTHWISE_CONV_2D(3x3 s2) ---- //
// Created by Graph::FuseStem; builtin_data = { conv1, depthwise } BuiltinParams;
// inputs [x (float32, or the int8 / uint8 input of the removed DEQUANTIZE), w1, b1, w2, b2].
struct StemBlob {
  BuiltinParams c1, dw;
};
struct StemParams {
  BuiltinParams c1, dw;
  // host copies of the (constant) filters and biases: the kernel takes them by value
  std::vector<float> w1, b1, w2, b2;
};
void* StemInit(TfLiteContext*, const char* buffer, size_t length) {
  auto* p = new StemParams();
  memset(&p->c1, 0, sizeof(p->c1));
  memset(&p->dw, 0, sizeof(p->dw));
  if (buffer && length >= sizeof(StemBlob)) {
    memcpy(&p->c1, buffer, sizeof(BuiltinParams));
    memcpy(&p->dw, buffer + sizeof(BuiltinParams), sizeof(BuiltinParams));
  }
  return p;
}
// Copy a constant float tensor to the host (device constants: one synchronous copy, at Prepare).
bool HostCopy(const TfLiteTensor* t, std::vector<float>* out) {
  out->clear();
  if (!t) return true;   // optional bias
  if (t->type != kTfLiteFloat32 || t->allocation_type != kTfLiteMmapRo || !t->data.raw) return false;
  out->resize(t->bytes / sizeof(float));
  if (OnDevice(t->data.raw))
    return cudaMemcpy(out->data(), t->data.raw, out->size() * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess;
  memcpy(out->data(), t->data.raw, out->size() * sizeof(float));
  return true;
}
void StemFree(TfLiteContext*, void* p) { delete static_cast<StemParams*>(p); }
void FillConvDesc(lce_f32_conv_desc* d, const BuiltinParams& bp, int batch, int h, int w, int cin,
                  const TfLiteTensor* filter) {
  d->batch = batch; d->in_h = h; d->in_w = w; d->in_c = cin;
  d->filter_h = filter->dims->data[1]; d->filter_w = filter->dims->data[2];
  d->stride_h = bp.stride_h; d->stride_w = bp.stride_w;
  d->dilation_h = bp.dilation_h; d->dilation_w = bp.dilation_w;
  d->padding = bp.padding; d->activation = bp.activation;
}
bool StemDescs(TfLiteContext* c, TfLiteNode* n, lce_f32_conv_desc* d1, lce_f32_conv_desc* d2) {
  const auto& sp = *static_cast<StemParams*>(n->user_data);
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const TfLiteTensor* w1 = T(c, n->inputs, 1);
  const TfLiteTensor* w2 = T(c, n->inputs, 3);
  if (!in || !w1 || !w2 || in->dims->size != 4 || w1->dims->size != 4 || w2->dims->size != 4)
    return false;
  int oh, ow;
  FillConvDesc(d1, sp.c1, in->dims->data[0], in->dims->data[1], in->dims->data[2],
               in->dims->data[3], w1);
  d1->out_c = w1->dims->data[0];
  if (lce_b200_f32_conv_out_shape(d1, &oh, &ow)) return false;
  FillConvDesc(d2, sp.dw, d1->batch, oh, ow, d1->out_c, w2);
  d2->out_c = w2->dims->data[3];
  return true;
}
TfLiteStatus StemPrepare(TfLiteContext* c, TfLiteNode* n) {
  lce_f32_conv_desc d1, d2;
  B_ENSURE(c, StemDescs(c, n, &d1, &d2), "fused stem: bad shapes");
  const TfLiteTensor* in = T(c, n->inputs, 0);
  B_ENSURE(c, in->type == kTfLiteFloat32 || in->type == kTfLiteInt8 || in->type == kTfLiteUInt8,
           "fused stem: the input must be float32, int8 or uint8");
  B_ENSURE(c, d1.in_c == 3, "fused stem: three input channels expected");
  auto& sp = *static_cast<StemParams*>(n->user_data);
  if (sp.w1.empty()) {
    B_ENSURE(c, HostCopy(T(c, n->inputs, 1), &sp.w1) && HostCopy(T(c, n->inputs, 2), &sp.b1) &&
                    HostCopy(T(c, n->inputs, 3), &sp.w2) && HostCopy(T(c, n->inputs, 4), &sp.b2) &&
                    sp.w1.size() == 16 * 27 && sp.w2.size() == 9 * 16 && (sp.b1.empty() || sp.b1.size() == 16) &&
                    (sp.b2.empty() || sp.b2.size() == 16),
             "fused stem: the filters and biases must be constant float tensors");
  }
  int oh, ow;
  B_CAPI(c, lce_b200_f32_conv_out_shape(&d2, &oh, &ow));
  return Resize(c, T(c, n->outputs, 0), {d2.batch, oh, ow, d2.out_c});
}
TfLiteStatus StemInvoke(TfLiteContext* c, TfLiteNode* n) {
  lce_f32_conv_desc d1, d2;
  StemDescs(c, n, &d1, &d2);
  const auto& sp = *static_cast<const StemParams*>(n->user_data);
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const int in_type = in->type == kTfLiteFloat32 ? LCE_T_FLOAT : in->type == kTfLiteInt8 ? LCE_T_INT8 : LCE_T_BOOL;
  B_CAPI(c, lce_b200_f32_stem_conv_dw(&d1, &d2, in_type, in->data.raw, in->params.scale, in->params.zero_point,
                                      sp.w1.data(), sp.b1.empty() ? nullptr : sp.b1.data(), sp.w2.data(),
                                      sp.b2.empty() ? nullptr : sp.b2.data(), T(c, n->outputs, 0)->data.f,
                                      lce_b200_get_stream()));
  return kTfLiteOk;
}

// ------------------------------ ADD / MUL / RELU ------------------------------ //
TfLiteStatus EltPrepare(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* a = T(c, n->inputs, 0);
  const TfLiteTensor* b = T(c, n->inputs, 1);
  B_ENSURE(c, a && b, "ADD/MUL: two inputs required");
  const int64_t na = Count(a), nb = Count(b);
  B_ENSURE(c, nb > 0 && na % nb == 0 &&
                  (na == nb || nb == a->dims->data[a->dims->size - 1]),
           "ADD/MUL: only same-shape or last-dimension broadcast is supported");
  TfLiteIntArray* dims = LceB200IntArrayCreate(a->dims->size);
  for (int i = 0; i < a->dims->size; ++i) dims->data[i] = a->dims->data[i];
  return c->ResizeTensor(c, T(c, n->outputs, 0), dims);
}
template <bool MUL>
TfLiteStatus EltInvoke(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* a = T(c, n->inputs, 0);
  const TfLiteTensor* b = T(c, n->inputs, 1);
  TfLiteTensor* out = T(c, n->outputs, 0);
  B_ENSURE(c, OnDevice(b->data.raw), "float builtins need device-resident operands");
  if (MUL) B_CAPI(c, lce_b200_f32_mul(a->data.f, b->data.f, out->data.f, Count(a), Count(b),
                                      P(n).activation, lce_b200_get_stream()));
  else B_CAPI(c, lce_b200_f32_add(a->data.f, b->data.f, out->data.f, Count(a), Count(b),
                                  P(n).activation, lce_b200_get_stream()));
  return kTfLiteOk;
}
TfLiteStatus SameShapePrepare(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* a = T(c, n->inputs, 0);
  TfLiteIntArray* dims = LceB200IntArrayCreate(a->dims->size);
  for (int i = 0; i < a->dims->size; ++i) dims->data[i] = a->dims->data[i];
  return c->ResizeTensor(c, T(c, n->outputs, 0), dims);
}
TfLiteStatus ReluInvoke(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* a = T(c, n->inputs, 0);
  B_CAPI(c, lce_b200_f32_activation(a->data.f, T(c, n->outputs, 0)->data.f, Count(a), LCE_ACT_RELU,
                                    lce_b200_get_stream()));
  return kTfLiteOk;
}

// --------------------------------- DEQUANTIZE -------------------------------- //
// int8 / uint8 -> float32 with the input tensor's (scale, zero_point): TF/lite/kernels/dequantize.cc.
TfLiteStatus DequantPrepare(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* a = T(c, n->inputs, 0);
  TfLiteTensor* o = T(c, n->outputs, 0);
  B_ENSURE(c, a && o && (a->type == kTfLiteInt8 || a->type == kTfLiteUInt8) && o->type == kTfLiteFloat32,
           "DEQUANTIZE: only int8 / uint8 -> float32 is supported");
  return SameShapePrepare(c, n);
}
TfLiteStatus DequantInvoke(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* a = T(c, n->inputs, 0);
  B_CAPI(c, lce_b200_dequantize_affine(a->type == kTfLiteInt8 ? LCE_T_INT8 : LCE_T_BOOL, a->data.raw,
                                       T(c, n->outputs, 0)->data.f, Count(a), a->params.scale,
                                       a->params.zero_point, lce_b200_get_stream()));
  return kTfLiteOk;
}

// ----------------------------------- MEAN ----------------------------------- //
TfLiteStatus MeanPrepare(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const TfLiteTensor* axes = T(c, n->inputs, 1);
  B_ENSURE(c, in && in->dims->size == 4, "MEAN: input must be 4-D");
  // only the global-average-pool form (axes = {1, 2}) is needed by these models
  bool ok = axes && axes->type == kTfLiteInt32 && Count(axes) == 2 && !OnDevice(axes->data.raw);
  if (ok) {
    const int a0 = axes->data.i32[0], a1 = axes->data.i32[1];
    ok = (a0 == 1 && a1 == 2) || (a0 == 2 && a1 == 1);
  }
  B_ENSURE(c, ok, "MEAN: only reduction over axes {1,2} is supported");
  if (P(n).keep_dims)
    return Resize(c, T(c, n->outputs, 0), {in->dims->data[0], 1, 1, in->dims->data[3]});
  return Resize(c, T(c, n->outputs, 0), {in->dims->data[0], in->dims->data[3]});
}
TfLiteStatus MeanInvoke(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  // P(n).activation: a RELU folded in front of the mean by Graph::FuseFloatGlue (0 = none)
  B_CAPI(c, lce_b200_f32_mean_hw_act(in->data.f, T(c, n->outputs, 0)->data.f, in->dims->data[0],
                                     in->dims->data[1], in->dims->data[2], in->dims->data[3],
                                     P(n).activation, lce_b200_get_stream()));
  return kTfLiteOk;
}

// ---------------------------------- SOFTMAX ---------------------------------- //
TfLiteStatus SoftmaxInvoke(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const int cols = in->dims->data[in->dims->size - 1];
  B_CAPI(c, lce_b200_f32_softmax(in->data.f, T(c, n->outputs, 0)->data.f,
                                 cols ? Count(in) / cols : 0, cols, P(n).beta,
                                 lce_b200_get_stream()));
  return kTfLiteOk;
}

// ---------------------------------- RESHAPE ---------------------------------- //
TfLiteStatus ReshapePrepare(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const BuiltinParams& p = P(n);
  std::vector<int> shape(p.new_shape, p.new_shape + p.n_new_shape);
  const TfLiteTensor* st = T(c, n->inputs, 1);
  if (shape.empty() && st && st->type == kTfLiteInt32 && !OnDevice(st->data.raw))
    shape.assign(st->data.i32, st->data.i32 + Count(st));
  int64_t known = 1;
  int wild = -1;
  for (size_t i = 0; i < shape.size(); ++i) {
    if (shape[i] == -1) wild = static_cast<int>(i);
    else known *= shape[i];
  }
  if (wild >= 0) shape[wild] = known ? static_cast<int>(Count(in) / known) : 0;
  int64_t total = 1;
  for (int d : shape) total *= d;
  B_ENSURE(c, total == Count(in), "RESHAPE: element count mismatch");
  TfLiteIntArray* dims = LceB200IntArrayCreate(static_cast<int>(shape.size()));
  for (size_t i = 0; i < shape.size(); ++i) dims->data[i] = shape[i];
  return c->ResizeTensor(c, T(c, n->outputs, 0), dims);
}
TfLiteStatus ReshapeInvoke(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  TfLiteTensor* out = T(c, n->outputs, 0);
  if (in->bytes && cudaMemcpyAsync(out->data.raw, in->data.raw, in->bytes, cudaMemcpyDeviceToDevice,
                                   static_cast<cudaStream_t>(lce_b200_get_stream())) != cudaSuccess) {
    c->ReportError(c, "RESHAPE: device copy failed");
    return kTfLiteError;
  }
  return kTfLiteOk;
}

// ------------------------------- PAD / PADV2 ------------------------------- //
// inputs: [x, paddings int32 [rank,2] (host constant), (PADV2: constant value, host scalar)].
// float32 and int32 (bitpacked) tensors of rank <= 4 (TF/lite/kernels/pad.cc).
bool PadGeometry(TfLiteContext* c, TfLiteNode* n, int32_t in4[4], int32_t before[4],
                 int32_t after[4], int* rank) {
  const TfLiteTensor* in = T(c, n->inputs, 0);
  const TfLiteTensor* pads = T(c, n->inputs, 1);
  if (!in || !pads || pads->type != kTfLiteInt32 || OnDevice(pads->data.raw) || !pads->data.raw)
    return false;
  const int r = in->dims->size;
  if (r < 1 || r > 4 || Count(pads) != 2 * r) return false;
  for (int i = 0; i < 4; ++i) { in4[i] = 1; before[i] = after[i] = 0; }
  for (int i = 0; i < r; ++i) {
    in4[4 - r + i] = in->dims->data[i];
    before[4 - r + i] = pads->data.i32[2 * i];
    after[4 - r + i] = pads->data.i32[2 * i + 1];
    if (before[4 - r + i] < 0 || after[4 - r + i] < 0) return false;
  }
  *rank = r;
  return true;
}
TfLiteStatus PadPrepare(TfLiteContext* c, TfLiteNode* n) {
  int32_t in4[4], before[4], after[4];
  int r = 0;
  const TfLiteTensor* in = T(c, n->inputs, 0);
  B_ENSURE(c, in && (in->type == kTfLiteFloat32 || in->type == kTfLiteInt32),
           "PAD: only float32 / int32 tensors are supported");
  B_ENSURE(c, PadGeometry(c, n, in4, before, after, &r),
           "PAD: paddings must be a constant int32 [rank, 2] tensor with rank <= 4");
  TfLiteIntArray* dims = LceB200IntArrayCreate(r);
  for (int i = 0; i < r; ++i) dims->data[i] = in4[4 - r + i] + before[4 - r + i] + after[4 - r + i];
  return c->ResizeTensor(c, T(c, n->outputs, 0), dims);
}
TfLiteStatus PadInvoke(TfLiteContext* c, TfLiteNode* n) {
  int32_t in4[4], before[4], after[4];
  int r = 0;
  B_ENSURE(c, PadGeometry(c, n, in4, before, after, &r), "PAD: bad paddings");
  uint32_t fill = 0;  // 0.0f / bitpacked +1
  const TfLiteTensor* v = T(c, n->inputs, 2);
  if (v) {
    B_ENSURE(c, v->data.raw && !OnDevice(v->data.raw) && Count(v) == 1 && v->bytes == 4,
             "PADV2: constant_values must be a constant 32-bit scalar");
    memcpy(&fill, v->data.raw, 4);
  }
  B_CAPI(c, lce_b200_pad4d_32(T(c, n->inputs, 0)->data.raw, T(c, n->outputs, 0)->data.raw, in4,
                              before, after, fill, lce_b200_get_stream()));
  return kTfLiteOk;
}

// ------------------------------ CONCATENATION ------------------------------ //
// Strided device copies (one cudaMemcpy2DAsync per input): no kernel of ours.
TfLiteStatus ConcatPrepare(TfLiteContext* c, TfLiteNode* n) {
  const TfLiteTensor* a = T(c, n->inputs, 0);
  B_ENSURE(c, a && n->inputs->size >= 1, "CONCATENATION: no inputs");
  const int r = a->dims->size;
  int axis = P(n).axis < 0 ? P(n).axis + r : P(n).axis;
  B_ENSURE(c, axis >= 0 && axis < r, "CONCATENATION: axis out of range");
  B_ENSURE(c, P(n).activation == 0, "CONCATENATION: fused activation is not supported");
  int sum = 0;
  for (int k = 0; k < n->inputs->size; ++k) {
    const TfLiteTensor* t = T(c, n->inputs, k);
    B_ENSURE(c, t && t->dims->size == r && t->type == a->type,
             "CONCATENATION: inputs must agree in rank and type");
    for (int i = 0; i < r; ++i)
      B_ENSURE(c, i == axis || t->dims->data[i] == a->dims->data[i],
               "CONCATENATION: inputs must agree outside the axis");
    sum += t->dims->data[axis];
  }
  TfLiteIntArray* dims = LceB200IntArrayCreate(r);
  for (int i = 0; i < r; ++i) dims->data[i] = i == axis ? sum : a->dims->data[i];
  return c->ResizeTensor(c, T(c, n->outputs, 0), dims);
}
TfLiteStatus ConcatInvoke(TfLiteContext* c, TfLiteNode* n) {
  TfLiteTensor* out = T(c, n->outputs, 0);
  const int r = out->dims->size;
  const int axis = P(n).axis < 0 ? P(n).axis + r : P(n).axis;
  size_t outer = 1, inner = out->bytes / std::max<int64_t>(Count(out), 1);  // element size
  for (int i = 0; i < axis; ++i) outer *= out->dims->data[i];
  for (int i = axis + 1; i < r; ++i) inner *= out->dims->data[i];
  const size_t out_pitch = inner * out->dims->data[axis];
  size_t off = 0;
  for (int k = 0; k < n->inputs->size; ++k) {
    const TfLiteTensor* t = T(c, n->inputs, k);
    const size_t w = inner * t->dims->data[axis];
    if (w && outer &&
        cudaMemcpy2DAsync(out->data.raw + off, out_pitch, t->data.raw, w, w, outer,
                          cudaMemcpyDefault,
                          static_cast<cudaStream_t>(lce_b200_get_stream())) != cudaSuccess) {
      c->ReportError(c, "CONCATENATION: device copy failed");
      return kTfLiteError;
    }
    off += w;
  }
  return kTfLiteOk;
}

}  // namespace

const TfLiteRegistration* FusedStemRegistration() {
  static TfLiteRegistration r = {StemInit, StemFree, StemPrepare, StemInvoke};
  return &r;
}
const TfLiteRegistration* FusedPoolDepthwiseRegistration() {
  static TfLiteRegistration r = {PoolDwInit, PoolDwFree, PoolDwPrepare, PoolDwInvoke};
  return &r;
}

void RegisterBuiltinOps(OpResolver* r) {
  static TfLiteRegistration conv = {Init, Free, ConvPrepare<false>, ConvInvoke<false>};
  static TfLiteRegistration dw = {Init, Free, ConvPrepare<true>, ConvInvoke<true>};
  static TfLiteRegistration fc = {Init, Free, FcPrepare, FcInvoke};
  static TfLiteRegistration maxp = {Init, Free, PoolPrepare, PoolInvoke<true>};
  static TfLiteRegistration avgp = {Init, Free, PoolPrepare, PoolInvoke<false>};
  static TfLiteRegistration add = {Init, Free, EltPrepare, EltInvoke<false>};
  static TfLiteRegistration mul = {Init, Free, EltPrepare, EltInvoke<true>};
  static TfLiteRegistration relu = {Init, Free, SameShapePrepare, ReluInvoke};
  static TfLiteRegistration mean = {Init, Free, MeanPrepare, MeanInvoke};
  static TfLiteRegistration softmax = {Init, Free, SameShapePrepare, SoftmaxInvoke};
  static TfLiteRegistration reshape = {Init, Free, ReshapePrepare, ReshapeInvoke};
  // BuiltinOperator codes, schema.fbs:259-325
  r->AddBuiltin(3, &conv);
  r->AddBuiltin(4, &dw);
  r->AddBuiltin(9, &fc);
  r->AddBuiltin(17, &maxp);
  r->AddBuiltin(1, &avgp);
  r->AddBuiltin(0, &add);
  r->AddBuiltin(18, &mul);
  r->AddBuiltin(19, &relu);
  r->AddBuiltin(40, &mean);
  r->AddBuiltin(25, &softmax);
  r->AddBuiltin(22, &reshape);
  static TfLiteRegistration pad = {Init, Free, PadPrepare, PadInvoke};
  static TfLiteRegistration concat = {Init, Free, ConcatPrepare, ConcatInvoke};
  r->AddBuiltin(34, &pad);   // PAD
  r->AddBuiltin(60, &pad);   // PADV2
  r->AddBuiltin(2, &concat);  // CONCATENATION
  static TfLiteRegistration dequant = {Init, Free, DequantPrepare, DequantInvoke};
  r->AddBuiltin(6, &dequant);  // DEQUANTIZE
}

}  // namespace lce_b200
