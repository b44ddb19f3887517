// oracle/abi_check.cc -- TEST INFRASTRUCTURE: compile-time proof that the ABI
// mirror in include/lce_b200_tflite.h lays out TfLiteTensor / TfLiteNode /
// TfLiteContext / TfLiteRegistration exactly like the reference's vendored
// tensorflow/lite/core/c/common.h (TF 2.16.1). Compiled (not run) by
// oracle/Makefile when /root/reference is present; tests/test_abi.py drives it.
#include <cstddef>

#include "tensorflow/lite/core/c/common.h"  // the real one (global namespace)

namespace mirror {
#include "../include/lce_b200_tflite.h"
}

#define SAME_SIZE(T) static_assert(sizeof(::T) == sizeof(mirror::T), "sizeof " #T)
#define SAME_OFF(T, m) \
  static_assert(offsetof(::T, m) == offsetof(mirror::T, m), "offsetof " #T "::" #m)

SAME_SIZE(TfLiteIntArray);
SAME_OFF(TfLiteIntArray, data);
SAME_SIZE(TfLiteQuantizationParams);
SAME_SIZE(TfLiteQuantization);
SAME_SIZE(TfLiteAffineQuantization);
SAME_OFF(TfLiteAffineQuantization, quantized_dimension);
SAME_SIZE(TfLitePtrUnion);

SAME_SIZE(TfLiteTensor);
SAME_OFF(TfLiteTensor, type);
SAME_OFF(TfLiteTensor, data);
SAME_OFF(TfLiteTensor, dims);
SAME_OFF(TfLiteTensor, params);
SAME_OFF(TfLiteTensor, allocation_type);
SAME_OFF(TfLiteTensor, bytes);
SAME_OFF(TfLiteTensor, allocation);
SAME_OFF(TfLiteTensor, name);
SAME_OFF(TfLiteTensor, buffer_handle);
SAME_OFF(TfLiteTensor, is_variable);
SAME_OFF(TfLiteTensor, quantization);
SAME_OFF(TfLiteTensor, dims_signature);

SAME_SIZE(TfLiteNode);
SAME_OFF(TfLiteNode, inputs);
SAME_OFF(TfLiteNode, outputs);
SAME_OFF(TfLiteNode, temporaries);
SAME_OFF(TfLiteNode, user_data);
SAME_OFF(TfLiteNode, custom_initial_data);
SAME_OFF(TfLiteNode, custom_initial_data_size);

SAME_SIZE(TfLiteContext);
SAME_OFF(TfLiteContext, tensors_size);
SAME_OFF(TfLiteContext, tensors);
SAME_OFF(TfLiteContext, impl_);
SAME_OFF(TfLiteContext, ResizeTensor);
SAME_OFF(TfLiteContext, ReportError);
SAME_OFF(TfLiteContext, AddTensors);
SAME_OFF(TfLiteContext, recommended_num_threads);
SAME_OFF(TfLiteContext, GetExternalContext);
SAME_OFF(TfLiteContext, SetExternalContext);
SAME_OFF(TfLiteContext, GetTensor);

SAME_SIZE(TfLiteRegistration);
SAME_OFF(TfLiteRegistration, init);
SAME_OFF(TfLiteRegistration, free);
SAME_OFF(TfLiteRegistration, prepare);
SAME_OFF(TfLiteRegistration, invoke);
SAME_OFF(TfLiteRegistration, builtin_code);
SAME_OFF(TfLiteRegistration, custom_name);
SAME_OFF(TfLiteRegistration, version);
SAME_OFF(TfLiteRegistration, inplace_operator);

static_assert(int(::kTfLiteFloat32) == int(mirror::kTfLiteFloat32), "");
static_assert(int(::kTfLiteInt32) == int(mirror::kTfLiteInt32), "");
static_assert(int(::kTfLiteInt8) == int(mirror::kTfLiteInt8), "");
static_assert(int(::kTfLiteBool) == int(mirror::kTfLiteBool), "");
static_assert(int(::kTfLiteMmapRo) == int(mirror::kTfLiteMmapRo), "");
static_assert(int(::kTfLiteArenaRw) == int(mirror::kTfLiteArenaRw), "");
static_assert(int(::kTfLiteDynamic) == int(mirror::kTfLiteDynamic), "");
static_assert(int(::kTfLiteCustom) == int(mirror::kTfLiteCustom), "");
static_assert(int(::kTfLiteAffineQuantization) == int(mirror::kTfLiteAffineQuantization), "");
static_assert(int(::kTfLiteCpuBackendContext) == int(mirror::kTfLiteCpuBackendContext), "");
static_assert(kTfLiteOptionalTensor == -1, "");

int lce_b200_abi_check_ok() { return 1; }
