/*
 * lce_b200_tflite.h -- the plugin surface: TFLite custom-op registrations for
 * LceQuantize / LceDequantize / LceBconv2d / LceBMaxPool2d backed by the CUDA
 * C-ABI of include/lce_b200.h.
 *
 * What it replaces (LCE = /root/reference/larq_compute_engine):
 *   LCE/tflite/kernels/lce_ops_register.h:16-53   factory functions + registrar
 *   LCE/tflite/kernels/bconv2d.cc:568-599         Register_BCONV_2D*()
 *   LCE/tflite/kernels/quantization.cc:149-159    Register_QUANTIZE/DEQUANTIZE()
 *   LCE/tflite/kernels/bmaxpool.cc:92-96          Register_BMAXPOOL_2D()
 *
 * ABI. The payload is TFLite's plain-C `TfLiteRegistration {init, free, prepare,
 * invoke, ...}` and the structs its callbacks receive (tensorflow/lite/core/c/
 * common.h:107-123, 471-582, 782-1008, 1082-1174 at TF 2.16.1). When building
 * inside a TFLite tree define LCE_B200_USE_TFLITE_HEADERS and the real header is
 * used. Stand-alone (this repository: TFLite cannot be built here) the section
 * below restates the layout of exactly the members this plugin touches; slots it
 * never calls are kept as opaque pointers of the same size. oracle/abi_check.cc
 * static_asserts every offset against the real header in the build container.
 */
#ifndef LCE_B200_TFLITE_H_
#define LCE_B200_TFLITE_H_

#ifdef LCE_B200_USE_TFLITE_HEADERS
#include "tensorflow/lite/core/c/common.h"
#else
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum TfLiteStatus { kTfLiteOk = 0, kTfLiteError = 1 } TfLiteStatus;

/* c_api_types.h:116-137 (only the element types this path reads) */
typedef enum TfLiteType {
  kTfLiteNoType = 0,
  kTfLiteFloat32 = 1,
  kTfLiteInt32 = 2,
  kTfLiteUInt8 = 3,
  kTfLiteInt64 = 4,
  kTfLiteBool = 6,
  kTfLiteInt8 = 9
} TfLiteType;

#define kTfLiteOptionalTensor (-1)

typedef struct TfLiteIntArray {
  int size;
  int data[];
} TfLiteIntArray;

typedef struct TfLiteFloatArray {
  int size;
  float data[];
} TfLiteFloatArray;

typedef struct TfLiteQuantizationParams {
  float scale;
  int32_t zero_point;
} TfLiteQuantizationParams;

typedef enum TfLiteQuantizationType {
  kTfLiteNoQuantization = 0,
  kTfLiteAffineQuantization = 1
} TfLiteQuantizationType;

typedef struct TfLiteQuantization {
  TfLiteQuantizationType type;
  void* params;
} TfLiteQuantization;

typedef struct TfLiteAffineQuantization {
  TfLiteFloatArray* scale;
  TfLiteIntArray* zero_point;
  int32_t quantized_dimension;
} TfLiteAffineQuantization;

typedef union TfLitePtrUnion {
  int32_t* i32;
  float* f;
  char* raw;
  const char* raw_const;
  uint8_t* uint8;
  bool* b;
  int8_t* int8;
  void* data;
} TfLitePtrUnion;

typedef enum TfLiteAllocationType {
  kTfLiteMemNone = 0,
  kTfLiteMmapRo,
  kTfLiteArenaRw,
  kTfLiteArenaRwPersistent,
  kTfLiteDynamic,
  kTfLitePersistentRo,
  kTfLiteCustom,
  kTfLiteVariantObject
} TfLiteAllocationType;

typedef int TfLiteBufferHandle;

typedef struct TfLiteTensor {
  TfLiteType type;
  TfLitePtrUnion data;
  TfLiteIntArray* dims;
  TfLiteQuantizationParams params;
  TfLiteAllocationType allocation_type;
  size_t bytes;
  const void* allocation;
  const char* name;
  void* delegate;
  TfLiteBufferHandle buffer_handle;
  bool data_is_stale;
  bool is_variable;
  TfLiteQuantization quantization;
  void* sparsity;
  const TfLiteIntArray* dims_signature;
} TfLiteTensor;

typedef struct TfLiteNode {
  TfLiteIntArray* inputs;
  TfLiteIntArray* outputs;
  TfLiteIntArray* intermediates;
  TfLiteIntArray* temporaries;
  void* user_data;
  void* builtin_data;
  const void* custom_initial_data;
  int custom_initial_data_size;
  void* delegate;
  bool might_have_side_effect;
} TfLiteNode;

typedef enum TfLiteExternalContextType {
  kTfLiteEigenContext = 0,
  kTfLiteGemmLowpContext = 1,
  kTfLiteEdgeTpuContext = 2,
  kTfLiteCpuBackendContext = 3,
  kTfLiteMaxExternalContexts = 4
} TfLiteExternalContextType;

struct TfLiteContext;
typedef struct TfLiteExternalContext {
  TfLiteExternalContextType type;
  TfLiteStatus (*Refresh)(struct TfLiteContext* context);
} TfLiteExternalContext;

/* Services used by this plugin: ResizeTensor (takes ownership of new_size),
 * ReportError, AddTensors, tensors / GetTensor, recommended_num_threads,
 * Get/SetExternalContext. Everything else is an opaque slot. */
typedef struct TfLiteContext {
  size_t tensors_size;
  void* GetExecutionPlan_;
  TfLiteTensor* tensors;
  void* impl_;
  TfLiteStatus (*ResizeTensor)(struct TfLiteContext*, TfLiteTensor* tensor,
                               TfLiteIntArray* new_size);
  void (*ReportError)(struct TfLiteContext*, const char* msg, ...);
  TfLiteStatus (*AddTensors)(struct TfLiteContext*, int tensors_to_add,
                             int* first_new_tensor_index);
  void* GetNodeAndRegistration_;
  void* ReplaceNodeSubsetsWithDelegateKernels_;
  int recommended_num_threads;
  TfLiteExternalContext* (*GetExternalContext)(struct TfLiteContext*,
                                               TfLiteExternalContextType);
  void (*SetExternalContext)(struct TfLiteContext*, TfLiteExternalContextType,
                             TfLiteExternalContext*);
  bool allow_fp32_relax_to_fp16;
  void* profiler;
  void* AllocatePersistentBuffer_;
  void* AllocateBufferForEval_;
  void* RequestScratchBufferInArena_;
  void* GetScratchBuffer_;
  void* ResizeTensorExplicit_;
  void* PreviewDelegatePartitioning_;
  TfLiteTensor* (*GetTensor)(const struct TfLiteContext* context, int tensor_idx);
  void* GetEvalTensor_;
  void* GetModelMetadata_;
  void* AcquireSubgraphContext_;
  void* ReleaseSubgraphContext_;
} TfLiteContext;

typedef struct TfLiteRegistration {
  void* (*init)(TfLiteContext* context, const char* buffer, size_t length);
  void (*free)(TfLiteContext* context, void* buffer);
  TfLiteStatus (*prepare)(TfLiteContext* context, TfLiteNode* node);
  TfLiteStatus (*invoke)(TfLiteContext* context, TfLiteNode* node);
  const char* (*profiling_string)(const TfLiteContext* context, const TfLiteNode* node);
  int32_t builtin_code;
  const char* custom_name;
  int version;
  void* registration_external;
  void* async_kernel;
  uint64_t inplace_operator;
} TfLiteRegistration;

/* TfLiteIntArrayCreate / Free (common.h:131-150, common.cc): malloc-based. */
TfLiteIntArray* LceB200IntArrayCreate(int size);
void LceB200IntArrayFree(TfLiteIntArray* a);

#ifdef __cplusplus
}
#endif
#endif /* LCE_B200_USE_TFLITE_HEADERS */

#ifdef __cplusplus
extern "C" {
#endif

/* extern "C" factories: each returns a pointer to a static TfLiteRegistration,
 * exactly like the reference's C++ factories (bconv2d.cc:568-590). */
TfLiteRegistration* lce_b200_Register_QUANTIZE(void);
TfLiteRegistration* lce_b200_Register_DEQUANTIZE(void);
TfLiteRegistration* lce_b200_Register_BCONV_2D(void);
TfLiteRegistration* lce_b200_Register_BCONV_2D_REF(void);
TfLiteRegistration* lce_b200_Register_BCONV_2D_OPT_BGEMM(void);
TfLiteRegistration* lce_b200_Register_BCONV_2D_OPT_INDIRECT_BGEMM(void);
TfLiteRegistration* lce_b200_Register_BMAXPOOL_2D(void);

/* The CUDA stream the ops launch on (a cudaStream_t as void*; default: the
 * legacy default stream). The reference threads a CpuBackendContext through
 * context->GetExternalContext (bconv2d.cc:459); a device host sets this once. */
void lce_b200_set_stream(void* stream);
void* lce_b200_get_stream(void);

#ifdef __cplusplus
}

/* C++ names identical to the reference's (lce_ops_register.h:16-21), so code
 * written against `compute_engine::tflite::Register_*` links unchanged. */
namespace compute_engine {
namespace tflite {
TfLiteRegistration* Register_QUANTIZE();
TfLiteRegistration* Register_DEQUANTIZE();
TfLiteRegistration* Register_BCONV_2D();
TfLiteRegistration* Register_BCONV_2D_REF();
TfLiteRegistration* Register_BCONV_2D_OPT_BGEMM();
TfLiteRegistration* Register_BCONV_2D_OPT_INDIRECT_BGEMM();
TfLiteRegistration* Register_BMAXPOOL_2D();

/* RegisterLCECustomOps (lce_ops_register.h:25-53) for any resolver type with an
 * AddCustom(const char*, const TfLiteRegistration*) member -- TFLite's
 * MutableOpResolver or this repository's lce_b200::OpResolver. */
template <class Resolver>
inline void RegisterLCECustomOps(Resolver* resolver, const bool use_reference_bconv = false,
                                 const bool use_indirect_bgemm = false) {
  resolver->AddCustom("LceQuantize", Register_QUANTIZE());
  resolver->AddCustom("LceDequantize", Register_DEQUANTIZE());
  if (use_reference_bconv) {
    resolver->AddCustom("LceBconv2d", Register_BCONV_2D_REF());
  } else if (use_indirect_bgemm) {
    resolver->AddCustom("LceBconv2d", Register_BCONV_2D_OPT_INDIRECT_BGEMM());
  } else {
    resolver->AddCustom("LceBconv2d", Register_BCONV_2D());
  }
  resolver->AddCustom("LceBMaxPool2d", Register_BMAXPOOL_2D());
}
}  // namespace tflite
}  // namespace compute_engine
#endif

#endif /* LCE_B200_TFLITE_H_ */
