// builtin_params.h -- parameter blob handed to this host's builtin registrations as
// node->builtin_data (filled by tflite_model.cc from the flatbuffer's builtin option
// tables, schema.fbs:806-1106). Private to this host: TFLite's own builtin_op_data.h
// structs are not part of the custom-op ABI.
#ifndef LCE_B200_HOST_BUILTIN_PARAMS_H_
#define LCE_B200_HOST_BUILTIN_PARAMS_H_
#include <cstdint>

namespace lce_b200 {
struct BuiltinParams {
  int32_t builtin_code;
  int32_t padding, stride_h, stride_w, dilation_h, dilation_w;
  int32_t filter_h, filter_w, depth_multiplier;
  int32_t activation;
  int32_t keep_dims;
  int32_t axis;
  float beta;
  int32_t n_new_shape;
  int32_t new_shape[8];
};
}  // namespace lce_b200
#endif
