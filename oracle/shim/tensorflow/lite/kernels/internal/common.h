// Shim (test infrastructure only): the real kernels/internal/common.h pulls in
// gemmlowp's fixedpoint headers, which are not on disk. The reference's core
// headers only need Offset/MatchingDim/RuntimeShape (types.h) and TfLiteRound
// (cppmath.h) from it.
#ifndef LCE_B200_ORACLE_SHIM_TFLITE_INTERNAL_COMMON_H_
#define LCE_B200_ORACLE_SHIM_TFLITE_INTERNAL_COMMON_H_
#include <algorithm>
#include "tensorflow/lite/kernels/internal/cppmath.h"
#include "tensorflow/lite/kernels/internal/types.h"
#endif
