// Shim (test infrastructure only): the reference's bitpack.h includes
// flatbuffers/base.h solely for this endianness macro (bitpack.h:14,219).
// flatbuffers itself is not vendored in /root/reference.
#ifndef LCE_B200_ORACLE_SHIM_FLATBUFFERS_BASE_H_
#define LCE_B200_ORACLE_SHIM_FLATBUFFERS_BASE_H_
#define FLATBUFFERS_LITTLEENDIAN 1
#endif
