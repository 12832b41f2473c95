// Internal host-side helpers shared by the translation units of libpromptir_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/promptir_b200.h"

// records a thread-local message and returns `code`
int pir_fail(int code, const char* fmt, ...);
// cudaGetLastError() after a launch -> PIR_OK / PIR_ERR_CUDA (with message)
int pir_check_launch(const char* what);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
// dims/box have `rank` entries, strides_bytes has rank-1 entries (dimension 0 is contiguous).
int pir_make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

// persistent pointwise GEMM (gemm_pw.cu): taps == 1, NHWC16 output
namespace pir { int pir_gemm_pw(const PirGemm* d, cudaStream_t stream); }
