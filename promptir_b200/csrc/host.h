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
// dense 3x3 over one halo'd tile (conv3x3.cu): 1 = shape not taken (fall back), else the launch status
namespace pir { int conv3x3_try(const PirGemm* d, cudaStream_t stream); }

// Programmatic dependent launch (PDL): a kernel launched with this attribute may start while the previous kernel on the stream is
// still draining; it runs its prologue (barrier init, TMEM allocation, descriptor prefetch, staging of static weights) and then
// executes griddepcontrol.wait (pir::pdl_wait) before touching anything the previous kernel produced.  Only kernels that contain
// that wait are launched through pir_launch.  Measured on B200 (same-box A/B, CUDA-graph replay): 29.03 ms per cfg2 step with the
// attribute, 28.54 ms without (the persistent kernels own the whole shared memory of an SM, so a dependent CTA cannot start early
// anyway, and the early-launch bookkeeping costs a little) -> OFF by default; PIR_PDL=1 turns it on.
bool pir_pdl_enabled();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device, so a process that uses a
// second GPU must set it there too.  Thread safe.  Returns false (with the error recorded) when the runtime refuses.
bool pir_smem_attr_once(const void* kernel, int bytes, const char* what);

template <class... KArgs, class... Args>
inline cudaError_t pir_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pir_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
