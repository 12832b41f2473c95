// Error plumbing, device check and TMA descriptor creation for libpromptir_b200.so.
#include "host.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <set>
#include <utility>

static thread_local char g_err[512] = "";

int pir_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int pir_check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return pir_fail(PIR_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return PIR_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int pir_make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return pir_fail(PIR_ERR_DRIVER, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i) gs[i - 1] = strides_bytes[i - 1];
  }
  const CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return pir_fail(PIR_ERR_DRIVER,
                    "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] base %p", (int)r,
                    rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                    (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                    rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
  return PIR_OK;
}

bool pir_smem_attr_once(const void* kernel, int bytes, const char* what) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { pir_fail(PIR_ERR_CUDA, "%s: no current device", what); return false; }
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({kernel, dev})) return true;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
    pir_fail(PIR_ERR_CUDA, "%s: cannot raise the dynamic shared memory limit to %d bytes", what, bytes);
    return false;
  }
  done.insert({kernel, dev});
  return true;
}

bool pir_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("PIR_PDL"); return e && e[0] == '1'; }();
  return on;
}

extern "C" int pir_abi_version(void) { return PIR_ABI_VERSION; }
extern "C" const char* pir_last_error(void) { return g_err; }

extern "C" int pir_check_device(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return pir_fail(PIR_ERR_CUDA, "no CUDA device");
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return pir_fail(PIR_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (p.major != 10) return pir_fail(PIR_ERR_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only", p.major, p.minor);
  return PIR_OK;
}
