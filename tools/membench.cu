// HBM micro-benchmarks that bound what the hot-path kernels can reach on this B200: plain LDG/STG streams and the
// TMA load / TMA store patterns the GEMM and stencil kernels use (128-byte rows at a pixel pitch).  Diagnostic only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/membench tools/membench.cu && tools/build/membench
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_read(const uint4* __restrict__ p, size_t n, uint32_t* sink) {
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}
__global__ void k_write(uint4* __restrict__ p, size_t n) {
  const uint4 v = make_uint4(threadIdx.x, 2, 3, 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_copy(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// TMA store stream: every CTA writes boxes {64 x 16-bit, 128 rows} (16 KB) of a [rows][cols] 16-bit matrix; `depth` stores in flight
__global__ void __launch_bounds__(128) k_tma_store(const __grid_constant__ CUtensorMap tm, int col_boxes, int row_boxes, int depth) {
  extern __shared__ __align__(1024) uint8_t sm[];
  for (int i = threadIdx.x; i < depth * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = col_boxes * row_boxes;
    int k = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++k) {
      const int rb = t / col_boxes, cb = t % col_boxes;
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(&tm), "r"(smem_u32(sm) + (k % depth) * 16384), "r"(cb * 64), "r"(rb * 128) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else if (depth == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// TMA load stream: ring of `depth` 16 KB boxes
__global__ void __launch_bounds__(128) k_tma_load(const __grid_constant__ CUtensorMap tm, int col_boxes, int row_boxes, int depth, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = col_boxes * row_boxes;
    int issued = 0, done = 0;
    int t_issue = blockIdx.x;
    while (true) {
      while (issued - done < depth && t_issue < total) {
        const int s = issued % depth;
        const int rb = t_issue / col_boxes, cb = t_issue % col_boxes;
        mbar_expect_tx(smem_u32(&bar[s]), 16384);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sm) + s * 16384), "l"(&tm), "r"(smem_u32(&bar[s])), "r"(cb * 64), "r"(rb * 128) : "memory");
        ++issued; t_issue += gridDim.x;
      }
      if (done == issued) break;
      mbar_wait(smem_u32(&bar[done % depth]), (done / depth) & 1);
      ++done;
    }
    if (sm[5] == 77 && sm[9000] == 3) *sink = 1;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t bytes = (size_t)2 << 30;
  uint8_t *a, *b;
  uint32_t* sink;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](const char* name, double gb, auto fn) {
    fn(); fn();
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0); fn(); cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("%-52s %8.3f ms  %8.1f GB/s\n", name, best, gb / best / 1e6);
    fflush(stdout);
  };
  const size_t n16 = bytes / 16;
  const double GB = (double)bytes;
  for (int bpsm : {4, 8, 16}) {
    char nm[96];
    snprintf(nm, sizeof nm, "LDG.128 read   2 GiB  grid %d x 256", 148 * bpsm);
    time(nm, GB, [&] { k_read<<<148 * bpsm, 256>>>((const uint4*)a, n16, sink); });
    snprintf(nm, sizeof nm, "STG.128 write  2 GiB  grid %d x 256", 148 * bpsm);
    time(nm, GB, [&] { k_write<<<148 * bpsm, 256>>>((uint4*)b, n16); });
    snprintf(nm, sizeof nm, "copy (r+w)     2+2 GiB grid %d x 256", 148 * bpsm);
    time(nm, 2 * GB, [&] { k_copy<<<148 * bpsm, 256>>>((const uint4*)a, (uint4*)b, n16); });
  }
  time("cudaMemcpyAsync D2D 2+2 GiB", 2 * GB, [&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, 0); });
  time("cudaMemsetAsync 2 GiB", GB, [&] { cudaMemsetAsync(b, 3, bytes, 0); });

  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  CK(cudaFuncSetAttribute(k_tma_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
  CK(cudaFuncSetAttribute(k_tma_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
  // matrices of 16-bit elements with `cols` columns (row = one pixel): cols = 64 (dense 128 B rows), 96, 288, 512
  for (int cols : {64, 96, 288, 512}) {
    const uint64_t rows = bytes / 2 / cols / 128 * 128;
    CUtensorMap tm;
    cuuint64_t gd[2] = {(cuuint64_t)cols, rows};
    cuuint64_t gs[1] = {(cuuint64_t)cols * 2};
    cuuint32_t bx[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    for (int store = 0; store < 2; ++store) {
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, store ? b : a, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int col_boxes = (cols + 63) / 64, row_boxes = (int)(rows / 128);
      const double moved = (double)rows * cols * 2;
      for (int depth : {2, 4, 8}) {
        for (int cps : {1, 2}) {
          char nm[96];
          snprintf(nm, sizeof nm, "TMA %s cols %3d depth %d ctas/SM %d", store ? "store" : "load ", cols, depth, cps);
          if (store) time(nm, moved, [&] { k_tma_store<<<148 * cps, 128, depth * 16384>>>(tm, col_boxes, row_boxes, depth); });
          else time(nm, moved, [&] { k_tma_load<<<148 * cps, 128, depth * 16384>>>(tm, col_boxes, row_boxes, depth, sink); });
        }
      }
    }
  }
  printf("done\n");
  return 0;
}
