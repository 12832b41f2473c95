// Probe: does a K-major SWIZZLE_128B tcgen05 operand descriptor work when its start address is shifted by a number of 128-byte
// rows that is NOT a multiple of 8 (i.e. not aligned to the 1024-byte swizzle atom)?  Needed for a 3x3 convolution that reads the
// nine taps as nine shifted windows of ONE halo'd shared-memory tile.  Tries base_offset = 0 and base_offset = (addr >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I promptir_b200/csrc tools/probes/umma_shift_probe.cu \
//        promptir_b200/csrc/build/host.o -o tools/probes/umma_shift_probe -lcudart
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "common.cuh"
#include "host.h"
using namespace pir;

constexpr int ROWS = 256, K = 64, N = 32;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, float* out,
                                              int shift, int mode) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar_full), 1); mbar_init(smem_u32(&bar_mma), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_smem), 32); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t xs = base, ws = base + ROWS * 128;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar_full), ROWS * 128 + N * 128);
    tma_load_2d(xs, &tmX, smem_u32(&bar_full), 0, 0);
    tma_load_2d(ws, &tmW, smem_u32(&bar_full), 0, 0);
  }
  mbar_wait(smem_u32(&bar_full), 0);
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t a_addr = xs + (uint32_t)shift * 128u;
      uint64_t ad = make_sdesc_sw128(a_addr, 16, 1024);
      if (mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;           // matrix base offset
      const uint64_t bd = make_sdesc_sw128(ws, 16, 1024);
      const uint32_t idesc = make_idesc_f16(1, 128, N, 0, 0);
      for (int k = 0; k < 4; ++k) umma_f16(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
      umma_commit(smem_u32(&bar_mma));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar_mma), 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < N; ++j) out[(warp * 32 + lane) * N + j] = __uint_as_float(v[j]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}

int main() {
  std::vector<__nv_bfloat16> hx(ROWS * K), hw(N * K);
  std::vector<float> fx(ROWS * K), fw(N * K);
  srand(1);
  for (int i = 0; i < ROWS * K; ++i) { float f = (rand() % 17 - 8) / 8.0f; hx[i] = __float2bfloat16(f); fx[i] = f; }
  for (int i = 0; i < N * K; ++i) { float f = (rand() % 9 - 4) / 4.0f; hw[i] = __float2bfloat16(f); fw[i] = f; }
  __nv_bfloat16 *dx, *dw; float* dout;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dw, hw.size() * 2); cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmX, tmW;
  { const uint64_t dims[2] = {K, ROWS}; const uint64_t st[1] = {K * 2}; const uint32_t box[2] = {64, ROWS};
    if (pir_make_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, st, box, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap X failed: %s\n", pir_last_error()); return 1; } }
  { const uint64_t dims[2] = {K, N}; const uint64_t st[1] = {K * 2}; const uint32_t box[2] = {64, N};
    if (pir_make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, st, box, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap W failed\n"); return 1; } }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> ho(128 * N);
  const int shifts[] = {0, 8, 1, 2, 3, 5, 9, 18, 19, 37};
  for (int mode = 0; mode < 2; ++mode)
    for (int s : shifts) {
      probe<<<1, 128, 48 * 1024>>>(tmX, tmW, dout, s, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, s, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)fx[(m + s) * K + k] * fw[n * K + k];
          worst = fmax(worst, fabs(ref - ho[m * N + n]));
        }
      printf("base_offset mode %d, shift %2d rows: max |err| = %.4g %s\n", mode, s, worst, worst < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
