// Issue-rate probe for the CUDA-core pipes the fused stencil kernels lean on (run on the device: tools/probes/fma_pipe_probe).
// Each variant is a fully unrolled loop of independent dependency chains, 4 warps per scheduler (512 threads, one CTA per SM);
// reports warp-instructions per cycle per scheduler (SMSP).  Built by `make -C tools/probes`.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define NCH 8

__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint32_t f2fp(float a, float b) { uint32_t r; asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmnmx2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t h[NCH]; float f[NCH]; uint64_t d[NCH]; uint32_t a[NCH];
  for (int i = 0; i < NCH; ++i) { h[i] = seed + i * 0x10001u; f[i] = (float)(seed + i); d[i] = ((uint64_t)(seed + i) << 32) | (seed * 3 + i); a[i] = seed ^ i; }
  const uint32_t hw = 0x3c003c00u ^ (seed & 1); const float fw = 1.0f + (float)(seed & 1); const uint64_t dw = ((uint64_t)__float_as_uint(fw) << 32) | __float_as_uint(fw);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (MODE == 0) { h[i] = hfma2(h[i], hw, h[(i + 1) % NCH]); }                                   // HFMA2 only (3 register sources)
      if (MODE == 1) { f[i] = ffma(f[i], fw, f[(i + 1) % NCH]); }                                     // FFMA only
      if (MODE == 2) { d[i] = ffma2(d[i], dw, d[(i + 1) % NCH]); }                                    // FFMA2 only
      if (MODE == 3) { h[i] = hfma2(h[i], hw, h[(i + 1) % NCH]); f[i] = ffma(f[i], fw, f[(i + 1) % NCH]); }                  // 1 : 1
      if (MODE == 4) { h[i] = hfma2(h[i], hw, h[(i + 1) % NCH]); f[i] = ffma(f[i], fw, f[(i + 1) % NCH]); f[i] = ffma(f[i], fw, f[(i + 3) % NCH]); }   // 1 : 2
      if (MODE == 5) { h[i] = hfma2(h[i], hw, h[(i + 1) % NCH]); a[i] = prmt(a[i], a[(i + 1) % NCH]); }                      // HFMA2 : PRMT 1 : 1
      if (MODE == 6) { h[i] = hfma2(h[i], hw, h[(i + 1) % NCH]); d[i] = ffma2(d[i], dw, d[(i + 1) % NCH]); }                 // HFMA2 : FFMA2 1 : 1
      if (MODE == 7) { a[i] = f2fp(f[i], f[(i + 1) % NCH]); }                                         // F2FP only
      if (MODE == 8) { h[i] = hfma2(h[i], hw, h[(i + 1) % NCH]); h[i] = hfma2(h[i], hw, h[(i + 2) % NCH]); a[i] = f2fp(f[i], f[(i + 1) % NCH]); }   // 2 H : 1 F2FP
      if (MODE == 9) { a[i] = hmnmx2(a[i], a[(i + 1) % NCH]); }                                       // HMNMX2 only
      if (MODE == 10) { a[i] = prmt(a[i], a[(i + 1) % NCH]); }                                        // PRMT only
      if (MODE == 11) { h[i] = hfma2(h[i], hw, h[(i + 1) % NCH]); f[i] = ffma(f[i], fw, f[(i + 1) % NCH]); a[i] = prmt(a[i], a[(i + 1) % NCH]); }   // H : F : PRMT
      if (MODE == 12) { h[i] = hfma2(h[i], hw, h[i]); }                                               // HFMA2, 2 distinct register sources
      if (MODE == 13) { f[i] = ffma(f[i], fw, f[i]); }                                                // FFMA, 2 distinct register sources
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < NCH; ++i) acc ^= h[i] ^ __float_as_uint(f[i]) ^ (uint32_t)d[i] ^ (uint32_t)(d[i] >> 32) ^ a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int per_iter, uint32_t* out, long long* cyc) {
  probe<MODE><<<148, 512>>>(out, cyc, 1u);
  cudaDeviceSynchronize();
  probe<MODE><<<148, 512>>>(out, cyc, 1u);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += (double)h[i];
  avg /= 148;
  const double instr = (double)ITERS * NCH * per_iter * 4;      // warp-instructions per scheduler (4 warps each)
  printf("%-34s %8.0f cycles  %.3f warp-instr/cycle/SMSP\n", name, avg, instr / avg);
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("HFMA2 (3 reg sources)", 1, out, cyc);
  run<12>("HFMA2 (2 reg sources)", 1, out, cyc);
  run<1>("FFMA (3 reg sources)", 1, out, cyc);
  run<13>("FFMA (2 reg sources)", 1, out, cyc);
  run<2>("FFMA2 f32x2", 1, out, cyc);
  run<3>("HFMA2 : FFMA 1:1", 2, out, cyc);
  run<4>("HFMA2 : FFMA 1:2", 3, out, cyc);
  run<5>("HFMA2 : PRMT 1:1", 2, out, cyc);
  run<6>("HFMA2 : FFMA2 1:1", 2, out, cyc);
  run<7>("F2FP.f16x2 only", 1, out, cyc);
  run<8>("HFMA2 : F2FP 2:1", 3, out, cyc);
  run<9>("HMNMX2 only", 1, out, cyc);
  run<10>("PRMT only", 1, out, cyc);
  run<11>("HFMA2 : FFMA : PRMT 1:1:1", 3, out, cyc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
