// Shared device-side helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers (inline PTX),
// 16-bit packing helpers.  Everything here is original; encodings follow the PTX ISA (tcgen05 shared-memory
// and instruction descriptors).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pir {

// ---------------------------------------------------------------------------------------------------
// element types: the whole pipeline is templated on a 16-bit storage type (bf16 or fp16)
// ---------------------------------------------------------------------------------------------------
struct BF16 { static constexpr int kFmt = 1; static constexpr unsigned short kOne = 0x3f80; };   // kFmt: tcgen05 kind::f16 a/b format code
struct FP16 { static constexpr int kFmt = 0; static constexpr unsigned short kOne = 0x3c00; };   // kOne: 1.0 in the 16-bit type

template <class T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<BF16>(float lo, float hi) {
  uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
// fp16 storage saturates at +-65504 (same instruction cost): an activation beyond the fp16 range clips instead of turning into inf
// and, one LayerNorm later, into NaN
template <> __device__ __forceinline__ uint32_t pack2<FP16>(float lo, float hi) {
  uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
template <class T> __device__ __forceinline__ float unpack_lo(uint32_t v);
template <class T> __device__ __forceinline__ float unpack_hi(uint32_t v);
template <> __device__ __forceinline__ float unpack_lo<BF16>(uint32_t v) { return __uint_as_float(v << 16); }
template <> __device__ __forceinline__ float unpack_hi<BF16>(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
template <> __device__ __forceinline__ float unpack_lo<FP16>(uint32_t v) {
  return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu)));
}
template <> __device__ __forceinline__ float unpack_hi<FP16>(uint32_t v) {
  return __half2float(__ushort_as_half((unsigned short)(v >> 16)));
}
template <class T> __device__ __forceinline__ unsigned short to16(float f);
template <> __device__ __forceinline__ unsigned short to16<BF16>(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
template <> __device__ __forceinline__ unsigned short to16<FP16>(float f) { return __half_as_ushort(__float2half_rn(f)); }
template <class T> __device__ __forceinline__ float from16(unsigned short v);
template <> __device__ __forceinline__ float from16<BF16>(unsigned short v) { return __uint_as_float(((uint32_t)v) << 16); }
template <> __device__ __forceinline__ float from16<FP16>(unsigned short v) { return __half2float(__ushort_as_half(v)); }

// mixed-precision FMA: d = a(16-bit half of a packed reg) * b(16-bit) + c(fp32).  SASS: FHFMA{.BF16}
template <class T> __device__ __forceinline__ float fma16(unsigned short a, unsigned short b, float c);
template <> __device__ __forceinline__ float fma16<BF16>(unsigned short a, unsigned short b, float c) {
  float r; asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(r) : "h"(a), "h"(b), "f"(c)); return r;
}
template <> __device__ __forceinline__ float fma16<FP16>(unsigned short a, unsigned short b, float c) {
  float r; asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(a), "h"(b), "f"(c)); return r;
}
__device__ __forceinline__ unsigned short lo16(uint32_t v) { return (unsigned short)(v & 0xffffu); }
__device__ __forceinline__ unsigned short hi16(uint32_t v) { return (unsigned short)(v >> 16); }

// ---------------------------------------------------------------------------------------------------
// shared-memory addressing + mbarrier
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) { }
}
// try_wait with a suspend-time hint: the thread is parked by the hardware until the phase completes or the hint (in ns) expires, so a
// waiting warp retries rarely instead of spinning through issue slots that its scheduler's other warps could use
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
  } while (!ok);
}
// for producer / issuer warps that wait most of the time: back off so the spin does not take issue slots from the
// compute warps sharing the scheduler
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}

// ---------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16/fp16 in, fp32 accumulate); issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
// 32 lanes x 64 consecutive fp32 columns -> 64 registers per thread
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]),
        "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]),
        "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Instruction descriptor, kind::f16, fp32 accumulate.  fmt: 0=f16 1=bf16.  major: 0=K-major 1=MN-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(int fmt, int M, int N, int a_major, int b_major) {
  return (1u << 4)                      // c_format = F32
         | ((uint32_t)fmt << 7)         // a_format
         | ((uint32_t)fmt << 10)        // b_format
         | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16)
         | ((uint32_t)(N >> 3) << 17)   // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}
// Shared-memory matrix descriptor (sm_100: version=1), 128-byte swizzle.
//   K-major : rows of 128 B (64 x 16-bit along K), 8-row groups SBO bytes apart (1024 when dense)
//   MN-major: rows of 128 B (64 x 16-bit along M/N), 8 K-rows per 1024 B atom, atoms along K SBO apart,
//             64-element groups along M/N LBO apart
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;               // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;               // SWIZZLE_128B
  return d;
}

// programmatic dependent launch: let the next kernel's CTAs start as this grid drains / wait for the previous grid's results
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Warp index through a shuffle from lane 0: the value is the same as threadIdx.x >> 5, but ptxas can now PROVE it warp-uniform, so the
// role branches of the warp-specialised kernels become uniform branches and addresses derived from it (tensor-memory columns,
// mbarriers, the global-memory descriptor) stay in uniform registers.  With the plain shift the fused stencil kernel carried 144
// R2UR moves (two in front of every store, one per tcgen05.ld), 96 registers and spills; with this 3, 87 and none.
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------------
// misc math
// ---------------------------------------------------------------------------------------------------
// erf-GELU  0.5 x (1 + erf(x / sqrt 2))  evaluated as  x * sigmoid(2 x (a + b u + c u^2)),  u = min(x^2, 25):
// a three-coefficient odd polynomial inside the logistic, fitted (minimax) to the exact erf form; |error| <= 2.6e-5
// absolute over all x (tests/test_oracle.py checks this bound), i.e. below half a unit in the last place of the 16-bit
// outputs it feeds.  9 instructions, two of them MUFU (ex2.approx, rcp.approx); -2*log2(e) is folded into the coefficients.
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
constexpr float kGeluA = -2.301121339544986f, kGeluB = -0.10677572399054727f, kGeluC = 0.0010142630610895519f;
__device__ __forceinline__ float gelu_erf(float x) {
  const float u = fminf(x * x, 25.0f);
  const float t = fmaf(u, fmaf(u, kGeluC, kGeluB), kGeluA);
  return x * rcp_approx(1.0f + ex2_approx(x * t));
}

}  // namespace pir
