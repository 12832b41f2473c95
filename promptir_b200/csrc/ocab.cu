// OCAB -- overlapping cross-attention of PromptXRestormer for sm_100a.    Reference: net/prompt_xrestormer.py:189-235 (OCAB),
// :25-73 (RelPosEmb: rel_to_abs / relative_logits_1d).
//
// One CTA per (8 x 8 query window, head): 64 queries x 144 keys x 16 channels.  The reference materialises the 12 x 12 key/value
// windows with nn.Unfold (2.25x the k/v tensors) and runs two bmm + ~10 elementwise kernels; here the window is gathered straight
// from the NHWC qkv tensor into shared memory (zero vectors outside the image, exactly what Unfold's zero padding produces --
// such keys still enter the softmax with logit = relative-position bias) and the whole head is finished in registers:
//   thread = (query i, part s in 0..3) owns keys j = s + 4t, t = 0..35, i.e. key row a = t / 3 and key column s + 4 (t % 3)
//   logit  = qs.k_j + qs.rel_w[kc - y + 11] + qs.rel_h[kr - x + 11]           (rel_to_abs reduces to this index shift)
//   softmax over the 4 x 36 logits of the query (two shuffle steps), out = sum p_j v_j reduced the same way.
// The relative-position terms need only 3 + 12 dot products per thread.  HBM traffic is the algorithmic minimum plus the 2.25x key
// halo, which stays in L2.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kOcWs = 8, kOcOws = 12, kOcDh = 16, kOcKeys = kOcOws * kOcOws, kOcRel = 2 * kOcOws - 1;
constexpr int kOcRow = 20;                 // shared-memory row stride in floats: 80 B keeps float4 reads of 4 consecutive rows conflict free

struct OcArgs {
  int H, W, heads, inner;
  const unsigned short* qkv; long long qpitch, qbs;
  const float* rel_h; const float* rel_w;
  unsigned short* out; long long opitch, obs;
};

template <class T>
__device__ __forceinline__ void load16(const unsigned short* p, float (&f)[16]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int q = 0; q < 8; ++q) { f[2 * q] = unpack_lo<T>(w[q]); f[2 * q + 1] = unpack_hi<T>(w[q]); }
}

template <class T>
__global__ void __launch_bounds__(256)
ocab_simt_kernel(const OcArgs a) {
  __shared__ __align__(16) float sK[kOcKeys][kOcRow];
  __shared__ __align__(16) float sV[kOcKeys][kOcRow];
  __shared__ __align__(16) float sRw[kOcRel][kOcDh];
  __shared__ __align__(16) float sRh[kOcRel][kOcDh];
  const int tid = threadIdx.x;
  const int nw = a.W / kOcWs;
  const int wy = blockIdx.x / nw, wx = blockIdx.x % nw;
  const int h = blockIdx.y, b = blockIdx.z;
  const unsigned short* base = a.qkv + (size_t)b * a.qbs;

  // ---- gather the 12 x 12 key / value window of this head (zero outside the image) ----
  for (int e = tid; e < kOcKeys * 4; e += 256) {
    const int j = e >> 2, which = (e >> 1) & 1, half = e & 1;
    const int py = wy * kOcWs - (kOcOws - kOcWs) / 2 + j / kOcOws, px = wx * kOcWs - (kOcOws - kOcWs) / 2 + j % kOcOws;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (py >= 0 && py < a.H && px >= 0 && px < a.W)
      v = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)py * a.W + px) * a.qpitch + (size_t)(1 + which) * a.inner + h * kOcDh + half * 8));
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    float* dst = (which ? sV[j] : sK[j]) + half * 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) { dst[2 * q] = unpack_lo<T>(w4[q]); dst[2 * q + 1] = unpack_hi<T>(w4[q]); }
  }
  for (int e = tid; e < kOcRel * kOcDh; e += 256) {
    sRw[e / kOcDh][e % kOcDh] = __ldg(a.rel_w + e);
    sRh[e / kOcDh][e % kOcDh] = __ldg(a.rel_h + e);
  }
  // ---- this thread's query ----
  const int i = tid >> 2, s = tid & 3;
  const int x = i >> 3, y = i & 7;                      // row / column inside the window
  const size_t qpix = (size_t)(wy * kOcWs + x) * a.W + wx * kOcWs + y;
  float qs[16];
  load16<T>(base + qpix * a.qpitch + h * kOcDh, qs);
#pragma unroll
  for (int d = 0; d < 16; ++d) qs[d] *= 0.25f;           // dim_head^-0.5
  __syncthreads();

  auto dot16 = [&](const float* r) {
    const float4 r0 = *reinterpret_cast<const float4*>(r), r1 = *reinterpret_cast<const float4*>(r + 4);
    const float4 r2 = *reinterpret_cast<const float4*>(r + 8), r3 = *reinterpret_cast<const float4*>(r + 12);
    float t0 = qs[0] * r0.x, t1 = qs[1] * r0.y, t2 = qs[2] * r0.z, t3 = qs[3] * r0.w;
    t0 = fmaf(qs[4], r1.x, t0); t1 = fmaf(qs[5], r1.y, t1); t2 = fmaf(qs[6], r1.z, t2); t3 = fmaf(qs[7], r1.w, t3);
    t0 = fmaf(qs[8], r2.x, t0); t1 = fmaf(qs[9], r2.y, t1); t2 = fmaf(qs[10], r2.z, t2); t3 = fmaf(qs[11], r2.w, t3);
    t0 = fmaf(qs[12], r3.x, t0); t1 = fmaf(qs[13], r3.y, t1); t2 = fmaf(qs[14], r3.z, t2); t3 = fmaf(qs[15], r3.w, t3);
    return (t0 + t1) + (t2 + t3);
  };
  float tw[3], th[12];
#pragma unroll
  for (int bb = 0; bb < 3; ++bb) tw[bb] = dot16(sRw[s + 4 * bb - y + kOcOws - 1]);
#pragma unroll
  for (int aa = 0; aa < 12; ++aa) th[aa] = dot16(sRh[aa - x + kOcOws - 1]);

  float lg[36];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 36; ++t) {
    const int j = s + 4 * t;                            // key row t / 3, key column s + 4 (t % 3)
    lg[t] = dot16(sK[j]) + tw[t % 3] + th[t / 3];
    mx = fmaxf(mx, lg[t]);
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 36; ++t) { lg[t] = exp2f((lg[t] - mx) * 1.4426950408889634f); sum += lg[t]; }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  float o[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll
  for (int t = 0; t < 36; ++t) {
    const float* r = sV[s + 4 * t];
    const float p = lg[t];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 v = *reinterpret_cast<const float4*>(r + 4 * q4);
      o[4 * q4] = fmaf(p, v.x, o[4 * q4]); o[4 * q4 + 1] = fmaf(p, v.y, o[4 * q4 + 1]);
      o[4 * q4 + 2] = fmaf(p, v.z, o[4 * q4 + 2]); o[4 * q4 + 3] = fmaf(p, v.w, o[4 * q4 + 3]);
    }
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    o[d] += __shfl_xor_sync(0xffffffffu, o[d], 1);
    o[d] += __shfl_xor_sync(0xffffffffu, o[d], 2);
    o[d] *= inv;
  }
  // part s stores channels [4s, 4s + 4) of the head (8 bytes): the four parts of a query write one 32-byte run
  float c4[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) c4[d] = s == 0 ? o[d] : (s == 1 ? o[4 + d] : (s == 2 ? o[8 + d] : o[12 + d]));
  uint2 pk;
  pk.x = pack2<T>(c4[0], c4[1]);
  pk.y = pack2<T>(c4[2], c4[3]);
  *reinterpret_cast<uint2*>(a.out + (size_t)b * a.obs + qpix * a.opitch + h * kOcDh + 4 * s) = pk;
}


// ------------------------------------------------------------------------------------------------------
// Tensor-core version (default).  One CTA of 4 warps per (window, head); warp w owns queries 16w .. 16w+15 (window rows 2w, 2w+1).
//   S = Qs K^T      : 18 x mma.m16n8k16 (A = Qs by ldmatrix, B = K rows by ldmatrix), fp32 accumulators stay in registers
//   + relative bias : Tw[i][r] = qs_i . rel_w[r], Th[i][r] = qs_i . rel_h[r] computed once per query in fp32 (19 + 19 dot products,
//                     the parameters stay fp32) and looked up per logit
//   softmax         : on the accumulator fragments, row reductions over the 4 lanes of a quad
//   O = P V         : P re-used in place as the A fragments (16-bit), V by ldmatrix.trans; 18 x mma
// The 64 x 144 x 16 products are far too small for a tcgen05 pipeline (one UMMA per product, TMEM round trip in between);
// warp-level MMA keeps everything in registers.
// ------------------------------------------------------------------------------------------------------
constexpr int kOcRowH = 24;                // 16-bit row stride (48 B): ldmatrix rows hit distinct banks
constexpr int kOcTs = 25;                  // fp32 row stride of the bias tables

template <class T> __device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<BF16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<FP16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <class T>
__global__ void __launch_bounds__(128)
ocab_kernel(const OcArgs a) {
  __shared__ __align__(16) unsigned short sQ[64][kOcRowH];
  __shared__ __align__(16) unsigned short sK[kOcKeys][kOcRowH];
  __shared__ __align__(16) unsigned short sV[kOcKeys][kOcRowH];
  __shared__ __align__(16) float sRel[2][kOcRel][kOcDh];       // [0] rel_w, [1] rel_h
  __shared__ float sT[2][64][kOcTs];                           // [0] Tw, [1] Th
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nw = a.W / kOcWs;
  const int wy = blockIdx.x / nw, wx = blockIdx.x % nw;
  const int h = blockIdx.y, b = blockIdx.z;
  const unsigned short* base = a.qkv + (size_t)b * a.qbs;

  // ---- gather: keys / values of the 12 x 12 window (zero outside the image), the 64 queries (scaled by 1/4: exact), rel tables ----
  for (int e = tid; e < kOcKeys * 4; e += 128) {
    const int j = e >> 2, which = (e >> 1) & 1, half = e & 1;
    const int py = wy * kOcWs - (kOcOws - kOcWs) / 2 + j / kOcOws, px = wx * kOcWs - (kOcOws - kOcWs) / 2 + j % kOcOws;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (py >= 0 && py < a.H && px >= 0 && px < a.W)
      v = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)py * a.W + px) * a.qpitch + (size_t)(1 + which) * a.inner + h * kOcDh + half * 8));
    *reinterpret_cast<uint4*>((which ? sV[j] : sK[j]) + half * 8) = v;
  }
  {
    const int i = tid >> 1, half = tid & 1;                    // 64 queries x 2 halves = 128 threads
    const size_t qpix = (size_t)(wy * kOcWs + (i >> 3)) * a.W + wx * kOcWs + (i & 7);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + qpix * a.qpitch + h * kOcDh + half * 8));
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    uint4 o;
    o.x = pack2<T>(unpack_lo<T>(w4[0]) * 0.25f, unpack_hi<T>(w4[0]) * 0.25f);
    o.y = pack2<T>(unpack_lo<T>(w4[1]) * 0.25f, unpack_hi<T>(w4[1]) * 0.25f);
    o.z = pack2<T>(unpack_lo<T>(w4[2]) * 0.25f, unpack_hi<T>(w4[2]) * 0.25f);
    o.w = pack2<T>(unpack_lo<T>(w4[3]) * 0.25f, unpack_hi<T>(w4[3]) * 0.25f);
    *reinterpret_cast<uint4*>(sQ[i] + half * 8) = o;
  }
  for (int e = tid; e < kOcRel * kOcDh; e += 128) {
    sRel[0][e / kOcDh][e % kOcDh] = __ldg(a.rel_w + e);
    sRel[1][e / kOcDh][e % kOcDh] = __ldg(a.rel_h + e);
  }
  __syncthreads();
  // ---- bias tables: thread (query i, which) -> T[which][i][r], r = 4..22 (the reachable shifts) ----
  {
    const int i = tid >> 1, which = tid & 1;
    float q[16];
#pragma unroll
    for (int d2 = 0; d2 < 8; ++d2) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(&sQ[i][2 * d2]);
      q[2 * d2] = unpack_lo<T>(w);
      q[2 * d2 + 1] = unpack_hi<T>(w);
    }
#pragma unroll
    for (int r = 4; r < kOcRel; ++r) {
      const float* rr = sRel[which][r];
      float t0 = 0.f, t1 = 0.f;
#pragma unroll
      for (int d = 0; d < 16; d += 2) { t0 = fmaf(q[d], rr[d], t0); t1 = fmaf(q[d + 1], rr[d + 1], t1); }
      sT[which][i][r] = t0 + t1;
    }
  }
  __syncthreads();

  // ---- S = Qs K^T ----
  const int g = lane >> 2, q4 = lane & 3;
  uint32_t af[4];
  ldsm_x4(af, smem_u32(&sQ[warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)][8 * (lane >> 4)]));
  float c[18][4];
#pragma unroll
  for (int n = 0; n < 18; ++n) { c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.f; }
#pragma unroll
  for (int n2 = 0; n2 < 9; ++n2) {                             // 16 keys per ldmatrix.x4: two n8 tiles
    uint32_t bf[4];
    ldsm_x4(bf, smem_u32(&sK[n2 * 16 + (lane & 7) + 8 * (lane >> 4)][8 * ((lane >> 3) & 1)]));
    mma16816<T>(c[2 * n2], af, bf[0], bf[1]);
    mma16816<T>(c[2 * n2 + 1], af, bf[2], bf[3]);
  }
  // ---- + relative-position bias, row max.  This thread: rows i0 = 16 warp + g (window row 2 warp, column g) and i0 + 8 ----
  const int i0 = warp * 16 + g;
  const float* tw0 = sT[0][i0] + (kOcOws - 1 - g);             // + kc
  const float* tw1 = sT[0][i0 + 8] + (kOcOws - 1 - g);
  const float* th0 = sT[1][i0] + (kOcOws - 1 - 2 * warp);      // + kr
  const float* th1 = sT[1][i0 + 8] + (kOcOws - 1 - (2 * warp + 1));
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < 18; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int v = (8 * n) % 12 + 2 * q4 + e;                 // key j = 8 n + 2 q4 + e -> (kr, kc) = (j / 12, j % 12)
      const int wrap = v >= 12 ? 1 : 0;
      const int kc = v - 12 * wrap, kr = (8 * n) / 12 + wrap;
      c[n][e] += tw0[kc] + th0[kr];
      c[n][2 + e] += tw1[kc] + th1[kr];
      m0 = fmaxf(m0, c[n][e]);
      m1 = fmaxf(m1, c[n][2 + e]);
    }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float s0 = 0.f, s1 = 0.f;
  const float L2E = 1.4426950408889634f;
#pragma unroll
  for (int n = 0; n < 18; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      c[n][e] = exp2f((c[n][e] - m0) * L2E); s0 += c[n][e];
      c[n][2 + e] = exp2f((c[n][2 + e] - m1) * L2E); s1 += c[n][2 + e];
    }
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  // ---- O = P V ----
  float o[2][4];
#pragma unroll
  for (int t = 0; t < 2; ++t) { o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 9; ++kk) {
    uint32_t pa[4];
    pa[0] = pack2<T>(c[2 * kk][0], c[2 * kk][1]);
    pa[1] = pack2<T>(c[2 * kk][2], c[2 * kk][3]);
    pa[2] = pack2<T>(c[2 * kk + 1][0], c[2 * kk + 1][1]);
    pa[3] = pack2<T>(c[2 * kk + 1][2], c[2 * kk + 1][3]);
    uint32_t vf[4];
    ldsm_x4_trans(vf, smem_u32(&sV[kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)][8 * (lane >> 4)]));
    mma16816<T>(o[0], pa, vf[0], vf[1]);
    mma16816<T>(o[1], pa, vf[2], vf[3]);
  }
  const float r0 = 1.0f / s0, r1 = 1.0f / s1;
  const size_t pix0 = (size_t)(wy * kOcWs + 2 * warp) * a.W + wx * kOcWs + g;
  unsigned short* op = a.out + (size_t)b * a.obs + h * kOcDh + 2 * q4;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    *reinterpret_cast<uint32_t*>(op + pix0 * a.opitch + 8 * t) = pack2<T>(o[t][0] * r0, o[t][1] * r0);
    *reinterpret_cast<uint32_t*>(op + (pix0 + a.W) * a.opitch + 8 * t) = pack2<T>(o[t][2] * r1, o[t][3] * r1);
  }
}


// ------------------------------------------------------------------------------------------------------
// Backward.  One CTA of 256 threads per (window, head); thread = (query i, part s) as in the SIMT forward.
//   recompute p = softmax(logits);  dP = dO V^T;  dS = p (dP - <p, dP>);
//   dQs_i = sum_j dS_ij (k_j + rel_w[kc - y + 11] + rel_h[kr - x + 11]),  dq = dQs / 4                  -> global (queries do not overlap)
//   dK_j = sum_i dS_ij qs_i,  dV_j = sum_i p_ij dO_i     (dS, p staged in shared memory, one thread per (key, half))
//        -> per-window partial in the workspace: overlapping windows share key pixels, a gather kernel adds the <= 4 contributions
//   d rel_w[r] = sum_i sum_{kc - y_i + 11 = r} (sum_kr dS_i[kr, kc]) qs_i   (and rel_h alike) -> per-CTA partial, reduced in fixed order
// fp32 throughout; deterministic (no atomics).
// ------------------------------------------------------------------------------------------------------
constexpr int kObKV = kOcKeys * 32;          // floats of dK|dV per (window, head): [144][16 dk | 16 dv]
constexpr int kObRel = 2 * kOcRel * kOcDh;   // floats of d rel_w | d rel_h per (window, head)

struct ObArgs {
  int H, W, heads, inner;
  const unsigned short* qkv; long long qpitch, qbs;
  const unsigned short* dout; long long dpitch, dbs;
  const float* rel_h; const float* rel_w;
  unsigned short* dqkv; long long gpitch, gbs;
  float* ws_kv;            // [B][windows][heads][144][32]
  float* ws_rel;           // [B][windows][heads][2][23][16]
};

template <class T>
__global__ void __launch_bounds__(256)
ocab_bwd_kernel(const ObArgs a) {
  extern __shared__ __align__(16) float ob_smem[];
  float (*sK)[kOcRow] = reinterpret_cast<float (*)[kOcRow]>(ob_smem);                      // [144][20]
  float (*sV)[kOcRow] = reinterpret_cast<float (*)[kOcRow]>(ob_smem + kOcKeys * kOcRow);
  float (*sRw)[kOcDh] = reinterpret_cast<float (*)[kOcDh]>(ob_smem + 2 * kOcKeys * kOcRow);
  float (*sRh)[kOcDh] = sRw + kOcRel;
  float (*sQ)[kOcDh] = sRh + kOcRel;                                                       // [64][16] scaled queries
  float (*sdO)[kOcDh] = sQ + 64;                                                           // [64][16]
  float (*sP)[kOcKeys + 1] = reinterpret_cast<float (*)[kOcKeys + 1]>(&sdO[64][0]);        // [64][145]
  float (*sdS)[kOcKeys + 1] = sP + 64;                                                     // [64][145]
  float (*sC)[kOcOws] = reinterpret_cast<float (*)[kOcOws]>(&sdS[64][0]);                  // [64][12] column sums of dS (per kc)
  float (*sR)[kOcOws] = sC + 64;                                                           // [64][12] row sums of dS (per kr)
  float (*sTw)[kOcOws] = sR + 64;                                                          // [64][12] qs . rel_w[kc - y + 11]
  float (*sTh)[kOcOws] = sTw + 64;                                                         // [64][12] qs . rel_h[kr - x + 11]
  const int tid = threadIdx.x;
  const int nw = a.W / kOcWs, nwin = (a.H / kOcWs) * nw;
  const int wy = blockIdx.x / nw, wx = blockIdx.x % nw;
  const int h = blockIdx.y, b = blockIdx.z;
  const unsigned short* base = a.qkv + (size_t)b * a.qbs;

  for (int e = tid; e < kOcKeys * 4; e += 256) {
    const int j = e >> 2, which = (e >> 1) & 1, half = e & 1;
    const int py = wy * kOcWs - (kOcOws - kOcWs) / 2 + j / kOcOws, px = wx * kOcWs - (kOcOws - kOcWs) / 2 + j % kOcOws;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (py >= 0 && py < a.H && px >= 0 && px < a.W)
      v = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)py * a.W + px) * a.qpitch + (size_t)(1 + which) * a.inner + h * kOcDh + half * 8));
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    float* dst = (which ? sV[j] : sK[j]) + half * 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) { dst[2 * q] = unpack_lo<T>(w4[q]); dst[2 * q + 1] = unpack_hi<T>(w4[q]); }
  }
  for (int e = tid; e < kOcRel * kOcDh; e += 256) {
    sRw[e / kOcDh][e % kOcDh] = __ldg(a.rel_w + e);
    sRh[e / kOcDh][e % kOcDh] = __ldg(a.rel_h + e);
  }
  const int i = tid >> 2, s = tid & 3;
  const int x = i >> 3, y = i & 7;
  const size_t qpix = (size_t)(wy * kOcWs + x) * a.W + wx * kOcWs + y;
  float qs[16], dO[16];
  load16<T>(base + qpix * a.qpitch + h * kOcDh, qs);
  load16<T>(a.dout + (size_t)b * a.dbs + qpix * a.dpitch + h * kOcDh, dO);
#pragma unroll
  for (int d = 0; d < 16; ++d) qs[d] *= 0.25f;
  if (s == 0) {
#pragma unroll
    for (int d = 0; d < 16; ++d) { sQ[i][d] = qs[d]; sdO[i][d] = dO[d]; }
  }
  __syncthreads();

  auto dot16 = [](const float (&u)[16], const float* r) {
    const float4 r0 = *reinterpret_cast<const float4*>(r), r1 = *reinterpret_cast<const float4*>(r + 4);
    const float4 r2 = *reinterpret_cast<const float4*>(r + 8), r3 = *reinterpret_cast<const float4*>(r + 12);
    float t0 = u[0] * r0.x, t1 = u[1] * r0.y, t2 = u[2] * r0.z, t3 = u[3] * r0.w;
    t0 = fmaf(u[4], r1.x, t0); t1 = fmaf(u[5], r1.y, t1); t2 = fmaf(u[6], r1.z, t2); t3 = fmaf(u[7], r1.w, t3);
    t0 = fmaf(u[8], r2.x, t0); t1 = fmaf(u[9], r2.y, t1); t2 = fmaf(u[10], r2.z, t2); t3 = fmaf(u[11], r2.w, t3);
    t0 = fmaf(u[12], r3.x, t0); t1 = fmaf(u[13], r3.y, t1); t2 = fmaf(u[14], r3.z, t2); t3 = fmaf(u[15], r3.w, t3);
    return (t0 + t1) + (t2 + t3);
  };
  // ---- bias tables of this query: part s computes the 3 key columns kc = s + 4 b and the 3 key rows a = s, s + 4, s + 8 ----
#pragma unroll
  for (int bb = 0; bb < 3; ++bb) {
    sTw[i][s + 4 * bb] = dot16(qs, sRw[s + 4 * bb - y + kOcOws - 1]);
    sTh[i][s + 4 * bb] = dot16(qs, sRh[s + 4 * bb - x + kOcOws - 1]);
    sC[i][s + 4 * bb] = 0.f;
  }
  __syncwarp();                                         // the four parts of a query share a warp
  // ---- recompute the softmax; logits / probabilities live in this thread's own slots of sP (runtime loops: registers stay small) ----
  float mx = -INFINITY;
#pragma unroll 2
  for (int t = 0; t < 36; ++t) {
    const int ka = t / 3, j = s + 4 * t;
    const float l = dot16(qs, sK[j]) + sTw[i][j - ka * kOcOws] + sTh[i][ka];
    sP[i][j] = l;
    mx = fmaxf(mx, l);
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float sum = 0.f;
#pragma unroll 4
  for (int t = 0; t < 36; ++t) { const float e = exp2f((sP[i][s + 4 * t] - mx) * 1.4426950408889634f); sP[i][s + 4 * t] = e; sum += e; }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  const float inv = 1.0f / sum;
  // ---- dP = dO . v_j (parked in sdS), <p, dP> ----
  float rowdot = 0.f;
#pragma unroll 2
  for (int t = 0; t < 36; ++t) {
    const int j = s + 4 * t;
    const float pj = sP[i][j] * inv;
    const float dPt = dot16(dO, sV[j]);
    sP[i][j] = pj;
    sdS[i][j] = dPt;
    rowdot = fmaf(pj, dPt, rowdot);
  }
  rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 1);
  rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 2);
  // ---- dS, dQs += dS k_j, column sums of dS per key column ----
  float dq[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) dq[d] = 0.f;
#pragma unroll 2
  for (int t = 0; t < 36; ++t) {
    const int ka = t / 3, j = s + 4 * t;
    const float dS = sP[i][j] * (sdS[i][j] - rowdot);
    sdS[i][j] = dS;
    sC[i][j - ka * kOcOws] += dS;                       // key column kc = j - 12 ka is owned by this part alone
    const float* kr = sK[j];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 kv = *reinterpret_cast<const float4*>(kr + 4 * q4);
      dq[4 * q4] = fmaf(dS, kv.x, dq[4 * q4]); dq[4 * q4 + 1] = fmaf(dS, kv.y, dq[4 * q4 + 1]);
      dq[4 * q4 + 2] = fmaf(dS, kv.z, dq[4 * q4 + 2]); dq[4 * q4 + 3] = fmaf(dS, kv.w, dq[4 * q4 + 3]);
    }
  }
  __syncwarp();
  // relative-position part of dQs: this part's 3 key columns and 3 key rows (row sums over all 12 columns of the query)
#pragma unroll
  for (int bb = 0; bb < 3; ++bb) {
    const int kk = s + 4 * bb;
    const float cbv = sC[i][kk];
    float rav = 0.f;
#pragma unroll
    for (int kc = 0; kc < kOcOws; ++kc) rav += sdS[i][kk * kOcOws + kc];
    sR[i][kk] = rav;
    const float* rw = sRw[kk - y + kOcOws - 1];
    const float* rh = sRh[kk - x + kOcOws - 1];
#pragma unroll
    for (int d = 0; d < 16; ++d) dq[d] = fmaf(cbv, rw[d], fmaf(rav, rh[d], dq[d]));
  }
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    dq[d] += __shfl_xor_sync(0xffffffffu, dq[d], 1);
    dq[d] += __shfl_xor_sync(0xffffffffu, dq[d], 2);
    dq[d] *= 0.25f;                                      // qs = q / 4
  }
  {
    float c4[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) c4[d] = s == 0 ? dq[d] : (s == 1 ? dq[4 + d] : (s == 2 ? dq[8 + d] : dq[12 + d]));
    uint2 pk;
    pk.x = pack2<T>(c4[0], c4[1]);
    pk.y = pack2<T>(c4[2], c4[3]);
    *reinterpret_cast<uint2*>(a.dqkv + (size_t)b * a.gbs + qpix * a.gpitch + h * kOcDh + 4 * s) = pk;
  }
  __syncthreads();
  // ---- dK_j = sum_i dS_ij qs_i,  dV_j = sum_i p_ij dO_i: task = (key j, dk | dv, half of the 16 channels) ----
  const size_t cta = ((size_t)b * nwin + blockIdx.x) * a.heads + h;
  float* okv = a.ws_kv + cta * kObKV;
  for (int task = tid; task < kOcKeys * 4; task += 256) {
    const int j = task % kOcKeys, sel = task / kOcKeys;   // sel: 0 dk lo, 1 dk hi, 2 dv lo, 3 dv hi
    const float (*coef)[kOcKeys + 1] = sel < 2 ? sdS : sP;
    const float (*vec)[kOcDh] = sel < 2 ? sQ : sdO;
    const int d0 = (sel & 1) * 8;
    float acc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = 0.f;
    for (int ii = 0; ii < 64; ++ii) {
      const float cf = coef[ii][j];
      const float4 v0 = *reinterpret_cast<const float4*>(&vec[ii][d0]), v1 = *reinterpret_cast<const float4*>(&vec[ii][d0 + 4]);
      acc[0] = fmaf(cf, v0.x, acc[0]); acc[1] = fmaf(cf, v0.y, acc[1]); acc[2] = fmaf(cf, v0.z, acc[2]); acc[3] = fmaf(cf, v0.w, acc[3]);
      acc[4] = fmaf(cf, v1.x, acc[4]); acc[5] = fmaf(cf, v1.y, acc[5]); acc[6] = fmaf(cf, v1.z, acc[6]); acc[7] = fmaf(cf, v1.w, acc[7]);
    }
    float* o = okv + (size_t)j * 32 + (sel >> 1) * 16 + d0;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  // ---- table gradients of this (window, head): task = (table, r, d) ----
  float* orel = a.ws_rel + cta * kObRel;
  for (int task = tid; task < kObRel; task += 256) {
    const int d = task % kOcDh, r = (task / kOcDh) % kOcRel, tbl = task / (kOcDh * kOcRel);   // tbl 0: rel_w (columns), 1: rel_h (rows)
    float acc = 0.f;
    for (int ii = 0; ii < 64; ++ii) {
      const int pos = tbl == 0 ? (ii & 7) : (ii >> 3);    // query column y / row x
      const int kk = r + pos - (kOcOws - 1);              // key column / row with kk - pos + 11 == r
      if (kk >= 0 && kk < kOcOws) acc = fmaf(tbl == 0 ? sC[ii][kk] : sR[ii][kk], sQ[ii][d], acc);
    }
    orel[task] = acc;
  }
}

// ------------------------------------------------------------------------------------------------------
// Backward on warp-level tensor-core MMA (default; the fp32 SIMT kernel above stays selectable with PIR_OCAB_BWD_SIMT=1).
// One CTA of 4 warps per (window, head); warp w owns queries 16w .. 16w+15 like the forward kernel.
//   phase A (per warp, accumulators in registers):
//       S = Qs K^T + bias -> P (18 mma);  dP = dO V^T (18 mma);  dS = P (dP - <P, dP>);  dQs = dS K (18 mma)
//       P and dS are staged as 16-bit [query][key] tiles; dS carries a power-of-two scale in the fp16 build (its values sit near
//       the fp16 subnormal range otherwise).  Column / row sums of dS (C_i[kc], R_i[kr]) and the relative-position part of dQs
//       are SIMT work on the warp's own 16 queries (a query's 144 keys all belong to one warp: no block barrier).
//   phase B (all warps): dK = dS^T Qs and dV = P^T dO, 9 key tiles each, A operand by ldmatrix.trans from the staged tiles (72 + 72 mma);
//       d rel_w = Cs^T Qs, d rel_h = Rs^T Qs with the shifted tables Cs[i][r] = C_i[r + y_i - 11]: one (table, 16-row tile) per warp.
// Workspace layout and the gather / table-reduce kernels are those of the SIMT version; fp32 accumulation, deterministic.
// ------------------------------------------------------------------------------------------------------
constexpr int kObPs = 152;                 // 16-bit row stride of the staged P / dS tiles (304 B: ldmatrix rows on distinct banks)
constexpr int kObCs = 24;                  // 16-bit row stride of the shifted column / row sum tables (23 used)
// 74.6 KB: three CTAs per SM.  The per-warp scratch (column / row sums, shifted tables) lives in the warp's own rows of the bias tables,
// which are dead once the warp has added the bias to its logits.
constexpr size_t kObMmaSmem = 2 * (2 * 64 * kOcRowH + 2 * kOcKeys * kOcRowH) + 4 * (2 * kOcRel * kOcDh + 2 * 64 * kOcTs) + 2 * (2 * 64 * kObPs);
static_assert(2 * 16 * kOcOws * 4 <= 16 * kOcTs * 4 && 2 * 16 * kObCs * 2 <= 16 * kOcTs * 4, "per-warp scratch must fit the warp's bias-table rows");

template <class T>
__global__ void __launch_bounds__(128, 3)
ocab_bwd_mma_kernel(const ObArgs a) {
  typedef unsigned short u16;
  extern __shared__ __align__(16) uint8_t obm[];
  u16 (*sQ)[kOcRowH] = reinterpret_cast<u16 (*)[kOcRowH]>(obm);                              // [64] scaled queries
  u16 (*sdO)[kOcRowH] = sQ + 64;                                                             // [64]
  u16 (*sK)[kOcRowH] = sdO + 64;                                                             // [144]
  u16 (*sV)[kOcRowH] = sK + kOcKeys;                                                         // [144]
  float (*sRel)[kOcRel][kOcDh] = reinterpret_cast<float (*)[kOcRel][kOcDh]>(sV + kOcKeys);   // [2]: rel_w, rel_h
  float (*sT)[64][kOcTs] = reinterpret_cast<float (*)[64][kOcTs]>(&sRel[2][0][0]);           // [2]: Tw, Th
  u16 (*sP)[kObPs] = reinterpret_cast<u16 (*)[kObPs]>(&sT[2][0][0]);                         // [64][144]
  u16 (*sdS)[kObPs] = sP + 64;                                                               // [64][144], scaled
  // per-warp scratch inside the bias tables (see kObMmaSmem): warp w' keeps C | R of its 16 queries ([16][12] fp32 each) in its rows of
  // Tw and the shifted 16-bit tables Cs | Rs ([2][16][24]) in its rows of Th
  auto wC = [&](int w2) { return &sT[0][w2 * 16][0]; };
  auto wCs = [&](int w2) { return reinterpret_cast<u16*>(&sT[1][w2 * 16][0]); };
  constexpr float kS = T::kFmt == 0 ? 256.0f : 1.0f, kInvS = 1.0f / kS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nw = a.W / kOcWs, nwin = (a.H / kOcWs) * nw;
  const int wy = blockIdx.x / nw, wx = blockIdx.x % nw;
  const int h = blockIdx.y, b = blockIdx.z;
  const unsigned short* base = a.qkv + (size_t)b * a.qbs;

  // ---- gather (as the forward kernel) + the 64 output gradients ----
  for (int e = tid; e < kOcKeys * 4; e += 128) {
    const int j = e >> 2, which = (e >> 1) & 1, half = e & 1;
    const int py = wy * kOcWs - (kOcOws - kOcWs) / 2 + j / kOcOws, px = wx * kOcWs - (kOcOws - kOcWs) / 2 + j % kOcOws;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (py >= 0 && py < a.H && px >= 0 && px < a.W)
      v = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)py * a.W + px) * a.qpitch + (size_t)(1 + which) * a.inner + h * kOcDh + half * 8));
    *reinterpret_cast<uint4*>((which ? sV[j] : sK[j]) + half * 8) = v;
  }
  {
    const int i = tid >> 1, half = tid & 1;
    const size_t qpix = (size_t)(wy * kOcWs + (i >> 3)) * a.W + wx * kOcWs + (i & 7);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + qpix * a.qpitch + h * kOcDh + half * 8));
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    uint4 o;
    o.x = pack2<T>(unpack_lo<T>(w4[0]) * 0.25f, unpack_hi<T>(w4[0]) * 0.25f);
    o.y = pack2<T>(unpack_lo<T>(w4[1]) * 0.25f, unpack_hi<T>(w4[1]) * 0.25f);
    o.z = pack2<T>(unpack_lo<T>(w4[2]) * 0.25f, unpack_hi<T>(w4[2]) * 0.25f);
    o.w = pack2<T>(unpack_lo<T>(w4[3]) * 0.25f, unpack_hi<T>(w4[3]) * 0.25f);
    *reinterpret_cast<uint4*>(sQ[i] + half * 8) = o;
    *reinterpret_cast<uint4*>(sdO[i] + half * 8) =
        __ldg(reinterpret_cast<const uint4*>(a.dout + (size_t)b * a.dbs + qpix * a.dpitch + h * kOcDh + half * 8));
  }
  for (int e = tid; e < kOcRel * kOcDh; e += 128) {
    sRel[0][e / kOcDh][e % kOcDh] = __ldg(a.rel_w + e);
    sRel[1][e / kOcDh][e % kOcDh] = __ldg(a.rel_h + e);
  }
  __syncthreads();
  {
    const int i = tid >> 1, which = tid & 1;
    float q[16];
#pragma unroll
    for (int d2 = 0; d2 < 8; ++d2) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(&sQ[i][2 * d2]);
      q[2 * d2] = unpack_lo<T>(w);
      q[2 * d2 + 1] = unpack_hi<T>(w);
    }
#pragma unroll
    for (int r = 4; r < kOcRel; ++r) {
      const float* rr = sRel[which][r];
      float t0 = 0.f, t1 = 0.f;
#pragma unroll
      for (int d = 0; d < 16; d += 2) { t0 = fmaf(q[d], rr[d], t0); t1 = fmaf(q[d + 1], rr[d + 1], t1); }
      sT[which][i][r] = t0 + t1;
    }
  }
  __syncthreads();

  // =============================== phase A: this warp's 16 queries ===============================
  const int g = lane >> 2, q4 = lane & 3;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  float c[18][4];
  {
    uint32_t af[4];
    ldsm_x4(af, smem_u32(&sQ[warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)][8 * (lane >> 4)]));
#pragma unroll
    for (int n = 0; n < 18; ++n) { c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.f; }
#pragma unroll
    for (int n2 = 0; n2 < 9; ++n2) {
      uint32_t bf[4];
      ldsm_x4(bf, smem_u32(&sK[n2 * 16 + (lane & 7) + 8 * (lane >> 4)][8 * ((lane >> 3) & 1)]));
      mma16816<T>(c[2 * n2], af, bf[0], bf[1]);
      mma16816<T>(c[2 * n2 + 1], af, bf[2], bf[3]);
    }
  }
  {
    const float* tw0 = sT[0][row0] + (kOcOws - 1 - g);
    const float* tw1 = sT[0][row1] + (kOcOws - 1 - g);
    const float* th0 = sT[1][row0] + (kOcOws - 1 - 2 * warp);
    const float* th1 = sT[1][row1] + (kOcOws - 1 - (2 * warp + 1));
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 18; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int v = (8 * n) % 12 + 2 * q4 + e;
        const int wrap = v >= 12 ? 1 : 0;
        const int kc = v - 12 * wrap, kr = (8 * n) / 12 + wrap;
        c[n][e] += tw0[kc] + th0[kr];
        c[n][2 + e] += tw1[kc] + th1[kr];
        m0 = fmaxf(m0, c[n][e]);
        m1 = fmaxf(m1, c[n][2 + e]);
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
    const float L2E = 1.4426950408889634f;
#pragma unroll
    for (int n = 0; n < 18; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        c[n][e] = exp2f((c[n][e] - m0) * L2E); s0 += c[n][e];
        c[n][2 + e] = exp2f((c[n][2 + e] - m1) * L2E); s1 += c[n][2 + e];
      }
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float r0 = 1.0f / s0, r1 = 1.0f / s1;
#pragma unroll
    for (int n = 0; n < 18; ++n) { c[n][0] *= r0; c[n][1] *= r0; c[n][2] *= r1; c[n][3] *= r1; }
  }
  // dP = dO V^T, <P, dP>, dS
  float dp[18][4];
  {
    uint32_t df[4];
    ldsm_x4(df, smem_u32(&sdO[warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)][8 * (lane >> 4)]));
#pragma unroll
    for (int n = 0; n < 18; ++n) { dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f; }
#pragma unroll
    for (int n2 = 0; n2 < 9; ++n2) {
      uint32_t bf[4];
      ldsm_x4(bf, smem_u32(&sV[n2 * 16 + (lane & 7) + 8 * (lane >> 4)][8 * ((lane >> 3) & 1)]));
      mma16816<T>(dp[2 * n2], df, bf[0], bf[1]);
      mma16816<T>(dp[2 * n2 + 1], df, bf[2], bf[3]);
    }
    float rd0 = 0.f, rd1 = 0.f;
#pragma unroll
    for (int n = 0; n < 18; ++n) {
      rd0 = fmaf(c[n][0], dp[n][0], fmaf(c[n][1], dp[n][1], rd0));
      rd1 = fmaf(c[n][2], dp[n][2], fmaf(c[n][3], dp[n][3], rd1));
    }
    rd0 += __shfl_xor_sync(0xffffffffu, rd0, 1); rd0 += __shfl_xor_sync(0xffffffffu, rd0, 2);
    rd1 += __shfl_xor_sync(0xffffffffu, rd1, 1); rd1 += __shfl_xor_sync(0xffffffffu, rd1, 2);
#pragma unroll
    for (int n = 0; n < 18; ++n) {
      dp[n][0] = c[n][0] * (dp[n][0] - rd0) * kS; dp[n][1] = c[n][1] * (dp[n][1] - rd0) * kS;
      dp[n][2] = c[n][2] * (dp[n][2] - rd1) * kS; dp[n][3] = c[n][3] * (dp[n][3] - rd1) * kS;
    }
  }
  // stage P and dS (16-bit), dQs = dS K on the tensor cores
  float dq[2][4];
#pragma unroll
  for (int t = 0; t < 2; ++t) { dq[t][0] = dq[t][1] = dq[t][2] = dq[t][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 9; ++kk) {
    uint32_t pa[4];
    pa[0] = pack2<T>(dp[2 * kk][0], dp[2 * kk][1]);
    pa[1] = pack2<T>(dp[2 * kk][2], dp[2 * kk][3]);
    pa[2] = pack2<T>(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
    pa[3] = pack2<T>(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
    *reinterpret_cast<uint32_t*>(&sdS[row0][16 * kk + 2 * q4]) = pa[0];
    *reinterpret_cast<uint32_t*>(&sdS[row1][16 * kk + 2 * q4]) = pa[1];
    *reinterpret_cast<uint32_t*>(&sdS[row0][16 * kk + 8 + 2 * q4]) = pa[2];
    *reinterpret_cast<uint32_t*>(&sdS[row1][16 * kk + 8 + 2 * q4]) = pa[3];
    *reinterpret_cast<uint32_t*>(&sP[row0][16 * kk + 2 * q4]) = pack2<T>(c[2 * kk][0], c[2 * kk][1]);
    *reinterpret_cast<uint32_t*>(&sP[row1][16 * kk + 2 * q4]) = pack2<T>(c[2 * kk][2], c[2 * kk][3]);
    *reinterpret_cast<uint32_t*>(&sP[row0][16 * kk + 8 + 2 * q4]) = pack2<T>(c[2 * kk + 1][0], c[2 * kk + 1][1]);
    *reinterpret_cast<uint32_t*>(&sP[row1][16 * kk + 8 + 2 * q4]) = pack2<T>(c[2 * kk + 1][2], c[2 * kk + 1][3]);
    uint32_t kf[4];
    ldsm_x4_trans(kf, smem_u32(&sK[kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)][8 * (lane >> 4)]));
    mma16816<T>(dq[0], pa, kf[0], kf[1]);
    mma16816<T>(dq[1], pa, kf[2], kf[3]);
  }
  __syncwarp();                                        // the warp's dS rows are complete; its bias-table rows are dead
  {
    // column sums C_i[kc] (even lanes) / row sums R_i[kr] (odd lanes) of dS for query i = 16 warp + lane / 2
    const int il = lane >> 1, i = warp * 16 + il, which = lane & 1;
    const int x = i >> 3, y = i & 7;
    const u16* ds = sdS[i];
    float* mine = wC(warp) + which * (16 * kOcOws) + il * kOcOws;          // C rows, then R rows
    if (which == 0) {
#pragma unroll
      for (int kc2 = 0; kc2 < kOcOws; kc2 += 2) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int kr = 0; kr < kOcOws; ++kr) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(ds + kr * kOcOws + kc2);
          s0 += unpack_lo<T>(w); s1 += unpack_hi<T>(w);
        }
        mine[kc2] = s0 * kInvS; mine[kc2 + 1] = s1 * kInvS;
      }
    } else {
#pragma unroll
      for (int kr = 0; kr < kOcOws; ++kr) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int kc2 = 0; kc2 < kOcOws; kc2 += 2) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(ds + kr * kOcOws + kc2);
          s0 += unpack_lo<T>(w); s1 += unpack_hi<T>(w);
        }
        mine[kr] = (s0 + s1) * kInvS;
      }
    }
    // shifted 16-bit tables for the table gradients: Cs[i][r] = C_i[r + y - 11], Rs[i][r] = R_i[r + x - 11] (zero outside 0..11)
    {
      const int pos = which ? x : y;
      u16* dst = wCs(warp) + which * (16 * kObCs) + il * kObCs;
#pragma unroll
      for (int r2 = 0; r2 < kObCs; r2 += 2) {
        const int k0 = r2 + pos - (kOcOws - 1), k1 = k0 + 1;
        const float v0 = (k0 >= 0 && k0 < kOcOws) ? mine[k0] * kS : 0.f;
        const float v1 = (k1 >= 0 && k1 < kOcOws) ? mine[k1] * kS : 0.f;
        *reinterpret_cast<uint32_t*>(dst + r2) = pack2<T>(v0, v1);
      }
    }
  }
  __syncwarp();
  {
    // relative-position part of dQs in the accumulator layout: rows row0 / row1 (window rows 2 warp / 2 warp + 1, column g),
    // channels 8 t + 2 q4 + {0, 1};  dQs_i += sum_kk C_i[kk] rel_w[kk - y + 11] + R_i[kk] rel_h[kk - x + 11]
    const float* c0p = wC(warp) + g * kOcOws;                 // query row0: il = g
    const float* c1p = c0p + 8 * kOcOws;                      // query row1: il = g + 8
    const float* r0p = c0p + 16 * kOcOws;
    const float* r1p = c1p + 16 * kOcOws;
#pragma unroll
    for (int kk = 0; kk < kOcOws; ++kk) {
      const float cv0 = c0p[kk], cv1 = c1p[kk], rv0 = r0p[kk], rv1 = r1p[kk];
      const float* rw = sRel[0][kk - g + kOcOws - 1] + 2 * q4;
      const float* rh0 = sRel[1][kk - 2 * warp + kOcOws - 1] + 2 * q4;
      const float* rh1 = sRel[1][kk - (2 * warp + 1) + kOcOws - 1] + 2 * q4;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float2 w2 = *reinterpret_cast<const float2*>(rw + 8 * t);
        const float2 h0 = *reinterpret_cast<const float2*>(rh0 + 8 * t), h1 = *reinterpret_cast<const float2*>(rh1 + 8 * t);
        dq[t][0] = fmaf(cv0 * kS, w2.x, fmaf(rv0 * kS, h0.x, dq[t][0])); dq[t][1] = fmaf(cv0 * kS, w2.y, fmaf(rv0 * kS, h0.y, dq[t][1]));
        dq[t][2] = fmaf(cv1 * kS, w2.x, fmaf(rv1 * kS, h1.x, dq[t][2])); dq[t][3] = fmaf(cv1 * kS, w2.y, fmaf(rv1 * kS, h1.y, dq[t][3]));
      }
    }
    // dq = dQs / 4 (qs = q / 4), unscaled
    const float sc = 0.25f * kInvS;
    const size_t pix0 = (size_t)(wy * kOcWs + 2 * warp) * a.W + wx * kOcWs + g;
    unsigned short* op = a.dqkv + (size_t)b * a.gbs + h * kOcDh + 2 * q4;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      *reinterpret_cast<uint32_t*>(op + pix0 * a.gpitch + 8 * t) = pack2<T>(dq[t][0] * sc, dq[t][1] * sc);
      *reinterpret_cast<uint32_t*>(op + (pix0 + a.W) * a.gpitch + 8 * t) = pack2<T>(dq[t][2] * sc, dq[t][3] * sc);
    }
  }
  __syncthreads();

  // =============================== phase B: contractions over the 64 queries ===============================
  const size_t cta = ((size_t)b * nwin + blockIdx.x) * a.heads + h;
  float* okv = a.ws_kv + cta * kObKV;
  uint32_t qf[4][4], of[4][4];                       // B operands: Qs / dO as [k = query][n = channel]
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    ldsm_x4_trans(qf[ks], smem_u32(&sQ[ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)][8 * (lane >> 4)]));
    ldsm_x4_trans(of[ks], smem_u32(&sdO[ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)][8 * (lane >> 4)]));
  }
  for (int mt = warp; mt < 9; mt += 4) {             // dK rows 16 mt .. 16 mt + 15
    float acc[2][4];
#pragma unroll
    for (int t = 0; t < 2; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t af[4];
      ldsm_x4_trans(af, smem_u32(&sdS[ks * 16 + (lane & 7) + 8 * (lane >> 4)][mt * 16 + 8 * ((lane >> 3) & 1)]));
      mma16816<T>(acc[0], af, qf[ks][0], qf[ks][1]);
      mma16816<T>(acc[1], af, qf[ks][2], qf[ks][3]);
    }
    float* o0 = okv + (size_t)(mt * 16 + g) * 32 + 2 * q4;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      *reinterpret_cast<float2*>(o0 + 8 * t) = make_float2(acc[t][0] * kInvS, acc[t][1] * kInvS);
      *reinterpret_cast<float2*>(o0 + 8 * 32 + 8 * t) = make_float2(acc[t][2] * kInvS, acc[t][3] * kInvS);
    }
  }
  for (int mt = (warp + 2) & 3; mt < 9; mt += 4) {   // dV rows (the other warps' turn: 5 / 4 / 5 / 4 tiles per warp)
    float acc[2][4];
#pragma unroll
    for (int t = 0; t < 2; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t af[4];
      ldsm_x4_trans(af, smem_u32(&sP[ks * 16 + (lane & 7) + 8 * (lane >> 4)][mt * 16 + 8 * ((lane >> 3) & 1)]));
      mma16816<T>(acc[0], af, of[ks][0], of[ks][1]);
      mma16816<T>(acc[1], af, of[ks][2], of[ks][3]);
    }
    float* o0 = okv + (size_t)(mt * 16 + g) * 32 + 16 + 2 * q4;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      *reinterpret_cast<float2*>(o0 + 8 * t) = make_float2(acc[t][0], acc[t][1]);
      *reinterpret_cast<float2*>(o0 + 8 * 32 + 8 * t) = make_float2(acc[t][2], acc[t][3]);
    }
  }
  {
    // table gradients: warp = (table, 16-row tile); rows r >= 23 are padding
    const int tbl = warp >> 1, mt = warp & 1;
    float acc[2][4];
#pragma unroll
    for (int t = 0; t < 2; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t af[4];
      // rows 16 ks .. 16 ks + 15 are warp ks' scratch; columns >= 23 are padding (they reach rows of the output that are not stored)
      ldsm_x4_trans(af, smem_u32(wCs(ks) + tbl * (16 * kObCs) + ((lane & 7) + 8 * (lane >> 4)) * kObCs + mt * 16 + 8 * ((lane >> 3) & 1)));
      mma16816<T>(acc[0], af, qf[ks][0], qf[ks][1]);
      mma16816<T>(acc[1], af, qf[ks][2], qf[ks][3]);
    }
    float* orel = a.ws_rel + cta * kObRel + (size_t)tbl * kOcRel * kOcDh;
    const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (r0 < kOcRel) *reinterpret_cast<float2*>(orel + r0 * kOcDh + 8 * t + 2 * q4) = make_float2(acc[t][0] * kInvS, acc[t][1] * kInvS);
      if (r1 < kOcRel) *reinterpret_cast<float2*>(orel + r1 * kOcDh + 8 * t + 2 * q4) = make_float2(acc[t][2] * kInvS, acc[t][3] * kInvS);
    }
  }
}

// gather: dk, dv of pixel (y, x), head h = sum over the key windows containing it.  one thread per (pixel, head, dk|dv)
template <class T>
__global__ void __launch_bounds__(256)
ocab_bwd_gather_kernel(const ObArgs a, long long total) {
  const int nw = a.W / kOcWs, nh = a.H / kOcWs, nwin = nh * nw;
  constexpr int pad = (kOcOws - kOcWs) / 2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int sel = (int)(e & 1);
    long long r = e >> 1;
    const int h = (int)(r % a.heads); r /= a.heads;
    const int x = (int)(r % a.W); r /= a.W;
    const int y = (int)(r % a.H);
    const int b = (int)(r / a.H);
    float acc[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[d] = 0.f;
    // windows wy with wy*8 - pad <= y <= wy*8 - pad + 11
    const int wy1 = min((y + pad) / kOcWs, nh - 1), wx1 = min((x + pad) / kOcWs, nw - 1);
    for (int wy = max(wy1 - 1, 0); wy <= wy1; ++wy) {
      const int kr = y - (wy * kOcWs - pad);
      if (kr < 0 || kr >= kOcOws) continue;
      for (int wx = max(wx1 - 1, 0); wx <= wx1; ++wx) {
        const int kc = x - (wx * kOcWs - pad);
        if (kc < 0 || kc >= kOcOws) continue;
        const float* p = a.ws_kv + ((((size_t)b * nwin + wy * nw + wx) * a.heads + h) * kOcKeys + kr * kOcOws + kc) * 32 + sel * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
          acc[4 * q] += v.x; acc[4 * q + 1] += v.y; acc[4 * q + 2] += v.z; acc[4 * q + 3] += v.w;
        }
      }
    }
    unsigned short* o = a.dqkv + (size_t)b * a.gbs + ((size_t)y * a.W + x) * a.gpitch + (size_t)(1 + sel) * a.inner + h * kOcDh;
    uint4 o0, o1;
    o0.x = pack2<T>(acc[0], acc[1]); o0.y = pack2<T>(acc[2], acc[3]); o0.z = pack2<T>(acc[4], acc[5]); o0.w = pack2<T>(acc[6], acc[7]);
    o1.x = pack2<T>(acc[8], acc[9]); o1.y = pack2<T>(acc[10], acc[11]); o1.z = pack2<T>(acc[12], acc[13]); o1.w = pack2<T>(acc[14], acc[15]);
    *reinterpret_cast<uint4*>(o) = o0;
    *reinterpret_cast<uint4*>(o + 8) = o1;
  }
}

// reduce the per-(window, head) table partials in a fixed order: one block per table entry group
__global__ void __launch_bounds__(256)
ocab_bwd_rel_kernel(const float* __restrict__ ws_rel, long long nparts, float inv_scale, float* __restrict__ dst_w, float* __restrict__ dst_h) {
  __shared__ double red[256];
  const int e = blockIdx.x;                               // 0 .. kObRel - 1
  double s = 0.0;
  for (long long p = threadIdx.x; p < nparts; p += 256) s += (double)ws_rel[p * kObRel + e];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float v = (float)red[0] * inv_scale;
    if (e < kOcRel * kOcDh) dst_w[e] = v; else dst_h[e - kOcRel * kOcDh] = v;
  }
}

}  // namespace pir

extern "C" int pir_ocab(const PirOcab* d, void* stream) {
  using namespace pir;
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_ocab: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->heads <= 0) return pir_fail(PIR_ERR_ARG, "pir_ocab: empty problem");
  if (d->ws != kOcWs || d->ows != kOcOws || d->dim_head != kOcDh)
    return pir_fail(PIR_ERR_UNSUPPORTED, "pir_ocab: built for window 8, overlapping window 12, head dim 16 (got %d, %d, %d)", d->ws, d->ows, d->dim_head);
  if ((d->H % kOcWs) || (d->W % kOcWs)) return pir_fail(PIR_ERR_ARG, "pir_ocab: H and W must be multiples of the window size");
  if ((d->qkv_pitch % 8) || (d->qkv_bstride % 8) || (d->out_pitch % 8) || (d->out_bstride % 8) || ((uintptr_t)d->qkv & 15) || ((uintptr_t)d->out & 15) ||
      !d->qkv || !d->out || !d->rel_h || !d->rel_w)
    return pir_fail(PIR_ERR_ARG, "pir_ocab: tensors missing or not vector aligned");
  if (d->heads > 65535 || d->B > 65535) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_ocab: grid too large");
  OcArgs a{};
  a.H = d->H; a.W = d->W; a.heads = d->heads; a.inner = d->heads * kOcDh;
  a.qkv = reinterpret_cast<const unsigned short*>(d->qkv); a.qpitch = d->qkv_pitch; a.qbs = d->qkv_bstride;
  a.rel_h = d->rel_h; a.rel_w = d->rel_w;
  a.out = reinterpret_cast<unsigned short*>(d->out); a.opitch = d->out_pitch; a.obs = d->out_bstride;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((d->H / kOcWs) * (d->W / kOcWs)), (unsigned)d->heads, (unsigned)d->B);
  static const bool simt = getenv("PIR_OCAB_SIMT") != nullptr;      // A/B: the fp32 SIMT version
  if (simt) {
    if (d->dtype == PIR_DTYPE_BF16) ocab_simt_kernel<BF16><<<grid, 256, 0, s>>>(a);
    else ocab_simt_kernel<FP16><<<grid, 256, 0, s>>>(a);
  } else {
    if (d->dtype == PIR_DTYPE_BF16) ocab_kernel<BF16><<<grid, 128, 0, s>>>(a);
    else ocab_kernel<FP16><<<grid, 128, 0, s>>>(a);
  }
  return pir_check_launch("pir_ocab");
}

extern "C" int64_t pir_ocab_bwd_ws_floats(int32_t B, int32_t H, int32_t W, int32_t heads) {
  if (B <= 0 || H <= 0 || W <= 0 || heads <= 0) return 0;
  return (int64_t)B * (H / pir::kOcWs) * (W / pir::kOcWs) * heads * (pir::kObKV + pir::kObRel);
}

extern "C" int pir_ocab_bwd(const PirOcabBwd* d, void* stream) {
  using namespace pir;
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_ocab_bwd: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->heads <= 0) return pir_fail(PIR_ERR_ARG, "pir_ocab_bwd: empty problem");
  if (d->ws_ != kOcWs || d->ows != kOcOws || d->dim_head != kOcDh)
    return pir_fail(PIR_ERR_UNSUPPORTED, "pir_ocab_bwd: built for window 8, overlapping window 12, head dim 16");
  if ((d->H % kOcWs) || (d->W % kOcWs)) return pir_fail(PIR_ERR_ARG, "pir_ocab_bwd: H and W must be multiples of the window size");
  if ((d->qkv_pitch % 8) || (d->qkv_bstride % 8) || (d->dout_pitch % 8) || (d->dout_bstride % 8) || (d->dqkv_pitch % 8) || (d->dqkv_bstride % 8) ||
      ((uintptr_t)d->qkv & 15) || ((uintptr_t)d->dout & 15) || ((uintptr_t)d->dqkv & 15) || ((uintptr_t)d->ws & 15) || !d->qkv || !d->dout || !d->dqkv ||
      !d->ws || !d->rel_h || !d->rel_w || !d->dst_rel_h || !d->dst_rel_w)
    return pir_fail(PIR_ERR_ARG, "pir_ocab_bwd: tensors missing or not 16-byte aligned");
  if (d->heads > 65535 || d->B > 65535) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_ocab_bwd: grid too large");
  const int nwin = (d->H / kOcWs) * (d->W / kOcWs);
  ObArgs a{};
  a.H = d->H; a.W = d->W; a.heads = d->heads; a.inner = d->heads * kOcDh;
  a.qkv = reinterpret_cast<const unsigned short*>(d->qkv); a.qpitch = d->qkv_pitch; a.qbs = d->qkv_bstride;
  a.dout = reinterpret_cast<const unsigned short*>(d->dout); a.dpitch = d->dout_pitch; a.dbs = d->dout_bstride;
  a.rel_h = d->rel_h; a.rel_w = d->rel_w;
  a.dqkv = reinterpret_cast<unsigned short*>(d->dqkv); a.gpitch = d->dqkv_pitch; a.gbs = d->dqkv_bstride;
  const long long nparts = (long long)d->B * nwin * d->heads;
  a.ws_kv = d->ws;
  a.ws_rel = d->ws + (size_t)nparts * kObKV;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = sizeof(float) * (2 * kOcKeys * kOcRow + 2 * kOcRel * kOcDh + 2 * 64 * kOcDh + 2 * 64 * (kOcKeys + 1) + 4 * 64 * kOcOws);
  const int fi = d->dtype == PIR_DTYPE_BF16 ? 1 : 0;
  if (!pir_smem_attr_once(fi ? reinterpret_cast<const void*>(ocab_bwd_kernel<BF16>) : reinterpret_cast<const void*>(ocab_bwd_kernel<FP16>), (int)smem,
                          "pir_ocab_bwd")) return PIR_ERR_CUDA;
  dim3 grid((unsigned)nwin, (unsigned)d->heads, (unsigned)d->B);
  static const bool simt = getenv("PIR_OCAB_BWD_SIMT") != nullptr;   // A/B: the fp32 SIMT version
  if (simt) {
    if (fi) ocab_bwd_kernel<BF16><<<grid, 256, smem, s>>>(a); else ocab_bwd_kernel<FP16><<<grid, 256, smem, s>>>(a);
  } else {
    if (!pir_smem_attr_once(fi ? reinterpret_cast<const void*>(ocab_bwd_mma_kernel<BF16>) : reinterpret_cast<const void*>(ocab_bwd_mma_kernel<FP16>),
                            (int)kObMmaSmem, "pir_ocab_bwd")) return PIR_ERR_CUDA;
    if (fi) ocab_bwd_mma_kernel<BF16><<<grid, 128, kObMmaSmem, s>>>(a); else ocab_bwd_mma_kernel<FP16><<<grid, 128, kObMmaSmem, s>>>(a);
  }
  if (int e = pir_check_launch("pir_ocab_bwd")) return e;
  const long long total = (long long)d->B * d->H * d->W * d->heads * 2;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (fi) ocab_bwd_gather_kernel<BF16><<<(unsigned)blocks, 256, 0, s>>>(a, total); else ocab_bwd_gather_kernel<FP16><<<(unsigned)blocks, 256, 0, s>>>(a, total);
  if (int e = pir_check_launch("pir_ocab_bwd(gather)")) return e;
  ocab_bwd_rel_kernel<<<kObRel, 256, 0, s>>>(a.ws_rel, nparts, d->inv_scale, d->dst_rel_w, d->dst_rel_h);
  return pir_check_launch("pir_ocab_bwd(tables)");
}
