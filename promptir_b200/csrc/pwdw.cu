// Fused  LayerNorm -> 1x1 conv -> depthwise 3x3 (-> GELU gate)  for sm_100a: one kernel for the MDTA qkv branch
// (net/model.py:60-63 + :111-112 + :120) and one for the GDFN front half (:60-63 + :88-90 + :96-97).  The widest
// activation of a TransformerBlock (3C resp. 2*hidden channels per pixel) never reaches HBM:
//
//   HBM traffic per pixel:  read C (x 1.3-1.5 for the halo, mostly L2 hits)  +  write 3C (qkv) or hidden (gated)
//   instead of              C + 2*(3C) + 3C                                 resp.  C + 2*(2 hidden) + hidden.
//
// One persistent CTA per SM walks (image, spatial tile) items.  Per item
//   * TMA brings the tile + 1-pixel halo of x (all K channels, 128B-swizzled k-blocks) into shared memory once;
//     pixels outside the image are the TMA out-of-bounds zero fill;
//   * per chunk of 64 pre-conv channels a tcgen05 GEMM (M = 128-pixel row groups, N = 64, fp32 accumulation in
//     TMEM, double buffered) runs one chunk ahead of the CUDA cores; its weights stream through a 2-stage TMA ring;
//   * eight compute warps drain the accumulators (LayerNorm fold: rstd*acc - rstd*mu*s[n] + t[n], zero outside the
//     image = the conv's zero padding) into a shared-memory fp16 tile, then run the 3x3 stencil from it with
//     packed HFMA2 (fp16 inputs, weights and accumulation: the intermediate keeps 11 mantissa bits, more than the
//     bf16 tensor it replaces; values saturate at +-65504), apply the exact-erf GELU gate in fp32 and store.
//
// Warp roles (320 threads): 0 TMA producer | 1 TMEM alloc + MMA issue | 2-9 statistics, drain, stencil, store.
#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kFwThreads = 320;
constexpr int kFwCompute = 256;
constexpr int kFwChunk = 64;                 // pre-conv channels per chunk (gate: 32 x1 + 32 x2)

struct FwArgs {
  int B, H, W, C;          // input tensor
  int n_pre;               // pre-conv channels (3C or 2*hp)
  int hp;                  // gate: gated channels (n_pre / 2); plain: unused
  int gate;
  int ln_mode;
  int nkb;                 // k-blocks of 64 input channels
  int n_chunks;
  int tiles_x, tiles_y, n_items;
  int dw_stride;           // channels per tap row of the depthwise weights (= n_pre)
  uint32_t off_b, off_conv, off_dw, off_vec, off_bias, off_stats;   // byte offsets from the 1024-aligned base
  const void* dw_w;        // [9][n_pre] fp16
  const float* dw_bias;    // [n_pre] or null
  const float* ln_s;       // [n_pre]
  const float* vec_t;      // [n_pre] or null
  void* out;
  long long out_pitch, out_bstride;
};

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
__device__ __forceinline__ uint32_t pack_f16_sat(float lo, float hi) {
  uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ void fw_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kFwCompute) : "memory"); }

// MT: 128-row groups of the halo'd tile held in shared memory; TW: interior tile width; R: interior rows per thread.
// interior tile = (R * 32 / TW) x TW pixels, halo'd tile = (R * 32 / TW + 2) x (TW + 2) <= MT * 128 rows.
template <class T, int MT, int TW, int R, bool GATE>
__global__ void __launch_bounds__(kFwThreads, 1)
pwdw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwArgs g) {
  constexpr int BANDS = 32 / TW;
  constexpr int TH = R * BANDS;
  constexpr int SW = TW + 2;
  constexpr int NPIX = (TH + 2) * SW;
  static_assert(NPIX <= MT * 128, "halo'd tile does not fit the M tiles");
  constexpr uint32_t A_KB_BYTES = MT * 128 * 128;           // one k-block of the halo'd tile
  constexpr uint32_t B_KB_BYTES = kFwChunk * 128;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull, bar_aempty;
  __shared__ __align__(8) uint64_t bar_bfull[2], bar_bempty[2];
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t b_stage_bytes = (uint32_t)g.nkb * B_KB_BYTES;
  uint8_t* conv_buf = base_ptr + g.off_conv;
  const unsigned short* sdw = reinterpret_cast<const unsigned short*>(base_ptr + g.off_dw);
  float* svec = reinterpret_cast<float*>(base_ptr + g.off_vec);              // [2][n_vec]: ln_s | vec_t
  float* sbias = reinterpret_cast<float*>(base_ptr + g.off_bias);            // [n_vec] depthwise bias
  float2* sstats = reinterpret_cast<float2*>(base_ptr + g.off_stats);        // [MT*128]: (-rstd*mu, rstd); rstd < 0 => outside
  const int n_vec = g.n_chunks * kFwChunk + g.hp;                            // covers every index the drain may touch

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    mbar_init(smem_u32(&bar_afull), 1); mbar_init(smem_u32(&bar_aempty), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_bfull[i]), 1); mbar_init(smem_u32(&bar_bempty[i]), 1);
      mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), MT == 3 ? 512 : 256); tmem_relinquish(); }
  // per-CTA constants: depthwise taps (fp16), LayerNorm fold vectors, biases
  {
    const unsigned short* src = reinterpret_cast<const unsigned short*>(g.dw_w);
    unsigned short* dst = const_cast<unsigned short*>(sdw);
    for (int i = threadIdx.x; i < 9 * n_vec; i += kFwThreads) {
      const int t = i / n_vec, c = i - t * n_vec;
      dst[i] = c < g.n_pre ? src[(size_t)t * g.dw_stride + c] : (unsigned short)0;
    }
    for (int i = threadIdx.x; i < n_vec; i += kFwThreads) {
      svec[i] = (g.ln_s && i < g.n_pre) ? g.ln_s[i] : 0.f;
      svec[n_vec + i] = (g.vec_t && i < g.n_pre) ? g.vec_t[i] : 0.f;
      sbias[i] = (g.dw_bias && i < g.n_pre) ? g.dw_bias[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int per_img = g.tiles_x * g.tiles_y;

  if (warp == 0) {
    // ========================================= TMA producer ==========================================
    uint32_t q = 0, it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      const int b = item / per_img, r = item % per_img;
      const int x0 = (r % g.tiles_x) * TW - 1, y0 = (r / g.tiles_x) * TH - 1;
      mbar_wait(smem_u32(&bar_aempty), (it & 1u) ^ 1u);
      if (lane == 0) {
        const uint32_t full = smem_u32(&bar_afull);
        mbar_expect_tx(full, (uint32_t)g.nkb * (uint32_t)(NPIX * 128));
        for (int kb = 0; kb < g.nkb; ++kb) tma_load_4d(base + (uint32_t)kb * A_KB_BYTES, &tmA, full, kb * 64, x0, y0, b);
      }
      __syncwarp();
      for (int c = 0; c < g.n_chunks; ++c, ++q) {
        const uint32_t st = q & 1u;
        mbar_wait(smem_u32(&bar_bempty[st]), ((q >> 1) & 1u) ^ 1u);
        if (lane == 0) {
          const uint32_t full = smem_u32(&bar_bfull[st]);
          const uint32_t dst = base + g.off_b + st * b_stage_bytes;
          const int r0 = g.gate ? c * 32 : c * 64, r1 = g.gate ? g.hp + c * 32 : c * 64 + 32;
          mbar_expect_tx(full, b_stage_bytes);
          for (int kb = 0; kb < g.nkb; ++kb) {
            tma_load_2d(dst + (uint32_t)kb * B_KB_BYTES, &tmB, full, kb * 64, r0);
            tma_load_2d(dst + (uint32_t)kb * B_KB_BYTES + 4096u, &tmB, full, kb * 64, r1);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ========================================= MMA issuer ============================================
    const uint32_t idesc = make_idesc_f16(T::kFmt, 128, kFwChunk, 0, 0);
    uint32_t q = 0, it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      mbar_wait(smem_u32(&bar_afull), it & 1u);
      for (int c = 0; c < g.n_chunks; ++c, ++q) {
        const uint32_t st = q & 1u, ph = (q >> 1) & 1u;
        mbar_wait(smem_u32(&bar_bfull[st]), ph);
        mbar_wait(smem_u32(&bar_tempty[st]), ph ^ 1u);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t b_src = base + g.off_b + st * b_stage_bytes;
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            const uint32_t d = tmem_base + st * (uint32_t)(MT * kFwChunk) + (uint32_t)(t * kFwChunk);
            for (int kb = 0; kb < g.nkb; ++kb) {
              const int rem = g.C - kb * 64;
              const int ksteps = rem >= 64 ? 4 : (rem + 15) >> 4;
              const uint32_t a_src = base + (uint32_t)kb * A_KB_BYTES + (uint32_t)t * 16384u;
              for (int k = 0; k < ksteps; ++k)
                umma_f16(d, make_sdesc_sw128(a_src + k * 32, 16, 1024), make_sdesc_sw128(b_src + kb * B_KB_BYTES + k * 32, 16, 1024),
                         idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&bar_bempty[st]));
          umma_commit(smem_u32(&bar_tfull[st]));
          if (c == g.n_chunks - 1) umma_commit(smem_u32(&bar_aempty));        // x tile no longer needed by the tensor core
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ statistics, drain, stencil, store ===============================
    const int ct = threadIdx.x - 64;                 // 0..255
    const int cw = ct >> 5;                          // compute warp 0..7
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may read
    const int sub = cw >> 2;                         // warps 2-5 drain columns [0,32) of a chunk, warps 6-9 columns [32,64)
    // stencil mapping
    const int cg = ct & 7;
    const int tx = (ct >> 3) % TW;
    const int band = (ct >> 3) / TW;
    uint32_t q = 0, it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      const int b = item / per_img, rr = item % per_img;
      const int x0 = (rr % g.tiles_x) * TW, y0 = (rr / g.tiles_x) * TH;
      // ---- per-pixel LayerNorm statistics of the halo'd tile (or just the inside/outside flag) ----
      mbar_wait(smem_u32(&bar_afull), it & 1u);
      for (int m = ct; m < NPIX; m += kFwCompute) {
        const int py = y0 - 1 + m / SW, px = x0 - 1 + m % SW;
        const bool inside = py >= 0 && py < g.H && px >= 0 && px < g.W;
        float2 st = make_float2(0.f, 1.f);
        if (g.ln_mode && inside) {
          float s1 = 0.f, s2 = 0.f;
          for (int kb = 0; kb < g.nkb; ++kb) {
            const uint8_t* a_row = base_ptr + (size_t)kb * A_KB_BYTES + (size_t)m * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 v = *reinterpret_cast<const uint4*>(a_row + ((j ^ (m & 7)) << 4));
              const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = unpack_lo<T>(w4[e]), bb = unpack_hi<T>(w4[e]);
                s1 += a + bb;
                s2 = fmaf(a, a, s2);
                s2 = fmaf(bb, bb, s2);
              }
            }
          }
          const float inv_k = 1.0f / (float)g.C;
          const float mu = s1 * inv_k;
          const float rstd = rsqrtf(fmaxf(fmaf(s2, inv_k, -mu * mu), 0.f) + 1e-5f);
          st = make_float2(g.ln_mode == 2 ? 0.f : -rstd * mu, rstd);
        }
        if (!inside) st.y = -1.f;
        sstats[m] = st;
      }
      fw_bar(1);

      for (int c = 0; c < g.n_chunks; ++c, ++q) {
        const uint32_t st = q & 1u, ph = (q >> 1) & 1u;
        // ---- drain: TMEM -> LayerNorm fold -> fp16 shared-memory tile ----
        mbar_wait(smem_u32(&bar_tfull[st]), ph);
        tc_fence_after();
        const int nb = GATE ? (sub ? g.hp + c * 32 : c * 32) : c * 64 + sub * 32;     // first pre-conv channel of this half
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const int m = t * 128 + quarter * 32 + lane;
          uint32_t acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + st * (uint32_t)(MT * kFwChunk) + (uint32_t)(t * kFwChunk + sub * 32), acc);
          const float2 ps = sstats[m < NPIX ? m : 0];
          tmem_ld_wait();
          if (m < NPIX) {
            const bool inside = ps.y >= 0.f;
            const float4* sv4 = reinterpret_cast<const float4*>(svec + nb);
            const float4* tv4 = reinterpret_cast<const float4*>(svec + n_vec + nb);
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 s4 = sv4[e], t4 = tv4[e];
              const float v0 = fmaf(ps.y, __uint_as_float(acc[4 * e + 0]), fmaf(ps.x, s4.x, t4.x));
              const float v1 = fmaf(ps.y, __uint_as_float(acc[4 * e + 1]), fmaf(ps.x, s4.y, t4.y));
              const float v2 = fmaf(ps.y, __uint_as_float(acc[4 * e + 2]), fmaf(ps.x, s4.z, t4.z));
              const float v3 = fmaf(ps.y, __uint_as_float(acc[4 * e + 3]), fmaf(ps.x, s4.w, t4.w));
              pk[2 * e] = inside ? pack_f16_sat(v0, v1) : 0u;
              pk[2 * e + 1] = inside ? pack_f16_sat(v2, v3) : 0u;
            }
            if (GATE) {
              // two planes [pixel][32 ch] (64-byte rows), 16-byte chunks swizzled by (m >> 1) & 3
              uint8_t* row = conv_buf + (size_t)sub * (MT * 128 * 64) + (size_t)m * 64;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(row + ((j ^ ((m >> 1) & 3)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            } else {
              // one plane [pixel][64 ch] (128-byte rows), 16-byte chunks swizzled by m & 7
              uint8_t* row = conv_buf + (size_t)m * 128;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(row + (((sub * 4 + j) ^ (m & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[st]));       // accumulator buffer may be overwritten
        fw_bar(2);                                                     // fp16 tile complete

        // ---- stencil from shared memory ----
        if (GATE) {
          const int ch = c * 32 + cg * 4;                              // gated channel of this thread
          const bool c_ok = ch < g.hp;
          uint2 w1[9], w2[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            w1[t] = *reinterpret_cast<const uint2*>(sdw + (size_t)t * n_vec + ch);
            w2[t] = *reinterpret_cast<const uint2*>(sdw + (size_t)t * n_vec + g.hp + ch);
          }
          float b1[4], b2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { b1[i] = sbias[ch + i]; b2[i] = sbias[g.hp + ch + i]; }
          const int x = x0 + tx;
          const bool ok = c_ok && x < g.W;
          unsigned short* outp = reinterpret_cast<unsigned short*>(g.out) + (size_t)b * g.out_bstride + ch;
          uint32_t p[3][2], qq[3][2];
#pragma unroll
          for (int r = 0; r < R + 2; ++r) {
            if (r < R) { p[r % 3][0] = p[r % 3][1] = 0u; qq[r % 3][0] = qq[r % 3][1] = 0u; }
            uint2 v1[3], v2[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int m = (band * R + r) * SW + tx + kx;
              const uint8_t* s = conv_buf + (size_t)m * 64 + ((((cg >> 1) ^ ((m >> 1) & 3)) << 4) | ((cg & 1) << 3));
              v1[kx] = *reinterpret_cast<const uint2*>(s);
              v2[kx] = *reinterpret_cast<const uint2*>(s + MT * 128 * 64);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int o = r - ky;
              if (o >= 0 && o < R) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  p[o % 3][0] = hfma2_u(v1[kx].x, w1[ky * 3 + kx].x, p[o % 3][0]);
                  p[o % 3][1] = hfma2_u(v1[kx].y, w1[ky * 3 + kx].y, p[o % 3][1]);
                  qq[o % 3][0] = hfma2_u(v2[kx].x, w2[ky * 3 + kx].x, qq[o % 3][0]);
                  qq[o % 3][1] = hfma2_u(v2[kx].y, w2[ky * 3 + kx].y, qq[o % 3][1]);
                }
              }
            }
            const int o = r - 2;
            if (o >= 0) {
              const int y = y0 + band * R + o;
              if (ok && y < g.H) {
                const float2 pa = h2_to_f2(p[o % 3][0]), pb = h2_to_f2(p[o % 3][1]);
                const float2 qa = h2_to_f2(qq[o % 3][0]), qb = h2_to_f2(qq[o % 3][1]);
                uint2 ov;
                ov.x = pack2<T>(gelu_erf(pa.x + b1[0]) * (qa.x + b2[0]), gelu_erf(pa.y + b1[1]) * (qa.y + b2[1]));
                ov.y = pack2<T>(gelu_erf(pb.x + b1[2]) * (qb.x + b2[2]), gelu_erf(pb.y + b1[3]) * (qb.y + b2[3]));
                *reinterpret_cast<uint2*>(outp + ((size_t)y * g.W + x) * g.out_pitch) = ov;
              }
            }
          }
        } else {
          const int ch = c * 64 + cg * 8;
          const bool c_ok = ch < g.n_pre;
          uint4 wt[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) wt[t] = *reinterpret_cast<const uint4*>(sdw + (size_t)t * n_vec + ch);
          float bb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) bb[i] = sbias[ch + i];
          const int x = x0 + tx;
          const bool ok = c_ok && x < g.W;
          unsigned short* outp = reinterpret_cast<unsigned short*>(g.out) + (size_t)b * g.out_bstride + ch;
          uint32_t p[3][4];
#pragma unroll
          for (int r = 0; r < R + 2; ++r) {
            if (r < R) { p[r % 3][0] = p[r % 3][1] = p[r % 3][2] = p[r % 3][3] = 0u; }
            uint4 v[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int m = (band * R + r) * SW + tx + kx;
              v[kx] = *reinterpret_cast<const uint4*>(conv_buf + (size_t)m * 128 + ((cg ^ (m & 7)) << 4));
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int o = r - ky;
              if (o >= 0 && o < R) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  const uint4 w = wt[ky * 3 + kx];
                  p[o % 3][0] = hfma2_u(v[kx].x, w.x, p[o % 3][0]);
                  p[o % 3][1] = hfma2_u(v[kx].y, w.y, p[o % 3][1]);
                  p[o % 3][2] = hfma2_u(v[kx].z, w.z, p[o % 3][2]);
                  p[o % 3][3] = hfma2_u(v[kx].w, w.w, p[o % 3][3]);
                }
              }
            }
            const int o = r - 2;
            if (o >= 0) {
              const int y = y0 + band * R + o;
              if (ok && y < g.H) {
                uint4 ov;
                uint32_t* op = &ov.x;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = h2_to_f2(p[o % 3][e]);
                  op[e] = pack2<T>(f.x + bb[2 * e], f.y + bb[2 * e + 1]);
                }
                *reinterpret_cast<uint4*>(outp + ((size_t)y * g.W + x) * g.out_pitch) = ov;
              }
            }
          }
        }
        fw_bar(3);                                                     // fp16 tile may be overwritten
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, MT == 3 ? 512 : 256); }
}

// ---------------------------------------------------------------------------------------------------
struct FwPlan {
  int mt, tw, rows;        // template configuration
  uint32_t smem;
  FwArgs g;
};

static int plan_pwdw(const PirPwDw* d, FwPlan* p) {
  FwArgs& g = p->g;
  g = FwArgs{};
  g.B = d->B; g.H = d->H; g.W = d->W; g.C = d->C;
  g.gate = d->gate;
  g.n_pre = d->gate ? 2 * d->N : d->N;
  g.hp = d->gate ? d->N : 0;
  g.ln_mode = d->ln_mode;
  g.nkb = (d->C + 63) / 64;
  g.n_chunks = d->gate ? (d->N + 31) / 32 : (d->N + 63) / 64;
  g.dw_stride = g.n_pre;
  const int n_vec = g.n_chunks * kFwChunk + g.hp;
  // shared-memory plan: [A: nkb x MT x 16 KB] [B ring: 2 x nkb x 8 KB] [fp16 tile: MT x 16 KB] [dw taps] [ln_s|vec_t] [dw bias] [stats]
  for (int mt = 3; mt >= 2; --mt) {
    uint32_t off = (uint32_t)g.nkb * mt * 16384u;
    g.off_b = off; off += 2u * g.nkb * 8192u;
    g.off_conv = off; off += (uint32_t)mt * 16384u;
    g.off_dw = off; off += (uint32_t)((9 * n_vec * 2 + 15) / 16 * 16);
    g.off_vec = off; off += 2u * n_vec * 4u;
    g.off_bias = off; off += (uint32_t)n_vec * 4u;
    g.off_stats = off; off += (uint32_t)mt * 128u * 8u;
    if (off + 1024u <= 227u * 1024u - 1024u) {
      p->mt = mt; p->smem = off + 1024u;
      if (mt == 3) { p->tw = d->W > 16 ? 32 : 16; p->rows = 8; }
      else { p->tw = 16; p->rows = 6; }
      const int th = p->rows * (32 / p->tw);
      g.tiles_x = (d->W + p->tw - 1) / p->tw;
      g.tiles_y = (d->H + th - 1) / th;
      g.n_items = g.tiles_x * g.tiles_y * d->B;
      return PIR_OK;
    }
  }
  return PIR_ERR_UNSUPPORTED;
}

template <class T, int MT, int TW, int R, bool GATE>
static int launch_cfg2(const FwPlan& p, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t stream) {
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(pwdw_kernel<T, MT, TW, R, GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024) != cudaSuccess)
      return pir_fail(PIR_ERR_CUDA, "pir_pwdw: cannot raise dynamic shared memory limit");
    set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int grid = p.g.n_items < num_sms ? p.g.n_items : num_sms;
  pwdw_kernel<T, MT, TW, R, GATE><<<grid, kFwThreads, p.smem, stream>>>(tmA, tmB, p.g);
  return pir_check_launch("pir_pwdw");
}
template <class T, int MT, int TW, int R>
static int launch_cfg(const PirPwDw* d, const FwPlan& p, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t stream) {
  return d->gate ? launch_cfg2<T, MT, TW, R, true>(p, tmA, tmB, stream) : launch_cfg2<T, MT, TW, R, false>(p, tmA, tmB, stream);
}

template <class T>
static int launch_pwdw(const PirPwDw* d, cudaStream_t stream) {
  FwPlan p;
  if (plan_pwdw(d, &p) != PIR_OK) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_pwdw: C = %d, N = %d does not fit the shared-memory plan", d->C, d->N);
  p.g.dw_w = d->dw_w; p.g.dw_bias = d->dw_bias; p.g.ln_s = d->ln_s; p.g.vec_t = d->vec_t;
  p.g.out = d->out; p.g.out_pitch = d->out_pitch; p.g.out_bstride = d->out_bstride;
  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB;
  {
    const int th = p.rows * (32 / p.tw);
    const uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->a_pitch * 2, (uint64_t)d->a_pitch * 2 * d->W, (uint64_t)d->a_bstride * 2};
    const uint32_t box[4] = {64, (uint32_t)(p.tw + 2), (uint32_t)(th + 2), 1};
    if (int e = pir_make_tmap(&tmA, dt, 4, d->a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t kpad = (uint64_t)p.g.nkb * 64;
    const uint64_t dims[2] = {kpad, (uint64_t)p.g.n_pre};
    const uint64_t strides[1] = {kpad * 2};
    const uint32_t box[2] = {64, 32};
    if (int e = pir_make_tmap(&tmB, dt, 2, d->w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  if (p.mt == 3 && p.tw == 32) return launch_cfg<T, 3, 32, 8>(d, p, tmA, tmB, stream);
  if (p.mt == 3) return launch_cfg<T, 3, 16, 8>(d, p, tmA, tmB, stream);
  return launch_cfg<T, 2, 16, 6>(d, p, tmA, tmB, stream);
}

}  // namespace pir

extern "C" int pir_pwdw_supported(int32_t C, int32_t N, int32_t gate) {
  PirPwDw d{};
  d.B = 1; d.H = 64; d.W = 64; d.C = C; d.N = N; d.gate = gate;
  pir::FwPlan p;
  if ((C % 8) || (N % 8)) return 0;
  return pir::plan_pwdw(&d, &p) == PIR_OK ? 1 : 0;
}

extern "C" int pir_pwdw(const PirPwDw* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_pwdw: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->N <= 0) return pir_fail(PIR_ERR_ARG, "pir_pwdw: empty problem");
  if ((d->C % 8) || (d->N % 8) || (d->a_pitch % 8) || (d->a_bstride % 8) || (d->out_pitch % 8) || (d->out_bstride % 8) ||
      ((uintptr_t)d->a & 15) || ((uintptr_t)d->w & 15) || ((uintptr_t)d->out & 15) || ((uintptr_t)d->dw_w & 15))
    return pir_fail(PIR_ERR_ARG, "pir_pwdw: channel counts / pitches / pointers are not 16-byte aligned");
  if (d->ln_mode && !d->ln_s) return pir_fail(PIR_ERR_ARG, "pir_pwdw: LayerNorm fold needs ln_s");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return d->dtype == PIR_DTYPE_BF16 ? pir::launch_pwdw<pir::BF16>(d, s) : pir::launch_pwdw<pir::FP16>(d, s);
}
