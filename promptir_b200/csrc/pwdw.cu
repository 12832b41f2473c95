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

struct FwArgs {
  int B, H, W, C;          // input tensor
  int n_pre;               // pre-conv channels (3C or 2*hp)
  int hp;                  // gate: gated channels (n_pre / 2); plain: 0
  int ln_mode;
  int nkb;                 // k-blocks of 64 input channels
  int n_chunks;
  int n_vec;               // length of the per-channel shared-memory vectors (covers every index a chunk may touch)
  int tiles_x, tiles_y, n_items;
  int has_bias;            // depthwise bias present
  uint32_t off_b, off_conv, off_dw, off_vec, off_bias, off_stats;   // byte offsets from the 1024-aligned base
  const void* dw_w;        // [9][n_pre] fp16
  const float* dw_bias;    // [n_pre] or null
  const float* ln_s;       // [n_pre]
  const float* vec_t;      // [n_pre] or null
  void* out;
  long long out_pitch, out_bstride;
};

__device__ __forceinline__ uint32_t pack_f16_sat(float lo, float hi) {
  uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// Compile-time shape of one kernel variant.
//   GATE : 512 compute threads, chunks of 64 pre-conv channels (32 of x1 + the matching 32 of x2), 4 gated channels/thread
//   plain: 384 compute threads, chunks of 48 pre-conv channels (3C is always a multiple of 48 here), 8 channels/thread
// Either way 64 pixel columns of threads: BANDS = 64 / TW row bands of R rows each.
template <int MT_, int TW_, int R_, bool GATE_>
struct FwCfg {
  static constexpr int MT = MT_, TW = TW_, R = R_;
  static constexpr bool GATE = GATE_;
  static constexpr int CG = GATE ? 8 : 6;                    // channel groups (threads) per pixel
  static constexpr int NCOMP = CG * 64;                      // compute threads
  static constexpr int NTHREADS = NCOMP + 64;                // + TMA warp + MMA warp
  static constexpr int CH = GATE ? 64 : 48;                  // pre-conv channels per chunk = UMMA N
  static constexpr int SLICES = CH / 16;                     // 16-column drain slices = compute warps per TMEM lane quarter
  static constexpr int BANDS = 64 / TW;
  static constexpr int TH = R * BANDS;
  static constexpr int SW = TW + 2;
  static constexpr int NPIX = (TH + 2) * SW;
  static constexpr int TMEM_COLS = 2 * MT * CH <= 256 ? 256 : 512;
  static constexpr uint32_t A_KB_BYTES = MT * 128 * 128;     // one k-block of the halo'd tile
  static constexpr uint32_t B_KB_BYTES = CH * 128;
  static constexpr int ROW_BYTES = GATE ? 64 : 96;           // fp16 tile row (per plane)
  static constexpr uint32_t CONV_BYTES = GATE ? 2u * MT * 128 * 64 : (uint32_t)MT * 128 * 96;
  static_assert(NPIX <= MT * 128, "halo'd tile does not fit the M tiles");
  static_assert(NCOMP / 32 == 4 * SLICES, "one compute warp per (TMEM lane quarter, 16-column slice)");
};

template <class T, class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, 1)
pwdw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwArgs g) {
  constexpr int MT = Cfg::MT, TW = Cfg::TW, R = Cfg::R, TH = Cfg::TH, SW = Cfg::SW, NPIX = Cfg::NPIX, CH = Cfg::CH;
  constexpr bool GATE = Cfg::GATE;
  constexpr int NCOMP = Cfg::NCOMP;
  constexpr uint32_t A_KB_BYTES = Cfg::A_KB_BYTES, B_KB_BYTES = Cfg::B_KB_BYTES;

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
  float2* sstats = reinterpret_cast<float2*>(base_ptr + g.off_stats);        // [MT*128]: (-rstd*mu, rstd)
  const int n_vec = g.n_vec;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    mbar_init(smem_u32(&bar_afull), 1); mbar_init(smem_u32(&bar_aempty), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_bfull[i]), 1); mbar_init(smem_u32(&bar_bempty[i]), 1);
      mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), NCOMP / 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), Cfg::TMEM_COLS); tmem_relinquish(); }
  // per-CTA constants: depthwise taps (fp16), LayerNorm fold vectors, biases
  {
    const unsigned short* src = reinterpret_cast<const unsigned short*>(g.dw_w);
    unsigned short* dst = const_cast<unsigned short*>(sdw);
    for (int i = threadIdx.x; i < 9 * n_vec; i += Cfg::NTHREADS) {
      const int t = i / n_vec, c = i - t * n_vec;
      dst[i] = c < g.n_pre ? src[(size_t)t * g.n_pre + c] : (unsigned short)0;
    }
    for (int i = threadIdx.x; i < n_vec; i += Cfg::NTHREADS) {
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
      mbar_wait_sleep(smem_u32(&bar_aempty), (it & 1u) ^ 1u);
      if (lane == 0) {
        const uint32_t full = smem_u32(&bar_afull);
        mbar_expect_tx(full, (uint32_t)g.nkb * (uint32_t)(NPIX * 128));
        for (int kb = 0; kb < g.nkb; ++kb) tma_load_4d(base + (uint32_t)kb * A_KB_BYTES, &tmA, full, kb * 64, x0, y0, b);
      }
      __syncwarp();
      for (int c = 0; c < g.n_chunks; ++c, ++q) {
        const uint32_t st = q & 1u;
        mbar_wait_sleep(smem_u32(&bar_bempty[st]), ((q >> 1) & 1u) ^ 1u);
        if (lane == 0) {
          const uint32_t full = smem_u32(&bar_bfull[st]);
          const uint32_t dst = base + g.off_b + st * b_stage_bytes;
          mbar_expect_tx(full, b_stage_bytes);
          for (int kb = 0; kb < g.nkb; ++kb) {
            if (GATE) {
              tma_load_2d(dst + (uint32_t)kb * B_KB_BYTES, &tmB, full, kb * 64, c * 32);
              tma_load_2d(dst + (uint32_t)kb * B_KB_BYTES + 4096u, &tmB, full, kb * 64, g.hp + c * 32);
            } else {
              tma_load_2d(dst + (uint32_t)kb * B_KB_BYTES, &tmB, full, kb * 64, c * CH);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ========================================= MMA issuer ============================================
    const uint32_t idesc = make_idesc_f16(T::kFmt, 128, CH, 0, 0);
    uint32_t q = 0, it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      mbar_wait_sleep(smem_u32(&bar_afull), it & 1u);
      for (int c = 0; c < g.n_chunks; ++c, ++q) {
        const uint32_t st = q & 1u, ph = (q >> 1) & 1u;
        mbar_wait_sleep(smem_u32(&bar_bfull[st]), ph);
        mbar_wait_sleep(smem_u32(&bar_tempty[st]), ph ^ 1u);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t b_src = base + g.off_b + st * b_stage_bytes;
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            const uint32_t d = tmem_base + st * (uint32_t)(MT * CH) + (uint32_t)(t * CH);
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
    const int ct = threadIdx.x - 64;                 // 0..NCOMP-1
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may read
    const int slice = (warp - 2) >> 2;               // 16-column slice of a chunk this warp drains
    // stencil mapping: channel group fastest, then the pixel column, then the row band
    const int cg = ct % Cfg::CG;
    const int tx = (ct / Cfg::CG) % TW;
    const int band = (ct / Cfg::CG) / TW;
    auto comp_bar = [](int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NCOMP) : "memory"); };
    uint32_t q = 0, it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      const int b = item / per_img, rr = item % per_img;
      const int x0 = (rr % g.tiles_x) * TW, y0 = (rr / g.tiles_x) * TH;
      // ---- per-pixel LayerNorm statistics of the halo'd tile; pixels outside the image get (0, 0) ----
      mbar_wait(smem_u32(&bar_afull), it & 1u);
      for (int m = ct; m < NPIX; m += NCOMP) {
        const int py = y0 - 1 + m / SW, px = x0 - 1 + m % SW;
        const bool inside = py >= 0 && py < g.H && px >= 0 && px < g.W;
        float2 st = make_float2(0.f, inside ? 1.f : 0.f);       // .y == 0 marks a pixel outside the image
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
        sstats[m] = st;
      }
      comp_bar(1);

      for (int c = 0; c < g.n_chunks; ++c, ++q) {
        const uint32_t st = q & 1u, ph = (q >> 1) & 1u;
        // ---- drain: TMEM -> LayerNorm fold -> fp16 shared-memory tile ----
        mbar_wait(smem_u32(&bar_tfull[st]), ph);
        tc_fence_after();
        // first pre-conv channel of this warp's 16-column slice, and its LayerNorm-fold vectors (same for every pixel)
        const int nb = GATE ? (slice < 2 ? c * 32 + slice * 16 : g.hp + c * 32 + (slice - 2) * 16) : c * CH + slice * 16;
        float sv[16], tv[16];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 s4 = reinterpret_cast<const float4*>(svec + nb)[e], t4 = reinterpret_cast<const float4*>(svec + n_vec + nb)[e];
          sv[4 * e] = s4.x; sv[4 * e + 1] = s4.y; sv[4 * e + 2] = s4.z; sv[4 * e + 3] = s4.w;
          tv[4 * e] = t4.x; tv[4 * e + 1] = t4.y; tv[4 * e + 2] = t4.z; tv[4 * e + 3] = t4.w;
        }
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const int m = t * 128 + quarter * 32 + lane;
          uint32_t acc[16];
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + st * (uint32_t)(MT * CH) + (uint32_t)(t * CH + slice * 16), acc);
          const float2 ps = sstats[m < NPIX ? m : 0];
          tmem_ld_wait();
          if (m < NPIX) {
            uint32_t pk[8];
            if (ps.y != 0.f) {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                pk[e] = pack_f16_sat(fmaf(ps.y, __uint_as_float(acc[2 * e]), fmaf(ps.x, sv[2 * e], tv[2 * e])),
                                     fmaf(ps.y, __uint_as_float(acc[2 * e + 1]), fmaf(ps.x, sv[2 * e + 1], tv[2 * e + 1])));
            } else {                                      // outside the image: the conv's zero padding
#pragma unroll
              for (int e = 0; e < 8; ++e) pk[e] = 0u;
            }
            // fp16 tile: GATE two planes [pixel][32 ch] (64-byte rows), plain one plane [pixel][48 ch] (96-byte rows)
            uint8_t* row = GATE ? conv_buf + (size_t)(slice >> 1) * (MT * 128 * 64) + (size_t)m * 64 + (slice & 1) * 32
                                : conv_buf + (size_t)m * 96 + slice * 32;
            *reinterpret_cast<uint4*>(row) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(row + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[st]));       // accumulator buffer may be overwritten
        comp_bar(2);                                                   // fp16 tile complete

        // ---- stencil from shared memory ----
        if (GATE) {
          const int ch = c * 32 + cg * 4;                              // gated channel of this thread
          uint2 w1[9], w2[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            w1[t] = *reinterpret_cast<const uint2*>(sdw + (size_t)t * n_vec + ch);
            w2[t] = *reinterpret_cast<const uint2*>(sdw + (size_t)t * n_vec + g.hp + ch);
          }
          const int x = x0 + tx;
          const bool ok = ch < g.hp && x < g.W;
          unsigned short* outp = reinterpret_cast<unsigned short*>(g.out) + (size_t)b * g.out_bstride + ch +
                                 ((size_t)(y0 + band * R) * g.W + x) * g.out_pitch;
          const size_t out_row = (size_t)g.W * g.out_pitch;
          const uint8_t* src = conv_buf + (size_t)((band * R) * SW + tx) * 64 + cg * 8;
          // the depthwise bias seeds the fp16 accumulators (zero when the conv has no bias)
          uint32_t b1[2] = {0u, 0u}, b2[2] = {0u, 0u};
          if (g.has_bias) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              b1[i] = pack_f16_sat(sbias[ch + 2 * i], sbias[ch + 2 * i + 1]);
              b2[i] = pack_f16_sat(sbias[g.hp + ch + 2 * i], sbias[g.hp + ch + 2 * i + 1]);
            }
          }
          uint32_t p[3][2], qq[3][2];
#pragma unroll
          for (int r = 0; r < R + 2; ++r) {
            if (r < R) { p[r % 3][0] = b1[0]; p[r % 3][1] = b1[1]; qq[r % 3][0] = b2[0]; qq[r % 3][1] = b2[1]; }
            uint2 v1[3], v2[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              v1[kx] = *reinterpret_cast<const uint2*>(src + (r * SW + kx) * 64);
              v2[kx] = *reinterpret_cast<const uint2*>(src + (r * SW + kx) * 64 + MT * 128 * 64);
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
              const float2 pa = h2_to_f2(p[o % 3][0]), pb = h2_to_f2(p[o % 3][1]);
              const float2 qa = h2_to_f2(qq[o % 3][0]), qb = h2_to_f2(qq[o % 3][1]);
              uint2 ov;
              ov.x = pack2<T>(gelu_erf(pa.x) * qa.x, gelu_erf(pa.y) * qa.y);
              ov.y = pack2<T>(gelu_erf(pb.x) * qb.x, gelu_erf(pb.y) * qb.y);
              if (ok && y0 + band * R + o < g.H) *reinterpret_cast<uint2*>(outp + o * out_row) = ov;
            }
          }
        } else {
          const int ch = c * CH + cg * 8;
          uint4 wt[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) wt[t] = *reinterpret_cast<const uint4*>(sdw + (size_t)t * n_vec + ch);
          const int x = x0 + tx;
          const bool ok = ch < g.n_pre && x < g.W;
          unsigned short* outp = reinterpret_cast<unsigned short*>(g.out) + (size_t)b * g.out_bstride + ch +
                                 ((size_t)(y0 + band * R) * g.W + x) * g.out_pitch;
          const size_t out_row = (size_t)g.W * g.out_pitch;
          const uint8_t* src = conv_buf + (size_t)((band * R) * SW + tx) * 96 + cg * 16;
          uint32_t bb[4] = {0u, 0u, 0u, 0u};             // the depthwise bias seeds the fp16 accumulators
          if (g.has_bias) {
#pragma unroll
            for (int i = 0; i < 4; ++i) bb[i] = pack_f16_sat(sbias[ch + 2 * i], sbias[ch + 2 * i + 1]);
          }
          uint32_t p[3][4];
#pragma unroll
          for (int r = 0; r < R + 2; ++r) {
            if (r < R) { p[r % 3][0] = bb[0]; p[r % 3][1] = bb[1]; p[r % 3][2] = bb[2]; p[r % 3][3] = bb[3]; }
            uint4 v[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) v[kx] = *reinterpret_cast<const uint4*>(src + (r * SW + kx) * 96);
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
              uint4 ov;
              uint32_t* op = &ov.x;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = h2_to_f2(p[o % 3][e]);
                op[e] = pack2<T>(f.x, f.y);
              }
              if (ok && y0 + band * R + o < g.H) *reinterpret_cast<uint4*>(outp + o * out_row) = ov;
            }
          }
        }
        comp_bar(3);                                                   // fp16 tile may be overwritten
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------
struct FwPlan {
  int mt, tw;              // template configuration
  uint32_t smem;
  FwArgs g;
};

static int plan_pwdw(const PirPwDw* d, FwPlan* p) {
  FwArgs& g = p->g;
  g = FwArgs{};
  g.B = d->B; g.H = d->H; g.W = d->W; g.C = d->C;
  const bool gate = d->gate != 0;
  const int ch = gate ? 64 : 48;
  g.n_pre = gate ? 2 * d->N : d->N;
  g.hp = gate ? d->N : 0;
  g.ln_mode = d->ln_mode;
  g.nkb = (d->C + 63) / 64;
  g.n_chunks = gate ? (d->N + 31) / 32 : (d->N + ch - 1) / ch;
  g.n_vec = gate ? g.hp + g.n_chunks * 32 : g.n_chunks * ch;
  g.has_bias = d->dw_bias ? 1 : 0;
  // shared-memory plan: [A: nkb x MT x 16 KB] [B ring: 2 x nkb x CH x 128 B] [fp16 tile] [dw taps] [ln_s|vec_t] [dw bias] [stats]
  for (int mt = 3; mt >= 2; --mt) {
    uint32_t off = (uint32_t)g.nkb * mt * 16384u;
    g.off_b = off; off += 2u * g.nkb * (uint32_t)ch * 128u;
    off = (off + 1023u) & ~1023u;
    g.off_conv = off; off += gate ? 2u * mt * 128u * 64u : (uint32_t)mt * 128u * 96u;
    g.off_dw = off; off += (uint32_t)((9 * g.n_vec * 2 + 15) / 16 * 16);
    g.off_vec = off; off += 2u * g.n_vec * 4u;
    g.off_bias = off; off += (uint32_t)g.n_vec * 4u;
    g.off_stats = off; off += (uint32_t)mt * 128u * 8u;
    if (off + 1024u <= 227u * 1024u - 1024u) {
      p->mt = mt; p->smem = off + 1024u;
      p->tw = (mt == 3 && d->W > 16) ? 32 : 16;
      const int th = mt == 3 ? (p->tw == 32 ? 8 : 16) : 12;
      g.tiles_x = (d->W + p->tw - 1) / p->tw;
      g.tiles_y = (d->H + th - 1) / th;
      g.n_items = g.tiles_x * g.tiles_y * d->B;
      return PIR_OK;
    }
  }
  return PIR_ERR_UNSUPPORTED;
}

template <class T, class Cfg>
static int launch_cfg(const FwPlan& p, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t stream) {
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(pwdw_kernel<T, Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024) != cudaSuccess)
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
  pwdw_kernel<T, Cfg><<<grid, Cfg::NTHREADS, p.smem, stream>>>(tmA, tmB, p.g);
  return pir_check_launch("pir_pwdw");
}

template <class T, bool GATE>
static int launch_pwdw(const PirPwDw* d, cudaStream_t stream) {
  FwPlan p;
  if (plan_pwdw(d, &p) != PIR_OK) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_pwdw: C = %d, N = %d does not fit the shared-memory plan", d->C, d->N);
  p.g.dw_w = d->dw_w; p.g.dw_bias = d->dw_bias; p.g.ln_s = d->ln_s; p.g.vec_t = d->vec_t;
  p.g.out = d->out; p.g.out_pitch = d->out_pitch; p.g.out_bstride = d->out_bstride;
  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB;
  {
    const int th = p.mt == 3 ? (p.tw == 32 ? 8 : 16) : 12;
    const uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->a_pitch * 2, (uint64_t)d->a_pitch * 2 * d->W, (uint64_t)d->a_bstride * 2};
    const uint32_t box[4] = {64, (uint32_t)(p.tw + 2), (uint32_t)(th + 2), 1};
    if (int e = pir_make_tmap(&tmA, dt, 4, d->a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t kpad = (uint64_t)p.g.nkb * 64;
    const uint64_t dims[2] = {kpad, (uint64_t)p.g.n_pre};
    const uint64_t strides[1] = {kpad * 2};
    const uint32_t box[2] = {64, GATE ? 32u : 48u};
    if (int e = pir_make_tmap(&tmB, dt, 2, d->w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  if (p.mt == 3 && p.tw == 32) return launch_cfg<T, FwCfg<3, 32, 4, GATE>>(p, tmA, tmB, stream);
  if (p.mt == 3) return launch_cfg<T, FwCfg<3, 16, 4, GATE>>(p, tmA, tmB, stream);
  return launch_cfg<T, FwCfg<2, 16, 3, GATE>>(p, tmA, tmB, stream);
}

}  // namespace pir

extern "C" int pir_pwdw_supported(int32_t C, int32_t N, int32_t gate) {
  PirPwDw d{};
  d.B = 1; d.H = 64; d.W = 64; d.C = C; d.N = N; d.gate = gate;
  pir::FwPlan p;
  if ((C % 8) || (N % 8) || C <= 0 || N <= 0) return 0;
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
  if (d->dtype == PIR_DTYPE_BF16)
    return d->gate ? pir::launch_pwdw<pir::BF16, true>(d, s) : pir::launch_pwdw<pir::BF16, false>(d, s);
  return d->gate ? pir::launch_pwdw<pir::FP16, true>(d, s) : pir::launch_pwdw<pir::FP16, false>(d, s);
}
