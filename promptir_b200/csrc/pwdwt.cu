// Fused  LayerNorm -> 1x1 conv -> depthwise 3x3 -> GELU gate  (GDFN front half, net/model.py:60-63 + :88-90 + :96-97), second
// generation: CHANNEL-MAJOR accumulators.  pwdw.cu computes D[pixel][channel] = X . W^T, so a thread that reads its TMEM lane owns
// ONE pixel and the 3x3 stencil has to go through a shared-memory round trip (drain, barrier, LDS).  Here the product is
// transposed,
//        D[channel][pixel] = W . X^T        (A = 128 weight rows, K-major;  B = the LayerNormed x tile, K-major, N = pixels)
// so a TMEM lane is a CHANNEL and the columns are the pixels of the halo'd tile in raster order: a thread reads rows of ITS channel
// straight from tensor memory (tcgen05.ld), packs horizontally adjacent pixels into half2 pairs and runs the whole 3x3 stencil, the
// GELU gate and the store from registers -- no shared-memory tile, no LDS, no group barriers.  x1 and x2 of the gate (channels c and
// hidden + c) are two accumulators of the SAME lane (two MMA chains against the two 128-row weight slabs).
//
// Per CTA (persistent, one per SM, 704 threads): warp 0 TMA producer | warp 1 MMA issuer | warps 2-17 stencil | warps 18-21 LayerNorm.
// Item = (image, 12 x 16 output tile); halo'd tile 14 x 18 = 252 pixels (256 rows in shared memory), two tile buffers: the LayerNorm
// warps normalise tile i + 1 in place (as in pwdw.cu) while tile i is multiplied and filtered.
// Per item, per block of 128 gated channels, three sub-units of 6 halo'd rows (4 output rows): N = 112 columns (6 x 18 = 108 used)
// per accumulator, two accumulators (x1 | x2) per TMEM buffer, two buffers (512 columns).  The B operand of a sub-unit is the x
// tile at row offset 72 * third (9 KiB: swizzle-atom aligned), so the overlapping rows are never copied.
// Stencil warp w: lane quarter q = w & 3 (channels 32 q .. 32 q + 31 of the block), column patch s = (w - 2) >> 2 (output columns
// 4 s .. 4 s + 3); per sub-unit a thread produces 4 x 4 gated outputs of one channel.  Whether a (warp, sub-unit) needs the border
// path (t added per in-image pixel, bounds-checked stores) is decided per warp and sub-unit, not per tile.
// Plain form (qkv, no gate): 256-channel blocks (two independent accumulators per lane); the last N % 128 <= 32 channels form ONE
// replicated unit per item (the same 32 weight rows in all four lane quarters against the whole tile, quarter q serving sub-unit q).
// What bounds it: the packed-fp16 FMA pipe of the CUDA cores (one HFMA2 per two cycles per scheduler), see DESIGN.md 3.2b.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kTwThreads = 704;
constexpr int kTwCompute = 512;                                            // stencil warps 2-17
constexpr int kTwLn = 128;                                                 // LayerNorm warps 18-21
constexpr int kTwTW = 16, kTwTH = 12, kTwSW = 18, kTwNPIX = 14 * 18;     // output tile, halo'd width, halo'd pixels
constexpr int kTwN = 112;                                                  // UMMA N of a sub-unit (6 x 18 = 108 columns used)
constexpr int kTwRowsPerThird = 4;

struct TwArgs {
  int B, H, W, C;
  int hp;                  // gate: gated channels (multiple of 128); plain: 0
  int n_rows;              // pre-conv channels = weight rows (gate: 2 hp; plain: N)
  int n_cb;                // channel blocks per item: gate hp / 128 (x1 | x2 of 128 gated channels), plain ceil(N / 256) (two 128-channel halves)
  int split;               // plain: channels >= split go to out2 (channel - split); N when there is one output tensor
  int n_main;              // plain: channels handled in 256-channel blocks (x 3 sub-units); gate: unused
  int n_rep;               // plain: the last n_rep <= 32 channels, replicated in all four lane quarters (one whole-tile unit per item)
  int w_rows_main;         // resident weight rows of the block part (the replicated 4 x 32 rows follow)
  int ln_mode;
  int tiles_x, tiles_y, n_items;
  uint32_t mg_per_img, mg_tiles_x;
  uint32_t off_w, off_tab, off_vt;    // byte offsets from the 1024-aligned base: weight ring, per-channel tap table, t vectors
  const void* dw_w;        // [9][2 hp] fp16
  const float* dw_bias;    // [2 hp] or null
  const float* vec_t;      // [2 hp] or null
  void* out;
  long long out_pitch, out_bstride;
  void* out2;
  long long out2_pitch, out2_bstride;
};

__device__ __forceinline__ uint32_t tw_div(uint32_t n, uint32_t magic) { return magic ? __umulhi(n, magic) : n; }
__device__ __forceinline__ uint32_t tw_pack_sat(float lo, float hi) {
  uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
__device__ __forceinline__ uint32_t tw_fma2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t tw_mul2(uint32_t a, uint32_t b) {
  uint32_t r; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t tw_min2(uint32_t a, uint32_t b) {
  uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t tw_max2(uint32_t a, uint32_t b) {
  uint32_t r; asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t tw_tanh2(uint32_t a) {
  uint32_t r; asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a)); return r;
}
__device__ __forceinline__ uint32_t tw_bcast_lo(uint32_t v) { uint32_t r; asm("prmt.b32 %0, %1, %1, 0x1010;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ uint32_t tw_bcast_hi(uint32_t v) { uint32_t r; asm("prmt.b32 %0, %1, %1, 0x3232;" : "=r"(r) : "r"(v)); return r; }
// gelu(p) * q on packed fp16 pairs: the three-coefficient erf-GELU fit of common.cuh in its tanh form (see pwdw.cu)
__device__ __forceinline__ uint32_t tw_gate2(uint32_t p, uint32_t q) {
  constexpr uint32_t kA = 0x3a613a61u, kB = 0x28bd28bdu, kC = 0x8dc28dc2u, k25 = 0x4e404e40u;
  const uint32_t u = tw_min2(tw_mul2(p, p), k25);
  const uint32_t t = tw_fma2(u, tw_fma2(u, kC, kB), kA);
  const uint32_t th = tw_tanh2(tw_mul2(p, t));
  // q already carries the factor 0.5.  The product saturates at +-65504 instead of overflowing to inf (and to NaN downstream)
  constexpr uint32_t kMax = 0x7bff7bffu, kMin = 0xfbfffbffu;
  return tw_max2(tw_min2(tw_mul2(tw_fma2(p, th, p), q), kMax), kMin);
}
__device__ __forceinline__ float2 tw_h2f2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
#ifndef PIR_TW_PARK
#define PIR_TW_PARK 1
#endif
#ifndef PIR_TW_LAZY_NS
#define PIR_TW_LAZY_NS 400
#endif
// Waits of the producer / issuer / LayerNorm warps (mostly idle) and of the stencil warps on a full accumulator: parked in the
// mbarrier unit (try_wait with a suspend-time hint) instead of polling, so a waiting warp costs no issue slots and wakes when the
// phase completes rather than at the next poll.
__device__ __forceinline__ void tw_wait_backoff(uint32_t bar, uint32_t parity) {
#if PIR_TW_PARK
  mbar_wait_parked(bar, parity);
#else
  while (!mbar_try_wait(bar, parity)) __nanosleep(40);
#endif
}
// latency-tolerant waits (TMA producer on a free tile buffer, LayerNorm warps on an arrived tile): poll rarely
__device__ __forceinline__ void tw_wait_lazy(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(PIR_TW_LAZY_NS);
}
__device__ __forceinline__ void tw_wait_full(uint32_t bar, uint32_t parity) {
#if PIR_TW_PARK
  mbar_wait_parked(bar, parity);
#else
  mbar_wait(bar, parity);
#endif
}
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tw_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}


// One sub-unit of one thread: 6 halo'd rows x 6 columns of its channel's two accumulators (x1 at tcol, x2 at tcol + 128) -> 4 x 4
// gated outputs.  BORDER: the tile touches the image border -- t is added to in-image pixels only (the zero padding of the depthwise
// conv must stay zero) and stores are bounds-checked; interior tiles take the straight-line path.
template <class T, bool GATE32, bool BORDER, int PITCH>
__device__ __forceinline__ void tw_subunit(uint32_t tcol, const uint32_t (&w1)[9], const uint32_t (&w2)[9], uint32_t seed1a, uint32_t seed1b,
                                           uint32_t seed2a, uint32_t seed2b,
                                           float2 tv, uint32_t col_in, int yo, int xo, int H, int W, unsigned short* orow, int pitch_rt,
                                           size_t row_stride) {
  constexpr int SW = kTwSW;
  const int pitch = PITCH ? PITCH : pitch_rt;           // compile-time pixel pitch (dense output): the four pixel stores share one address
  uint32_t a1[3][2], a2[3][2];                             // accumulator rows (mod 3) x 2 pixel pairs, per tensor
  uint32_t r1[8], r2[8];
  tw_ld8(tcol, r1); tw_ld8(tcol + 128u, r2);
#pragma unroll
  for (int ri = 0; ri < 6; ++ri) {
    tmem_ld_wait();
    if (BORDER) {
      const int py = yo - 1 + ri;
      if (py >= 0 && py < H) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          if (col_in & (1u << j)) {
            r1[j] = __float_as_uint(__uint_as_float(r1[j]) + tv.x);
            r2[j] = __float_as_uint(__uint_as_float(r2[j]) + tv.y);
          }
        }
      }
    }
    // pairs of horizontally adjacent pixels: E_j = (v[2j], v[2j+1]), O_j = (v[2j+1], v[2j+2])
    uint32_t e1[3], o1[2], e2[3], o2[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      e1[j] = tw_pack_sat(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
      e2[j] = tw_pack_sat(__uint_as_float(r2[2 * j]), __uint_as_float(r2[2 * j + 1]));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      o1[j] = tw_pack_sat(__uint_as_float(r1[2 * j + 1]), __uint_as_float(r1[2 * j + 2]));
      o2[j] = tw_pack_sat(__uint_as_float(r2[2 * j + 1]), __uint_as_float(r2[2 * j + 2]));
    }
    if (ri < 5) { tw_ld8(tcol + (uint32_t)((ri + 1) * SW), r1); tw_ld8(tcol + 128u + (uint32_t)((ri + 1) * SW), r2); }
    if (ri < 4) { a1[ri % 3][0] = seed1a; a1[ri % 3][1] = seed1b; a2[ri % 3][0] = seed2a; a2[ri % 3][1] = seed2b; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ro = ri - ky;
      if (ro >= 0 && ro < 4) {
        uint32_t* p = a1[ro % 3];
        uint32_t* qq = a2[ro % 3];
        p[0] = tw_fma2(e1[0], w1[ky * 3], p[0]); p[0] = tw_fma2(o1[0], w1[ky * 3 + 1], p[0]); p[0] = tw_fma2(e1[1], w1[ky * 3 + 2], p[0]);
        p[1] = tw_fma2(e1[1], w1[ky * 3], p[1]); p[1] = tw_fma2(o1[1], w1[ky * 3 + 1], p[1]); p[1] = tw_fma2(e1[2], w1[ky * 3 + 2], p[1]);
        qq[0] = tw_fma2(e2[0], w2[ky * 3], qq[0]); qq[0] = tw_fma2(o2[0], w2[ky * 3 + 1], qq[0]); qq[0] = tw_fma2(e2[1], w2[ky * 3 + 2], qq[0]);
        qq[1] = tw_fma2(e2[1], w2[ky * 3], qq[1]); qq[1] = tw_fma2(o2[1], w2[ky * 3 + 1], qq[1]); qq[1] = tw_fma2(e2[2], w2[ky * 3 + 2], qq[1]);
      }
    }
    if (ri >= 2) {
      const int ro = ri - 2;
      uint32_t g0, g1;
      if (GATE32) {
        const float2 pa = tw_h2f2(a1[ro % 3][0]), pb = tw_h2f2(a1[ro % 3][1]);
        const float2 qa = tw_h2f2(a2[ro % 3][0]), qb = tw_h2f2(a2[ro % 3][1]);
        g0 = pack2<T>(gelu_erf(pa.x) * qa.x, gelu_erf(pa.y) * qa.y);
        g1 = pack2<T>(gelu_erf(pb.x) * qb.x, gelu_erf(pb.y) * qb.y);
      } else {
        g0 = tw_gate2(a1[ro % 3][0], a2[ro % 3][0]);
        g1 = tw_gate2(a1[ro % 3][1], a2[ro % 3][1]);
        if (T::kFmt == 1) {                                   // bf16 storage: fp16 pairs -> bf16 pairs
          const float2 fa = tw_h2f2(g0), fb = tw_h2f2(g1);
          g0 = pack2<T>(fa.x, fa.y); g1 = pack2<T>(fb.x, fb.y);
        }
      }
      if (!BORDER) {
        orow[0] = (unsigned short)(g0 & 0xffffu);
        orow[pitch] = (unsigned short)(g0 >> 16);
        orow[2 * pitch] = (unsigned short)(g1 & 0xffffu);
        orow[3 * pitch] = (unsigned short)(g1 >> 16);
      } else if (yo + ro < H) {
        if (xo + 0 < W) orow[0] = (unsigned short)(g0 & 0xffffu);
        if (xo + 1 < W) orow[pitch] = (unsigned short)(g0 >> 16);
        if (xo + 2 < W) orow[2 * pitch] = (unsigned short)(g1 & 0xffffu);
        if (xo + 3 < W) orow[3 * pitch] = (unsigned short)(g1 >> 16);
      }
      orow += row_stride;
      asm volatile("" : "+l"(orow));
    }
  }
}

// Plain (no gate) sub-unit of one thread and ONE accumulator: 6 halo'd rows x 6 columns of its channel -> 4 x 4 outputs of the
// depthwise conv, stored as 16-bit.  Same structure as tw_subunit.
template <class T, bool BORDER, int PITCH = 0>
__device__ __forceinline__ void tw_plain(uint32_t tcol, const uint32_t (&w)[9], uint32_t seeda, uint32_t seedb, float tv, uint32_t col_in, int yo, int xo,
                                         int H, int W, unsigned short* orow, int pitch_rt, size_t row_stride, bool st_ok = true) {
  constexpr int SW = kTwSW;
  const int pitch = PITCH ? PITCH : pitch_rt;           // compile-time pixel pitch: the four pixel stores of a row share one address
  uint32_t a[3][2];
  uint32_t r[8];
  tw_ld8(tcol, r);
#pragma unroll
  for (int ri = 0; ri < 6; ++ri) {
    tmem_ld_wait();
    if (BORDER) {
      const int py = yo - 1 + ri;
      if (py >= 0 && py < H) {
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (col_in & (1u << j)) r[j] = __float_as_uint(__uint_as_float(r[j]) + tv);
      }
    }
    uint32_t e[3], o[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) e[j] = tw_pack_sat(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
#pragma unroll
    for (int j = 0; j < 2; ++j) o[j] = tw_pack_sat(__uint_as_float(r[2 * j + 1]), __uint_as_float(r[2 * j + 2]));
    if (ri < 5) tw_ld8(tcol + (uint32_t)((ri + 1) * SW), r);
    if (ri < 4) { a[ri % 3][0] = seeda; a[ri % 3][1] = seedb; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ro = ri - ky;
      if (ro >= 0 && ro < 4) {
        uint32_t* p = a[ro % 3];
        p[0] = tw_fma2(e[0], w[ky * 3], p[0]); p[0] = tw_fma2(o[0], w[ky * 3 + 1], p[0]); p[0] = tw_fma2(e[1], w[ky * 3 + 2], p[0]);
        p[1] = tw_fma2(e[1], w[ky * 3], p[1]); p[1] = tw_fma2(o[1], w[ky * 3 + 1], p[1]); p[1] = tw_fma2(e[2], w[ky * 3 + 2], p[1]);
      }
    }
    if (ri >= 2) {
      const int ro = ri - 2;
      uint32_t g0 = a[ro % 3][0], g1 = a[ro % 3][1];
      if (T::kFmt == 1) {                                     // bf16 storage: fp16 pairs -> bf16 pairs
        const float2 fa = tw_h2f2(g0), fb = tw_h2f2(g1);
        g0 = pack2<T>(fa.x, fa.y); g1 = pack2<T>(fb.x, fb.y);
      }
      if (!BORDER) {
        if (st_ok) {                                          // lanes without a channel (replicated unit with fewer than 32 channels)
          orow[0] = (unsigned short)(g0 & 0xffffu);
          orow[pitch] = (unsigned short)(g0 >> 16);
          orow[2 * pitch] = (unsigned short)(g1 & 0xffffu);
          orow[3 * pitch] = (unsigned short)(g1 >> 16);
        }
      } else if (yo + ro < H) {
        if (xo + 0 < W) orow[0] = (unsigned short)(g0 & 0xffffu);
        if (xo + 1 < W) orow[pitch] = (unsigned short)(g0 >> 16);
        if (xo + 2 < W) orow[2 * pitch] = (unsigned short)(g1 & 0xffffu);
        if (xo + 3 < W) orow[3 * pitch] = (unsigned short)(g1 >> 16);
      }
      orow += row_stride;
      asm volatile("" : "+l"(orow));
    }
  }
}

// Shared-memory matrix descriptor with the swizzle mode as a parameter (2: 128-byte rows, 4: 64-byte rows); see make_sdesc_sw128
__device__ __forceinline__ uint64_t tw_sdesc(uint32_t sbo_bytes, uint32_t layout) {
  return ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

// Shared-memory plan (bytes from the 1024-aligned base):
//   [x tile: 2 x NKB k-blocks x 256 rows x KBB B] [weights, resident: NKB k-blocks x 2 hp rows x KBB B] [tap table: hp x 80 B] [t: hp x 8 B]
// KBB = bytes of a k-block row: 128 (64 channels, 128-byte swizzle; C = 48) or 64 (32 channels, 64-byte swizzle; C = 96 = 3 x 32, so
// nothing is padded, two x tiles and the WHOLE weight matrix fit next to each other: the weights are fetched once per CTA and the
// next item's tile is loaded and LayerNormed while the current one is multiplied).
template <class T, int NKB, int KBB, bool GATE, bool GATE32, int PITCH>
__global__ void __launch_bounds__(kTwThreads, 1)
pwdwt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWr,
             const TwArgs g) {
  constexpr int TW = kTwTW, TH = kTwTH, SW = kTwSW, NPIX = kTwNPIX;
  constexpr int NA = 2;
  constexpr int KCH = KBB / 2;                                 // channels per k-block
  constexpr uint32_t A_KB = 256 * KBB;                         // one k-block of the halo'd x tile (256 rows)
  constexpr uint32_t X_BYTES = NKB * A_KB;
  constexpr uint32_t SW_LAYOUT = KBB == 128 ? 2u : 4u;
  constexpr uint32_t SBO = 8 * KBB;                            // 8-row swizzle atom
  const int w_rows_main = GATE ? 2 * g.hp : g.w_rows_main;     // resident weight rows (plain: padded to whole 128-row slabs, TMA zero fill)
  const int w_rows = w_rows_main + ((!GATE && g.n_rep) ? 128 : 0);   // + the replicated slab: 4 x (the last n_rep channels, zero padded to 32 rows)
  const uint32_t w_kb_bytes = (uint32_t)w_rows * KBB;          // one k-block of the resident weights

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[NA], bar_aready[NA], bar_aempty[NA];
  __shared__ __align__(8) uint64_t bar_wfull;
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  uint4* stab = reinterpret_cast<uint4*>(base_ptr + g.off_tab);          // [hp][5] x 16 B: w1[9] w2[9] seed1i seed2i bias1 bias2 0 0 | seed1L seed2L seed1R seed2R 0 0 0 0 (fp16)
  float2* svt = reinterpret_cast<float2*>(base_ptr + g.off_vt);          // [hp] (t1, t2)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmWr);
    for (int i = 0; i < NA; ++i) {
      mbar_init(smem_u32(&bar_afull[i]), 1); mbar_init(smem_u32(&bar_aready[i]), kTwLn / 32); mbar_init(smem_u32(&bar_aempty[i]), 1);
    }
    mbar_init(smem_u32(&bar_wfull), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), kTwCompute / 32); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), 512); tmem_relinquish(); }
  // per-CTA tables: taps, accumulator seeds and t of the two channels of every (block, lane) slot.  Slot e = block * 128 + lane:
  // gate: channels (e, hp + e) = x1 | x2 of gated channel e; plain: channels (block * 256 + lane, + 128)
  {
    const unsigned short* src = reinterpret_cast<const unsigned short*>(g.dw_w);
    unsigned short* tab = reinterpret_cast<unsigned short*>(stab);
    const int n_pre = g.n_rows;
    // packed gate: 0.5 p (1 + tanh(.)) q = (p tanh(.) + p) (0.5 q); the 0.5 rides on the x2 branch's taps and bias (exact, a power of two)
    const float qs = (GATE && !GATE32) ? 0.5f : 1.0f;
    const int n_slots = g.n_cb * 128 + ((!GATE && g.n_rep) ? 32 : 0);
    for (int i = threadIdx.x; i < n_slots; i += kTwThreads) {
      // plain: slots past the blocks belong to the replicated unit (one channel per lane, no second accumulator)
      const bool rslot = !GATE && i >= g.n_cb * 128;
      const int c1 = GATE ? i : rslot ? g.n_main + (i - g.n_cb * 128) : (i >> 7) * 256 + (i & 127), c2 = GATE ? g.hp + i : c1 + 128;
      const bool v1 = GATE ? true : rslot ? (i - g.n_cb * 128) < g.n_rep : c1 < g.n_main;
      const bool v2 = GATE ? true : (!rslot && c2 < g.n_main);
      float ws1 = 0.f, ws2 = 0.f, cl1 = 0.f, cl2 = 0.f, cr1 = 0.f, cr2 = 0.f;
      for (int tap = 0; tap < 9; ++tap) {
        const unsigned short a = v1 ? src[(size_t)tap * n_pre + c1] : (unsigned short)0;
        const float bf = v2 ? __half2float(__ushort_as_half(src[(size_t)tap * n_pre + c2])) * qs : 0.f;
        tab[i * 40 + tap] = a; tab[i * 40 + 9 + tap] = __half_as_ushort(__float2half_rn(bf));
        ws1 += __half2float(__ushort_as_half(a)); ws2 += bf;
        if (tap % 3 == 0) { cl1 += __half2float(__ushort_as_half(a)); cl2 += bf; }      // left column of taps (dx = -1)
        if (tap % 3 == 2) { cr1 += __half2float(__ushort_as_half(a)); cr2 += bf; }      // right column (dx = +1)
      }
      const float b1 = (g.dw_bias && v1) ? g.dw_bias[c1] : 0.f, b2 = ((g.dw_bias && v2) ? g.dw_bias[c2] : 0.f) * qs;
      const float t1 = (g.vec_t && v1) ? g.vec_t[c1] : 0.f, t2 = (g.vec_t && v2) ? g.vec_t[c2] : 0.f;
      tab[i * 40 + 18] = __half_as_ushort(__float2half_rn(b1 + t1 * ws1));      // interior tiles: the conv of the constant t is a constant
      tab[i * 40 + 19] = __half_as_ushort(__float2half_rn(b2 + t2 * ws2));
      tab[i * 40 + 20] = __half_as_ushort(__float2half_rn(b1));
      tab[i * 40 + 21] = __half_as_ushort(__float2half_rn(b2));
      tab[i * 40 + 22] = 0; tab[i * 40 + 23] = 0;
      // output columns whose left / right neighbour column lies outside the image (its pixels are zero padding and get no t): the
      // constant loses that column of taps.  With these seeds an x-edge patch runs the straight-line path like an interior one.
      tab[i * 40 + 24] = __half_as_ushort(__float2half_rn(b1 + t1 * (ws1 - cl1)));
      tab[i * 40 + 25] = __half_as_ushort(__float2half_rn(b2 + t2 * (ws2 - cl2)));
      tab[i * 40 + 26] = __half_as_ushort(__float2half_rn(b1 + t1 * (ws1 - cr1)));
      tab[i * 40 + 27] = __half_as_ushort(__float2half_rn(b2 + t2 * (ws2 - cr2)));
      // (slots are 80 bytes apart: with 64 the 16-byte reads of eight consecutive channels fall on two bank groups, a 4-way conflict)
      for (int z = 28; z < 40; ++z) tab[i * 40 + z] = 0;
      svt[i] = make_float2(t1, t2);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int per_img = g.tiles_x * g.tiles_y;
  auto item_geo = [&](int item, int& b, int& x0, int& y0) {
    b = (int)tw_div((uint32_t)item, g.mg_per_img);
    const int rr = item - b * per_img;
    const int ty = (int)tw_div((uint32_t)rr, g.mg_tiles_x);
    x0 = (rr - ty * g.tiles_x) * TW; y0 = ty * TH;
  };

  if (warp == 0) {
    // ========================================= TMA producer ==========================================
    if (elect_one()) {                               // the whole weight matrix, once
      const uint32_t full = smem_u32(&bar_wfull);
      mbar_expect_tx(full, (uint32_t)NKB * w_kb_bytes);
      for (int kb = 0; kb < NKB; ++kb) {
        for (int r0 = 0; r0 < w_rows_main; r0 += 128)
          tma_load_2d(base + g.off_w + (uint32_t)kb * w_kb_bytes + (uint32_t)r0 * KBB, &tmW, full, kb * KCH, r0);
        if (w_rows > w_rows_main)                      // the same 32-row box four times: rows past the tensor are zero filled
          for (int j = 0; j < 4; ++j)
            tma_load_2d(base + g.off_w + (uint32_t)kb * w_kb_bytes + (uint32_t)(w_rows_main + 32 * j) * KBB, &tmWr, full, kb * KCH, g.n_main);
      }
    }
    __syncwarp();
    uint32_t it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      int b, x0, y0;
      item_geo(item, b, x0, y0);
      const uint32_t ab = it % NA;
      tw_wait_lazy(smem_u32(&bar_aempty[ab]), ((it / NA) & 1u) ^ 1u);
      if (elect_one()) {
        const uint32_t full = smem_u32(&bar_afull[ab]);
        mbar_expect_tx(full, (uint32_t)NKB * (uint32_t)(NPIX * KBB));
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) tma_load_4d(base + ab * X_BYTES + (uint32_t)kb * A_KB, &tmA, full, kb * KCH, x0 - 1, y0 - 1, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ========================================= MMA issuer ============================================
    const uint64_t desc_hi = tw_sdesc(SBO, SW_LAYOUT);
    const uint32_t idesc = make_idesc_f16(T::kFmt, 128, kTwN, 0, 0);
    tw_wait_backoff(smem_u32(&bar_wfull), 0);
    uint32_t it = 0, tq = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      const uint32_t ab = it % NA;
      tw_wait_backoff(smem_u32(&bar_aready[ab]), (it / NA) & 1u);    // a spinning issuer warp takes issue slots from the compute warps of its scheduler
      tc_fence_after();
      for (int cb = 0; cb < g.n_cb; ++cb) {
        for (int third = 0; third < 3; ++third, ++tq) {
          const uint32_t tb = tq & 1u;
          tw_wait_backoff(smem_u32(&bar_tempty[tb]), ((tq >> 1) & 1u) ^ 1u);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d1 = tmem_base + tb * 256u, d2 = d1 + 128u;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
              const int rem = g.C - kb * KCH;
              const int ksteps = rem >= KCH ? KCH / 16 : (rem + 15) >> 4;
              const uint32_t x_lo = (base + ab * X_BYTES + (uint32_t)kb * A_KB + (uint32_t)third * (72u * KBB)) >> 4;
              const int row1 = GATE ? cb * 128 : cb * 256, row2 = GATE ? g.hp + cb * 128 : cb * 256 + 128;
              const uint32_t w_lo = (base + g.off_w + (uint32_t)kb * w_kb_bytes) >> 4;
              const uint64_t xd = desc_hi | (uint64_t)(x_lo & 0x3fffu);
              const uint64_t w1d = desc_hi | (uint64_t)((w_lo + (((uint32_t)row1 * KBB) >> 4)) & 0x3fffu);
              const uint64_t w2d = desc_hi | (uint64_t)((w_lo + (((uint32_t)row2 * KBB) >> 4)) & 0x3fffu);
              const bool second = GATE || row2 < g.n_main;         // plain: the last block may have no second half
              for (int k = 0; k < ksteps; ++k) {
                const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                umma_f16(d1, w1d + (uint64_t)(2 * k), xd + (uint64_t)(2 * k), idesc, acc);
                if (second) umma_f16(d2, w2d + (uint64_t)(2 * k), xd + (uint64_t)(2 * k), idesc, acc);
              }
            }
            umma_commit(smem_u32(&bar_tfull[tb]));
            if (third == 2 && cb == g.n_cb - 1 && (GATE || g.n_rep == 0)) umma_commit(smem_u32(&bar_aempty[ab]));
          }
          __syncwarp();
        }
      }
      if (!GATE && g.n_rep) {
        // replicated unit: the last <= 32 channels in all four lane quarters against the WHOLE halo'd tile (N = 256 columns of one
        // buffer); lane quarter q then serves sub-unit q, so the odd channels cost one accumulator pass instead of three
        const uint32_t tb = tq & 1u;
        tw_wait_backoff(smem_u32(&bar_tempty[tb]), ((tq >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc_r = make_idesc_f16(T::kFmt, 128, 256, 0, 0);
          const uint32_t d1 = tmem_base + tb * 256u;
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            const int rem = g.C - kb * KCH;
            const int ksteps = rem >= KCH ? KCH / 16 : (rem + 15) >> 4;
            const uint32_t x_lo = (base + ab * X_BYTES + (uint32_t)kb * A_KB) >> 4;
            const uint32_t w_lo = (base + g.off_w + (uint32_t)kb * w_kb_bytes + (uint32_t)w_rows_main * KBB) >> 4;
            const uint64_t xd = desc_hi | (uint64_t)(x_lo & 0x3fffu);
            const uint64_t wd = desc_hi | (uint64_t)(w_lo & 0x3fffu);
            for (int k = 0; k < ksteps; ++k) umma_f16(d1, wd + (uint64_t)(2 * k), xd + (uint64_t)(2 * k), idesc_r, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_tfull[tb]));
          umma_commit(smem_u32(&bar_aempty[ab]));
        }
        __syncwarp();
        ++tq;
      }
    }
  } else if (warp >= 2 + kTwCompute / 32) {
    // ============================ LayerNorm warps: normalise the halo'd x tile in place ============================
    // Four warps of their own instead of a phase of the stencil warps: their LDS / fp32-FMA / shuffle stream fills issue slots the
    // stencil warps (bound by the fp16 FMA pipe) leave empty, and the stencil warps never stop for a tile's LayerNorm.
    const int ct = threadIdx.x - (64 + kTwCompute);  // 0..127
    constexpr int LNR = (NPIX * 4 + kTwLn - 1) / kTwLn;
    const float inv_k = 1.0f / (float)g.C;
    auto layernorm_tile = [&](int x0, int y0, uint32_t itn) {
      const uint32_t ab = itn % NA;
      uint8_t* a_tile = base_ptr + (size_t)ab * X_BYTES;
      tw_wait_lazy(smem_u32(&bar_afull[ab]), (itn / NA) & 1u);        // sleeps between polls: these warps run a tile ahead and mostly wait
      if (g.ln_mode) {
        // Two tasks per thread at a time, their rows held in registers between the statistics and the normalisation: the loads of
        // both are in flight together (in place, the compiler cannot move a task's loads above the previous task's stores) and
        // every value is read from shared memory once.  A LayerNorm warp is latency bound (~0.1 IPC); with one task at a time the
        // four warps needed ~12 K cycles per tile and paced the plain (qkv) kernel and the C = 48 gate.
        constexpr int CPT = KBB / 64;                      // 16-byte chunks of a k-block row per thread
        constexpr int NV = NKB * CPT;
        constexpr int LNP = KBB == 128 ? 4 : 2;              // tasks in flight per thread (C = 96: three 16-byte chunks each; four would spill)
#pragma unroll 1
        for (int r = 0; r < LNR; r += LNP) {
          uint4 v[LNP][NV];
          int mrow[LNP], part_[LNP];
          bool act[LNP];
#pragma unroll
          for (int u = 0; u < LNP; ++u) {
            const int task = (r + u) * kTwLn + ct;
            const int m = task >> 2, part = task & 3;
            const int my = m / SW;
            const int py = y0 - 1 + my, px = x0 - 1 + (m - my * SW);
            mrow[u] = m; part_[u] = part;
            act[u] = (r + u) < LNR && m < NPIX && py >= 0 && py < g.H && px >= 0 && px < g.W;
            // a thread owns CPT physical 16-byte chunks of every k-block row: 2 part + (e ^ (m & 1)) of 8 (128-byte rows; the row parity
            // keeps the quarter-warp's LDS.128 conflict free) or chunk `part` of 4 (64-byte rows)
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
              for (int e = 0; e < CPT; ++e) {
                const int j = KBB == 128 ? 2 * part + (e ^ (m & 1)) : part;
                v[u][kb * CPT + e] = act[u] ? *reinterpret_cast<const uint4*>(a_tile + (size_t)kb * A_KB + (size_t)m * KBB + (j << 4))
                                            : make_uint4(0u, 0u, 0u, 0u);
              }
          }
          float rstd[LNP], shift[LNP];
#pragma unroll
          for (int u = 0; u < LNP; ++u) {
            float s1 = 0.f, s2 = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
              const uint32_t w4[4] = {v[u][q].x, v[u][q].y, v[u][q].z, v[u][q].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const unsigned short lo = lo16(w4[i]), hi = hi16(w4[i]);
                s1 = fma16<T>(lo, T::kOne, s1);
                s1b = fma16<T>(hi, T::kOne, s1b);
                s2 = fma16<T>(lo, lo, s2);
                s2b = fma16<T>(hi, hi, s2b);
              }
            }
            s1 += s1b; s2 += s2b;
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
            const float mu = s1 * inv_k;
            rstd[u] = rsqrtf(fmaxf(fmaf(s2, inv_k, -mu * mu), 0.f) + 1e-5f);
            shift[u] = g.ln_mode == 2 ? 0.f : -rstd[u] * mu;
          }
#pragma unroll
          for (int u = 0; u < LNP; ++u) {
            if (!act[u]) continue;
            const int m = mrow[u], part = part_[u];
            const int sw = KBB == 128 ? (m & 7) : ((m >> 1) & 3);               // logical chunk = physical chunk ^ sw
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
              const int valid = g.C - kb * KCH;                                 // the zero padding above C (last k-block) must stay zero
#pragma unroll
              for (int e = 0; e < CPT; ++e) {
                const int j = KBB == 128 ? 2 * part + (e ^ (m & 1)) : part;
                if ((j ^ sw) * 8 < valid) {
                  uint4 o = v[u][kb * CPT + e];
                  uint32_t* w4 = &o.x;
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    w4[i] = pack2<T>(fmaf(unpack_lo<T>(w4[i]), rstd[u], shift[u]), fmaf(unpack_hi<T>(w4[i]), rstd[u], shift[u]));
                  *reinterpret_cast<uint4*>(a_tile + (size_t)kb * A_KB + (size_t)m * KBB + (j << 4)) = o;
                }
              }
            }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_aready[ab]));
    };
    uint32_t it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      int b, x0, y0;
      item_geo(item, b, x0, y0);
      layernorm_tile(x0, y0, it);
    }
  } else {
    // ============================ stencil + gate + store from TMEM ============================
    const int q = warp & 3;                          // TMEM lane quarter of this warp
    const int s = (warp - 2) >> 2;                   // column patch: output columns 4 s .. 4 s + 3
    uint32_t tq = 0;
    int b = 0, x0 = 0, y0 = 0;
    if ((int)blockIdx.x < g.n_items) item_geo(blockIdx.x, b, x0, y0);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(4 * s);
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      // Border handling is decided per (warp, sub-unit), not per tile: a warp's column patch needs its six input columns inside the
      // image, a sub-unit its six input rows.  On a 256 x 256 image 20 % of the tiles touch the border but only 6 % of the
      // (warp, sub-unit) pairs do, and the sub-units of the last tile row that lie wholly below the image are skipped.
      const bool cols_ok = x0 - 1 + 4 * s >= 0 && x0 + 4 * s + 4 < g.W;
      // x-edge patches: exactly the left-most (right-most) input column is outside the image and all four output columns are inside.
      // They would make ONE of the four column patches 1.5x slower than the other three for every sub-unit of an edge tile, and the
      // other twelve warps wait for it at every accumulator hand-over; edge seeds (see the table) put them on the straight-line path.
      const bool left_edge = x0 - 1 + 4 * s == -1 && x0 + 4 * s + 4 < g.W;
      const bool right_edge = x0 + 4 * s + 4 == g.W && x0 - 1 + 4 * s >= 0;
      // validity of this thread's six input columns (halo'd columns 4 s .. 4 s + 5) and its four output columns
      uint32_t col_in = 0;
#pragma unroll
      for (int j = 0; j < 6; ++j) { const int px = x0 - 1 + 4 * s + j; if (px >= 0 && px < g.W) col_in |= 1u << j; }
      const int xo = x0 + 4 * s;                       // first output column
      unsigned short* out_img = reinterpret_cast<unsigned short*>(g.out) + (size_t)b * g.out_bstride;
      const int pitch = (int)g.out_pitch;
      const size_t row_stride = (size_t)g.W * g.out_pitch;

      for (int cb = 0; cb < g.n_cb; ++cb) {
        const int ch = cb * 128 + q * 32 + lane;       // table slot of this thread = its gated channel (gate)
        // taps (packed pairs), seeds
        const uint4 ta = stab[ch * 5], tb4 = stab[ch * 5 + 1], tc = stab[ch * 5 + 2];
        // halves: ta = w1[0..7]; tb4 = w1[8] w2[0..6]; tc = w2[7] w2[8] s1i s2i b1 b2 0 0
        uint32_t w1[9], w2[9];
        w1[0] = tw_bcast_lo(ta.x); w1[1] = tw_bcast_hi(ta.x); w1[2] = tw_bcast_lo(ta.y); w1[3] = tw_bcast_hi(ta.y);
        w1[4] = tw_bcast_lo(ta.z); w1[5] = tw_bcast_hi(ta.z); w1[6] = tw_bcast_lo(ta.w); w1[7] = tw_bcast_hi(ta.w);
        w1[8] = tw_bcast_lo(tb4.x);
        w2[0] = tw_bcast_hi(tb4.x); w2[1] = tw_bcast_lo(tb4.y); w2[2] = tw_bcast_hi(tb4.y); w2[3] = tw_bcast_lo(tb4.z);
        w2[4] = tw_bcast_hi(tb4.z); w2[5] = tw_bcast_lo(tb4.w); w2[6] = tw_bcast_hi(tb4.w); w2[7] = tw_bcast_lo(tc.x);
        w2[8] = tw_bcast_hi(tc.x);
        const float2 tv = svt[ch];
        const bool add_t = g.vec_t != nullptr;

        // running output pointer(s) of this thread: (first output row of the sub-unit, first output column, channel)
        unsigned short* orow = out_img + ((size_t)y0 * g.W + xo) * g.out_pitch + ch;
        unsigned short* orow2 = nullptr;
        int pitch2 = 0;
        size_t row_stride2 = 0;
        bool va = true, vb = true;
        if (!GATE) {
          // plain: two independent channels per thread, each routed to `out` (below split) or `out2`
          const int ca = cb * 256 + q * 32 + lane, cbn = ca + 128;
          va = ca < g.n_main; vb = cbn < g.n_main;
          unsigned short* img2 = reinterpret_cast<unsigned short*>(g.out2) + (size_t)b * g.out2_bstride;
          const size_t pix = (size_t)y0 * g.W + xo;
          orow = ca < g.split ? out_img + pix * g.out_pitch + ca : img2 + pix * g.out2_pitch + (ca - g.split);
          orow2 = cbn < g.split ? out_img + pix * g.out_pitch + cbn : img2 + pix * g.out2_pitch + (cbn - g.split);
          pitch2 = cbn < g.split ? (int)g.out_pitch : (int)g.out2_pitch;
          row_stride2 = (size_t)g.W * pitch2;
        }
        const int pitch1 = (GATE || cb * 256 + q * 32 + lane < g.split) ? pitch : (int)g.out2_pitch;
        const size_t row_stride1 = GATE ? row_stride : (size_t)g.W * pitch1;
        const bool any_a = GATE || __any_sync(0xffffffffu, va), any_b = GATE || __any_sync(0xffffffffu, vb);
        const bool all_a = GATE || __all_sync(0xffffffffu, va), all_b = GATE || __all_sync(0xffffffffu, vb);   // tcgen05.ld is warp-collective: branch per warp only
        for (int third = 0; third < 3; ++third, ++tq) {
          const uint32_t tb = tq & 1u;
          tw_wait_full(smem_u32(&bar_tfull[tb]), (tq >> 1) & 1u);
          tc_fence_after();
          const uint32_t tcol = t_lane + tb * 256u;
          const int yo = y0 + third * kTwRowsPerThird;            // first output row of the sub-unit
          const bool rows_ok = yo >= 1 && yo + 4 < g.H;
          const bool xl = rows_ok && left_edge, xr = rows_ok && right_edge;
          const bool interior = (cols_ok && rows_ok) || xl || xr;   // straight-line path: no per-pixel t, no bounds checks
          // interior: the conv of the constant t is the constant t * sum(taps) and rides on the seed; border: t is added per in-image pixel
          const uint32_t seed1 = interior ? tw_bcast_lo(tc.y) : tw_bcast_lo(tc.z);
          const uint32_t seed2 = interior ? tw_bcast_hi(tc.y) : tw_bcast_hi(tc.z);
          uint32_t seed1a = seed1, seed1b = seed1, seed2a = seed2, seed2b = seed2;   // per accumulator: output pairs (0, 1) and (2, 3)
          if (xl || xr) {                                           // rare: fetched here so nothing stays live across the loop
            const uint4 td = stab[ch * 5 + 3];
            if (xl) { seed1a = __byte_perm(td.x, tc.y, 0x5410); seed2a = __byte_perm(td.x, tc.y, 0x7632); }    // (edge, interior)
            if (xr) { seed1b = __byte_perm(tc.y, td.y, 0x5410); seed2b = __byte_perm(tc.y, td.y, 0x7632); }    // (interior, edge)
          }
          if (yo >= g.H) {
            // nothing to store (last tile row of an image whose height is not a multiple of the tile): release the accumulators
          } else if (GATE) {
            if (interior) tw_subunit<T, GATE32, false, PITCH>(tcol, w1, w2, seed1a, seed1b, seed2a, seed2b, tv, 0u, yo, xo, g.H, g.W, orow, pitch, row_stride);
            else tw_subunit<T, GATE32, true, PITCH>(tcol, w1, w2, seed1a, seed1b, seed2a, seed2b, add_t ? tv : make_float2(0.f, 0.f), col_in, yo, xo, g.H, g.W, orow, pitch, row_stride);
          } else {
            // channels past the end of the tensor (last block) are skipped per warp; a partially valid warp masks its stores by
            // pretending the row is outside the image (H = 0 on the border path)
            if (any_a) {
              if (interior && all_a) {
                // the networks' two output tensors have pixel pitches PITCH (q|k) and PITCH / 2 (v), a warp's channels go to one of them
                if constexpr (PITCH != 0) {
                  if (cb * 256 + q * 32 < g.split) tw_plain<T, false, PITCH>(tcol, w1, seed1a, seed1b, 0.f, 0u, yo, xo, g.H, g.W, orow, pitch1, row_stride1);
                  else tw_plain<T, false, PITCH / 2>(tcol, w1, seed1a, seed1b, 0.f, 0u, yo, xo, g.H, g.W, orow, pitch1, row_stride1);
                } else {
                  tw_plain<T, false>(tcol, w1, seed1a, seed1b, 0.f, 0u, yo, xo, g.H, g.W, orow, pitch1, row_stride1);
                }
              }
              else tw_plain<T, true>(tcol, w1, seed1a, seed1b, (add_t && !interior) ? tv.x : 0.f, interior ? 0x3fu : col_in, yo, xo, va ? g.H : 0, g.W, orow, pitch1, row_stride1);
            }
            if (any_b) {
              if (interior && all_b) {
                if constexpr (PITCH != 0) {
                  if (cb * 256 + 128 + q * 32 < g.split) tw_plain<T, false, PITCH>(tcol + 128u, w2, seed2a, seed2b, 0.f, 0u, yo, xo, g.H, g.W, orow2, pitch2, row_stride2);
                  else tw_plain<T, false, PITCH / 2>(tcol + 128u, w2, seed2a, seed2b, 0.f, 0u, yo, xo, g.H, g.W, orow2, pitch2, row_stride2);
                } else {
                  tw_plain<T, false>(tcol + 128u, w2, seed2a, seed2b, 0.f, 0u, yo, xo, g.H, g.W, orow2, pitch2, row_stride2);
                }
              }
              else tw_plain<T, true>(tcol + 128u, w2, seed2a, seed2b, (add_t && !interior) ? tv.y : 0.f, interior ? 0x3fu : col_in, yo, xo, vb ? g.H : 0, g.W, orow2, pitch2, row_stride2);
            }
          }
          if (GATE) orow += 4 * row_stride;
          else { orow += 4 * row_stride1; orow2 += 4 * row_stride2; }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[tb]));
        }
      }
      if (!GATE && g.n_rep) {
        // replicated unit: warp (q, s) runs sub-unit q of the last n_rep channels (lane = channel), quarter 3 has nothing to do
        const uint32_t tb = tq & 1u;
        tw_wait_full(smem_u32(&bar_tfull[tb]), (tq >> 1) & 1u);
        tc_fence_after();
        const int yo = y0 + q * kTwRowsPerThird;
        if (q < 3 && yo < g.H) {
          const int slot = g.n_cb * 128 + lane;
          const uint4 ta = stab[slot * 5], tb4 = stab[slot * 5 + 1], tc = stab[slot * 5 + 2], td = stab[slot * 5 + 3];
          uint32_t w1[9];
          w1[0] = tw_bcast_lo(ta.x); w1[1] = tw_bcast_hi(ta.x); w1[2] = tw_bcast_lo(ta.y); w1[3] = tw_bcast_hi(ta.y);
          w1[4] = tw_bcast_lo(ta.z); w1[5] = tw_bcast_hi(ta.z); w1[6] = tw_bcast_lo(ta.w); w1[7] = tw_bcast_hi(ta.w);
          w1[8] = tw_bcast_lo(tb4.x);
          const int c = g.n_main + lane;
          const bool v = lane < g.n_rep;
          const bool rows_ok = yo >= 1 && yo + 4 < g.H;
          const bool xl = rows_ok && left_edge, xr = rows_ok && right_edge;
          const bool interior = (cols_ok && rows_ok) || xl || xr;
          uint32_t seeda, seedb;
          seeda = seedb = interior ? tw_bcast_lo(tc.y) : tw_bcast_lo(tc.z);
          if (xl) seeda = __byte_perm(td.x, tc.y, 0x5410);
          if (xr) seedb = __byte_perm(tc.y, td.y, 0x5410);
          const float tvx = svt[slot].x;
          unsigned short* img2 = reinterpret_cast<unsigned short*>(g.out2) + (size_t)b * g.out2_bstride;
          const size_t pix = (size_t)yo * g.W + xo;
          const int pr = c < g.split ? pitch : (int)g.out2_pitch;
          unsigned short* op = c < g.split ? out_img + pix * g.out_pitch + c : img2 + pix * g.out2_pitch + (c - g.split);
          const uint32_t tcol = t_lane + tb * 256u + (uint32_t)(q * 4 * SW);
          if (interior) {
            // dense outputs: the replicated channels (the last of the 3 C) all belong to v
            if constexpr (PITCH != 0) tw_plain<T, false, PITCH / 2>(tcol, w1, seeda, seedb, 0.f, 0u, yo, xo, g.H, g.W, op, pr, (size_t)g.W * pr, v);
            else tw_plain<T, false>(tcol, w1, seeda, seedb, 0.f, 0u, yo, xo, g.H, g.W, op, pr, (size_t)g.W * pr, v);
          }
          else tw_plain<T, true>(tcol, w1, seeda, seedb, g.vec_t != nullptr ? tvx : 0.f, col_in, yo, xo, v ? g.H : 0, g.W, op, pr, (size_t)g.W * pr);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[tb]));
        ++tq;
      }
      if (item + (int)gridDim.x < g.n_items) item_geo(item + gridDim.x, b, x0, y0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------
template <class T, int NKB, int KBB, bool GATE, bool GATE32, int PITCH>
static int tw_launch(const TwArgs& g, uint32_t smem, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmWr, cudaStream_t stream) {
  if (!pir_smem_attr_once(reinterpret_cast<const void*>(pwdwt_kernel<T, NKB, KBB, GATE, GATE32, PITCH>), (int)(227 * 1024 - 1024), "pir_pwdw")) return PIR_ERR_CUDA;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int grid = g.n_items < num_sms ? g.n_items : num_sms;
  pwdwt_kernel<T, NKB, KBB, GATE, GATE32, PITCH><<<dim3(grid), dim3(kTwThreads), smem, stream>>>(tmA, tmW, tmWr, g);
  return pir_check_launch("pir_pwdw (channel-major)");
}

// k-block geometry: C <= 64 -> one 64-channel block (128-byte rows); C = 96 -> three 32-channel blocks (64-byte rows, no padding)
static bool tw_shape(int C, int* nkb, int* kbb) {
  if (C <= 64) { *nkb = 1; *kbb = 128; return true; }
  if (C == 96) { *nkb = 3; *kbb = 64; return true; }
  return false;
}

// C <= 64 or C == 96; gate: hidden a multiple of 128; the plan (two x tiles + ALL weights + tables) fits shared memory
// plain form: the last N % 128 channels, when there are at most 32 of them (qkv: 144 = 128 + 16, 288 = 256 + 32), do not get a
// 128-row block of their own (three of four lane quarters would idle through three sub-units) but ONE replicated unit per item
static void tw_plain_split(int N, int* n_main, int* n_rep) {
  const int r = N % 128;
  *n_rep = (r > 0 && r <= 32 && N > 128) ? r : 0;
  *n_main = N - *n_rep;
}

static uint32_t tw_smem(const PirPwDw* d, int nkb, int kbb, TwArgs* g) {
  int n_main = d->N, n_rep = 0;
  if (!d->gate) tw_plain_split(d->N, &n_main, &n_rep);
  const int n_cb = d->gate ? d->N / 128 : (n_main + 255) / 256;
  const int w_rows_main = d->gate ? 2 * d->N : (n_main + 127) / 128 * 128;
  const int w_rows = w_rows_main + (n_rep ? 128 : 0);
  const uint32_t slots = (uint32_t)n_cb * 128u + (n_rep ? 32u : 0u);
  if (g) { g->n_main = n_main; g->n_rep = n_rep; g->n_cb = n_cb; g->w_rows_main = w_rows_main; }
  uint32_t off = 2u * nkb * 256u * kbb;
  if (g) g->off_w = off;
  off += (uint32_t)nkb * w_rows * kbb;
  if (g) g->off_tab = off;
  off += slots * 80u;
  if (g) g->off_vt = off;
  off += slots * 8u;
  return off + 1024u;
}

bool pwdwt_supported(const PirPwDw* d) {
  // bit 0: gate, bit 1: plain.  Default: gate only -- measured on B200 (B = 16, 256 x 256): the plain form leaves lane quarters idle
  // in the last channel block (288 = 256 + 32, 144 = 128 + 16 channels) and runs at 495 / 351 us against 398 / 233 us of pwdw.cu
  static const int mode = [] { const char* e = getenv("PIR_PWDW_T"); return e ? atoi(e) : 3; }();
  if (!(mode & (d->gate ? 1 : 2))) return false;
  int nkb, kbb;
  if (!tw_shape(d->C, &nkb, &kbb)) return false;
  if (d->gate && d->N % 128 != 0) return false;
  return tw_smem(d, nkb, kbb, nullptr) <= 227u * 1024u - 1024u;
}

template <class T>
static int tw_run(const PirPwDw* d, cudaStream_t stream) {
  TwArgs g{};
  const bool gate = d->gate != 0;
  g.B = d->B; g.H = d->H; g.W = d->W; g.C = d->C;
  g.hp = gate ? d->N : 0; g.n_rows = gate ? 2 * d->N : d->N;
  g.ln_mode = d->ln_mode;
  g.tiles_x = (d->W + kTwTW - 1) / kTwTW; g.tiles_y = (d->H + kTwTH - 1) / kTwTH;
  g.n_items = g.tiles_x * g.tiles_y * d->B;
  {
    auto magic = [](uint32_t dv) { return dv <= 1 ? 0u : (uint32_t)((0x100000000ull + dv - 1) / dv); };
    const uint64_t per_img = (uint64_t)g.tiles_x * g.tiles_y;
    if ((uint64_t)g.n_items * per_img >= 0x100000000ull) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_pwdw: too many tiles");
    g.mg_per_img = magic((uint32_t)per_img); g.mg_tiles_x = magic((uint32_t)g.tiles_x);
  }
  int nkb = 0, kbb = 0;
  tw_shape(d->C, &nkb, &kbb);
  const int kch = kbb / 2;
  const uint32_t smem = tw_smem(d, nkb, kbb, &g);
  if (smem > 227u * 1024u - 1024u) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_pwdw: shared-memory plan does not fit");
  g.dw_w = d->dw_w; g.dw_bias = d->dw_bias; g.vec_t = d->vec_t;
  g.out = d->out; g.out_pitch = d->out_pitch; g.out_bstride = d->out_bstride;
  const bool two = !gate && d->out2 != nullptr && d->split > 0 && d->split < d->N;
  g.split = two ? d->split : d->N;
  g.out2 = two ? d->out2 : d->out; g.out2_pitch = two ? d->out2_pitch : d->out_pitch; g.out2_bstride = two ? d->out2_bstride : d->out_bstride;
  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = kbb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->a_pitch * 2, (uint64_t)d->a_pitch * 2 * d->W, (uint64_t)d->a_bstride * 2};
    const uint32_t box[4] = {(uint32_t)kch, (uint32_t)kTwSW, (uint32_t)(kTwTH + 2), 1};
    if (int e = pir_make_tmap(&tmA, dt, 4, d->a, dims, strides, box, sw)) return e;
  }
  {
    const uint64_t kpad = (uint64_t)((d->C + 63) / 64) * 64;           // row length of the packed weights (packing.kpad_of)
    const uint64_t dims[2] = {kpad, (uint64_t)g.n_rows};
    const uint64_t strides[1] = {kpad * 2};
    const uint32_t box[2] = {(uint32_t)kch, 128};
    if (int e = pir_make_tmap(&tmW, dt, 2, d->w, dims, strides, box, sw)) return e;
  }
  CUtensorMap tmWr = tmW;
  if (g.n_rep) {                                                        // 32-row boxes for the replicated slab
    const uint64_t kpad = (uint64_t)((d->C + 63) / 64) * 64;
    const uint64_t dims[2] = {kpad, (uint64_t)g.n_rows};
    const uint64_t strides[1] = {kpad * 2};
    const uint32_t box[2] = {(uint32_t)kch, 32};
    if (int e = pir_make_tmap(&tmWr, dt, 2, d->w, dims, strides, box, sw)) return e;
  }
  if (!gate) {
    // dense q|k (pitch 2 C) and v (pitch C) tensors: those pitches are compiled in (store addresses become immediates)
    const bool dense2 = two && d->split == 2 * d->C && d->out_pitch == 2 * d->C && d->out2_pitch == d->C && d->N == 3 * d->C;
    if (nkb == 1) return dense2 && d->C == 48 ? tw_launch<T, 1, 128, false, false, 96>(g, smem, tmA, tmW, tmWr, stream)
                                              : tw_launch<T, 1, 128, false, false, 0>(g, smem, tmA, tmW, tmWr, stream);
    return dense2 && d->C == 96 ? tw_launch<T, 3, 64, false, false, 192>(g, smem, tmA, tmW, tmWr, stream)
                                : tw_launch<T, 3, 64, false, false, 0>(g, smem, tmA, tmW, tmWr, stream);
  }
  // fp32 erf-GELU gate for fp16 storage only on request (PIR_PWDW_GATE32=1, A/B): the packed fp16 gate measured 8.3e-4 against
  // 9.0e-4 max-abs on the cfg2 forward and is 0.75 ms per step faster
  static const bool gate32 = [] { const char* e = getenv("PIR_PWDW_GATE32"); return e && e[0] == '1'; }();
  const bool g32 = T::kFmt == 0 && gate32;
  // the networks write a dense gated tensor (pitch == hidden): those two pitches are compiled in
  if (nkb == 1) {
    if (d->out_pitch == 128) return g32 ? tw_launch<T, 1, 128, true, true, 128>(g, smem, tmA, tmW, tmWr, stream) : tw_launch<T, 1, 128, true, false, 128>(g, smem, tmA, tmW, tmWr, stream);
    return g32 ? tw_launch<T, 1, 128, true, true, 0>(g, smem, tmA, tmW, tmWr, stream) : tw_launch<T, 1, 128, true, false, 0>(g, smem, tmA, tmW, tmWr, stream);
  }
  if (d->out_pitch == 256) return g32 ? tw_launch<T, 3, 64, true, true, 256>(g, smem, tmA, tmW, tmWr, stream) : tw_launch<T, 3, 64, true, false, 256>(g, smem, tmA, tmW, tmWr, stream);
  return g32 ? tw_launch<T, 3, 64, true, true, 0>(g, smem, tmA, tmW, tmWr, stream) : tw_launch<T, 3, 64, true, false, 0>(g, smem, tmA, tmW, tmWr, stream);
}

int pwdwt_run(const PirPwDw* d, cudaStream_t stream) {
  return d->dtype == PIR_DTYPE_BF16 ? tw_run<BF16>(d, stream) : tw_run<FP16>(d, stream);
}

}  // namespace pir
