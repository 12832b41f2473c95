// Fused  LayerNorm -> 1x1 conv -> depthwise 3x3 (-> GELU gate)  for sm_100a: one kernel for the MDTA qkv branch
// (net/model.py:60-63 + :111-112 + :120) and one for the GDFN front half (:60-63 + :88-90 + :96-97).  The widest
// activation of a TransformerBlock (3C resp. 2*hidden channels per pixel) never reaches HBM:
//
//   HBM traffic per pixel:  read C (x 1.3-1.5 for the halo, mostly L2 hits)  +  write 3C (qkv) or hidden (gated)
//   instead of              C + 2*(3C) + 3C                                 resp.  C + 2*(2 hidden) + hidden.
//
// One persistent CTA per SM walks (image, spatial tile) items.  Per item
//   * TMA brings the tile + 1-pixel halo of x (all K channels, 128B-swizzled k-blocks) into shared memory; pixels
//     outside the image are the TMA out-of-bounds zero fill;
//   * the compute warps LayerNorm the tile IN PLACE (per-pixel mean / rstd over the K channels; gamma is folded into
//     the weights, beta into the per-channel vector t), so the GEMM consumes normalised 16-bit rows and zero rows stay
//     zero -- which is exactly the zero padding the depthwise conv wants around the image;
//   * per chunk of 32 pre-conv channels a tcgen05 GEMM (M = 128-pixel row groups, N = 32, fp32 accumulation in TMEM)
//     is issued by one thread into one of two accumulator buffers per group; weights stream through a 4-stage TMA ring;
//   * the 16 compute warps form two groups that take alternate chunks and run half a period apart (while one drains
//     TMEM -- LDTM / F2FP / STS -- the other runs its stencil -- LDS / HFMA2 / MUFU).  A group drains its accumulator
//     buffer into its own shared-memory fp16 tile (+ t[n] inside the image), then runs the 3x3 stencil from it with
//     packed HFMA2 (fp16 inputs, taps and accumulation: the intermediate keeps 11 mantissa bits, more than the bf16
//     tensor it replaces; values saturate at +-65504), applies the erf-GELU gate in fp32 and stores 16-bit NHWC.
//
// Warp roles (576 threads): 0 TMA producer | 1 TMEM alloc + MMA issue | 2-17 compute warps in 2 groups of 8 or, where shared memory
// allows a fourth fp16 tile, 4 groups of 4 (each thread then walks 6 output rows: 8 input rows per 6 outputs instead of 5 per 3).
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

// PIR_PWDW_DBG what-if switches cost a few predicated branches per chunk: compiled in only with -DPIR_PWDW_DEBUG
#ifdef PIR_PWDW_DEBUG
#define FW_DBG(g) ((g).dbg)
#else
#define FW_DBG(g) 0
#endif

namespace pir {

constexpr int kFwThreads = 576;
constexpr int kFwCompute = 512;
constexpr int kFwBStages = 2;               // weight TMA ring (one super-chunk per stage)
constexpr int kFwChunk = 32;                 // pre-conv channels per chunk = UMMA N (gate: 16 of x1 + the matching 16 of x2)

struct FwArgs {
  int B, H, W, C;          // input tensor
  int n_pre;               // pre-conv channels (3C or 2*hp)
  int hp;                  // gate: gated channels (n_pre / 2); plain: 0
  int ln_mode;
  int nkb;                 // k-blocks of 64 input channels
  int n_chunks;
  int n_vec;               // length of the per-channel shared-memory vectors (covers every index a chunk may touch)
  int tiles_x, tiles_y, n_items;
  uint32_t mg_per_img, mg_tiles_x;
  int has_bias;            // depthwise bias present
  int has_t;               // additive per-channel vector present
  int scb;                 // chunks per super-chunk (<= SC of the configuration)
  int rot;                 // rotate the group <-> chunk assignment by the super-chunk index too
  int dbg;                 // diagnostics (PIR_PWDW_DBG): 1 skip the stencil, 2 skip the drain arithmetic, 4 skip the stores, 8 skip LayerNorm, 16 skip TMEM loads
  uint32_t off_b, off_conv, off_dw, off_vec, off_bias;   // byte offsets from the 1024-aligned base
  const void* dw_w;        // [9][n_pre] fp16
  const float* dw_bias;    // [n_pre] or null
  const float* vec_t;      // [n_pre] or null
  void* out;
  long long out_pitch, out_bstride;
  void* out2;              // plain only: channels >= split go here (channel - split); == out / n_pre when there is one tensor
  long long out2_pitch, out2_bstride;
  int split;
};

__device__ __forceinline__ uint32_t pack_f16_sat(float lo, float hi) {
  uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t hmul2_u(uint32_t a, uint32_t b) {
  uint32_t r; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t hmin2_u(uint32_t a, uint32_t b) {
  uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t htanh2_u(uint32_t a) {
  uint32_t r; asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a)); return r;
}
// gelu(p) * q on packed fp16 pairs, for 16-bit outputs with 8 mantissa bits (bf16): the same three-coefficient erf-GELU fit as
// gelu_erf() in its tanh form, 0.5 p (1 + tanh(p (a + b u + c u^2))), u = min(p^2, 25), evaluated in fp16 (MUFU.TANH.F16).
// Its error (~5e-4 * |p q|) stays 4-8x below the bf16 rounding of the result; fp16 outputs take the fp32 path instead.
__device__ __forceinline__ uint32_t gelu_gate_h2(uint32_t p, uint32_t q) {
  constexpr uint32_t kA = 0x3a613a61u, kB = 0x28bd28bdu, kC = 0x8dc28dc2u, k25 = 0x4e404e40u, kHalf = 0x38003800u;   // 0.7975, 0.037006, -3.5152e-4, 25, 0.5
  const uint32_t u = hmin2_u(hmul2_u(p, p), k25);
  const uint32_t t = hfma2_u(u, hfma2_u(u, kC, kB), kA);
  const uint32_t th = htanh2_u(hmul2_u(p, t));
  const uint32_t hx = hmul2_u(p, kHalf);
  // the product saturates at +-65504 instead of overflowing to inf (and to NaN downstream)
  constexpr uint32_t kMax = 0x7bff7bffu, kMin = 0xfbfffbffu;
  uint32_t r = hmin2_u(hmul2_u(hfma2_u(hx, th, hx), q), kMax);
  asm("max.f16x2 %0, %0, %1;" : "+r"(r) : "r"(kMin));
  return r;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

__device__ __forceinline__ uint32_t fw_div(uint32_t n, uint32_t magic) { return magic ? __umulhi(n, magic) : n; }

// Compile-time shape of one kernel variant.  A compute group has 256 threads = 4 channel groups x 64 pixel columns;
// the 64 columns are BANDS = 64 / TW row bands of R rows over a TW-wide tile.
//   GATE : 4 gated channels per thread (16 gated channels per chunk), fp16 tile = two planes of 32-byte rows
//   plain: 8 channels per thread       (32 channels per chunk),      fp16 tile = one plane of 64-byte rows
template <int MT_, int TW_, int R_, bool GATE_, int NA_, int SC_, int NG_ = 2, int NKB_ = 0>
struct FwCfg {
  static constexpr int NKB = NKB_;
  static constexpr int MT = MT_, TW = TW_, R = R_;
  static constexpr int NG = NG_;                             // compute groups: 2 x 8 warps or 4 x 4 warps, one chunk in flight per group
  static constexpr int GT = kFwCompute / NG_;                // threads per group = 4 channel groups x GT/4 pixel columns
  static constexpr bool GATE = GATE_;
  static constexpr int NA = NA_;
  static constexpr int SC = SC_;                             // chunks per super-chunk: one tcgen05.mma covers SC * 32 channels
  static constexpr int SCN = SC * kFwChunk;                  // UMMA N of a full super-chunk                             // x-tile buffers: 2 = the next item's tile is loaded and normalised ahead
  static constexpr int BANDS = GT / 4 / TW;
  static constexpr int TH = R * BANDS;
  static constexpr int SW = TW + 2;
  static constexpr int NPIX = (TH + 2) * SW;
  static constexpr int TMEM_COLS = 2 * MT * SCN <= 256 ? 256 : 512;          // 2 super-chunk buffers x MT x SC x 32 columns
  static_assert(2 * MT * SCN <= 512, "accumulators do not fit TMEM");
  static constexpr uint32_t A_KB_BYTES = MT * 128 * 128;     // one k-block of the halo'd tile
  static constexpr uint32_t B_KB_BYTES = SCN * 128;          // one k-block of a super-chunk's weights
  static constexpr uint32_t CONV_BYTES = MT * 128 * 64;      // fp16 tile of one group (32 channels per pixel)
  static_assert(NPIX <= MT * 128, "halo'd tile does not fit the M tiles");
};

template <class T, class Cfg>
__global__ void __launch_bounds__(kFwThreads, 1)
pwdw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwArgs g) {
  constexpr int MT = Cfg::MT, TW = Cfg::TW, R = Cfg::R, TH = Cfg::TH, SW = Cfg::SW, NPIX = Cfg::NPIX;
  constexpr bool GATE = Cfg::GATE;
  constexpr int NA = Cfg::NA, SC = Cfg::SC, SCN = Cfg::SCN, NG = Cfg::NG, GT = Cfg::GT;
  static_assert(NG == 2 || (NG == 4 && SC == 4), "four groups take one chunk each of a 4-chunk super-chunk");
  constexpr uint32_t A_KB_BYTES = Cfg::A_KB_BYTES, B_KB_BYTES = Cfg::B_KB_BYTES;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[2], bar_aready[2], bar_aempty[2];
  __shared__ __align__(8) uint64_t bar_bfull[kFwBStages], bar_bempty[kFwBStages];
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];          // super-chunk accumulator buffers
  __shared__ uint32_t tmem_base_smem;

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  constexpr int NKB = Cfg::NKB;
  const int nkb = NKB ? NKB : g.nkb;
  const uint32_t b_stage_bytes = (uint32_t)nkb * B_KB_BYTES;
  const unsigned short* sdw = reinterpret_cast<const unsigned short*>(base_ptr + g.off_dw);
  float* svec = reinterpret_cast<float*>(base_ptr + g.off_vec);              // [n_vec] additive vector t
  const uint32_t* sseed = reinterpret_cast<const uint32_t*>(base_ptr + g.off_bias);   // [2][n_chunks * 32] fp16 accumulator seeds
  const int n_vec = g.n_vec;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_afull[i]), 1); mbar_init(smem_u32(&bar_aready[i]), kFwCompute / 32); mbar_init(smem_u32(&bar_aempty[i]), 1);
    }
    for (int i = 0; i < kFwBStages; ++i) { mbar_init(smem_u32(&bar_bfull[i]), 1); mbar_init(smem_u32(&bar_bempty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), kFwCompute / 32); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), Cfg::TMEM_COLS); tmem_relinquish(); }
  // per-CTA constants.  Depthwise taps are regrouped chunk-major so a thread reaches all its taps from ONE base address:
  //   GATE : [chunk][x1 | x2][tap][16 ch]   plain: [chunk][tap][32 ch]     (fp16, 576 bytes per chunk)
  // svec: additive vector t.  Two fp16 accumulator seeds per pre-conv channel (same regrouping, 32 channels per chunk):
  //   sseed[0] = depthwise bias                              (tiles touching the image border: t is added in the drain)
  //   sseed[1] = depthwise bias + t[n] * sum of the 9 taps   (interior tiles: the conv of the constant t is a constant)
  {
    const unsigned short* src = reinterpret_cast<const unsigned short*>(g.dw_w);
    unsigned short* dst = const_cast<unsigned short*>(sdw);
    for (int i = threadIdx.x; i < g.n_chunks * 288; i += kFwThreads) {
      const int c = i / 288, r = i - c * 288;
      int tap, n;
      if (GATE) { const int half = r / 144, rr = r - half * 144; tap = rr >> 4; n = (half ? g.hp : 0) + c * 16 + (rr & 15); if (c * 16 + (rr & 15) >= g.hp) n = g.n_pre; }
      else { tap = r >> 5; n = c * 32 + (r & 31); }
      dst[i] = n < g.n_pre ? src[(size_t)tap * g.n_pre + n] : (unsigned short)0;
    }
    for (int i = threadIdx.x; i < n_vec; i += kFwThreads) svec[i] = (g.vec_t && i < g.n_pre) ? g.vec_t[i] : 0.f;
    unsigned short* seed = reinterpret_cast<unsigned short*>(base_ptr + g.off_bias);
    for (int i = threadIdx.x; i < g.n_chunks * 32; i += kFwThreads) {
      const int c = i >> 5, r = i & 31;
      int n;
      if (GATE) { n = (r >= 16 ? g.hp : 0) + c * 16 + (r & 15); if (c * 16 + (r & 15) >= g.hp) n = g.n_pre; }
      else n = i;
      float bias = 0.f, tsum = 0.f;
      if (n < g.n_pre) {
        if (g.dw_bias) bias = g.dw_bias[n];
        if (g.vec_t) {
          float wsum = 0.f;
          for (int tap = 0; tap < 9; ++tap) wsum += __half2float(__ushort_as_half(src[(size_t)tap * g.n_pre + n]));
          tsum = g.vec_t[n] * wsum;
        }
      }
      seed[i] = __half_as_ushort(__float2half_rn(bias));
      seed[g.n_chunks * 32 + i] = __half_as_ushort(__float2half_rn(bias + tsum));
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();                                    // barrier init, TMEM allocation and the staging of the (static) depthwise taps / seeds above
                                                 // overlapped the previous kernel's tail; x itself is only read from here on
  const int per_img = g.tiles_x * g.tiles_y;
  auto item_geo = [&](int item, int& b, int& x0, int& y0) {
    b = (int)fw_div((uint32_t)item, g.mg_per_img);
    const int rr = item - b * per_img;
    const int ty = (int)fw_div((uint32_t)rr, g.mg_tiles_x);
    x0 = (rr - ty * g.tiles_x) * TW; y0 = ty * TH;
  };
  const int n_super = (g.n_chunks + g.scb - 1) / g.scb;
  const int scb = g.scb;                                                   // chunks per super-chunk
  const uint32_t a_buf_bytes = (uint32_t)nkb * A_KB_BYTES;               // one x-tile buffer

  if (warp == 0) {
    // ========================================= TMA producer ==========================================
    // order: A(0) | A(1) B(0,*) | A(2) B(1,*) | ...   (NA == 2: the next item's tile is requested before this item's weights)
    auto load_a = [&](int item, uint32_t it) {
      int b, x0, y0;
      item_geo(item, b, x0, y0);
      x0 -= 1; y0 -= 1;
      const uint32_t ab = it % NA;
      mbar_wait_sleep(smem_u32(&bar_aempty[ab]), ((it / NA) & 1u) ^ 1u);
      if (elect_one()) {
        const uint32_t full = smem_u32(&bar_afull[ab]);
        mbar_expect_tx(full, (uint32_t)nkb * (uint32_t)(NPIX * 128));
        for (int kb = 0; kb < nkb; ++kb) tma_load_4d(base + ab * a_buf_bytes + (uint32_t)kb * A_KB_BYTES, &tmA, full, kb * 64, x0, y0, b);
      }
      __syncwarp();
    };
    uint32_t q = 0, it = 0;
    if (NA == 2 && (int)blockIdx.x < g.n_items) load_a(blockIdx.x, 0);
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      if (NA == 2) {
        if (item + (int)gridDim.x < g.n_items) load_a(item + gridDim.x, it + 1);
      } else {
        load_a(item, it);
      }
      for (int sc = 0; sc < n_super; ++sc, ++q) {
        const uint32_t st = q % kFwBStages;
        const int nvalid = min(scb, g.n_chunks - sc * scb);
        mbar_wait_sleep(smem_u32(&bar_bempty[st]), ((q / kFwBStages) & 1u) ^ 1u);
        if (elect_one()) {
          const uint32_t full = smem_u32(&bar_bfull[st]);
          const uint32_t dst = base + g.off_b + st * b_stage_bytes;
          mbar_expect_tx(full, (uint32_t)nkb * (uint32_t)nvalid * 4096u);
          for (int kb = 0; kb < nkb; ++kb) {
            for (int j = 0; j < nvalid; ++j) {               // chunk c occupies rows [32 j, 32 j + 32) of the stage
              const int c = sc * scb + j;
              const int r0 = GATE ? c * 16 : c * 32, r1 = GATE ? g.hp + c * 16 : c * 32 + 16;
              tma_load_2d(dst + (uint32_t)kb * B_KB_BYTES + (uint32_t)j * 4096u, &tmB, full, kb * 64, r0);
              tma_load_2d(dst + (uint32_t)kb * B_KB_BYTES + (uint32_t)j * 4096u + 2048u, &tmB, full, kb * 64, r1);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ========================================= MMA issuer ============================================
    // one tcgen05.mma covers a whole super-chunk (SC chunks = SC * 32 accumulator columns per 128-pixel row group): small-N
    // MMAs are latency bound (~110 cycles each whatever N), so N is made as large as TMEM double buffering allows
    const uint64_t desc_hi = make_sdesc_sw128(0, 16, 1024);           // everything but the start address
    uint32_t q = 0, it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      const uint32_t ab = it % NA;
      mbar_wait(smem_u32(&bar_aready[ab]), (it / NA) & 1u);
      tc_fence_after();
      for (int sc = 0; sc < n_super; ++sc, ++q) {
        const uint32_t st = q % kFwBStages, ph = (q / kFwBStages) & 1u;
        const uint32_t sb = q & 1u;
        const int nvalid = min(scb, g.n_chunks - sc * scb);
        const uint32_t idesc = make_idesc_f16(T::kFmt, 128, nvalid * kFwChunk, 0, 0);
        mbar_wait(smem_u32(&bar_bfull[st]), ph);
        mbar_wait(smem_u32(&bar_tempty[sb]), ((q >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_lo = (base + g.off_b + st * b_stage_bytes) >> 4;
          for (int kb = 0; kb < nkb; ++kb) {
            const int rem = g.C - kb * 64;
            const int ksteps = rem >= 64 ? 4 : (rem + 15) >> 4;
            const uint64_t bd = desc_hi | (uint64_t)((b_lo + (uint32_t)kb * (B_KB_BYTES >> 4)) & 0x3fffu);
            for (int k = 0; k < ksteps; ++k) {         // 32 bytes (16 elements) along K per step: +2 in the address field
#pragma unroll
              for (int t = 0; t < MT; ++t) {           // row groups innermost: consecutive MMAs hit different accumulators
                const uint32_t d = tmem_base + sb * (uint32_t)(MT * SCN) + (uint32_t)(t * SCN);
                const uint32_t a_lo = (base + ab * a_buf_bytes + (uint32_t)kb * A_KB_BYTES + (uint32_t)t * 16384u) >> 4;
                umma_f16(d, (desc_hi | (uint64_t)(a_lo & 0x3fffu)) + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(smem_u32(&bar_bempty[st]));
          umma_commit(smem_u32(&bar_tfull[sb]));
          if (sc == n_super - 1) umma_commit(smem_u32(&bar_aempty[ab]));      // x tile no longer needed by the tensor core
        }
        __syncwarp();
      }
    }
  } else {
    // ============================ LayerNorm in place, drain, stencil, store =============================
    const int ct = threadIdx.x - 64;                 // 0..511
    const int grp = ct / GT;                         // compute group
    const int gt = ct % GT;                          // thread within the group
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may read
    // 16-column halves of a chunk this warp drains (gate: 0 = x1, 1 = x2): 8-warp groups split them over two warps per quarter,
    // 4-warp groups take both
    const int slice0 = NG == 2 ? (((warp - 2) >> 2) & 1) : 0;
    constexpr int kSliceStep = NG == 2 ? 2 : 1;
    // stencil mapping: channel group fastest, then the pixel column, then the row band
    const int cg = gt & 3;
    const int tx = (gt >> 2) % TW;
    const int band = (gt >> 2) / TW;
    uint8_t* conv_buf = base_ptr + g.off_conv + (size_t)grp * Cfg::CONV_BYTES;
    // shared-memory swizzle of the fp16 tile, keyed on the halo'd column xh so that the drain's 16-byte stores (one pixel
    // per lane) and the stencil's loads are both bank-conflict free:
    //   GATE  (32-byte rows): 16-byte half h of pixel (yh, xh) lives at half h ^ ((xh >> 2) & 1)
    //   plain (64-byte rows): 16-byte chunk j lives at chunk j ^ ((xh >> 1) & 3)
    uint32_t rd_off[3];                              // this thread's byte offset inside a pixel row, per horizontal tap
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xh = tx + kx;
      rd_off[kx] = GATE ? (uint32_t)((((cg >> 1) ^ ((xh >> 2) & 1)) << 4) | ((cg & 1) << 3)) : (uint32_t)((cg ^ ((xh >> 1) & 3)) << 4);
    }
    // per-thread constants of the tile geometry (no divisions inside the item loop): LayerNorm tasks (pixel m = task / 4 -> row, column
    // in the halo'd tile) and the pixels this thread drains
    constexpr int LNR = (NPIX * 4 + kFwCompute - 1) / kFwCompute;
    int ln_yx[LNR];
#pragma unroll
    for (int r = 0; r < LNR; ++r) { const int m = (r * kFwCompute + ct) >> 2; ln_yx[r] = ((m / SW) << 8) | (m % SW); }
    int dr_yx[MT];
#pragma unroll
    for (int t = 0; t < MT; ++t) { const int m = t * 128 + quarter * 32 + lane; dr_yx[t] = ((m / SW) << 8) | (m % SW); }
    const int nbar = 1 + grp * 2;                    // named barriers of this group: nbar, nbar + 1
    auto grp_bar = [](int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(GT) : "memory"); };
    // ---- LayerNorm of an item's halo'd tile in place; pixels outside the image stay zero ----
    auto layernorm_tile = [&](int item, uint32_t itn) {
      int b_, x0, y0;
      item_geo(item, b_, x0, y0);
      const uint32_t ab = itn % NA;
      uint8_t* a_tile = base_ptr + (size_t)ab * a_buf_bytes;
      mbar_wait(smem_u32(&bar_afull[ab]), (itn / NA) & 1u);
      if (g.ln_mode && !(FW_DBG(g) & 8)) {
        // four threads per pixel; thread `part` owns the physical 16-byte chunks 2*part + (e ^ (m & 1)), e = 0, 1, of every
        // k-block (the row parity term keeps the quarter-warp's LDS.128 conflict free); statistics meet through shuffles
#pragma unroll
        for (int r = 0; r < LNR; ++r) {
          const int task = r * kFwCompute + ct;
          const int m = task >> 2, part = task & 3;
          const int py = y0 - 1 + (ln_yx[r] >> 8), px = x0 - 1 + (ln_yx[r] & 255);
          const bool act = m < NPIX && py >= 0 && py < g.H && px >= 0 && px < g.W;
          float s1 = 0.f, s2 = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
          for (int kb = 0; kb < (NKB ? NKB : 3); ++kb) {
            if (act && kb < nkb) {
              const uint8_t* a_row = a_tile + (size_t)kb * A_KB_BYTES + (size_t)m * 128;
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const uint4 v = *reinterpret_cast<const uint4*>(a_row + ((2 * part + (e ^ (m & 1))) << 4));
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {                 // 16-bit x 16-bit + fp32 (FHFMA): no unpack instructions
                  const unsigned short lo = lo16(w4[i]), hi = hi16(w4[i]);
                  s1 = fma16<T>(lo, T::kOne, s1);
                  s1b = fma16<T>(hi, T::kOne, s1b);
                  s2 = fma16<T>(lo, lo, s2);
                  s2b = fma16<T>(hi, hi, s2b);
                }
              }
            }
          }
          s1 += s1b; s2 += s2b;
          s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
          const float inv_k = 1.0f / (float)g.C;
          const float mu = s1 * inv_k;
          const float rstd = rsqrtf(fmaxf(fmaf(s2, inv_k, -mu * mu), 0.f) + 1e-5f);
          const float shift = g.ln_mode == 2 ? 0.f : -rstd * mu;            // BiasFree: numerator not centred
#pragma unroll
          for (int kb = 0; kb < (NKB ? NKB : 3); ++kb) {
            if (act && kb < nkb) {
              uint8_t* a_row = a_tile + (size_t)kb * A_KB_BYTES + (size_t)m * 128;
              const int valid = min(64, g.C - kb * 64);                       // channels of this k-block that exist
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                // physical chunk j holds logical chunk j ^ (m & 7): the zero padding above C must stay zero
                const int j = 2 * part + (e ^ (m & 1));
                if ((j ^ (m & 7)) * 8 < valid) {
                  uint4 v = *reinterpret_cast<const uint4*>(a_row + (j << 4));
                  uint32_t* w4 = &v.x;
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    w4[i] = pack2<T>(fmaf(unpack_lo<T>(w4[i]), rstd, shift), fmaf(unpack_hi<T>(w4[i]), rstd, shift));
                  *reinterpret_cast<uint4*>(a_row + (j << 4)) = v;
                }
              }
            }
          }
        }
      }
      fence_proxy_async();                           // generic-proxy writes of the tile -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_aready[ab]));
    };
    uint32_t it = 0, uses = 0;
    if (NA == 2 && (int)blockIdx.x < g.n_items) layernorm_tile(blockIdx.x, 0);
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      int b, x0, y0;
      item_geo(item, b, x0, y0);
      if (NA == 2) {
        if (item + (int)gridDim.x < g.n_items) layernorm_tile(item + gridDim.x, it + 1);   // next item's tile, one item ahead
      } else {
        layernorm_tile(item, it);
      }

      // interior tile: the halo is inside the image too, so t[n] can ride on the accumulator seed instead of the drain
      const bool interior = x0 >= 1 && y0 >= 1 && x0 + TW + 1 <= g.W && y0 + TH + 1 <= g.H;
      // inside-the-image flags of the pixels this thread drains (border tiles: t[n] is added inside only)
      bool inside[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const int m = t * 128 + quarter * 32 + lane;
        const int py = y0 - 1 + (dr_yx[t] >> 8), px = x0 - 1 + (dr_yx[t] & 255);
        inside[t] = m < NPIX && py >= 0 && py < g.H && px >= 0 && px < g.W;
      }
      // this thread's first output pixel of the item (the chunk loops only add the channel offset)
      unsigned short* out_item = reinterpret_cast<unsigned short*>(g.out) + (size_t)b * g.out_bstride +
                                 ((size_t)(y0 + band * R) * g.W + (x0 + tx)) * g.out_pitch;
      const size_t out_row = (size_t)g.W * g.out_pitch;
      // plain kernel with two output tensors (q|k and v): second base pointer / row stride, selected per chunk
      unsigned short* out_item2 = GATE ? nullptr : reinterpret_cast<unsigned short*>(g.out2) + (size_t)b * g.out2_bstride +
                                                   ((size_t)(y0 + band * R) * g.W + (x0 + tx)) * g.out2_pitch - g.split;
      const size_t out_row2 = (size_t)g.W * g.out2_pitch;

      for (int sc = 0; sc < n_super; ++sc, ++uses) {
      // the group's chunks of this super-chunk: j = j0, j0 + 2, ... (j0 alternates per item so odd chunk counts balance out)
      const uint32_t sb = uses & 1u;
      mbar_wait(smem_u32(&bar_tfull[sb]), (uses >> 1) & 1u);
      tc_fence_after();
      const int nvalid = min(scb, g.n_chunks - sc * scb);
      for (int j = (grp + (int)it + (g.rot ? sc : 0)) % NG; j < nvalid; j += NG) {
        const int c = sc * scb + j;
        // ---- drain: TMEM -> (+ t[n]) -> fp16 shared-memory tile of this group ----
#pragma unroll
        for (int slice = slice0; slice < 2; slice += kSliceStep) {
        const int nb = GATE ? (slice ? g.hp + c * 16 : c * 16) : c * 32 + slice * 16;    // first pre-conv channel of this warp's slice
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const int m = t * 128 + quarter * 32 + lane;
          uint32_t acc[16];
          if (!(FW_DBG(g) & 16)) {
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + sb * (uint32_t)(MT * SCN) + (uint32_t)(t * SCN + j * kFwChunk + slice * 16), acc);
            tmem_ld_wait();
          }
          if (m < NPIX && !(FW_DBG(g) & 2)) {
            uint32_t pk[8];
            if (g.has_t && !interior && inside[t]) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 t4 = reinterpret_cast<const float4*>(svec + nb)[e];
                pk[2 * e] = pack_f16_sat(__uint_as_float(acc[4 * e]) + t4.x, __uint_as_float(acc[4 * e + 1]) + t4.y);
                pk[2 * e + 1] = pack_f16_sat(__uint_as_float(acc[4 * e + 2]) + t4.z, __uint_as_float(acc[4 * e + 3]) + t4.w);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) pk[e] = pack_f16_sat(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1]));
            }
            const int xh = dr_yx[t] & 255;
            if (GATE) {
              uint8_t* row = conv_buf + (size_t)slice * (MT * 128 * 32) + (size_t)m * 32;
              const int sw = (xh >> 2) & 1;
              *reinterpret_cast<uint4*>(row + ((0 ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(row + ((1 ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
              uint8_t* row = conv_buf + (size_t)m * 64;
              const int sw = (xh >> 1) & 3;
              *reinterpret_cast<uint4*>(row + (((slice * 2) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(row + (((slice * 2 + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
        }
        grp_bar(nbar);                                                 // fp16 tile complete

        // ---- stencil from shared memory ----
        if (FW_DBG(g) & 1) {
        } else if (GATE) {
          const int ch = c * 16 + cg * 4;                              // gated channel of this thread
          const uint8_t* wbase = reinterpret_cast<const uint8_t*>(sdw) + (size_t)c * 576 + cg * 8;
          uint2 w1[9], w2[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            w1[t] = *reinterpret_cast<const uint2*>(wbase + t * 32);
            w2[t] = *reinterpret_cast<const uint2*>(wbase + 288 + t * 32);
          }
          // accumulator seeds: depthwise bias (+ the conv of the constant t on interior tiles)
          const uint32_t* sd = sseed + (interior ? g.n_chunks * 16 : 0) + c * 16 + cg * 2;
          const uint32_t b1[2] = {sd[0], sd[1]}, b2[2] = {sd[8], sd[9]};
          const int x = x0 + tx;
          const bool ok = ch < g.hp && x < g.W;
          unsigned short* outp = out_item + ch;
          const uint8_t* src = conv_buf + (size_t)((band * R) * SW + tx) * 32;
          uint32_t p[3][2], qq[3][2];
#pragma unroll
          for (int r = 0; r < R + 2; ++r) {
            if (r < R) { p[r % 3][0] = b1[0]; p[r % 3][1] = b1[1]; qq[r % 3][0] = b2[0]; qq[r % 3][1] = b2[1]; }
            uint2 v1[3], v2[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              v1[kx] = *reinterpret_cast<const uint2*>(src + rd_off[kx] + (r * SW + kx) * 32);
              v2[kx] = *reinterpret_cast<const uint2*>(src + rd_off[kx] + (r * SW + kx) * 32 + MT * 128 * 32);
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
              uint2 ov;
              // packed fp16 GELU gate for both storage types (fp16: measured 8.3e-4 max-abs on the cfg2 forward against 9.0e-4 with
              // the fp32 erf form, and faster); bf16 storage converts the fp16 pairs
              const uint32_t ga = gelu_gate_h2(p[o % 3][0], qq[o % 3][0]), gb = gelu_gate_h2(p[o % 3][1], qq[o % 3][1]);
              if (T::kFmt == 1) {
                const float2 fa = h2_to_f2(ga), fb = h2_to_f2(gb);
                ov.x = pack2<T>(fa.x, fa.y);
                ov.y = pack2<T>(fb.x, fb.y);
              } else {
                ov.x = ga; ov.y = gb;
              }
              if (ok && y0 + band * R + o < g.H && !(FW_DBG(g) & 4)) *reinterpret_cast<uint2*>(outp + o * out_row) = ov;
            }
          }
        } else {
          const int ch = c * 32 + cg * 8;
          const uint8_t* wbase = reinterpret_cast<const uint8_t*>(sdw) + (size_t)c * 576 + cg * 16;
          uint4 wt[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) wt[t] = *reinterpret_cast<const uint4*>(wbase + t * 64);
          const uint32_t* sd = sseed + (interior ? g.n_chunks * 16 : 0) + c * 16 + cg * 4;
          const uint32_t bb[4] = {sd[0], sd[1], sd[2], sd[3]};
          const int x = x0 + tx;
          const bool ok = ch < g.n_pre && x < g.W;
          // q|k and v may be two dense tensors (split is a multiple of 8, so a thread's 8 channels never straddle it)
          const bool second = ch >= g.split;
          unsigned short* outp = (second ? out_item2 : out_item) + ch;
          const size_t orow_stride = second ? out_row2 : out_row;
          const uint8_t* src = conv_buf + (size_t)((band * R) * SW + tx) * 64;
          uint32_t p[3][4];
#pragma unroll
          for (int r = 0; r < R + 2; ++r) {
            if (r < R) { p[r % 3][0] = bb[0]; p[r % 3][1] = bb[1]; p[r % 3][2] = bb[2]; p[r % 3][3] = bb[3]; }
            uint4 v[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) v[kx] = *reinterpret_cast<const uint4*>(src + rd_off[kx] + (r * SW + kx) * 64);
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
              if (ok && y0 + band * R + o < g.H && !(FW_DBG(g) & 4)) *reinterpret_cast<uint4*>(outp + o * orow_stride) = ov;
            }
          }
        }
        grp_bar(nbar + 1);                                             // fp16 tile may be overwritten
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[sb]));          // this warp is done with the super-chunk's accumulators
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------
struct FwPlan {
  int mt, tw, na, sc, ng;  // template configuration
  uint32_t smem;
  FwArgs g;
};

static int tile_h(int mt, int tw) { return mt == 3 ? (tw == 32 ? 8 : 16) : 12; }

static bool plan_fits(const PirPwDw* d, FwPlan* p, int mt, int na, int sc, int ng = 2) {
  FwArgs& g = p->g;
  if (2 * mt * sc * kFwChunk > 512) return false;                      // TMEM
  if (ng == 4 && (sc != 4 || mt != 2)) return false;
  // shared-memory plan: [A: na x nkb x MT x 16 KB] [B ring: 2 x nkb x SC x 4 KB] [fp16 tiles: 2 x MT x 8 KB] [dw taps] [t] [accumulator seeds]
  uint32_t off = (uint32_t)na * g.nkb * mt * 16384u;
  g.off_b = off; off += (uint32_t)kFwBStages * g.nkb * (uint32_t)sc * 4096u;
  g.off_conv = off; off += (uint32_t)ng * (uint32_t)mt * 128u * 64u;
  g.off_dw = off; off += (uint32_t)g.n_chunks * 576u;
  g.off_vec = off; off += (uint32_t)g.n_vec * 4u;
  g.off_bias = off; off += (uint32_t)g.n_chunks * 128u;
  if (off + 1024u > 227u * 1024u - 1024u) return false;
  p->mt = mt; p->na = na; p->sc = sc; p->ng = ng; p->smem = off + 1024u;
  p->tw = (mt == 3 && d->W > 16) ? 32 : 16;
  const int th = tile_h(mt, p->tw);
  g.tiles_x = (d->W + p->tw - 1) / p->tw;
  g.tiles_y = (d->H + th - 1) / th;
  g.n_items = g.tiles_x * g.tiles_y * d->B;
  {
    auto magic = [](uint32_t dv) { return dv <= 1 ? 0u : (uint32_t)((0x100000000ull + dv - 1) / dv); };
    const uint64_t per_img = (uint64_t)g.tiles_x * g.tiles_y;
    if ((uint64_t)g.n_items * per_img >= 0x100000000ull) return false;
    g.mg_per_img = magic((uint32_t)per_img); g.mg_tiles_x = magic((uint32_t)g.tiles_x);
  }
  return true;
}

static int plan_pwdw(const PirPwDw* d, FwPlan* p) {
  FwArgs& g = p->g;
  g = FwArgs{};
  g.B = d->B; g.H = d->H; g.W = d->W; g.C = d->C;
  const bool gate = d->gate != 0;
  g.n_pre = gate ? 2 * d->N : d->N;
  g.hp = gate ? d->N : 0;
  g.ln_mode = d->ln_mode;
  g.nkb = (d->C + 63) / 64;
  g.n_chunks = gate ? (d->N + 15) / 16 : (d->N + 31) / 32;
  g.n_vec = gate ? g.hp + g.n_chunks * 16 : g.n_chunks * 32;
  g.has_bias = d->dw_bias ? 1 : 0;
  g.has_t = d->vec_t ? 1 : 0;
  { const char* e = getenv("PIR_PWDW_DBG"); g.dbg = e ? atoi(e) : 0; }
  // configurations <MT row groups, NA x-tile buffers, SC chunks per MMA>, best first.  PIR_PWDW_CFG="<mt><na><sc>" forces one.
  static const char* force = getenv("PIR_PWDW_CFG");
  if (force && force[0] && force[1] && force[2])
    return plan_fits(d, p, force[0] - '0', force[1] - '0', force[2] - '0', force[3] ? force[3] - '0' : 2) ? PIR_OK : PIR_ERR_UNSUPPORTED;
  static const int pref[][4] = {{2, 2, 4, 4}, {2, 1, 4, 4}, {2, 2, 4, 2}, {2, 1, 4, 2}, {2, 2, 2, 2}, {2, 1, 2, 2}, {3, 1, 2, 2}};
  static const bool no4 = getenv("PIR_PWDW_NG2") != nullptr;            // A/B: never use the four-group variants
  for (const auto& c : pref)
    if (!(no4 && c[3] == 4) && plan_fits(d, p, c[0], c[1], c[2], c[3])) return PIR_OK;
  return PIR_ERR_UNSUPPORTED;
}

template <class T, class Cfg>
static int launch_cfg(const FwPlan& p, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t stream) {
  if (!pir_smem_attr_once(reinterpret_cast<const void*>(pwdw_kernel<T, Cfg>), (int)(227 * 1024 - 1024), "pir_pwdw")) return PIR_ERR_CUDA;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int grid = p.g.n_items < num_sms ? p.g.n_items : num_sms;
  pir_launch(pwdw_kernel<T, Cfg>, dim3(grid), dim3(kFwThreads), p.smem, stream, tmA, tmB, p.g);
  return pir_check_launch("pir_pwdw");
}

template <class T, bool GATE>
static int launch_pwdw(const PirPwDw* d, cudaStream_t stream) {
  FwPlan p;
  if (plan_pwdw(d, &p) != PIR_OK) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_pwdw: C = %d, N = %d does not fit the shared-memory plan", d->C, d->N);
  p.g.dw_w = d->dw_w; p.g.dw_bias = d->dw_bias; p.g.vec_t = d->vec_t;
  {
    // equal-sized super-chunks (9 chunks run as 3 + 3 + 3, not 4 + 4 + 1) with the group <-> chunk assignment rotated by the
    // super-chunk index, so the idle group moves around: 440 -> 420 us on the C = 96 qkv kernel (same-box A/B).  PIR_PWDW_BAL=0 restores
    // the plain split (bit 0: equal sizes, bit 1: rotation).
    static const char* bal = getenv("PIR_PWDW_BAL");
    const int mode = bal ? atoi(bal) : 3;
    const int n_super = (p.g.n_chunks + p.sc - 1) / p.sc;
    p.g.scb = (mode & 1) ? (p.g.n_chunks + n_super - 1) / n_super : p.sc;
    p.g.rot = (mode & 2) ? 1 : 0;
  }
  p.g.out = d->out; p.g.out_pitch = d->out_pitch; p.g.out_bstride = d->out_bstride;
  {
    const bool two = !d->gate && d->out2 != nullptr && d->split > 0 && d->split < d->N;
    if (two && (d->split % 8 || d->out2_pitch % 8 || d->out2_bstride % 8 || ((uintptr_t)d->out2 & 15)))
      return pir_fail(PIR_ERR_ARG, "pir_pwdw: split / out2 pitch / out2 pointer are not 16-byte aligned");
    p.g.split = two ? d->split : p.g.n_pre;
    p.g.out2 = two ? d->out2 : d->out; p.g.out2_pitch = two ? d->out2_pitch : d->out_pitch; p.g.out2_bstride = two ? d->out2_bstride : d->out_bstride;
  }
  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB;
  {
    const int th = tile_h(p.mt, p.tw);
    const uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->a_pitch * 2, (uint64_t)d->a_pitch * 2 * d->W, (uint64_t)d->a_bstride * 2};
    const uint32_t box[4] = {64, (uint32_t)(p.tw + 2), (uint32_t)(th + 2), 1};
    if (int e = pir_make_tmap(&tmA, dt, 4, d->a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t kpad = (uint64_t)p.g.nkb * 64;
    const uint64_t dims[2] = {kpad, (uint64_t)p.g.n_pre};
    const uint64_t strides[1] = {kpad * 2};
    const uint32_t box[2] = {64, 16};
    if (int e = pir_make_tmap(&tmB, dt, 2, d->w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  static const bool generic = getenv("PIR_PWDW_NKB0") != nullptr;
  if (!generic) {
    if (p.g.nkb == 1 && p.mt == 2 && p.na == 2 && p.sc == 4 && p.ng == 4) return launch_cfg<T, FwCfg<2, 16, 6, GATE, 2, 4, 4, 1>>(p, tmA, tmB, stream);
    if (p.g.nkb == 2 && p.mt == 2 && p.na == 1 && p.sc == 4 && p.ng == 4) return launch_cfg<T, FwCfg<2, 16, 6, GATE, 1, 4, 4, 2>>(p, tmA, tmB, stream);
    if (p.g.nkb == 3 && p.mt == 2 && p.na == 1 && p.sc == 2 && p.ng == 2) return launch_cfg<T, FwCfg<2, 16, 3, GATE, 1, 2, 2, 3>>(p, tmA, tmB, stream);
  }
  if (p.mt == 2 && p.na == 2 && p.sc == 4 && p.ng == 4) return launch_cfg<T, FwCfg<2, 16, 6, GATE, 2, 4, 4>>(p, tmA, tmB, stream);
  if (p.mt == 2 && p.na == 1 && p.sc == 4 && p.ng == 4) return launch_cfg<T, FwCfg<2, 16, 6, GATE, 1, 4, 4>>(p, tmA, tmB, stream);
  if (p.mt == 2 && p.na == 2 && p.sc == 4) return launch_cfg<T, FwCfg<2, 16, 3, GATE, 2, 4>>(p, tmA, tmB, stream);
  if (p.mt == 2 && p.na == 1 && p.sc == 4) return launch_cfg<T, FwCfg<2, 16, 3, GATE, 1, 4>>(p, tmA, tmB, stream);
  if (p.mt == 2 && p.na == 2 && p.sc == 2) return launch_cfg<T, FwCfg<2, 16, 3, GATE, 2, 2>>(p, tmA, tmB, stream);
  if (p.mt == 2 && p.na == 1 && p.sc == 2) return launch_cfg<T, FwCfg<2, 16, 3, GATE, 1, 2>>(p, tmA, tmB, stream);
  if (p.mt == 3 && p.na == 1 && p.sc == 2 && p.tw == 32) return launch_cfg<T, FwCfg<3, 32, 4, GATE, 1, 2>>(p, tmA, tmB, stream);
  if (p.mt == 3 && p.na == 1 && p.sc == 2) return launch_cfg<T, FwCfg<3, 16, 4, GATE, 1, 2>>(p, tmA, tmB, stream);
  return pir_fail(PIR_ERR_UNSUPPORTED, "pir_pwdw: configuration <%d,%d,%d> is not compiled", p.mt, p.na, p.sc);
}

// channel-major variant (pwdwt.cu)
bool pwdwt_supported(const PirPwDw* d);
int pwdwt_run(const PirPwDw* d, cudaStream_t stream);

}  // namespace pir

extern "C" int pir_pwdw_supported(int32_t C, int32_t N, int32_t gate) {
  PirPwDw d{};
  d.B = 1; d.H = 64; d.W = 64; d.C = C; d.N = N; d.gate = gate;
  pir::FwPlan p;
  if ((C % 8) || (N % 8) || C <= 0 || N <= 0) return 0;
  return pir::plan_pwdw(&d, &p) == PIR_OK ? 1 : 0;
}

extern "C" int pir_pwdw_split_supported(int32_t C, int32_t N) {
  PirPwDw d{};
  d.B = 1; d.H = 64; d.W = 64; d.C = C; d.N = N; d.gate = 0;
  if ((C % 8) || (N % 8) || C <= 0 || N <= 0) return 0;
  pir::FwPlan p;
  return (pir::pwdwt_supported(&d) || pir::plan_pwdw(&d, &p) == PIR_OK) ? 1 : 0;
}

extern "C" int pir_pwdw(const PirPwDw* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_pwdw: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->N <= 0) return pir_fail(PIR_ERR_ARG, "pir_pwdw: empty problem");
  if ((d->C % 8) || (d->N % 8) || (d->a_pitch % 8) || (d->a_bstride % 8) || (d->out_pitch % 8) || (d->out_bstride % 8) ||
      ((uintptr_t)d->a & 15) || ((uintptr_t)d->w & 15) || ((uintptr_t)d->out & 15) || ((uintptr_t)d->dw_w & 15))
    return pir_fail(PIR_ERR_ARG, "pir_pwdw: channel counts / pitches / pointers are not 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (pir::pwdwt_supported(d)) return pir::pwdwt_run(d, s);
  if (d->dtype == PIR_DTYPE_BF16)
    return d->gate ? pir::launch_pwdw<pir::BF16, true>(d, s) : pir::launch_pwdw<pir::BF16, false>(d, s);
  return d->gate ? pir::launch_pwdw<pir::FP16, true>(d, s) : pir::launch_pwdw<pir::FP16, false>(d, s);
}
