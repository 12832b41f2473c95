// Weight gradients of the 1x1 / dense 3x3 convolutions for sm_100a.    Reference: autograd of nn.Conv2d in net/model.py
// (conv backward w.r.t. weight), called from loss.backward() in train.py:37-46.
//
//   dW[n][k][tap] = sum over pixels p of  dY[p, n] * X[p + off(tap), k]
//
// The contraction runs over PIXELS (up to B*H*W = 524 288 at the training config) while the result is a small
// [M x N] matrix: the same shape of problem as the MDTA Gram (mdta.cu), so it is built the same way -- a split-K tcgen05 GEMM
// whose operands are MN-major (channels contiguous) 64-pixel tiles taken straight from the two NHWC tensors by TMA, fp32
// accumulation in TMEM, one fp32 partial per split (deterministic; pir_wgrad_finalize reduces them into the parameter's
// gradient).  3x3 taps are nine independent Grams against the shifted X (4-D TMA boxes, zero fill outside the image).
// The operand with fewer channels takes the 128-row side of the MMA and the wider one the 256-column side ("swap"), and the
// accumulator is stored transposed when needed so the workspace layout never depends on that choice.
// Spare warps add up the columns of `a` from the same shared-memory stages (bias / LayerNorm-beta gradients).
#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kWgThreads = 192;            // warp0 TMA, warp1 MMA, warps 2-5 column sums + epilogue
constexpr int kWgPix = 64;                 // pixels per stage (4 UMMA K-steps of 16)
constexpr int kWgGroupBytes = kWgPix * 128;
constexpr int kWgMaxStages = 6;

struct WgArgs {
  int M, N, taps;
  int spatial, swap, per_image, splits;
  int tw, th, tiles_x, cpi;    // pixel chunks per image
  int total_chunks, ups, cps;  // chunks per split: over the batch / per image
  int rblocks, cblocks;
  int stages;
  int colsum_on;
  float* ws;
  float* colsum;
};

template <class T>
__global__ void __launch_bounds__(kWgThreads)
wgrad_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC, const WgArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kWgMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kWgMaxStages];
  __shared__ __align__(8) uint64_t bar_accum;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red[4][4][64];

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int rb = blockIdx.x / g.cblocks;
  const int cb = blockIdx.x % g.cblocks;
  const int part = blockIdx.y;
  const int tap = blockIdx.z;
  const int Rdim = g.swap ? g.N : g.M;
  const int Cdim = g.swap ? g.M : g.N;
  const int r0 = rb * 128, c0 = cb * 256;
  const int rgroups = min(2, (Rdim - r0 + 63) / 64);
  const int cgroups = min(4, (Cdim - c0 + 63) / 64);
  const int block_n = cgroups * 64;
  const uint32_t stage_bytes = (uint32_t)(rgroups + cgroups) * kWgGroupBytes;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  int u_begin, u_end;
  if (g.per_image) {
    const int b = part / g.splits, s = part % g.splits;
    u_begin = b * g.cpi + min(s * g.cps, g.cpi);
    u_end = b * g.cpi + min((s + 1) * g.cps, g.cpi);
  } else {
    u_begin = min(part * g.ups, g.total_chunks);
    u_end = min(u_begin + g.ups, g.total_chunks);
  }
  const int nst = u_end - u_begin;
  const uint32_t tmem_cols = block_n <= 64 ? 64 : (block_n <= 128 ? 128 : 256);
  const int dy = g.taps == 9 ? tap / 3 - 1 : 0, dx = g.taps == 9 ? tap % 3 - 1 : 0;
  // column sums of `a`: a is the row operand unless swapped
  const bool cs_here = g.colsum_on && tap == 0 && (g.swap ? rb == 0 : cb == 0);
  const int cs_first = g.swap ? rgroups : 0;
  const int cs_groups = g.swap ? cgroups : rgroups;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), cs_here ? 5 : 1);
    }
    mbar_init(smem_u32(&bar_accum), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_smem), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nst; ++it) {
      mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
      if (elect_one()) {
        const uint32_t full = smem_u32(&bar_full[stage]);
        const uint32_t dst = smem_base + (uint32_t)stage * stage_bytes;
        mbar_expect_tx(full, stage_bytes);
        const int u = u_begin + it;
        const int b = u / g.cpi, q = u % g.cpi;
        if (g.spatial) {
          const int x0 = (q % g.tiles_x) * g.tw, y0 = (q / g.tiles_x) * g.th;
          // the shift belongs to operand b: the row operand when swapped, the column operand otherwise
          const int rx = x0 + (g.swap ? dx : 0), ry = y0 + (g.swap ? dy : 0);
          const int cx = x0 + (g.swap ? 0 : dx), cy = y0 + (g.swap ? 0 : dy);
          for (int gi = 0; gi < rgroups; ++gi) tma_load_4d(dst + gi * kWgGroupBytes, &tmR, full, r0 + gi * 64, rx, ry, b);
          for (int gi = 0; gi < cgroups; ++gi) tma_load_4d(dst + (rgroups + gi) * kWgGroupBytes, &tmC, full, c0 + gi * 64, cx, cy, b);
        } else {
          const int p = q * kWgPix;
          for (int gi = 0; gi < rgroups; ++gi) tma_load_3d(dst + gi * kWgGroupBytes, &tmR, full, r0 + gi * 64, p, b);
          for (int gi = 0; gi < cgroups; ++gi) tma_load_3d(dst + (rgroups + gi) * kWgGroupBytes, &tmC, full, c0 + gi * 64, p, b);
        }
      }
      __syncwarp();
      if (++stage == g.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(T::kFmt, 128, block_n, 1, 1);      // both operands MN-major
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nst; ++it) {
      mbar_wait(smem_u32(&bar_full[stage]), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_src = smem_base + (uint32_t)stage * stage_bytes;
        const uint32_t b_src = a_src + (uint32_t)rgroups * kWgGroupBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {       // 16 pixels (two 8-row swizzle atoms = 2048 B) per MMA
          const uint64_t ad = make_sdesc_sw128(a_src + k * 2048, kWgGroupBytes, 1024);
          const uint64_t bd = make_sdesc_sw128(b_src + k * 2048, kWgGroupBytes, 1024);
          umma_f16(tmem_base, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar_empty[stage]));
        if (it == nst - 1) umma_commit(smem_u32(&bar_accum));
      }
      __syncwarp();
      if (++stage == g.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    const int wq = warp - 2;
    if (cs_here) {
      // ---- column sums of a: warp wq takes pixel rows [16 wq, 16 wq + 16) of every group of a; a lane owns one 16-byte chunk
      //      (8 channels) of rows 4i + lane/8 ----
      const int rsub = lane >> 3, ch8 = lane & 7;
      float ss[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) ss[i][e] = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      const uint8_t* base = smem_raw + (smem_base - smem_u32(smem_raw));
      for (int it = 0; it < nst; ++it) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        const uint8_t* st = base + (size_t)stage * stage_bytes + (size_t)cs_first * kWgGroupBytes;
        uint4 v[4][4];
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          if (gi < cs_groups) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = wq * 16 + i * 4 + rsub;
              v[gi][i] = *reinterpret_cast<const uint4*>(st + gi * kWgGroupBytes + row * 128 + ((ch8 ^ (row & 7)) << 4));
            }
          }
        }
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          if (gi < cs_groups) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t w4[4] = {v[gi][i].x, v[gi][i].y, v[gi][i].z, v[gi][i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                ss[gi][2 * e] += unpack_lo<T>(w4[e]);
                ss[gi][2 * e + 1] += unpack_hi<T>(w4[e]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_empty[stage]));
        if (++stage == g.stages) { stage = 0; phase ^= 1u; }
      }
#pragma unroll
      for (int gi = 0; gi < 4; ++gi) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float t = ss[gi][e];
          t += __shfl_xor_sync(0xffffffffu, t, 8);
          t += __shfl_xor_sync(0xffffffffu, t, 16);
          if (rsub == 0) red[wq][gi][ch8 * 8 + e] = t;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");        // the four statistics warps only (cs_here is block-uniform)
      const int a0 = g.swap ? c0 : r0;                      // first channel of a held by this CTA
      const int t = threadIdx.x - 64;                       // 0..127
      for (int e = t; e < cs_groups * 64; e += 128) {
        const int gi = e >> 6, ch = e & 63;
        const int m = a0 + gi * 64 + ch;
        if (m < g.M) g.colsum[(size_t)part * g.M + m] = red[0][gi][ch] + red[1][gi][ch] + red[2][gi][ch] + red[3][gi][ch];
      }
    }
    // ---- epilogue: TMEM -> fp32 partial.  Lane = accumulator row (channel of the row operand) ----
    const int quarter = warp & 3;
    const int row = r0 + quarter * 32 + lane;
    if (nst > 0) {
      mbar_wait(smem_u32(&bar_accum), 0);
      tc_fence_after();
    }
    float* out = g.ws + ((size_t)part * g.taps + tap) * (size_t)g.M * g.N;
    const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool warp_live = r0 + quarter * 32 < Rdim;        // warp-uniform
    for (int cc = 0; cc < block_n; cc += 16) {
      if (!warp_live || c0 + cc >= Cdim) continue;
      uint32_t acc[16];
      if (nst > 0) {
        tmem_ld16(taddr_row + (uint32_t)cc, acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = 0u;
      }
      if (!g.swap) {                                        // row = m, columns = n: float4 along n
        if (row < g.M) {
          float* o = out + (size_t)row * g.N + c0 + cc;
#pragma unroll
          for (int q = 0; q < 16; q += 4)
            if (c0 + cc + q < g.N)
              *reinterpret_cast<float4*>(o + q) = make_float4(__uint_as_float(acc[q]), __uint_as_float(acc[q + 1]),
                                                              __uint_as_float(acc[q + 2]), __uint_as_float(acc[q + 3]));
        }
      } else {                                              // row = n, columns = m: lanes are consecutive n -> coalesced
        if (row < g.N) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int m = c0 + cc + q;
            if (m < g.M) out[(size_t)m * g.N + row] = __uint_as_float(acc[q]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------
// finalize: reduce the partials into the parameter gradient (+ LayerNorm gamma/beta and bias gradients)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int prow_of(int r, int half, int half_pad) { return r < half ? r : r - half + half_pad; }

__device__ __forceinline__ float sum_parts(const float* p, size_t stride, int P) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int i = 0;
  for (; i + 4 <= P; i += 4) {
    s0 += p[(size_t)i * stride];
    s1 += p[(size_t)(i + 1) * stride];
    s2 += p[(size_t)(i + 2) * stride];
    s3 += p[(size_t)(i + 3) * stride];
  }
  for (; i < P; ++i) s0 += p[(size_t)i * stride];
  return (s0 + s1) + (s2 + s3);
}

// plain weights: 32 consecutive (r, k) elements x 8 split lanes per block, all taps
__global__ void __launch_bounds__(256)
wgrad_fin_kernel(const PirWgradFin f) {
  __shared__ float red[8][33];
  const int le = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const long long e = (long long)blockIdx.x * 32 + le;
  const bool ok = e < (long long)f.R * f.Cc;
  const int r = ok ? (int)(e / f.Cc) : 0, k = ok ? (int)(e % f.Cc) : 0;
  const int pr = prow_of(r, f.half, f.half_pad);
  const size_t pstride = (size_t)f.taps * f.M * f.N;
  for (int t = 0; t < f.taps; ++t) {
    const float* p0 = f.ws + ((size_t)t * f.M + pr) * f.N + k;
    float s0 = 0.f, s1 = 0.f;
    if (ok) {
      int p = sl;
      for (; p + 8 < f.P; p += 16) { s0 += p0[(size_t)p * pstride]; s1 += p0[(size_t)(p + 8) * pstride]; }
      if (p < f.P) s0 += p0[(size_t)p * pstride];
    }
    red[sl][le] = s0 + s1;
    __syncthreads();
    if (sl == 0 && ok) {
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) g += red[q][le];
      f.dst_w[((size_t)r * f.Cc + k) * f.taps + t] = g * f.inv_scale;
    }
    __syncthreads();
  }
}

// LayerNorm-folded weights, step 1: reduce the P partials IN PLACE into partial 0 (and colsum 0).  A block takes 32 consecutive
// elements x 8 split lanes (coalesced 128-byte rows per split; the serial chain is P / 8 loads, not P); blocks [0, nmain) take the
// matrix, the rest the column sums.
__global__ void __launch_bounds__(256)
wgrad_reduce_inplace_kernel(float* __restrict__ ws, float* __restrict__ colsum, int P, long long MN, int M, int nmain) {
  __shared__ float red[8][33];
  const int le = threadIdx.x & 31, sl = threadIdx.x >> 5;
  float* base;
  long long e, n;
  if ((int)blockIdx.x < nmain) { base = ws; n = MN; e = (long long)blockIdx.x * 32 + le; }
  else { base = colsum; n = M; e = (long long)((int)blockIdx.x - nmain) * 32 + le; }
  float s0 = 0.f, s1 = 0.f;
  if (e < n) {
    int p = sl;
    for (; p + 8 < P; p += 16) { s0 += base[(size_t)p * n + e]; s1 += base[(size_t)(p + 8) * n + e]; }
    if (p < P) s0 += base[(size_t)p * n + e];
  }
  red[sl][le] = s0 + s1;
  __syncthreads();
  if (sl == 0 && e < n) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][le];
    base[e] = t;
  }
}

// step 2: dW = gamma G + beta s, and the column reductions dgamma[k] = sum_r W G, dbeta[k] = sum_r W s.  One block per 8 input channels
// x 32 row lanes (32-byte segments stay whole; the row loop is R / 32 long); block 0 also writes dbias.
__global__ void __launch_bounds__(256)
wgrad_ln_apply_kernel(const PirWgradFin f) {
  __shared__ float sg[32][9], sb[32][9];
  const int kc = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int k = blockIdx.x * 8 + kc;
  float dg = 0.f, db = 0.f;
  if (k < f.Cc) {
    const float gam = f.gamma[k], bet = f.beta ? f.beta[k] : 0.f;
#pragma unroll 4
    for (int r = rl; r < f.R; r += 32) {
      const int pr = prow_of(r, f.half, f.half_pad);
      const float G = f.ws[(size_t)pr * f.N + k] * f.inv_scale;
      const float s = f.colsum ? f.colsum[pr] * f.inv_scale : 0.f;
      const float w = f.w[(size_t)r * f.Cc + k];
      f.dst_w[(size_t)r * f.Cc + k] = fmaf(gam, G, bet * s);
      dg = fmaf(w, G, dg);
      db = fmaf(w, s, db);
    }
  }
  sg[rl][kc] = dg;
  sb[rl][kc] = db;
  __syncthreads();
  if (rl == 0 && k < f.Cc) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) { a += sg[w][kc]; b += sb[w][kc]; }
    f.dst_gamma[k] = a;
    if (f.dst_beta) f.dst_beta[k] = b;
  }
  if (blockIdx.x == 0 && f.dst_bias)
    for (int r = threadIdx.x; r < f.R; r += 256) f.dst_bias[r] = f.colsum[prow_of(r, f.half, f.half_pad)] * f.inv_scale;
}

// plain weights: bias gradient, one thread per row
__global__ void __launch_bounds__(256)
wgrad_fin_bias_kernel(const PirWgradFin f) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r < f.R) f.dst_bias[r] = sum_parts(f.colsum + prow_of(r, f.half, f.half_pad), (size_t)f.M, f.P) * f.inv_scale;
}

// ------------------------------------------------------------------------------------------------------
static int pow2_ceil_wg(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct WgPlan { int spatial, swap, tw, th, tiles_x, cpi, rblocks, cblocks; };

static WgPlan wg_plan(int H, int W, int M, int N, int taps) {
  WgPlan p{};
  p.spatial = taps == 9;
  if (p.spatial) {
    p.tw = pow2_ceil_wg(W) > 64 ? 64 : (pow2_ceil_wg(W) < 8 ? 8 : pow2_ceil_wg(W));
    p.th = kWgPix / p.tw;
    p.tiles_x = (W + p.tw - 1) / p.tw;
    p.cpi = p.tiles_x * ((H + p.th - 1) / p.th);
  } else {
    p.tw = p.th = p.tiles_x = 0;
    p.cpi = (H * W + kWgPix - 1) / kWgPix;
  }
  const int plain = ((M + 127) / 128) * ((N + 255) / 256), swapped = ((N + 127) / 128) * ((M + 255) / 256);
  p.swap = swapped < plain;
  p.rblocks = ((p.swap ? N : M) + 127) / 128;
  p.cblocks = ((p.swap ? M : N) + 255) / 256;
  return p;
}

template <class T>
static int launch_wgrad(const PirWgrad* d, cudaStream_t stream) {
  const WgPlan p = wg_plan(d->H, d->W, d->M, d->N, d->taps);
  WgArgs g{};
  g.M = d->M; g.N = d->N; g.taps = d->taps;
  g.spatial = p.spatial; g.swap = p.swap; g.per_image = d->per_image; g.splits = d->splits;
  g.tw = p.tw; g.th = p.th; g.tiles_x = p.tiles_x; g.cpi = p.cpi;
  g.total_chunks = d->B * p.cpi;
  g.ups = (g.total_chunks + d->splits - 1) / d->splits;
  g.cps = (p.cpi + d->splits - 1) / d->splits;
  g.rblocks = p.rblocks; g.cblocks = p.cblocks;
  g.colsum_on = d->colsum ? 1 : 0;
  g.ws = d->ws; g.colsum = d->colsum;
  const int Rdim = p.swap ? d->N : d->M, Cdim = p.swap ? d->M : d->N;
  const int groups = (Rdim >= 128 ? 2 : (Rdim + 63) / 64) + (Cdim >= 256 ? 4 : (Cdim + 63) / 64);
  const size_t stage_bytes = (size_t)groups * kWgGroupBytes;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  const int max_it = d->per_image ? g.cps : g.ups;
  if (stages > max_it) stages = max_it < 1 ? 1 : max_it;
  g.stages = stages;
  const size_t smem = stages * stage_bytes + 1024;

  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB;
  auto make = [&](CUtensorMap* tm, const void* base, int C, int64_t pitch, int64_t bstride) -> int {
    if (p.spatial) {
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
      const uint64_t strides[3] = {(uint64_t)pitch * 2, (uint64_t)pitch * 2 * d->W, (uint64_t)bstride * 2};
      const uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1};
      return pir_make_tmap(tm, dt, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)d->H * d->W, (uint64_t)d->B};
    const uint64_t strides[2] = {(uint64_t)pitch * 2, (uint64_t)bstride * 2};
    const uint32_t box[3] = {64, (uint32_t)kWgPix, 1};
    return pir_make_tmap(tm, dt, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  if (int e = make(&tmA, d->a, d->M, d->a_pitch, d->a_bstride)) return e;
  if (int e = make(&tmB, d->b, d->N, d->b_pitch, d->b_bstride)) return e;
  if (!pir_smem_attr_once(reinterpret_cast<const void*>(wgrad_kernel<T>), (int)(220 * 1024), "pir_wgrad")) return PIR_ERR_CUDA;
  const int P = d->per_image ? d->B * d->splits : d->splits;
  dim3 grid((unsigned)(p.rblocks * p.cblocks), (unsigned)P, (unsigned)d->taps);
  if (p.swap) wgrad_kernel<T><<<grid, kWgThreads, smem, stream>>>(tmB, tmA, g);
  else wgrad_kernel<T><<<grid, kWgThreads, smem, stream>>>(tmA, tmB, g);
  return pir_check_launch("pir_wgrad");
}

}  // namespace pir

extern "C" int pir_wgrad_splits(int32_t B, int32_t HW, int32_t M, int32_t N, int32_t taps, int32_t per_image) {
  // One CTA per SM (the TMA ring fills shared memory): pick the split count that fills whole waves of 148 CTAs best while every
  // split keeps at least 4 chunks of 64 pixels.  (The chunk count of the spatial 3x3 path is >= this estimate.)
  if (B <= 0 || HW <= 0 || M <= 0 || N <= 0) return 1;
  const int plain = ((M + 127) / 128) * ((N + 255) / 256), swapped = ((N + 127) / 128) * ((M + 255) / 256);
  const int tiles = (swapped < plain ? swapped : plain) * (taps == 9 ? 9 : 1) * (per_image ? B : 1);
  const int cpi = (HW + 63) / 64;
  const int chunks = per_image ? cpi : B * cpi;
  int cap = chunks / 4;
  if (cap < 1) cap = 1;
  if (cap > 128) cap = 128;
  int best = 1;
  double best_u = 0.0;
  for (int s = 1; s <= cap; ++s) {
    const int total = tiles * s;
    if (total > 2 * 148 && s > 1) break;
    const int waves = (total + 147) / 148;
    const double u = (double)total / (148.0 * waves);
    if (u > best_u + 1e-9) { best_u = u; best = s; }
  }
  return best;
}

extern "C" int pir_wgrad(const PirWgrad* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_wgrad: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->M <= 0 || d->N <= 0 || d->splits <= 0) return pir_fail(PIR_ERR_ARG, "pir_wgrad: empty problem");
  if (d->taps != 1 && d->taps != 9) return pir_fail(PIR_ERR_ARG, "pir_wgrad: taps must be 1 or 9");
  if ((d->M % 8) || (d->N % 8) || (d->a_pitch % 8) || (d->b_pitch % 8) || (d->a_bstride % 8) || (d->b_bstride % 8) ||
      ((uintptr_t)d->a & 15) || ((uintptr_t)d->b & 15))
    return pir_fail(PIR_ERR_ARG, "pir_wgrad: channel counts / pitches must be multiples of 8 and pointers 16-byte aligned");
  if (!d->ws || ((uintptr_t)d->ws & 15)) return pir_fail(PIR_ERR_ARG, "pir_wgrad: workspace missing or misaligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return d->dtype == PIR_DTYPE_BF16 ? pir::launch_wgrad<pir::BF16>(d, s) : pir::launch_wgrad<pir::FP16>(d, s);
}

extern "C" int pir_wgrad_finalize(const PirWgradFin* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_wgrad_finalize: null descriptor");
  if (d->P <= 0 || d->R <= 0 || d->Cc <= 0 || d->M <= 0 || d->N <= 0 || !d->ws || !d->dst_w) return pir_fail(PIR_ERR_ARG, "pir_wgrad_finalize: empty problem");
  if (d->Cc > d->N || (d->R > d->half ? d->R - d->half + d->half_pad : d->R) > d->M) return pir_fail(PIR_ERR_ARG, "pir_wgrad_finalize: parameter larger than the partials");
  if (d->gamma && (d->taps != 1 || !d->w || !d->dst_gamma)) return pir_fail(PIR_ERR_ARG, "pir_wgrad_finalize: LayerNorm fold needs taps == 1, w and dst_gamma");
  if ((d->beta || d->dst_bias || d->dst_beta) && !d->colsum) return pir_fail(PIR_ERR_ARG, "pir_wgrad_finalize: beta / bias gradients need the column sums");
  if ((d->dst_beta != nullptr) != (d->beta != nullptr)) return pir_fail(PIR_ERR_ARG, "pir_wgrad_finalize: beta and dst_beta go together");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (d->gamma) {                                         // LayerNorm-folded 1x1 conv: reduce in place, then apply + column sums
    const long long MN = (long long)d->M * d->N;
    const int nmain = (int)((MN + 31) / 32), ncs = d->colsum ? (d->M + 31) / 32 : 0;
    pir::wgrad_reduce_inplace_kernel<<<(unsigned)(nmain + ncs), 256, 0, s>>>(d->ws, d->colsum, d->P, MN, d->M, nmain);
    if (int e = pir_check_launch("pir_wgrad_finalize(reduce)")) return e;
    pir::wgrad_ln_apply_kernel<<<(unsigned)((d->Cc + 7) / 8), 256, 0, s>>>(*d);
    return pir_check_launch("pir_wgrad_finalize(ln)");
  }
  const long long total = (long long)d->R * d->Cc;
  pir::wgrad_fin_kernel<<<(unsigned)((total + 31) / 32), 256, 0, s>>>(*d);
  if (int e = pir_check_launch("pir_wgrad_finalize")) return e;
  if (d->dst_bias) {
    pir::wgrad_fin_bias_kernel<<<(unsigned)((d->R + 255) / 256), 256, 0, s>>>(*d);
    return pir_check_launch("pir_wgrad_finalize(bias)");
  }
  return PIR_OK;
}
