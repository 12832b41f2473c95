// MDTA (multi-DConv-head transposed attention) core for sm_100a.           Reference: net/model.py:121-137.
//
// The attention matrix of MDTA is c x c over CHANNELS (c = C / heads); the pixel axis (up to 65 536) is only
// the contraction length.  So the work splits into
//   (1) pir_mdta_gram     : G[b] = Q[b]^T K[b] over the pixels -- a split-K tcgen05 GEMM whose operands are
//                           MN-major (channels contiguous) tiles of the NHWC qkv tensor, 64 pixels per stage,
//                           fp32 accumulation in TMEM; the spare epilogue warps accumulate the squared L2
//                           norms of every q / k channel from the same shared-memory stages.  Per-split
//                           partials go to a workspace (no atomics -> deterministic).
//   (2) pir_mdta_finalize : reduce partials, logits = G / (|q_i| |k_j|) * temperature, softmax over j,
//                           then fold into project_out:  Wf[b][o][hc+j] = sum_i Wo[o][hc+i] A[b,h][i][j]
//                           so that (attn @ v) and project_out become ONE pointwise GEMM on v (pir_gemm with
//                           per-image weights) -- the attention output is never materialised.  One clustered
//                           kernel for head dims <= 192 (mdta_finalize_fused_kernel), softmax + fold kernels above.
#include <cooperative_groups.h>
#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kGramThreads = 192;          // warp0 TMA, warp1 MMA, warps 2-5 norms + epilogue
constexpr int kGramPix = 64;               // pixels per stage (4 UMMA K-steps of 16)
constexpr int kGroupBytes = kGramPix * 128;  // one 64-channel x 64-pixel SW128 tile
constexpr int kGramMaxStages = 6;

struct GramArgs {
  int HW, C, heads;
  int splits, chunk;           // pixels per split (multiple of 64)
  int nblocks_n;               // column blocks (256 channels each)
  int stages;
  float* ws_gram;              // [B][splits][C][c]  (c = C / heads: only the heads' diagonal blocks)
  float* ws_norm;              // [B][splits][2][C]
};

template <class T>
__global__ void __launch_bounds__(kGramThreads)
mdta_gram_kernel(const __grid_constant__ CUtensorMap tmQKV, const GramArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kGramMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kGramMaxStages];
  __shared__ __align__(8) uint64_t bar_accum;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red[4][6][64];

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int mb = blockIdx.x / g.nblocks_n;
  const int nb = blockIdx.x % g.nblocks_n;
  const int split = blockIdx.y;
  const int b = blockIdx.z;
  const int m0 = mb * 128;
  const int n0 = nb * 256;
  const int cdim = g.C / g.heads;
  // skip blocks that do not touch any head's diagonal c x c block (block-uniform -> safe early exit)
  {
    const int m1 = min(m0 + 127, g.C - 1), n1 = min(n0 + 255, g.C - 1);
    const int hlo = max(m0 / cdim, n0 / cdim), hhi = min(m1 / cdim, n1 / cdim);
    if (hlo > hhi) return;
  }
  const int agroups = min(2, (g.C - m0 + 63) / 64);
  const int bgroups = min(4, (g.C - n0 + 63) / 64);
  const int block_n = bgroups * 64;
  const uint32_t stage_bytes = (uint32_t)(agroups + bgroups) * kGroupBytes;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int p_begin = split * g.chunk;
  const int p_end = min(p_begin + g.chunk, g.HW);
  const int nst = p_end > p_begin ? (p_end - p_begin + kGramPix - 1) / kGramPix : 0;
  const uint32_t tmem_cols = block_n <= 64 ? 64 : (block_n <= 128 ? 128 : 256);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 5);
    }
    mbar_init(smem_u32(&bar_accum), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_smem), tmem_cols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nst; ++it) {
      mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
      if (elect_one()) {
        const uint32_t full = smem_u32(&bar_full[stage]);
        const uint32_t dst = smem_base + (uint32_t)stage * stage_bytes;
        mbar_expect_tx(full, stage_bytes);
        const int p = p_begin + it * kGramPix;
        for (int gi = 0; gi < agroups; ++gi) tma_load_3d(dst + gi * kGroupBytes, &tmQKV, full, m0 + gi * 64, p, b);
        for (int gi = 0; gi < bgroups; ++gi)
          tma_load_3d(dst + (agroups + gi) * kGroupBytes, &tmQKV, full, g.C + n0 + gi * 64, p, b);
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
        const uint32_t b_src = a_src + (uint32_t)agroups * kGroupBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {       // 16 pixels (two 8-row swizzle atoms = 2048 B) per MMA
          const uint64_t ad = make_sdesc_sw128(a_src + k * 2048, kGroupBytes, 1024);
          const uint64_t bd = make_sdesc_sw128(b_src + k * 2048, kGroupBytes, 1024);
          umma_f16(tmem_base, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar_empty[stage]));
        if (it == nst - 1) umma_commit(smem_u32(&bar_accum));
      }
      __syncwarp();
      if (++stage == g.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ---- squared norms: warp w takes pixel rows [16w', 16w'+16) of every group.  A lane owns one 16-byte chunk
    //      (8 channels) of rows 4i + lane/8: all LDS.128 of a stage are issued before the FMAs that consume them ----
    const int wq = warp - 2;
    const int rsub = lane >> 3, ch8 = lane & 7;
    const int ngroups = agroups + bgroups;
    float ss[6][8];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) ss[i][e] = 0.f;
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint8_t* base = smem_raw + (smem_base - smem_u32(smem_raw));
      for (int it = 0; it < nst; ++it) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        const uint8_t* st = base + (size_t)stage * stage_bytes;
        uint4 v[6][4];
#pragma unroll
        for (int gi = 0; gi < 6; ++gi) {
          if (gi < ngroups) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = wq * 16 + i * 4 + rsub;
              v[gi][i] = *reinterpret_cast<const uint4*>(st + gi * kGroupBytes + row * 128 + ((ch8 ^ (row & 7)) << 4));
            }
          }
        }
#pragma unroll
        for (int gi = 0; gi < 6; ++gi) {
          if (gi < ngroups) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t w4[4] = {v[gi][i].x, v[gi][i].y, v[gi][i].z, v[gi][i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x = unpack_lo<T>(w4[e]), y = unpack_hi<T>(w4[e]);
                ss[gi][2 * e] = fmaf(x, x, ss[gi][2 * e]);
                ss[gi][2 * e + 1] = fmaf(y, y, ss[gi][2 * e + 1]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_empty[stage]));
        if (++stage == g.stages) { stage = 0; phase ^= 1u; }
      }
    }
#pragma unroll
    for (int gi = 0; gi < 6; ++gi) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = ss[gi][e];
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        if (rsub == 0) red[wq][gi][ch8 * 8 + e] = t;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");        // the four statistics warps only
    {
      float* nrm = g.ws_norm + ((size_t)(b * g.splits + split) * 2) * g.C;
      const int t = threadIdx.x - 64;                     // 0..127
      for (int e = t; e < ngroups * 64; e += 128) {
        const int gi = e >> 6, ch = e & 63;
        const float s = red[0][gi][ch] + red[1][gi][ch] + red[2][gi][ch] + red[3][gi][ch];
        if (gi < agroups) {
          const int cq = m0 + gi * 64 + ch;
          if (nb == m0 / 256 && cq < g.C) nrm[cq] = s;          // the column block holding this row block's diagonal
        } else {
          const int ck = n0 + (gi - agroups) * 64 + ch;
          if (mb == n0 / 128 && ck < g.C) nrm[g.C + ck] = s;    // the row block holding this column block's diagonal
        }
      }
    }
    // ---- epilogue: TMEM -> fp32 partial Gram, only the c x c diagonal block of the row's head is kept:
    //      ws_gram[b][split][i][j - head(i)*c]  ----
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int i = m0 + row;
    const int hi = i / cdim;
    const int jlo = hi * cdim, jhi = jlo + cdim;                       // this row's valid column range
    float* gout = g.ws_gram + ((size_t)(b * g.splits + split) * g.C + (size_t)i) * cdim;
    // columns any row of this warp needs (warp-uniform, so the aligned TMEM loads stay converged)
    const int wlo = ((m0 + quarter * 32) / cdim) * cdim;
    const int whi = (min(m0 + quarter * 32 + 31, g.C - 1) / cdim + 1) * cdim;
    if (nst > 0) {
      mbar_wait(smem_u32(&bar_accum), 0);
      tc_fence_after();
    }
    const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool vec4 = (cdim & 3) == 0;
    for (int c0 = 0; c0 < block_n; c0 += 16) {
      const int j0 = n0 + c0;
      if (j0 + 16 <= wlo || j0 >= whi || m0 + quarter * 32 >= g.C) continue;      // warp-uniform
      uint32_t acc[16];
      if (nst > 0) {
        tmem_ld16(taddr_row + (uint32_t)c0, acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = 0u;
      }
      if (i < g.C) {
        if (vec4) {
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const int j = j0 + q;
            if (j >= jlo && j + 3 < jhi)
              *reinterpret_cast<float4*>(gout + (j - jlo)) = make_float4(__uint_as_float(acc[q]), __uint_as_float(acc[q + 1]),
                                                                         __uint_as_float(acc[q + 2]), __uint_as_float(acc[q + 3]));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int j = j0 + q;
            if (j >= jlo && j < jhi) gout[j - jlo] = __uint_as_float(acc[q]);
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
// finalize, two-kernel path (head dims > 192, or PIR_MDTA_FUSED=0).  1: softmax rows.  grid (ceil(C/8), B), 256 threads:
// one warp per attention row (i = channel of q).
// ------------------------------------------------------------------------------------------------------
// Sums the split-K partials of four items at a time, eight loads in flight each (32 independent loads per thread and
// round); per item the additions happen in one fixed order (four accumulators over the multiple-of-4 prefix of the
// splits, the remainder onto the first), so every finalize path produces the same bits.  A null pointer skips the item.
__device__ __forceinline__ void sum_splits_x4(const float* const (&p)[4], const size_t (&stride)[4], int splits, float (&out)[4]) {
  float acc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x) acc[x][0] = acc[x][1] = acc[x][2] = acc[x][3] = 0.f;
  const int tail = splits & ~3;
  for (int sp = 0; sp < splits; sp += 8) {
    float l[4][8];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int u = 0; u < 8; ++u) l[x][u] = (p[x] != nullptr && sp + u < splits) ? p[x][(size_t)(sp + u) * stride[x]] : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool main = sp + u < tail;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        if (main) acc[x][u & 3] += l[x][u];
        else acc[x][0] += l[x][u];
      }
    }
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) out[x] = (acc[x][0] + acc[x][1]) + (acc[x][2] + acc[x][3]);
}

template <int NT>
__global__ void __launch_bounds__(256)
mdta_softmax_kernel(const float* __restrict__ ws_gram, const float* __restrict__ ws_norm, const float* __restrict__ temperature,
                    float* __restrict__ attn, int C, int heads, int splits) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;         // global q channel
  const int b = blockIdx.y;
  if (r >= C) return;
  const int c = C / heads;
  const int h = r / c, i = r - h * c;
  const float* gbase = ws_gram + ((size_t)b * splits * C + r) * c;             // + sp * C * c
  const float* nbase = ws_norm + (size_t)b * splits * 2 * C;                   // + sp * 2 * C
  const float temp = temperature[h];
  // c <= 32 * NT: up to NT columns per lane (NT = 8 for the usual 48-wide heads, 24 for the single-head prompt blocks).
  // Two columns (their k norm and Gram sums) per batch of loads.
  float v[NT];
  float mx = -INFINITY;
  float qn = 1.f;
#pragma unroll
  for (int t = 0; t < NT; t += 2) {
    const int j0 = lane + t * 32, j1 = lane + (t + 1) * 32;
    const bool ok0 = j0 < c, ok1 = t + 1 < NT && j1 < c;
    if (t * 32 >= c) {                                                  // whole warp past the head dimension
      v[t] = -INFINITY;
      if (t + 1 < NT) v[t + 1] = -INFINITY;
      continue;
    }
    const float* p[4] = {ok0 ? nbase + C + h * c + j0 : nullptr, ok0 ? gbase + j0 : nullptr,
                         ok1 ? nbase + C + h * c + j1 : nullptr, ok1 ? gbase + j1 : nullptr};
    const size_t st[4] = {(size_t)2 * C, (size_t)C * c, (size_t)2 * C, (size_t)C * c};
    float out[4];
    if (t == 0) {                                                     // q norm: every lane of the row needs it
      const float* pq[4] = {nbase + r, nullptr, nullptr, nullptr};
      float oq[4];
      sum_splits_x4(pq, st, splits, oq);
      qn = fmaxf(sqrtf(oq[0]), 1e-12f);                               // F.normalize: x / max(||x||, eps)
    }
    sum_splits_x4(p, st, splits, out);
    v[t] = -INFINITY;
    if (ok0) { v[t] = out[1] / (qn * fmaxf(sqrtf(out[0]), 1e-12f)) * temp; mx = fmaxf(mx, v[t]); }
    if (t + 1 < NT) {
      v[t + 1] = -INFINITY;
      if (ok1) { v[t + 1] = out[3] / (qn * fmaxf(sqrtf(out[2]), 1e-12f)) * temp; mx = fmaxf(mx, v[t + 1]); }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int j = lane + t * 32;
    if (j < c) { v[t] = expf(v[t] - mx); sum += v[t]; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  float* arow = attn + ((size_t)(b * heads + h) * c + i) * c;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int j = lane + t * 32;
    if (j < c) arow[j] = v[t] * inv;
  }
}

// ------------------------------------------------------------------------------------------------------
// finalize 2: Wf[b][o][h*c + j] = sum_i Wo[o][h*c + i] * A[b,h][i][j].  32x32 output tiles, K chunks of 32.
// grid (ceil(C/32) o-tiles, heads * ceil(c/32) j-tiles, B), block (32, 8).
// ------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
mdta_fold_kernel(const float* __restrict__ wo, const float* __restrict__ attn, unsigned short* __restrict__ wfold, int C,
                 int heads, int kpad) {
  __shared__ float sW[32][33];
  __shared__ float sA[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int c = C / heads;
  const int jt_per_head = (c + 31) / 32;
  const int h = blockIdx.y / jt_per_head;
  const int j0 = (blockIdx.y % jt_per_head) * 32;
  const int o0 = blockIdx.x * 32;
  const int b = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float* A = attn + (size_t)(b * heads + h) * c * c;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i0 = 0; i0 < c; i0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int rr = ty + r * 8;
      const int o = o0 + rr, i = i0 + tx;
      sW[rr][tx] = (o < C && i < c) ? wo[(size_t)o * C + h * c + i] : 0.f;
      const int ii = i0 + rr, j = j0 + tx;
      sA[rr][tx] = (ii < c && j < c) ? A[(size_t)ii * c + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a = sA[k][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sW[ty + r * 8][k], a, acc[r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int o = o0 + ty + r * 8, j = j0 + tx;
    if (o < C && j < c) wfold[((size_t)b * C + o) * kpad + h * c + j] = to16<T>(acc[r]);
  }
}

// ------------------------------------------------------------------------------------------------------
// finalize, fused (head dim <= 192): one CTA per (o-chunk, head, image) reduces the head's c x c Gram block and
// norms into shared memory, runs the softmax there (one warp per row; same arithmetic and summation order as
// mdta_softmax_kernel, so the two paths agree bit for bit), writes the attention matrix for the backward pass and
// folds it into Wo:  Wf[b][o][h*c + j] = sum_i Wo[o][h*c + i] * A[i][j]  for its o rows.  One launch instead of two and
// no 32x32-tile round trips through L2: lanes run over j, each warp keeps 8 o rows in registers, Wo tiles are staged
// transposed so the 8 weights of a k step are two broadcast float4 reads.
// ------------------------------------------------------------------------------------------------------
constexpr int kFinThreads = 512;
constexpr int kFinRows = (kFinThreads / 32) * 8;   // o rows per pass
constexpr int kFinIC = 64;                         // k (= i) chunk staged per pass
constexpr int kFinWPitch = kFinRows + 4;

template <class T, int NT>
__global__ void __launch_bounds__(kFinThreads)
mdta_finalize_fused_kernel(const float* __restrict__ ws_gram, const float* __restrict__ ws_norm,
                           const float* __restrict__ temperature, const float* __restrict__ wo, float* __restrict__ attn,
                           unsigned short* __restrict__ wfold, int C, int heads, int splits, int kpad, int o_per_cta) {
  extern __shared__ __align__(16) float fin_smem[];
  constexpr int kPre = kFinRows * kFinIC / kFinThreads;              // Wo values each thread stages per chunk
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int nrank = (int)gridDim.x, rank = (int)blockIdx.x;          // the cluster spans grid.x: one CTA per o-chunk
  const int c = C / heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  float* sG = fin_smem;                                              // [c][c] Gram sums, then probabilities
  float* sN = sG + c * c;                                            // [2c]: ||q_i||^2, ||k_j||^2; then 32 * NT floats of slack (masked lanes)
  float* sW = sG + ((c * c + 2 * c + 32 * NT + 3) & ~3);             // [kFinIC][kFinWPitch], 16-byte aligned
  const int o_begin = rank * o_per_cta;
  const int o_end = o_begin + o_per_cta < C ? o_begin + o_per_cta : C;

  // Wo tile (rows ob.., k chunk i0..) -> registers; it only depends on weights, so the first one is fetched before
  // anything else and each next one while the current chunk is being multiplied.
  float pre[kPre];
  auto fetch_w = [&](int ob, int i0) {
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int e = tid + u * kFinThreads;
      const int oo = e / kFinIC, ii = e % kFinIC;
      const int o = ob + oo, i = i0 + ii;
      pre[u] = (o < o_end && i < c) ? wo[(size_t)o * C + h * c + i] : 0.f;
    }
  };
  fetch_w(o_begin, 0);
  pdl_launch_dependents();
  pdl_wait();

  // ---- 1a: split-K partials of this head's Gram block and norms -> shared memory.  One SM pulls only ~40 GB/s out of
  // L2 with this access pattern, so the c*c + 2c sums are divided over the CTAs of the cluster (item e lives at sG[e] in
  // its owner) and exchanged through distributed shared memory. ----
  const int ng = c * c, n = ng + 2 * c;
  const int per = (n + nrank - 1) / nrank;
  {
    const size_t gstride = (size_t)C * c;
    const float* gb = ws_gram + ((size_t)b * splits * C + (size_t)h * c) * c;
    const float* nb = ws_norm + (size_t)b * splits * 2 * C;
    const int lo = rank * per, hi = lo + per < n ? lo + per : n;
    for (int e0 = lo + tid; e0 < hi; e0 += 4 * kFinThreads) {
      const float* p[4];
      size_t st[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int e = e0 + x * kFinThreads;
        p[x] = nullptr; st[x] = 0;
        if (e < hi && e < ng) { p[x] = gb + e; st[x] = gstride; }
        else if (e < hi) {
          const int t = e - ng;
          p[x] = nb + (t < c ? h * c + t : C + h * c + (t - c)); st[x] = (size_t)2 * C;
        }
      }
      float out[4];
      sum_splits_x4(p, st, splits, out);
#pragma unroll
      for (int x = 0; x < 4; ++x)
        if (p[x] != nullptr) sG[e0 + x * kFinThreads] = out[x];
    }
  }
  if (nrank > 1) {
    cluster.sync();
    const int lo = rank * per, hi = lo + per;
    for (int e0 = tid; e0 < n; e0 += 4 * kFinThreads) {               // four independent remote loads in flight per thread
      float v[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int e = e0 + x * kFinThreads;
        v[x] = (e < n && (e < lo || e >= hi)) ? cluster.map_shared_rank(sG, e / per)[e] : 0.f;
      }
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int e = e0 + x * kFinThreads;
        if (e < n && (e < lo || e >= hi)) sG[e] = v[x];
      }
    }
    cluster.sync();                                                   // nobody overwrites (1b) or leaves before all peers have read
  } else {
    __syncthreads();
  }
  // ---- 1b: softmax rows, in place ----
  {
    const float temp = temperature[h];
    for (int i = warp; i < c; i += kFinThreads / 32) {
      const float qn = fmaxf(sqrtf(sN[i]), 1e-12f);                  // F.normalize: x / max(||x||, eps)
      float v[NT];
      float mx = -INFINITY;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int j = lane + t * 32;
        v[t] = -INFINITY;
        if (j < c) {
          const float kn = fmaxf(sqrtf(sN[c + j]), 1e-12f);
          v[t] = sG[i * c + j] / (qn * kn) * temp;
          mx = fmaxf(mx, v[t]);
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int j = lane + t * 32;
        if (j < c) { v[t] = expf(v[t] - mx); sum += v[t]; }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.0f / sum;
      float* arow = attn + ((size_t)(b * heads + h) * c + i) * c;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int j = lane + t * 32;
        if (j < c) {
          const float p = v[t] * inv;
          sG[i * c + j] = p;
          if (blockIdx.x == 0) arow[j] = p;          // kept for pir_mdta_bwd
        }
      }
    }
  }
  // ---- 2: fold into Wo ----
  for (int ob = o_begin; ob < o_end; ob += kFinRows) {
    float acc[8][NT];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[r][t] = 0.f;
    for (int i0 = 0; i0 < c; i0 += kFinIC) {
      __syncthreads();                              // sG complete (first pass) / previous sW tile consumed
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int e = tid + u * kFinThreads;
        sW[(e % kFinIC) * kFinWPitch + e / kFinIC] = pre[u];
      }
      __syncthreads();
      if (i0 + kFinIC < c) fetch_w(ob, i0 + kFinIC);
      else if (ob + kFinRows < o_end) fetch_w(ob + kFinRows, 0);
      if (ob + warp * 8 < o_end) {
        const int imax = c - i0 < kFinIC ? c - i0 : kFinIC;
#pragma unroll 4
        for (int ii = 0; ii < imax; ++ii) {
          const float4 w0 = *reinterpret_cast<const float4*>(sW + ii * kFinWPitch + warp * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(sW + ii * kFinWPitch + warp * 8 + 4);
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          float a[NT];
#pragma unroll
          for (int t = 0; t < NT; ++t) a[t] = sG[(i0 + ii) * c + lane + t * 32];       // lanes with j >= c read slack: never stored
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[r][t] = fmaf(w[r], a[t], acc[r][t]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int o = ob + warp * 8 + r;
      if (o < o_end) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int j = lane + t * 32;
          if (j < c) wfold[((size_t)b * C + o) * kpad + h * c + j] = to16<T>(acc[r][t]);
        }
      }
    }
  }
}

static size_t fin_smem_bytes(int c, int nt) {
  return ((size_t)((c * c + 2 * c + 32 * nt + 3) & ~3) + (size_t)kFinIC * kFinWPitch) * sizeof(float);
}

static bool fin_fused_ok(int C, int heads) {
  static const bool off = [] { const char* e = getenv("PIR_MDTA_FUSED"); return e && e[0] == '0'; }();   // A/B: softmax + fold kernels
  return !off && heads > 0 && C % heads == 0 && C / heads <= 192;
}

template <class T, int NT>
static int launch_fin_fused(const PirMdta* d, cudaStream_t s, const float* ws_gram, const float* ws_norm, float* attn) {
  const int c = d->C / d->heads;
  const size_t smem = fin_smem_bytes(c, NT);
  if (smem > 220 * 1024) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_mdta_finalize: head dim %d needs %zu bytes of shared memory", c, smem);
  if (!pir_smem_attr_once(reinterpret_cast<const void*>(mdta_finalize_fused_kernel<T, NT>), 220 * 1024, "pir_mdta_finalize")) return PIR_ERR_CUDA;
  // one cluster per (head, image); its CTAs share the split-K reduction and take C / size output rows each
  // (a multiple of 8 = one warp's rows).  Enough CTAs to cover the SMs when B * heads alone does not.
  const int pairs = d->B * d->heads;
  // Clusters of 8 only when a handful of them exist (not every GPC can place two at a time: 16 clusters of 8 ran in two
  // waves, 20 us instead of 14); otherwise up to 4 CTAs per cluster while the grid stays within one wave.
  int cl = 1;
  while (cl < 4 && pairs * cl * 2 <= 148) cl *= 2;
  if (pairs <= 8) cl = 8;
  const int o_per_cta = ((d->C + cl - 1) / cl + 7) / 8 * 8;
  const int kpad = (d->C + 63) / 64 * 64;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cl, (unsigned)d->heads, (unsigned)d->B);
  cfg.blockDim = dim3(kFinThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pir_pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, mdta_finalize_fused_kernel<T, NT>, ws_gram, ws_norm, d->temperature, d->wo, attn,
                     reinterpret_cast<unsigned short*>(d->wfold), (int)d->C, (int)d->heads, (int)d->splits, kpad, o_per_cta);
  return pir_check_launch("pir_mdta_finalize(fused)");
}

template <class T>
static int launch_fin_fused_t(const PirMdta* d, cudaStream_t s, const float* ws_gram, const float* ws_norm, float* attn) {
  const int c = d->C / d->heads;
  if (c <= 64) return launch_fin_fused<T, 2>(d, s, ws_gram, ws_norm, attn);
  if (c <= 96) return launch_fin_fused<T, 3>(d, s, ws_gram, ws_norm, attn);
  return launch_fin_fused<T, 6>(d, s, ws_gram, ws_norm, attn);
}

// ------------------------------------------------------------------------------------------------------
static int gram_nblocks_n(int C) { return (C + 255) / 256; }
static int gram_nblocks_m(int C) { return (C + 127) / 128; }

template <class T>
static int launch_gram(const PirMdta* d, cudaStream_t stream) {
  GramArgs g{};
  g.HW = d->HW; g.C = d->C; g.heads = d->heads; g.splits = d->splits;
  g.chunk = ((d->HW + d->splits - 1) / d->splits + kGramPix - 1) / kGramPix * kGramPix;
  g.nblocks_n = gram_nblocks_n(d->C);
  g.ws_gram = d->ws;
  g.ws_norm = d->ws + (size_t)d->B * d->splits * d->C * (d->C / d->heads);
  const int groups = (d->C >= 128 ? 2 : (d->C + 63) / 64) + (d->C >= 256 ? 4 : (d->C + 63) / 64);
  const size_t stage_bytes = (size_t)groups * kGroupBytes;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > kGramMaxStages) stages = kGramMaxStages;
  const int max_it = g.chunk / kGramPix;
  if (stages > max_it) stages = max_it < 1 ? 1 : max_it;
  g.stages = stages;
  const size_t smem = stages * stage_bytes + 1024;

  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const uint64_t dims[3] = {(uint64_t)3 * d->C, (uint64_t)d->HW, (uint64_t)d->B};
  const uint64_t strides[2] = {(uint64_t)d->qkv_pitch * 2, (uint64_t)d->qkv_bstride * 2};
  const uint32_t box[3] = {64, (uint32_t)kGramPix, 1};
  CUtensorMap tm;
  if (int e = pir_make_tmap(&tm, dt, 3, d->qkv, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  if (!pir_smem_attr_once(reinterpret_cast<const void*>(mdta_gram_kernel<T>), (int)(220 * 1024), "pir_mdta_gram")) return PIR_ERR_CUDA;
  dim3 grid((unsigned)(gram_nblocks_m(d->C) * g.nblocks_n), (unsigned)d->splits, (unsigned)d->B);
  pir_launch(mdta_gram_kernel<T>, grid, dim3(kGramThreads), smem, stream, tm, g);
  return pir_check_launch("pir_mdta_gram");
}

static int check_mdta(const PirMdta* d, const char* who) {
  if (!d) return pir_fail(PIR_ERR_ARG, "%s: null descriptor", who);
  if (d->B <= 0 || d->HW <= 0 || d->C <= 0 || d->heads <= 0 || d->splits <= 0) return pir_fail(PIR_ERR_ARG, "%s: empty problem", who);
  if (d->C % d->heads) return pir_fail(PIR_ERR_ARG, "%s: C must be divisible by heads", who);
  if (d->C / d->heads > 768) return pir_fail(PIR_ERR_UNSUPPORTED, "%s: head dim > 768", who);
  if ((d->C % 8) || (d->qkv_pitch % 8) || (d->qkv_bstride % 8) || ((uintptr_t)d->qkv & 15)) return pir_fail(PIR_ERR_ARG, "%s: qkv not 16-byte aligned", who);
  if (!d->ws) return pir_fail(PIR_ERR_ARG, "%s: workspace missing", who);
  return PIR_OK;
}

}  // namespace pir

extern "C" int pir_mdta_splits(int32_t B, int32_t HW, int32_t C) {
  // The Gram kernel runs one CTA per SM (its TMA ring fills shared memory), so pick the split-K factor that
  // fills whole waves of 148 CTAs best; each split keeps at least 256 pixels.
  const int units = B * ((C + 127) / 128) * ((C + 255) / 256);          // CTAs per split (upper bound)
  const int cap = HW / 256 > 1 ? (HW / 256 > 64 ? 64 : HW / 256) : 1;
  int best = 1;
  double best_u = 0.0;
  for (int s = 1; s <= cap; ++s) {
    const int total = units * s;
    if (total > 2 * 148 && s > 1) break;
    const int waves = (total + 147) / 148;
    const double u = (double)total / (148.0 * waves);
    if (u > best_u + 1e-9) { best_u = u; best = s; }
  }
  return best;
}

extern "C" int64_t pir_mdta_ws_floats(int32_t B, int32_t C, int32_t splits) {
  // partial Grams + partial norms + attention matrices (upper bound: heads = 1)
  return (int64_t)B * splits * ((int64_t)C * C + 2 * C) + (int64_t)B * C * C;
}

extern "C" int pir_mdta_gram(const PirMdta* d, void* stream) {
  if (int e = pir::check_mdta(d, "pir_mdta_gram")) return e;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return d->dtype == PIR_DTYPE_BF16 ? pir::launch_gram<pir::BF16>(d, s) : pir::launch_gram<pir::FP16>(d, s);
}

extern "C" int pir_mdta_finalize_kernels(int32_t C, int32_t heads) { return pir::fin_fused_ok(C, heads) ? 1 : 2; }

extern "C" int pir_mdta_finalize(const PirMdta* d, void* stream) {
  if (int e = pir::check_mdta(d, "pir_mdta_finalize")) return e;
  if (!d->temperature || !d->wo || !d->wfold) return pir_fail(PIR_ERR_ARG, "pir_mdta_finalize: missing pointers");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int c = d->C / d->heads;
  const float* ws_gram = d->ws;
  const float* ws_norm = d->ws + (size_t)d->B * d->splits * d->C * c;
  float* attn = d->ws + (size_t)d->B * d->splits * ((size_t)d->C * c + 2 * d->C);
  if (pir::fin_fused_ok(d->C, d->heads))
    return d->dtype == PIR_DTYPE_BF16 ? pir::launch_fin_fused_t<pir::BF16>(d, s, ws_gram, ws_norm, attn)
                                      : pir::launch_fin_fused_t<pir::FP16>(d, s, ws_gram, ws_norm, attn);
  if (c <= 256) pir_launch(pir::mdta_softmax_kernel<8>, dim3((d->C + 7) / 8, d->B), dim3(256), 0, s, ws_gram, ws_norm, d->temperature, attn, d->C, d->heads, d->splits);
  else pir_launch(pir::mdta_softmax_kernel<24>, dim3((d->C + 7) / 8, d->B), dim3(256), 0, s, ws_gram, ws_norm, d->temperature, attn, d->C, d->heads, d->splits);
  if (int e = pir_check_launch("pir_mdta_finalize(softmax)")) return e;
  const int kpad = (d->C + 63) / 64 * 64;
  dim3 grid((d->C + 31) / 32, d->heads * ((c + 31) / 32), d->B);
  if (d->dtype == PIR_DTYPE_BF16)
    pir_launch(pir::mdta_fold_kernel<pir::BF16>, grid, dim3(32, 8), 0, s, d->wo, attn, reinterpret_cast<unsigned short*>(d->wfold), d->C, d->heads, kpad);
  else
    pir_launch(pir::mdta_fold_kernel<pir::FP16>, grid, dim3(32, 8), 0, s, d->wo, attn, reinterpret_cast<unsigned short*>(d->wfold), d->C, d->heads, kpad);
  return pir_check_launch("pir_mdta_finalize(fold)");
}
