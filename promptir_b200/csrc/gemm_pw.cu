// Persistent pointwise (1x1 conv) GEMM for sm_100a with NHWC 16-bit output -- the kernel behind K1/K4/K5/K7
// and the standalone reduce convs (net/model.py:88,92,111,113,294-313 plus the LayerNorm :60-63 and the
// residual adds :193-194).  Same math as gemm_tcgen05.cu, restructured so the HBM pipe never drains:
//
//   * one CTA per SM, each walking its share of the 128-pixel tiles (static round-robin);
//   * the weight matrix stays RESIDENT in shared memory when it fits (all level-1/2 layers), otherwise it is
//     streamed with the activations through the TMA ring;
//   * accumulators are double buffered in TMEM (2 x 256 fp32 columns): the MMA warp fills one buffer while the
//     epilogue warps drain the other;
//   * the epilogue converts to 16-bit into 128B-swizzled staging slabs and leaves through TMA stores (fully
//     coalesced, asynchronous, clipped at the tensor edge); the residual comes in the same way -- a dedicated
//     warp prefetches residual slabs by TMA into the very staging slot the result is written back to;
//   * LayerNorm statistics are accumulated by four dedicated warps from the activation stages already in shared
//     memory and handed to the epilogue through a double-buffered smem mailbox.
//
// Warp roles (480 threads): 0 TMA producer | 1 TMEM alloc + MMA issue | 2-5 LN statistics |
//   6-9 and 10-13 two independent epilogue warpgroups (alternate 64-column slabs, own staging slots) |
//   14 residual prefetch.  (One warpgroup alone is latency bound: TMEM load -> FMA -> pack -> STS chains.)
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kPwThreads = 480;
constexpr int kPwBlockM = 128;
constexpr int kPwBlockK = 64;
constexpr int kPwMaxStages = 8;
constexpr int kPwMaxSlots = 4;                     // staging slabs of 128 rows x 64 columns (16 KB): 2 per epilogue group
constexpr uint32_t kPwATile = kPwBlockM * kPwBlockK * 2;
constexpr uint32_t kPwSlab = 128 * 128;

struct PwArgs {
  int hw, B, N, K;
  int nkb;            // k-blocks
  int n_chunks, NC;   // column chunks of NC (<= 256) accumulator columns; NC % 64 == 0 when n_chunks > 1
  int stages;
  int resident;       // weights resident in smem
  int w_batched;
  int ln_mode;
  int has_res;
  int res_pf;         // residual tiles prefetched into L2 ahead of the slot loads (0 = off; PIR_GEMM_RESPF)
  int rotate;         // streamed weights: per-CTA rotation of the column-chunk order (PIR_GEMM_ROTATE=0 turns it off)
  int m_tiles;        // per image
  int n_alloc;        // weight rows held per k-block: ceil(N16 / 64) * 64
  uint32_t off_ring, off_slots, off_vec, off_stats;   // byte offsets from the 1024-aligned smem base
  const float* ln_s;
  const float* vec_t;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// L2 prefetch of a tensor box (no shared memory, no completion): used to pull a residual tile towards L2 a few tiles ahead
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(m), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// This kernel is memory bound: its warps mostly wait, and a polling loop is pure power (the cfg2 step runs into the software power
// cap).  Latency-tolerant waits (TMA producer, residual prefetch) poll every 256 ns, the others every 32 ns.
__device__ __forceinline__ void pw_wait_lazy(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) __nanosleep(256); }
__device__ __forceinline__ void pw_wait_nap(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) __nanosleep(32); }

template <class T>
__global__ void __launch_bounds__(kPwThreads, 1)
gemm_pw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const PwArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kPwMaxStages], bar_empty[kPwMaxStages];
  __shared__ __align__(8) uint64_t bar_bfull;
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ __align__(8) uint64_t bar_sfull[2], bar_sempty[2];          // LN statistics mailbox
  __shared__ __align__(8) uint64_t bar_slot_full[kPwMaxSlots], bar_slot_empty[kPwMaxSlots];
  __shared__ uint32_t tmem_base_smem;

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base + g.off_ring;
  const uint32_t slots = base + g.off_slots;
  float* svec = reinterpret_cast<float*>(base_ptr + g.off_vec);           // [2][n_alloc]: ln_s, vec_t
  float2* sstats = reinterpret_cast<float2*>(base_ptr + g.off_stats);     // [2][128]: (-rstd*mu, rstd)
  const uint32_t b_chunk_bytes = (uint32_t)((g.NC + 63) / 64) * 8192u;   // whole 64-row boxes
  const uint32_t stage_bytes = kPwATile + (g.resident ? 0u : b_chunk_bytes);
  const int total_tiles = g.B * g.m_tiles;
  // Streamed (non-resident) weights: every CTA walks the column chunks of a tile in its own rotation, so at any moment the CTAs
  // pull DIFFERENT parts of the weight matrix out of L2 instead of all 128-148 of them hammering the same lines (the level-4
  // GEMMs stream a 0.9-1.6 MB matrix per CTA and ran at ~35 GB/s per SM).  All roles use the same order.
  const int rot = (g.resident || g.rotate == 0) ? 0 : (int)(blockIdx.x % (unsigned)g.n_chunks);
  auto chunk_of = [&](int c) { const int cc = c + rot; return cc >= g.n_chunks ? cc - g.n_chunks : cc; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmO);
    if (g.has_res) tma_prefetch_desc(&tmR);
    const uint32_t empty_count = 1u + (g.ln_mode ? 4u : 0u);
    for (int s = 0; s < g.stages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), empty_count); }
    mbar_init(smem_u32(&bar_bfull), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), 2);
      mbar_init(smem_u32(&bar_sfull[i]), 128); mbar_init(smem_u32(&bar_sempty[i]), 256);
    }
    for (int i = 0; i < kPwMaxSlots; ++i) { mbar_init(smem_u32(&bar_slot_full[i]), 1); mbar_init(smem_u32(&bar_slot_empty[i]), 1); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < g.n_alloc; i += kPwThreads) {
    svec[i] = (g.ln_s && i < g.N) ? g.ln_s[i] : 0.f;
    svec[g.n_alloc + i] = (g.vec_t && i < g.N) ? g.vec_t[i] : 0.f;
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();                                    // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ========================================= TMA producer ==========================================
    if (g.resident && elect_one()) {
      const uint32_t bf = smem_u32(&bar_bfull);
      mbar_expect_tx(bf, (uint32_t)g.nkb * (uint32_t)g.n_alloc * 128u);
      for (int kb = 0; kb < g.nkb; ++kb)
        for (int r = 0; r < g.n_alloc; r += 64)
          tma_load_3d(base + ((uint32_t)kb * g.n_alloc + (uint32_t)r) * 128u, &tmB, bf, kb * kPwBlockK, r, 0);
    }
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int b = t / g.m_tiles, m0 = (t % g.m_tiles) * kPwBlockM;
      const int reps = g.resident ? 1 : g.n_chunks;
      for (int ci = 0; ci < reps; ++ci) {
        const int c = chunk_of(ci);
        const int ncols = min(g.NC, ((g.N + 15) / 16 * 16) - c * g.NC);
        const int bboxes = g.resident ? 0 : (ncols + 63) / 64;
        for (int kb = 0; kb < g.nkb; ++kb) {
          pw_wait_lazy(smem_u32(&bar_empty[stage]), phase ^ 1u);
          if (elect_one()) {
            const uint32_t full = smem_u32(&bar_full[stage]);
            const uint32_t dst = ring + (uint32_t)stage * stage_bytes;
            mbar_expect_tx(full, kPwATile + (uint32_t)bboxes * 8192u);
            tma_load_3d(dst, &tmA, full, kb * kPwBlockK, m0, b);
            for (int r = 0; r < bboxes; ++r)
              tma_load_3d(dst + kPwATile + r * 8192u, &tmB, full, kb * kPwBlockK, c * g.NC + r * 64, g.w_batched ? b : 0);
          }
          __syncwarp();
          if (++stage == g.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ========================================= MMA issuer ============================================
    if (g.resident) mbar_wait(smem_u32(&bar_bfull), 0);
    int stage = 0; uint32_t phase = 0;      // ring position of the current tile's first k-block
    uint32_t q = 0;                          // accumulator chunk counter
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      for (int ci = 0; ci < g.n_chunks; ++ci, ++q) {
        const int c = chunk_of(ci);
        const uint32_t buf = q & 1u;
        const int ncols = min(g.NC, ((g.N + 15) / 16 * 16) - c * g.NC);
        const uint32_t idesc = make_idesc_f16(T::kFmt, kPwBlockM, ncols, 0, 0);
        pw_wait_nap(smem_u32(&bar_tempty[buf]), ((q >> 1) & 1u) ^ 1u);
        tc_fence_after();
        int st = stage; uint32_t ph = phase;
        for (int kb = 0; kb < g.nkb; ++kb) {
          pw_wait_nap(smem_u32(&bar_full[st]), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_src = ring + (uint32_t)st * stage_bytes;
            const uint32_t b_src = g.resident ? base + ((uint32_t)kb * g.n_alloc + (uint32_t)c * g.NC) * 128u : a_src + kPwATile;
            const int rem = g.K - kb * kPwBlockK;
            const int ksteps = rem >= kPwBlockK ? 4 : (rem + 15) >> 4;
            for (int k = 0; k < ksteps; ++k)
              umma_f16(tmem_base + buf * 256u, make_sdesc_sw128(a_src + k * 32, 16, 1024), make_sdesc_sw128(b_src + k * 32, 16, 1024),
                       idesc, (kb | k) != 0 ? 1u : 0u);
            if (!g.resident || ci == g.n_chunks - 1) umma_commit(smem_u32(&bar_empty[st]));   // stage no longer needed
            if (kb == g.nkb - 1) umma_commit(smem_u32(&bar_tfull[buf]));
          }
          __syncwarp();
          if (++st == g.stages) { st = 0; ph ^= 1u; }
        }
        if (!g.resident || ci == g.n_chunks - 1) { stage = st; phase = ph; }     // resident: chunks re-walk the same stages
      }
    }
  } else if (warp < 6) {
    // ========================================= LN statistics ==========================================
    if (g.ln_mode) {
      const int row = (warp & 3) * 32 + lane;
      int stage = 0; uint32_t phase = 0;
      uint32_t ti = 0;
      const float inv_k = 1.0f / (float)g.K;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
        float s1 = 0.f, s2 = 0.f;
        const int reps = g.resident ? 1 : g.n_chunks;
        for (int c = 0; c < reps; ++c) {
          for (int kb = 0; kb < g.nkb; ++kb) {
            pw_wait_nap(smem_u32(&bar_full[stage]), phase);
            if (c == 0) {
              const uint8_t* a_src = base_ptr + g.off_ring + (size_t)stage * stage_bytes + (size_t)row * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 v = *reinterpret_cast<const uint4*>(a_src + ((j ^ (row & 7)) << 4));
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
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_empty[stage]));
            if (++stage == g.stages) { stage = 0; phase ^= 1u; }
          }
          if (c == 0) {
            const float mu = s1 * inv_k;
            const float rstd = rsqrtf(fmaxf(fmaf(s2, inv_k, -mu * mu), 0.f) + 1e-5f);
            const uint32_t mb = ti & 1u;
            pw_wait_nap(smem_u32(&bar_sempty[mb]), ((ti >> 1) & 1u) ^ 1u);
            sstats[mb * 128 + row] = make_float2(g.ln_mode == 2 ? 0.f : -rstd * mu, rstd);
            mbar_arrive(smem_u32(&bar_sfull[mb]));          // release semantics order the store above
          }
        }
      }
    }
  } else if (warp < 14) {
    // ========================================= epilogue (two warpgroups) ==============================
    const int grp = (warp - 6) >> 2;                 // slabs with (u & 1) == grp belong to this group
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool lead_warp = ((warp - 6) & 3) == 0;     // its elected lane issues the TMA stores (always the same lane)
    const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int n16 = (g.N + 15) / 16 * 16;
    uint32_t q = 0, u = 0, ti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const int b = t / g.m_tiles, m0 = (t % g.m_tiles) * kPwBlockM;
      float a_scale = 1.f, a_shift = 0.f;
      if (g.ln_mode) {
        const uint32_t mb = ti & 1u;
        pw_wait_nap(smem_u32(&bar_sfull[mb]), (ti >> 1) & 1u);
        const float2 st = sstats[mb * 128 + row];
        a_shift = st.x; a_scale = st.y;
        mbar_arrive(smem_u32(&bar_sempty[mb]));
      }
      for (int ci = 0; ci < g.n_chunks; ++ci, ++q) {
        const int c = chunk_of(ci);
        const uint32_t buf = q & 1u;
        const int ncols = min(g.NC, n16 - c * g.NC);
        // both groups observe every chunk (even one whose slabs all belong to the other group) so neither can run
        // more than one mbarrier phase ahead on bar_tempty
        pw_wait_nap(smem_u32(&bar_tfull[buf]), (q >> 1) & 1u);
        tc_fence_after();
        for (int s0 = 0; s0 < ncols; s0 += 64, ++u) {
          if ((int)(u & 1u) != grp) continue;
          const uint32_t j = u >> 1;
          const uint32_t slot = (uint32_t)grp * 2u + (j & 1u), sph = (j >> 1) & 1u;
          uint32_t acc[64];
          const int ncol0 = c * g.NC + s0;
          const int nvalid = min(64, ncols - s0);            // multiple of 16
          if (nvalid == 64) {
            tmem_ld64(taddr_row + buf * 256u + (uint32_t)s0, acc);
          } else {
#pragma unroll
            for (int part = 0; part < 4; ++part) {
              if (part * 16 < nvalid) {
                uint32_t a16[16];
                tmem_ld16(taddr_row + buf * 256u + (uint32_t)(s0 + part * 16), a16);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[part * 16 + i] = a16[i];
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[part * 16 + i] = 0u;
              }
            }
          }
          if (g.has_res) pw_wait_nap(smem_u32(&bar_slot_full[slot]), sph);
          else pw_wait_nap(smem_u32(&bar_slot_empty[slot]), sph ^ 1u);
          tmem_ld_wait();
          uint8_t* srow = base_ptr + g.off_slots + (size_t)slot * kPwSlab + (size_t)row * 128;
          const float4* sv4 = reinterpret_cast<const float4*>(svec + ncol0);
          const float4* tv4 = reinterpret_cast<const float4*>(svec + g.n_alloc + ncol0);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {                    // 8 columns = one 16-byte chunk of the slab row
            float v[8];
            const float4 t0 = tv4[ch * 2], t1 = tv4[ch * 2 + 1];
            const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            if (g.ln_mode) {
              const float4 s0v = sv4[ch * 2], s1v = sv4[ch * 2 + 1];
              const float ss[8] = {s0v.x, s0v.y, s0v.z, s0v.w, s1v.x, s1v.y, s1v.z, s1v.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaf(a_scale, __uint_as_float(acc[ch * 8 + i]), fmaf(a_shift, ss[i], tt[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[ch * 8 + i]) + tt[i];
            }
            uint4* dst = reinterpret_cast<uint4*>(srow + ((ch ^ (row & 7)) << 4));
            if (g.has_res) {
              const uint4 rv = *dst;
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) { v[2 * e] += unpack_lo<T>(rw[e]); v[2 * e + 1] += unpack_hi<T>(rw[e]); }
            }
            uint4 ov;
            ov.x = pack2<T>(v[0], v[1]); ov.y = pack2<T>(v[2], v[3]); ov.z = pack2<T>(v[4], v[5]); ov.w = pack2<T>(v[6], v[7]);
            *dst = ov;
          }
          fence_proxy_async();                                 // generic-proxy smem writes -> visible to the TMA store
          named_bar(2 + grp, 128);
          if (lead_warp && elect_one()) {
            tma_store_3d(&tmO, slots + slot * kPwSlab, ncol0, m0, b);
            bulk_commit();
            bulk_wait_read<1>();                               // this group's previous store has finished reading smem
            if (j >= 1) mbar_arrive(smem_u32(&bar_slot_empty[(uint32_t)grp * 2u + ((j + 1u) & 1u)]));
          }
        }
        // release the accumulator buffer: every thread of the group has finished its TMEM loads (tcgen05.wait::ld above)
        tc_fence_before();
        named_bar(4 + grp, 128);
        if (lead_warp && elect_one()) mbar_arrive(smem_u32(&bar_tempty[buf]));
      }
    }
    if (lead_warp && elect_one()) bulk_wait_all();                               // smem must outlive the last store
  } else {
    // ========================================= residual prefetch ======================================
    if (g.has_res) {
      const int n16 = (g.N + 15) / 16 * 16;
      uint32_t u = 0;
      // A slab can only be requested when its staging slot is free again (two slots per epilogue group), so with ~4 us of loaded
      // DRAM latency per request each group turned one slab around per ~5 K cycles whatever the ring depth.  The tiles this CTA will
      // need g.res_pf iterations from now are pulled into L2 here; the slot loads then see L2 latency.  Measured (cfg2 forward): K4
      // 2.41 -> 2.10 ms and K7 3.08 -> 3.01 ms with one tile ahead; two or more ahead, or prefetching the activation tiles as well,
      // is slower (K7 already runs at ~90 % of the copy bandwidth: extra requests only add DRAM contention).
      auto prefetch_tile = [&](int tp) {
        if (tp >= total_tiles || !elect_one()) return;
        const int b2 = tp / g.m_tiles, m2 = (tp % g.m_tiles) * kPwBlockM;
        for (int c0 = 0; c0 < n16; c0 += 64) tma_prefetch_3d(&tmR, c0, m2, b2);
      };
      for (int k = 0; k < g.res_pf; ++k) prefetch_tile((int)blockIdx.x + k * (int)gridDim.x);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t / g.m_tiles, m0 = (t % g.m_tiles) * kPwBlockM;
        if (g.res_pf) prefetch_tile(t + g.res_pf * (int)gridDim.x);
        for (int ci = 0; ci < g.n_chunks; ++ci) {
          const int c = chunk_of(ci);
          const int ncols = min(g.NC, n16 - c * g.NC);
          for (int s0 = 0; s0 < ncols; s0 += 64, ++u) {
            const uint32_t grp = u & 1u, j = u >> 1;
            const uint32_t slot = grp * 2u + (j & 1u), sph = (j >> 1) & 1u;
            pw_wait_lazy(smem_u32(&bar_slot_empty[slot]), sph ^ 1u);
            if (elect_one()) {
              const uint32_t full = smem_u32(&bar_slot_full[slot]);
              mbar_expect_tx(full, kPwSlab);
              tma_load_3d(slots + slot * kPwSlab, &tmR, full, c * g.NC + s0, m0, b);
            }
            __syncwarp();
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------
static int g_num_sms = 0;

template <class T>
static int launch_pw(const PirGemm* d, cudaStream_t stream) {
  PwArgs g{};
  g.hw = d->H * d->W; g.B = d->B; g.N = d->N; g.K = d->K;
  g.nkb = (d->K + kPwBlockK - 1) / kPwBlockK;
  const int kpad = g.nkb * kPwBlockK;
  const int n16 = (d->N + 15) / 16 * 16;
  if (n16 <= 256) { g.NC = n16; g.n_chunks = 1; } else { g.NC = 256; g.n_chunks = (n16 + 255) / 256; }
  g.n_alloc = (n16 + 63) / 64 * 64;
  g.w_batched = d->w_batched; g.ln_mode = d->ln_mode; g.has_res = d->res ? 1 : 0;
  {
    static const int rot_env = [] { const char* e = getenv("PIR_GEMM_ROTATE"); return e ? atoi(e) : 1; }();
    g.rotate = rot_env;
    static const int pf_env = [] { const char* e = getenv("PIR_GEMM_RESPF"); const int v = e ? atoi(e) : 1; return v < 0 ? 0 : (v > 8 ? 8 : v); }();
    g.res_pf = g.has_res ? pf_env : 0;
  }
  g.m_tiles = (g.hw + kPwBlockM - 1) / kPwBlockM;
  g.ln_s = d->ln_s; g.vec_t = d->vec_t;

  // shared-memory plan: [resident weights] [A (+B) ring] [4 staging slabs] [ln_s | vec_t] [stats mailbox]
  const uint32_t budget = 225 * 1024;
  const uint32_t fixed = 2u * g.n_alloc * 4u + 2048u + (uint32_t)kPwMaxSlots * kPwSlab;
  const uint32_t bres = (uint32_t)g.nkb * g.n_alloc * 128u;
  g.resident = 0;
  if (!d->w_batched && g.nkb <= kPwMaxStages && bres + fixed + (uint32_t)g.nkb * kPwATile <= budget) {
    int st = (int)((budget - bres - fixed) / kPwATile);       // >= nkb: at least one whole tile of A in flight
    if (st > 2 * g.nkb) st = 2 * g.nkb;
    if (st > kPwMaxStages) st = kPwMaxStages;
    g.resident = 1; g.stages = st;
  }
  uint32_t ring_bytes;
  if (g.resident) {
    ring_bytes = g.stages * kPwATile;
    g.off_ring = bres;
  } else {
    const uint32_t sb = kPwATile + (uint32_t)((g.NC + 63) / 64) * 8192u;
    int st = (int)((budget - fixed) / sb);
    if (st > 4) st = 4;
    if (st < 2) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_gemm: N too large for the shared-memory plan");
    g.stages = st;
    ring_bytes = st * sb;
    g.off_ring = 0;
  }
  g.off_slots = g.off_ring + ring_bytes;
  g.off_vec = g.off_slots + (uint32_t)kPwMaxSlots * kPwSlab;
  g.off_stats = g.off_vec + 2u * g.n_alloc * 4u;
  const size_t smem = (size_t)g.off_stats + 2048 + 1024;

  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB, tmO, tmR;
  {
    const uint64_t dims[3] = {(uint64_t)d->K, (uint64_t)g.hw, (uint64_t)d->B};
    const uint64_t strides[2] = {(uint64_t)d->a_pitch * 2, (uint64_t)d->a_bstride * 2};
    const uint32_t box[3] = {(uint32_t)kPwBlockK, (uint32_t)kPwBlockM, 1};
    if (int e = pir_make_tmap(&tmA, dt, 3, d->a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t nb = d->w_batched ? (uint64_t)d->B : 1;
    const uint64_t dims[3] = {(uint64_t)kpad, (uint64_t)d->N, nb};
    const uint64_t strides[2] = {(uint64_t)kpad * 2, (uint64_t)kpad * 2 * (uint64_t)d->N};
    const uint32_t box[3] = {(uint32_t)kPwBlockK, 64, 1};
    if (int e = pir_make_tmap(&tmB, dt, 3, d->w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t dims[3] = {(uint64_t)d->N, (uint64_t)g.hw, (uint64_t)d->B};
    const uint64_t strides[2] = {(uint64_t)d->out_pitch * 2, (uint64_t)d->out_bstride * 2};
    const uint32_t box[3] = {64, 128, 1};
    if (int e = pir_make_tmap(&tmO, dt, 3, d->out, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    tmR = tmO;
    if (d->res) {
      const uint64_t rstrides[2] = {(uint64_t)d->res_pitch * 2, (uint64_t)d->res_bstride * 2};
      if (int e = pir_make_tmap(&tmR, dt, 3, d->res, dims, rstrides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }
  }
  if (!pir_smem_attr_once(reinterpret_cast<const void*>(gemm_pw_kernel<T>), (int)(227 * 1024 - 512), "pir_gemm")) return PIR_ERR_CUDA;
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const int total_tiles = d->B * g.m_tiles;
  const int grid = total_tiles < g_num_sms ? total_tiles : g_num_sms;
  if (pir_launch(gemm_pw_kernel<T>, dim3(grid), dim3(kPwThreads), smem, stream, tmA, tmB, tmO, tmR, g) != cudaSuccess) {}
  return pir_check_launch("pir_gemm(pw)");
}

// entry used by pir_gemm for taps == 1 && out_mode == NHWC16
int pir_gemm_pw(const PirGemm* d, cudaStream_t stream) {
  if ((d->N % 8) || (d->out_pitch % 8) || (d->out_bstride % 8) || ((uintptr_t)d->out & 15))
    return pir_fail(PIR_ERR_ARG, "pir_gemm: NHWC16 output needs N, pitch multiples of 8 and 16-byte alignment");
  if (d->res && ((d->res_pitch % 8) || (d->res_bstride % 8) || ((uintptr_t)d->res & 15)))
    return pir_fail(PIR_ERR_ARG, "pir_gemm: residual not 16-byte aligned");
  return d->dtype == PIR_DTYPE_BF16 ? launch_pw<BF16>(d, stream) : launch_pw<FP16>(d, stream);
}

}  // namespace pir
