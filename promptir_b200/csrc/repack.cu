// pir_repack: every derived weight cache of an engine rebuilt from the live fp32 parameters in ONE launch.
//
// The fp32 nn.Parameters stay the canonical storage (optimizers, state_dict); the kernels consume 16-bit K-major copies with the
// LayerNorm affine folded in (net/model.py:60-63 into :88,:111), zero padding to the TMA/UMMA granules, the GDFN [x1 | x2] padded
// channel space, tap-major 3x3 layouts and -- for the backward -- transposed / tap-flipped variants.  promptir_b200/packing.py
// states those layouts in torch (it is what the CPU wiring tests run); this file is the device implementation: a table of
// PirPackJob records (built once per engine, resident in HBM) and one kernel whose CTAs each produce one destination row of one
// job.  A training step therefore refreshes all caches with a single ~0.1 ms launch after optimizer.step() (train.py:52-56)
// instead of ~10^3 small framework kernels.
//
//   POINTWISE  dst[d][c] = round16( W(r, s) * gamma ),  r = inv_row(d), s = inv_col(c); zero where unmapped
//              ln_s[d]   = sum_c float(dst[d][c])                (row sum of the ROUNDED weights: the LayerNorm mean term then cancels exactly)
//              vec_t[d]  = sum_s W(r, s) * beta[s] + bias[r]
//   CONV3X3    dst[n][tap * kpad + cin] = round16( w[n][cin][tap] )      (transpose: w[cin][n][.], flip: tap -> 8 - tap)
//   DEPTHWISE  dst[tap][d] = round16( w[inv(d)][tap] )                   (flip: tap -> 8 - tap)
//   VEC        dst[d] = v[inv(d)]                                         (fp32)
//   PROMPT     dst[l][s][t][dch] = p[l][dch][s][t]                        (fp32, channels last)
// inv(): identity below `split`; [split, hp) is padding; hp + j maps to split + j  (packing.gdfn_maps).
#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kRpThreads = 128;

__device__ __forceinline__ int rp_inv(int d, int n_src, int split, int hp) {
  // destination index -> source index or -1 (padding)
  if (split <= 0) return d < n_src ? d : -1;
  if (d < split) return d;
  if (d < hp) return -1;
  const int s = d - hp + split;
  return s < n_src ? s : -1;
}

__device__ __forceinline__ unsigned short rp_round(float v, int dtype) {
  return dtype == PIR_DTYPE_BF16 ? to16<BF16>(v) : to16<FP16>(v);
}
__device__ __forceinline__ float rp_widen(unsigned short h, int dtype) {
  return dtype == PIR_DTYPE_BF16 ? from16<BF16>(h) : from16<FP16>(h);
}

__device__ __forceinline__ float rp_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kRpThreads / 32; ++i) s += red[i];
  return s;
}

__global__ void __launch_bounds__(kRpThreads)
repack_kernel(const PirPackJob* __restrict__ jobs, const int32_t* __restrict__ first_row, int n_jobs) {
  __shared__ float red[kRpThreads / 32];
  // job of this CTA: last j with first_row[j] <= blockIdx.x   (first_row has n_jobs + 1 entries)
  int lo = 0, hi = n_jobs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (first_row[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const PirPackJob j = jobs[lo];
  const int d = (int)blockIdx.x - first_row[lo];
  const int tid = threadIdx.x;

  if (j.kind == PIR_PACK_POINTWISE) {
    const int r = rp_inv(d, j.n, j.row_split, j.row_hp);
    unsigned short* dst = reinterpret_cast<unsigned short*>(j.dst) + (size_t)d * j.k_pad;
    float s_sum = 0.f, t_sum = 0.f;
    for (int c = tid; c < j.k_pad; c += kRpThreads) {
      const int s = r >= 0 ? rp_inv(c, j.k, j.col_split, j.col_hp) : -1;
      float v = 0.f;
      if (s >= 0) {
        const float w = j.transpose ? j.src[(size_t)s * j.n + r] : j.src[(size_t)r * j.k + s];
        v = j.gamma_axis == 1 ? w * j.gamma[s] : (j.gamma_axis == 2 ? w * j.gamma[r] : w);
        if (j.beta) t_sum = fmaf(w, j.beta[s], t_sum);
      }
      const unsigned short h = rp_round(v, j.dst_dtype);
      dst[c] = h;
      s_sum += rp_widen(h, j.dst_dtype);
    }
    if (j.ln_s) { const float s = rp_block_sum(s_sum, red); if (tid == 0) j.ln_s[d] = s; }
    if (j.vec_t) {
      const float t = j.beta ? rp_block_sum(t_sum, red) : 0.f;
      if (tid == 0) j.vec_t[d] = r >= 0 ? t + (j.bias ? j.bias[r] : 0.f) : 0.f;
    }
  } else if (j.kind == PIR_PACK_CONV3X3) {
    // one destination row = output channel d: 9 taps x k_pad input channels
    unsigned short* dst = reinterpret_cast<unsigned short*>(j.dst) + (size_t)d * 9 * j.k_pad;
    for (int i = tid; i < 9 * j.k_pad; i += kRpThreads) {
      const int tap = i / j.k_pad, cin = i - tap * j.k_pad;
      float v = 0.f;
      if (cin < j.k) {
        const int st = j.flip ? 8 - tap : tap;
        v = j.transpose ? j.src[((size_t)cin * j.n + d) * 9 + st] : j.src[((size_t)d * j.k + cin) * 9 + st];
      }
      dst[i] = rp_round(v, j.dst_dtype);
    }
  } else if (j.kind == PIR_PACK_DEPTHWISE) {
    // one destination row = tap d: n_total channels
    unsigned short* dst = reinterpret_cast<unsigned short*>(j.dst) + (size_t)d * j.n_total;
    const int st = j.flip ? 8 - d : d;
    for (int c = tid; c < j.n_total; c += kRpThreads) {
      const int s = rp_inv(c, j.n, j.row_split, j.row_hp);
      dst[c] = rp_round(s >= 0 ? j.src[(size_t)s * 9 + st] : 0.f, j.dst_dtype);
    }
  } else if (j.kind == PIR_PACK_VEC) {
    // rows of kRpThreads elements
    const int c = d * kRpThreads + tid;
    if (c < j.n_total) {
      const int s = rp_inv(c, j.n, j.row_split, j.row_hp);
      reinterpret_cast<float*>(j.dst)[c] = s >= 0 ? j.src[s] : 0.f;
    }
  } else if (j.kind == PIR_PACK_PROMPT) {
    // source [L = n][D = k][S*S = n_total]; one destination row = (l, s, t): D channels
    const int ss = j.n_total;
    const int l = d / ss, st = d - l * ss;
    float* dst = reinterpret_cast<float*>(j.dst) + (size_t)d * j.k;
    for (int c = tid; c < j.k; c += kRpThreads) dst[c] = j.src[((size_t)l * j.k + c) * ss + st];
  }
}

}  // namespace pir

extern "C" int64_t pir_repack_rows(const PirPackJob* job) {
  if (!job) return -1;
  switch (job->kind) {
    case PIR_PACK_POINTWISE: return job->n_total;
    case PIR_PACK_CONV3X3: return job->n;
    case PIR_PACK_DEPTHWISE: return 9;
    case PIR_PACK_VEC: return (job->n_total + pir::kRpThreads - 1) / pir::kRpThreads;
    case PIR_PACK_PROMPT: return (int64_t)job->n * job->n_total;
    default: return -1;
  }
}

extern "C" int pir_repack(const PirPackJob* jobs_dev, const int32_t* first_row_dev, int32_t n_jobs, int32_t n_rows, void* stream) {
  if (!jobs_dev || !first_row_dev) return pir_fail(PIR_ERR_ARG, "pir_repack: null job table");
  if (n_jobs <= 0 || n_rows <= 0) return pir_fail(PIR_ERR_ARG, "pir_repack: empty job table");
  pir::repack_kernel<<<dim3((unsigned)n_rows), dim3(pir::kRpThreads), 0, reinterpret_cast<cudaStream_t>(stream)>>>(jobs_dev, first_row_dev, n_jobs);
  return pir_check_launch("pir_repack");
}
