// PromptGenBlock front half (net/model.py:226-232) and OverlapPatchEmbed (net/model.py:206) for sm_100a,
// plus the fused tile blend of tiled inference (demo.py:43-47).  All three are small HBM/L2-bound SIMT kernels.
#include "common.cuh"
#include "host.h"

namespace pir {

// ------------------------------------------------------------------------------------------------------
// (1) global average pool, stage 1: per-(image, pixel chunk) channel sums.  block = (C/8) x PL threads.
// ------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
pool_partial_kernel(const unsigned short* __restrict__ x, long long pitch, long long bstride, int HW, int C, int chunk,
                    float* __restrict__ partial) {
  extern __shared__ float sred[];            // [PL][C]
  pdl_launch_dependents();
  pdl_wait();
  const int groups = C >> 3;
  const int PL = blockDim.x / groups;
  const int cg = threadIdx.x % groups;
  const int pl = threadIdx.x / groups;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(p0 + chunk, HW);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (pl < PL) {
    const unsigned short* xb = x + (size_t)b * bstride + cg * 8;
    for (int p = p0 + pl; p < p1; p += PL) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)p * pitch));
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { acc[2 * q] += unpack_lo<T>(w4[q]); acc[2 * q + 1] += unpack_hi<T>(w4[q]); }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sred[pl * C + cg * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < PL; ++r) s += sred[r * C + c];
    partial[((size_t)b * gridDim.x + blockIdx.x) * C + c] = s;
  }
}

// ------------------------------------------------------------------------------------------------------
// (2) linear + softmax over the L prompt components, weighted component sum, bilinear resize
//     (align_corners = False), 16-bit NHWC store.  Each thread: one pixel x 8 channels.
// ------------------------------------------------------------------------------------------------------
constexpr int kMaxL = 8;

template <class T>
__global__ void __launch_bounds__(256)
prompt_mix_kernel(const float* __restrict__ partial, int nchunks, int HW, int C, const float* __restrict__ lin_w,
                  const float* __restrict__ lin_b, int L, const float* __restrict__ prompt, int D, int S, int H, int W,
                  unsigned short* __restrict__ out, long long pitch, long long bstride, float* __restrict__ weights_out, int align) {
  extern __shared__ float semb[];            // [C]
  __shared__ float slog[kMaxL];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  // pooled mean: every block re-derives it from the per-chunk partials (up to 64 of them): spread the chain over 4 threads per channel
  // with two accumulators each -- it is pure L2 latency, and every block of the grid pays it
  {
    float* spart = semb + C;                       // [4][C]
    for (int idx = threadIdx.x; idx < 4 * C; idx += blockDim.x) {
      const int c = idx % C, part = idx / C;
      float s0 = 0.f, s1 = 0.f;
      int k = part;
      for (; k + 4 < nchunks; k += 8) { s0 += partial[((size_t)b * nchunks + k) * C + c]; s1 += partial[((size_t)b * nchunks + k + 4) * C + c]; }
      if (k < nchunks) s0 += partial[((size_t)b * nchunks + k) * C + c];
      spart[part * C + c] = s0 + s1;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) semb[c] = (spart[c] + spart[C + c] + spart[2 * C + c] + spart[3 * C + c]) / (float)HW;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < L) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(semb[c], lin_w[(size_t)warp * C + c], s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) slog[warp] = s + lin_b[warp];
  }
  __syncthreads();
  float wgt[kMaxL];
  {
    float mx = -INFINITY;
    for (int l = 0; l < L; ++l) mx = fmaxf(mx, slog[l]);
    float sum = 0.f;
    for (int l = 0; l < L; ++l) { wgt[l] = expf(slog[l] - mx); sum += wgt[l]; }
    const float inv = 1.0f / sum;
    for (int l = 0; l < L; ++l) wgt[l] *= inv;
  }
  if (weights_out && blockIdx.x == 0 && threadIdx.x < L) weights_out[b * L + threadIdx.x] = wgt[threadIdx.x];

  const int groups = D >> 3;
  const long long total = (long long)H * W * groups;
  // align_corners = False: src = scale * (dst + 0.5) - 0.5, scale = S / H;  True: src = dst * (S - 1) / (H - 1)   (ATen upsample_bilinear2d)
  const float sh = align ? (H > 1 ? (float)(S - 1) / (float)(H - 1) : 0.f) : (float)S / (float)H;
  const float sw = align ? (W > 1 ? (float)(S - 1) / (float)(W - 1) : 0.f) : (float)S / (float)W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int dg = (int)(e % groups);
    const int p = (int)(e / groups);
    const int y = p / W, x = p % W;
    float fy = align ? sh * (float)y : fmaxf(sh * ((float)y + 0.5f) - 0.5f, 0.f);
    float fx = align ? sw * (float)x : fmaxf(sw * ((float)x + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)fy, S - 1), x0 = min((int)fx, S - 1);
    const int y1 = y0 + (y0 < S - 1 ? 1 : 0), x1 = x0 + (x0 < S - 1 ? 1 : 0);
    const float ly = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), lx = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int l = 0; l < L; ++l) {
      const float* P = prompt + ((size_t)l * S * S) * D + dg * 8;
      const float4* p00 = reinterpret_cast<const float4*>(P + ((size_t)y0 * S + x0) * D);
      const float4* p01 = reinterpret_cast<const float4*>(P + ((size_t)y0 * S + x1) * D);
      const float4* p10 = reinterpret_cast<const float4*>(P + ((size_t)y1 * S + x0) * D);
      const float4* p11 = reinterpret_cast<const float4*>(P + ((size_t)y1 * S + x1) * D);
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const float4 a = __ldg(p00 + hlf), bq = __ldg(p01 + hlf), cq = __ldg(p10 + hlf), dq = __ldg(p11 + hlf);
        const float wl = wgt[l];
        acc[hlf * 4 + 0] = fmaf(wl, w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * dq.x, acc[hlf * 4 + 0]);
        acc[hlf * 4 + 1] = fmaf(wl, w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * dq.y, acc[hlf * 4 + 1]);
        acc[hlf * 4 + 2] = fmaf(wl, w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * dq.z, acc[hlf * 4 + 2]);
        acc[hlf * 4 + 3] = fmaf(wl, w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * dq.w, acc[hlf * 4 + 3]);
      }
    }
    uint4 ov;
    ov.x = pack2<T>(acc[0], acc[1]); ov.y = pack2<T>(acc[2], acc[3]);
    ov.z = pack2<T>(acc[4], acc[5]); ov.w = pack2<T>(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + (size_t)b * bstride + (size_t)p * pitch + dg * 8) = ov;
  }
}

// ------------------------------------------------------------------------------------------------------
// PromptGenBlock in ONE launch (net/model.py:226-232): pool -> linear -> softmax -> weighted prompt sum -> bilinear resize.
// The global mean pool needs every pixel of an image before the first output pixel can be written, so the persistent grid (one
// CTA per SM at most: all CTAs are co-resident) meets once at a device-wide barrier:
//   phase 1  per-(image, pixel chunk) partial channel sums -> ws (same layout as the two-kernel path: pir_prompt_bwd reads it);
//   barrier  arrive counter in `sync` (zero on entry; the last CTA to finish the kernel resets it, so CUDA-graph replays work);
//   phase 2  per (image, TS x TS output tile): the component weights of the image (pooled mean, linear, softmax: a few hundred
//            FMAs, redone when the CTA moves to another image), the MIX at source resolution for the source patch the tile needs
//            (sum_l w_l prompt_l, L float4 loads per source vector, kept in shared memory as fp32), then the bilinear taps from
//            shared memory.  The two-kernel path mixed at OUTPUT resolution: 40 L2 loads per output vector instead of ~2.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prompt_grid_barrier(int* counter, int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1);
    while (atomicAdd(counter, 0) < nblocks) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
}

struct PromptFusedArgs {
  const unsigned short* x; long long x_pitch, x_bstride;
  int B, H, W, C, L, D, S, nchunks, chunk;
  float* partial;
  const float* lin_w; const float* lin_b; const float* prompt;
  unsigned short* out; long long out_pitch, out_bstride;
  float* weights_out;
  int align, ts, pr, tiles_x, tiles_y;
  int* sync;
};

constexpr int kPfThreads = 1024;      // one CTA per SM, many loads in flight: both phases stream (pool: x, mix/resize: out)
template <class T>
__global__ void __launch_bounds__(kPfThreads)
prompt_fused_kernel(const PromptFusedArgs g) {
  extern __shared__ float smem_f[];
  __shared__ float slog[kMaxL];
  const int C = g.C, HW = g.H * g.W, D = g.D, S = g.S, L = g.L;
  // ---------------- phase 1: pooled partial sums ----------------
  {
    float* sred = smem_f;                       // [PL][C]
    const int groups = C >> 3;
    const int PL = kPfThreads / groups;
    const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
    for (int work = blockIdx.x; work < g.B * g.nchunks; work += gridDim.x) {
      const int b = work / g.nchunks, ck = work - b * g.nchunks;
      const int p0 = ck * g.chunk, p1 = min(p0 + g.chunk, HW);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      if (pl < PL) {
        const unsigned short* xb = g.x + (size_t)b * g.x_bstride + cg * 8;
        for (int p = p0 + pl; p < p1; p += PL) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)p * g.x_pitch));
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { acc[2 * q] += unpack_lo<T>(w4[q]); acc[2 * q + 1] += unpack_hi<T>(w4[q]); }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) sred[pl * C + cg * 8 + i] = acc[i];
      }
      __syncthreads();
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sm = 0.f;
        for (int r = 0; r < PL; ++r) sm += sred[r * C + c];
        g.partial[((size_t)b * g.nchunks + ck) * C + c] = sm;
      }
      __syncthreads();
    }
  }
  prompt_grid_barrier(g.sync, (int)gridDim.x);

  // ---------------- phase 2: weights, mix at source resolution, bilinear ----------------
  float* semb = smem_f;                         // [C] pooled mean, then [4][C] scratch
  float* smix = smem_f + 5 * C;                 // [pr * pr][D]
  const int per_img = g.tiles_x * g.tiles_y;
  const int n_items = g.B * per_img;
  const int per_cta = (n_items + (int)gridDim.x - 1) / (int)gridDim.x;      // contiguous ranges: a CTA mostly stays inside one image
  const int i0 = blockIdx.x * per_cta, i1 = min(i0 + per_cta, n_items);
  const float sh = g.align ? (g.H > 1 ? (float)(S - 1) / (float)(g.H - 1) : 0.f) : (float)S / (float)g.H;
  const float sw = g.align ? (g.W > 1 ? (float)(S - 1) / (float)(g.W - 1) : 0.f) : (float)S / (float)g.W;
  auto src = [&](float scale, int dst) { return g.align ? scale * (float)dst : fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f); };
  float wgt[kMaxL];
  int cur_b = -1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int item = i0; item < i1; ++item) {
    const int b = item / per_img, t = item - b * per_img;
    const int ty = t / g.tiles_x, tx = t - ty * g.tiles_x;
    if (b != cur_b) {
      cur_b = b;
      float* spart = semb + C;
      for (int idx = threadIdx.x; idx < 4 * C; idx += blockDim.x) {
        const int c = idx % C, part = idx / C;
        float s0 = 0.f, s1 = 0.f;
        int k = part;
        for (; k + 4 < g.nchunks; k += 8) {
          s0 += __ldcg(g.partial + ((size_t)b * g.nchunks + k) * C + c);
          s1 += __ldcg(g.partial + ((size_t)b * g.nchunks + k + 4) * C + c);
        }
        if (k < g.nchunks) s0 += __ldcg(g.partial + ((size_t)b * g.nchunks + k) * C + c);
        spart[part * C + c] = s0 + s1;
      }
      __syncthreads();
      for (int c = threadIdx.x; c < C; c += blockDim.x) semb[c] = (spart[c] + spart[C + c] + spart[2 * C + c] + spart[3 * C + c]) / (float)HW;
      __syncthreads();
      if (warp < L) {
        float sm = 0.f;
        for (int c = lane; c < C; c += 32) sm = fmaf(semb[c], g.lin_w[(size_t)warp * C + c], sm);
#pragma unroll
        for (int o = 16; o; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
        if (lane == 0) slog[warp] = sm + g.lin_b[warp];
      }
      __syncthreads();
      float mx = -INFINITY;
      for (int l = 0; l < L; ++l) mx = fmaxf(mx, slog[l]);
      float sum = 0.f;
      for (int l = 0; l < L; ++l) { wgt[l] = expf(slog[l] - mx); sum += wgt[l]; }
      const float inv = 1.0f / sum;
      for (int l = 0; l < L; ++l) wgt[l] *= inv;
      if (g.weights_out && t == 0 && threadIdx.x < L) g.weights_out[b * L + threadIdx.x] = wgt[threadIdx.x];
      __syncthreads();                           // slog is rewritten for the next image
    }
    // source patch of this tile
    const int oy0 = ty * g.ts, ox0 = tx * g.ts;
    const int oy1 = min(oy0 + g.ts, g.H) - 1, ox1 = min(ox0 + g.ts, g.W) - 1;
    const int py0 = min((int)src(sh, oy0), S - 1), px0 = min((int)src(sw, ox0), S - 1);
    const int py1 = min(min((int)src(sh, oy1), S - 1) + 1, S - 1), px1 = min(min((int)src(sw, ox1), S - 1) + 1, S - 1);
    const int ph = py1 - py0 + 1, pw = px1 - px0 + 1;
    const int d4 = D >> 2;
    for (int e = threadIdx.x; e < ph * pw * d4; e += blockDim.x) {
      const int dq = e % d4, pp = e / d4;
      const int yy = pp / pw, xx = pp - yy * pw;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int l = 0; l < L; ++l) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g.prompt + (((size_t)l * S + (py0 + yy)) * S + (px0 + xx)) * D) + dq);
        acc.x = fmaf(wgt[l], v.x, acc.x); acc.y = fmaf(wgt[l], v.y, acc.y); acc.z = fmaf(wgt[l], v.z, acc.z); acc.w = fmaf(wgt[l], v.w, acc.w);
      }
      reinterpret_cast<float4*>(smix + (size_t)pp * D)[dq] = acc;
    }
    __syncthreads();
    const int groups = D >> 3;
    const int tw = ox1 - ox0 + 1, th = oy1 - oy0 + 1;
    for (int e = threadIdx.x; e < th * tw * groups; e += blockDim.x) {
      const int dg = e % groups, pp = e / groups;
      const int y = oy0 + pp / tw, x = ox0 + pp % tw;
      const float fy = src(sh, y), fx = src(sw, x);
      const int y0 = min((int)fy, S - 1), x0 = min((int)fx, S - 1);
      const int y1 = y0 + (y0 < S - 1 ? 1 : 0), x1 = x0 + (x0 < S - 1 ? 1 : 0);
      const float ly = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), lx = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      const float* q00 = smix + ((size_t)(y0 - py0) * pw + (x0 - px0)) * D + dg * 8;
      const float* q01 = smix + ((size_t)(y0 - py0) * pw + (x1 - px0)) * D + dg * 8;
      const float* q10 = smix + ((size_t)(y1 - py0) * pw + (x0 - px0)) * D + dg * 8;
      const float* q11 = smix + ((size_t)(y1 - py0) * pw + (x1 - px0)) * D + dg * 8;
      float acc[8];
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const float4 a = reinterpret_cast<const float4*>(q00)[hlf], bq = reinterpret_cast<const float4*>(q01)[hlf];
        const float4 cq = reinterpret_cast<const float4*>(q10)[hlf], dq = reinterpret_cast<const float4*>(q11)[hlf];
        // same association as the two-kernel path for each component would differ in rounding only; here the mix comes first
        acc[hlf * 4 + 0] = w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * dq.x;
        acc[hlf * 4 + 1] = w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * dq.y;
        acc[hlf * 4 + 2] = w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * dq.z;
        acc[hlf * 4 + 3] = w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * dq.w;
      }
      uint4 ov;
      ov.x = pack2<T>(acc[0], acc[1]); ov.y = pack2<T>(acc[2], acc[3]);
      ov.z = pack2<T>(acc[4], acc[5]); ov.w = pack2<T>(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(g.out + (size_t)b * g.out_bstride + ((size_t)y * g.W + x) * g.out_pitch + dg * 8) = ov;
    }
    __syncthreads();
  }
  // the last CTA to get here re-arms the barrier for the next launch (graph replay)
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(g.sync + 1, 1) == (int)gridDim.x - 1) { g.sync[0] = 0; g.sync[1] = 0; __threadfence(); }
  }
}

static int pool_chunks(int HW) {
  int n = (HW + 255) / 256;
  return n > 64 ? 64 : (n < 1 ? 1 : n);
}

template <class T>
static int launch_prompt(const PirPrompt* d, cudaStream_t s) {
  const int HW = d->H * d->W;
  const int nchunks = pool_chunks(HW);
  const int chunk = (HW + nchunks - 1) / nchunks;
  const int groups = d->C / 8;
  if (groups > 256) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_prompt_gen: C > 2048");
  const int PL = 256 / groups;
  const int threads = PL * groups;
  static const bool two = [] { const char* e = getenv("PIR_PROMPT_FUSED"); return e && e[0] == '0'; }();   // A/B: the two-kernel path
  if (d->sync && !two && (d->D % 8) == 0) {
    // one launch: output tile size from the shared-memory budget of the source patch (fp32 mix of (ceil(ts * scale) + 2)^2 x D)
    const float sc_h = d->align_corners ? (d->H > 1 ? (float)(d->S - 1) / (float)(d->H - 1) : 0.f) : (float)d->S / (float)d->H;
    const float sc_w = d->align_corners ? (d->W > 1 ? (float)(d->S - 1) / (float)(d->W - 1) : 0.f) : (float)d->S / (float)d->W;
    const float sc = sc_h > sc_w ? sc_h : sc_w;
    int ts = 16, pr = 0;
    size_t smem = 0;
    for (; ts >= 2; ts >>= 1) {
      pr = (int)((float)(ts - 1) * sc) + 3;
      if (pr > d->S) pr = d->S;
      smem = ((size_t)5 * d->C + (size_t)pr * pr * d->D) * sizeof(float);
      const size_t pool = (size_t)(kPfThreads / groups) * d->C * sizeof(float);
      if (pool > smem) smem = pool;
      if (smem <= 160 * 1024) break;
    }
    if (ts >= 2) {
      if (!pir_smem_attr_once(reinterpret_cast<const void*>(prompt_fused_kernel<T>), (int)(160 * 1024), "pir_prompt_gen")) return PIR_ERR_CUDA;
      static int num_sms = 0;
      if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
      }
      PromptFusedArgs g{};
      g.x = reinterpret_cast<const unsigned short*>(d->x); g.x_pitch = d->x_pitch; g.x_bstride = d->x_bstride;
      g.B = d->B; g.H = d->H; g.W = d->W; g.C = d->C; g.L = d->L; g.D = d->D; g.S = d->S; g.nchunks = nchunks; g.chunk = chunk;
      g.partial = d->ws; g.lin_w = d->lin_w; g.lin_b = d->lin_b; g.prompt = d->prompt;
      g.out = reinterpret_cast<unsigned short*>(d->out); g.out_pitch = d->out_pitch; g.out_bstride = d->out_bstride;
      g.weights_out = d->weights_out; g.align = d->align_corners; g.ts = ts; g.pr = pr;
      g.tiles_x = (d->W + ts - 1) / ts; g.tiles_y = (d->H + ts - 1) / ts;
      g.sync = d->sync;
      const int items = d->B * g.tiles_x * g.tiles_y, pool_items = d->B * nchunks;
      int grid = items > pool_items ? items : pool_items;
      if (grid > num_sms) grid = num_sms;            // every CTA must be resident: they meet at a device-wide barrier
      prompt_fused_kernel<T><<<dim3(grid), dim3(kPfThreads), smem, s>>>(g);
      return pir_check_launch("pir_prompt_gen(fused)");
    }
  }
  pir_launch(pool_partial_kernel<T>, dim3(nchunks, d->B), dim3(threads), (size_t)PL * d->C * sizeof(float), s,
             reinterpret_cast<const unsigned short*>(d->x), d->x_pitch, d->x_bstride, HW, d->C, chunk, d->ws);
  if (int e = pir_check_launch("pir_prompt_gen(pool)")) return e;
  const long long total = (long long)HW * (d->D / 8);
  // every block re-derives the component weights from the pooled partials, so keep the grid to a few blocks per SM in total
  int blocks = (int)((total + 255) / 256);
  const int cap = (148 * 4 + d->B - 1) / d->B;
  if (blocks > cap) blocks = cap;
  pir_launch(prompt_mix_kernel<T>, dim3(blocks, d->B), dim3(256), 5 * d->C * sizeof(float), s,
      d->ws, nchunks, HW, d->C, d->lin_w, d->lin_b, d->L, d->prompt, d->D, d->S, d->H, d->W,
      reinterpret_cast<unsigned short*>(d->out), d->out_pitch, d->out_bstride, d->weights_out, d->align_corners);
  return pir_check_launch("pir_prompt_gen(mix)");
}

// ------------------------------------------------------------------------------------------------------
// OverlapPatchEmbed: dense 3x3 (pad 1) from the fp32 NCHW image to 16-bit NHWC.  K = 9*Cin = 27: SIMT.
// block = (Cout/8) x 32 pixels.
// ------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
patch_embed_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias, int H, int W,
                   int Cin, int Cout, unsigned short* __restrict__ out, long long pitch, long long bstride) {
  extern __shared__ float sw[];              // [Cin*9][Cout]
  const int groups = Cout >> 3;
  for (int e = threadIdx.x; e < Cin * 9 * Cout; e += blockDim.x) {
    const int co = e % Cout, k = e / Cout;   // k = ci*9 + tap
    sw[e] = w[(size_t)co * Cin * 9 + k];
  }
  __syncthreads();
  const int g = threadIdx.x % groups;
  const int pl = threadIdx.x / groups;
  const int ppb = blockDim.x / groups;
  const int b = blockIdx.y;
  // grid-stride over pixel groups: the 3x3xCinxCout weights are staged once per (persistent) block, not once per 42 pixels
  for (long long p = (long long)blockIdx.x * ppb + pl; p < (long long)H * W; p += (long long)gridDim.x * ppb) {
  const int y = (int)(p / W), x = (int)(p % W);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = bias ? bias[g * 8 + i] : 0.f;
  const float* ib = img + (size_t)b * Cin * H * W;
  for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(ib + ((size_t)ci * H + yy) * W + xx) : 0.f;
      const float* ws = sw + (size_t)(ci * 9 + t) * Cout + g * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(v, ws[i], acc[i]);
    }
  }
  uint4 ov;
  ov.x = pack2<T>(acc[0], acc[1]); ov.y = pack2<T>(acc[2], acc[3]);
  ov.z = pack2<T>(acc[4], acc[5]); ov.w = pack2<T>(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(out + (size_t)b * bstride + (size_t)p * pitch + g * 8) = ov;
  }
}


// Fast path for the shapes the networks use (Cin = 3, Cout = 48): one thread = one pixel x ALL output channels.  The 27 inputs of the
// pixel are loaded once (coalesced across the warp: 32 consecutive pixels per plane row), every weight is a shared-memory broadcast
// (the whole warp reads the same address), 27 x 48 FMAs per pixel in registers, six 16-byte stores.  ~13x fewer instructions per
// output than the generic kernel above, which re-loads the inputs once per 8-channel group.
template <class T, int CIN, int COUT>
__global__ void __launch_bounds__(128)
patch_embed_px_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias, int H, int W,
                      unsigned short* __restrict__ out, long long pitch, long long bstride) {
  __shared__ __align__(16) float sw[CIN * 9][COUT];
  for (int e = threadIdx.x; e < CIN * 9 * COUT; e += blockDim.x) {
    const int co = e % COUT, k = e / COUT;
    sw[k][co] = w[(size_t)co * CIN * 9 + k];
  }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();                                   // (the output buffer may still be read by the previous forward's tail)
  const int b = blockIdx.y;
  const float* ib = img + (size_t)b * CIN * H * W;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < (long long)H * W; p += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(p / W), x = (int)(p % W);
    float v[CIN * 9];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        v[ci * 9 + t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(ib + ((size_t)ci * H + yy) * W + xx) : 0.f;
      }
    unsigned short* op = out + (size_t)b * bstride + (size_t)p * pitch;
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 16) {                  // 16 output channels at a time keeps the accumulators in registers
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = bias ? __ldg(bias + c0 + i) : 0.f;
#pragma unroll
      for (int k = 0; k < CIN * 9; ++k) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = *reinterpret_cast<const float4*>(&sw[k][c0 + 4 * q]);
          acc[4 * q] = fmaf(v[k], w4.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v[k], w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v[k], w4.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v[k], w4.w, acc[4 * q + 3]);
        }
      }
      uint4 o0, o1;
      o0.x = pack2<T>(acc[0], acc[1]); o0.y = pack2<T>(acc[2], acc[3]); o0.z = pack2<T>(acc[4], acc[5]); o0.w = pack2<T>(acc[6], acc[7]);
      o1.x = pack2<T>(acc[8], acc[9]); o1.y = pack2<T>(acc[10], acc[11]); o1.z = pack2<T>(acc[12], acc[13]); o1.w = pack2<T>(acc[14], acc[15]);
      *reinterpret_cast<uint4*>(op + c0) = o0;
      *reinterpret_cast<uint4*>(op + c0 + 8) = o1;
    }
  }
}

template <class T>
static int launch_patch_embed(const PirPatchEmbed* d, cudaStream_t s) {
  if (d->Cin == 3 && d->Cout == 48) {
    const long long hw = (long long)d->H * d->W;
    long long gx = (hw + 127) / 128;
    const long long cap = (148 * 12 + d->B - 1) / d->B;
    if (gx > cap) gx = cap;
    pir_launch(patch_embed_px_kernel<T, 3, 48>, dim3((unsigned)gx, d->B), dim3(128), 0, s, d->img, d->w, d->bias, d->H, d->W,
               reinterpret_cast<unsigned short*>(d->out), d->out_pitch, d->out_bstride);
    return pir_check_launch("pir_patch_embed");
  }
  const int groups = d->Cout / 8;
  const int ppb = 256 / groups;
  const int threads = ppb * groups;
  const long long HW = (long long)d->H * d->W;
  const size_t smem = (size_t)d->Cin * 9 * d->Cout * sizeof(float);
  if (smem > 48 * 1024) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_patch_embed: weight does not fit 48 KB of shared memory");
  long long gx = (HW + ppb - 1) / ppb;
  const long long cap = (148 * 8 + d->B - 1) / d->B;                 // ~8 resident blocks per SM over the whole batch
  if (gx > cap) gx = cap;
  patch_embed_kernel<T><<<dim3((unsigned)gx, d->B), threads, smem, s>>>(
      d->img, d->w, d->bias, d->H, d->W, d->Cin, d->Cout, reinterpret_cast<unsigned short*>(d->out), d->out_pitch, d->out_bstride);
  return pir_check_launch("pir_patch_embed");
}

// ------------------------------------------------------------------------------------------------------
// tile blend (demo.py:36-47): out = clamp(sum of covering tiles / hit count, 0, 1), tiles summed in the
// reference's loop order (rows outer, columns inner) -> deterministic, no atomics.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_blend_kernel(const float* __restrict__ tiles, int ny, int nx, const int* __restrict__ ys, const int* __restrict__ xs,
                  int C, int th, int tw, float* __restrict__ out, int H, int W) {
  const long long n = (long long)C * H * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % W);
    const int y = (int)((e / W) % H);
    const int c = (int)(e / ((long long)W * H));
    float acc = 0.f, cnt = 0.f;
    for (int iy = 0; iy < ny; ++iy) {
      const int ly = y - ys[iy];
      if (ly < 0 || ly >= th) continue;
      for (int ix = 0; ix < nx; ++ix) {
        const int lx = x - xs[ix];
        if (lx < 0 || lx >= tw) continue;
        acc += tiles[(((size_t)(iy * nx + ix) * C + c) * th + ly) * tw + lx];
        cnt += 1.f;
      }
    }
    out[e] = fminf(fmaxf(acc / cnt, 0.f), 1.f);
  }
}

}  // namespace pir

extern "C" int64_t pir_prompt_ws_floats(int32_t B, int32_t HW, int32_t C) {
  return (int64_t)B * pir::pool_chunks(HW) * C;
}

extern "C" int pir_prompt_gen(const PirPrompt* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_prompt_gen: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0) return pir_fail(PIR_ERR_ARG, "pir_prompt_gen: empty problem");
  if ((d->C % 8) || (d->D % 8) || (d->x_pitch % 8) || (d->out_pitch % 8) || ((uintptr_t)d->x & 15) || ((uintptr_t)d->out & 15) ||
      ((uintptr_t)d->prompt & 15))
    return pir_fail(PIR_ERR_ARG, "pir_prompt_gen: channels / pitches / pointers are not vector aligned");
  if (d->L < 1 || d->L > pir::kMaxL) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_prompt_gen: 1 <= L <= 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return d->dtype == PIR_DTYPE_BF16 ? pir::launch_prompt<pir::BF16>(d, s) : pir::launch_prompt<pir::FP16>(d, s);
}

extern "C" int pir_prompt_gen_kernels(const PirPrompt* d) {
  static const bool two = [] { const char* e = getenv("PIR_PROMPT_FUSED"); return e && e[0] == '0'; }();
  return (d && d->sync && !two && (d->D % 8) == 0) ? 1 : 2;
}

extern "C" int pir_patch_embed(const PirPatchEmbed* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_patch_embed: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0) return pir_fail(PIR_ERR_ARG, "pir_patch_embed: empty problem");
  if ((d->Cout % 8) || d->Cout > 2048 || (d->out_pitch % 8) || ((uintptr_t)d->out & 15)) return pir_fail(PIR_ERR_ARG, "pir_patch_embed: Cout / pitch not vector aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return d->dtype == PIR_DTYPE_BF16 ? pir::launch_patch_embed<pir::BF16>(d, s) : pir::launch_patch_embed<pir::FP16>(d, s);
}

extern "C" int pir_tile_blend(const float* tiles, int32_t ny, int32_t nx, const int32_t* ys, const int32_t* xs, int32_t C,
                              int32_t th, int32_t tw, float* out, int32_t H, int32_t W, void* stream) {
  if (!tiles || !ys || !xs || !out || ny <= 0 || nx <= 0 || C <= 0) return pir_fail(PIR_ERR_ARG, "pir_tile_blend: bad arguments");
  const long long n = (long long)C * H * W;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pir::tile_blend_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tiles, ny, nx, ys, xs, C, th, tw, out, H, W);
  return pir_check_launch("pir_tile_blend");
}
