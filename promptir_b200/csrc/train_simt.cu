// Training-side SIMT kernels for sm_100a: everything of the backward pass that is NOT a tensor-core contraction.
// Reference: autograd of net/model.py (LayerNorm :39-41,60-63; GDFN gate :96-97; MDTA :123-137; PixelShuffle/Unshuffle
// :165,175; PromptGenBlock :226-232) as driven by train.py:37-46.  All of them are HBM/L2-bound: 16-byte vector loads and
// stores on NHWC 16-bit tensors, fp32 arithmetic, deterministic reductions (no atomics).
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kMaxLTrain = 8;

template <class T> __device__ __forceinline__ void unpack8t(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) { f[2 * q] = unpack_lo<T>(w[q]); f[2 * q + 1] = unpack_hi<T>(w[q]); }
}
template <class T> __device__ __forceinline__ uint4 pack8t(const float (&f)[8]) {
  uint4 v;
  v.x = pack2<T>(f[0], f[1]); v.y = pack2<T>(f[2], f[3]); v.z = pack2<T>(f[4], f[5]); v.w = pack2<T>(f[6], f[7]);
  return v;
}

// ------------------------------------------------------------------------------------------------------
// LayerNorm forward / backward.  LPP lanes cooperate on one pixel, each owning up to 3 chunks of 8 channels.
// ------------------------------------------------------------------------------------------------------
struct LnArgs {
  int npix_img, B, C, nch, biasfree;       // nch = chunks per lane
  const unsigned short* x; long long x_pitch, x_bstride;
  unsigned short* xhat; long long xh_pitch, xh_bstride;
  float* rstd;
  unsigned short* g; long long g_pitch, g_bstride;
};

template <int LPP> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPP / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <class T, int LPP>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const LnArgs a) {
  const int sub = threadIdx.x % LPP;
  const int ppb = blockDim.x / LPP;
  const long long total = (long long)a.B * a.npix_img;
  const long long gstride = (long long)gridDim.x * ppb;
  const int chunks = a.C >> 3;
  const float invC = 1.0f / (float)a.C;
  const long long rounds = (total + gstride - 1) / gstride;
  for (long long it = 0; it < rounds; ++it) {
    // all lanes of a warp iterate together (shuffles); out-of-range pixels are masked
    const long long p = it * gstride + (long long)blockIdx.x * ppb + threadIdx.x / LPP;
    const bool live = p < total;
    const int b = live ? (int)(p / a.npix_img) : 0;
    const int pi = live ? (int)(p % a.npix_img) : 0;
    const unsigned short* xp = a.x + (size_t)b * a.x_bstride + (size_t)pi * a.x_pitch;
    float v[3][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int ch = sub + i * LPP;
      const bool ok = live && i < a.nch && ch < chunks;
      if (ok) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(xp + ch * 8));
        unpack8t<T>(u, v[i]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
    const float mu = group_sum<LPP>(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int ch = sub + i * LPP;
      if (i < a.nch && ch < chunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mu; q = fmaf(d, d, q); }
      }
    }
    const float rstd = rsqrtf(group_sum<LPP>(q) * invC + 1e-5f);
    if (!live) continue;
    const float sh = a.biasfree ? 0.f : mu;
    unsigned short* op = a.xhat + (size_t)b * a.xh_bstride + (size_t)pi * a.xh_pitch;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int ch = sub + i * LPP;
      if (i < a.nch && ch < chunks) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - sh) * rstd;
        *reinterpret_cast<uint4*>(op + ch * 8) = pack8t<T>(o);
      }
    }
    if (sub == 0) a.rstd[p] = rstd;
  }
}

template <class T, int LPP>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const LnArgs a) {
  const int sub = threadIdx.x % LPP;
  const int ppb = blockDim.x / LPP;
  const long long total = (long long)a.B * a.npix_img;
  const long long gstride = (long long)gridDim.x * ppb;
  const int chunks = a.C >> 3;
  const float invC = 1.0f / (float)a.C;
  const long long rounds = (total + gstride - 1) / gstride;
  for (long long it = 0; it < rounds; ++it) {
    const long long p = it * gstride + (long long)blockIdx.x * ppb + threadIdx.x / LPP;
    const bool live = p < total;
    const int b = live ? (int)(p / a.npix_img) : 0;
    const int pi = live ? (int)(p % a.npix_img) : 0;
    const unsigned short* dp = a.x + (size_t)b * a.x_bstride + (size_t)pi * a.x_pitch;
    const unsigned short* hp = a.xhat + (size_t)b * a.xh_bstride + (size_t)pi * a.xh_pitch;
    float d[3][8], h[3][8];
    float sd = 0.f, sdh = 0.f, sh = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int ch = sub + i * LPP;
      if (live && i < a.nch && ch < chunks) {
        unpack8t<T>(__ldg(reinterpret_cast<const uint4*>(dp + ch * 8)), d[i]);
        unpack8t<T>(__ldg(reinterpret_cast<const uint4*>(hp + ch * 8)), h[i]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) { d[i][e] = 0.f; h[i][e] = 0.f; }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { sd += d[i][e]; sdh = fmaf(d[i][e], h[i][e], sdh); sh += h[i][e]; }
    }
    const float md = group_sum<LPP>(sd) * invC, mdh = group_sum<LPP>(sdh) * invC, mh = group_sum<LPP>(sh) * invC;
    if (!live) continue;
    const float rstd = a.rstd[p];
    unsigned short* gp = a.g + (size_t)b * a.g_bstride + (size_t)pi * a.g_pitch;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int ch = sub + i * LPP;
      if (i < a.nch && ch < chunks) {
        float gv[8];
        unpack8t<T>(*reinterpret_cast<const uint4*>(gp + ch * 8), gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dx = a.biasfree ? (d[i][e] - (h[i][e] - mh) * mdh) : (d[i][e] - md - h[i][e] * mdh);
          gv[e] = fmaf(rstd, dx, gv[e]);
        }
        *reinterpret_cast<uint4*>(gp + ch * 8) = pack8t<T>(gv);
      }
    }
  }
}

static void ln_plan(int C, int* lpp, int* nch) {
  // lanes per pixel in {8, 16, 32}, chunks per lane <= 3: least wasted lane-chunks
  const int chunks = C / 8;
  int best = 1 << 30;
  *lpp = 32; *nch = 3;
  for (int l = 8; l <= 32; l <<= 1) {
    const int n = (chunks + l - 1) / l;
    if (n > 3) continue;
    if (l * n - chunks < best) { best = l * n - chunks; *lpp = l; *nch = n; }
  }
}

template <class T>
static int launch_ln(const PirLn* d, bool bwd, cudaStream_t s) {
  LnArgs a{};
  a.npix_img = d->H * d->W; a.B = d->B; a.C = d->C; a.biasfree = d->ln_mode == PIR_LN_BIASFREE;
  a.x = reinterpret_cast<const unsigned short*>(d->x); a.x_pitch = d->x_pitch; a.x_bstride = d->x_bstride;
  a.xhat = reinterpret_cast<unsigned short*>(d->xhat); a.xh_pitch = d->xh_pitch; a.xh_bstride = d->xh_bstride;
  a.rstd = d->rstd;
  a.g = reinterpret_cast<unsigned short*>(d->g); a.g_pitch = d->g_pitch; a.g_bstride = d->g_bstride;
  int lpp;
  ln_plan(d->C, &lpp, &a.nch);
  const long long total = (long long)d->B * a.npix_img;
  const int ppb = 256 / lpp;
  long long blocks = (total + ppb - 1) / ppb;
  if (blocks > 148 * 8) blocks = 148 * 8;
#define PIR_LN_LAUNCH(L)                                                          \
  if (bwd) ln_bwd_kernel<T, L><<<(unsigned)blocks, 256, 0, s>>>(a);               \
  else ln_fwd_kernel<T, L><<<(unsigned)blocks, 256, 0, s>>>(a)
  if (lpp == 8) { PIR_LN_LAUNCH(8); } else if (lpp == 16) { PIR_LN_LAUNCH(16); } else { PIR_LN_LAUNCH(32); }
#undef PIR_LN_LAUNCH
  return pir_check_launch(bwd ? "pir_ln_bwd" : "pir_ln_fwd");
}

static int check_ln(const PirLn* d, bool bwd, const char* who) {
  if (!d) return pir_fail(PIR_ERR_ARG, "%s: null descriptor", who);
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0) return pir_fail(PIR_ERR_ARG, "%s: empty problem", who);
  if (d->ln_mode != PIR_LN_WITHBIAS && d->ln_mode != PIR_LN_BIASFREE) return pir_fail(PIR_ERR_ARG, "%s: ln_mode", who);
  if ((d->C % 8) || d->C > 768) return pir_fail(PIR_ERR_UNSUPPORTED, "%s: C must be a multiple of 8 and <= 768", who);
  if ((d->x_pitch % 8) || (d->xh_pitch % 8) || (d->x_bstride % 8) || (d->xh_bstride % 8) || ((uintptr_t)d->x & 15) || ((uintptr_t)d->xhat & 15) ||
      !d->x || !d->xhat || !d->rstd)
    return pir_fail(PIR_ERR_ARG, "%s: tensors missing or not 16-byte aligned", who);
  if (bwd && (!d->g || (d->g_pitch % 8) || (d->g_bstride % 8) || ((uintptr_t)d->g & 15))) return pir_fail(PIR_ERR_ARG, "%s: g missing or misaligned", who);
  return PIR_OK;
}

// ------------------------------------------------------------------------------------------------------
// GDFN gate backward (exact erf GELU), in place on y = [y1 | y2]
// ------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
gate_bwd_kernel(unsigned short* __restrict__ y, long long ypitch, long long ybs, const unsigned short* __restrict__ dg, long long gpitch,
                long long gbs, int HW, int C, long long total) {
  const int chunks = C >> 3;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % chunks);
    const long long p = e / chunks;
    const int b = (int)(p / HW), pi = (int)(p % HW);
    unsigned short* yp = y + (size_t)b * ybs + (size_t)pi * ypitch + ch * 8;
    float y1[8], y2[8], d[8];
    unpack8t<T>(*reinterpret_cast<const uint4*>(yp), y1);
    unpack8t<T>(*reinterpret_cast<const uint4*>(yp + C), y2);
    unpack8t<T>(__ldg(reinterpret_cast<const uint4*>(dg + (size_t)b * gbs + (size_t)pi * gpitch + ch * 8)), d);
    float o1[8], o2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float cdf = 0.5f * (1.0f + erff(y1[i] * 0.7071067811865476f));
      const float pdf = expf(-0.5f * y1[i] * y1[i]) * 0.3989422804014327f;
      o1[i] = d[i] * y2[i] * fmaf(y1[i], pdf, cdf);
      o2[i] = d[i] * y1[i] * cdf;
    }
    *reinterpret_cast<uint4*>(yp) = pack8t<T>(o1);
    *reinterpret_cast<uint4*>(yp + C) = pack8t<T>(o2);
  }
}

// ------------------------------------------------------------------------------------------------------
// PixelShuffle(2) / PixelUnshuffle(2): channel index of the 4c side = ch*4 + i*2 + j  <->  pixel (2y+i, 2x+j), channel ch
// one thread: one low-resolution pixel x 8 channels of the c side = 32 contiguous channels of the 4c side
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
shuffle_kernel(const unsigned short* __restrict__ in, long long ipitch, long long ibs, unsigned short* __restrict__ out, long long opitch,
               long long obs, int H, int W, int c, int up, long long total) {
  const int chunks = c >> 3;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % chunks);
    long long p = e / chunks;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    unsigned short v[4][8];                       // [i*2+j][channel of the c side]
    if (up) {                                     // read 32 channels of the 4c side
      const unsigned short* ip = in + (size_t)b * ibs + ((size_t)y * W + x) * ipitch + ch * 32;
      unsigned short w[32];
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(w + q * 8) = __ldg(reinterpret_cast<const uint4*>(ip + q * 8));
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int ij = 0; ij < 4; ++ij) v[ij][k] = w[k * 4 + ij];
#pragma unroll
      for (int ij = 0; ij < 4; ++ij) {
        unsigned short* op = out + (size_t)b * obs + ((size_t)(2 * y + (ij >> 1)) * (2 * W) + 2 * x + (ij & 1)) * opitch + ch * 8;
        *reinterpret_cast<uint4*>(op) = *reinterpret_cast<const uint4*>(v[ij]);
      }
    } else {
#pragma unroll
      for (int ij = 0; ij < 4; ++ij) {
        const unsigned short* ip = in + (size_t)b * ibs + ((size_t)(2 * y + (ij >> 1)) * (2 * W) + 2 * x + (ij & 1)) * ipitch + ch * 8;
        *reinterpret_cast<uint4*>(v[ij]) = __ldg(reinterpret_cast<const uint4*>(ip));
      }
      unsigned short w[32];
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int ij = 0; ij < 4; ++ij) w[k * 4 + ij] = v[ij][k];
      unsigned short* op = out + (size_t)b * obs + ((size_t)y * W + x) * opitch + ch * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(op + q * 8) = *reinterpret_cast<const uint4*>(w + q * 8);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// g[b,p,c] += v[b][c]
// ------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
bcast_add_kernel(unsigned short* __restrict__ g, long long pitch, long long bs, const float* __restrict__ v, int HW, int C, long long total) {
  const int chunks = C >> 3;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % chunks);
    const long long p = e / chunks;
    const int b = (int)(p / HW), pi = (int)(p % HW);
    unsigned short* gp = g + (size_t)b * bs + (size_t)pi * pitch + ch * 8;
    float f[8];
    unpack8t<T>(*reinterpret_cast<const uint4*>(gp), f);
    const float* vp = v + (size_t)b * C + ch * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += __ldg(vp + i);
    *reinterpret_cast<uint4*>(gp) = pack8t<T>(f);
  }
}

// fp32 NCHW -> 16-bit NHWC, 8 channels per pixel (C <= 8, the rest zero)
template <class T>
__global__ void __launch_bounds__(256)
to_nhwc16_kernel(const float* __restrict__ src, unsigned short* __restrict__ out, long long pitch, long long bs, int C, int HW, float scale,
                 long long total) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(p / HW), pi = (int)(p % HW);
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = c < C ? __ldg(src + ((size_t)b * C + c) * HW + pi) * scale : 0.f;
    *reinterpret_cast<uint4*>(out + (size_t)b * bs + (size_t)pi * pitch) = pack8t<T>(f);
  }
}

// ------------------------------------------------------------------------------------------------------
// depthwise 3x3 weight gradient  dw[c][tap] = sum_p dy[p, c] x[p + off(tap), c]  (+ bias gradient sum_p dy).
// Same structure as the forward stencil (dwconv.cu): persistent CTAs per 64-channel chunk walk (image, 8 x 32 pixel tile) items;
// one 4-D TMA box brings the x tile with its halo and one the dy tile into a double buffer (zero fill outside the image = the
// conv's zero padding, and zero dy outside = no contribution).  A thread owns a pixel column and 8 channels, walks down the rows
// with the last three dy rows in registers and keeps its 9 + 1 fp32 accumulators x 8 channels for the CTA's whole lifetime
// (FHFMA: 16-bit x 16-bit + fp32); one deterministic block reduction at the end writes ws[cta][10][C].
// ------------------------------------------------------------------------------------------------------
struct DwwArgs {
  int H, W, C;
  int tiles_x, tiles_y, n_sp;
  float* ws;
};

template <class T, int TH>
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, const DwwArgs a) {
  constexpr int SW = 34, RB = 128;
  constexpr uint32_t XT = (TH + 2) * SW * RB, DT = TH * 32 * RB, BUF = XT + DT;
  extern __shared__ uint8_t dww_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ float red[8][10][64];
  uint8_t* sm = dww_raw + ((128u - (smem_u32(dww_raw) & 127u)) & 127u);
  const int tid = threadIdx.x, cg = tid & 7, tx = tid >> 3, warp = tid >> 5;
  const int c0 = blockIdx.y * 64;
  const uint32_t bar_a[2] = {smem_u32(&bar[0]), smem_u32(&bar[1])};
  const int per_img = a.tiles_x * a.tiles_y;
  auto issue = [&](int sp, int buf) {      // thread 0 only
    const int b = sp / per_img, r = sp % per_img;
    const int x0 = (r % a.tiles_x) * 32, y0 = (r / a.tiles_x) * TH;
    mbar_expect_tx(bar_a[buf], BUF);
    tma_load_4d(smem_u32(sm) + buf * BUF, &tmX, bar_a[buf], c0, x0 - 1, y0 - 1, b);
    tma_load_4d(smem_u32(sm) + buf * BUF + XT, &tmD, bar_a[buf], c0, x0, y0, b);
  };
  if (tid == 0) {
    mbar_init(bar_a[0], 1);
    mbar_init(bar_a[1], 1);
    fence_barrier_init();
    if ((int)blockIdx.x < a.n_sp) issue(blockIdx.x, 0);
  }
  float acc[10][8];
#pragma unroll
  for (int t = 0; t < 10; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  __syncthreads();
  int it = 0;
  for (int sp = blockIdx.x; sp < a.n_sp; sp += gridDim.x, ++it) {
    const int buf = it & 1;
    if (tid == 0 && sp + (int)gridDim.x < a.n_sp) issue(sp + gridDim.x, buf ^ 1);
    mbar_wait(bar_a[buf], (it >> 1) & 1);
    const uint8_t* xs = sm + (size_t)buf * BUF + (size_t)tx * RB + cg * 16;
    const uint8_t* ds = sm + (size_t)buf * BUF + XT + (size_t)tx * RB + cg * 16;
    uint4 dyv[3];
#pragma unroll
    for (int r = 0; r < TH + 2; ++r) {
      if (r < TH) {
        dyv[r % 3] = *reinterpret_cast<const uint4*>(ds + (size_t)r * 32 * RB);
        const uint32_t dq[4] = {dyv[r % 3].x, dyv[r % 3].y, dyv[r % 3].z, dyv[r % 3].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[9][2 * q] += unpack_lo<T>(dq[q]); acc[9][2 * q + 1] += unpack_hi<T>(dq[q]); }
      }
      uint4 xv[3];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) xv[kx] = *reinterpret_cast<const uint4*>(xs + ((size_t)r * SW + kx) * RB);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int o = r - ky;                        // x row r is tap row ky of output row o
        if (o >= 0 && o < TH) {
          const uint32_t dq[4] = {dyv[o % 3].x, dyv[o % 3].y, dyv[o % 3].z, dyv[o % 3].w};
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint32_t xw[4] = {xv[kx].x, xv[kx].y, xv[kx].z, xv[kx].w};
            float* ac = acc[ky * 3 + kx];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              ac[2 * q] = fma16<T>(lo16(xw[q]), lo16(dq[q]), ac[2 * q]);
              ac[2 * q + 1] = fma16<T>(hi16(xw[q]), hi16(dq[q]), ac[2 * q + 1]);
            }
          }
        }
      }
    }
    __syncthreads();                 // everyone is done with this buffer before it is refilled
  }
  // reduce over the 4 pixel columns of a warp, then over the 8 warps
#pragma unroll
  for (int t = 0; t < 10; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[t][i];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((tid & 31) < 8) red[warp][t][cg * 8 + i] = v;
    }
  __syncthreads();
  for (int e = tid; e < 10 * 64; e += 256) {
    const int t = e >> 6, ch = e & 63;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][t][ch];
    if (c0 + ch < a.C) a.ws[((size_t)blockIdx.x * 10 + t) * a.C + c0 + ch] = s;
  }
}

// 32 consecutive (tap, channel) elements x 8 partial lanes per block: the serial chain over the per-CTA partials is parts / 8 long
__global__ void __launch_bounds__(256)
dw_wgrad_fin_kernel(const float* __restrict__ ws, int parts, int C, int R, int half, int half_pad, float inv_scale, float* __restrict__ dst_w,
                    float* __restrict__ dst_bias) {
  __shared__ float red[8][33];
  const int le = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + le;                  // element = t * R + r  (rows fastest: consecutive channels -> coalesced)
  const bool ok = e < R * 10;
  const int t = ok ? e / R : 0, r = ok ? e % R : 0;
  const int pr = r < half ? r : r - half + half_pad;
  float s0 = 0.f, s1 = 0.f;
  if (ok) {
    const float* p0 = ws + (size_t)t * C + pr;
    int p = sl;
    for (; p + 8 < parts; p += 16) { s0 += p0[(size_t)p * 10 * C]; s1 += p0[(size_t)(p + 8) * 10 * C]; }
    if (p < parts) s0 += p0[(size_t)p * 10 * C];
  }
  red[sl][le] = s0 + s1;
  __syncthreads();
  if (sl == 0 && ok) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q][le];
    s *= inv_scale;
    if (t < 9) dst_w[(size_t)r * 9 + t] = s;
    else if (dst_bias) dst_bias[r] = s;
  }
}

// ------------------------------------------------------------------------------------------------------
// MDTA backward on the channel x channel matrices.  Scratch layout (floats), per pir_mdta_bwd_ws_floats():
//   nrm [B][2C] | dWf [B][C][C] | dcos [B][C][c] | cosm [B][C][c] | rq [B][C] | rk [B][C] | dTp [B][C] | dWoP [B][C][C]
// ------------------------------------------------------------------------------------------------------
struct MbArgs {
  int B, C, heads, c, sf, sb;
  const float* gram;      // forward: [B][sf][C][c]
  const float* norm;      // forward: [B][sf][2][C]
  const float* attn;      // forward: [B][heads][c][c]
  const float* ws_b;      // [B*sb][C][C]
  const float* colsum_b;  // [B*sb][C] or null
  const float* temperature; const float* wo;
  float inv_scale;
  float *nrm, *dWf, *dcos, *cosm, *rq, *rk, *dTp, *dWoP;
  float *dst_wo, *dst_temp, *dst_bias;
};

// eight loads in flight per round; even terms onto s0, odd terms onto s1, in order
__device__ __forceinline__ float sum_strided(const float* p, size_t stride, int n) {
  float s0 = 0.f, s1 = 0.f;
  for (int i = 0; i < n; i += 8) {
    float l[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) l[u] = i + u < n ? p[(size_t)(i + u) * stride] : 0.f;
    s0 += l[0]; s1 += l[1]; s0 += l[2]; s1 += l[3]; s0 += l[4]; s1 += l[5]; s0 += l[6]; s1 += l[7];
  }
  return s0 + s1;
}

// k1: norms, the per-image dWfold and the forward Gram sums (raw, into the cos buffer; k2 normalises them).  grid (ceil(C*C/256), B)
__global__ void __launch_bounds__(256) mdta_bwd_reduce_kernel(const MbArgs a) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e < a.C * a.C) a.dWf[(size_t)b * a.C * a.C + e] = sum_strided(a.ws_b + (size_t)b * a.sb * a.C * a.C + e, (size_t)a.C * a.C, a.sb);
  if (e < a.C * a.c) a.cosm[(size_t)b * a.C * a.c + e] = sum_strided(a.gram + (size_t)b * a.sf * a.C * a.c + e, (size_t)a.C * a.c, a.sf);
  if (e < 2 * a.C)
    a.nrm[(size_t)b * 2 * a.C + e] = fmaxf(sqrtf(sum_strided(a.norm + (size_t)b * a.sf * 2 * a.C + e, (size_t)2 * a.C, a.sf)), 1e-12f);
}

// Strided, batched fp32 GEMM for the channel x channel products of the MDTA backward (three per block, up to 704 x 176 x 704 per
// image): out[z][m][n] = sum_k A[z][m][k] B[z][k][n], z = (image, head).  64 x 64 tile, K chunks of 16, 256 threads x (4 x 4) outputs.
// Operand strides are free (the matrices are transposed / head-sliced views), so tile loads pick the unit-stride axis for coalescing.
struct SgArgs {
  int M, N, K, heads;
  const float* A; long long a_sm, a_sk, a_sb, a_sh;
  const float* B; long long b_sk, b_sn, b_sb, b_sh;
  float* out; unsigned short* out16; long long o_sm, o_sn, o_sb, o_sh;
};

// Up to three independent problems per launch (the three products of pir_mdta_bwd only depend on its first kernel): a flat grid,
// each problem owns a contiguous range of blocks = (image, head) x its own m x n tiles.
struct SgMulti {
  SgArgs g[3];
  int bend[3];        // exclusive prefix ends of each problem's block range
  int nt_m[3], nt_n[3];
};

template <class T>
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const SgMulti mp) {
  __shared__ __align__(16) float As[16][64];
  __shared__ __align__(16) float Bs[16][64];
  const int which = (int)blockIdx.x < mp.bend[0] ? 0 : ((int)blockIdx.x < mp.bend[1] ? 1 : 2);
  const SgArgs& g = mp.g[which];
  const int idx = (int)blockIdx.x - (which ? mp.bend[which - 1] : 0);
  const int tiles = mp.nt_m[which] * mp.nt_n[which];
  const int z = idx / tiles, tile = idx % tiles, b = z / g.heads, h = z % g.heads;
  const int m0 = (tile / mp.nt_n[which]) * 64, n0 = (tile % mp.nt_n[which]) * 64;
  const float* A = g.A + (size_t)b * g.a_sb + (size_t)h * g.a_sh;
  const float* Bm = g.B + (size_t)b * g.b_sb + (size_t)h * g.b_sh;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // tile coordinates of the four A and four B elements this thread stages per K chunk (fixed for the whole loop)
  int am[4], ak[4], bn[4], bk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (g.a_sm == 1) { am[i] = t & 63; ak[i] = (t >> 6) + 4 * i; } else { ak[i] = t & 15; am[i] = (t >> 4) + 16 * i; }
    if (g.b_sn == 1) { bn[i] = t & 63; bk[i] = (t >> 6) + 4 * i; } else { bk[i] = t & 15; bn[i] = (t >> 4) + 16 * i; }
  }
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ra[i] = (m0 + am[i] < g.M && k0 + ak[i] < g.K) ? A[(size_t)(m0 + am[i]) * g.a_sm + (size_t)(k0 + ak[i]) * g.a_sk] : 0.f;
      rb[i] = (n0 + bn[i] < g.N && k0 + bk[i] < g.K) ? Bm[(size_t)(k0 + bk[i]) * g.b_sk + (size_t)(n0 + bn[i]) * g.b_sn] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[ak[i]][am[i]] = ra[i]; Bs[bk[i]][bn[i]] = rb[i]; }
    __syncthreads();
    if (k0 + 16 < g.K) fetch(k0 + 16);              // next chunk's loads fly while this one is multiplied
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      const size_t o = (size_t)b * g.o_sb + (size_t)h * g.o_sh + (size_t)m * g.o_sm + (size_t)n * g.o_sn;
      if (g.out16) g.out16[o] = to16<T>(acc[i][j]);
      else g.out[o] = acc[i][j];
    }
  }
}

// Tensor-core form of the same strided batched GEMM (default): fp32 operands are split into TF32 high and low parts at fragment
// load and every product is three mma.m16n8k8 (lo x hi, hi x lo, hi x hi: "3xTF32", fp32-level accuracy, error ~2^-21 relative).
// 128 threads, 64 x 64 tile, warp w owns the 32 x 32 quadrant (w >> 1, w & 1); K chunks of 16 staged through shared memory with the
// next chunk's global loads in flight, rows padded to 72 floats so the fragment reads (lane = (g, q4) -> word 8 q4 + g) are
// conflict free.  PIR_SGEMM_SIMT=1 selects the FFMA kernel above (A/B).
__device__ __forceinline__ uint32_t f2tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <class T>
__global__ void __launch_bounds__(128, 4) tgemm_strided_kernel(const SgMulti mp) {
  constexpr int LD = 72;
  __shared__ __align__(16) float As[16][LD];
  __shared__ __align__(16) float Bs[16][LD];
  const int which = (int)blockIdx.x < mp.bend[0] ? 0 : ((int)blockIdx.x < mp.bend[1] ? 1 : 2);
  const SgArgs& g = mp.g[which];
  const int idx = (int)blockIdx.x - (which ? mp.bend[which - 1] : 0);
  const int tiles = mp.nt_m[which] * mp.nt_n[which];
  const int z = idx / tiles, tile = idx % tiles, b = z / g.heads, h = z % g.heads;
  const int m0 = (tile / mp.nt_n[which]) * 64, n0 = (tile % mp.nt_n[which]) * 64;
  const float* A = g.A + (size_t)b * g.a_sb + (size_t)h * g.a_sh;
  const float* Bm = g.B + (size_t)b * g.b_sb + (size_t)h * g.b_sh;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, gq = lane >> 2, q4 = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f; }
  // tile coordinates of the eight A and eight B elements this thread stages per K chunk: the unit-stride axis goes across lanes
  const bool a_mfast = g.a_sm == 1, b_nfast = g.b_sn == 1;
  auto a_m = [&](int i) { return a_mfast ? (t & 63) : (t >> 4) + 8 * i; };
  auto a_k = [&](int i) { return a_mfast ? (t >> 6) + 2 * i : (t & 15); };
  auto b_n = [&](int i) { return b_nfast ? (t & 63) : (t >> 4) + 8 * i; };
  auto b_k = [&](int i) { return b_nfast ? (t >> 6) + 2 * i : (t & 15); };
  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int am = a_m(i), ak = a_k(i), bn = b_n(i), bk = b_k(i);
      ra[i] = (m0 + am < g.M && k0 + ak < g.K) ? A[(size_t)(m0 + am) * g.a_sm + (size_t)(k0 + ak) * g.a_sk] : 0.f;
      rb[i] = (n0 + bn < g.N && k0 + bk < g.K) ? Bm[(size_t)(k0 + bk) * g.b_sk + (size_t)(n0 + bn) * g.b_sn] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { As[a_k(i)][a_m(i)] = ra[i]; Bs[b_k(i)][b_n(i)] = rb[i]; }
    __syncthreads();
    if (k0 + 16 < g.K) fetch(k0 + 16);              // next chunk's loads fly while this one is multiplied
#pragma unroll
    for (int ks = 0; ks < 16; ks += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float v[4] = {As[ks + q4][wm + 16 * mt + gq], As[ks + q4][wm + 16 * mt + gq + 8], As[ks + q4 + 4][wm + 16 * mt + gq],
                            As[ks + q4 + 4][wm + 16 * mt + gq + 8]};
#pragma unroll
        for (int e = 0; e < 4; ++e) { ah[mt][e] = f2tf32(v[e]); al[mt][e] = f2tf32(v[e] - __uint_as_float(ah[mt][e])); }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float v[2] = {Bs[ks + q4][wn + 8 * nt + gq], Bs[ks + q4 + 4][wn + 8 * nt + gq]};
#pragma unroll
        for (int e = 0; e < 2; ++e) { bh[nt][e] = f2tf32(v[e]); bl[nt][e] = f2tf32(v[e] - __uint_as_float(bh[nt][e])); }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma_tf32(acc[mt][nt], al[mt], bh[nt]);
          mma_tf32(acc[mt][nt], ah[mt], bl[nt]);
          mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = m0 + wm + 16 * mt + gq + 8 * (e >> 1), n = n0 + wn + 8 * nt + 2 * q4 + (e & 1);
        if (m >= g.M || n >= g.N) continue;
        const size_t o = (size_t)b * g.o_sb + (size_t)h * g.o_sh + (size_t)m * g.o_sm + (size_t)n * g.o_sn;
        if (g.out16) g.out16[o] = to16<T>(acc[mt][nt][e]);
        else g.out[o] = acc[mt][nt][e];
      }
}

// k2: one warp per attention row (b, r = h*c + i): softmax backward of dA (in place -> dcos), cos, rq, temperature partial.  c <= 32 NT
template <int NT>
__global__ void __launch_bounds__(256) mdta_bwd_rows_kernel(const MbArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  const int b = blockIdx.y;
  if (r >= a.C) return;
  const int c = a.c, C = a.C;
  const int h = r / c, i = r - h * c;
  const float qn = a.nrm[(size_t)b * 2 * C + r];
  const float T = a.temperature[h];
  float dA[NT], A[NT], cs[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {                       // dA = Wo_h^T dWf_h was left in the dcos buffer by the SGEMM
    const int j = lane + t * 32;
    dA[t] = j < c ? a.dcos[((size_t)b * C + r) * c + j] : 0.f;
  }
  float dot = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int j = lane + t * 32;
    A[t] = 0.f; cs[t] = 0.f;
    if (j < c) {
      A[t] = a.attn[((size_t)(b * a.heads + h) * c + i) * c + j];
      const float kn = a.nrm[(size_t)b * 2 * C + C + h * c + j];
      cs[t] = a.cosm[((size_t)b * C + r) * c + j] / (qn * kn);
      dot = fmaf(dA[t], A[t], dot);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  float dT = 0.f, rq = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int j = lane + t * 32;
    if (j < c) {
      const float dS = A[t] * (dA[t] - dot);
      dT = fmaf(dS, cs[t], dT);
      const float dc = dS * T;
      rq = fmaf(dc, cs[t], rq);
      a.dcos[((size_t)b * C + r) * c + j] = dc;
      a.cosm[((size_t)b * C + r) * c + j] = cs[t];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { dT += __shfl_xor_sync(0xffffffffu, dT, o); rq += __shfl_xor_sync(0xffffffffu, rq, o); }
  if (lane == 0) {
    a.dTp[(size_t)b * C + r] = dT;
    a.rq[(size_t)b * C + r] = rq / (qn * qn);
  }
}

// k3: rk[b][h*c + j] = sum_i dcos[i][j] cos[i][j] / kn_j^2.  one warp per (b, channel): lanes over i, fixed-order butterfly
__global__ void __launch_bounds__(256) mdta_bwd_cols_kernel(const MbArgs a) {
  const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= a.B * a.C) return;
  const int b = e / a.C, ch = e % a.C;
  const int h = ch / a.c, j = ch - h * a.c;
  const size_t base = ((size_t)b * a.C + h * a.c) * a.c + j;
  float s = 0.f;
  for (int i = lane; i < a.c; i += 32) s = fmaf(a.dcos[base + (size_t)i * a.c], a.cosm[base + (size_t)i * a.c], s);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const float kn = a.nrm[(size_t)b * 2 * a.C + a.C + ch];
    a.rk[e] = s / (kn * kn);
  }
}

// k4: wqk[b] (softmax / cosine / normalisation backward as a 2C x 2C matrix).  grid (ceil(2C * 2C / 8 / 256), B): eight consecutive k
// per thread (2C % 8 == 0), one 16-byte store; most groups lie outside the head blocks and are plain zeros.
template <class T>
__global__ void __launch_bounds__(256) mdta_bwd_weights_kernel(const MbArgs a, unsigned short* __restrict__ wqk, int kpad2) {
  const int b = blockIdx.y;
  const int C = a.C, c = a.c;
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  const int groups = 2 * C / 8;
  if (e >= (long long)2 * C * groups) return;
  const float* nrm = a.nrm + (size_t)b * 2 * C;
  const int n = (int)(e / groups), k0 = (int)(e % groups) * 8;
  float v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int k = k0 + u;
    float x = 0.f;
    if (n < C) {
      if (k < C) { if (k == n) x = -a.rq[(size_t)b * C + n]; }
      else {
        const int kk = k - C;
        if (kk / c == n / c) x = a.dcos[((size_t)b * C + n) * c + kk % c] / (nrm[n] * nrm[C + kk]);
      }
    } else {
      const int jg = n - C;
      if (k < C) { if (k / c == jg / c) x = a.dcos[((size_t)b * C + k) * c + jg % c] / (nrm[k] * nrm[C + jg]); }
      else if (k == n) x = -a.rk[(size_t)b * C + jg];
    }
    // fp16 storage: 1 / (|q| |k|) factors of degenerate (single-pixel) levels times the loss scale can leave the fp16 range -> saturate
    v[u] = T::kFmt ? x : fminf(fmaxf(x, -65504.f), 65504.f);
  }
  uint4 o;
  o.x = (uint32_t)to16<T>(v[0]) | ((uint32_t)to16<T>(v[1]) << 16);
  o.y = (uint32_t)to16<T>(v[2]) | ((uint32_t)to16<T>(v[3]) << 16);
  o.z = (uint32_t)to16<T>(v[4]) | ((uint32_t)to16<T>(v[5]) << 16);
  o.w = (uint32_t)to16<T>(v[6]) | ((uint32_t)to16<T>(v[7]) << 16);
  *reinterpret_cast<uint4*>(wqk + ((size_t)b * 2 * C + n) * kpad2 + k0) = o;
}

// k5b + k6, one launch: block 0 = temperature gradient (one warp per head), then ceil(C/256) blocks of project_out bias gradient
// (when there is a bias), then ceil(C*C/256) blocks of dWo = inv * sum_b dWoP[b]
__global__ void __launch_bounds__(256) mdta_bwd_small_kernel(const MbArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbias = a.dst_bias ? (a.C + 255) / 256 : 0;
  if (blockIdx.x == 0) {
    if (warp < a.heads) {
      float s = 0.f;
#pragma unroll 8
      for (int e = lane; e < a.B * a.c; e += 32) s += a.dTp[(size_t)(e / a.c) * a.C + warp * a.c + e % a.c];
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) a.dst_temp[warp] = s * a.inv_scale;
    }
  } else if ((int)blockIdx.x <= nbias) {
    const int o = (blockIdx.x - 1) * 256 + threadIdx.x;
    if (o < a.C) a.dst_bias[o] = sum_strided(a.colsum_b + o, (size_t)a.C, a.B * a.sb) * a.inv_scale;
  } else {
    const int e = (blockIdx.x - 1 - nbias) * 256 + threadIdx.x;
    if (e < a.C * a.C) a.dst_wo[e] = sum_strided(a.dWoP + e, (size_t)a.C * a.C, a.B) * a.inv_scale;
  }
}

// ------------------------------------------------------------------------------------------------------
// PromptGenBlock backward.  scratch: dmix [B][S][S][D] | dw [B][L] | emb [B][C]
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int y, float scale, int S, int align, int* y0, int* y1, float* ly) {
  const float fy = align ? scale * (float)y : fmaxf(scale * ((float)y + 0.5f) - 0.5f, 0.f);
  *y0 = min((int)fy, S - 1);
  *y1 = *y0 + (*y0 < S - 1 ? 1 : 0);
  *ly = fminf(fmaxf(fy - (float)*y0, 0.f), 1.f);
}

// dmix[b][s][t][d] = sum over destination pixels whose bilinear footprint contains (s, t).  one thread per (b, s, t, 8 channels)
template <class T>
__global__ void __launch_bounds__(256)
prompt_bwd_dmix_kernel(const unsigned short* __restrict__ dup, long long pitch, long long bs, int H, int W, int D, int S, float* __restrict__ dmix,
                       long long total, int align) {
  const int chunks = D >> 3;
  const float sh = align ? (H > 1 ? (float)(S - 1) / (float)(H - 1) : 0.f) : (float)S / (float)H;
  const float sw = align ? (W > 1 ? (float)(S - 1) / (float)(W - 1) : 0.f) : (float)S / (float)W;
  // destination pixels whose footprint can contain a source index: within ~1/scale of its centre (the whole axis if scale is 0)
  const int ry = sh > 0.f ? (int)ceilf(1.0f / sh) + 1 : H, rx = sw > 0.f ? (int)ceilf(1.0f / sw) + 1 : W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % chunks);
    long long p = e / chunks;
    const int t = (int)(p % S); p /= S;
    const int s = (int)(p % S);
    const int b = (int)(p / S);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const int yc = sh > 0.f ? (int)(((float)s + (align ? 0.f : 0.5f)) / sh) : 0, xc = sw > 0.f ? (int)(((float)t + (align ? 0.f : 0.5f)) / sw) : 0;
    for (int y = max(yc - 2 * ry, 0); y <= min(yc + 2 * ry, H - 1); ++y) {
      int y0, y1; float ly;
      bilinear_src(y, sh, S, align, &y0, &y1, &ly);
      const float wy = (y0 == s ? 1.f - ly : 0.f) + (y1 == s ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int x = max(xc - 2 * rx, 0); x <= min(xc + 2 * rx, W - 1); ++x) {
        int x0, x1; float lx;
        bilinear_src(x, sw, S, align, &x0, &x1, &lx);
        const float wx = (x0 == t ? 1.f - lx : 0.f) + (x1 == t ? lx : 0.f);
        if (wx == 0.f) continue;
        float f[8];
        unpack8t<T>(__ldg(reinterpret_cast<const uint4*>(dup + (size_t)b * bs + ((size_t)y * W + x) * pitch + ch * 8)), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(wy * wx, f[i], acc[i]);
      }
    }
    float* o = dmix + (((size_t)b * S + s) * S + t) * D + ch * 8;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// dw[b][l] = <dmix[b], prompt[l]>.  grid (L, B), block 256
__global__ void __launch_bounds__(256)
prompt_bwd_dot_kernel(const float* __restrict__ dmix, const float* __restrict__ prompt, long long n, int L, float* __restrict__ dw) {
  __shared__ float red[8];
  const int l = blockIdx.x, b = blockIdx.y;
  const float4* a = reinterpret_cast<const float4*>(dmix + (size_t)b * n);
  const float4* p = reinterpret_cast<const float4*>(prompt + (size_t)l * n);
  float s = 0.f;
  for (long long i = threadIdx.x; i < n / 4; i += 256) {
    const float4 u = a[i], v = __ldg(p + i);
    s += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    dw[b * L + l] = t;
  }
}

// softmax / linear / mean-pool backward.  one block; B*L <= 8*... small
__global__ void __launch_bounds__(256)
prompt_bwd_small_kernel(const float* __restrict__ dw, const float* __restrict__ wts, const float* __restrict__ pool_ws, int nchunks, int B, int L,
                        int C, int HW, const float* __restrict__ lin_w, float inv_scale, float* __restrict__ dlog_s, float* __restrict__ demb,
                        float* __restrict__ dst_lin_w, float* __restrict__ dst_lin_b) {
  // dlog[b][l] = w (dw - sum_l w dw)
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float dot = 0.f;
    for (int l = 0; l < L; ++l) dot = fmaf(wts[b * L + l], dw[b * L + l], dot);
    for (int l = 0; l < L; ++l) dlog_s[b * L + l] = wts[b * L + l] * (dw[b * L + l] - dot);
  }
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dlog_s[b * L + l];
    dst_lin_b[l] = s * inv_scale;
  }
  const float invHW = 1.0f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float gl[kMaxLTrain];
    for (int l = 0; l < L; ++l) gl[l] = 0.f;
    for (int b = 0; b < B; ++b) {
      float emb = 0.f;
      for (int k = 0; k < nchunks; ++k) emb += pool_ws[((size_t)b * nchunks + k) * C + c];
      emb *= invHW;
      float de = 0.f;
      for (int l = 0; l < L; ++l) {
        gl[l] = fmaf(dlog_s[b * L + l], emb, gl[l]);
        de = fmaf(dlog_s[b * L + l], lin_w[(size_t)l * C + c], de);
      }
      demb[(size_t)b * C + c] = de * invHW;
    }
    for (int l = 0; l < L; ++l) dst_lin_w[(size_t)l * C + c] = gl[l] * inv_scale;
  }
}

// dprompt[0][l][d][s][t] = inv * sum_b w[b][l] dmix[b][s][t][d].  one thread per (d, s*S + t), all L
__global__ void __launch_bounds__(256)
prompt_bwd_param_kernel(const float* __restrict__ dmix, const float* __restrict__ wts, int B, int L, int D, int SS, float inv_scale,
                        float* __restrict__ dst) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= (long long)D * SS) return;
  const int st = (int)(e % SS), d = (int)(e / SS);
  float acc[kMaxLTrain];
  for (int l = 0; l < L; ++l) acc[l] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float v = dmix[((size_t)b * SS + st) * D + d];
    for (int l = 0; l < L; ++l) acc[l] = fmaf(wts[b * L + l], v, acc[l]);
  }
  for (int l = 0; l < L; ++l) dst[((size_t)l * D + d) * SS + st] = acc[l] * inv_scale;
}

static inline unsigned grid_for(long long total, int threads = 256, int cap = 148 * 16) {
  long long b = (total + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace pir

using namespace pir;

#define PIR_BY_DTYPE(dt, expr_bf, expr_fp) \
  do { if ((dt) == PIR_DTYPE_BF16) { expr_bf; } else { expr_fp; } } while (0)

extern "C" int pir_ln_fwd(const PirLn* d, void* stream) {
  if (int e = check_ln(d, false, "pir_ln_fwd")) return e;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return d->dtype == PIR_DTYPE_BF16 ? launch_ln<BF16>(d, false, s) : launch_ln<FP16>(d, false, s);
}

extern "C" int pir_ln_bwd(const PirLn* d, void* stream) {
  if (int e = check_ln(d, true, "pir_ln_bwd")) return e;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return d->dtype == PIR_DTYPE_BF16 ? launch_ln<BF16>(d, true, s) : launch_ln<FP16>(d, true, s);
}

extern "C" int pir_gate_bwd(const PirGateBwd* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_gate_bwd: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0) return pir_fail(PIR_ERR_ARG, "pir_gate_bwd: empty problem");
  if ((d->C % 8) || (d->y_pitch % 8) || (d->dg_pitch % 8) || (d->y_bstride % 8) || (d->dg_bstride % 8) || ((uintptr_t)d->y & 15) ||
      ((uintptr_t)d->dg & 15) || !d->y || !d->dg)
    return pir_fail(PIR_ERR_ARG, "pir_gate_bwd: tensors missing or not 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)d->B * d->H * d->W * (d->C / 8);
  auto* y = reinterpret_cast<unsigned short*>(d->y);
  auto* dg = reinterpret_cast<const unsigned short*>(d->dg);
  PIR_BY_DTYPE(d->dtype,
               (gate_bwd_kernel<BF16><<<grid_for(total), 256, 0, s>>>(y, d->y_pitch, d->y_bstride, dg, d->dg_pitch, d->dg_bstride, d->H * d->W, d->C, total)),
               (gate_bwd_kernel<FP16><<<grid_for(total), 256, 0, s>>>(y, d->y_pitch, d->y_bstride, dg, d->dg_pitch, d->dg_bstride, d->H * d->W, d->C, total)));
  return pir_check_launch("pir_gate_bwd");
}

extern "C" int pir_pixel_shuffle(const PirShuffle* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_pixel_shuffle: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0) return pir_fail(PIR_ERR_ARG, "pir_pixel_shuffle: empty problem");
  if ((d->C % 32) || (d->in_pitch % 8) || (d->out_pitch % 8) || (d->in_bstride % 8) || (d->out_bstride % 8) || ((uintptr_t)d->in & 15) ||
      ((uintptr_t)d->out & 15) || !d->in || !d->out)
    return pir_fail(PIR_ERR_ARG, "pir_pixel_shuffle: the 4c side needs C %% 32 == 0 and 16-byte aligned tensors");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int c = d->C / 4;
  const long long total = (long long)d->B * d->H * d->W * (c / 8);
  shuffle_kernel<<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const unsigned short*>(d->in), d->in_pitch, d->in_bstride,
                                                 reinterpret_cast<unsigned short*>(d->out), d->out_pitch, d->out_bstride, d->H, d->W, c, d->up, total);
  return pir_check_launch("pir_pixel_shuffle");
}

extern "C" int pir_bcast_add(const PirBcastAdd* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_bcast_add: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0) return pir_fail(PIR_ERR_ARG, "pir_bcast_add: empty problem");
  if ((d->C % 8) || (d->g_pitch % 8) || (d->g_bstride % 8) || ((uintptr_t)d->g & 15) || !d->g || !d->v)
    return pir_fail(PIR_ERR_ARG, "pir_bcast_add: tensors missing or not 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)d->B * d->H * d->W * (d->C / 8);
  auto* g = reinterpret_cast<unsigned short*>(d->g);
  PIR_BY_DTYPE(d->dtype, (bcast_add_kernel<BF16><<<grid_for(total), 256, 0, s>>>(g, d->g_pitch, d->g_bstride, d->v, d->H * d->W, d->C, total)),
               (bcast_add_kernel<FP16><<<grid_for(total), 256, 0, s>>>(g, d->g_pitch, d->g_bstride, d->v, d->H * d->W, d->C, total)));
  return pir_check_launch("pir_bcast_add");
}

extern "C" int pir_nchw32_to_nhwc16(const PirToNhwc16* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_nchw32_to_nhwc16: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0) return pir_fail(PIR_ERR_ARG, "pir_nchw32_to_nhwc16: empty problem");
  if (d->Cpad != 8 || d->C > 8) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_nchw32_to_nhwc16: C <= 8 = Cpad only");
  if ((d->out_pitch % 8) || (d->out_bstride % 8) || ((uintptr_t)d->out & 15) || !d->out || !d->src)
    return pir_fail(PIR_ERR_ARG, "pir_nchw32_to_nhwc16: tensors missing or not 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)d->B * d->H * d->W;
  auto* o = reinterpret_cast<unsigned short*>(d->out);
  PIR_BY_DTYPE(d->dtype, (to_nhwc16_kernel<BF16><<<grid_for(total), 256, 0, s>>>(d->src, o, d->out_pitch, d->out_bstride, d->C, d->H * d->W, d->scale, total)),
               (to_nhwc16_kernel<FP16><<<grid_for(total), 256, 0, s>>>(d->src, o, d->out_pitch, d->out_bstride, d->C, d->H * d->W, d->scale, total)));
  return pir_check_launch("pir_nchw32_to_nhwc16");
}

static void dww_plan(int B, int H, int W, int C, int* tiles_x, int* tiles_y, int* n_sp, int* chunks, int* workers) {
  constexpr int TH = 8;
  *tiles_x = (W + 31) / 32;
  *tiles_y = (H + TH - 1) / TH;
  *n_sp = *tiles_x * *tiles_y * B;
  *chunks = (C + 63) / 64;
  int w = (148 + *chunks - 1) / *chunks;            // one CTA per SM (the double buffer fills shared memory)
  if (w > *n_sp) w = *n_sp;
  *workers = w < 1 ? 1 : w;
}

extern "C" int pir_dw_wgrad_parts(int32_t B, int32_t H, int32_t W, int32_t C) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0) return 1;
  int tx, ty, n_sp, chunks, workers;
  dww_plan(B, H, W, C, &tx, &ty, &n_sp, &chunks, &workers);
  return workers;
}

template <class T>
static int launch_dw_wgrad(const PirDwWgrad* d, cudaStream_t s) {
  constexpr int TH = 8;
  DwwArgs a{};
  int chunks, workers;
  dww_plan(d->B, d->H, d->W, d->C, &a.tiles_x, &a.tiles_y, &a.n_sp, &chunks, &workers);
  if (workers != d->parts) return pir_fail(PIR_ERR_ARG, "pir_dw_wgrad: parts must be pir_dw_wgrad_parts(B, H, W, C) = %d", workers);
  a.H = d->H; a.W = d->W; a.C = d->C; a.ws = d->ws;
  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
  CUtensorMap tmX, tmD;
  {
    const uint64_t strides[3] = {(uint64_t)d->x_pitch * 2, (uint64_t)d->x_pitch * 2 * d->W, (uint64_t)d->x_bstride * 2};
    const uint32_t box[4] = {64, 34, TH + 2, 1};
    if (int e = pir_make_tmap(&tmX, dt, 4, d->x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
  }
  {
    const uint64_t strides[3] = {(uint64_t)d->dy_pitch * 2, (uint64_t)d->dy_pitch * 2 * d->W, (uint64_t)d->dy_bstride * 2};
    const uint32_t box[4] = {64, 32, TH, 1};
    if (int e = pir_make_tmap(&tmD, dt, 4, d->dy, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
  }
  const size_t smem = (size_t)2 * ((TH + 2) * 34 + TH * 32) * 128 + 128;
  if (!pir_smem_attr_once(reinterpret_cast<const void*>(dw_wgrad_kernel<T, TH>), (int)((int)smem), "pir_dw_wgrad")) return PIR_ERR_CUDA;
  dim3 grid((unsigned)workers, (unsigned)chunks);
  dw_wgrad_kernel<T, TH><<<grid, 256, smem, s>>>(tmX, tmD, a);
  return pir_check_launch("pir_dw_wgrad");
}

extern "C" int pir_dw_wgrad(const PirDwWgrad* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_dw_wgrad: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->parts <= 0 || d->R <= 0) return pir_fail(PIR_ERR_ARG, "pir_dw_wgrad: empty problem");
  if ((d->C % 8) || (d->x_pitch % 8) || (d->dy_pitch % 8) || (d->x_bstride % 8) || (d->dy_bstride % 8) || ((uintptr_t)d->x & 15) ||
      ((uintptr_t)d->dy & 15) || !d->x || !d->dy || !d->ws || !d->dst_w)
    return pir_fail(PIR_ERR_ARG, "pir_dw_wgrad: tensors missing or not 16-byte aligned");
  if ((d->R > d->half ? d->R - d->half + d->half_pad : d->R) > d->C) return pir_fail(PIR_ERR_ARG, "pir_dw_wgrad: parameter larger than the tensor");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (int e = d->dtype == PIR_DTYPE_BF16 ? launch_dw_wgrad<BF16>(d, s) : launch_dw_wgrad<FP16>(d, s)) return e;
  dw_wgrad_fin_kernel<<<(unsigned)((d->R * 10 + 31) / 32), 256, 0, s>>>(d->ws, d->parts, d->C, d->R, d->half, d->half_pad, d->inv_scale, d->dst_w,
                                                                         d->dst_bias);
  return pir_check_launch("pir_dw_wgrad(finalize)");
}

extern "C" int64_t pir_mdta_bwd_ws_floats(int32_t B, int32_t C, int32_t heads) {
  if (B <= 0 || C <= 0 || heads <= 0) return 0;
  const int64_t c = C / heads;
  return (int64_t)B * (2 * C + 2 * (int64_t)C * C + 2 * C * c + 3 * C);
}

extern "C" int pir_mdta_bwd(const PirMdtaBwd* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_mdta_bwd: null descriptor");
  if (d->B <= 0 || d->C <= 0 || d->heads <= 0 || d->splits_f <= 0 || d->splits_b <= 0) return pir_fail(PIR_ERR_ARG, "pir_mdta_bwd: empty problem");
  if (d->C % d->heads) return pir_fail(PIR_ERR_ARG, "pir_mdta_bwd: C must be divisible by heads");
  if (d->C % 8) return pir_fail(PIR_ERR_ARG, "pir_mdta_bwd: C must be a multiple of 8");
  if (d->C / d->heads > 768 || d->heads > 8) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_mdta_bwd: head dim > 768 or heads > 8");
  if (!d->ws_f || !d->ws_b || !d->temperature || !d->wo || !d->scratch || !d->wft || !d->wqk || !d->dst_wo || !d->dst_temp)
    return pir_fail(PIR_ERR_ARG, "pir_mdta_bwd: missing pointers");
  if (d->dst_bias && !d->colsum_b) return pir_fail(PIR_ERR_ARG, "pir_mdta_bwd: bias gradient needs the column sums");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MbArgs a{};
  const int B = d->B, C = d->C, c = C / d->heads;
  a.B = B; a.C = C; a.heads = d->heads; a.c = c; a.sf = d->splits_f; a.sb = d->splits_b;
  a.gram = d->ws_f;
  a.norm = d->ws_f + (size_t)B * a.sf * C * c;
  a.attn = d->ws_f + (size_t)B * a.sf * ((size_t)C * c + 2 * C);
  a.ws_b = d->ws_b; a.colsum_b = d->colsum_b; a.temperature = d->temperature; a.wo = d->wo; a.inv_scale = d->inv_scale;
  float* p = d->scratch;
  a.nrm = p; p += (size_t)B * 2 * C;
  a.dWf = p; p += (size_t)B * C * C;
  a.dcos = p; p += (size_t)B * C * c;
  a.cosm = p; p += (size_t)B * C * c;
  a.rq = p; p += (size_t)B * C;
  a.rk = p; p += (size_t)B * C;
  a.dTp = p; p += (size_t)B * C;
  a.dWoP = p;
  a.dst_wo = d->dst_wo; a.dst_temp = d->dst_temp; a.dst_bias = d->dst_bias;
  const bool bf = d->dtype == PIR_DTYPE_BF16;
  const int kpad1 = (C + 63) / 64 * 64, kpad2 = (2 * C + 63) / 64 * 64;
  const int heads = d->heads;
  mdta_bwd_reduce_kernel<<<dim3((C * C + 255) / 256, B), 256, 0, s>>>(a);
  if (int e = pir_check_launch("pir_mdta_bwd(reduce)")) return e;
  {
    SgMulti mp{};
    {   // dA_h[i][j] = sum_o Wo[o][hc+i] dWf[b][o][hc+j]  -> dcos buffer
      SgArgs& g = mp.g[0];
      g.M = c; g.N = c; g.K = C; g.heads = heads;
      g.A = d->wo; g.a_sm = 1; g.a_sk = C; g.a_sb = 0; g.a_sh = c;
      g.B = a.dWf; g.b_sk = C; g.b_sn = 1; g.b_sb = (long long)C * C; g.b_sh = c;
      g.out = a.dcos; g.o_sm = c; g.o_sn = 1; g.o_sb = (long long)C * c; g.o_sh = (long long)c * c;
    }
    {   // wft[b][hc+j][o] = sum_i A[b,h][i][j] Wo[o][hc+i]
      SgArgs& g = mp.g[1];
      g.M = c; g.N = C; g.K = c; g.heads = heads;
      g.A = a.attn; g.a_sm = 1; g.a_sk = c; g.a_sb = (long long)heads * c * c; g.a_sh = (long long)c * c;
      g.B = d->wo; g.b_sk = 1; g.b_sn = C; g.b_sb = 0; g.b_sh = c;
      g.out16 = reinterpret_cast<unsigned short*>(d->wft); g.o_sm = kpad1; g.o_sn = 1; g.o_sb = (long long)C * kpad1; g.o_sh = (long long)c * kpad1;
    }
    {   // dWoP[b][o][hc+i] = sum_j dWf[b][o][hc+j] A[b,h][i][j]
      SgArgs& g = mp.g[2];
      g.M = C; g.N = c; g.K = c; g.heads = heads;
      g.A = a.dWf; g.a_sm = C; g.a_sk = 1; g.a_sb = (long long)C * C; g.a_sh = c;
      g.B = a.attn; g.b_sk = 1; g.b_sn = c; g.b_sb = (long long)heads * c * c; g.b_sh = (long long)c * c;
      g.out = a.dWoP; g.o_sm = C; g.o_sn = 1; g.o_sb = (long long)C * C; g.o_sh = c;
    }
    int total = 0;
    for (int i = 0; i < 3; ++i) {
      mp.nt_m[i] = (mp.g[i].M + 63) / 64; mp.nt_n[i] = (mp.g[i].N + 63) / 64;
      total += mp.nt_m[i] * mp.nt_n[i] * B * heads;
      mp.bend[i] = total;
    }
    static const bool simt = getenv("PIR_SGEMM_SIMT") != nullptr;      // A/B: the FFMA kernel
    if (simt) {
      if (bf) sgemm_strided_kernel<BF16><<<total, 256, 0, s>>>(mp); else sgemm_strided_kernel<FP16><<<total, 256, 0, s>>>(mp);
    } else {
      if (bf) tgemm_strided_kernel<BF16><<<total, 128, 0, s>>>(mp); else tgemm_strided_kernel<FP16><<<total, 128, 0, s>>>(mp);
    }
    if (int e = pir_check_launch("pir_mdta_bwd(dA, fold, dwo part)")) return e;
  }
  if (c <= 256) mdta_bwd_rows_kernel<8><<<dim3((C + 7) / 8, B), 256, 0, s>>>(a);
  else mdta_bwd_rows_kernel<24><<<dim3((C + 7) / 8, B), 256, 0, s>>>(a);
  if (int e = pir_check_launch("pir_mdta_bwd(rows)")) return e;
  mdta_bwd_cols_kernel<<<(B * C + 7) / 8, 256, 0, s>>>(a);
  if (int e = pir_check_launch("pir_mdta_bwd(cols)")) return e;
  dim3 gw((unsigned)(((long long)2 * C * (2 * C / 8) + 255) / 256), B);
  if (bf) mdta_bwd_weights_kernel<BF16><<<gw, 256, 0, s>>>(a, reinterpret_cast<unsigned short*>(d->wqk), kpad2);
  else mdta_bwd_weights_kernel<FP16><<<gw, 256, 0, s>>>(a, reinterpret_cast<unsigned short*>(d->wqk), kpad2);
  if (int e = pir_check_launch("pir_mdta_bwd(weights)")) return e;
  mdta_bwd_small_kernel<<<1 + (d->dst_bias ? (C + 255) / 256 : 0) + (C * C + 255) / 256, 256, 0, s>>>(a);
  return pir_check_launch("pir_mdta_bwd(dwo, small)");
}

extern "C" int64_t pir_prompt_bwd_ws_floats(int32_t B, int32_t L, int32_t D, int32_t S) {
  if (B <= 0 || L <= 0 || D <= 0 || S <= 0) return 0;
  return (int64_t)B * ((int64_t)S * S * D + 2 * L);
}

extern "C" int pir_prompt_bwd(const PirPromptBwd* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_prompt_bwd: null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->D <= 0 || d->S <= 0) return pir_fail(PIR_ERR_ARG, "pir_prompt_bwd: empty problem");
  if (d->L < 1 || d->L > kMaxLTrain) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_prompt_bwd: 1 <= L <= 8");
  if ((d->D % 8) || (d->dup_pitch % 8) || (d->dup_bstride % 8) || ((uintptr_t)d->dup & 15) || ((uintptr_t)d->scratch & 15))
    return pir_fail(PIR_ERR_ARG, "pir_prompt_bwd: tensors not 16-byte aligned");
  if (!d->dup || !d->prompt || !d->weights || !d->pool_ws || !d->lin_w || !d->scratch || !d->demb || !d->dst_prompt || !d->dst_lin_w || !d->dst_lin_b)
    return pir_fail(PIR_ERR_ARG, "pir_prompt_bwd: missing pointers");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)d->S * d->S * d->D;
  float* dmix = d->scratch;
  float* dw = dmix + (size_t)d->B * n;
  float* dlog = dw + (size_t)d->B * d->L;
  const long long total = (long long)d->B * d->S * d->S * (d->D / 8);
  auto* dup = reinterpret_cast<const unsigned short*>(d->dup);
  PIR_BY_DTYPE(d->dtype, (prompt_bwd_dmix_kernel<BF16><<<grid_for(total), 256, 0, s>>>(dup, d->dup_pitch, d->dup_bstride, d->H, d->W, d->D, d->S, dmix, total, d->align_corners)),
               (prompt_bwd_dmix_kernel<FP16><<<grid_for(total), 256, 0, s>>>(dup, d->dup_pitch, d->dup_bstride, d->H, d->W, d->D, d->S, dmix, total, d->align_corners)));
  if (int e = pir_check_launch("pir_prompt_bwd(dmix)")) return e;
  prompt_bwd_dot_kernel<<<dim3(d->L, d->B), 256, 0, s>>>(dmix, d->prompt, n, d->L, dw);
  if (int e = pir_check_launch("pir_prompt_bwd(dot)")) return e;
  const int HW = d->H * d->W;
  int nchunks = (HW + 255) / 256;                       // pool_chunks() of pir_prompt_gen
  nchunks = nchunks > 64 ? 64 : (nchunks < 1 ? 1 : nchunks);
  prompt_bwd_small_kernel<<<1, 256, 0, s>>>(dw, d->weights, d->pool_ws, nchunks, d->B, d->L, d->C, HW, d->lin_w, d->inv_scale, dlog, d->demb,
                                            d->dst_lin_w, d->dst_lin_b);
  if (int e = pir_check_launch("pir_prompt_bwd(small)")) return e;
  prompt_bwd_param_kernel<<<(unsigned)(((long long)d->D * d->S * d->S + 255) / 256), 256, 0, s>>>(dmix, d->weights, d->B, d->L, d->D, d->S * d->S,
                                                                                                   d->inv_scale, d->dst_prompt);
  return pir_check_launch("pir_prompt_bwd(param)");
}
