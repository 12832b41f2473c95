// On-device evaluation I/O around the forward: the reference's test loop pads on the device but scores every image
// on the host through skimage (test.py:100-114, utils/val_utils.py:50-66) and synthesises noise in numpy (utils/dataset_utils.py:195-198,
// utils/degradation_utils.py:21-26).  Once the forward takes milliseconds those host steps dominate, so they are kernels here:
//   pir_mirror_pad   : torch.cat([x, flip(x)])[:Hp] on both axes                                   (test.py:100-105)
//   pir_psnr_ssim    : skimage.metrics.peak_signal_noise_ratio / structural_similarity (7x7 uniform window, sample covariance,
//                      K1 = 0.01, K2 = 0.03, data_range = 1, border of 3 cropped, mean over channels) of clip(x, 0, 1) images
//   pir_add_noise    : clip(clean + sigma * N(0,1), 0, 255).astype(uint8) / 255 with a counter-based Philox stream
// All three are HBM-bound element-wise / small-stencil kernels: coalesced fp32 rows, shared-memory separable box sums, fp64 only for
// the final per-image accumulation (deterministic two-stage reduction).
#include <curand_kernel.h>

#include "common.cuh"
#include "host.h"

namespace pir {

__global__ void __launch_bounds__(256)
mirror_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int H, int W, int Hp, int Wp) {
  const long long total = (long long)planes * Hp * Wp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % Wp);
    const long long r = e / Wp;
    const int y = (int)(r % Hp);
    const long long p = r / Hp;
    const int sy = y < H ? y : 2 * H - 1 - y, sx = x < W ? x : 2 * W - 1 - x;
    out[e] = __ldg(in + (p * H + sy) * W + sx);
  }
}

constexpr int kSsT = 32, kSsR = 3, kSsIn = kSsT + 2 * kSsR;      // 32 x 32 window centres per CTA, 38 x 38 input pixels

// grid (tiles_x * tiles_y, C, B); partial[(b*C + c)*tiles + tile] = {sum of S over the valid centres, sum of squared error of the block}
__global__ void __launch_bounds__(256)
psnr_ssim_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int tiles_x, double2* __restrict__ partial) {
  __shared__ float sA[kSsIn][kSsIn + 1], sB[kSsIn][kSsIn + 1];
  __shared__ float sH[5][kSsIn][kSsT + 1];
  __shared__ double red[2][8];
  const int tile = blockIdx.x, c = blockIdx.y, img = blockIdx.z, C = gridDim.y;
  const int y0 = (tile / tiles_x) * kSsT, x0 = (tile % tiles_x) * kSsT;
  const float* pa = a + ((size_t)img * C + c) * H * W;
  const float* pb = b + ((size_t)img * C + c) * H * W;
  double se = 0.0, ss = 0.0;
  for (int e = threadIdx.x; e < kSsIn * kSsIn; e += 256) {
    const int r = e / kSsIn, q = e % kSsIn;
    const int y = y0 - kSsR + r, x = x0 - kSsR + q;
    float va = 0.f, vb = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      va = fminf(fmaxf(__ldg(pa + (size_t)y * W + x), 0.f), 1.f);         // np.clip(., 0, 1)   val_utils.py:52-53
      vb = fminf(fmaxf(__ldg(pb + (size_t)y * W + x), 0.f), 1.f);
      if (r >= kSsR && r < kSsR + kSsT && q >= kSsR && q < kSsR + kSsT) { const double d = (double)va - (double)vb; se += d * d; }
    }
    sA[r][q] = va;
    sB[r][q] = vb;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kSsIn * kSsT; e += 256) {               // horizontal 7-sums of x, y, xx, yy, xy
    const int r = e / kSsT, q = e % kSsT;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
    for (int d = 0; d < 7; ++d) {
      const float u = sA[r][q + d], v = sB[r][q + d];
      s0 += u; s1 += v; s2 = fmaf(u, u, s2); s3 = fmaf(v, v, s3); s4 = fmaf(u, v, s4);
    }
    sH[0][r][q] = s0; sH[1][r][q] = s1; sH[2][r][q] = s2; sH[3][r][q] = s3; sH[4][r][q] = s4;
  }
  __syncthreads();
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f, inv = 1.0f / 49.0f, covn = 49.0f / 48.0f;
  for (int e = threadIdx.x; e < kSsT * kSsT; e += 256) {
    const int i = e / kSsT, j = e % kSsT;
    const int cy = y0 + i, cx = x0 + j;
    if (cy < kSsR || cy >= H - kSsR || cx < kSsR || cx >= W - kSsR) continue;      // crop(S, 3)
    float s[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float t = 0.f;
#pragma unroll
      for (int d = 0; d < 7; ++d) t += sH[q][i + d][j];
      s[q] = t * inv;
    }
    const float ux = s[0], uy = s[1];
    const float vx = covn * (s[2] - ux * ux), vy = covn * (s[3] - uy * uy), vxy = covn * (s[4] - ux * uy);
    const float S = ((2.f * ux * uy + C1) * (2.f * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    ss += (double)S;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); se += __shfl_xor_sync(0xffffffffu, se, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ss; red[1][threadIdx.x >> 5] = se; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int w = 0; w < 8; ++w) { s0 += red[0][w]; s1 += red[1][w]; }
    partial[((size_t)img * C + c) * gridDim.x + tile] = make_double2(s0, s1);
  }
}

// one block per image: out[b] = {psnr, ssim}
__global__ void __launch_bounds__(256)
psnr_ssim_final_kernel(const double2* __restrict__ partial, int C, int tiles, int H, int W, double* __restrict__ out) {
  __shared__ double red[2][256];
  const int b = blockIdx.x;
  double ss = 0.0, se = 0.0;
  for (int e = threadIdx.x; e < C * tiles; e += 256) { const double2 v = partial[(size_t)b * C * tiles + e]; ss += v.x; se += v.y; }
  red[0][threadIdx.x] = ss; red[1][threadIdx.x] = se;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) { red[0][threadIdx.x] += red[0][threadIdx.x + o]; red[1][threadIdx.x] += red[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mse = red[1][0] / ((double)C * H * W);
    out[2 * b] = 10.0 * log10(1.0 / mse);
    out[2 * b + 1] = red[0][0] / ((double)C * (double)(H - 2 * kSsR) * (double)(W - 2 * kSsR));
  }
}

__global__ void __launch_bounds__(256)
add_noise_kernel(const float* __restrict__ clean255, float* __restrict__ out, long long n, float sigma, unsigned long long seed) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = 4 consecutive elements = one Philox call
  if (q * 4 >= n) return;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)q, 0, &st);
  const float4 z = curand_normal4(&st);
  const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long e = q * 4 + i;
    if (e < n) out[e] = floorf(fminf(fmaxf(__ldg(clean255 + e) + zz[i] * sigma, 0.f), 255.f)) * (1.0f / 255.0f);   // .astype(uint8) / 255
  }
}

}  // namespace pir

extern "C" int pir_mirror_pad(const float* in, float* out, int32_t planes, int32_t H, int32_t W, int32_t Hp, int32_t Wp, void* stream) {
  if (!in || !out || planes <= 0 || H <= 0 || W <= 0) return pir_fail(PIR_ERR_ARG, "pir_mirror_pad: bad arguments");
  if (Hp < H || Wp < W || Hp > 2 * H || Wp > 2 * W) return pir_fail(PIR_ERR_ARG, "pir_mirror_pad: padded size must be in [size, 2*size]");
  const long long total = (long long)planes * Hp * Wp;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pir::mirror_pad_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, planes, H, W, Hp, Wp);
  return pir_check_launch("pir_mirror_pad");
}

extern "C" int64_t pir_psnr_ssim_ws_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)B * C * ((H + pir::kSsT - 1) / pir::kSsT) * ((W + pir::kSsT - 1) / pir::kSsT) * 16;
}

extern "C" int pir_psnr_ssim(const float* restored, const float* clean, int32_t B, int32_t C, int32_t H, int32_t W, void* ws, double* out,
                             void* stream) {
  if (!restored || !clean || !ws || !out || B <= 0 || C <= 0) return pir_fail(PIR_ERR_ARG, "pir_psnr_ssim: bad arguments");
  if (H < 7 || W < 7) return pir_fail(PIR_ERR_ARG, "pir_psnr_ssim: images must be at least 7 x 7 (the SSIM window), as skimage requires");
  if (B > 65535 || C > 65535) return pir_fail(PIR_ERR_UNSUPPORTED, "pir_psnr_ssim: grid too large");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int tx = (W + pir::kSsT - 1) / pir::kSsT, ty = (H + pir::kSsT - 1) / pir::kSsT;
  pir::psnr_ssim_partial_kernel<<<dim3(tx * ty, C, B), 256, 0, s>>>(restored, clean, H, W, tx, reinterpret_cast<double2*>(ws));
  if (int e = pir_check_launch("pir_psnr_ssim(partial)")) return e;
  pir::psnr_ssim_final_kernel<<<B, 256, 0, s>>>(reinterpret_cast<const double2*>(ws), C, tx * ty, H, W, out);
  return pir_check_launch("pir_psnr_ssim(final)");
}

extern "C" int pir_add_noise(const float* clean255, float* out, int64_t n, float sigma, uint64_t seed, void* stream) {
  if (!clean255 || !out || n <= 0) return pir_fail(PIR_ERR_ARG, "pir_add_noise: bad arguments");
  const long long quads = (n + 3) / 4;
  pir::add_noise_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(clean255, out, n, sigma, seed);
  return pir_check_launch("pir_add_noise");
}
