// Depthwise 3x3 stencils (pad 1) on NHWC 16-bit tensors for sm_100a, shared-memory staged by TMA.
//
//   plain : out[c] = dw3x3(in)[c]                                  MDTA qkv_dwconv   (net/model.py:112,120)
//   gated : out[c] = gelu_erf(dw3x3(in)[c]) * dw3x3(in)[C + c]     GDFN dwconv+gate  (net/model.py:90,96-97)
//
// Persistent CTAs: blockIdx.y fixes a channel chunk (its nine taps stay in registers for the CTA's lifetime),
// blockIdx.x strides over the (image, tile row, tile column) list.  Tiles are TH x 32 pixels; a single 4-D TMA box
// {chunk, 34, TH+2, 1} brings a tile plus its halo into one of two shared-memory buffers while the previous
// tile is being computed (double buffering); the image border's zero padding is the TMA out-of-bounds fill.  Every
// thread owns one pixel column and one vector of channels (8 for the plain kernel = LDS.128, 4 per half for
// the gated kernel = LDS.64), walks down the TH+2 input rows and scatters each row into three rotating fp32
// accumulator rows, so every input value is read from shared memory once per horizontal tap and the nine
// taps are register-resident packed 16-bit pairs consumed directly by mixed-precision FMAs
// (fma.rn.f32.{bf16,f16} -> SASS FHFMA): 9 FMAs per output element and no unpack instructions.
// These kernels are HBM-bound: algorithmic bytes = (Cin + Cout) * 2 per pixel.
#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kDwTW = 32;          // tile width in pixels (one lane group per pixel column)

struct DwArgs {
  int H, W, C;                     // C = output channels
  int tiles_x, tiles_y, n_sp;      // spatial tiles per image row / column, total over the batch
  const void* w;                   // [9][Cin] 16-bit, tap-major
  const float* bias;               // [Cin] or null
  void* out;
  long long out_pitch, out_bstride;
  const void* dg;                  // gate backward only: gradient of the gated output [B,H,W,C]
  long long dg_pitch, dg_bstride;
};

// ------------------------------------------------------------------------------------------------------
// plain: CG channel groups of 8 channels per tile (CG = 8 -> 64 ch, CG = 6 -> 48 ch); block = CG * 32 threads
// ------------------------------------------------------------------------------------------------------
template <class T, int CG, int TH>
__global__ void __launch_bounds__(CG * 32)
dwconv_plain_kernel(const __grid_constant__ CUtensorMap tmIn, const DwArgs a) {
  constexpr int CC = CG * 8;
  constexpr int ROW_BYTES = CC * 2;
  constexpr int SW = kDwTW + 2;
  extern __shared__ uint8_t dw_smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  uint8_t* dw_smem = dw_smem_raw + ((128u - (smem_u32(dw_smem_raw) & 127u)) & 127u);   // TMA destination alignment

  const int tid = threadIdx.x;
  const int cg = tid % CG;
  const int tx = tid / CG;
  const int c0 = blockIdx.y * CC;
  constexpr uint32_t TILE_BYTES = (TH + 2) * SW * ROW_BYTES;
  const uint32_t bar_a[2] = {smem_u32(&bar[0]), smem_u32(&bar[1])};
  const int per_img = a.tiles_x * a.tiles_y;

  auto issue = [&](int sp, int buf) {      // thread 0 only
    const int b = sp / per_img, r = sp % per_img;
    mbar_expect_tx(bar_a[buf], TILE_BYTES);
    tma_load_4d(smem_u32(dw_smem) + buf * TILE_BYTES, &tmIn, bar_a[buf], c0, (r % a.tiles_x) * kDwTW - 1, (r / a.tiles_x) * TH - 1, b);
  };
  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(bar_a[0], 1);
    mbar_init(bar_a[1], 1);
    fence_barrier_init();
    pdl_wait();
    if ((int)blockIdx.x < a.n_sp) issue(blockIdx.x, 0);
  }
  // taps for this thread's 8 channels, kept as packed 16-bit pairs (9 x 4 registers)
  const int c = c0 + cg * 8;
  const bool c_ok = c < a.C;
  uint4 wt[9];
  float binit[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) binit[i] = 0.f;
  if (c_ok) {
    const unsigned short* wp = reinterpret_cast<const unsigned short*>(a.w) + c;
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)t * a.C));
    if (a.bias) {
#pragma unroll
      for (int i = 0; i < 8; ++i) binit[i] = __ldg(a.bias + c + i);
    }
  } else {
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();                 // barrier init visible to all waiters
  pdl_wait();

  int it = 0;
  for (int sp = blockIdx.x; sp < a.n_sp; sp += gridDim.x, ++it) {
    const int buf = it & 1;
    if (tid == 0 && sp + (int)gridDim.x < a.n_sp) issue(sp + gridDim.x, buf ^ 1);   // buffer freed by the barrier below
    const int b = sp / per_img, rr = sp % per_img;
    const int x0 = (rr % a.tiles_x) * kDwTW, y0 = (rr / a.tiles_x) * TH;
    mbar_wait(bar_a[buf], (it >> 1) & 1);

    const int x = x0 + tx;
    const bool x_ok = x < a.W;
    unsigned short* outp = reinterpret_cast<unsigned short*>(a.out) + (size_t)b * a.out_bstride + c;
    const uint8_t* col = dw_smem + (size_t)buf * TILE_BYTES + (size_t)tx * ROW_BYTES + cg * 16;

    float acc[3][8];
#pragma unroll
    for (int r = 0; r < TH + 2; ++r) {
      // input row r feeds output rows r (tap row 0), r-1 (tap row 1), r-2 (tap row 2)
      if (r < TH) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[r % 3][i] = binit[i];
      }
      uint4 v[3];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) v[kx] = *reinterpret_cast<const uint4*>(col + ((size_t)r * SW + kx) * ROW_BYTES);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int o = r - ky;
        if (o >= 0 && o < TH) {
          float* ac = acc[o % 3];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint4 w = wt[ky * 3 + kx];
            const uint32_t vv[4] = {v[kx].x, v[kx].y, v[kx].z, v[kx].w};
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              ac[2 * q] = fma16<T>(lo16(vv[q]), lo16(ww[q]), ac[2 * q]);
              ac[2 * q + 1] = fma16<T>(hi16(vv[q]), hi16(ww[q]), ac[2 * q + 1]);
            }
          }
        }
      }
      const int o = r - 2;            // this output row is now complete
      if (o >= 0) {
        const int y = y0 + o;
        if (c_ok && x_ok && y < a.H) {
          const float* ac = acc[o % 3];
          uint4 ov;
          ov.x = pack2<T>(ac[0], ac[1]); ov.y = pack2<T>(ac[2], ac[3]);
          ov.z = pack2<T>(ac[4], ac[5]); ov.w = pack2<T>(ac[6], ac[7]);
          *reinterpret_cast<uint4*>(outp + ((size_t)y * a.W + x) * a.out_pitch) = ov;
        }
      }
    }
    __syncthreads();                 // everyone is done with this buffer before it is refilled
  }
}

// ------------------------------------------------------------------------------------------------------
// gated: 32 output channels per tile (two 32-channel boxes: x1 at c0, x2 at C + c0); 4 channels per thread
// ------------------------------------------------------------------------------------------------------
// BWD = true: the same stencil recomputes y1, y2 (fp32) and the epilogue is the gate's backward instead:
//   out[c] = dg * y2 * (Phi(y1) + y1 phi(y1)),  out[C + c] = dg * y1 * Phi(y1)     (exact erf GELU; out has 2C channels)
template <class T, int TH, bool BWD>
__global__ void __launch_bounds__(256)
dwconv_gate_kernel(const __grid_constant__ CUtensorMap tmIn, const DwArgs a) {
  constexpr int CC = 32;
  constexpr int ROW_BYTES = CC * 2;
  constexpr int SW = kDwTW + 2;
  constexpr int HALF_BYTES = (TH + 2) * SW * ROW_BYTES;
  extern __shared__ uint8_t dw_smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  uint8_t* dw_smem = dw_smem_raw + ((128u - (smem_u32(dw_smem_raw) & 127u)) & 127u);   // TMA destination alignment

  const int tid = threadIdx.x;
  const int cg = tid & 7;
  const int tx = tid >> 3;
  const int c0 = blockIdx.y * CC;
  constexpr uint32_t TILE_BYTES = 2 * HALF_BYTES;
  const uint32_t bar_a[2] = {smem_u32(&bar[0]), smem_u32(&bar[1])};
  const int per_img = a.tiles_x * a.tiles_y;

  auto issue = [&](int sp, int buf) {      // thread 0 only
    const int b = sp / per_img, r = sp % per_img;
    const int xx = (r % a.tiles_x) * kDwTW - 1, yy = (r / a.tiles_x) * TH - 1;
    const uint32_t dst = smem_u32(dw_smem) + buf * TILE_BYTES;
    mbar_expect_tx(bar_a[buf], TILE_BYTES);
    tma_load_4d(dst, &tmIn, bar_a[buf], c0, xx, yy, b);
    tma_load_4d(dst + HALF_BYTES, &tmIn, bar_a[buf], a.C + c0, xx, yy, b);
  };
  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(bar_a[0], 1);
    mbar_init(bar_a[1], 1);
    fence_barrier_init();
    pdl_wait();
    if ((int)blockIdx.x < a.n_sp) issue(blockIdx.x, 0);
  }
  const int c = c0 + cg * 4;
  const bool c_ok = c < a.C;
  uint2 w1[9], w2[9];
  float b1[4] = {0.f, 0.f, 0.f, 0.f}, b2[4] = {0.f, 0.f, 0.f, 0.f};
  if (c_ok) {
    const unsigned short* wp = reinterpret_cast<const unsigned short*>(a.w) + c;
    const size_t cin = (size_t)2 * a.C;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      w1[t] = __ldg(reinterpret_cast<const uint2*>(wp + (size_t)t * cin));
      w2[t] = __ldg(reinterpret_cast<const uint2*>(wp + (size_t)t * cin + a.C));
    }
    if (a.bias) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { b1[i] = __ldg(a.bias + c + i); b2[i] = __ldg(a.bias + a.C + c + i); }
    }
  } else {
#pragma unroll
    for (int t = 0; t < 9; ++t) { w1[t] = make_uint2(0, 0); w2[t] = make_uint2(0, 0); }
  }
  __syncthreads();
  pdl_wait();

  int it = 0;
  for (int sp = blockIdx.x; sp < a.n_sp; sp += gridDim.x, ++it) {
    const int buf = it & 1;
    if (tid == 0 && sp + (int)gridDim.x < a.n_sp) issue(sp + gridDim.x, buf ^ 1);
    const int b = sp / per_img, rr = sp % per_img;
    const int x0 = (rr % a.tiles_x) * kDwTW, y0 = (rr / a.tiles_x) * TH;
    mbar_wait(bar_a[buf], (it >> 1) & 1);

    const int x = x0 + tx;
    const bool x_ok = x < a.W;
    unsigned short* outp = reinterpret_cast<unsigned short*>(a.out) + (size_t)b * a.out_bstride + c;
    const uint8_t* col = dw_smem + (size_t)buf * TILE_BYTES + (size_t)tx * ROW_BYTES + cg * 8;

    // gate backward: this thread's dg values of the whole tile are fetched up front, so their latency hides behind the stencil
    uint2 dgv[BWD ? TH : 1];
    if (BWD) {
#pragma unroll
      for (int o = 0; o < TH; ++o) {
        const int y = y0 + o;
        dgv[o] = (c_ok && x_ok && y < a.H)
                     ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(a.dg) + (size_t)b * a.dg_bstride +
                                                            ((size_t)y * a.W + x) * a.dg_pitch + c))
                     : make_uint2(0u, 0u);
      }
    }
    float p[3][4], q[3][4];
#pragma unroll
    for (int r = 0; r < TH + 2; ++r) {
      if (r < TH) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { p[r % 3][i] = b1[i]; q[r % 3][i] = b2[i]; }
      }
      uint2 v1[3], v2[3];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint8_t* s = col + ((size_t)r * SW + kx) * ROW_BYTES;
        v1[kx] = *reinterpret_cast<const uint2*>(s);
        v2[kx] = *reinterpret_cast<const uint2*>(s + HALF_BYTES);
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int o = r - ky;
        if (o >= 0 && o < TH) {
          float* pp = p[o % 3];
          float* qq = q[o % 3];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint2 wa = w1[ky * 3 + kx], wb = w2[ky * 3 + kx];
            pp[0] = fma16<T>(lo16(v1[kx].x), lo16(wa.x), pp[0]);
            pp[1] = fma16<T>(hi16(v1[kx].x), hi16(wa.x), pp[1]);
            pp[2] = fma16<T>(lo16(v1[kx].y), lo16(wa.y), pp[2]);
            pp[3] = fma16<T>(hi16(v1[kx].y), hi16(wa.y), pp[3]);
            qq[0] = fma16<T>(lo16(v2[kx].x), lo16(wb.x), qq[0]);
            qq[1] = fma16<T>(hi16(v2[kx].x), hi16(wb.x), qq[1]);
            qq[2] = fma16<T>(lo16(v2[kx].y), lo16(wb.y), qq[2]);
            qq[3] = fma16<T>(hi16(v2[kx].y), hi16(wb.y), qq[3]);
          }
        }
      }
      const int o = r - 2;
      if (o >= 0) {
        const int y = y0 + o;
        if (c_ok && x_ok && y < a.H) {
          const float* pp = p[o % 3];
          const float* qq = q[o % 3];
          if (!BWD) {
            uint2 ov;
            ov.x = pack2<T>(gelu_erf(pp[0]) * qq[0], gelu_erf(pp[1]) * qq[1]);
            ov.y = pack2<T>(gelu_erf(pp[2]) * qq[2], gelu_erf(pp[3]) * qq[3]);
            *reinterpret_cast<uint2*>(outp + ((size_t)y * a.W + x) * a.out_pitch) = ov;
          } else {
            const uint2 dv = dgv[BWD ? o : 0];
            const float d[4] = {unpack_lo<T>(dv.x), unpack_hi<T>(dv.x), unpack_lo<T>(dv.y), unpack_hi<T>(dv.y)};
            float o1[4], o2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              // Phi from the same logistic-polynomial fit the forward gate uses (gelu = x * Phi: |x dPhi| <= 2.6e-5), phi from ex2.approx
              const float u = fminf(pp[i] * pp[i], 25.0f);
              const float cdf = rcp_approx(1.0f + ex2_approx(pp[i] * fmaf(u, fmaf(u, kGeluC, kGeluB), kGeluA)));
              const float pdf = ex2_approx(-0.7213475204444817f * pp[i] * pp[i]) * 0.3989422804014327f;
              o1[i] = d[i] * qq[i] * fmaf(pp[i], pdf, cdf);
              o2[i] = d[i] * pp[i] * cdf;
            }
            unsigned short* op = outp + ((size_t)y * a.W + x) * a.out_pitch;
            *reinterpret_cast<uint2*>(op) = make_uint2(pack2<T>(o1[0], o1[1]), pack2<T>(o1[2], o1[3]));
            *reinterpret_cast<uint2*>(op + a.C) = make_uint2(pack2<T>(o2[0], o2[1]), pack2<T>(o2[2], o2[3]));
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------
template <class T>
static int launch_dwconv(const PirDwConv* d, cudaStream_t stream) {
  constexpr int TH = 8;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0) return pir_fail(PIR_ERR_ARG, "pir_dwconv3x3: empty problem");
  const int cin = d->gate ? 2 * d->C : d->C;
  const int cmul = d->gate ? 4 : 8;
  if ((d->C % cmul) || (cin % 8) || (d->in_pitch % 8) || (d->in_bstride % 8) || (d->out_pitch % cmul) ||
      ((uintptr_t)d->in & 15) || ((uintptr_t)d->out & 15) || ((uintptr_t)d->w & 15))
    return pir_fail(PIR_ERR_ARG, "pir_dwconv3x3: channel counts / pitches / pointers are not vector aligned");
  if (d->gate && (d->C % 8)) return pir_fail(PIR_ERR_ARG, "pir_dwconv3x3: gated C must be a multiple of 8");

  DwArgs a{};
  a.H = d->H; a.W = d->W; a.C = d->C;
  a.tiles_x = (d->W + kDwTW - 1) / kDwTW;
  a.tiles_y = (d->H + TH - 1) / TH;
  a.n_sp = a.tiles_x * a.tiles_y * d->B;
  auto workers_for = [&](int chunks, int occ) {   // `occ` resident CTAs per SM, split evenly over the channel chunks
    int w = (occ * 148 + chunks - 1) / chunks;
    if (w > a.n_sp) w = a.n_sp;
    return w < 1 ? 1 : w;
  };
  a.w = d->w; a.bias = d->bias; a.out = d->out; a.out_pitch = d->out_pitch; a.out_bstride = d->out_bstride;
  a.dg = d->dg; a.dg_pitch = d->dg_pitch; a.dg_bstride = d->dg_bstride;
  if (d->gate == 2 && (!d->dg || (d->dg_pitch % 4) || (d->dg_bstride % 4) || ((uintptr_t)d->dg & 7)))
    return pir_fail(PIR_ERR_ARG, "pir_dwconv3x3: gate backward needs dg (8-byte aligned)");

  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
  const uint64_t strides[3] = {(uint64_t)d->in_pitch * 2, (uint64_t)d->in_pitch * 2 * d->W, (uint64_t)d->in_bstride * 2};
  CUtensorMap tm;
  if (d->gate) {
    const uint32_t box[4] = {32, kDwTW + 2, TH + 2, 1};
    if (int e = pir_make_tmap(&tm, dt, 4, d->in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    const size_t smem = (size_t)2 * 2 * (TH + 2) * (kDwTW + 2) * 64 + 128;
    const int bw = d->gate == 2 ? 1 : 0;
    if (!pir_smem_attr_once(bw ? reinterpret_cast<const void*>(dwconv_gate_kernel<T, TH, true>) : reinterpret_cast<const void*>(dwconv_gate_kernel<T, TH, false>),
                            (int)smem, "pir_dwconv3x3")) return PIR_ERR_CUDA;
    const int chunks = (d->C + 31) / 32;
    dim3 grid((unsigned)workers_for(chunks, 2), (unsigned)chunks, 1);
    if (bw) pir_launch(dwconv_gate_kernel<T, TH, true>, grid, dim3(256), smem, stream, tm, a);
    else pir_launch(dwconv_gate_kernel<T, TH, false>, grid, dim3(256), smem, stream, tm, a);
  } else {
    const bool use48 = (d->C % 64 != 0) && (d->C % 48 == 0);
    const int cc = use48 ? 48 : 64;
    const uint32_t box[4] = {(uint32_t)cc, kDwTW + 2, TH + 2, 1};
    if (int e = pir_make_tmap(&tm, dt, 4, d->in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    const size_t smem = (size_t)2 * (TH + 2) * (kDwTW + 2) * cc * 2 + 128;
    const int chunks = (d->C + cc - 1) / cc;
    dim3 grid((unsigned)workers_for(chunks, use48 ? 3 : 2), (unsigned)chunks, 1);
    if (use48) {
      if (!pir_smem_attr_once(reinterpret_cast<const void*>(dwconv_plain_kernel<T, 6, TH>), (int)smem, "pir_dwconv3x3")) return PIR_ERR_CUDA;
      pir_launch(dwconv_plain_kernel<T, 6, TH>, grid, dim3(6 * 32), smem, stream, tm, a);
    } else {
      if (!pir_smem_attr_once(reinterpret_cast<const void*>(dwconv_plain_kernel<T, 8, TH>), (int)smem, "pir_dwconv3x3")) return PIR_ERR_CUDA;
      pir_launch(dwconv_plain_kernel<T, 8, TH>, grid, dim3(8 * 32), smem, stream, tm, a);
    }
  }
  return pir_check_launch("pir_dwconv3x3");
}

}  // namespace pir

extern "C" int pir_dwconv3x3(const PirDwConv* d, void* stream) {
  if (!d) return pir_fail(PIR_ERR_ARG, "pir_dwconv3x3: null descriptor");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return d->dtype == PIR_DTYPE_BF16 ? pir::launch_dwconv<pir::BF16>(d, s) : pir::launch_dwconv<pir::FP16>(d, s);
}
