// Dense 3x3 convolution (pad 1) as nine tcgen05 GEMMs over ONE halo'd shared-memory tile  (nn.Conv2d(k=3): net/model.py:164, 174, 206,
// 223, 320; with PixelUnshuffle / PixelShuffle / "+ inp_img" folded into the store: :165, :175, :377).
//
// gemm_tcgen05.cu fetches every tap as its own shifted TMA box: the activation crosses L2 -> shared memory nine times and every
// 128-pixel tile is a CTA of its own (1.3 ms of the cfg2 step for convolutions whose HBM time is 0.15 ms).  Here
//   * a persistent CTA walks (image, TH x 16-pixel tile) items; TMA brings the tile + 1-pixel halo ONCE ((TH + 2) x 18 pixels, all input
//     channels, 128-byte swizzled k-blocks, out-of-image pixels = zero fill = the conv's padding), double buffered;
//   * the accumulator rows are the RASTER positions of the halo'd tile (row m = y * 18 + x): for tap (dy, dx) the A operand is the same
//     tile read from row offset dy * 18 + dx.  A K-major SWIZZLE_128B descriptor may start at any 128-byte row: the swizzle is a
//     function of the absolute shared-memory address, which TMA used when it wrote the tile (tools/probes/umma_shift_probe.cu checks
//     this on the device).  Rows whose x >= 16 are halo columns: computed and discarded (2 of 18);
//   * the weights of the CTA's output-channel slice (all nine taps) stay resident in shared memory; wide convolutions are split into
//     channel slices across CTAs (the activation tile is then read once per slice, from L2);
//   * fp32 accumulators in TMEM, double buffered: the epilogue of item i overlaps the MMAs of item i + 1.  The K loop of an item (nine
//     taps x k-steps of 16 channels) is dealt round-robin onto FOUR independent accumulators per item (two row groups x two partial
//     sums, or one row group x four) that the epilogue adds: back-to-back MMAs into one accumulator serialise at ~100 cycles each
//     for these narrow N, and the issuing thread's loop is fully unrolled with compile-time descriptor offsets for the same reason.
// Warp roles (192 threads): 0 TMA producer | 1 MMA issuer | 2-5 epilogue (TMEM lane = raster position = one pixel per thread).
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace pir {

constexpr int kC3Threads = 192;
constexpr int kC3TW = 16, kC3SW = 18;

struct C3Args {
  int B, H, W, K, N;       // K input channels, N output channels
  int nkb;                 // k-blocks of 64 input channels
  int ns, slice_n;         // output-channel slices across CTAs, padded channels per slice (multiple of 16, <= 64)
  int th;                  // output rows per item (7 * MT)
  int tiles_x, tiles_y, n_items, groups;     // groups = CTAs per slice
  uint32_t mg_per_img, mg_tiles_x;
  uint32_t x_rows;         // shared-memory rows per x k-block (covers the largest shifted window)
  int nbuf;                // x-tile ring depth (2..4): loads run nbuf - 1 items ahead of the MMAs
  uint32_t off_w;          // byte offset of the resident weights
  int out_mode;
  void* out;
  long long out_pitch, out_bstride;
  const void* res;
  long long res_pitch, res_bstride;
  const float* vec_t;      // [N] bias or null
  const float* img;        // fp32 NCHW input image (PIR_OUT_FINAL_NCHW32)
};

__device__ __forceinline__ uint32_t c3_div(uint32_t n, uint32_t magic) { return magic ? __umulhi(n, magic) : n; }

// KSL: 16-channel k-steps of the LAST k-block (earlier k-blocks have 4): with MT and NKB it makes the issuing thread's loop straight-line
template <class T, int MT, int NKB, int KSL>
__global__ void __launch_bounds__(kC3Threads, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const C3Args g) {
  constexpr int TW = kC3TW, SW = kC3SW, TH = 7 * MT;
  constexpr int NP = 4 / MT;                                    // partial accumulators per row group
  constexpr uint32_t X_ROWS = (uint32_t)((2 * kC3SW + 2 + MT * 128 + 7) / 8 * 8);
  constexpr uint32_t X_KB = X_ROWS * 128u, X_BYTES = NKB * X_KB;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[4], bar_aempty[4], bar_wfull, bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t x_kb_bytes = X_KB, x_bytes = X_BYTES;
  const uint32_t w_tile = (uint32_t)g.slice_n * 128u;            // one (tap, k-block) tile of the weight slice
  const int slice = blockIdx.x % g.ns, gidx = blockIdx.x / g.ns;
  const int n0 = slice * g.slice_n;
  const int per_img = g.tiles_x * g.tiles_y;
  auto item_geo = [&](int item, int& b, int& x0, int& y0) {
    b = (int)c3_div((uint32_t)item, g.mg_per_img);
    const int rr = item - b * per_img;
    const int ty = (int)c3_div((uint32_t)rr, g.mg_tiles_x);
    x0 = (rr - ty * g.tiles_x) * TW; y0 = ty * TH;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW);
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&bar_afull[i]), 1); mbar_init(smem_u32(&bar_aempty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), 4); }
    mbar_init(smem_u32(&bar_wfull), 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ========================================= TMA producer ==========================================
    if (elect_one()) {                               // the weight slice: 9 taps x nkb k-blocks x slice_n rows, once
      const uint32_t full = smem_u32(&bar_wfull);
      mbar_expect_tx(full, 9u * (uint32_t)g.nkb * w_tile);
      for (int tap = 0; tap < 9; ++tap)
        for (int kb = 0; kb < g.nkb; ++kb)
          tma_load_2d(base + g.off_w + (uint32_t)(tap * g.nkb + kb) * w_tile, &tmW, full, (tap * g.nkb + kb) * 64, n0);
    }
    __syncwarp();
    uint32_t it = 0;
    for (int item = gidx; item < g.n_items; item += g.groups, ++it) {
      int b, x0, y0;
      item_geo(item, b, x0, y0);
      const uint32_t ab = it % (uint32_t)g.nbuf;
      mbar_wait_sleep(smem_u32(&bar_aempty[ab]), ((it / (uint32_t)g.nbuf) & 1u) ^ 1u);
      if (elect_one()) {
        const uint32_t full = smem_u32(&bar_afull[ab]);
        mbar_expect_tx(full, (uint32_t)g.nkb * (uint32_t)((TH + 2) * SW * 128));
        for (int kb = 0; kb < g.nkb; ++kb) tma_load_4d(base + ab * x_bytes + (uint32_t)kb * x_kb_bytes, &tmA, full, kb * 64, x0 - 1, y0 - 1, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ========================================= MMA issuer ============================================
    const uint64_t desc_hi = make_sdesc_sw128(0, 16, 1024);
    const uint32_t idesc = make_idesc_f16(T::kFmt, 128, g.slice_n, 0, 0);
    mbar_wait_sleep(smem_u32(&bar_wfull), 0);
    uint32_t it = 0;
    for (int item = gidx; item < g.n_items; item += g.groups, ++it) {
      const uint32_t ab = it % (uint32_t)g.nbuf, tb = it & 1u;
      mbar_wait(smem_u32(&bar_afull[ab]), (it / (uint32_t)g.nbuf) & 1u);
      mbar_wait(smem_u32(&bar_tempty[tb]), ((it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t xd0 = desc_hi | (uint64_t)(((base + ab * X_BYTES) >> 4) & 0x3fffu);
        const uint64_t wd0 = desc_hi | (uint64_t)(((base + g.off_w) >> 4) & 0x3fffu);
        const uint32_t d0 = tmem_base + tb * 256u;
        const uint32_t w_step = w_tile >> 4;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t shift = (uint32_t)((tap / 3) * SW + tap % 3);        // raster offset of the tap's window
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            const uint64_t wd = wd0 + (uint64_t)((uint32_t)(tap * NKB + kb) * w_step);
#pragma unroll
            for (int k = 0; k < (kb == NKB - 1 ? KSL : 4); ++k) {
              const int p = (tap + k) % NP;           // partial sum this k-step goes to
              // first k-step a partial receives = the smallest (tap, kb, k) with (tap + k) % NP == p: k-steps of tap 0, then taps < NP
              const bool first = kb == 0 && ((tap == 0 && k < NP) || (k == 0 && tap < NP && tap >= (NKB == 1 ? KSL : 4)));
#pragma unroll
              for (int t = 0; t < MT; ++t) {          // row groups innermost: consecutive MMAs hit different accumulators
                const uint32_t a_off = ((uint32_t)kb * X_KB + (shift + (uint32_t)t * 128u) * 128u) >> 4;
                umma_f16(d0 + (uint32_t)((t * NP + p) * 64), xd0 + (uint64_t)(a_off + 2 * k), wd + (uint64_t)(2 * k), idesc, first ? 0u : 1u);
              }
            }
          }
        }
        // every partial receives at least one k-step: taps 0 .. NP-1 with k = 0 cover p = 0 .. NP-1
        umma_commit(smem_u32(&bar_aempty[ab]));       // x tile free when the MMAs retire
        umma_commit(smem_u32(&bar_tfull[tb]));
      }
      __syncwarp();
    }
  } else {
    // ========================================= epilogue ==============================================
    const int quarter = warp & 3;
    uint32_t it = 0;
    for (int item = gidx; item < g.n_items; item += g.groups, ++it) {
      int b, x0, y0;
      item_geo(item, b, x0, y0);
      const uint32_t ab = it & 1u;
      mbar_wait(smem_u32(&bar_tfull[ab]), (it >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const int m = t * 128 + quarter * 32 + lane;
        const int ry = m / SW, rx = m - ry * SW;
        const int py = y0 + ry, px = x0 + rx;
        const bool valid = rx < TW && ry < TH && py < g.H && px < g.W;
        const int pix = py * g.W + px;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * 256u + (uint32_t)(t * NP * 64);
        for (int c0 = 0; c0 < g.slice_n; c0 += 16) {
          const int n = n0 + c0;
          if (n >= g.N) break;                       // warp-uniform
          float v[16];
          {
            uint32_t acc[NP][16];
#pragma unroll
            for (int p = 0; p < NP; ++p) tmem_ld16(taddr + (uint32_t)(p * 64 + c0), acc[p]);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float sum = __uint_as_float(acc[0][i]);
#pragma unroll
              for (int p = 1; p < NP; ++p) sum += __uint_as_float(acc[p][i]);
              v[i] = sum;
            }
          }
          if (g.vec_t) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += (n + i < g.N) ? __ldg(g.vec_t + n + i) : 0.f;
          }
          if (g.out_mode == PIR_OUT_UNSHUFFLE16) {
            // PixelUnshuffle(2): out[b, y/2, x/2, 4n + 2(y&1) + (x&1)] = conv[b, y, x, n].  Raster neighbours (m, m ^ 1) are the two x
            // parities of one output pixel (SW and the block bases are even): they swap halves of their 16 channels, so each lane
            // stores 4-byte pairs for 8 channels instead of 2-byte scalars for 16
            const bool odd = (lane & 1) != 0;
            float mine[8], theirs[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float send = odd ? v[i] : v[8 + i];                 // even lanes keep channels 0-7, odd lanes 8-15
              theirs[i] = __shfl_xor_sync(0xffffffffu, send, 1);
              mine[i] = odd ? v[8 + i] : v[i];
            }
            if (valid) {
              const int nb = n + (odd ? 8 : 0);
              unsigned short* o = reinterpret_cast<unsigned short*>(g.out) + (size_t)b * g.out_bstride +
                                  ((size_t)(py >> 1) * (g.W >> 1) + (px >> 1)) * g.out_pitch + (py & 1) * 2;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (nb + i < g.N)
                  *reinterpret_cast<uint32_t*>(o + (size_t)(nb + i) * 4) = odd ? pack2<T>(theirs[i], mine[i]) : pack2<T>(mine[i], theirs[i]);
            }
          } else if (!valid) {
            // halo column / outside the image: nothing to store (the warp stays converged for the next tcgen05.ld)
          } else if (g.out_mode == PIR_OUT_NHWC16) {
            unsigned short* o = reinterpret_cast<unsigned short*>(g.out) + (size_t)b * g.out_bstride + (size_t)pix * g.out_pitch + n;
            const unsigned short* r = g.res ? reinterpret_cast<const unsigned short*>(g.res) + (size_t)b * g.res_bstride + (size_t)pix * g.res_pitch + n : nullptr;
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              if (n + h8 * 8 < g.N) {
                float* vv = v + h8 * 8;
                if (r) {
                  const uint4 rv = *reinterpret_cast<const uint4*>(r + h8 * 8);
                  const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) { vv[2 * q] += unpack_lo<T>(rw[q]); vv[2 * q + 1] += unpack_hi<T>(rw[q]); }
                }
                uint4 ov;
                ov.x = pack2<T>(vv[0], vv[1]); ov.y = pack2<T>(vv[2], vv[3]);
                ov.z = pack2<T>(vv[4], vv[5]); ov.w = pack2<T>(vv[6], vv[7]);
                *reinterpret_cast<uint4*>(o + h8 * 8) = ov;
              }
            }
          } else if (g.out_mode == PIR_OUT_SHUFFLE16) {
            // PixelShuffle(2): out[b, 2y+i, 2x+j, c] = conv[b, y, x, 4c + 2i + j]: the 16 columns of this thread are four consecutive
            // output channels for each of the four sub-pixels -> one 8-byte store per sub-pixel
            unsigned short* o = reinterpret_cast<unsigned short*>(g.out) + (size_t)b * g.out_bstride;
            const size_t W2 = (size_t)g.W * 2;
            if (n + 16 <= g.N) {
#pragma unroll
              for (int sub = 0; sub < 4; ++sub) {
                const size_t dp = ((size_t)(2 * py + (sub >> 1))) * W2 + (size_t)(2 * px + (sub & 1));
                uint2 ov;
                ov.x = pack2<T>(v[sub], v[4 + sub]);
                ov.y = pack2<T>(v[8 + sub], v[12 + sub]);
                *reinterpret_cast<uint2*>(o + dp * g.out_pitch + (n >> 2)) = ov;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int nn = n + i;
                if (nn < g.N) {
                  const size_t dp = ((size_t)(2 * py + ((nn >> 1) & 1))) * W2 + (size_t)(2 * px + (nn & 1));
                  o[dp * g.out_pitch + (nn >> 2)] = to16<T>(v[i]);
                }
              }
            }
          } else if (g.out_mode == PIR_OUT_FINAL_NCHW32) {
            float* o = reinterpret_cast<float*>(g.out) + (size_t)b * g.out_bstride;
            const float* im = g.img + (size_t)b * g.out_bstride;
            const size_t hw = (size_t)g.H * g.W;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n + i < g.N) {
                const size_t off = (size_t)(n + i) * hw + pix;
                o[off] = v[i] + im[off];
              }
          } else {  // PIR_OUT_NHWC32
            float* o = reinterpret_cast<float*>(g.out) + (size_t)b * g.out_bstride + (size_t)pix * g.out_pitch + n;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n + i < g.N) o[i] = v[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[ab]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------
struct C3Plan { int mt, ns, slice_n; uint32_t smem; C3Args g; };

// Shapes this kernel takes: <= 128 input channels (two k-blocks), 16-byte aligned 16-bit NHWC input, no per-image weights, and a
// channel slice (<= 64 channels, all nine taps) that fits shared memory next to two x tiles.  PIR_CONV3=0 turns it off (A/B).
static bool c3_plan(const PirGemm* d, C3Plan* p) {
  static const bool off = [] { const char* e = getenv("PIR_CONV3"); return e && e[0] == '0'; }();
  if (off || d->taps != 9 || d->w_batched || d->ln_mode || d->K > 128 || (d->K % 8)) return false;
  if (d->out_mode == PIR_OUT_UNSHUFFLE16 && (((d->H | d->W) & 1) || (d->out_pitch & 1) || (d->out_bstride & 1) || ((uintptr_t)d->out & 3))) return false;
  if (d->out_mode == PIR_OUT_SHUFFLE16 && ((d->out_pitch & 3) || (d->out_bstride & 3) || ((uintptr_t)d->out & 7))) return false;
  C3Args& g = p->g;
  g = C3Args{};
  g.B = d->B; g.H = d->H; g.W = d->W; g.K = d->K; g.N = d->N;
  g.nkb = (d->K + 63) / 64;
  const int n16 = (d->N + 15) / 16 * 16;
  p->mt = g.nkb == 1 ? 2 : 1;
  g.x_rows = (uint32_t)((2 * kC3SW + 2 + p->mt * 128 + 7) / 8 * 8);
  const uint32_t x1 = (uint32_t)g.nkb * g.x_rows * 128u, x2 = 2u * x1;
  const uint32_t budget = 227u * 1024u - 2048u;
  // widest slice (multiple of 16, <= 64) whose nine taps fit; fewer slices = fewer re-reads of the activation tile
  int sn = n16 < 64 ? n16 : 64;
  while (sn >= 16 && x2 + 9u * g.nkb * (uint32_t)sn * 128u + 1024u > budget) sn -= 16;
  if (sn < 16) return false;
  g.slice_n = sn; g.ns = (n16 + sn - 1) / sn;
  p->ns = g.ns; p->slice_n = sn;
  const uint32_t wbytes = 9u * g.nkb * (uint32_t)sn * 128u;
  g.nbuf = 2;
  while (g.nbuf < 4 && (uint32_t)(g.nbuf + 1) * x1 + wbytes + 1024u <= budget) ++g.nbuf;
  g.off_w = (uint32_t)g.nbuf * x1;
  p->smem = g.off_w + wbytes + 1024u;
  g.th = 7 * p->mt;
  g.tiles_x = (d->W + kC3TW - 1) / kC3TW; g.tiles_y = (d->H + g.th - 1) / g.th;
  g.n_items = g.tiles_x * g.tiles_y * d->B;
  auto magic = [](uint32_t dv) { return dv <= 1 ? 0u : (uint32_t)((0x100000000ull + dv - 1) / dv); };
  const uint64_t per_img = (uint64_t)g.tiles_x * g.tiles_y;
  if ((uint64_t)g.n_items * per_img >= 0x100000000ull) return false;
  g.mg_per_img = magic((uint32_t)per_img); g.mg_tiles_x = magic((uint32_t)g.tiles_x);
  return true;
}

template <class T, int MT, int NKB, int KSL>
static int c3_launch(const C3Plan& p, const CUtensorMap& tmA, const CUtensorMap& tmW, cudaStream_t stream) {
  if (!pir_smem_attr_once(reinterpret_cast<const void*>(conv3x3_kernel<T, MT, NKB, KSL>), (int)(227 * 1024 - 1024), "pir_gemm (3x3)")) return PIR_ERR_CUDA;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  C3Args g = p.g;
  int groups = num_sms / g.ns;
  if (groups < 1) groups = 1;
  if (groups > g.n_items) groups = g.n_items;
  g.groups = groups;
  conv3x3_kernel<T, MT, NKB, KSL><<<dim3((unsigned)(groups * g.ns)), dim3(kC3Threads), p.smem, stream>>>(tmA, tmW, g);
  return pir_check_launch("pir_gemm (3x3, halo tile)");
}

template <class T>
static int c3_run(const PirGemm* d, C3Plan& p, cudaStream_t stream) {
  C3Args& g = p.g;
  g.out_mode = d->out_mode; g.out = d->out; g.out_pitch = d->out_pitch; g.out_bstride = d->out_bstride;
  g.res = d->res; g.res_pitch = d->res_pitch; g.res_bstride = d->res_bstride;
  g.vec_t = d->vec_t; g.img = d->img;
  const CUtensorMapDataType dt = T::kFmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->a_pitch * 2, (uint64_t)d->a_pitch * 2 * d->W, (uint64_t)d->a_bstride * 2};
    const uint32_t box[4] = {64, (uint32_t)kC3SW, (uint32_t)(g.th + 2), 1};
    if (int e = pir_make_tmap(&tmA, dt, 4, d->a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  {
    const uint64_t ktot = (uint64_t)9 * g.nkb * 64;
    const uint64_t dims[2] = {ktot, (uint64_t)d->N};
    const uint64_t strides[1] = {ktot * 2};
    const uint32_t box[2] = {64, (uint32_t)g.slice_n};
    if (int e = pir_make_tmap(&tmW, dt, 2, d->w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  const int rem = d->K - (g.nkb - 1) * 64;
  const int ksl = rem >= 64 ? 4 : (rem + 15) >> 4;                    // k-steps of the last k-block
  if (p.mt == 2) {
    switch (ksl) {
      case 1: return c3_launch<T, 2, 1, 1>(p, tmA, tmW, stream);
      case 2: return c3_launch<T, 2, 1, 2>(p, tmA, tmW, stream);
      case 3: return c3_launch<T, 2, 1, 3>(p, tmA, tmW, stream);
      default: return c3_launch<T, 2, 1, 4>(p, tmA, tmW, stream);
    }
  }
  switch (ksl) {
    case 1: return c3_launch<T, 1, 2, 1>(p, tmA, tmW, stream);
    case 2: return c3_launch<T, 1, 2, 2>(p, tmA, tmW, stream);
    case 3: return c3_launch<T, 1, 2, 3>(p, tmA, tmW, stream);
    default: return c3_launch<T, 1, 2, 4>(p, tmA, tmW, stream);
  }
}

// entry for gemm_tcgen05.cu: returns 1 when the shape is not taken (caller falls back), else the launch status (<= 0)
int conv3x3_try(const PirGemm* d, cudaStream_t stream) {
  C3Plan p;
  if (!c3_plan(d, &p)) return 1;
  return d->dtype == PIR_DTYPE_BF16 ? c3_run<BF16>(d, p, stream) : c3_run<FP16>(d, p, stream);
}

}  // namespace pir
