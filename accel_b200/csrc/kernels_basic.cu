// Bandwidth-side kernels of the Accel hot path: stems, pooling, deformable im2col, flow-guided warp,
// layout conversion, score fusion + x16 upsampling + argmax.  sm_100a only.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace accel {

namespace {

// ------------------------------------------------------------------------------------------------
// 7x7 / s2 / p3 stem, fp32 NCHW source, 64 output channels (R101/R50 `conv1`, R18/34 `conv0`,
// FlowNet `flow_conv1` with its 2x2 average pool and /255 folded into the patch load).
// CTA = 8 x 32 output pixels; the input patch and all weights live in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int ST_TH = 8, ST_TW = 32;
constexpr int ST_PH = (ST_TH - 1) * 2 + 7;   // 21
constexpr int ST_PW = (ST_TW - 1) * 2 + 7;   // 69
constexpr int ST_PWP = ST_PW + 2;            // 71: odd row stride

__global__ void __launch_bounds__(256) stem_kernel(const StemParams P) {
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                              // [Cin*49][64]
  float* patch = smem + P.Cin * 49 * 64;         // [Cin][ST_PH][ST_PWP]
  const int tid = threadIdx.x;
  const int Hp = P.pool ? P.Hs / 2 : P.Hs, Wp = P.pool ? P.Ws / 2 : P.Ws;   // size the conv sees

  pdl_trigger();
  for (int i = tid; i < P.Cin * 49 * 64; i += 256) {
    const int n = i & 63, k = i >> 6;            // weight (n, c, ky, kx) -> ws[k][n], k = c*49 + ky*7 + kx
    ws[i] = P.weight[(size_t)n * P.Cin * 49 + k];
  }
  pdl_wait();
  const int oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  for (int i = tid; i < P.Cin * ST_PH * ST_PW; i += 256) {
    const int c = i / (ST_PH * ST_PW);
    const int r = i - c * (ST_PH * ST_PW);
    const int py = r / ST_PW, px = r - py * ST_PW;
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < Hp && ix >= 0 && ix < Wp) {
      const float* src = (c < 3 ? P.src0 : P.src1) + (size_t)(c < 3 ? c : c - 3) * P.Hs * P.Ws;
      if (P.pool) {
        const float* q = src + (size_t)(2 * iy) * P.Ws + 2 * ix;
        v = ((q[0] + q[1]) + (q[P.Ws] + q[P.Ws + 1])) * 0.25f;
      } else {
        v = src[(size_t)iy * P.Ws + ix];
      }
      v = fmaf(v, P.in_scale[c], P.in_shift[c]);
    }
    patch[(c * ST_PH + py) * ST_PWP + px] = v;
  }
  __syncthreads();

  const int ty = tid >> 5, tx = tid & 31;
  float acc[64];
#pragma unroll
  for (int n = 0; n < 64; ++n) acc[n] = 0.f;
  for (int c = 0; c < P.Cin; ++c) {
    for (int ky = 0; ky < 7; ++ky) {
      const float* prow = patch + (c * ST_PH + ty * 2 + ky) * ST_PWP + tx * 2;
      const float* wrow = ws + (size_t)((c * 7 + ky) * 7) * 64;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float a = prow[kx];
        const float4* w4 = reinterpret_cast<const float4*>(wrow + kx * 64);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 w = w4[q];
          acc[4 * q + 0] = fmaf(a, w.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(a, w.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(a, w.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(a, w.w, acc[4 * q + 3]);
        }
      }
    }
  }
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy < P.Ho && ox < P.Wo) {
    const int pix = oy * P.Wo + ox;
#pragma unroll
    for (int g = 0; g < 8; ++g) epilogue_store<8>(P.epi, pix, g * 8, acc + g * 8);
  }
}

// ------------------------------------------------------------------------------------------------
// Pooling over the split format: one thread per (output pixel, 8 channels).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_kernel(const PoolParams P) {
  pdl_trigger();
  pdl_wait();
  const int groups = (P.C + 7) / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P.Ho * P.Wo * groups) return;
  const int g = (int)(idx % groups);
  const int p = (int)(idx / groups);
  const int oy = p / P.Wo, ox = p - oy * P.Wo;
  float best[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) best[i] = P.is_max ? -INFINITY : 0.f;
  for (int ky = 0; ky < P.kernel; ++ky) {
    const int iy = oy * P.stride - P.pad + ky;
    if (iy < 0 || iy >= P.Hin) continue;
    for (int kx = 0; kx < P.kernel; ++kx) {
      const int ix = ox * P.stride - P.pad + kx;
      if (ix < 0 || ix >= P.Win) continue;
      float v[8];
      const size_t off = ((size_t)iy * P.Win + ix) * P.in_ld + g * 8;
      load8(P.in_hi + off, P.in_lo + off, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) best[i] = P.is_max ? fmaxf(best[i], v[i]) : best[i] + v[i];
    }
  }
  if (!P.is_max) {
    const float inv = 1.f / (float)(P.kernel * P.kernel);
#pragma unroll
    for (int i = 0; i < 8; ++i) best[i] *= inv;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = g * 8 + i;
    if (c >= P.C) best[i] = 0.f;
    else if (P.scale) best[i] = apply_act(fmaf(best[i], P.scale[c], P.shift[c]), P.act);
  }
  const size_t o = (size_t)p * P.out_ld + g * 8;
  store8(P.out_hi + o, P.out_lo + o, best);
}

// ------------------------------------------------------------------------------------------------
// Deformable im2col, DCNv1 rule (oracle/ops.py:deformable_convolution): sample = 0 unless
// 0 <= p < size; inside, bilinear with the high neighbour clamped to size-1.
// One thread per (pixel, tap, 8 channels); the four neighbours are 16-byte NHWC loads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dcn_col_kernel(const DcnColParams P) {
  pdl_trigger();
  pdl_wait();
  const int groups = P.C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P.H * P.W * 9 * groups) return;
  const int g = (int)(idx % groups);
  const int t = (int)((idx / groups) % 9);
  const int p = (int)(idx / ((long long)groups * 9));
  const int oy = p / P.W, ox = p - oy * P.W;
  const int cpg = P.C / P.dg;
  const int dgi = (g * 8) / cpg;
  const int ti = t / 3, tj = t - ti * 3;
  const size_t plane = (size_t)P.H * P.W;
  const float offy = P.offset[(size_t)(dgi * 18 + 2 * t) * plane + p];
  const float offx = P.offset[(size_t)(dgi * 18 + 2 * t + 1) * plane + p];
  float py = (float)(oy - P.pad) + (float)(ti * P.dilate) + offy;
  float px = (float)(ox - P.pad) + (float)(tj * P.dilate) + offx;
  float out[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (py >= 0.f && px >= 0.f && py < (float)P.H && px < (float)P.W) {
    int y0 = (int)floorf(py), x0 = (int)floorf(px);
    int y1, x1;
    if (y0 >= P.H - 1) { y0 = y1 = P.H - 1; py = (float)y0; } else { y1 = y0 + 1; }
    if (x0 >= P.W - 1) { x0 = x1 = P.W - 1; px = (float)x0; } else { x1 = x0 + 1; }
    const float ly = py - (float)y0, lx = px - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
    float a[8], b[8], c[8], d[8];
    const int c0 = g * 8;
    size_t o;
    o = ((size_t)y0 * P.W + x0) * P.in_ld + c0; load8(P.in_hi + o, P.in_lo + o, a);
    o = ((size_t)y0 * P.W + x1) * P.in_ld + c0; load8(P.in_hi + o, P.in_lo + o, b);
    o = ((size_t)y1 * P.W + x0) * P.in_ld + c0; load8(P.in_hi + o, P.in_lo + o, c);
    o = ((size_t)y1 * P.W + x1) * P.in_ld + c0; load8(P.in_hi + o, P.in_lo + o, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = ((a[i] * w00 + b[i] * w01) + c[i] * w10) + d[i] * w11;
  }
  const size_t oo = (size_t)p * P.col_ld + (size_t)t * P.C + g * 8;
  store8(P.col_hi + oo, P.col_lo + oo, out);
}

// Warp-per-(pixel, tap) variant: the lanes walk the 8-channel groups, so the sampling position, the clamping and the
// four bilinear weights -- shared by every channel of a deformable group -- are set up once per deformable group
// instead of once per 8 channels, and no 64-bit index division is left (the one-thread-per-(pixel, tap, 8 channels)
// kernel above spends ~510 instructions per thread, most of them on exactly that; ncu: issue slots 73 % busy).
// Same loads (16 bytes per lane, 512 contiguous bytes per warp and corner), same arithmetic order, same stores.
__global__ void __launch_bounds__(256) dcn_col_warp_kernel(const DcnColParams P) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);        // (pixel, tap) index
  const int npix = P.H * P.W;
  if (w >= npix * 9) return;
  const int p = w / 9, t = w - p * 9;
  const int oy = p / P.W, ox = p - oy * P.W;
  const int ti = t / 3, tj = t - ti * 3;
  const int groups = P.C >> 3, gpd = groups / P.dg;                          // 8-channel groups, groups per deformable group
  const size_t plane = (size_t)npix;
  const float by = (float)(oy - P.pad) + (float)(ti * P.dilate), bx = (float)(ox - P.pad) + (float)(tj * P.dilate);
  int cur_dg = -1;
  bool inside = false;
  size_t o00 = 0, o01 = 0, o10 = 0, o11 = 0;
  float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
  for (int g = lane; g < groups; g += 32) {
    const int dgi = g / gpd;
    if (dgi != cur_dg) {
      cur_dg = dgi;
      float py = by + P.offset[(size_t)(dgi * 18 + 2 * t) * plane + p];
      float px = bx + P.offset[(size_t)(dgi * 18 + 2 * t + 1) * plane + p];
      inside = py >= 0.f && px >= 0.f && py < (float)P.H && px < (float)P.W;
      if (inside) {
        int y0 = (int)floorf(py), x0 = (int)floorf(px);
        int y1, x1;
        if (y0 >= P.H - 1) { y0 = y1 = P.H - 1; py = (float)y0; } else { y1 = y0 + 1; }
        if (x0 >= P.W - 1) { x0 = x1 = P.W - 1; px = (float)x0; } else { x1 = x0 + 1; }
        const float ly = py - (float)y0, lx = px - (float)x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        w00 = hy * hx; w01 = hy * lx; w10 = ly * hx; w11 = ly * lx;
        o00 = ((size_t)y0 * P.W + x0) * P.in_ld; o01 = ((size_t)y0 * P.W + x1) * P.in_ld;
        o10 = ((size_t)y1 * P.W + x0) * P.in_ld; o11 = ((size_t)y1 * P.W + x1) * P.in_ld;
      }
    }
    float out[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int c0 = g * 8;
    if (inside) {
      float a[8], b[8], c[8], d[8];
      load8(P.in_hi + o00 + c0, P.in_lo + o00 + c0, a);
      load8(P.in_hi + o01 + c0, P.in_lo + o01 + c0, b);
      load8(P.in_hi + o10 + c0, P.in_lo + o10 + c0, c);
      load8(P.in_hi + o11 + c0, P.in_lo + o11 + c0, d);
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = ((a[i] * w00 + b[i] * w01) + c[i] * w10) + d[i] * w11;
    }
    const size_t oo = (size_t)p * P.col_ld + (size_t)t * P.C + c0;
    store8(P.col_hi + oo, P.col_lo + oo, out);
  }
}

// Tiled variant (default since load8 / store8 issue 128-bit accesses; ACCEL_DCN_TILED=0 = the per-thread kernel): a CTA owns an 8 x 4 block of output pixels and 128 channels (of one deformable group), a
// half-warp one (pixel, tap), a lane eight channels.  The 9 taps x 4 corners of neighbouring pixels overlap heavily (the taps sit
// `dilate` pixels apart, the learned offsets move them by a few pixels), so with all of them gathered by the same SM the
// L1 cache serves most of the 128-byte corner reads: the per-thread kernel above spreads one pixel's taps over several
// SMs and fetches every corner from L2 (ncu: l1tex 90 % busy, 604 MB L2 -> SM per 64 x 128 x 512 layer).  Sampling
// position, clamping and weights are computed once per (pixel, tap, group); same arithmetic order, same stores.
constexpr int DCN_TBX = 8, DCN_TBY = 4;   // output pixels per CTA (x, y)
constexpr int DCN_CB = 128;          // channels per CTA: 16 lanes x 8 channels
__global__ void __launch_bounds__(256) dcn_col_tile_kernel(const DcnColParams P) {
  pdl_trigger();
  pdl_wait();
  const int tiles_x = (P.W + DCN_TBX - 1) / DCN_TBX;
  const int ty0 = (blockIdx.x / tiles_x) * DCN_TBY, tx0 = (blockIdx.x % tiles_x) * DCN_TBX;
  const int cb = blockIdx.y * DCN_CB;                              // this CTA's DCN_CB channels (inside one deformable group)
  const int cpg = P.C / P.dg;
  const int dgi = cb / cpg;
  const int half = threadIdx.x >> 4, hl = threadIdx.x & 15;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);            // the 16 lanes of this half-warp
  const size_t plane = (size_t)P.H * P.W;
  const float* offp = P.offset + (size_t)(dgi * 18) * plane;
  const int c0 = cb + hl * 8;
  // a half-warp walks whole pixels: the pixel's 18 offsets are fetched once (lane l: offsets l and 16 + (l & 1)) and
  // handed round by shuffles, so the nine taps are nine independent gather / interpolate / store groups
  for (int pl = half; pl < DCN_TBX * DCN_TBY; pl += 16) {
    const int oy = ty0 + pl / DCN_TBX, ox = tx0 + pl % DCN_TBX;
    if (oy >= P.H || ox >= P.W) continue;                          // (uniform per half-warp)
    const int p = oy * P.W + ox;
    const float o_a = __ldg(offp + (size_t)hl * plane + p);
    const float o_b = __ldg(offp + (size_t)(16 + (hl & 1)) * plane + p);
    const float by = (float)(oy - P.pad), bx = (float)(ox - P.pad);
    const size_t orow = (size_t)p * P.col_ld + c0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ti = t / 3, tj = t - ti * 3;
      const float offy = t < 8 ? __shfl_sync(hmask, o_a, 2 * t, 16) : __shfl_sync(hmask, o_b, 0, 16);
      const float offx = t < 8 ? __shfl_sync(hmask, o_a, 2 * t + 1, 16) : __shfl_sync(hmask, o_b, 1, 16);
      float py = by + (float)(ti * P.dilate) + offy;
      float px = bx + (float)(tj * P.dilate) + offx;
      float out[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (py >= 0.f && px >= 0.f && py < (float)P.H && px < (float)P.W) {
        int y0 = (int)floorf(py), x0 = (int)floorf(px);
        int y1, x1;
        if (y0 >= P.H - 1) { y0 = y1 = P.H - 1; py = (float)y0; } else { y1 = y0 + 1; }
        if (x0 >= P.W - 1) { x0 = x1 = P.W - 1; px = (float)x0; } else { x1 = x0 + 1; }
        const float ly = py - (float)y0, lx = px - (float)x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
        const size_t o00 = ((size_t)y0 * P.W + x0) * P.in_ld + c0, o01 = ((size_t)y0 * P.W + x1) * P.in_ld + c0;
        const size_t o10 = ((size_t)y1 * P.W + x0) * P.in_ld + c0, o11 = ((size_t)y1 * P.W + x1) * P.in_ld + c0;
        float a[8], b[8], c[8], d[8];
        load8(P.in_hi + o00, P.in_lo + o00, a);
        load8(P.in_hi + o01, P.in_lo + o01, b);
        load8(P.in_hi + o10, P.in_lo + o10, c);
        load8(P.in_hi + o11, P.in_lo + o11, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = ((a[i] * w00 + b[i] * w01) + c[i] * w10) + d[i] * w11;
      }
      store8(P.col_hi + orow + (size_t)t * P.C, P.col_lo + orow + (size_t)t * P.C, out);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Flow-guided warp: GridGenerator(transform_type='warp') + BilinearSampler
// (dff_deeplab/symbols/accel_18.py:174-175).  The sampling position is computed with the same fp32
// operation sequence as the MXNet operators (normalise to [-1,1], de-normalise), un-fused, so the
// warped feature is bit-identical to the oracle's.  CTA = 32 consecutive pixels x 64 channels: each
// lane owns a pixel (its 2x2 neighbourhood + weights are computed once and reused down the channel
// axis, reads and NCHW writes are coalesced along x), then the tile is transposed through shared
// memory into the split NHWC copy the task head consumes.
// ------------------------------------------------------------------------------------------------
// One thread = one output pixel; its four tap pointers and bilinear weights are computed once and reused
// down the channel axis.  Channels are walked in batches of 8: the 32 gathers of a batch are issued before
// any use (32 independent 128-byte requests per warp in flight) and, when the plane size is a compile-time
// constant, every one of them is `[tap pointer + immediate]` -- no per-channel address arithmetic, ~12
// instructions per output element -- so the kernel is pure HBM streaming: 2*C*H*W*4 + 2*H*W*4 algorithmic
// bytes.  Reads and NCHW writes are coalesced along x.  The grid is sized to one resident wave: blockIdx.x =
// block of 128 pixels, blockIdx.y strides over the 8-channel batches.
constexpr int WP_THREADS = 128, WP_UNROLL = 8;

template <int NPIX>
__global__ void __launch_bounds__(WP_THREADS) warp_kernel(const float* __restrict__ feat, const float* __restrict__ flow,
                                                          float* __restrict__ out, int C, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const int npix = NPIX > 0 ? NPIX : H * W;
  const int p = blockIdx.x * WP_THREADS + threadIdx.x;
  if (p >= npix) return;
  const int y = p / W, x = p - y * W;
  const float sx = (float)(W - 1) / 2.0f, sy = (float)(H - 1) / 2.0f;
  const float gx = __fsub_rn(__fdiv_rn(__fadd_rn(flow[p], (float)x), sx), 1.0f);
  const float gy = __fsub_rn(__fdiv_rn(__fadd_rn(flow[npix + p], (float)y), sy), 1.0f);
  const float xr = __fmul_rn(__fadd_rn(gx, 1.0f), sx);
  const float yr = __fmul_rn(__fadd_rn(gy, 1.0f), sy);
  const float xf = floorf(xr), yf = floorf(yr);
  const float wx0 = __fsub_rn(1.0f, __fsub_rn(xr, xf)), wy0 = __fsub_rn(1.0f, __fsub_rn(yr, yf));
  const float wx1 = __fsub_rn(1.0f, wx0), wy1 = __fsub_rn(1.0f, wy0);
  const bool x0ok = xf >= 0.f && xf <= (float)(W - 1), x1ok = xf + 1.f >= 0.f && xf + 1.f <= (float)(W - 1);
  const bool y0ok = yf >= 0.f && yf <= (float)(H - 1), y1ok = yf + 1.f >= 0.f && yf + 1.f <= (float)(H - 1);
  // clamp before the int conversion would overflow for wild flows
  const bool sane = fabsf(xr) < 1e9f && fabsf(yr) < 1e9f;
  int i00 = 0, i01 = 0, i10 = 0, i11 = 0;
  if (sane) {
    const int xi0 = min(max((int)xf, 0), W - 1), xi1 = min(max((int)xf + 1, 0), W - 1);
    const int yi0 = min(max((int)yf, 0), H - 1), yi1 = min(max((int)yf + 1, 0), H - 1);
    i00 = yi0 * W + xi0; i01 = yi0 * W + xi1; i10 = yi1 * W + xi0; i11 = yi1 * W + xi1;
  }
  const float w00 = (sane && y0ok && x0ok) ? __fmul_rn(wy0, wx0) : kSkipTap;
  const float w01 = (sane && y0ok && x1ok) ? __fmul_rn(wy0, wx1) : kSkipTap;
  const float w10 = (sane && y1ok && x0ok) ? __fmul_rn(wy1, wx0) : kSkipTap;
  const float w11 = (sane && y1ok && x1ok) ? __fmul_rn(wy1, wx1) : kSkipTap;

  const int nbatch = (C + WP_UNROLL - 1) / WP_UNROLL;
  for (int bt = blockIdx.y; bt < nbatch; bt += gridDim.y) {
    const int c0 = bt * WP_UNROLL;
    const size_t base = (size_t)c0 * npix;
    const float* p00 = feat + base + i00;
    const float* p01 = feat + base + i01;
    const float* p10 = feat + base + i10;
    const float* p11 = feat + base + i11;
    float* po = out + base + p;
    float a[WP_UNROLL], b[WP_UNROLL], c[WP_UNROLL], d[WP_UNROLL];
    if (c0 + WP_UNROLL <= C) {
#pragma unroll
      for (int k = 0; k < WP_UNROLL; ++k) {
        a[k] = __ldg(p00 + (size_t)k * npix);
        b[k] = __ldg(p01 + (size_t)k * npix);
        c[k] = __ldg(p10 + (size_t)k * npix);
        d[k] = __ldg(p11 + (size_t)k * npix);
      }
#pragma unroll
      for (int k = 0; k < WP_UNROLL; ++k) {
        // taps with a zero weight (out of range, as MXNet's BilinearSampler skips them) contribute by VALUE zero: an Inf /
        // NaN in the clamped neighbour must not become NaN here
        float v = __fmul_rn(keep_tap(w00) ? a[k] : 0.f, w00);
        v = __fadd_rn(v, __fmul_rn(keep_tap(w01) ? b[k] : 0.f, w01));
        v = __fadd_rn(v, __fmul_rn(keep_tap(w10) ? c[k] : 0.f, w10));
        v = __fadd_rn(v, __fmul_rn(keep_tap(w11) ? d[k] : 0.f, w11));
        po[(size_t)k * npix] = v;
      }
    } else {
      for (int k = 0; c0 + k < C; ++k) {
        float v = __fmul_rn(keep_tap(w00) ? __ldg(p00 + (size_t)k * npix) : 0.f, w00);
        v = __fadd_rn(v, __fmul_rn(keep_tap(w01) ? __ldg(p01 + (size_t)k * npix) : 0.f, w01));
        v = __fadd_rn(v, __fmul_rn(keep_tap(w10) ? __ldg(p10 + (size_t)k * npix) : 0.f, w10));
        v = __fadd_rn(v, __fmul_rn(keep_tap(w11) ? __ldg(p11 + (size_t)k * npix) : 0.f, w11));
        po[(size_t)k * npix] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 NCHW <-> split NHWC.  nchw -> split: CTA = 32 pixels x 64 channels; lanes read along x
// (coalesced), the tile is transposed through shared memory and leaves as 128 contiguous bytes per
// pixel and plane.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_split_kernel(const float* __restrict__ src, int C, int npix,
                                                            __half* hi, __half* lo, int ld, const float* __restrict__ bias,
                                                            int act) {
  __shared__ float tile[64][33];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + wid * 8 + k, p = p0 + lane;
    v[k] = (c < C && p < npix) ? __ldg(src + (size_t)c * npix + p) : 0.f;
    if (bias && c < C) v[k] = apply_act(v[k] + __ldg(bias + c), act);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) tile[wid * 8 + k][lane] = v[k];
  __syncthreads();
  const int pl = threadIdx.x >> 3, g = threadIdx.x & 7;
  const int p = p0 + pl;
  // only the tensor's own channels (rounded up to the 8-channel store): the destination may be a channel view of a
  // wider concat buffer whose other channels belong to another producer
  if (p < npix && c0 + g * 8 < C && c0 + g * 8 < ld) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = tile[g * 8 + i][pl];
    const size_t off = (size_t)p * ld + c0 + g * 8;
    store8(hi + off, lo + off, o);
  }
}

__global__ void __launch_bounds__(256) split_to_nchw_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                                            int ld, int C, int npix, float* dst) {
  __shared__ float tile[32][33];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (threadIdx.x < 128) {
    const int pl = threadIdx.x >> 2, g = threadIdx.x & 3;
    const int p = p0 + pl;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p < npix && c0 + g * 8 < ld) {
      const size_t o = (size_t)p * ld + c0 + g * 8;
      load8(hi + o, lo + o, v);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[g * 8 + i][pl] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + wid * 4 + k, p = p0 + lane;
    if (c < C && p < npix) dst[(size_t)c * npix + p] = tile[wid * 4 + k][lane];
  }
}

// ------------------------------------------------------------------------------------------------
// FlowNet `upsample_flow*`: Deconvolution(2 -> 2, k4, s2, p0) + Crop(1,1) == transposed conv pad 1.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upflow_kernel(const UpflowParams P) {
  pdl_trigger();
  pdl_wait();
  const int OW = 2 * P.W, OH = 2 * P.H;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= OW * OH) return;
  const int oy = idx / OW, ox = idx - oy * OW;
  float acc[2] = {P.bias[0], P.bias[1]};
  for (int ky = 0; ky < 4; ++ky) {
    const int ty = oy + 1 - ky;
    if (ty < 0 || (ty & 1)) continue;
    const int iy = ty >> 1;
    if (iy >= P.H) continue;
    for (int kx = 0; kx < 4; ++kx) {
      const int tx = ox + 1 - kx;
      if (tx < 0 || (tx & 1)) continue;
      const int ix = tx >> 1;
      if (ix >= P.W) continue;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const float v = P.flow[(size_t)ci * P.H * P.W + iy * P.W + ix];
        acc[0] = fmaf(v, P.weight[(ci * 2 + 0) * 16 + ky * 4 + kx], acc[0]);
        acc[1] = fmaf(v, P.weight[(ci * 2 + 1) * 16 + ky * 4 + kx], acc[1]);
      }
    }
  }
#pragma unroll
  for (int co = 0; co < 2; ++co) {
    __half h, l;
    split_f32(acc[co], h, l);
    P.out_hi[(size_t)idx * P.out_ld + co] = h;
    P.out_lo[(size_t)idx * P.out_ld + co] = l;
  }
}

// ------------------------------------------------------------------------------------------------
// Score-level fusion at feature resolution.  Both x16 upsamplings use the same channel-independent
// bilinear kernel, so correction(concat(up(a), up(b))) == up(Wa a + Wb b) + bias: the 38 -> 19 conv
// (accel_18.py:229-235) runs on the 64x128 maps and the bias is added after interpolation.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fuse_lowres_kernel(const FuseParams P) {
  extern __shared__ float wsm[];
  pdl_trigger();
  for (int i = threadIdx.x; i < P.K * 2 * P.K; i += blockDim.x) wsm[i] = P.w[i];
  __syncthreads();
  pdl_wait();
  const int npix = P.h * P.w_;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  float in[64];
  for (int j = 0; j < P.K; ++j) {
    in[j] = P.a[(size_t)j * npix + p];
    in[P.K + j] = P.b[(size_t)j * npix + p];
  }
  for (int c = 0; c < P.K; ++c) {
    float acc = 0.f;
    for (int j = 0; j < 2 * P.K; ++j) acc = fmaf(wsm[c * 2 * P.K + j], in[j], acc);
    P.out[(size_t)c * npix + p] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// x16 bilinear upsampling (grouped 32x32/s16 Deconvolution + Crop(8,8), accel_18.py:193-197) fused
// with the per-pixel argmax (demo.py:238,245,252).  out[Y] = sum_i s[i] * w1[Y + 8 - 16 i] with
// w1[k] = 1 - |k/16 - 31/32|: two taps per axis, zero beyond the map (no clamping).  One thread owns
// four horizontally adjacent pixels (they share the same source columns); labels leave as uchar4.
// The fp32 score volume is written only when the caller asks for it (parity mode).
// ------------------------------------------------------------------------------------------------
// KC > 0: class count known at compile time (19 for Cityscapes): the class loop is fully unrolled, so the 4 x KC
// low-resolution loads of a thread are all in flight before the first use.
template <int KC, int MINB = 1>
__global__ void __launch_bounds__(256, MINB) tail_kernel(const TailParams P) {
  pdl_trigger();
  pdl_wait();
  const int OW = P.w * P.factor, OH = P.h * P.factor;
  const int qw = OW / 4;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= qw * OH) return;
  const int Y = idx / qw, X0 = (idx - Y * qw) * 4;
  const int f = P.factor, half = f / 2;
  const float inv = 1.f / (float)f, cen = (float)(2 * f - 1) / (float)(2 * f);
  const int i0 = (Y + half) / f, ky = (Y + half) - i0 * f;
  const float wy0 = (i0 < P.h) ? 1.f - fabsf((float)ky * inv - cen) : 0.f;            // source row i0
  const float wy1 = (i0 >= 1) ? 1.f - fabsf((float)(ky + f) * inv - cen) : 0.f;       // source row i0-1
  const int j0 = (X0 + half) / f, kx = (X0 + half) - j0 * f;
  const bool c0ok = j0 < P.w, c1ok = j0 >= 1;
  float wx0[4], wx1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    wx0[q] = c0ok ? 1.f - fabsf((float)(kx + q) * inv - cen) : 0.f;
    wx1[q] = c1ok ? 1.f - fabsf((float)(kx + q + f) * inv - cen) : 0.f;
  }
  const int r0 = min(i0, P.h - 1), r1 = max(i0 - 1, 0), q0 = min(j0, P.w - 1), q1 = max(j0 - 1, 0);
  const size_t plane = (size_t)P.h * P.w, oplane = (size_t)OH * OW;
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int arg[4] = {0, 0, 0, 0};
  const int K = KC > 0 ? KC : P.K;
#pragma unroll
  for (int c = 0; c < K; ++c) {
    const float* s = P.score + c * plane;
    const float s00 = s[r0 * P.w + q0], s01 = s[r0 * P.w + q1], s10 = s[r1 * P.w + q0], s11 = s[r1 * P.w + q1];
    const float b = P.bias ? P.bias[c] : 0.f;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = ((s00 * (wy0 * wx0[q]) + s01 * (wy0 * wx1[q])) + s10 * (wy1 * wx0[q])) + s11 * (wy1 * wx1[q]) + b;
      if (v[q] > best[q]) { best[q] = v[q]; arg[q] = c; }      // strict: ties keep the lowest class
    }
    if (P.score_out)
      *reinterpret_cast<float4*>(P.score_out + c * oplane + (size_t)Y * OW + X0) = make_float4(v[0], v[1], v[2], v[3]);
  }
  *reinterpret_cast<uchar4*>(P.label + (size_t)Y * OW + X0) =
      make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
}

// Band version of the tail for the shape that matters (factor 16, KC classes): every 16 x 16 output cell
// {Y+8 in [16 i0, 16 i0 + 16)} x {X+8 in [16 j0, 16 j0 + 16)} interpolates the SAME 2 x 2 low-resolution neighbourhood,
// so the interpolation is evaluated separably -- per class and quad column once `top = wx0 s[i0][j0] + wx1 s[i0][j0-1]`
// and `bot` (row i0-1), then two FMAs per pixel -- instead of four loads and four multiply-adds per pixel and class.
// CTA = 16 output rows x 512 columns; its 3 x 34 x KC source window is staged in shared memory (clamped indices, like
// the kernel above: weights, not values, are zeroed beyond the map); thread = one quad column x 8 rows (one source-row
// pair), 32 running (best, arg) pairs in registers; labels leave as uchar4, scores (parity mode) as float4.
// FUSE: the low-resolution scores are `fuse_w` (KC x 2KC, corr_weight) applied to concat(fuse_a, fuse_b) -- the
// score-level fusion of accel_18.py:229-235 moved to feature resolution (fuse_lowres_kernel) -- evaluated here for the
// CTA's own 3 x 34 source window, with that kernel's operation order (bit-identical), instead of by a separate launch.
template <int KC, bool SCORES, bool FUSE>
__global__ void __launch_bounds__(256, 2) tail_band_kernel(const TailParams P) {
  constexpr int F = 16, TW = 512, NC = TW / F + 2;
  __shared__ float ss[KC + 1][3][NC];                       // + one dummy class: the loop prefetches one ahead
  __shared__ float sbias[KC];
  __shared__ float sraw[FUSE ? 2 * KC : 1][3 * NC];
  __shared__ float sw[FUSE ? KC : 1][2 * KC];
  pdl_trigger();
  if (FUSE)
    for (int i = threadIdx.x; i < KC * 2 * KC; i += 256) sw[i / (2 * KC)][i % (2 * KC)] = P.fuse_w[i];
  pdl_wait();
  const int OW = P.w * F, OH = P.h * F;
  const int m = blockIdx.y, kc = blockIdx.x;
  const size_t plane = (size_t)P.h * P.w, oplane = (size_t)OH * OW;
  if (FUSE) {
    for (int i = threadIdx.x; i < 2 * KC * 3 * NC; i += 256) {
      const int c = i / (3 * NC), rem = i - c * 3 * NC, r = rem / NC, j = rem - r * NC;
      const int sy = min(max(m - 1 + r, 0), P.h - 1), sx = min(max(kc * (TW / F) - 1 + j, 0), P.w - 1);
      const float* src = c < KC ? P.fuse_a + c * plane : P.fuse_b + (c - KC) * plane;
      sraw[c][rem] = src[(size_t)sy * P.w + sx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < KC * 3 * NC; i += 256) {
      const int c = i / (3 * NC), px = i - c * 3 * NC;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 2 * KC; ++j) acc = fmaf(sw[c][j], sraw[j][px], acc);
      (&ss[c][0][0])[px] = acc;
    }
  } else {
    for (int i = threadIdx.x; i < KC * 3 * NC; i += 256) {
      const int c = i / (3 * NC), rem = i - c * 3 * NC, r = rem / NC, j = rem - r * NC;
      const int sy = min(max(m - 1 + r, 0), P.h - 1), sx = min(max(kc * (TW / F) - 1 + j, 0), P.w - 1);
      ss[c][r][j] = P.score[c * plane + (size_t)sy * P.w + sx];
    }
  }
  if (threadIdx.x < 3 * NC) ss[KC][threadIdx.x / NC][threadIdx.x % NC] = 0.f;
  if (threadIdx.x < KC) sbias[threadIdx.x] = P.bias ? P.bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int qx = threadIdx.x & 127, rh = threadIdx.x >> 7;
  const int X0 = kc * TW + 4 * qx;
  if (X0 >= OW) return;
  const float inv = 1.f / (float)F, cen = (float)(2 * F - 1) / (float)(2 * F);
  const int j0 = (X0 + F / 2) / F, kx = (X0 + F / 2) - j0 * F;
  const int lc0 = j0 - (kc * (TW / F) - 1);                 // shared-memory column of source column j0 (j0-1: lc0-1)
  const bool c0ok = j0 < P.w, c1ok = j0 >= 1;
  float wx0[4], wx1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    wx0[q] = c0ok ? 1.f - fabsf((float)(kx + q) * inv - cen) : 0.f;
    wx1[q] = c1ok ? 1.f - fabsf((float)(kx + q + F) * inv - cen) : 0.f;
  }
  const int Yb = m * F + rh * 8;                            // first of this thread's 8 rows
  const int i0 = m + rh, ky0 = rh ? 0 : 8;                  // (Y + 8) / 16 and (Y + 8) % 16 of row Yb
  float wy0[8], wy1[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    wy0[r] = (i0 < P.h) ? 1.f - fabsf((float)(ky0 + r) * inv - cen) : 0.f;
    wy1[r] = (i0 >= 1) ? 1.f - fabsf((float)(ky0 + r + F) * inv - cen) : 0.f;
  }
  float best[8][4];
  unsigned arg[8];                                          // four 8-bit class ids per row
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    arg[r] = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) best[r][q] = -INFINITY;
  }
  const float* s0 = &ss[0][1 + rh][lc0];                    // source row i0   (class stride 3 * NC)
  const float* s1 = &ss[0][rh][lc0];                        // source row i0-1
  float a0 = s0[0], a1 = s0[-1], b0 = s1[0], b1 = s1[-1];
  float* so = SCORES ? P.score_out + (size_t)Yb * OW + X0 : nullptr;
  // the class loop stays rolled: its body (8 rows x 4 pixels) is ~4 KB of code that all 19 iterations reuse from the
  // instruction cache; fully unrolled it is > 100 KB of straight-line code and instruction fetch becomes the limit
#pragma unroll 1
  for (int c = 0; c < KC; ++c) {
    float top[4], bot[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      top[q] = __fmaf_rn(a1, wx1[q], __fmul_rn(a0, wx0[q]));
      bot[q] = __fmaf_rn(b1, wx1[q], __fmul_rn(b0, wx0[q]));
    }
    const float b = sbias[c];
    s0 += 3 * NC; s1 += 3 * NC;                             // next class's taps, in flight under this class's math
    a0 = s0[0]; a1 = s0[-1]; b0 = s1[0]; b1 = s1[-1];
    const unsigned cb = (unsigned)c * 0x01010101u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q] = __fadd_rn(__fmaf_rn(wy1[r], bot[q], __fmul_rn(wy0[r], top[q])), b);
        if (v[q] > best[r][q]) {                            // strict: ties keep the lowest class
          best[r][q] = v[q];
          arg[r] = (arg[r] & ~(0xffu << (8 * q))) | (cb & (0xffu << (8 * q)));
        }
      }
      if (SCORES && Yb + r < OH) *reinterpret_cast<float4*>(so + (size_t)r * OW) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (SCORES) so += oplane;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (Yb + r < OH) *reinterpret_cast<unsigned*>(P.label + (size_t)(Yb + r) * OW + X0) = arg[r];
}

}  // namespace

cudaError_t launch_stem(const StemParams& P, cudaStream_t stream) {
  const size_t smem = ((size_t)P.Cin * 49 * 64 + (size_t)P.Cin * ST_PH * ST_PWP) * sizeof(float);
  if (first_time_on_device(ONCE_STEM)) {
    cudaError_t e = cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((P.Wo + ST_TW - 1) / ST_TW, (P.Ho + ST_TH - 1) / ST_TH);
  return launch_k(stem_kernel, grid, dim3(256), smem, stream, P);
}

cudaError_t launch_pool(const PoolParams& P, cudaStream_t stream) {
  const long long work = (long long)P.Ho * P.Wo * ((P.C + 7) / 8);
  return launch_k(pool_kernel, dim3((unsigned)((work + 255) / 256)), dim3(256), 0, stream, P);
}

cudaError_t launch_dcn_col(const DcnColParams& P, cudaStream_t stream) {
  static int warp_variant = -1;
  if (warp_variant < 0) { const char* e = getenv("ACCEL_DCN_WARP"); warp_variant = (e && e[0] == '1') ? 1 : 0; }
  if (warp_variant && P.C % 8 == 0 && (P.C / 8) % P.dg == 0) {
    const long long warps = (long long)P.H * P.W * 9;
    return launch_k(dcn_col_warp_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, stream, P);
  }
  static int tiled = -1;                         // ACCEL_DCN_TILED=0: the one-thread-per-(pixel, tap, 8 channels) kernel
  if (tiled < 0) { const char* e = getenv("ACCEL_DCN_TILED"); tiled = (e && e[0] == '0') ? 0 : 1; }
  if (tiled && P.C % P.dg == 0 && (P.C / P.dg) % DCN_CB == 0) {
    const int tiles = ((P.W + DCN_TBX - 1) / DCN_TBX) * ((P.H + DCN_TBY - 1) / DCN_TBY);
    return launch_k(dcn_col_tile_kernel, dim3((unsigned)tiles, (unsigned)(P.C / DCN_CB)), dim3(256), 0, stream, P);
  }
  const long long work = (long long)P.H * P.W * 9 * (P.C / 8);
  return launch_k(dcn_col_kernel, dim3((unsigned)((work + 255) / 256)), dim3(256), 0, stream, P);
}

template <int NPIX>
static cudaError_t launch_warp_t(const WarpParams& P, cudaStream_t stream) {
  static int slots = 0;                         // resident CTAs of this instantiation on the whole device
  if (!slots) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, warp_kernel<NPIX>, WP_THREADS, 0);
    slots = sms * (per_sm > 0 ? per_sm : 1);
  }
  const int bx = (P.H * P.W + WP_THREADS - 1) / WP_THREADS;
  const int nbatch = (P.C + WP_UNROLL - 1) / WP_UNROLL;
  int by = slots / bx;
  if (by < 1) by = 1;
  if (by > nbatch) by = nbatch;
  return launch_k(warp_kernel<NPIX>, dim3(bx, by), dim3(WP_THREADS), 0, stream, P.feat, P.flow, P.out_nchw, P.C, P.H, P.W);
}

static bool env_gather_only() {
  static int v = -1;
  if (v < 0) { const char* s = getenv("ACCEL_WARP_GATHER"); v = (s && atoi(s)) ? 1 : 0; }
  return v == 1;
}

cudaError_t launch_warp(const WarpParams& P, cudaStream_t stream) {
  // One pass for both consumers where the fused kernel fits (warp_staged.cu: fp32 NCHW `warping_feat_output` and / or
  // the split NHWC copy the task head loads); otherwise the warped feature lands as fp32 NCHW first and the split
  // copy is a second, L2-fed pass.
  if (!env_gather_only() && warp_fused_supported(P)) {
    cudaError_t e = launch_warp_fused(P, stream);
    if (e != cudaErrorNotSupported) return e;
    cudaGetLastError();
  }
  if (!P.out_nchw) return cudaErrorInvalidValue;
  const int npix = P.H * P.W;
  cudaError_t e = env_gather_only() ? cudaErrorNotSupported : launch_warp_staged(P, stream);
  if (e == cudaSuccess) {
    if (!P.out_hi) return e;
    return launch_nchw_to_split(P.out_nchw, P.C, P.H, P.W, P.out_hi, P.out_lo, P.out_ld, stream, P.bias, P.act);
  }
  if (e != cudaErrorNotSupported) return e;
  cudaGetLastError();
  if (npix == 64 * 128) e = launch_warp_t<64 * 128>(P, stream);          // 1024 x 2048 frames
  else if (npix == 32 * 64) e = launch_warp_t<32 * 64>(P, stream);       // 512 x 1024
  else e = launch_warp_t<0>(P, stream);
  if (e != cudaSuccess || !P.out_hi) return e;
  return launch_nchw_to_split(P.out_nchw, P.C, P.H, P.W, P.out_hi, P.out_lo, P.out_ld, stream, P.bias, P.act);
}

cudaError_t launch_nchw_to_split(const float* src, int C, int H, int W, __half* hi, __half* lo, int ld,
                                 cudaStream_t stream, const float* bias, int act) {
  dim3 grid((H * W + 31) / 32, (C + 63) / 64);
  return launch_k(nchw_to_split_kernel, grid, dim3(256), 0, stream, src, C, H * W, hi, lo, ld, bias, act);
}

// ------------------------------------------------------------------------------------------------
// Finalize-time weight composition (shortcut iii): a 1x1 convolution that directly follows a
// 4x4/s2 transposed convolution (accel_18.py:204-213: `18_feat_upsampling` -> `18_fc6`, no bias / activation in
// between) is one transposed convolution with weights  out[c][n][t] = sum_m conv1x1[n][m] * deconv[c][m][t].
// For a fixed input channel c the deconv slab [mid][16] is contiguous: block = (64 n) x (16 t) outputs of one c,
// thread = 4 n x 1 t, fp32 operands staged in shared memory, fp64 accumulation.  Runs once per handle.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fold_deconv_1x1_kernel(const float* __restrict__ deconv,
                                                              const float* __restrict__ conv1x1, float* __restrict__ out,
                                                              int mid, int cout) {
  __shared__ float sw[64][33];       // conv1x1[n0 + i][m0 + j]
  __shared__ float sd[32][16];       // deconv[c][m0 + j][t]
  const int c = blockIdx.y, n0 = blockIdx.x * 64;
  const int t = threadIdx.x & 15, ng = threadIdx.x >> 4;          // n = n0 + ng + 16 * k
  const float* dslab = deconv + (size_t)c * mid * 16;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int m0 = 0; m0 < mid; m0 += 32) {
    for (int i = threadIdx.x; i < 64 * 32; i += 256) {
      const int r = i >> 5, j = i & 31;
      sw[r][j] = (n0 + r < cout && m0 + j < mid) ? conv1x1[(size_t)(n0 + r) * mid + m0 + j] : 0.f;
    }
    for (int i = threadIdx.x; i < 32 * 16; i += 256)
      sd[i >> 4][i & 15] = (m0 + (i >> 4) < mid) ? dslab[(size_t)m0 * 16 + i] : 0.f;
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const double d = (double)sd[j][t];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fma((double)sw[ng + 16 * k][j], d, acc[k]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int n = n0 + ng + 16 * k;
    if (n < cout) out[((size_t)c * cout + n) * 16 + t] = (float)acc[k];
  }
}

cudaError_t launch_fold_deconv_1x1(const float* deconv, const float* conv1x1, float* out, int cin, int mid, int cout,
                                   cudaStream_t stream) {
  fold_deconv_1x1_kernel<<<dim3((cout + 63) / 64, cin), 256, 0, stream>>>(deconv, conv1x1, out, mid, cout);
  cudaError_t e = cudaGetLastError();
  return e != cudaSuccess ? e : cudaStreamSynchronize(stream);
}

cudaError_t launch_split_to_nchw(const __half* hi, const __half* lo, int ld, int C, int H, int W, float* dst,
                                 cudaStream_t stream) {
  dim3 grid((H * W + 31) / 32, (C + 31) / 32);
  return launch_k(split_to_nchw_kernel, grid, dim3(256), 0, stream, hi, lo, ld, C, H * W, dst);
}

cudaError_t launch_upflow(const UpflowParams& P, cudaStream_t stream) {
  const int work = 4 * P.H * P.W;
  return launch_k(upflow_kernel, dim3((work + 255) / 256), dim3(256), 0, stream, P);
}

cudaError_t launch_fuse_lowres(const FuseParams& P, cudaStream_t stream) {
  if (P.K > 32) return cudaErrorInvalidValue;
  const int npix = P.h * P.w_;
  return launch_k(fuse_lowres_kernel, dim3((npix + 127) / 128), dim3(128), (size_t)P.K * 2 * P.K * sizeof(float), stream, P);
}

bool tail_band_supported(int K, int factor) {
  const char* e = getenv("ACCEL_TAIL_BAND");
  return !(e && e[0] == '0') && K == 19 && factor == 16;
}

cudaError_t launch_tail(const TailParams& P, cudaStream_t stream) {
  const int work = (P.w * P.factor / 4) * (P.h * P.factor);
  static int minb = -1, band = -1;
  if (minb < 0) { const char* e = getenv("ACCEL_TAIL_MINB"); minb = e && *e ? atoi(e) : 4; }
  if (band < 0) { const char* e = getenv("ACCEL_TAIL_BAND"); band = !(e && e[0] == '0'); }
  if (band && P.K == 19 && P.factor == 16) {
    const dim3 grid((P.w * 16 + 511) / 512, P.h);
    if (P.fuse_w)
      return P.score_out ? launch_k(tail_band_kernel<19, true, true>, grid, dim3(256), 0, stream, P)
                         : launch_k(tail_band_kernel<19, false, true>, grid, dim3(256), 0, stream, P);
    return P.score_out ? launch_k(tail_band_kernel<19, true, false>, grid, dim3(256), 0, stream, P)
                       : launch_k(tail_band_kernel<19, false, false>, grid, dim3(256), 0, stream, P);
  }
  if (P.fuse_w) return cudaErrorInvalidValue;              // only the band kernel evaluates the fusion itself
  if (P.K == 19 && minb == 4) return launch_k(tail_kernel<19, 4>, dim3((work + 255) / 256), dim3(256), 0, stream, P);
  if (P.K == 19 && minb == 6) return launch_k(tail_kernel<19, 6>, dim3((work + 255) / 256), dim3(256), 0, stream, P);
  if (P.K == 19) return launch_k(tail_kernel<19>, dim3((work + 255) / 256), dim3(256), 0, stream, P);
  return launch_k(tail_kernel<0>, dim3((work + 255) / 256), dim3(256), 0, stream, P);
}

}  // namespace accel
