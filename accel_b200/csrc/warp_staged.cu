// Flow-guided warp with shared-memory staging of the source rows (sm_100a).
//
//   GridGenerator(transform_type='warp') + BilinearSampler  (dff_deeplab/symbols/accel_18.py:174-175)
//   out[c, y, x] = sum over the 2x2 neighbourhood of  w * feat_key[c, y + dy, x + dx],  zeros outside
//
// The plain gather kernel (kernels_basic.cu) fetches each source value ~2.4 times through L2 (four taps, L1 hit
// rate ~50 %), which makes it L2-fabric bound at ~0.55 of the HBM roofline.  Here a CTA owns TH full output rows
// and a strided set of 8-channel batches.  The flow field over its rows fixes, once, the band of source rows
// [r0, r0 + R) that its 2x2 neighbourhoods touch; for every channel that band is ONE contiguous run of R*W floats
// in NCHW memory, so it is fetched with one `cp.async.bulk` (TMA 1-D bulk copy) into a 3-stage shared-memory ring,
// each source byte crossing L2->SM once per row band.  The four taps are then gathered from shared memory and the
// NCHW rows are written coalesced.  Per-pixel tap offsets and bilinear weights are computed once per CTA and kept
// in registers for all channels.  Arithmetic and operation order are those of the gather kernel (and the oracle):
// the result is bit-identical.  When the band does not fit (wild flow fields), the CTA gathers from global memory.
#include <string.h>

#include <algorithm>

#include <mutex>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace accel {

namespace {

using namespace tc;

constexpr int WS_STAGES = 3;
constexpr int WS_TILE = 1024;     // output pixels per CTA: TH * W <= WS_THREADS * WS_PPT = 1024

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

struct Tap {
  int o00;        // offset of the (y0, x0) tap: (yi0 - r0) * W + xi0 when staged, yi0 * W + xi0 otherwise
  int dx1, dyw;   // xi1 - xi0 (0 / 1), (yi1 - yi0) * W (0 / W)
  float w00, w01, w10, w11;
};

// WS_CB = channels per pipeline stage
template <int WS_THREADS, int WS_PPT, int WS_CB>
__global__ void __launch_bounds__(WS_THREADS) warp_kernel_staged(const float* __restrict__ feat, const float* __restrict__ flow,
                                                                 float* __restrict__ out, int C, int H, int W, int TH, int RMAX) {
  extern __shared__ __align__(128) float ring[];            // [WS_STAGES][WS_CB][RMAX * W]
  __shared__ __align__(8) uint64_t bars[WS_STAGES];
  __shared__ int s_min[WS_THREADS / 32], s_max[WS_THREADS / 32];

  pdl_trigger();
  pdl_wait();                                               // the flow field is the previous kernel's output
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int npix = H * W;
  const int y0 = blockIdx.x * TH;
  const int rows = min(TH, H - y0);
  const int tile = rows * W;

  // ---- per-pixel sampling positions (same fp32 operation sequence as MXNet: normalise to [-1,1], de-normalise)
  Tap tap[WS_PPT];
  int ymin = H, ymax = -1;
  const float sx = (float)(W - 1) / 2.0f, sy = (float)(H - 1) / 2.0f;
#pragma unroll
  for (int k = 0; k < WS_PPT; ++k) {
    const int q = tid + k * WS_THREADS;
    Tap t;
    t.o00 = 0; t.dx1 = 0; t.dyw = 0; t.w00 = t.w01 = t.w10 = t.w11 = kSkipTap;
    if (q < tile) {
      const int y = y0 + q / W, x = q - (q / W) * W;
      const int p = y * W + x;
      const float gx = __fsub_rn(__fdiv_rn(__fadd_rn(__ldg(flow + p), (float)x), sx), 1.0f);
      const float gy = __fsub_rn(__fdiv_rn(__fadd_rn(__ldg(flow + npix + p), (float)y), sy), 1.0f);
      const float xr = __fmul_rn(__fadd_rn(gx, 1.0f), sx);
      const float yr = __fmul_rn(__fadd_rn(gy, 1.0f), sy);
      const float xf = floorf(xr), yf = floorf(yr);
      const float wx0 = __fsub_rn(1.0f, __fsub_rn(xr, xf)), wy0 = __fsub_rn(1.0f, __fsub_rn(yr, yf));
      const float wx1 = __fsub_rn(1.0f, wx0), wy1 = __fsub_rn(1.0f, wy0);
      const bool x0ok = xf >= 0.f && xf <= (float)(W - 1), x1ok = xf + 1.f >= 0.f && xf + 1.f <= (float)(W - 1);
      const bool y0ok = yf >= 0.f && yf <= (float)(H - 1), y1ok = yf + 1.f >= 0.f && yf + 1.f <= (float)(H - 1);
      const bool sane = fabsf(xr) < 1e9f && fabsf(yr) < 1e9f;       // keep the int conversion defined for wild flows
      int xi0 = 0, xi1 = 0, yi0 = y, yi1 = y;                        // dead taps (all weights 0) point at the pixel itself
      if (sane && (x0ok || x1ok) && (y0ok || y1ok)) {
        xi0 = min(max((int)xf, 0), W - 1); xi1 = min(max((int)xf + 1, 0), W - 1);
        yi0 = min(max((int)yf, 0), H - 1); yi1 = min(max((int)yf + 1, 0), H - 1);
        t.w00 = (y0ok && x0ok) ? __fmul_rn(wy0, wx0) : kSkipTap;
        t.w01 = (y0ok && x1ok) ? __fmul_rn(wy0, wx1) : kSkipTap;
        t.w10 = (y1ok && x0ok) ? __fmul_rn(wy1, wx0) : kSkipTap;
        t.w11 = (y1ok && x1ok) ? __fmul_rn(wy1, wx1) : kSkipTap;
      }
      t.o00 = yi0 * W + xi0; t.dx1 = xi1 - xi0; t.dyw = (yi1 - yi0) * W;
      ymin = min(ymin, yi0); ymax = max(ymax, yi1);
    }
    tap[k] = t;
  }
  // ---- band of source rows this CTA touches
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if (lane == 0) { s_min[wid] = ymin; s_max[wid] = ymax; }
  if (tid == 0) {
    for (int s = 0; s < WS_STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < WS_THREADS / 32; ++i) { ymin = min(ymin, s_min[i]); ymax = max(ymax, s_max[i]); }
  const int r0 = ymin, R = ymax - ymin + 1;
  const int nbatch = (C + WS_CB - 1) / WS_CB;

  if (R <= RMAX) {
    // ================================ staged path ================================
    const int band = R * W;                                   // floats per channel in the ring
    const uint32_t band_bytes = (uint32_t)band * 4u;
    const int slot = RMAX * W;
#pragma unroll
    for (int k = 0; k < WS_PPT; ++k) tap[k].o00 -= r0 * W;
    const uint32_t ring0 = smem_u32(ring), bar0 = smem_u32(&bars[0]);
    auto issue = [&](int bt, int s) {                          // one thread: WS_CB bulk copies -> stage s
      const int c0 = bt * WS_CB;
      const int nc = min(WS_CB, C - c0);
      const uint32_t bar = bar0 + 8u * s;
      mbar_arrive_expect_tx(bar, band_bytes * (uint32_t)nc);
      for (int c = 0; c < nc; ++c)
        bulk_load(ring0 + (uint32_t)((s * WS_CB + c) * slot) * 4u, feat + (size_t)(c0 + c) * npix + (size_t)r0 * W, band_bytes, bar);
    };
    int issued = 0;                                            // batches issued so far (this CTA's sequence)
    if (tid == 0)
      for (int bt = blockIdx.y; bt < nbatch && issued < WS_STAGES - 1; bt += gridDim.y, ++issued) issue(bt, issued);
    int it = 0;
    for (int bt = blockIdx.y; bt < nbatch; bt += gridDim.y, ++it) {
      const int s = it % WS_STAGES;
      if (tid == 0) {                                          // refill the stage everyone left at the end of iteration it-1
        const int nb = bt + (WS_STAGES - 1) * gridDim.y;
        if (nb < nbatch) issue(nb, (it + WS_STAGES - 1) % WS_STAGES);
      }
      mbar_wait(bar0 + 8u * s, (uint32_t)((it / WS_STAGES) & 1));
      const int c0 = bt * WS_CB;
      const int nc = min(WS_CB, C - c0);
      const float* st = ring + (size_t)s * WS_CB * slot;
#pragma unroll
      for (int k = 0; k < WS_PPT; ++k) {
        const int q = tid + k * WS_THREADS;
        if (q < tile) {
          const Tap t = tap[k];
          float* po = out + (size_t)c0 * npix + (size_t)y0 * W + q;
          if (nc == WS_CB) {
            float a[WS_CB], b[WS_CB], c[WS_CB], d[WS_CB];
#pragma unroll
            for (int ch = 0; ch < WS_CB; ++ch) {
              const float* sp = st + ch * slot + t.o00;
              a[ch] = sp[0]; b[ch] = sp[t.dx1]; c[ch] = sp[t.dyw]; d[ch] = sp[t.dyw + t.dx1];
            }
#pragma unroll
            for (int ch = 0; ch < WS_CB; ++ch) {
              float v = __fmul_rn(keep_tap(t.w00) ? a[ch] : 0.f, t.w00);      // zero-weight (out-of-range) taps: skipped by value
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w01) ? b[ch] : 0.f, t.w01));
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w10) ? c[ch] : 0.f, t.w10));
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w11) ? d[ch] : 0.f, t.w11));
              po[(size_t)ch * npix] = v;
            }
          } else {
            for (int ch = 0; ch < nc; ++ch) {
              const float* sp = st + ch * slot + t.o00;
              float v = __fmul_rn(keep_tap(t.w00) ? sp[0] : 0.f, t.w00);
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w01) ? sp[t.dx1] : 0.f, t.w01));
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w10) ? sp[t.dyw] : 0.f, t.w10));
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w11) ? sp[t.dyw + t.dx1] : 0.f, t.w11));
              po[(size_t)ch * npix] = v;
            }
          }
        }
      }
      __syncthreads();                                         // stage s is free for the refill of iteration it+1
    }
  } else {
    // ================================ gather path (band too tall) ================================
    for (int bt = blockIdx.y; bt < nbatch; bt += gridDim.y) {
      const int c0 = bt * WS_CB;
      const int nc = min(WS_CB, C - c0);
#pragma unroll
      for (int k = 0; k < WS_PPT; ++k) {
        const int q = tid + k * WS_THREADS;
        if (q < tile) {
          const Tap t = tap[k];
          float* po = out + (size_t)c0 * npix + (size_t)y0 * W + q;
          const float* sp0 = feat + (size_t)c0 * npix + t.o00;
          if (nc == WS_CB) {
            float a[WS_CB], b[WS_CB], c[WS_CB], d[WS_CB];
#pragma unroll
            for (int ch = 0; ch < WS_CB; ++ch) {
              const float* sp = sp0 + (size_t)ch * npix;
              a[ch] = __ldg(sp); b[ch] = __ldg(sp + t.dx1); c[ch] = __ldg(sp + t.dyw); d[ch] = __ldg(sp + t.dyw + t.dx1);
            }
#pragma unroll
            for (int ch = 0; ch < WS_CB; ++ch) {
              float v = __fmul_rn(keep_tap(t.w00) ? a[ch] : 0.f, t.w00);      // zero-weight (out-of-range) taps: skipped by value
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w01) ? b[ch] : 0.f, t.w01));
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w10) ? c[ch] : 0.f, t.w10));
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w11) ? d[ch] : 0.f, t.w11));
              po[(size_t)ch * npix] = v;
            }
          } else {
            for (int ch = 0; ch < nc; ++ch) {
              const float* sp = sp0 + (size_t)ch * npix;
              float v = __fmul_rn(keep_tap(t.w00) ? __ldg(sp) : 0.f, t.w00);
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w01) ? __ldg(sp + t.dx1) : 0.f, t.w01));
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w10) ? __ldg(sp + t.dyw) : 0.f, t.w10));
              v = __fadd_rn(v, __fmul_rn(keep_tap(t.w11) ? __ldg(sp + t.dyw + t.dx1) : 0.f, t.w11));
              po[(size_t)ch * npix] = v;
            }
          }
        }
      }
    }
  }
}

// ================================================================================================
// Fused warp: ONE pass over the source writes both consumers' copies -- the fp32 NCHW `warping_feat_output` the
// chained schedule carries to the next frame, and the split-fp16 NHWC tensor the first conv of the fusion head (`fc6`,
// or Accel-101's `corr`) loads through TMA -- so the warped feature is never re-read to change its layout
// (nchw_to_split_kernel: 64 MiB read + 64 MiB written per cur frame) and `fc6` finds its operand already in place.
//   * CTA = WF_CONS output pixels (TH full rows) x a strided set of 32-channel groups.
//   * A dedicated producer warp streams the CTA's source row band, 4 channels per stage, into a 4-stage ring with
//     `cp.async.bulk`; per-stage full / empty mbarriers -- no CTA-wide barrier inside a group.
//   * Each consumer thread owns one pixel: taps and weights in registers, four channels per stage; the fp32 values
//     leave as coalesced NCHW rows, their hi/lo fp16 split (+ bias / activation for the commuted L head) is kept in
//     registers for the 8 stages of a group and then written as one 64-byte row of a SWIZZLE_64B staging tile that
//     two TMA tensor stores (hi, lo) move to the NHWC tensor.
// Same fp32 operation order as the other warp kernels and the oracle: bit-identical `warping_feat_output`; the split
// copy equals nchw_to_split_kernel's (same split_pair on the same channel pairs).  Out-of-range taps are skipped by
// VALUE (a zero is blended instead of the clamped neighbour), as MXNet's BilinearSampler does: an Inf next to the
// border cannot turn into NaN.
// ================================================================================================
constexpr int WF_CB = 4, WF_G = 32;           // channels per ring stage / per staged NHWC group
constexpr int WF_NMAP = 12;                   // source tensor maps: box {W, R, WF_CB} for band heights R = 1 .. WF_NMAP

struct alignas(64) WarpFusedParams {
  CUtensorMap o_hi, o_lo;                    // split NHWC view {C, W, H}, box {32, W, TH}, SWIZZLE_64B
  CUtensorMap src[WF_NMAP];                  // fp32 NCHW source {W, H, C}: src[R-1] has box {W, R, WF_CB} -- a whole stage in ONE request
  const float* feat;
  const float* flow;
  float* out_nchw;
  const float* bias;
  int act, has_split;
  int C, H, W, TH, RMAX;
  int debug;                                 // ACCEL_WARP_FUSED_DEBUG (timing decomposition only): 1 no NCHW stores, 2 no TMA stores
  int nmap;                                  // bands up to this height use src[]; taller ones one cp.async.bulk per channel
  int ring_floats, max_stages;               // ring capacity; the CTA cuts it into min(max_stages, capacity / band) stages
};

// WF_CONS = consumer threads = pixels per CTA tile (the producer warp comes on top); WF_STAGES = deepest ring (mbarrier pairs):
// the ring is cut into stages of the band the CTA's flow vectors actually span (R rows, known after the tap set-up), not of
// the worst case RMAX, so that a smooth field keeps twice the bytes in flight
template <int WF_CONS, int WF_STAGES>
__global__ void __launch_bounds__(WF_CONS + 32) warp_kernel_fused(const __grid_constant__ WarpFusedParams P) {
  extern __shared__ __align__(1024) uint8_t wf_smem[];          // [staging hi 16 KB | lo 16 KB][ring]
  __shared__ __align__(8) uint64_t wf_bars[2 * WF_STAGES];
  __shared__ int s_min[WF_CONS / 32], s_max[WF_CONS / 32];

  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool producer = wid == WF_CONS / 32;
  const int H = P.H, W = P.W, C = P.C, npix = H * W;
  const int y0 = blockIdx.x * P.TH;
  const int rows = min(P.TH, H - y0);
  const int tile = rows * W;
  const uint32_t smem0 = (smem_u32(wf_smem) + 1023u) & ~1023u;
  const uint32_t stg_hi = smem0, stg_lo = smem0 + WF_CONS * 64u, ring0 = smem0 + 2u * WF_CONS * 64u;
  const float* ring = reinterpret_cast<const float*>(wf_smem + (ring0 - smem_u32(wf_smem)));
  const uint32_t full0 = smem_u32(&wf_bars[0]), empty0 = smem_u32(&wf_bars[WF_STAGES]);

  Tap t;
  t.o00 = 0; t.dx1 = 0; t.dyw = 0; t.w00 = t.w01 = t.w10 = t.w11 = kSkipTap;
  int ymin = H, ymax = -1;
  const bool live = !producer && tid < tile;
  if (live) {
    const float sx = (float)(W - 1) / 2.0f, sy = (float)(H - 1) / 2.0f;
    const int y = y0 + tid / W, x = tid - (tid / W) * W;
    const int p = y * W + x;
    const float gx = __fsub_rn(__fdiv_rn(__fadd_rn(__ldg(P.flow + p), (float)x), sx), 1.0f);
    const float gy = __fsub_rn(__fdiv_rn(__fadd_rn(__ldg(P.flow + npix + p), (float)y), sy), 1.0f);
    const float xr = __fmul_rn(__fadd_rn(gx, 1.0f), sx);
    const float yr = __fmul_rn(__fadd_rn(gy, 1.0f), sy);
    const float xf = floorf(xr), yf = floorf(yr);
    const float wx0 = __fsub_rn(1.0f, __fsub_rn(xr, xf)), wy0 = __fsub_rn(1.0f, __fsub_rn(yr, yf));
    const float wx1 = __fsub_rn(1.0f, wx0), wy1 = __fsub_rn(1.0f, wy0);
    const bool x0ok = xf >= 0.f && xf <= (float)(W - 1), x1ok = xf + 1.f >= 0.f && xf + 1.f <= (float)(W - 1);
    const bool y0ok = yf >= 0.f && yf <= (float)(H - 1), y1ok = yf + 1.f >= 0.f && yf + 1.f <= (float)(H - 1);
    const bool sane = fabsf(xr) < 1e9f && fabsf(yr) < 1e9f;
    int xi0 = 0, xi1 = 0, yi0 = y, yi1 = y;
    if (sane && (x0ok || x1ok) && (y0ok || y1ok)) {
      xi0 = min(max((int)xf, 0), W - 1); xi1 = min(max((int)xf + 1, 0), W - 1);
      yi0 = min(max((int)yf, 0), H - 1); yi1 = min(max((int)yf + 1, 0), H - 1);
      t.w00 = (y0ok && x0ok) ? __fmul_rn(wy0, wx0) : kSkipTap;
      t.w01 = (y0ok && x1ok) ? __fmul_rn(wy0, wx1) : kSkipTap;
      t.w10 = (y1ok && x0ok) ? __fmul_rn(wy1, wx0) : kSkipTap;
      t.w11 = (y1ok && x1ok) ? __fmul_rn(wy1, wx1) : kSkipTap;
    }
    t.o00 = yi0 * W + xi0; t.dx1 = xi1 - xi0; t.dyw = (yi1 - yi0) * W;
    ymin = yi0; ymax = yi1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if (!producer && lane == 0) { s_min[wid] = ymin; s_max[wid] = ymax; }
  if (tid == 0) {
    for (int s = 0; s < WF_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, WF_CONS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  ymin = H; ymax = -1;
#pragma unroll
  for (int i = 0; i < WF_CONS / 32; ++i) { ymin = min(ymin, s_min[i]); ymax = max(ymax, s_max[i]); }
  const int r0 = ymin, R = max(ymax - ymin + 1, 1);
  const bool staged = R <= P.RMAX;
  const int ngroups = C / WF_G;
  const int slot = R * W;                                       // floats per channel slot of the ring
  const int nst = staged ? min(min(WF_STAGES, P.max_stages), P.ring_floats / (WF_CB * slot)) : 1;
  const uint32_t band_bytes = (uint32_t)(R * W) * 4u;

  if (producer) {
    if (staged && lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int g = blockIdx.y; g < ngroups; g += gridDim.y)
        for (int b = 0; b < WF_G / WF_CB; ++b) {
          const int c0 = g * WF_G + b * WF_CB;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t fb = full0 + 8 * s;
          mbar_arrive_expect_tx(fb, band_bytes * WF_CB);
          if (R <= P.nmap) {                                    // one tensor request per stage: {W, R rows, WF_CB channels}
            tma_load_3d(ring0 + (uint32_t)(s * WF_CB * slot) * 4u, &P.src[R - 1], fb, 0, r0, c0);
          } else {
#pragma unroll
            for (int c = 0; c < WF_CB; ++c)
              bulk_load(ring0 + (uint32_t)((s * WF_CB + c) * slot) * 4u, P.feat + (size_t)(c0 + c) * npix + (size_t)r0 * W, band_bytes, fb);
          }
          if (++s == nst) { s = 0; ph ^= 1; }
        }
    }
    return;
  }

  // ---------------------------------------------------------------- consumers
  if (staged) t.o00 -= r0 * W;
  const bool n00 = keep_tap(t.w00), n01 = keep_tap(t.w01), n10 = keep_tap(t.w10), n11 = keep_tap(t.w11);
  const int o01 = t.o00 + t.dx1, o10 = t.o00 + t.dyw, o11 = t.o00 + t.dyw + t.dx1;
  const bool fast = staged && __all_sync(0xffffffffu, live && n00 && n01 && n10 && n11);
  int s = 0;
  uint32_t ph = 0;
  for (int g = blockIdx.y; g < ngroups; g += gridDim.y) {
    uint32_t wh[16], wl[16];
#pragma unroll
    for (int b = 0; b < WF_G / WF_CB; ++b) {
      const int c0 = g * WF_G + b * WF_CB;
      float v[WF_CB];
      if (staged) {
        mbar_wait(full0 + 8 * s, ph);
        if (fast) {                                             // every tap of every lane is in range: no predicates, no zero fills
          const float* st = ring + (size_t)s * WF_CB * slot;    // warp-uniform: the four tap offsets stay in registers
#pragma unroll
          for (int ch = 0; ch < WF_CB; ++ch) {
            const float* sp = st + ch * slot;
            float x = __fmul_rn(sp[t.o00], t.w00);
            x = __fadd_rn(x, __fmul_rn(sp[o01], t.w01));
            x = __fadd_rn(x, __fmul_rn(sp[o10], t.w10));
            v[ch] = __fadd_rn(x, __fmul_rn(sp[o11], t.w11));
          }
        } else {
          const float* st = ring + (size_t)s * WF_CB * slot + t.o00;
#pragma unroll
          for (int ch = 0; ch < WF_CB; ++ch) {
            const float* sp = st + ch * slot;
            const float a = n00 ? sp[0] : 0.f, bq = n01 ? sp[t.dx1] : 0.f, cq = n10 ? sp[t.dyw] : 0.f, d = n11 ? sp[t.dyw + t.dx1] : 0.f;
            float x = __fmul_rn(a, t.w00);
            x = __fadd_rn(x, __fmul_rn(bq, t.w01));
            x = __fadd_rn(x, __fmul_rn(cq, t.w10));
            v[ch] = __fadd_rn(x, __fmul_rn(d, t.w11));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if (++s == nst) { s = 0; ph ^= 1; }
      } else {
        const float* sp0 = P.feat + (size_t)c0 * npix + t.o00;
#pragma unroll
        for (int ch = 0; ch < WF_CB; ++ch) {
          const float* sp = sp0 + (size_t)ch * npix;
          const float a = (live && n00) ? __ldg(sp) : 0.f, bq = (live && n01) ? __ldg(sp + t.dx1) : 0.f;
          const float cq = (live && n10) ? __ldg(sp + t.dyw) : 0.f, d = (live && n11) ? __ldg(sp + t.dyw + t.dx1) : 0.f;
          float x = __fmul_rn(a, t.w00);
          x = __fadd_rn(x, __fmul_rn(bq, t.w01));
          x = __fadd_rn(x, __fmul_rn(cq, t.w10));
          v[ch] = __fadd_rn(x, __fmul_rn(d, t.w11));
        }
      }
      if (live && P.out_nchw && !(P.debug & 1)) {
        float* po = P.out_nchw + (size_t)c0 * npix + (size_t)y0 * W + tid;
#pragma unroll
        for (int ch = 0; ch < WF_CB; ++ch) po[(size_t)ch * npix] = v[ch];
      }
      if (P.has_split) {
        if (P.bias) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(P.bias + c0));
          v[0] = apply_act(v[0] + bb.x, P.act); v[1] = apply_act(v[1] + bb.y, P.act);
          v[2] = apply_act(v[2] + bb.z, P.act); v[3] = apply_act(v[3] + bb.w, P.act);
        }
        split_pair(v[0], v[1], wh[2 * b], wl[2 * b]);
        split_pair(v[2], v[3], wh[2 * b + 1], wl[2 * b + 1]);
      }
    }
    if (P.has_split) {
      if (tid == 0) bulk_wait_read0();                          // the previous group's stores are done with the tile
      asm volatile("bar.sync 1, %0;" ::"n"(WF_CONS) : "memory");
      stage_row64(stg_hi, tid, wh);
      stage_row64(stg_lo, tid, wl);
      fence_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(WF_CONS) : "memory");
      if (tid == 0 && !(P.debug & 2)) {
        tma_store_3d(&P.o_hi, stg_hi, g * WF_G, 0, y0);       // rows past the last image row are clipped by the map
        tma_store_3d(&P.o_lo, stg_lo, g * WF_G, 0, y0);
        bulk_commit();
      }
    }
  }
  if (tid == 0 && P.has_split) bulk_wait0();
}

}  // namespace

static int wf_cons() {
  static const int v = [] { const int c = env_int("ACCEL_WARP_FUSED_CONS", 256); return c == 512 ? 512 : 256; }();
  return v;
}

bool warp_fused_supported(const WarpParams& P) {
  const int W = P.W, cons = wf_cons();
  if (env_int("ACCEL_WARP_FUSED", 1) == 0) return false;
  if (W < 8 || W > cons || (cons % W) != 0 || (W & 3)) return false;              // TH full rows = `cons` pixels
  if (P.C % WF_G) return false;
  if (((uintptr_t)P.feat & 15) != 0) return false;
  if (P.out_hi && (((uintptr_t)P.out_hi & 15) || ((uintptr_t)P.out_lo & 15) || (P.out_ld % 8))) return false;
  if (P.bias && ((uintptr_t)P.bias & 15)) return false;
  return P.out_hi != nullptr;          // fp32-NCHW-only launches stay on warp_kernel_staged (measured faster there)
}

template <int CONS, int STAGES>
static cudaError_t launch_fused_variant(const WarpParams& P, int once_slot, cudaStream_t stream) {
  int dev = 0, sms = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (first_time_on_device(once_slot)) {
    if (cudaFuncSetAttribute(warp_kernel_fused<CONS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 2048) != cudaSuccess)
      return cudaErrorNotSupported;
  }
  WarpFusedParams F;
  memset(&F, 0, sizeof(F));
  F.feat = P.feat; F.flow = P.flow; F.out_nchw = P.out_nchw; F.bias = P.bias; F.act = P.act;
  F.C = P.C; F.H = P.H; F.W = P.W;
  F.TH = CONS / P.W;
  if (F.TH > P.H) F.TH = P.H;
  F.has_split = P.out_hi ? 1 : 0;
  // `per_sm` CTAs per SM: staging (CONS x 64 B x 2 planes) + the ring; a band may take up to half the ring (two stages)
  const int per_sm = std::min(std::max(env_int("ACCEL_WARP_FUSED_PER_SM", 2), 1), 4);
  const int budget = max_smem / per_sm - 4096 - 2 * CONS * 64;
  if (budget < 2 * WF_CB * P.W * 4 * (F.TH + 2)) return cudaErrorNotSupported;
  F.ring_floats = budget / 4;
  F.debug = env_int("ACCEL_WARP_FUSED_DEBUG", 0);
  F.max_stages = std::min(std::max(env_int("ACCEL_WARP_FUSED_NST", STAGES), 2), STAGES);
  int rmax = F.ring_floats / (2 * WF_CB * P.W);
  const int cap = env_int("ACCEL_WARP_RMAX", 0);
  if (cap > 0 && rmax > cap) rmax = cap;
  if (rmax > P.H) rmax = P.H;
  if (rmax < F.TH + 2 && rmax < P.H) return cudaErrorNotSupported;
  F.RMAX = rmax;
  if (P.out_hi) {
    char err[256];
    const cuuint64_t e = sizeof(__half);
    cuuint64_t dims[3] = {(cuuint64_t)P.C, (cuuint64_t)P.W, (cuuint64_t)P.H};
    cuuint64_t str[2] = {(cuuint64_t)P.out_ld * e, (cuuint64_t)P.out_ld * P.W * e};
    cuuint32_t box[3] = {WF_G, (cuuint32_t)P.W, (cuuint32_t)F.TH};
    if (!encode(&F.o_hi, P.out_hi, 3, dims, str, box, err, sizeof(err), CU_TENSOR_MAP_SWIZZLE_64B) ||
        !encode(&F.o_lo, P.out_lo, 3, dims, str, box, err, sizeof(err), CU_TENSOR_MAP_SWIZZLE_64B))
      return cudaErrorInvalidValue;
  }
  F.nmap = 0;
  if (env_int("ACCEL_WARP_FUSED_TMAP", 1) != 0 && P.W <= 256) {
    char err[256];
    const cuuint64_t dims[3] = {(cuuint64_t)P.W, (cuuint64_t)P.H, (cuuint64_t)P.C};
    const cuuint64_t str[2] = {(cuuint64_t)P.W * 4, (cuuint64_t)P.W * P.H * 4};
    const int nmap = std::min(std::min(WF_NMAP, rmax), P.H);
    int ok = 1;
    for (int r = 1; r <= nmap && ok; ++r) {
      const cuuint32_t box[3] = {(cuuint32_t)P.W, (cuuint32_t)r, (cuuint32_t)WF_CB};
      ok = encode(&F.src[r - 1], P.feat, 3, dims, str, box, err, sizeof(err), CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    }
    if (ok) F.nmap = nmap;
  }
  const size_t smem = 1024 + 2 * (size_t)CONS * 64 + (size_t)F.ring_floats * 4;
  const int bx = (P.H + F.TH - 1) / F.TH;
  const int ngroups = P.C / WF_G;
  int by = (sms * per_sm) / bx;
  if (by < 1) by = 1;
  if (by > ngroups) by = ngroups;
  const int per = (ngroups + by - 1) / by;                       // equal number of groups per CTA: no straggler wave
  by = (ngroups + per - 1) / per;
  {
    const int force = env_int("ACCEL_WARP_FUSED_BY", 0);         // tuning aid: CTAs per row tile, unequal group counts allowed
    if (force > 0 && force <= ngroups) by = force;
  }
  return launch_k(warp_kernel_fused<CONS, STAGES>, dim3(bx, by), dim3(CONS + 32), smem, stream, F);
}

cudaError_t launch_warp_fused(const WarpParams& P, cudaStream_t stream) {
  if (wf_cons() == 512) return launch_fused_variant<512, 16>(P, ONCE_WARP_FUSED_1, stream);
  return launch_fused_variant<256, 16>(P, ONCE_WARP_FUSED, stream);
}

template <int THREADS, int PPT, int CB>
static cudaError_t launch_variant(const WarpParams& P, int th, int max_smem, int sms, int stage_cap, cudaStream_t stream) {
  if (first_time_on_device(ONCE_WARP_STAGED_0 + (THREADS == 256 ? 0 : 2) + (CB == 8 ? 0 : 1))) {
    if (cudaFuncSetAttribute(warp_kernel_staged<THREADS, PPT, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             max_smem - 2048) != cudaSuccess)
      return cudaErrorNotSupported;
  }
  const int W = P.W, H = P.H;
  // ring sized for `per_sm` resident CTAs: CB = 4 -> two CTAs per SM (one CTA's barrier bubbles hide behind the other)
  const int per_sm = CB <= 4 ? 2 : 1;
  const int budget = max_smem / per_sm - 4096;
  int rmax = budget / (WS_STAGES * CB * W * 4);
  if (rmax > H) rmax = H;
  if (stage_cap > 0 && rmax > stage_cap) rmax = stage_cap;
  if (rmax < th + 2 && rmax < H) return cudaErrorNotSupported;
  const size_t smem = (size_t)WS_STAGES * CB * rmax * W * 4;
  const int bx = (H + th - 1) / th;
  const int nbatch = (P.C + CB - 1) / CB;
  int by = (sms * per_sm) / bx;
  if (by < 1) by = 1;
  if (by > nbatch) by = nbatch;
  return launch_k(warp_kernel_staged<THREADS, PPT, CB>, dim3(bx, by), dim3(THREADS), smem, stream, P.feat, P.flow, P.out_nchw, P.C,
                  H, W, th, rmax);
}

// Returns cudaErrorNotSupported when the shape does not suit the staged kernel (caller falls back to the gather kernel).
cudaError_t launch_warp_staged(const WarpParams& P, cudaStream_t stream) {
  const int W = P.W, H = P.H;
  if (W < 8 || (W & 3) || W > WS_TILE) return cudaErrorNotSupported;
  if (((uintptr_t)P.feat & 15) != 0) return cudaErrorNotSupported;
  // per-device launch configuration, filled once per device under a lock (several GPUs may be driven from the threads
  // of one process: loader.pred_eval_multiprocess)
  struct DevCfg { int max_smem = 0, sms = 0; };
  static std::mutex mu;
  static DevCfg cfgs[64];
  static int threads = 512, cb = 4, stage_cap = 0, th_cap = 0;
  static bool env_read = false;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorNotSupported;
  int max_smem, sms;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!env_read) {
      threads = env_int("ACCEL_WARP_THREADS", 512);          // tuning aids (tools/bench_warp.py)
      cb = env_int("ACCEL_WARP_CB", 4);
      stage_cap = env_int("ACCEL_WARP_RMAX", 0);
      th_cap = env_int("ACCEL_WARP_TH", 0);
      env_read = true;
    }
    DevCfg& c = cfgs[dev];
    if (!c.max_smem) {
      cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&c.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    max_smem = c.max_smem;
    sms = c.sms;
  }
  int th = WS_TILE / W;
  if (th > H) th = H;
  if (th_cap > 0 && th > th_cap) th = th_cap;
  if (threads == 256) {
    if (cb == 8) return launch_variant<256, 4, 8>(P, th, max_smem, sms, stage_cap, stream);
    return launch_variant<256, 4, 4>(P, th, max_smem, sms, stage_cap, stream);
  }
  if (cb == 8) return launch_variant<512, 2, 8>(P, th, max_smem, sms, stage_cap, stream);
  return launch_variant<512, 2, 4>(P, th, max_smem, sms, stage_cap, stream);
}

}  // namespace accel
