// Frame ingest and metric kernels either side of the per-frame graphs:
//   * preprocess_kernel : lib/utils/image.py:224-235 `transform` -- uint8 BGR HWC -> fp32 RGB NCHW minus
//     PIXEL_MEANS -- so a frame crosses PCIe as 3 bytes per pixel instead of 12.
//   * confusion_kernel  : dff_deeplab/demo.py:50-53 `fast_hist` -- the num_classes^2 confusion counts of a
//     label map against its ground truth, accumulated on the device (int64), read back once per video.
// Both are plain HBM streaming kernels; sm_100a only.
#include "common.cuh"
#include "kernels.h"

namespace accel {

namespace {

// One thread = 4 horizontally adjacent pixels: 12 source bytes as three aligned 32-bit loads, one float4
// store per colour plane (coalesced 512 B per warp and plane).  The subtraction is done in double and
// rounded once to fp32, exactly like numpy's `im[:, :, 2 - i] - pixel_means[2 - i]` (float64) followed by
// mx.nd.array's float32 conversion (demo.py:185).
__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ src, int npix4, size_t plane,
                                                         double mean_b, double mean_g, double mean_r,
                                                         float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix4) return;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src) + (size_t)i * 3;
  const uint32_t w0 = __ldg(s), w1 = __ldg(s + 1), w2 = __ldg(s + 2);
  // bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
  const uint32_t b[4] = {w0 & 255u, (w0 >> 24) & 255u, (w1 >> 16) & 255u, (w2 >> 8) & 255u};
  const uint32_t g[4] = {(w0 >> 8) & 255u, w1 & 255u, (w1 >> 24) & 255u, (w2 >> 16) & 255u};
  const uint32_t r[4] = {(w0 >> 16) & 255u, (w1 >> 8) & 255u, w2 & 255u, (w2 >> 24) & 255u};
  float4 R, G, B;
  R.x = (float)((double)r[0] - mean_r); R.y = (float)((double)r[1] - mean_r);
  R.z = (float)((double)r[2] - mean_r); R.w = (float)((double)r[3] - mean_r);
  G.x = (float)((double)g[0] - mean_g); G.y = (float)((double)g[1] - mean_g);
  G.z = (float)((double)g[2] - mean_g); G.w = (float)((double)g[3] - mean_g);
  B.x = (float)((double)b[0] - mean_b); B.y = (float)((double)b[1] - mean_b);
  B.z = (float)((double)b[2] - mean_b); B.w = (float)((double)b[3] - mean_b);
  float4* o = reinterpret_cast<float4*>(out) + i;
  o[0] = R;                                  // channel 0 = R  (im[:, :, 2])
  o[plane / 4] = G;                          // channel 1 = G
  o[plane / 2] = B;                          // channel 2 = B
}

// scalar tail / unaligned variant (frame sizes that are not a multiple of 4 pixels, or odd base pointers)
__global__ void __launch_bounds__(256) preprocess_scalar_kernel(const uint8_t* __restrict__ src, size_t npix,
                                                                double mean_b, double mean_g, double mean_r,
                                                                float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  out[i] = (float)((double)src[3 * i + 2] - mean_r);
  out[npix + i] = (float)((double)src[3 * i + 1] - mean_g);
  out[2 * npix + i] = (float)((double)src[3 * i] - mean_b);
}

// cv2.resize(im, None, None, fx, fy, INTER_LINEAR) for 8-bit BGR (lib/utils/image.py:211), bit for bit: OpenCV's fixed-point
// separable bilinear.  Per destination column: fx = (float)((dx + 0.5) / fx_scale - 0.5), sx = floor, frac clamped to 0 at
// both borders, coefficients rounded (half to even) to 11 bits; per destination row the same WITHOUT clamping the fraction
// (both rows then clip to the border row, OpenCV's `clip(sy, 0, ssize.height)`); the horizontal pass keeps 11 fractional
// bits and the vertical pass is ((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2.  An exact 2x decimation is
// what OpenCV turns into INTER_AREA: the rounded mean of the 2x2 block.  Pinned against cv2 4.13 (tests/golden/
// make_resize_vectors.py, tests/test_gpu_io.py).  One thread per destination pixel (3 bytes); HBM streaming.
__global__ void __launch_bounds__(256) resize_linear_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst,
                                                            int dh, int dw, double scale_x, double scale_y, int area2) {
  pdl_trigger();
  pdl_wait();
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= dw || dy >= dh) return;
  uint8_t* o = dst + ((size_t)dy * dw + dx) * 3;
  if (area2) {
    const uint8_t* p0 = src + ((size_t)(2 * dy) * sw + 2 * dx) * 3;
    const uint8_t* p1 = p0 + (size_t)sw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (uint8_t)((p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
    return;
  }
  // un-fused double multiply / subtract, as the host code OpenCV runs (an FMA here flips a coefficient now and then)
  float fx = (float)__dsub_rn(__dmul_rn((double)dx + 0.5, scale_x), 0.5);
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) { fx = 0.f; sx = 0; }
  if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
  const int a0 = __float2int_rn((1.f - fx) * 2048.f), a1 = __float2int_rn(fx * 2048.f);
  const int sx1 = min(sx + 1, sw - 1);
  float fy = (float)__dsub_rn(__dmul_rn((double)dy + 0.5, scale_y), 0.5);
  const int sy = (int)floorf(fy);
  fy -= (float)sy;
  const int b0 = __float2int_rn((1.f - fy) * 2048.f), b1 = __float2int_rn(fy * 2048.f);
  const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
  const uint8_t* r0 = src + (size_t)y0 * sw * 3;
  const uint8_t* r1 = src + (size_t)y1 * sw * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = r0[sx * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;
    const int h1 = r1[sx * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
    o[c] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
  }
}

// fast_hist: hist[label * n + pred] += 1 for every pixel with label < n (labels are uint8, so `label >= 0`
// always holds; 255 = Cityscapes "ignore").  pred >= n cannot come out of the argmax; such pixels are
// dropped rather than corrupting a neighbouring bin.  Per-CTA histogram in shared memory (32-bit), 16
// pixels per thread per iteration as one 128-bit load of each map, flushed with 64-bit global atomics.
constexpr int CM_MAX = 32;
__global__ void __launch_bounds__(256) confusion_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ label,
                                                        size_t n16, size_t n, int K, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[CM_MAX * CM_MAX];
  pdl_trigger();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  pdl_wait();
  const uint4* p4 = reinterpret_cast<const uint4*>(pred);
  const uint4* l4 = reinterpret_cast<const uint4*>(label);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 pv = __ldg(p4 + i), lv = __ldg(l4 + i);
    const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const unsigned l = (lw[w] >> (8 * b)) & 255u, p = (pw[w] >> (8 * b)) & 255u;
        if (l < (unsigned)K && p < (unsigned)K) atomicAdd(&sh[l * K + p], 1u);
      }
  }
  // the last (n mod 16) pixels
  if (blockIdx.x == 0)
    for (size_t i = n16 * 16 + threadIdx.x; i < n; i += blockDim.x) {
      const unsigned l = label[i], p = pred[i];
      if (l < (unsigned)K && p < (unsigned)K) atomicAdd(&sh[l * K + p], 1u);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

}  // namespace

cudaError_t launch_preprocess(const uint8_t* bgr_hwc, int H, int W, const double mean_bgr[3], float* out,
                              cudaStream_t stream) {
  const size_t npix = (size_t)H * W;
  if (npix == 0) return cudaSuccess;
  const bool vec = (npix % 4 == 0) && (((uintptr_t)bgr_hwc & 3) == 0) && (((uintptr_t)out & 15) == 0);
  if (vec) {
    const int npix4 = (int)(npix / 4);
    return launch_k(preprocess_kernel, dim3((npix4 + 255) / 256), dim3(256), 0, stream, bgr_hwc, npix4, npix, mean_bgr[0],
                    mean_bgr[1], mean_bgr[2], out);
  } else {
    return launch_k(preprocess_scalar_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, stream, bgr_hwc, npix,
                    mean_bgr[0], mean_bgr[1], mean_bgr[2], out);
  }
  return cudaGetLastError();
}

void resize_linear_size(int sh, int sw, double fx, double fy, int* dh, int* dw) {
  *dw = (int)nearbyint((double)sw * fx);          // saturate_cast<int>(ssize.width * inv_scale_x): round half to even
  *dh = (int)nearbyint((double)sh * fy);
}

cudaError_t launch_resize_linear(const uint8_t* src, int sh, int sw, double fx, double fy, uint8_t* dst, cudaStream_t stream) {
  int dh, dw;
  resize_linear_size(sh, sw, fx, fy, &dh, &dw);
  if (dh < 1 || dw < 1 || sh < 1 || sw < 1) return cudaErrorInvalidValue;
  const double scale_x = 1.0 / fx, scale_y = 1.0 / fy;
  // OpenCV: INTER_LINEAR with an exact integer 2x decimation in both directions is computed as INTER_AREA
  const int area2 = (scale_x == 2.0 && scale_y == 2.0 && sw == 2 * dw && sh == 2 * dh) ? 1 : 0;
  return launch_k(resize_linear_kernel, dim3((dw + 255) / 256, dh), dim3(256), 0, stream, src, sh, sw, dst, dh, dw, scale_x, scale_y,
                  area2);
}

cudaError_t launch_confusion(const uint8_t* pred, const uint8_t* label, size_t n, int K, unsigned long long* hist,
                             cudaStream_t stream) {
  if (K < 1 || K > CM_MAX) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  const bool vec = (((uintptr_t)pred | (uintptr_t)label) & 15) == 0;
  const size_t n16 = vec ? n / 16 : 0;
  size_t blocks = (n16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 2 * 148) blocks = 2 * 148;          // two CTAs per SM; each flushes K*K atomics once
  return launch_k(confusion_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, pred, label, n16, n, K, hist);
}

}  // namespace accel
