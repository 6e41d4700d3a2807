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
