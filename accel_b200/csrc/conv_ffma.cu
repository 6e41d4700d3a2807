// CUDA-core (FFMA) implicit-GEMM convolution over the split-fp16 NHWC format.
//
// This is the fp32 "safety net" contraction: layers the tcgen05 kernel does not take (odd shapes,
// tiny maps, narrow outputs) run here.  Same operands, same epilogue, same packed weights.
#include "common.cuh"
#include "kernels.h"

namespace accel {

namespace {

constexpr int BM = 64;   // output pixels per CTA
constexpr int BN = 64;   // output channels per CTA
constexpr int BK = 32;   // channels of one tap per main-loop step
constexpr int LDS = BM + 4;

__device__ __forceinline__ void cvt8(const Half8& a, const Half8& b, float out[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 fa = __half22float2(a.v[i]);
    float2 fb = __half22float2(b.v[i]);
    out[2 * i] = fa.x + fb.x;
    out[2 * i + 1] = fa.y + fb.y;
  }
}

__global__ void __launch_bounds__(256) conv_ffma_kernel(const ConvParams P) {
  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Bs[BK][LDS];

  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int m_load = tid & 63;       // pixel (A) / channel row (B) this thread stages
  const int j_load = tid >> 6;       // which 8-wide k sub-chunk
  const int tx = tid & 15, ty = tid >> 4;

  const int npix = P.Ho * P.Wo;
  const int p0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-side: the output pixel this thread loads for
  const int p_load = p0 + m_load;
  const bool p_ok = p_load < npix;
  const int oy = p_ok ? p_load / P.Wo : 0;
  const int ox = p_ok ? p_load - oy * P.Wo : 0;
  const int iy0 = oy * P.stride, ix0 = ox * P.stride;

  const int chunks = P.Cin_pad / BK;
  const int iters = P.ntaps * chunks;
  const int it_begin = (int)(((long long)iters * blockIdx.z) / P.splits);
  const int it_end = (int)(((long long)iters * (blockIdx.z + 1)) / P.splits);

  const __half* wrow_hi = P.w_hi + (size_t)(n0 + m_load) * P.Kpad + j_load * 8;
  const __half* wrow_lo = P.w_lo + (size_t)(n0 + m_load) * P.Kpad + j_load * 8;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  Half8 ra_hi, ra_lo, rb_hi, rb_lo;
  auto fetch = [&](int it) {
    const int t = it / chunks;
    const int k0 = (it - t * chunks) * BK + j_load * 8;
    const int iy = iy0 + P.dy[t], ix = ix0 + P.dx[t];
    const bool ok = p_ok && iy >= 0 && iy < P.Hin && ix >= 0 && ix < P.Win && k0 < P.Cin;
    if (ok) {
      const size_t off = ((size_t)iy * P.Win + ix) * P.in_ld + k0;
      ra_hi = *reinterpret_cast<const Half8*>(P.in_hi + off);
      ra_lo = *reinterpret_cast<const Half8*>(P.in_lo + off);
    } else {
      ra_hi = Half8{};
      ra_lo = Half8{};
    }
    const size_t woff = (size_t)t * P.Cin_pad + (it - t * chunks) * BK;
    rb_hi = *reinterpret_cast<const Half8*>(wrow_hi + woff);
    rb_lo = *reinterpret_cast<const Half8*>(wrow_lo + woff);
  };

  if (it_begin < it_end) fetch(it_begin);
  for (int it = it_begin; it < it_end; ++it) {
    float fa[8], fb[8];
    cvt8(ra_hi, ra_lo, fa);
    cvt8(rb_hi, rb_lo, fb);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      As[j_load * 8 + i][m_load] = fa[i];
      Bs[j_load * 8 + i][m_load] = fb[i];
    }
    __syncthreads();
    if (it + 1 < it_end) fetch(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  const int c0 = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty * 4 + i;
    if (p >= npix) continue;
    if (P.splits > 1) {
      float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(P.partial + ((size_t)blockIdx.z * npix + p) * P.Cout_pad + c0) = v;
    } else if (c0 < P.epi.Cout) {
      const int y = p / P.Wo, x = p - y * P.Wo;
      const int pix = (y * P.epi.osy + P.epi.ooy) * P.epi.OWf + x * P.epi.osx + P.epi.oox;
      epilogue_store<4>(P.epi, pix, c0, acc[i]);
    }
  }
}

// Deterministic split-K tail: sum the partial slabs in split order, then the shared epilogue.
__global__ void __launch_bounds__(256) splitk_epilogue_kernel(const float* __restrict__ partial, int splits, int npix,
                                                              int Cout_pad, int Wo, const Epilogue epi) {
  pdl_trigger();
  pdl_wait();
  const int groups = (epi.Cout + 3) / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)npix * groups) return;
  const int p = (int)(idx / groups);
  const int c0 = (int)(idx - (long long)p * groups) * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < splits; ++s) {
    const float4 t = *reinterpret_cast<const float4*>(partial + ((size_t)s * npix + p) * Cout_pad + c0);
    v[0] += t.x;
    v[1] += t.y;
    v[2] += t.z;
    v[3] += t.w;
  }
  const int y = p / Wo, x = p - y * Wo;
  const int pix = (y * epi.osy + epi.ooy) * epi.OWf + x * epi.osx + epi.oox;
  epilogue_store<4>(epi, pix, c0, v);
}

// Narrow-output convolution (Cout <= 8: FlowNet's 2-channel flow heads, ...flownet_deeplab.py:1774-1803).
// One warp per output pixel; lanes stride the (tap, channel) axis 8 channels at a time and the
// partial dot products meet in a shuffle tree.  Output is fp32 planar (and/or split NHWC).
template <int NOUT>
__global__ void __launch_bounds__(256) conv_narrow_kernel(const ConvParams P) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int npix = P.Ho * P.Wo;
  if (warp >= npix) return;
  const int oy = warp / P.Wo, ox = warp - oy * P.Wo;
  float acc[NOUT];
#pragma unroll
  for (int n = 0; n < NOUT; ++n) acc[n] = 0.f;
  const int c8 = (P.Cin + 7) / 8;
  for (int t = 0; t < P.ntaps; ++t) {
    const int iy = oy * P.stride + P.dy[t], ix = ox * P.stride + P.dx[t];
    if (iy < 0 || iy >= P.Hin || ix < 0 || ix >= P.Win) continue;
    const size_t base = ((size_t)iy * P.Win + ix) * P.in_ld;
    for (int g = lane; g < c8; g += 32) {
      float a[8];
      load8(P.in_hi + base + g * 8, P.in_lo + base + g * 8, a);
#pragma unroll
      for (int n = 0; n < NOUT; ++n) {
        float w[8];
        const size_t woff = (size_t)n * P.Kpad + (size_t)t * P.Cin_pad + g * 8;
        load8(P.w_hi + woff, P.w_lo + woff, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[n] = fmaf(a[i], w[i], acc[n]);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NOUT; ++n)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  if (lane == 0) {
    const Epilogue& e = P.epi;
    const int pix = (oy * e.osy + e.ooy) * e.OWf + ox * e.osx + e.oox;
    const size_t plane = (size_t)e.OHf * e.OWf;
#pragma unroll
    for (int n = 0; n < NOUT; ++n) {
      if (n >= e.Cout) break;
      float v = fmaf(acc[n], e.scale ? e.scale[n] : 1.f, e.shift ? e.shift[n] : 0.f);
      v = apply_act(v, e.act);
      if (e.out_nchw) e.out_nchw[n * plane + pix] = v;
      if (e.out_hi) {
        __half h, l;
        split_f32(v, h, l);
        e.out_hi[(size_t)pix * e.out_ld + n] = h;
        e.out_lo[(size_t)pix * e.out_ld + n] = l;
      }
    }
  }
}

}  // namespace

int ffma_pick_splits(const ConvParams& P, int num_sms) {
  const int npix = P.Ho * P.Wo;
  const long long tiles = (long long)((npix + BM - 1) / BM) * (P.Cout_pad / BN);
  const int iters = P.ntaps * (P.Cin_pad / BK);
  if (tiles >= 2LL * num_sms || iters < 16) return 1;
  long long s = (4LL * num_sms + tiles - 1) / tiles;
  if (s > iters / 8) s = iters / 8;
  if (s > 32) s = 32;
  return s < 1 ? 1 : (int)s;
}

size_t ffma_partial_bytes(const ConvParams& P, int splits) {
  return splits > 1 ? (size_t)splits * P.Ho * P.Wo * P.Cout_pad * sizeof(float) : 0;
}

cudaError_t launch_conv_ffma(const ConvParams& P, cudaStream_t stream) {
  const int npix = P.Ho * P.Wo;
  dim3 grid((npix + BM - 1) / BM, P.Cout_pad / BN, P.splits);
  cudaError_t e = launch_k(conv_ffma_kernel, grid, dim3(256), 0, stream, P);
  if (e == cudaSuccess && P.splits > 1) {
    const long long work = (long long)npix * ((P.epi.Cout + 3) / 4);
    e = launch_k(splitk_epilogue_kernel, dim3((unsigned)((work + 255) / 256)), dim3(256), 0, stream, (const float*)P.partial,
                 P.splits, npix, P.Cout_pad, P.Wo, P.epi);
  }
  return e;
}

cudaError_t launch_splitk_epilogue(const float* partial, int splits, int npix, int Cout_pad, int Wo, const Epilogue& epi,
                                   cudaStream_t stream) {
  const long long work = (long long)npix * ((epi.Cout + 3) / 4);
  return launch_k(splitk_epilogue_kernel, dim3((unsigned)((work + 255) / 256)), dim3(256), 0, stream, partial, splits, npix,
                  Cout_pad, Wo, epi);
}

cudaError_t launch_conv_narrow(const ConvParams& P, cudaStream_t stream) {
  const int npix = P.Ho * P.Wo;
  const unsigned blocks = (unsigned)(((long long)npix * 32 + 255) / 256);
  if (P.epi.Cout <= 2) return launch_k(conv_narrow_kernel<2>, dim3(blocks), dim3(256), 0, stream, P);
  if (P.epi.Cout <= 4) return launch_k(conv_narrow_kernel<4>, dim3(blocks), dim3(256), 0, stream, P);
  return launch_k(conv_narrow_kernel<8>, dim3(blocks), dim3(256), 0, stream, P);
}

}  // namespace accel
