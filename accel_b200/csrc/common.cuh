// Shared device-side types for the Accel hot-path kernels (sm_100a only).
//
// Internal activation format ("split fp16"): every activation tensor is NHWC with two fp16 planes,
// hi = rn16(x) and lo = rn16(x - hi), so hi + lo carries ~22 mantissa bits.  The tensor-core convs
// consume the planes directly (fp16x3: hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM), the
// CUDA-core kernels add them back to fp32 on load.  fp32 NCHW only exists at the C-ABI boundary
// (frames, feat_key / feat_out, score volumes) -- the layouts the reference's Predictor exchanges
// (dff_deeplab/core/tester.py:32-35, demo.py:184).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

namespace accel {

enum ActKind { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// A plan is a chain of several hundred short kernels on one stream.  Every kernel is launched with the
// programmatic-stream-serialization attribute (launch_k below; the edges survive CUDA-graph capture), calls
// pdl_trigger() first thing -- so the next kernel's CTAs may be scheduled, and run their prologue (barrier
// init, TMEM allocation, descriptor prefetch, index arithmetic), on SMs as they drain -- and calls pdl_wait()
// before its first global-memory access that is not a constant parameter / weight.  pdl_wait() returns once
// the preceding kernel has completed and flushed, so the chain's memory ordering is that of a plain stream;
// because EVERY kernel waits before it reads or writes activations, the guarantee is transitive.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// True the first time it is called for (slot, current device): cudaFuncSetAttribute is per device, and one process may
// drive several GPUs (pred_eval_multiprocess uses one predictor pair per GPU from threads).
bool first_time_on_device(int slot);
enum { ONCE_CONV_TC = 0, ONCE_STEM_TC, ONCE_STEM, ONCE_WARP_STAGED_0, ONCE_WARP_STAGED_1, ONCE_WARP_STAGED_2, ONCE_WARP_STAGED_3, ONCE_WARP_FUSED, ONCE_WARP_FUSED_1, ONCE_SLOTS };

bool pdl_enabled();   // graph.cu: ACCEL_PDL=1 turns it on (measured neutral on B200 for these plans, so off by default)

// Same, as thread-block clusters of `cluster_x` CTAs along x (grid.x must be a multiple of it).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

constexpr int kMaxTaps = 25;   // 5x5 is the largest generic kernel (FlowNet conv2/conv3); 7x7 stems are separate

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LEAKY) return v > 0.f ? v : v * 0.1f;   // LeakyReLU(slope=0.1), ...flownet_deeplab.py:1755
  return v;
}

// Flow-guided warp: an out-of-range tap carries the weight -0.0f, which no in-range tap can have (its weights are products
// of values in [0, 1]).  Out-of-range taps are skipped by VALUE, as MXNet's BilinearSampler skips them -- the clamped
// neighbour they point at may hold Inf -- while an in-range tap whose weight happens to be +0 still multiplies, as it does
// there (0 * Inf = NaN in both).
// The same activation over a register array with the switch taken ONCE (apply_act per element compiles to a branch per
// element when `act` is a run-time value: 32 taken branches per 32-channel chunk in the conv epilogues).
template <int N>
__device__ __forceinline__ void apply_act_n(float (&v)[N], int act) {
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (act == ACT_LEAKY) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * 0.1f;
  }
}

constexpr float kSkipTap = -0.0f;
__device__ __forceinline__ bool keep_tap(float w) { return __float_as_uint(w) != 0x80000000u; }

// Two values at a time: hi2 = rn16x2(v), lo2 = rn16x2(v - hi2), both saturating to +-65504 (one packed convert each
// instead of clamp + scalar converts: 3 instructions per element instead of 7 -- the split is on the critical path
// of every epilogue).  For |v| <= 65504 this is exactly hi = rn16(v), lo = rn16(v - hi).
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi2, uint32_t& lo2) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(v1), "f"(v0));
  const __half2 h = *reinterpret_cast<const __half2*>(&hi2);
  const float2 f = __half22float2(h);
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(v1 - f.y), "f"(v0 - f.x));
}

__device__ __forceinline__ void split_f32(float v, __half& hi, __half& lo) {
  uint32_t h2, l2;
  split_pair(v, 0.f, h2, l2);
  hi = __ushort_as_half((unsigned short)(h2 & 0xffffu));
  lo = __ushort_as_half((unsigned short)(l2 & 0xffffu));
}

struct alignas(16) Half8 {
  __half2 v[4];
};

// 8 consecutive channels of one pixel: hi + lo -> fp32.  Explicit 128-bit accesses: a copy of the Half8 struct was
// scalarised by the compiler into four 32-bit loads / stores (cuobjdump: 34 LDG.E.32 in dcn_col_kernel), i.e. four times the
// L1 data-pipe wavefronts -- which is what that kernel is bound by (ncu: l1tex__data_pipe_lsu_wavefronts 90 % of peak).
__device__ __forceinline__ void load8(const __half* hi, const __half* lo, float out[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(hi);
  const uint4 b = *reinterpret_cast<const uint4*>(lo);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[i]));
    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[i]));
    out[2 * i] = fa.x + fb.x;
    out[2 * i + 1] = fa.y + fb.y;
  }
}

__device__ __forceinline__ void store8(__half* hi, __half* lo, const float v[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair(v[2 * i], v[2 * i + 1], h[i], l[i]);
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

struct alignas(8) Half4 {
  __half2 v[2];
};

__device__ __forceinline__ void load4(const __half* hi, const __half* lo, float out[4]) {
  Half4 a = *reinterpret_cast<const Half4*>(hi);
  Half4 b = *reinterpret_cast<const Half4*>(lo);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float2 fa = __half22float2(a.v[i]);
    float2 fb = __half22float2(b.v[i]);
    out[2 * i] = fa.x + fb.x;
    out[2 * i + 1] = fa.y + fb.y;
  }
}

__device__ __forceinline__ void store4(__half* hi, __half* lo, const float v[4]) {
  Half4 a, b;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint32_t h2, l2;
    split_pair(v[2 * i], v[2 * i + 1], h2, l2);
    a.v[i] = *reinterpret_cast<__half2*>(&h2);
    b.v[i] = *reinterpret_cast<__half2*>(&l2);
  }
  *reinterpret_cast<Half4*>(hi) = a;
  *reinterpret_cast<Half4*>(lo) = b;
}

// What every conv-like kernel does with an accumulator once the contraction is finished:
//   v = acc * scale[c] + shift[c]  (+ residual)  -> act  -> out (split NHWC) and/or out_nchw (fp32)
//   out2 = act2(v * scale2[c] + shift2[c])           (pre-activation ResNets: next unit's bn1+relu)
// Output pixel (y, x) of the kernel's own loop space lands at (y*osy + ooy, x*osx + oox) of the full
// output map (OHf x OWf); transposed-conv phases use osy = osx = 2.
struct Epilogue {
  const float* scale;
  const float* shift;
  int act;
  const __half* res_hi;
  const __half* res_lo;
  int res_ld;
  __half* out_hi;
  __half* out_lo;
  int out_ld;
  const float* scale2;
  const float* shift2;
  int act2;
  __half* out2_hi;
  __half* out2_lo;
  int out2_ld;
  float* out_nchw;
  float* raw_nchw;           // fp32 NCHW copy of acc * scale (no shift, no residual, no activation), or nullptr
  int osy, osx, ooy, oox;
  int OHf, OWf;
  int Cout;
  // Batched plans stack their frames along H (OHf = frames * per-frame height): the split NHWC tensors are then
  // simply taller, but fp32 planar outputs stay per frame, [frame][Cout][nchw_hw]:
  int nchw_hw;               // pixels of ONE frame of the full output map (0 = OHf * OWf: a single frame)
  int nchw_nb;               // only the first nchw_nb frames receive the fp32 copy (0 = all)
};

// Offset of channel 0 of full-map pixel `pix` in an fp32 planar output (channel c sits c * hw further); false = this
// frame gets no fp32 copy.
__device__ __forceinline__ bool nchw_base(const Epilogue& e, int pix, size_t& base, size_t& hw) {
  hw = e.nchw_hw ? (size_t)e.nchw_hw : (size_t)e.OHf * e.OWf;
  const int b = (int)((size_t)pix / hw);
  if (e.nchw_nb && b >= e.nchw_nb) return false;
  base = (size_t)b * e.Cout * hw + ((size_t)pix - (size_t)b * hw);
  return true;
}

// Applies the epilogue to NV (4 or 8) consecutive channels [c0, c0+NV) of full-map pixel `pix`.
template <int NV>
__device__ __forceinline__ void epilogue_store(const Epilogue& e, int pix, int c0, float v[NV]) {
  float r[NV];
  if (e.res_hi) {
    if (NV == 8) load8(e.res_hi + (size_t)pix * e.res_ld + c0, e.res_lo + (size_t)pix * e.res_ld + c0, r);
    else load4(e.res_hi + (size_t)pix * e.res_ld + c0, e.res_lo + (size_t)pix * e.res_ld + c0, r);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = c0 + i;
    float x = 0.f;
    if (c < e.Cout) {
      float s = e.scale ? e.scale[c] : 1.f;
      float b = e.shift ? e.shift[c] : 0.f;
      if (e.raw_nchw) {
        size_t nb_, hw_;
        if (nchw_base(e, pix, nb_, hw_)) e.raw_nchw[nb_ + (size_t)c * hw_] = v[i] * s;
      }
      x = fmaf(v[i], s, b);
      if (e.res_hi) x += r[i];
      x = apply_act(x, e.act);
    }
    v[i] = x;
  }
  if (e.out_hi) {
    if (NV == 8) store8(e.out_hi + (size_t)pix * e.out_ld + c0, e.out_lo + (size_t)pix * e.out_ld + c0, v);
    else store4(e.out_hi + (size_t)pix * e.out_ld + c0, e.out_lo + (size_t)pix * e.out_ld + c0, v);
  }
  if (e.out_nchw) {
    size_t base, plane;
    if (nchw_base(e, pix, base, plane)) {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (c0 + i < e.Cout) e.out_nchw[base + (size_t)(c0 + i) * plane] = v[i];
    }
  }
  if (e.out2_hi) {
    float w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = c0 + i;
      float x = 0.f;
      if (c < e.Cout) x = apply_act(fmaf(v[i], e.scale2[c], e.shift2[c]), e.act2);
      w[i] = x;
    }
    if (NV == 8) store8(e.out2_hi + (size_t)pix * e.out2_ld + c0, e.out2_lo + (size_t)pix * e.out2_ld + c0, w);
    else store4(e.out2_hi + (size_t)pix * e.out2_ld + c0, e.out2_lo + (size_t)pix * e.out2_ld + c0, w);
  }
}

// Geometry + operands of one convolution-like contraction over the internal format.
//   out[y, x, n] = sum_t sum_c in[y*stride + dy[t], x*stride + dx[t], c] * w[n][t*Cin_pad + c]
// Weights are packed K-major per output channel (rows of Kpad = ntaps*Cin_pad fp16, hi and lo planes),
// rows padded to a multiple of 64 output channels and Cin_pad to a multiple of 64 with zeros.
struct ConvParams {
  const __half* in_hi;
  const __half* in_lo;
  int in_ld, Hin, Win, Cin;
  const __half* w_hi;
  const __half* w_lo;
  int Kpad, Cin_pad, Cout_pad;
  int ntaps;
  int8_t dy[kMaxTaps];
  int8_t dx[kMaxTaps];
  int stride;
  int Ho, Wo;          // loop space (== output size except for transposed-conv phases), ONE frame
  int nb;              // frames (0/1 = one): input, outputs and residual hold `nb` frames stacked densely along H
  int splits;          // split-K factor (1 = epilogue in-kernel)
  float* partial;      // [splits][Ho*Wo][Cout_pad] fp32 when splits > 1
  int ext_outputs;     // 1: a launch may add caller-owned fp32 outputs (launch_conv_tc_ext) to `epi`
  Epilogue epi;
};

}  // namespace accel
