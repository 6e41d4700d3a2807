// Host-callable launchers for the hot-path kernels.  Internal to the library (the public surface is
// include/accel_b200.h).
#pragma once
#include <stdio.h>
#include <vector>

#include "common.cuh"

namespace accel {

// ---- convolution engines -------------------------------------------------------------------------
int ffma_pick_splits(const ConvParams& P, int num_sms);
size_t ffma_partial_bytes(const ConvParams& P, int splits);
cudaError_t launch_conv_ffma(const ConvParams& P, cudaStream_t stream);
cudaError_t launch_conv_narrow(const ConvParams& P, cudaStream_t stream);

// tcgen05 / TMEM / TMA implicit-GEMM (conv_tc.cu)
struct TcPlan;   // opaque: tensor maps + tiling chosen at plan time
bool tc_supported(const ConvParams& P);
bool tc_batchable(const ConvParams& P);   // P.nb frames in one launch (frames stacked along H)?
TcPlan* tc_plan_create(const ConvParams& P, int num_sms, char* err, int errlen);
void tc_plan_destroy(TcPlan* plan);
void tc_trace_dump(FILE* f);      // tuning aid: timeline of CTA 0 (ACCEL_TC_DEBUG bit 2048)
void tc_trace_reset();
size_t tc_plan_partial_bytes(const TcPlan* plan);
void tc_plan_set_partial(TcPlan* plan, float* partial);
// ext_nchw: optional fp32 NCHW destination that replaces the plan's epilogue out_nchw for this launch
// ext_raw: optional fp32 NCHW destination of acc * scale (Epilogue::raw_nchw) for this launch
cudaError_t launch_conv_tc_ext(const TcPlan* plan, float* ext_nchw, float* ext_raw, cudaStream_t stream);
int tc_plan_launches(const TcPlan* plan);

// ---- stems (7x7 / stride 2 / pad 3, 64 output channels, fp32 NCHW source) ------------------------
struct StemParams {
  const float* src0;       // (3, Hs, Ws) fp32 planar: the frame
  const float* src1;       // second frame (FlowNet: data_key) or nullptr
  int Hs, Ws;              // source size
  int pool;                // 1: average 2x2 blocks of src/255 first (FlowNet `resize_data`), 0: none
  int Cin;                 // 3, or 6 when src1 is set
  float in_scale[6];       // per-input-channel affine applied to in-bounds samples (R18/34 `bn_data`)
  float in_shift[6];
  const float* weight;     // fp32 (64, Cin, 7, 7)
  Epilogue epi;            // out = split NHWC (64, Ho, Wo)
  int Ho, Wo;
};
cudaError_t launch_stem(const StemParams& P, cudaStream_t stream);   // CUDA-core cross-check path

// tcgen05 stem (stem_tc.cu): software-built A operand, weights packed by stem_tc_pack_weights
struct StemTcPlan;
int stem_tc_kpad(int cin);
void stem_tc_pack_weights(const float* w, int cin, std::vector<__half>& hi, std::vector<__half>& lo,
                          std::vector<float>& prescale);
bool stem_tc_supported(const StemParams& S);
StemTcPlan* stem_tc_plan_create(const StemParams& S, const __half* w_hi, const __half* w_lo, int num_sms, char* err,
                                int errlen);
void stem_tc_plan_destroy(StemTcPlan* plan);
cudaError_t launch_stem_tc(const StemTcPlan* plan, const float* src0, const float* src1, cudaStream_t stream);

// ---- pooling ---------------------------------------------------------------------------------------
struct PoolParams {
  const __half* in_hi;
  const __half* in_lo;
  int in_ld, Hin, Win, C;
  __half* out_hi;
  __half* out_lo;
  int out_ld, Ho, Wo;
  int kernel, stride, pad;   // max: window clipped at the border; avg: exact tiles only
  int is_max;
  const float* scale;        // optional per-channel affine + activation on the pooled value
  const float* shift;        // (pre-activation ResNets: the next unit's bn1 + relu)
  int act;
};
cudaError_t launch_pool(const PoolParams& P, cudaStream_t stream);

// ---- deformable im2col (DCNv1) ----------------------------------------------------------------------
struct DcnColParams {
  const __half* in_hi;
  const __half* in_lo;
  int in_ld, H, W, C;
  const float* offset;       // fp32 planar (dg*18, H, W)
  int dg, dilate, pad;       // 3x3, stride 1
  __half* col_hi;            // split NHWC (9*C, H, W): channel = tap*C + c
  __half* col_lo;
  int col_ld;
};
cudaError_t launch_dcn_col(const DcnColParams& P, cudaStream_t stream);

// ---- flow-guided warp (GridGenerator(warp) + BilinearSampler) -----------------------------------------
struct WarpParams {
  const float* feat;         // fp32 NCHW (C, H, W)
  const float* flow;         // fp32 planar (2, H, W): dx, dy in feature-grid pixels
  int C, H, W;
  float* out_nchw;           // fp32 NCHW warped feature (`warping_feat_output`) or nullptr
  __half* out_hi;            // split NHWC copy feeding the task head, or nullptr
  __half* out_lo;
  int out_ld;
  const float* bias;         // split copy only: + bias[c], then `act` (commuted L head: relu(warp(W*F) + fc6_bias)); or nullptr
  int act;
};
// Writes the fp32 NCHW warped feature (when out_nchw is set) and the split NHWC copy (when out_hi is set).  Shapes the
// fused kernel supports do both in ONE pass over the source (warp_kernel_fused, warp_staged.cu); others run the staged /
// gather kernel and, for the split copy, nchw_to_split_kernel over its output (out_nchw is then required).
cudaError_t launch_warp(const WarpParams& P, cudaStream_t stream);
bool warp_fused_supported(const WarpParams& P);
cudaError_t launch_warp_fused(const WarpParams& P, cudaStream_t stream);
// shared-memory staged variant (warp_staged.cu); cudaErrorNotSupported = shape unsuitable, use the gather kernel
cudaError_t launch_warp_staged(const WarpParams& P, cudaStream_t stream);

// ---- weight preparation (finalize time) --------------------------------------------------------------------
// out[c][n][t] = sum_m conv1x1[n][m] * deconv[c][m][t]  (t = ky*4+kx; deconv (Cin, mid, 4, 4), conv1x1 (cout, mid),
// out (Cin, cout, 4, 4)); fp32 operands, fp64 accumulation.
cudaError_t launch_fold_deconv_1x1(const float* deconv, const float* conv1x1, float* out, int cin, int mid, int cout,
                                   cudaStream_t stream);

// ---- layout conversion --------------------------------------------------------------------------------
// bias / act: optional per-channel bias added and activation applied on the way (nullptr / ACT_NONE = plain copy)
cudaError_t launch_nchw_to_split(const float* src, int C, int H, int W, __half* hi, __half* lo, int ld,
                                 cudaStream_t stream, const float* bias = nullptr, int act = 0);
cudaError_t launch_split_to_nchw(const __half* hi, const __half* lo, int ld, int C, int H, int W, float* dst,
                                 cudaStream_t stream);

// ---- tiny 2->2 channel 4x4/s2/p1 transposed conv (FlowNet `upsample_flow*`) ----------------------------
struct UpflowParams {
  const float* flow;         // fp32 planar (2, H, W)
  int H, W;
  float weight[2 * 2 * 16];  // (cin, cout, 4, 4)
  float bias[2];
  __half* out_hi;            // split NHWC view (2 channels) of the 2H x 2W concat buffer
  __half* out_lo;
  int out_ld;
};
cudaError_t launch_upflow(const UpflowParams& P, cudaStream_t stream);

// ---- score fusion + x16 bilinear upsampling + argmax ---------------------------------------------------
struct FuseParams {          // low-res 1x1 fusion: out[c] = sum_j wa[c][j] a[j] + sum_j wb[c][j] b[j]
  const float* a;            // fp32 planar (K, h, w)
  const float* b;
  const float* w;            // fp32 (K, 2K) row-major: `corr_weight`
  float* out;                // fp32 planar (K, h, w)
  int K, h, w_;
};
cudaError_t launch_fuse_lowres(const FuseParams& P, cudaStream_t stream);

struct TailParams {
  const float* score;        // fp32 planar (K, h, w) low-res scores
  const float* bias;         // per-class bias added after interpolation (corr_bias) or nullptr
  int K, h, w;
  int factor;                // 16
  uint8_t* label;            // (h*factor, w*factor)
  float* score_out;          // fp32 planar (K, h*factor, w*factor) or nullptr
  // optional: `score` = fuse_w (K, 2K) applied to concat(fuse_a, fuse_b) (FuseParams), evaluated inside the band
  // kernel for each CTA's source window instead of by a launch of its own (only where tail_band_supported())
  const float* fuse_a;
  const float* fuse_b;
  const float* fuse_w;
};
bool tail_band_supported(int K, int factor);
cudaError_t launch_tail(const TailParams& P, cudaStream_t stream);

// ---- frame ingest + metric (kernels_io.cu) ---------------------------------------------------------------
// lib/utils/image.py:224-235 transform(): uint8 BGR HWC -> fp32 RGB NCHW minus PIXEL_MEANS (B, G, R order)
cudaError_t launch_preprocess(const uint8_t* bgr_hwc, int H, int W, const double mean_bgr[3], float* out,
                              cudaStream_t stream);
// lib/utils/image.py:211 cv2.resize(im, None, None, fx, fy, cv2.INTER_LINEAR) on (H, W, 3) uint8, bit-exact vs OpenCV
void resize_linear_size(int sh, int sw, double fx, double fy, int* dh, int* dw);
cudaError_t launch_resize_linear(const uint8_t* src, int sh, int sw, double fx, double fy, uint8_t* dst, cudaStream_t stream);
// dff_deeplab/demo.py:50-53 fast_hist(): hist[label * K + pred] += 1 where label < K (device int64 K x K)
cudaError_t launch_confusion(const uint8_t* pred, const uint8_t* label, size_t n, int K, unsigned long long* hist,
                             cudaStream_t stream);

}  // namespace accel
