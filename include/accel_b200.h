/* accel_b200 -- C ABI of the B200-native Accel (dff_deeplab) per-frame hot path.
 *
 * Everything the reference does per video frame between "two preprocessed frames are in device
 * memory" and "a label map exists" happens behind these calls.  File:line citations are relative
 * to the reference tree (SamvitJ/Accel @ d1d7bb1).
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers to dense fp32 NCHW data (batch 1), exactly the arrays
 *     the reference's Predictor exchanges (dff_deeplab/core/tester.py:22-35, demo.py:184), except
 *     where a parameter is documented as host memory.
 *   - Calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream).  A handle is bound to one device and is NOT thread-safe: its graph cache, side streams and
 *     profiling state are mutated by every forward, so use one handle per host thread (the reference runs
 *     one predictor pair per GPU thread too, tester.py:309-313).
 *   - Every function returns 0 on success, non-zero on failure; accel_last_error() returns the
 *     message.  No C++ exception crosses this boundary.  There is no CPU fallback: without a CUDA
 *     device accel_create succeeds only far enough to enumerate parameters, and every compute entry
 *     point fails with an error.
 */
#ifndef ACCEL_B200_H_
#define ACCEL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct AccelHandle AccelHandle;

/* Which `accel_<version>` symbol file the cur-frame graph follows
 * (dff_deeplab/symbols/accel_{18,34,50,101}.py; demo.py:126-132 `--version`).
 * ACCEL_VERSION_DFF = the L branch alone: FlowNet + warp + task head (Deep Feature Flow). */
enum { ACCEL_VERSION_DFF = 0, ACCEL_VERSION_18 = 18, ACCEL_VERSION_34 = 34, ACCEL_VERSION_50 = 50,
       ACCEL_VERSION_101 = 101 };

/* accel_create flags */
enum { ACCEL_FLAG_NO_TENSOR_CORES = 1 /* run every contraction on the CUDA-core (fp32 FFMA) kernels */,
       ACCEL_FLAG_NO_GRAPH = 2        /* launch kernels one by one instead of replaying a CUDA graph */ };

typedef struct AccelConfig {
  int version;      /* ACCEL_VERSION_* */
  int height;       /* frame size; multiples of 128 (config.SCALES 1024x2048, dff_deeplab_vid_demo.yaml:5-7) */
  int width;
  int num_classes;  /* 19 (cfg.dataset.NUM_CLASSES) */
  int device;       /* CUDA device ordinal */
  int flags;
} AccelConfig;

/* Replaces Predictor.__init__ -> MutableModule.bind (tester.py:22-30): builds the key-frame and
 * cur-frame plans for the configured frame size. */
int accel_create(const AccelConfig* config, AccelHandle** out);
void accel_destroy(AccelHandle* h);
/* Message of the last failure on `h` (or of the last failed accel_create when h == NULL). */
const char* accel_last_error(const AccelHandle* h);

/* Parameter inventory: the `arg:`/`aux:` tensors of the two reference checkpoints this version
 * reads (lib/utils/load_model.py:15-30, demo.py:192-195), by reference name. */
int accel_param_count(const AccelHandle* h);
int accel_param_info(const AccelHandle* h, int index, const char** name, int64_t shape[4], int* ndim);

/* Replaces MutableModule.init_params(arg_params, aux_params) (tester.py:30) for one tensor.
 * `data` is HOST memory, fp32, dense, in the reference's own layout; shape must match. */
int accel_set_param(AccelHandle* h, const char* name, const float* data, const int64_t* shape, int ndim);
/* Folds BatchNorm statistics, splits and packs the weights, uploads them, allocates workspaces.
 * Fails if a parameter is missing.  Called implicitly by the first forward. */
int accel_finalize(AccelHandle* h);

/* Key-frame graph, get_key_test_symbol (accel_18.py:121-159): R101-DCN -> fc6 -> score -> x16
 * upsampling -> crop, plus the argmax of demo.py:238.
 *   data       in  (1,3,H,W)
 *   feat_out   out (1,2048,H/16,W/16)  `res5c_relu_output`;            may be NULL
 *   score_out  out (1,19,H,W)          `croped_score_output`;          may be NULL (production)
 *   label_out  out (H,W) uint8         argmax over classes (demo.py:238,252); may be NULL */
int accel_key_forward(AccelHandle* h, const float* data, float* feat_out, float* score_out, uint8_t* label_out,
                      void* stream);

/* Cur-frame graph, get_cur_test_symbol (accel_18.py:161-239, accel_101.py:144-193): FlowNet(data,
 * data_key) -> warp(feat_key) -> L head [-> R branch -> fusion] -> argmax (demo.py:243-245).
 *   feat_key   in  (1,2048,H/16,W/16)  previous frame's feature (chained, demo.py:241) or the key
 *                                      frame's (un-chained, tester.py:252-256)
 *   feat_out   out `warping_feat_output`; must not alias feat_key;     may be NULL
 *   score_out  out `correction_output` (18/34/50) or `croped_score_output` (101, dff); may be NULL */
int accel_cur_forward(AccelHandle* h, const float* data, const float* data_key, const float* feat_key,
                      float* feat_out, float* score_out, uint8_t* label_out, void* stream);

/* The same two graphs with the L head commuted through the warp (DFF, Accel-18/34/50 only).  `fc6` is a per-pixel
 * linear map and GridGenerator(warp)+BilinearSampler a per-channel linear one (accel_18.py:174-183), so
 *     fc6(warp(F)) = warp(W_fc6 * F) + b_fc6
 * accel_key_forward_lin additionally emits g_out = W_fc6 * res5c_relu (1,1024,H/16,W/16; no bias, no ReLU);
 * accel_cur_forward_lin warps g_key instead of the 2048-channel feature (half the bytes, no fc6 GEMM) and returns
 * the warped g_out for the next frame of the chained schedule (may be NULL; must not alias g_key).  Same score
 * volume / label map as accel_cur_forward up to fp32 re-association (parity tests hold both to the same 1e-3). */
int accel_key_forward_lin(AccelHandle* h, const float* data, float* feat_out, float* g_out, float* score_out,
                          uint8_t* label_out, void* stream);
int accel_cur_forward_lin(AccelHandle* h, const float* data, const float* data_key, const float* g_key, float* g_out,
                          float* score_out, uint8_t* label_out, void* stream);

/* The correction network of Accel-18/34/50 with its own DeepLab head, alone (accel_18.py:199-227 without the L branch
 * and the fusion; accel_50.py:195-216): the plain DeepLab-<v> segmentation of ONE frame, i.e. what deeplab/test.py's
 * Predictor computes per image (deeplab/core/tester.py:84-85 takes the argmax of its softmax, which is the argmax of
 * this score volume).  BASELINE config 1's network.  score_out (1,19,H,W) may be NULL; label_out (H,W) uint8. */
int accel_rbranch_forward(AccelHandle* h, const float* data, float* score_out, uint8_t* label_out, void* stream);

/* Whole key interval in one call, for callers that hold the interval's frames up front (dff_deeplab/demo.py:165-185
 * preloads the whole clip; precedent for batching frames through the nets: dff_rfcn/demo_batch.py:76-96).
 * accel_plan_interval (before accel_finalize) builds the plan for `interval` frames (2..16) of the CHAINED schedule
 * (demo.py:228-250): frame 0 = key graph, frame t = cur graph with data_key = frame t-1 and feat_key = frame t-1's
 * (warped) feature.  accel_interval_forward runs it: everything that depends on one frame alone (R101 of the key frame,
 * FlowNet of every frame pair, the correction network of every cur frame) runs as concurrent chains, each on its
 * share of the SMs; only the warps and the heads / fusion behind them are sequential.  Same arithmetic per layer as
 * the frame-by-frame calls (tile shapes may differ, so results agree to fp32 re-association, not bit for bit).
 *   frames     HOST array of `interval` DEVICE pointers, each (1,3,H,W) fp32
 *   score_out  NULL, or HOST array of `interval` DEVICE pointers (1,19,H,W), entries may be NULL
 *   label_out  HOST array of `interval` DEVICE pointers (H,W) uint8 */
int accel_plan_interval(AccelHandle* h, int interval);
int accel_interval_forward(AccelHandle* h, const float* const* frames, float* const* score_out, uint8_t* const* label_out,
                           void* stream);

/* Per-layer parity / debugging aid: the internal (split fp16 NHWC) output that layer `op_name` (the reference's layer
 * name, e.g. "res4b7_branch2c", "conv3_1", "fc6") of plan `plan` ("key", "cur", "flow", "interval", ...) holds after the
 * last forward, converted to fp32 NCHW into `out` (DEVICE memory; NULL = only report the shape (frames, C, H, W)). */
int accel_debug_fetch(AccelHandle* h, const char* plan, const char* op_name, float* out, int64_t shape[4], void* stream);

/* CUDA-graph cache of the handle: every distinct set of caller pointers is captured once (a miss costs a stream
 * capture + instantiate, ~100x a replay); callers that see `misses` grow per frame are passing fresh temporaries. */
int accel_graph_cache_stats(const AccelHandle* h, uint64_t* hits, uint64_t* misses);

/* FlowNet-S alone, get_flownet (resnet_v1_101_flownet_deeplab.py:1751-1808): flow_out (1,2,H/16,W/16),
 * channel 0 = dx, 1 = dy in feature-grid pixels, already multiplied by 2.5. */
int accel_flownet(AccelHandle* h, const float* data, const float* data_key, float* flow_out, void* stream);

/* ---- operator-level entry points (no handle; mirror the MXNet built-ins the symbols call) ---- */

/* mx.sym.GridGenerator(data=flow, transform_type='warp') + mx.sym.BilinearSampler(data=feat, grid)
 * (accel_18.py:174-175).  feat, out: (1,C,H,W); flow: (1,2,H,W). */
int accel_warp(const float* feat, const float* flow, float* out, int channels, int height, int width,
               void* stream);

/* The same warp, also emitting the operand the first conv of the fusion head reads (accel_18.py:176-183: the warped
 * feature goes straight into `fc6`): the split-fp16 NHWC copy, out_hi[y][x][c] = fp16(v), out_lo[y][x][c] =
 * fp16(v - out_hi), both (H,W,C) __half, 16-byte aligned, C a multiple of 32.  One pass over `feat`
 * (warp_kernel_fused) where the shape allows, else the warp followed by a layout pass.  `out` (fp32 NCHW) is required. */
int accel_warp_split(const float* feat, const float* flow, float* out, void* out_hi, void* out_lo, int channels,
                     int height, int width, void* stream);

/* Concat(dim=1) -> correction 1x1 (2K -> K, +bias) on the x16-upsampled, cropped score maps, then
 * argmax (accel_18.py:193-197,223-235; demo.py:245).  score_a/score_b: low-res (1,K,h,w) outputs of
 * the `score` / `<v>_score` convs; corr_weight (K,2K) and corr_bias (K) are DEVICE pointers, NULL
 * for both = no fusion (key / Accel-101 tail: score_b ignored).  label: (16h,16w) uint8;
 * score_full: (1,K,16h,16w) or NULL. */
int accel_fuse_argmax(const float* score_a, const float* score_b, const float* corr_weight,
                      const float* corr_bias, int num_classes, int h, int w, uint8_t* label,
                      float* score_full, void* stream);

/* lib/utils/image.py:224-235 transform(im, pixel_means) as demo.py:170-175 applies it to every decoded
 * frame: `bgr_hwc` is the (H,W,3) uint8 BGR image cv2.imread/resize returns, in DEVICE memory;
 * pixel_means_bgr = config.network.PIXEL_MEANS (B, G, R; dff_deeplab_vid_demo.yaml:13-16), HOST doubles.
 * out (1,3,H,W) fp32: out[0,i] = im[:,:,2-i] - pixel_means[2-i], computed in float64 and rounded once to
 * float32 like numpy + mx.nd.array do (bit-exact).  Lets a frame cross PCIe as 3 bytes per pixel. */
int accel_preprocess(const uint8_t* bgr_hwc, int height, int width, const double pixel_means_bgr[3], float* out,
                     void* stream);

/* cv2.resize(im, None, None, fx=im_scale, fy=im_scale, interpolation=cv2.INTER_LINEAR) of lib/utils/image.py:211 (the
 * `resize()` every decoded frame goes through, demo.py:173), on the device, bit for bit (OpenCV's fixed-point separable
 * bilinear; an exact 2x decimation is the rounded 2x2 mean, as in OpenCV).  src/dst: (H,W,3) uint8 BGR in DEVICE memory;
 * accel_resize_size gives the destination size (round-half-even of size * scale, as OpenCV). */
int accel_resize_size(int src_height, int src_width, double fx, double fy, int* dst_height, int* dst_width);
int accel_resize_bgr(const uint8_t* src_hwc, int src_height, int src_width, double fx, double fy, uint8_t* dst_hwc, void* stream);

/* fast_hist(pred, label, n) of dff_deeplab/demo.py:50-53 on the device: for every pixel with label < n,
 * hist[label * n + pred] += 1.  pred/label: `count` uint8 values (label 255 = ignore); hist: n*n int64 in
 * DEVICE memory, ACCUMULATED into (`hist += curr_hist`, demo.py:272) -- zero it before the first frame. */
int accel_confusion(const uint8_t* pred, const uint8_t* label, size_t count, int num_classes, int64_t* hist,
                    void* stream);

/* The DeepLab task head at feature resolution, accel_18.py:177-191 (L branch) / :208-221 (R branch):
 *   relu_fc6 = ReLU(Convolution(feat, fc6_weight (mid,cin,1,1), fc6_bias)),  score = Convolution(relu_fc6,
 *   score_weight (num_classes,mid,1,1), score_bias)
 * feat (1,cin,height,width) and score_lowres (1,num_classes,height,width) are fp32 NCHW in DEVICE memory, the four
 * parameter arrays HOST memory in MXNet layout.  The x16 `upsampling` + Crop + argmax that follow are
 * accel_fuse_argmax.  Builds and runs a two-layer plan on the fly (operator-level entry point: weights are packed
 * per call; the whole-graph entry points keep them resident). */
int accel_head(const float* feat, int cin, int height, int width, const float* fc6_weight, const float* fc6_bias, int mid,
               const float* score_weight, const float* score_bias, int num_classes, float* score_lowres, int device,
               char* err, int errlen);

/* One convolution-like layer through the same kernels the graphs use; parity-test hook.
 *   kind: 0 Convolution, 1 Deconvolution(k4,s2,p1 after crop), 2 DeformableConvolution(3x3,s1),
 *         3 the 7x7/s2/p3 stem over one fp32 NCHW frame (cin 3), 4 FlowNet's stem over the frame pair
 *         (`in`, `offset`), 2x2 average-pooled and divided by 255 first (cin 6; hin/win = frame size)
 *   in (1,cin,hin,win) device; weight/scale/shift HOST (weight in MXNet layout; scale/shift per
 *   output channel, NULL = 1/0); offset (device, deformable only); residual (device, NCHW, or NULL)
 *   act: 0 none, 1 relu, 2 leaky(0.1); engine: 0 auto, 1 CUDA-core, 2 tcgen05
 *   out (1,cout,hout,wout) device. */
int accel_conv_layer(int kind, const float* in, int cin, int hin, int win, const float* weight, int cout,
                     int ksize, int stride, int pad, int dilate, int deform_groups, const float* offset,
                     const float* scale, const float* shift, int act, const float* residual, int engine,
                     float* out, int device, char* err, int errlen);

/* Number of kernel launches issued by the most recent forward on `h` (bench.py's gpu_launches). */
int accel_last_launch_count(const AccelHandle* h);
/* Per-stage device time of the most recent forward, when profiling was enabled with
 * accel_set_profiling(h, 1): fills up to `cap` (name, milliseconds) pairs, returns the count. */
int accel_set_profiling(AccelHandle* h, int enabled);
int accel_stage_times(AccelHandle* h, const char** names, float* ms, int cap);
/* Same, per layer (one entry per op of the plan, in launch order): name = "<stage>/<reference layer name>",
 * device milliseconds, and the layer's reference-graph flops (2*MAC; 0 for non-contractions). */
int accel_op_times(AccelHandle* h, const char** names, float* ms, double* flops, int cap);

#ifdef __cplusplus
}
#endif
#endif /* ACCEL_B200_H_ */
