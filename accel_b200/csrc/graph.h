// Host-side plan builder and executor: turns an Accel version + frame size into ordered kernel
// launches over the split-fp16 NHWC activation format.  Pure host logic until finalize().
#pragma once
#include <map>
#include <set>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.h"

namespace accel {

constexpr int kMaxInterval = 16;     // frames of one key interval the whole-interval plan accepts
enum Ext { X_NONE = 0, X_DATA, X_DATA_KEY, X_FEAT_KEY, X_FEAT_OUT, X_SCORE_OUT, X_LABEL_OUT, X_FLOW_OUT, X_AUX_IN,
           X_AUX_OUT, X_G_KEY, X_G_OUT,              // X_G_*: fc6's linear part W*F (1,1024,h,w) of the commuted L head
           X_FRAME0,                                 // whole-interval plan: frame i of the interval = X_FRAME0 + i
           X_LABEL0 = X_FRAME0 + kMaxInterval,       //   its label map
           X_SCORE0 = X_LABEL0 + kMaxInterval,       //   its score volume (optional)
           X_COUNT = X_SCORE0 + kMaxInterval };

struct Tensor {
  int C = 0, H = 0, W = 0;
  int ld = 0;        // channel stride of the underlying buffer
  int buf = -1;      // buffer index
  int coff = 0;      // channel offset inside the buffer (zero-copy concat views)
  bool f32 = false;  // fp32 planar (C,H,W) instead of split NHWC
  int nb = 1;        // frames stacked densely in the buffer ([nb][H][W][ld] / [nb][C][H][W]): batched plans
  int foff = 0;      // first frame of this view inside the buffer
};

struct Buffer {
  size_t elems = 0;  // fp16 elements per plane, or fp32 elements
  bool f32 = false;
  __half* hi = nullptr;
  __half* lo = nullptr;
  float* f = nullptr;
};

struct ParamSpec {
  std::string name;
  std::vector<int64_t> shape;
};

struct EpiSpec {
  std::string bn;            // BatchNorm after the contraction (name prefix), or
  float eps = 1e-5f;
  std::string bias;          // a bias parameter name
  float mul = 1.f;           // multiplies the result (FlowNet's `* 2.5`)
  int act = ACT_NONE;
  int res = -1;              // residual tensor id (added before act)
  std::string bn2;           // second output: act2(bn2(v))
  float eps2 = 2e-5f;
  int act2 = ACT_RELU;
  int out2 = -1;
  int out_f32 = -1;          // also/only write fp32 planar tensor
  int f32_frames = 0;        // batched plans: only the first f32_frames frames get the fp32 copy (0 = all)
  int ext_out = X_NONE;      // also write an external fp32 NCHW output when the caller passed one
  int ext_raw = X_NONE;      // external fp32 NCHW copy of the LINEAR part (acc * scale: no bias, no activation)
  bool no_split_out = false;
};

enum OpType { OP_STEM, OP_CONV, OP_POOL, OP_DCN_COL, OP_WARP, OP_UPFLOW, OP_FUSE, OP_TAIL, OP_TO_SPLIT, OP_TO_NCHW, OP_COPY_F32 };
enum Engine { ENG_AUTO = 0, ENG_FFMA = 1, ENG_TC = 2, ENG_NARROW = 3 };

struct Op {
  OpType type;
  std::string stage;         // profiling bucket
  std::string name;
  // symbolic operands (tensor ids / parameter names), resolved into the kernel params at finalize
  int in = -1, in2 = -1, out = -1;
  std::string weight, weight2;
  EpiSpec epi;
  int cout = 0, ksize = 0, stride = 1, pad = 0, dilate = 1;
  int kind = 0;              // OP_CONV: 0 conv, 1 deconv phase set (expanded), 2 1x1 over dcn columns
  int phase_y = 0, phase_x = 0;
  int pool_max = 0;
  int dg = 1;
  int ext_in0 = X_NONE, ext_in1 = X_NONE, ext_out = X_NONE, ext_out2 = X_NONE;
  int engine = ENG_AUTO;
  std::string bn_in;         // stem: input BatchNorm (bn_data)
  int stem_pool = 0;
  bool skip = false;         // OP_FUSE whose work the following band tail kernel does itself (TailParams::fuse_*)
  int src_warp = 0;          // OP_TO_SPLIT: read the warp op's fp32 output (caller's feat_out or the scratch)
  int src_f32 = -1;          // OP_WARP / OP_TO_SPLIT: internal fp32 planar source tensor (instead of an external pointer)
  int dst_f32 = -1;          // OP_WARP: internal fp32 planar destination tensor (whole-interval plan: the chained features)
  bool fuse_split = false;   // OP_WARP: the OP_TO_SPLIT that follows is done by the fused warp kernel when the shape allows
  int sm_budget = 0;         // > 0: plan this op's persistent kernels for that many SMs (a chain of the whole-interval plan)
  std::string split_bias;    // OP_TO_SPLIT: per-channel bias added + `split_act` applied during the conversion
  int split_act = 0;
  const float* split_bias_dev = nullptr;
  float stem_in_mul = 1.f;
  // resolved
  ConvParams conv{};
  StemParams stem{};
  PoolParams pool{};
  DcnColParams dcn{};
  WarpParams warp{};
  UpflowParams upflow{};
  FuseParams fuse{};
  TailParams tail{};
  TcPlan* tc = nullptr;
  std::vector<ConvParams> frame_convs;   // a batched tensor run frame by frame (layers the batched kernel cannot take)
  std::vector<TcPlan*> frame_tc;
  int frame = 0;             // OP_STEM of a batched plan: which frame of the output tensor this launch writes
  StemTcPlan* stem_tc = nullptr;
  int partial_buf = -1;
  double flops = 0.0;
  // Parallel section: ops that share `par_group` (> 0) are mutually independent across `par_branch` values and are
  // issued on forked streams (parallel kernel nodes once captured); ops of one branch keep their order.
  // `par_width` = number of branches that run tensor-core convs at once: the planner gives each 1/par_width of the SMs.
  int par_group = 0, par_branch = 0, par_width = 1;
  // Lane 1 = a long chain independent of lane 0 (the correction branch next to FlowNet + warp + L head): issued on its
  // own stream from the start of the plan; the first op with `join_lane` waits for it.
  int lane = 0, join_lane = 0;
};

struct OpTime {
  std::string name;
  float ms;
  double flops;
};

class Graph {
 public:
  explicit Graph(int device, int flags) : device_(device), flags_(flags) {}
  ~Graph();

  // ---- building ----
  int new_tensor(int C, int H, int W, bool f32 = false, int nb = 1);
  int new_view(int base, int coff, int C);
  int frame_view(int base, int first_frame, int nb);     // frames [first_frame, first_frame + nb) of a batched tensor
  int add_param(const std::string& name, std::vector<int64_t> shape);
  std::vector<Op>& seq(const std::string& which) { return seqs_[which]; }

  int stem(std::vector<Op>& s, const std::string& stage, int ext0, int ext1, int Hs, int Ws, bool pool,
           float in_mul, const std::string& bn_in, const std::string& wname, int cin, EpiSpec e, int out = -1, int frame = 0);
  int conv(std::vector<Op>& s, const std::string& stage, int in, const std::string& wname, int cout, int k, int stride,
           int pad, int dil, EpiSpec e, int out = -1);
  // fold_1x1: name of a 1x1 convolution (weight (cout, mid, 1, 1)) that follows the transposed conv (weight
  // (Cin, mid, 4, 4)) with nothing in between: the two linear maps are composed once at finalize into one
  // (Cin, cout, 4, 4) transposed conv (shortcut iii: 18_fc6 o 18_feat_upsampling).
  int deconv4(std::vector<Op>& s, const std::string& stage, int in, const std::string& wname, int cout, EpiSpec e,
              int out = -1, const std::string& fold_1x1 = "", int mid = 0);
  int dcn(std::vector<Op>& s, const std::string& stage, int in, int offset_f32, const std::string& wname, int cout,
          int dg, EpiSpec e);
  int pool(std::vector<Op>& s, const std::string& stage, int in, int k, int stride, int pad, bool is_max, bool full,
           EpiSpec post = EpiSpec());
  void require_bilinear(const std::string& name, int num_classes);
  void warp(std::vector<Op>& s, int ext_feat, int flow_f32, int out_split, int ext_out, const std::string& bias = "",
            int act = 0);
  // whole-interval plan: source and destination are internal fp32 planar tensors (the chained warped features)
  void warp_internal(std::vector<Op>& s, int src_f32, int flow_f32, int dst_f32, int out_split);
  void upflow(std::vector<Op>& s, int flow_f32, const std::string& wname, const std::string& bname, int out_view);
  void fuse(std::vector<Op>& s, int a_f32, int b_f32, const std::string& wname, int out_f32);
  void tail(std::vector<Op>& s, int score_f32, const std::string& bias_name, int ext_label, int ext_score);
  void to_split(std::vector<Op>& s, int ext_in, int out);
  void to_nchw(std::vector<Op>& s, int in, int ext_out);
  void copy_f32(std::vector<Op>& s, int ext_in, int out_f32);

  // ---- parameters ----
  const std::vector<ParamSpec>& params() const { return params_; }
  bool set_param(const std::string& name, const float* data, const int64_t* shape, int ndim, std::string* err);

  // ---- execution ----
  bool finalize(std::string* err);
  bool finalized() const { return finalized_; }
  bool run(const std::string& which, void* const ext[X_COUNT], cudaStream_t stream, std::string* err);
  bool run_eager(const std::string& which, void* const ext[X_COUNT], cudaStream_t stream, std::string* err);
  // Debugging / per-layer parity aid: the split NHWC output of op `op_name` of plan `which` (as left by the last run) as
  // fp32 NCHW into `dst` (device); shape receives (frames, C, H, W).
  bool fetch_op_output(const std::string& which, const std::string& op_name, float* dst, int64_t shape[4], cudaStream_t stream,
                       std::string* err);
  int last_launches() const { return last_launches_; }
  void cache_stats(unsigned long long* hits, unsigned long long* misses) const { *hits = cache_hits_; *misses = cache_misses_; }
  void set_profiling(bool on) { profiling_ = on; }
  const std::vector<std::pair<std::string, float>>& stage_times();
  std::vector<OpTime> op_times();   // per launch group of the last profiled run
  const Tensor& tensor(int id) const { return tensors_[id]; }
  size_t frame_elems(const Tensor& t) const { return t.f32 ? (size_t)t.ld * t.H * t.W : (size_t)t.H * t.W * t.ld; }
  __half* hi_ptr(const Tensor& t) const { return bufs_[t.buf].hi + (size_t)t.foff * frame_elems(t) + t.coff; }
  __half* lo_ptr(const Tensor& t) const { return bufs_[t.buf].lo + (size_t)t.foff * frame_elems(t) + t.coff; }
  float* f_ptr(const Tensor& t) const { return bufs_[t.buf].f + (size_t)t.foff * frame_elems(t); }
  int flags() const { return flags_; }
  int num_sms() const { return num_sms_; }
  int num_sms_hint() const;            // SM count of the handle's device when one is visible (before finalize), else 148

 private:
  bool resolve_conv(Op& op, std::string* err);
  bool make_scale_shift(const EpiSpec& e, int cout, const std::vector<float>& prescale, float** scale, float** shift,
                        std::string* err);
  bool make_scale_shift2(const EpiSpec& e, int cout, float** scale, float** shift, std::string* err);
  const std::vector<float>* host_param(const std::string& name, std::string* err) const;
  float* upload(const std::vector<float>& v);
  void* dev_alloc(size_t bytes);
  ActKind dummy_ = ACT_NONE;

  int device_, flags_;
  int num_sms_ = 148;
  bool finalized_ = false;
  bool profiling_ = false;
  int last_launches_ = 0;
  std::vector<Tensor> tensors_;
  std::vector<Buffer> bufs_;
  std::vector<ParamSpec> params_;
  std::unordered_map<std::string, int> param_index_;
  std::unordered_map<std::string, std::vector<float>> host_;
  std::map<std::string, std::vector<Op>> seqs_;
  std::vector<void*> allocs_;
  // profiling
  std::vector<cudaEvent_t> events_;
  std::vector<std::string> event_stage_;
  std::vector<const Op*> event_op_;
  std::vector<std::pair<std::string, float>> times_;
  std::vector<std::string> bilinear_checks_;
  struct PackedWeights { __half* hi = nullptr; __half* lo = nullptr; std::vector<float> prescale; };
  std::map<std::string, PackedWeights> packed_;          // device copies of packed conv weights, shared between plans
  struct FoldSpec { std::string out, deconv, conv1x1; int cin, mid, cout; };
  std::vector<FoldSpec> folds_;          // derived parameters: out = conv1x1 o deconv, computed at finalize
  int label_scratch_ = -1;
  uint8_t* label_scratch_ptr_ = nullptr;
  size_t label_scratch_bytes_ = 0;
  float* warp_scratch_ = nullptr;
  struct CachedGraph {
    std::string which;
    void* ext[X_COUNT];
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
    unsigned long long stamp = 0;
  };
  std::vector<CachedGraph> graph_cache_;
  std::set<std::string> warmed_;
  cudaStream_t capture_stream_ = nullptr;
  static constexpr int kMaxBranches = 2 * kMaxInterval;
  unsigned long long cache_hits_ = 0, cache_misses_ = 0;
  cudaStream_t side_[kMaxBranches] = {nullptr};
  cudaEvent_t fork_ev_ = nullptr, join_ev_[kMaxBranches] = {nullptr};
  cudaStream_t lane_stream_ = nullptr;
  cudaEvent_t lane_fork_ev_ = nullptr, lane_join_ev_ = nullptr;
  int next_group_ = 0;
 public:
  int new_par_group() { return ++next_group_; }
  bool branches_enabled() const;
 private:
  unsigned long long graph_clock_ = 0;   // fp32 NCHW warped feature when the caller passes no feat_out

 public:
  void need_label_scratch(size_t bytes) { if (bytes > label_scratch_bytes_) label_scratch_bytes_ = bytes; }
};

// Accel graphs (nets.cu)
bool build_accel(Graph& g, int version, int H, int W, int num_classes, std::string* err);
// Whole-interval plan "interval": key frame + (interval-1) chained cur frames issued as ONE launch sequence in which the
// per-frame chains (R101 of the key frame, FlowNet of every frame pair, the correction branch of every cur frame) are
// parallel branches, each planned for its share of the SMs (nets.cu).
bool build_interval(Graph& g, int version, int H, int W, int num_classes, int interval, std::string* err);

}  // namespace accel
