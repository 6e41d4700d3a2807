// Plan builder / executor.  See graph.h.
#include <stdlib.h>
#include "graph.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <mutex>

namespace accel {

bool first_time_on_device(int slot) {
  static std::mutex mu;
  static bool done[ONCE_SLOTS][64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || slot < 0 || slot >= ONCE_SLOTS) return true;
  std::lock_guard<std::mutex> lock(mu);
  const bool first = !done[slot][dev];
  done[slot][dev] = true;
  return first;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ACCEL_PDL");
    v = (e && *e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}


static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

Graph::~Graph() {
  for (auto& kv : seqs_)
    for (auto& op : kv.second) {
      if (op.tc) tc_plan_destroy(op.tc);
      for (TcPlan* p : op.frame_tc)
        if (p) tc_plan_destroy(p);
    }
  for (auto& kv : seqs_)
    for (auto& op : kv.second)
      if (op.stem_tc) stem_tc_plan_destroy(op.stem_tc);
  for (auto& c : graph_cache_) cudaGraphExecDestroy(c.exec);
  if (capture_stream_) cudaStreamDestroy(capture_stream_);
  for (int i = 0; i < kMaxBranches; ++i) {
    if (side_[i]) cudaStreamDestroy(side_[i]);
    if (join_ev_[i]) cudaEventDestroy(join_ev_[i]);
  }
  if (fork_ev_) cudaEventDestroy(fork_ev_);
  if (lane_stream_) cudaStreamDestroy(lane_stream_);
  if (lane_fork_ev_) cudaEventDestroy(lane_fork_ev_);
  if (lane_join_ev_) cudaEventDestroy(lane_join_ev_);
  for (void* p : allocs_) cudaFree(p);
  for (auto e : events_) cudaEventDestroy(e);
}

int Graph::num_sms_hint() const {
  int n = 0, sms = 0;
  if (cudaGetDeviceCount(&n) == cudaSuccess && device_ >= 0 && device_ < n &&
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_) == cudaSuccess && sms > 0)
    return sms;
  cudaGetLastError();
  return 148;
}

int Graph::new_tensor(int C, int H, int W, bool f32, int nb) {
  Buffer b;
  Tensor t;
  t.C = C; t.H = H; t.W = W; t.f32 = f32;
  t.nb = nb < 1 ? 1 : nb;
  t.ld = f32 ? C : round_up(C, 16);   // 32-byte pixel rows: the conv epilogues move 256-bit sectors
  b.f32 = f32;
  b.elems = (size_t)H * W * t.ld * t.nb;
  bufs_.push_back(b);
  t.buf = (int)bufs_.size() - 1;
  tensors_.push_back(t);
  return (int)tensors_.size() - 1;
}

int Graph::new_view(int base, int coff, int C) {
  Tensor t = tensors_[base];
  t.coff = tensors_[base].coff + coff;
  t.C = C;
  tensors_.push_back(t);
  return (int)tensors_.size() - 1;
}

int Graph::frame_view(int base, int first_frame, int nb) {
  Tensor t = tensors_[base];
  t.foff = tensors_[base].foff + first_frame;
  t.nb = nb;
  tensors_.push_back(t);
  return (int)tensors_.size() - 1;
}

int Graph::add_param(const std::string& name, std::vector<int64_t> shape) {
  auto it = param_index_.find(name);
  if (it != param_index_.end()) return it->second;
  params_.push_back({name, std::move(shape)});
  param_index_[name] = (int)params_.size() - 1;
  return (int)params_.size() - 1;
}

static void add_bn_params(Graph& g, const std::string& bn, int c) {
  g.add_param(bn + "_gamma", {c});
  g.add_param(bn + "_beta", {c});
  g.add_param(bn + "_moving_mean", {c});
  g.add_param(bn + "_moving_var", {c});
}

static void register_epi_params(Graph& g, const EpiSpec& e, int cout) {
  if (!e.bn.empty()) add_bn_params(g, e.bn, cout);
  if (!e.bias.empty()) g.add_param(e.bias, {cout});
  if (!e.bn2.empty()) add_bn_params(g, e.bn2, cout);
}

int Graph::stem(std::vector<Op>& s, const std::string& stage, int ext0, int ext1, int Hs, int Ws, bool pool,
                float in_mul, const std::string& bn_in, const std::string& wname, int cin, EpiSpec e, int out, int frame) {
  Op op{};
  op.type = OP_STEM;
  op.stage = stage;
  op.name = wname;
  op.ext_in0 = ext0;
  op.ext_in1 = ext1;
  op.weight = wname + "_weight";
  op.bn_in = bn_in;
  op.stem_pool = pool ? 1 : 0;
  op.stem_in_mul = in_mul;
  op.cout = 64;
  op.ksize = 7;
  const int Hc = pool ? Hs / 2 : Hs, Wc = pool ? Ws / 2 : Ws;
  const int Ho = (Hc + 6 - 7) / 2 + 1, Wo = (Wc + 6 - 7) / 2 + 1;
  if (!bn_in.empty()) add_bn_params(*this, bn_in, cin);      // declared before the conv weight, as the symbol does
  add_param(op.weight, {64, cin, 7, 7});
  register_epi_params(*this, e, 64);
  op.epi = e;
  op.in = cin;            // stems carry the channel count here
  op.stem.Hs = Hs;
  op.stem.Ws = Ws;
  op.out = out >= 0 ? out : new_tensor(64, Ho, Wo);
  op.frame = frame;
  op.flops = 2.0 * 64 * cin * 49 * Ho * Wo;
  s.push_back(op);
  return op.out;
}

int Graph::conv(std::vector<Op>& s, const std::string& stage, int in, const std::string& wname, int cout, int k,
                int stride, int pad, int dil, EpiSpec e, int out) {
  const Tensor ti = tensors_[in];
  Op op{};
  op.type = OP_CONV;
  op.stage = stage;
  op.name = wname;
  op.kind = 0;
  op.in = in;
  op.weight = wname + "_weight";
  op.cout = cout; op.ksize = k; op.stride = stride; op.pad = pad; op.dilate = dil;
  add_param(op.weight, {cout, ti.C, k, k});
  register_epi_params(*this, e, cout);
  const int Ho = (ti.H + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
  const int Wo = (ti.W + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
  if (out < 0 && !e.no_split_out) out = new_tensor(cout, Ho, Wo, false, ti.nb);
  op.out = out;
  op.epi = e;
  op.flops = 2.0 * cout * ti.C * k * k * Ho * Wo * ti.nb;
  op.conv.Ho = Ho;
  op.conv.Wo = Wo;
  s.push_back(op);
  return out;
}

int Graph::deconv4(std::vector<Op>& s, const std::string& stage, int in, const std::string& wname_in, int cout,
                   EpiSpec e, int out, const std::string& fold_1x1, int mid) {
  const Tensor ti = tensors_[in];
  std::string wname = wname_in;
  if (fold_1x1.empty()) {
    add_param(wname + "_weight", {ti.C, cout, 4, 4});
  } else {
    // the caller's two parameters stay the public ones; the composed weight is private to the handle
    add_param(wname_in + "_weight", {ti.C, mid, 4, 4});
    add_param(fold_1x1 + "_weight", {cout, mid, 1, 1});
    wname = fold_1x1 + "*" + wname_in;
    bool seen = false;
    for (const FoldSpec& f : folds_) seen = seen || f.out == wname + "_weight";
    if (!seen) folds_.push_back({wname + "_weight", wname_in + "_weight", fold_1x1 + "_weight", ti.C, mid, cout});
  }
  register_epi_params(*this, e, cout);
  if (out < 0) out = new_tensor(cout, 2 * ti.H, 2 * ti.W, false, ti.nb);
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      Op op{};
      op.type = OP_CONV;
      op.stage = stage;
      op.name = wname;
      op.kind = 1;
      op.phase_y = py; op.phase_x = px;
      op.in = in;
      op.out = out;
      op.weight = wname + "_weight";
      op.cout = cout; op.ksize = 4; op.stride = 1;
      op.epi = e;
      op.flops = 2.0 * cout * ti.C * 4 * ti.H * ti.W * ti.nb;
      op.conv.Ho = ti.H;
      op.conv.Wo = ti.W;
      s.push_back(op);
    }
  return out;
}

int Graph::dcn(std::vector<Op>& s, const std::string& stage, int in, int offset_f32, const std::string& wname,
               int cout, int dg, EpiSpec e) {
  const Tensor ti = tensors_[in];
  const int col = new_tensor(9 * ti.C, ti.H, ti.W, false, ti.nb);
  Op g{};
  g.type = OP_DCN_COL;
  g.stage = stage;
  g.name = wname + "(im2col)";
  g.in = in; g.in2 = offset_f32; g.out = col; g.dg = dg; g.dilate = 2; g.pad = 2;
  s.push_back(g);
  Op op{};
  op.type = OP_CONV;
  op.stage = stage;
  op.name = wname;
  op.kind = 2;
  op.in = col;
  op.weight = wname + "_weight";
  op.cout = cout; op.ksize = 3; op.stride = 1;
  add_param(op.weight, {cout, ti.C, 3, 3});
  register_epi_params(*this, e, cout);
  op.out = new_tensor(cout, ti.H, ti.W, false, ti.nb);
  op.epi = e;
  op.flops = 2.0 * cout * ti.C * 9 * ti.H * ti.W * ti.nb;
  op.conv.Ho = ti.H;
  op.conv.Wo = ti.W;
  s.push_back(op);
  return op.out;
}

void Graph::require_bilinear(const std::string& name, int num_classes) {
  add_param(name, {num_classes, 1, 32, 32});
  bilinear_checks_.push_back(name);
}

int Graph::pool(std::vector<Op>& s, const std::string& stage, int in, int k, int stride, int pad, bool is_max,
                bool full, EpiSpec post) {
  const Tensor ti = tensors_[in];
  auto osz = [&](int n) {
    const int num = n + 2 * pad - k;
    int o = (full ? (num + stride - 1) / stride : num / stride) + 1;
    if (full && (o - 1) * stride >= n + pad) --o;
    return o;
  };
  Op op{};
  op.type = OP_POOL;
  op.stage = stage;
  op.name = is_max ? "maxpool" : "avgpool";
  op.in = in;
  op.ksize = k; op.stride = stride; op.pad = pad; op.pool_max = is_max ? 1 : 0;
  op.epi = post;
  register_epi_params(*this, post, ti.C);
  op.out = new_tensor(ti.C, osz(ti.H), osz(ti.W), false, ti.nb);
  s.push_back(op);
  return op.out;
}

void Graph::warp(std::vector<Op>& s, int ext_feat, int flow_f32, int out_split, int ext_out, const std::string& bias,
                 int act) {
  Op op{};
  op.type = OP_WARP;
  op.stage = "warp";
  op.name = "warping_feat";
  op.ext_in0 = ext_feat;
  op.in = flow_f32;
  op.out = out_split;        // only sizes the scratch; the split copy is the next op
  op.ext_out = ext_out;
  s.push_back(op);
  // fp32 NCHW warped feature -> split NHWC input of the task head (reads mostly hit L2)
  Op cv{};
  cv.type = OP_TO_SPLIT;
  cv.stage = "warp_to_head";
  cv.name = "warping_feat(nchw->split)";
  cv.ext_in0 = ext_out;
  cv.src_warp = 1;
  cv.out = out_split;
  cv.split_bias = bias;       // commuted L head: relu(warp(W*F) + fc6_bias) while converting
  cv.split_act = act;
  if (!bias.empty()) add_param(bias, {tensors_[out_split].C});
  s.push_back(cv);
}

void Graph::warp_internal(std::vector<Op>& s, int src_f32, int flow_f32, int dst_f32, int out_split) {
  Op op{};
  op.type = OP_WARP;
  op.stage = "warp";
  op.name = "warping_feat";
  op.in = flow_f32;
  op.out = out_split;
  op.src_f32 = src_f32;
  op.dst_f32 = dst_f32;
  s.push_back(op);
  Op cv{};
  cv.type = OP_TO_SPLIT;
  cv.stage = "warp_to_head";
  cv.name = "warping_feat(nchw->split)";
  cv.src_f32 = dst_f32;
  cv.out = out_split;
  s.push_back(cv);
}

void Graph::upflow(std::vector<Op>& s, int flow_f32, const std::string& wname, const std::string& bname,
                   int out_view) {
  Op op{};
  op.type = OP_UPFLOW;
  op.stage = "flownet";
  op.name = wname;
  op.in = flow_f32;
  op.out = out_view;
  op.weight = wname + "_weight";
  op.weight2 = bname;
  add_param(op.weight, {2, 2, 4, 4});
  add_param(bname, {2});
  s.push_back(op);
}

void Graph::fuse(std::vector<Op>& s, int a_f32, int b_f32, const std::string& wname, int out_f32) {
  Op op{};
  op.type = OP_FUSE;
  op.stage = "tail";
  op.name = wname;
  op.in = a_f32; op.in2 = b_f32; op.out = out_f32;
  op.weight = wname + "_weight";
  const int K = tensors_[a_f32].C;
  add_param(op.weight, {K, 2 * K, 1, 1});
  s.push_back(op);
}

void Graph::tail(std::vector<Op>& s, int score_f32, const std::string& bias_name, int ext_label, int ext_score) {
  Op op{};
  op.type = OP_TAIL;
  op.stage = "tail";
  op.name = "upsample+argmax";
  op.in = score_f32;
  op.weight2 = bias_name;
  if (!bias_name.empty()) add_param(bias_name, {tensors_[score_f32].C});
  op.ext_out = ext_label;
  op.ext_out2 = ext_score;
  need_label_scratch((size_t)tensors_[score_f32].H * 16 * tensors_[score_f32].W * 16);
  s.push_back(op);
}

void Graph::to_split(std::vector<Op>& s, int ext_in, int out) {
  Op op{};
  op.type = OP_TO_SPLIT;
  op.stage = "io";
  op.name = "nchw->split";
  op.ext_in0 = ext_in;
  op.out = out;
  s.push_back(op);
}

void Graph::to_nchw(std::vector<Op>& s, int in, int ext_out) {
  Op op{};
  op.type = OP_TO_NCHW;
  op.stage = "io";
  op.name = "split->nchw";
  op.in = in;
  op.ext_out = ext_out;
  s.push_back(op);
}

void Graph::copy_f32(std::vector<Op>& s, int ext_in, int out_f32) {
  Op op{};
  op.type = OP_COPY_F32;
  op.stage = "io";
  op.name = "copy";
  op.ext_in0 = ext_in;
  op.out = out_f32;
  s.push_back(op);
}

// -------------------------------------------------------------------------------------------------
bool Graph::set_param(const std::string& name, const float* data, const int64_t* shape, int ndim, std::string* err) {
  auto it = param_index_.find(name);
  if (it == param_index_.end()) {
    *err = "unknown parameter '" + name + "'";
    return false;
  }
  const ParamSpec& ps = params_[it->second];
  bool ok = (int)ps.shape.size() == ndim;
  size_t n = 1;
  for (int i = 0; ok && i < ndim; ++i) {
    ok = ps.shape[i] == shape[i];
    n *= (size_t)shape[i];
  }
  if (!ok) {
    *err = "shape mismatch for parameter '" + name + "'";     // cf. check_parameter_shapes, lib/utils/symbol.py:43-55
    return false;
  }
  if (finalized_) {
    *err = "parameters are frozen after accel_finalize";
    return false;
  }
  host_[name].assign(data, data + n);
  return true;
}

const std::vector<float>* Graph::host_param(const std::string& name, std::string* err) const {
  auto it = host_.find(name);
  if (it == host_.end()) {
    *err = "parameter '" + name + "' was never set";
    return nullptr;
  }
  return &it->second;
}

void* Graph::dev_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  allocs_.push_back(p);
  return p;
}

float* Graph::upload(const std::vector<float>& v) {
  float* d = (float*)dev_alloc(v.size() * sizeof(float));
  if (d) cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice);
  return d;
}

static bool bn_fold(const Graph& g, const std::unordered_map<std::string, std::vector<float>>& host,
                    const std::string& bn, float eps, bool fix_gamma, int c, std::vector<float>& scale,
                    std::vector<float>& shift, std::string* err) {
  const char* suf[4] = {"_gamma", "_beta", "_moving_mean", "_moving_var"};
  const std::vector<float>* v[4];
  for (int i = 0; i < 4; ++i) {
    auto it = host.find(bn + suf[i]);
    if (it == host.end() || (int)it->second.size() != c) {
      *err = "parameter '" + bn + suf[i] + "' was never set";
      return false;
    }
    v[i] = &it->second;
  }
  scale.resize(c);
  shift.resize(c);
  for (int i = 0; i < c; ++i) {
    const float gamma = fix_gamma ? 1.f : (*v[0])[i];
    const float inv = gamma / sqrtf((*v[3])[i] + eps);
    scale[i] = inv;
    shift[i] = (*v[1])[i] - (*v[2])[i] * inv;
  }
  (void)g;
  return true;
}

bool Graph::make_scale_shift(const EpiSpec& e, int cout, const std::vector<float>& prescale, float** scale,
                             float** shift, std::string* err) {
  std::vector<float> sc(cout, 1.f), sh(cout, 0.f);
  if (!e.bn.empty()) {
    if (!bn_fold(*this, host_, e.bn, e.eps, false, cout, sc, sh, err)) return false;
  } else if (!e.bias.empty()) {
    const std::vector<float>* b = host_param(e.bias, err);
    if (!b) return false;
    sh = *b;
  }
  for (int i = 0; i < cout; ++i) {
    sc[i] *= e.mul * (prescale.empty() ? 1.f : prescale[i]);
    sh[i] *= e.mul;
  }
  *scale = upload(sc);
  *shift = upload(sh);
  return *scale && *shift;
}

bool Graph::make_scale_shift2(const EpiSpec& e, int cout, float** scale, float** shift, std::string* err) {
  std::vector<float> sc, sh;
  if (!bn_fold(*this, host_, e.bn2, e.eps2, false, cout, sc, sh, err)) return false;
  *scale = upload(sc);
  *shift = upload(sh);
  return *scale && *shift;
}

// Packs one conv's weights K-major per output channel, split into fp16 hi/lo after scaling each
// row by a power of two so that its largest entry sits in [1, 2) (keeps `lo` out of the fp16
// subnormal range; the factor goes back in through the epilogue scale).
static void pack_weights(const float* w, int kind, int cout, int cin, int k, int py, int px, int cout_pad,
                         int cin_pad, int ntaps, std::vector<__half>& hi, std::vector<__half>& lo,
                         std::vector<float>& prescale) {
  const size_t kpad = (size_t)ntaps * cin_pad;
  hi.assign((size_t)cout_pad * kpad, __float2half_rn(0.f));
  lo.assign((size_t)cout_pad * kpad, __float2half_rn(0.f));
  prescale.assign(cout, 1.f);
  std::vector<float> row(kpad);
  for (int n = 0; n < cout; ++n) {
    std::fill(row.begin(), row.end(), 0.f);
    if (kind == 0) {
      for (int c = 0; c < cin; ++c)
        for (int t = 0; t < k * k; ++t) row[(size_t)t * cin_pad + c] = w[((size_t)n * cin + c) * k * k + t];
    } else if (kind == 1) {
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          const int ky = py == 0 ? (a == 0 ? 1 : 3) : (a == 0 ? 2 : 0);
          const int kx = px == 0 ? (b == 0 ? 1 : 3) : (b == 0 ? 2 : 0);
          const int t = a * 2 + b;
          for (int c = 0; c < cin; ++c)
            row[(size_t)t * cin_pad + c] = w[(((size_t)c * cout + n) * 4 + ky) * 4 + kx];
        }
    } else {   // 1x1 over deformable columns: K index = tap*C + c; `cin` here is C
      for (int c = 0; c < cin; ++c)
        for (int t = 0; t < 9; ++t) row[(size_t)t * cin + c] = w[((size_t)n * cin + c) * 9 + t];
    }
    float m = 0.f;
    for (float v : row) m = fmaxf(m, fabsf(v));
    float mult = 1.f;
    if (m > 0.f && isfinite(m)) {
      const int ex = ilogbf(m);
      mult = ldexpf(1.f, -ex);
      prescale[n] = ldexpf(1.f, ex);
    }
    __half* ph = hi.data() + (size_t)n * kpad;
    __half* pl = lo.data() + (size_t)n * kpad;
    for (size_t i = 0; i < kpad; ++i) {
      const float v = row[i] * mult;
      const __half h = __float2half_rn(v);
      ph[i] = h;
      pl[i] = __float2half_rn(v - __half2float(h));
    }
  }
}

bool Graph::resolve_conv(Op& op, std::string* err) {
  const Tensor& ti = tensors_[op.in];
  ConvParams& P = op.conv;
  const int Ho = P.Ho, Wo = P.Wo;
  P = ConvParams{};
  P.Ho = Ho; P.Wo = Wo;
  P.in_hi = hi_ptr(ti);
  P.in_lo = lo_ptr(ti);
  P.in_ld = ti.ld; P.Hin = ti.H; P.Win = ti.W; P.Cin = ti.C;
  P.stride = op.stride;
  int wcin = ti.C;            // channel count in the weight tensor
  if (op.kind == 0) {
    P.ntaps = op.ksize * op.ksize;
    if (P.ntaps > kMaxTaps) { *err = "kernel too large for the generic conv: " + op.name; return false; }
    for (int ky = 0; ky < op.ksize; ++ky)
      for (int kx = 0; kx < op.ksize; ++kx) {
        P.dy[ky * op.ksize + kx] = (int8_t)(ky * op.dilate - op.pad);
        P.dx[ky * op.ksize + kx] = (int8_t)(kx * op.dilate - op.pad);
      }
    P.Cin_pad = round_up(ti.C, 64);
  } else if (op.kind == 1) {
    P.ntaps = 4;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        P.dy[a * 2 + b] = (int8_t)(op.phase_y == 0 ? (a == 0 ? 0 : -1) : (a == 0 ? 0 : 1));
        P.dx[a * 2 + b] = (int8_t)(op.phase_x == 0 ? (b == 0 ? 0 : -1) : (b == 0 ? 0 : 1));
      }
    P.Cin_pad = round_up(ti.C, 64);
  } else {
    P.ntaps = 1;
    P.dy[0] = P.dx[0] = 0;
    wcin = ti.C / 9;
    P.Cin_pad = round_up(ti.C, 64);
  }
  P.Kpad = P.ntaps * P.Cin_pad;
  P.Cout_pad = round_up(op.cout, 64);

  const std::vector<float>* w = host_param(op.weight, err);
  if (!w) return false;
  // One packed device copy per (parameter, packing): the key / cur / cur_lin plans repeat FlowNet and the R branch,
  // and Accel-101's correction net shares the key net's weights (accel_101.py:161) -- packing them once also lets
  // concurrently running plans hit the same lines in L2.
  const std::string pkey = op.weight + "|" + std::to_string(op.kind) + "|" + std::to_string(op.phase_y) + std::to_string(op.phase_x) +
                           "|" + std::to_string(op.cout) + "|" + std::to_string(wcin) + "|" + std::to_string(op.ksize) + "|" +
                           std::to_string(P.Cout_pad) + "|" + std::to_string(P.Cin_pad) + "|" + std::to_string(P.ntaps);
  auto pit = packed_.find(pkey);
  if (pit == packed_.end()) {
    PackedWeights pw;
    std::vector<__half> hi, lo;
    pack_weights(w->data(), op.kind, op.cout, wcin, op.ksize, op.phase_y, op.phase_x, P.Cout_pad,
                 op.kind == 2 ? P.Cin_pad : P.Cin_pad, P.ntaps, hi, lo, pw.prescale);
    pw.hi = (__half*)dev_alloc(hi.size() * sizeof(__half));
    pw.lo = (__half*)dev_alloc(lo.size() * sizeof(__half));
    if (!pw.hi || !pw.lo) { *err = "out of device memory packing " + op.name; return false; }
    cudaMemcpy(pw.hi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice);
    cudaMemcpy(pw.lo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice);
    pit = packed_.emplace(pkey, std::move(pw)).first;
  }
  const std::vector<float>& prescale = pit->second.prescale;
  P.w_hi = pit->second.hi;
  P.w_lo = pit->second.lo;

  Epilogue& E = P.epi;
  E = Epilogue{};
  float *sc = nullptr, *sh = nullptr;
  if (!make_scale_shift(op.epi, op.cout, prescale, &sc, &sh, err)) return false;
  E.scale = sc; E.shift = sh; E.act = op.epi.act; E.Cout = op.cout;
  if (op.epi.res >= 0) {
    const Tensor& tr = tensors_[op.epi.res];
    E.res_hi = hi_ptr(tr);
    E.res_lo = lo_ptr(tr);
    E.res_ld = tr.ld;
  }
  if (op.out >= 0) {
    const Tensor& to = tensors_[op.out];
    E.out_hi = hi_ptr(to);
    E.out_lo = lo_ptr(to);
    E.out_ld = to.ld;
  }
  if (op.epi.out2 >= 0) {
    const Tensor& t2 = tensors_[op.epi.out2];
    float *s2 = nullptr, *h2 = nullptr;
    if (!make_scale_shift2(op.epi, op.cout, &s2, &h2, err)) return false;
    E.scale2 = s2; E.shift2 = h2; E.act2 = op.epi.act2;
    E.out2_hi = hi_ptr(t2);
    E.out2_lo = lo_ptr(t2);
    E.out2_ld = t2.ld;
  }
  if (op.epi.out_f32 >= 0) E.out_nchw = f_ptr(tensors_[op.epi.out_f32]);
  if (op.kind == 1) {
    E.osy = E.osx = 2; E.ooy = op.phase_y; E.oox = op.phase_x;
    E.OHf = 2 * ti.H; E.OWf = 2 * ti.W;
  } else {
    E.osy = E.osx = 1; E.ooy = E.oox = 0;
    E.OHf = Ho; E.OWf = Wo;
  }
  E.nchw_hw = E.OHf * E.OWf;
  E.nchw_nb = op.epi.f32_frames;
  P.ext_outputs = (op.epi.ext_out != X_NONE || op.epi.ext_raw != X_NONE) ? 1 : 0;

  // frames of a batched plan: every operand holds the same number of frames, stacked densely
  const int nb = ti.nb;
  {
    bool same = true;
    if (op.out >= 0) same = same && tensors_[op.out].nb == nb;
    if (op.epi.res >= 0) same = same && tensors_[op.epi.res].nb == nb;
    if (op.epi.out2 >= 0) same = same && tensors_[op.epi.out2].nb == nb;
    if (op.epi.out_f32 >= 0) same = same && (tensors_[op.epi.out_f32].nb == nb || (op.epi.f32_frames > 0 && tensors_[op.epi.out_f32].nb >= op.epi.f32_frames));
    if (!same) { *err = "operands of '" + op.name + "' disagree on the number of frames"; return false; }
    if (nb > 1 && (op.epi.ext_out != X_NONE || op.epi.ext_raw != X_NONE)) { *err = "batched layer '" + op.name + "' cannot write a caller tensor"; return false; }
  }

  // engine
  int eng = op.engine;
  if (eng == ENG_AUTO) {
    if (!(flags_ & 1) && tc_supported(P)) eng = ENG_TC;
    else if (op.cout <= 8) eng = ENG_NARROW;
    else eng = ENG_FFMA;
  }
  op.engine = eng;
  const int sm_budget = op.sm_budget > 0 ? std::min(op.sm_budget, num_sms_)
                        : (op.par_group && op.par_width > 1 && branches_enabled()) ? std::max(8, num_sms_ / op.par_width) : num_sms_;
  auto plan_one = [&](ConvParams& Q, TcPlan** plan) -> bool {
    if (eng == ENG_TC) {
      char msg[256] = {0};
      *plan = tc_plan_create(Q, sm_budget, msg, sizeof(msg));
      if (!*plan) { *err = std::string("tcgen05 plan failed for ") + op.name + ": " + msg; return false; }
      const size_t pb = tc_plan_partial_bytes(*plan);
      if (pb) {
        float* part = (float*)dev_alloc(pb);
        if (!part) { *err = "out of device memory (split-K workspace)"; return false; }
        tc_plan_set_partial(*plan, part);
      }
    } else if (eng == ENG_FFMA) {
      Q.splits = ffma_pick_splits(Q, num_sms_);
      if (Q.splits > 1) {
        Q.partial = (float*)dev_alloc(ffma_partial_bytes(Q, Q.splits));
        if (!Q.partial) { *err = "out of device memory (split-K workspace)"; return false; }
      }
    } else {
      Q.splits = 1;
    }
    return true;
  };
  P.nb = nb;
  static const bool no_batch = [] { const char* e = getenv("ACCEL_TC_BATCH"); return e && e[0] == '0'; }();
  if (nb > 1 && eng == ENG_TC && tc_batchable(P) && !no_batch) {
    // ONE launch over all frames: the loop space is the frames stacked along H
    E.OHf *= nb;
    return plan_one(P, &op.tc);
  }
  if (nb == 1) return plan_one(P, &op.tc);
  // frame by frame (a layer the batched kernel cannot take: padded stride-2 convs, partial tiles, CUDA-core engines)
  P.nb = 1;
  op.frame_convs.assign(nb, P);
  op.frame_tc.assign(nb, nullptr);
  for (int b = 0; b < nb; ++b) {
    ConvParams& Q = op.frame_convs[b];
    Q.in_hi += (size_t)b * frame_elems(ti); Q.in_lo += (size_t)b * frame_elems(ti);
    Epilogue& F = Q.epi;
    if (op.out >= 0) { const size_t fs = frame_elems(tensors_[op.out]); F.out_hi += b * fs; F.out_lo += b * fs; }
    if (op.epi.res >= 0) { const size_t fs = frame_elems(tensors_[op.epi.res]); F.res_hi += b * fs; F.res_lo += b * fs; }
    if (op.epi.out2 >= 0) { const size_t fs = frame_elems(tensors_[op.epi.out2]); F.out2_hi += b * fs; F.out2_lo += b * fs; }
    if (F.out_nchw) {
      if (op.epi.f32_frames > 0 && b >= op.epi.f32_frames) F.out_nchw = nullptr;
      else F.out_nchw += (size_t)b * op.cout * F.nchw_hw;
    }
    F.nchw_nb = 0;
    if (!plan_one(Q, &op.frame_tc[b])) return false;
  }
  return true;
}

bool Graph::finalize(std::string* err) {
  if (finalized_) return true;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    *err = "no CUDA device: accel_b200 has no CPU fallback";
    return false;
  }
  if (cudaSetDevice(device_) != cudaSuccess) { *err = "cudaSetDevice failed"; return false; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device_);
  num_sms_ = prop.multiProcessorCount;
  if (prop.major != 10) {
    *err = "accel_b200 is built for sm_100a only (found sm_" + std::to_string(prop.major * 10 + prop.minor) + ")";
    return false;
  }
  for (auto& b : bufs_) {
    if (b.f32) {
      b.f = (float*)dev_alloc(b.elems * sizeof(float));
      if (!b.f) { *err = "out of device memory"; return false; }
      cudaMemset(b.f, 0, b.elems * sizeof(float));
    } else {
      b.hi = (__half*)dev_alloc(b.elems * sizeof(__half));
      b.lo = (__half*)dev_alloc(b.elems * sizeof(__half));
      if (!b.hi || !b.lo) { *err = "out of device memory"; return false; }
      cudaMemset(b.hi, 0, b.elems * sizeof(__half));
      cudaMemset(b.lo, 0, b.elems * sizeof(__half));
    }
  }
  if (label_scratch_bytes_) {
    label_scratch_ptr_ = (uint8_t*)dev_alloc(label_scratch_bytes_);
    if (!label_scratch_ptr_) { *err = "out of device memory"; return false; }
  }
  // The x16 score upsampling is executed as its closed form; insist the checkpoint's fixed kernel
  // really is the MXNet bilinear initialiser (deeplab/symbols/resnet_v1_101_deeplab.py:820-828).
  for (const std::string& name : bilinear_checks_) {
    const std::vector<float>* w = host_param(name, err);
    if (!w) return false;
    const size_t n = w->size() / 1024;
    for (size_t c = 0; c < n; ++c)
      for (int y = 0; y < 32; ++y)
        for (int x = 0; x < 32; ++x) {
          const float ref = (1.f - fabsf(x / 16.f - 31.f / 32.f)) * (1.f - fabsf(y / 16.f - 31.f / 32.f));
          if (fabsf((*w)[c * 1024 + y * 32 + x] - ref) > 1e-6f) {
            *err = "'" + name + "' is not the fixed bilinear x16 kernel; the fused upsampling tail does not apply";
            return false;
          }
        }
  }
  // derived parameters: (1x1 conv) o (4x4/s2 transposed conv) composed into one transposed conv, on the device in
  // fp32 products with fp64 accumulation
  for (const FoldSpec& f : folds_) {
    const std::vector<float>* wd = host_param(f.deconv, err);
    const std::vector<float>* wc = wd ? host_param(f.conv1x1, err) : nullptr;
    if (!wd || !wc) return false;
    float *dd = nullptr, *dc = nullptr, *dout = nullptr;
    const size_t nout = (size_t)f.cin * f.cout * 16;
    bool ok = cudaMalloc(&dd, wd->size() * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&dc, wc->size() * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&dout, nout * sizeof(float)) == cudaSuccess;
    std::vector<float>& composed = host_[f.out];
    composed.resize(nout);
    ok = ok && cudaMemcpy(dd, wd->data(), wd->size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(dc, wc->data(), wc->size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess &&
         launch_fold_deconv_1x1(dd, dc, dout, f.cin, f.mid, f.cout, nullptr) == cudaSuccess &&
         cudaMemcpy(composed.data(), dout, nout * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(dd); cudaFree(dc); cudaFree(dout);
    if (!ok) { *err = "composing '" + f.out + "' failed: " + cudaGetErrorString(cudaGetLastError()); return false; }
  }
  for (auto& kv : seqs_) {
    for (auto& op : kv.second) {
      switch (op.type) {
        case OP_CONV:
          if (!resolve_conv(op, err)) return false;
          break;
        case OP_STEM: {
          StemParams& S = op.stem;
          const int cin = op.in;
          const int Hs = S.Hs, Ws = S.Ws;
          S = StemParams{};
          S.Hs = Hs; S.Ws = Ws; S.pool = op.stem_pool; S.Cin = cin;
          for (int c = 0; c < 6; ++c) { S.in_scale[c] = op.stem_in_mul; S.in_shift[c] = 0.f; }
          if (!op.bn_in.empty()) {
            std::vector<float> sc, sh;
            if (!bn_fold(*this, host_, op.bn_in, 2e-5f, true, cin, sc, sh, err)) return false;   // fix_gamma=True, eps 2e-5
            for (int c = 0; c < cin; ++c) { S.in_scale[c] = sc[c]; S.in_shift[c] = sh[c]; }
          }
          const std::vector<float>* w = host_param(op.weight, err);
          if (!w) return false;
          S.weight = upload(*w);
          const Tensor& to = tensors_[op.out];
          S.Ho = to.H; S.Wo = to.W;
          const bool use_tc = !(flags_ & 1) && stem_tc_supported(S);
          std::vector<float> prescale;
          __half *dhi = nullptr, *dlo = nullptr;
          if (use_tc) {
            std::vector<__half> hi, lo;
            stem_tc_pack_weights(w->data(), cin, hi, lo, prescale);
            dhi = (__half*)dev_alloc(hi.size() * sizeof(__half));
            dlo = (__half*)dev_alloc(lo.size() * sizeof(__half));
            if (!dhi || !dlo) { *err = "out of device memory packing " + op.name; return false; }
            cudaMemcpy(dhi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice);
            cudaMemcpy(dlo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice);
          }
          float *sc = nullptr, *sh = nullptr;
          if (!make_scale_shift(op.epi, 64, prescale, &sc, &sh, err)) return false;
          Epilogue& E = S.epi;
          E.scale = sc; E.shift = sh; E.act = op.epi.act; E.Cout = 64;
          E.out_hi = hi_ptr(to) + (size_t)op.frame * frame_elems(to);
          E.out_lo = lo_ptr(to) + (size_t)op.frame * frame_elems(to);
          E.out_ld = to.ld;
          E.osy = E.osx = 1; E.OHf = to.H; E.OWf = to.W;
          if (use_tc) {
            char msg[256] = {0};
            op.stem_tc = stem_tc_plan_create(S, dhi, dlo, op.sm_budget > 0 ? std::min(op.sm_budget, num_sms_) : num_sms_, msg, sizeof(msg));
            if (!op.stem_tc) { *err = std::string("tcgen05 stem plan failed for ") + op.name + ": " + msg; return false; }
          }
          break;
        }
        case OP_POOL: {
          const Tensor& ti = tensors_[op.in];
          const Tensor& to = tensors_[op.out];
          PoolParams& Q = op.pool;
          Q.in_hi = hi_ptr(ti); Q.in_lo = lo_ptr(ti);
          Q.in_ld = ti.ld; Q.Hin = ti.H; Q.Win = ti.W; Q.C = ti.C;
          Q.out_hi = hi_ptr(to); Q.out_lo = lo_ptr(to);
          Q.out_ld = to.ld; Q.Ho = to.H; Q.Wo = to.W;
          Q.kernel = op.ksize; Q.stride = op.stride; Q.pad = op.pad; Q.is_max = op.pool_max;
          Q.scale = Q.shift = nullptr; Q.act = op.epi.act;
          if (!op.epi.bn.empty()) {
            float *sc = nullptr, *sh = nullptr;
            if (!make_scale_shift(op.epi, ti.C, {}, &sc, &sh, err)) return false;
            Q.scale = sc; Q.shift = sh;
          }
          break;
        }
        case OP_DCN_COL: {
          const Tensor& ti = tensors_[op.in];
          const Tensor& tf = tensors_[op.in2];
          const Tensor& to = tensors_[op.out];
          DcnColParams& D = op.dcn;
          D.in_hi = hi_ptr(ti); D.in_lo = lo_ptr(ti);
          D.in_ld = ti.ld; D.H = ti.H; D.W = ti.W; D.C = ti.C;
          D.offset = f_ptr(tf);
          D.dg = op.dg; D.dilate = op.dilate; D.pad = op.pad;
          D.col_hi = hi_ptr(to); D.col_lo = lo_ptr(to); D.col_ld = to.ld;
          if (ti.C % (8 * op.dg) != 0) { *err = "deformable conv needs C % (8*dg) == 0"; return false; }
          break;
        }
        case OP_WARP: {
          const Tensor& tf = tensors_[op.in];
          WarpParams& Wp = op.warp;
          if (op.src_f32 >= 0) {                               // whole-interval plan: internal fp32 source / destination
            const Tensor& ts = tensors_[op.src_f32];
            Wp.feat = f_ptr(ts);
            Wp.out_nchw = f_ptr(tensors_[op.dst_f32]);
            Wp.flow = f_ptr(tf);
            Wp.C = ts.C; Wp.H = tf.H; Wp.W = tf.W;
            break;
          }
          if (!warp_scratch_ && op.out >= 0) {
            const Tensor& to = tensors_[op.out];
            warp_scratch_ = (float*)dev_alloc((size_t)to.C * to.H * to.W * sizeof(float));
            if (!warp_scratch_) { *err = "out of device memory"; return false; }
          }
          Wp.flow = f_ptr(tf);
          Wp.H = tf.H; Wp.W = tf.W;
          if (op.out >= 0) Wp.C = tensors_[op.out].C;
          break;
        }
        case OP_UPFLOW: {
          const Tensor& tf = tensors_[op.in];
          const Tensor& to = tensors_[op.out];
          UpflowParams& U = op.upflow;
          U.flow = f_ptr(tf); U.H = tf.H; U.W = tf.W;
          const std::vector<float>* w = host_param(op.weight, err);
          const std::vector<float>* b = w ? host_param(op.weight2, err) : nullptr;
          if (!w || !b) return false;
          memcpy(U.weight, w->data(), sizeof(U.weight));
          memcpy(U.bias, b->data(), sizeof(U.bias));
          U.out_hi = hi_ptr(to); U.out_lo = lo_ptr(to); U.out_ld = to.ld;
          break;
        }
        case OP_FUSE: {
          const Tensor& ta = tensors_[op.in];
          FuseParams& F = op.fuse;
          F.a = f_ptr(ta); F.b = f_ptr(tensors_[op.in2]); F.out = f_ptr(tensors_[op.out]);
          F.K = ta.C; F.h = ta.H; F.w_ = ta.W;
          const std::vector<float>* w = host_param(op.weight, err);
          if (!w) return false;
          F.w = upload(*w);
          break;
        }
        case OP_TAIL: {
          const Tensor& ts = tensors_[op.in];
          TailParams& T = op.tail;
          T.score = f_ptr(ts); T.K = ts.C; T.h = ts.H; T.w = ts.W; T.factor = 16;
          T.bias = nullptr;
          if (!op.weight2.empty()) {
            const std::vector<float>* b = host_param(op.weight2, err);
            if (!b) return false;
            T.bias = upload(*b);
          }
          break;
        }
        case OP_TO_SPLIT:
          if (!op.split_bias.empty()) {
            const std::vector<float>* b = host_param(op.split_bias, err);
            if (!b) return false;
            op.split_bias_dev = upload(*b);
          }
          break;
        case OP_TO_NCHW:
        case OP_COPY_F32:
          break;
      }
    }
  }
  // score-level fusion directly followed by the tail: the band tail kernel applies `corr_weight` to its own source
  // window, so the separate low-resolution launch is dropped (ACCEL_TAIL_FUSE=0 keeps it)
  {
    const char* tf = getenv("ACCEL_TAIL_FUSE");
    if (!(tf && tf[0] == '0'))
      for (auto& kv : seqs_) {
        std::vector<Op>& ops = kv.second;
        for (size_t i = 0; i + 1 < ops.size(); ++i) {
          Op &f = ops[i], &t = ops[i + 1];
          if (f.type != OP_FUSE || t.type != OP_TAIL || t.in != f.out || !tail_band_supported(t.tail.K, t.tail.factor)) continue;
          f.skip = true;
          t.tail.fuse_a = f.fuse.a; t.tail.fuse_b = f.fuse.b; t.tail.fuse_w = f.fuse.w;
        }
      }
  }
  // warp followed by the layout conversion of its own output: the fused warp kernel writes the split NHWC copy itself
  // (decided per launch: the source pointer is the caller's)
  for (auto& kv : seqs_) {
    std::vector<Op>& ops = kv.second;
    for (size_t i = 0; i + 1 < ops.size(); ++i) {
      Op &wp = ops[i], &cv = ops[i + 1];
      if (wp.type != OP_WARP || cv.type != OP_TO_SPLIT) continue;
      if (!(cv.src_warp || (cv.src_f32 >= 0 && cv.src_f32 == wp.dst_f32))) continue;
      const Tensor& to = tensors_[cv.out];
      wp.warp.out_hi = hi_ptr(to);
      wp.warp.out_lo = lo_ptr(to);
      wp.warp.out_ld = to.ld;
      wp.warp.bias = cv.split_bias_dev;
      wp.warp.act = cv.split_act;
      if (wp.warp.C == 0) wp.warp.C = to.C;
      wp.fuse_split = true;
    }
  }
  if (cudaDeviceSynchronize() != cudaSuccess) {
    *err = std::string("CUDA error during finalize: ") + cudaGetErrorString(cudaGetLastError());
    return false;
  }
  host_.clear();
  for (auto& kv : packed_) std::vector<float>().swap(kv.second.prescale);
  finalized_ = true;
  return true;
}

// CUDA-graph front end: the launch sequence of a plan is fixed once finalized, so each distinct set of
// caller pointers is captured once (on the handle's own capture stream: the legacy default stream cannot
// be captured) and replayed with one cudaGraphLaunch on the caller's stream afterwards.
bool Graph::run(const std::string& which, void* const ext[X_COUNT], cudaStream_t stream, std::string* err) {
  if ((flags_ & 2) || profiling_) return run_eager(which, ext, stream, err);
  if (!finalized_ && !finalize(err)) return false;
  if (!warmed_.count(which)) {                       // first use of a plan: eager (one-time kernel attribute setup)
    if (!run_eager(which, ext, stream, err)) return false;
    warmed_.insert(which);
    return true;
  }
  for (auto& c : graph_cache_) {
    if (c.which == which && memcmp(c.ext, ext, sizeof(c.ext)) == 0) {
      c.stamp = ++graph_clock_;
      ++cache_hits_;
      last_launches_ = c.launches;
      cudaError_t ce = cudaGraphLaunch(c.exec, stream);
      if (ce != cudaSuccess) { *err = std::string("cudaGraphLaunch failed: ") + cudaGetErrorString(ce); return false; }
      return true;
    }
  }
  if (!capture_stream_ && cudaStreamCreateWithFlags(&capture_stream_, cudaStreamNonBlocking) != cudaSuccess) {
    *err = "cudaStreamCreate failed";
    return false;
  }
  ++cache_misses_;                                  // a new pointer set: one stream capture + instantiate (~100x a replay)
  cudaError_t ce = cudaStreamBeginCapture(capture_stream_, cudaStreamCaptureModeThreadLocal);
  if (ce != cudaSuccess) { *err = std::string("cudaStreamBeginCapture failed: ") + cudaGetErrorString(ce); return false; }
  const bool ok = run_eager(which, ext, capture_stream_, err);
  cudaGraph_t graph = nullptr;
  ce = cudaStreamEndCapture(capture_stream_, &graph);
  if (!ok) { if (graph) cudaGraphDestroy(graph); return false; }
  if (ce != cudaSuccess || !graph) { *err = std::string("cudaStreamEndCapture failed: ") + cudaGetErrorString(ce); return false; }
  CachedGraph c;
  c.which = which;
  memcpy(c.ext, ext, sizeof(c.ext));
  c.launches = last_launches_;
  c.stamp = ++graph_clock_;
  ce = cudaGraphInstantiate(&c.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) { *err = std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ce); return false; }
  if (graph_cache_.size() >= 48) {                   // evict the least recently used capture
    size_t lru = 0;
    for (size_t i = 1; i < graph_cache_.size(); ++i)
      if (graph_cache_[i].stamp < graph_cache_[lru].stamp) lru = i;
    cudaGraphExecDestroy(graph_cache_[lru].exec);
    graph_cache_[lru] = c;
  } else {
    graph_cache_.push_back(c);
  }
  ce = cudaGraphLaunch(c.exec, stream);
  if (ce != cudaSuccess) { *err = std::string("cudaGraphLaunch failed: ") + cudaGetErrorString(ce); return false; }
  return true;
}

bool Graph::branches_enabled() const {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ACCEL_BRANCHES");
    v = (e && *e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

bool Graph::run_eager(const std::string& which, void* const ext[X_COUNT], cudaStream_t main_stream, std::string* err) {
  auto it = seqs_.find(which);
  if (it == seqs_.end()) { *err = "no such graph: " + which; return false; }
  if (!finalized_ && !finalize(err)) return false;
  std::vector<Op>& ops = it->second;
  int launches = 0;
  size_t ev = 0;
  if (profiling_) {
    event_stage_.clear();
    event_op_.clear();
    while (events_.size() < ops.size() + 1) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      events_.push_back(e);
    }
    cudaEventRecord(events_[ev++], main_stream);
  }
  cudaError_t ce = cudaSuccess;
  // fork / join of parallel sections (see Op::par_group).  Works the same on a real stream and under capture,
  // where it turns the branches into parallel kernel nodes of the graph.
  const bool use_branches = branches_enabled() && !profiling_;
  int cur_group = 0;
  unsigned used = 0;                                       // bit b: side stream b received work in this section
  auto join_all = [&]() -> cudaError_t {
    for (int b = 0; b < kMaxBranches; ++b)
      if (used & (1u << b)) {
        cudaError_t e = cudaEventRecord(join_ev_[b], side_[b]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(main_stream, join_ev_[b], 0);
        if (e != cudaSuccess) return e;
      }
    used = 0;
    return cudaSuccess;
  };
  bool lane_open = false;
  if (use_branches) {
    for (auto& op : ops) lane_open = lane_open || op.lane == 1;
    if (lane_open) {
      if (!lane_stream_) {
        cudaStreamCreateWithFlags(&lane_stream_, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&lane_fork_ev_, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&lane_join_ev_, cudaEventDisableTiming);
      }
      ce = cudaEventRecord(lane_fork_ev_, main_stream);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(lane_stream_, lane_fork_ev_, 0);
      if (ce != cudaSuccess) { *err = std::string("lane fork failed: ") + cudaGetErrorString(ce); return false; }
    }
  }
  auto join_lane = [&]() -> cudaError_t {
    if (!lane_open) return cudaSuccess;
    lane_open = false;
    cudaError_t e = cudaEventRecord(lane_join_ev_, lane_stream_);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(main_stream, lane_join_ev_, 0);
    return e;
  };
  const Op* skip_split = nullptr;
  for (auto& op : ops) {
    cudaStream_t stream = main_stream;
    if (use_branches && op.lane == 1 && lane_open) {
      stream = lane_stream_;
    } else if (use_branches) {
      if (op.join_lane && (ce = join_lane()) != cudaSuccess) { *err = std::string("lane join failed: ") + cudaGetErrorString(ce); return false; }
      if (op.par_group != cur_group) {
        if (cur_group && (ce = join_all()) != cudaSuccess) { *err = std::string("join failed: ") + cudaGetErrorString(ce); return false; }
        cur_group = op.par_group;
        if (cur_group) {
          if (!fork_ev_) cudaEventCreateWithFlags(&fork_ev_, cudaEventDisableTiming);
          ce = cudaEventRecord(fork_ev_, main_stream);
          if (ce != cudaSuccess) { *err = std::string("fork failed: ") + cudaGetErrorString(ce); return false; }
        }
      }
      if (cur_group && op.par_branch > 0 && op.par_branch <= kMaxBranches) {
        const int b = op.par_branch - 1;
        if (!side_[b]) {
          cudaStreamCreateWithFlags(&side_[b], cudaStreamNonBlocking);
          cudaEventCreateWithFlags(&join_ev_[b], cudaEventDisableTiming);
        }
        if (!(used & (1u << b))) {
          ce = cudaStreamWaitEvent(side_[b], fork_ev_, 0);
          if (ce != cudaSuccess) { *err = std::string("fork failed: ") + cudaGetErrorString(ce); return false; }
          used |= 1u << b;
        }
        stream = side_[b];
      }
    }
    switch (op.type) {
      case OP_STEM: {
        StemParams S = op.stem;
        S.src0 = (const float*)ext[op.ext_in0];
        S.src1 = op.ext_in1 != X_NONE ? (const float*)ext[op.ext_in1] : nullptr;
        if (!S.src0 || (op.ext_in1 != X_NONE && !S.src1)) { *err = "missing input frame"; return false; }
        ce = op.stem_tc ? launch_stem_tc(op.stem_tc, S.src0, S.src1, stream) : launch_stem(S, stream);
        ++launches;
        break;
      }
      case OP_CONV: {
        float* ext_nchw = op.epi.ext_out != X_NONE ? (float*)ext[op.epi.ext_out] : nullptr;
        float* ext_raw = op.epi.ext_raw != X_NONE ? (float*)ext[op.epi.ext_raw] : nullptr;
        const int nf = op.frame_convs.empty() ? 1 : (int)op.frame_convs.size();     // > 1: batched tensor, frame by frame
        for (int b = 0; b < nf && ce == cudaSuccess; ++b) {
          TcPlan* plan = op.frame_convs.empty() ? op.tc : op.frame_tc[b];
          if (op.engine == ENG_TC) {
            ce = launch_conv_tc_ext(plan, ext_nchw, ext_raw, stream);
            launches += tc_plan_launches(plan);
          } else {
            ConvParams P = op.frame_convs.empty() ? op.conv : op.frame_convs[b];
            if (ext_nchw) P.epi.out_nchw = ext_nchw;
            P.epi.raw_nchw = ext_raw;
            if (op.engine == ENG_NARROW) {
              ce = launch_conv_narrow(P, stream);
              ++launches;
            } else {
              ce = launch_conv_ffma(P, stream);
              launches += P.splits > 1 ? 2 : 1;
            }
          }
        }
        break;
      }
      case OP_POOL: {
        const Tensor& ti = tensors_[op.in];
        const Tensor& to = tensors_[op.out];
        for (int b = 0; b < ti.nb && ce == cudaSuccess; ++b) {       // frame by frame (window clipping is per frame)
          PoolParams Q = op.pool;
          Q.in_hi += (size_t)b * frame_elems(ti); Q.in_lo += (size_t)b * frame_elems(ti);
          Q.out_hi += (size_t)b * frame_elems(to); Q.out_lo += (size_t)b * frame_elems(to);
          ce = launch_pool(Q, stream);
          ++launches;
        }
        break;
      }
      case OP_DCN_COL: {
        const Tensor& ti = tensors_[op.in];
        const Tensor& tf = tensors_[op.in2];
        const Tensor& to = tensors_[op.out];
        for (int b = 0; b < ti.nb && ce == cudaSuccess; ++b) {       // frame by frame (sampling bounds are per frame)
          DcnColParams D = op.dcn;
          D.in_hi += (size_t)b * frame_elems(ti); D.in_lo += (size_t)b * frame_elems(ti);
          D.offset += (size_t)b * frame_elems(tf);
          D.col_hi += (size_t)b * frame_elems(to); D.col_lo += (size_t)b * frame_elems(to);
          ce = launch_dcn_col(D, stream);
          ++launches;
        }
        break;
      }
      case OP_WARP: {
        WarpParams Wp = op.warp;
        if (op.src_f32 < 0) {
          Wp.feat = (const float*)ext[op.ext_in0];
          Wp.out_nchw = op.ext_out != X_NONE ? (float*)ext[op.ext_out] : nullptr;
          if (!Wp.feat) { *err = "missing feat_key"; return false; }
          if (Wp.out_nchw == Wp.feat) { *err = "feat_out must not alias feat_key"; return false; }
        }
        // one pass for both copies when the fused kernel takes this shape; the conversion op that follows is then skipped
        const bool with_split = op.fuse_split && warp_fused_supported(Wp);
        if (with_split) {
          skip_split = &op + 1;
        } else {
          Wp.out_hi = Wp.out_lo = nullptr;
          Wp.bias = nullptr;
          if (!Wp.out_nchw) Wp.out_nchw = warp_scratch_;        // caller does not keep `warping_feat_output`
        }
        ce = launch_warp(Wp, stream);
        ++launches;
        break;
      }
      case OP_UPFLOW: {
        const Tensor& tf = tensors_[op.in];
        const Tensor& to = tensors_[op.out];
        for (int b = 0; b < tf.nb && ce == cudaSuccess; ++b) {
          UpflowParams U = op.upflow;
          U.flow += (size_t)b * frame_elems(tf);
          U.out_hi += (size_t)b * frame_elems(to); U.out_lo += (size_t)b * frame_elems(to);
          ce = launch_upflow(U, stream);
          ++launches;
        }
        break;
      }
      case OP_FUSE:
        if (op.skip) break;                                  // evaluated inside the tail kernel that follows
        ce = launch_fuse_lowres(op.fuse, stream);
        ++launches;
        break;
      case OP_TAIL: {
        TailParams T = op.tail;
        T.label = op.ext_out != X_NONE && ext[op.ext_out] ? (uint8_t*)ext[op.ext_out] : label_scratch_ptr_;
        T.score_out = op.ext_out2 != X_NONE ? (float*)ext[op.ext_out2] : nullptr;
        ce = launch_tail(T, stream);
        ++launches;
        break;
      }
      case OP_TO_SPLIT: {
        if (&op == skip_split) break;                        // written by the fused warp kernel
        const Tensor& to = tensors_[op.out];
        const float* src = op.src_f32 >= 0 ? f_ptr(tensors_[op.src_f32]) : (const float*)ext[op.ext_in0];
        if (!src && op.src_warp) src = warp_scratch_;
        if (!src) { *err = "missing input tensor"; return false; }
        ce = launch_nchw_to_split(src, to.C, to.H, to.W, hi_ptr(to), lo_ptr(to), to.ld,
                                  stream, op.split_bias_dev, op.split_act);
        ++launches;
        break;
      }
      case OP_COPY_F32: {
        const Tensor& to = tensors_[op.out];
        const void* src = ext[op.ext_in0];
        if (!src) { *err = "missing input tensor"; return false; }
        ce = cudaMemcpyAsync(f_ptr(to), src, (size_t)to.C * to.H * to.W * sizeof(float), cudaMemcpyDeviceToDevice,
                             stream);
        break;
      }
      case OP_TO_NCHW: {
        const Tensor& ti = tensors_[op.in];
        float* dst = (float*)ext[op.ext_out];
        if (dst) {
          ce = launch_split_to_nchw(hi_ptr(ti), lo_ptr(ti), ti.ld, ti.C, ti.H, ti.W, dst,
                                    stream);
          ++launches;
        }
        break;
      }
    }
    if (ce != cudaSuccess) {
      *err = "launch of '" + op.name + "' failed: " + cudaGetErrorString(ce);
      return false;
    }
    if (profiling_) {
      cudaEventRecord(events_[ev++], stream);
      event_stage_.push_back(op.stage);
      event_op_.push_back(&op);
    }
    static const bool debug_sums = [] { const char* e = getenv("ACCEL_DEBUG_SUMS"); return e && e[0] == '1'; }();
    if (debug_sums && op.out >= 0 && !tensors_[op.out].f32) {
      // debugging aid: checksum of every op's split output (hi plane, real channels), printed in launch order
      cudaStreamSynchronize(stream);
      const Tensor& to = tensors_[op.out];
      std::vector<__half> hbuf((size_t)to.H * to.W * to.ld);
      cudaMemcpy(hbuf.data(), hi_ptr(to) - to.coff, hbuf.size() * sizeof(__half), cudaMemcpyDeviceToHost);
      double sum = 0.0, asum = 0.0, amax = 0.0;
      for (size_t px = 0; px < (size_t)to.H * to.W; ++px)
        for (int c = 0; c < to.C; ++c) {
          const double v = (double)__half2float(hbuf[px * to.ld + to.coff + c]);
          sum += v; asum += fabs(v); amax = std::max(amax, fabs(v));
        }
      // |hi| == 65504: the split format clamped an activation (cvt.rn.satfinite) -- a checkpoint whose activations leave
      // the fp16 range would silently lose them; this is where to look for it
      fprintf(stderr, "ACCEL_SUM %s/%s C=%d coff=%d ld=%d sum=%.6e abs=%.6e max=%.6e%s\n", op.stage.c_str(), op.name.c_str(), to.C,
              to.coff, to.ld, sum, asum, amax, amax >= 65504.0 ? "  SATURATED (fp16 range)" : "");
    }
  }
  if (cur_group && (ce = join_all()) != cudaSuccess) { *err = std::string("join failed: ") + cudaGetErrorString(ce); return false; }
  if ((ce = join_lane()) != cudaSuccess) { *err = std::string("lane join failed: ") + cudaGetErrorString(ce); return false; }
  last_launches_ = launches;
  return true;
}

bool Graph::fetch_op_output(const std::string& which, const std::string& op_name, float* dst, int64_t shape[4],
                            cudaStream_t stream, std::string* err) {
  auto it = seqs_.find(which);
  if (it == seqs_.end()) { *err = "no such graph: " + which; return false; }
  if (!finalized_) { *err = "not finalized"; return false; }
  const Op* found = nullptr;
  for (const Op& op : it->second)
    if (op.name == op_name && op.out >= 0 && !tensors_[op.out].f32) found = &op;      // the last op of that name
  if (!found) { *err = "no op '" + op_name + "' with a split output in plan '" + which + "'"; return false; }
  const Tensor& t = tensors_[found->out];
  shape[0] = t.nb; shape[1] = t.C; shape[2] = t.H; shape[3] = t.W;
  if (!dst) return true;
  for (int b = 0; b < t.nb; ++b) {
    cudaError_t ce = launch_split_to_nchw(hi_ptr(t) + (size_t)b * frame_elems(t), lo_ptr(t) + (size_t)b * frame_elems(t), t.ld, t.C,
                                          t.H, t.W, dst + (size_t)b * t.C * t.H * t.W, stream);
    if (ce != cudaSuccess) { *err = std::string("split_to_nchw: ") + cudaGetErrorString(ce); return false; }
  }
  return true;
}

const std::vector<std::pair<std::string, float>>& Graph::stage_times() {
  times_.clear();
  if (event_stage_.empty()) return times_;
  cudaEventSynchronize(events_[event_stage_.size()]);
  for (size_t i = 0; i < event_stage_.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, events_[i], events_[i + 1]);
    bool found = false;
    for (auto& kv : times_)
      if (kv.first == event_stage_[i]) { kv.second += ms; found = true; break; }
    if (!found) times_.push_back({event_stage_[i], ms});
  }
  return times_;
}

std::vector<OpTime> Graph::op_times() {
  std::vector<OpTime> out;
  if (event_stage_.empty()) return out;
  cudaEventSynchronize(events_[event_stage_.size()]);
  for (size_t i = 0; i < event_op_.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, events_[i], events_[i + 1]);
    const Op* op = event_op_[i];
    std::string nm = op->stage + "/" + op->name;
    if (op->type == OP_CONV && op->kind == 1) nm += "[" + std::to_string(op->phase_y) + std::to_string(op->phase_x) + "]";
    out.push_back({nm, ms, op->flops});
  }
  return out;
}

}  // namespace accel
