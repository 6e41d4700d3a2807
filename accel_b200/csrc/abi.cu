// extern "C" surface declared in include/accel_b200.h.  No C++ exception leaves this file.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>

#include "../../include/accel_b200.h"
#include "graph.h"

using namespace accel;

struct AccelHandle {
  std::vector<OpTime> ops;
  AccelConfig cfg;
  int interval = 0;            // frames of the whole-interval plan (accel_plan_interval), 0 = not planned
  Graph* graph;
  std::string err;
  const char* stage_names[32];
  std::vector<std::pair<std::string, float>> stages;
};

static std::string g_create_err;
static std::mutex g_mu;

#define ACCEL_TRY try {
#define ACCEL_CATCH(h)                                                        \
  }                                                                           \
  catch (const std::bad_alloc&) { if (h) (h)->err = "out of host memory"; return 2; } \
  catch (const std::exception& e) { if (h) (h)->err = e.what(); return 3; }  \
  catch (...) { if (h) (h)->err = "unknown C++ exception"; return 4; }

extern "C" int accel_create(const AccelConfig* config, AccelHandle** out) {
  AccelHandle* h = nullptr;
  try {
    if (!config || !out) { std::lock_guard<std::mutex> l(g_mu); g_create_err = "null argument"; return 1; }
    h = new AccelHandle();
    h->cfg = *config;
    h->graph = new Graph(config->device, config->flags);
    std::string err;
    if (!build_accel(*h->graph, config->version, config->height, config->width, config->num_classes, &err)) {
      std::lock_guard<std::mutex> l(g_mu);
      g_create_err = err;
      delete h->graph;
      delete h;
      return 1;
    }
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    std::lock_guard<std::mutex> l(g_mu);
    g_create_err = e.what();
  } catch (...) {
    std::lock_guard<std::mutex> l(g_mu);
    g_create_err = "unknown C++ exception";
  }
  if (h) { delete h->graph; delete h; }
  return 4;
}

extern "C" void accel_destroy(AccelHandle* h) {
  if (!h) return;
  try {
    delete h->graph;
    delete h;
  } catch (...) {
  }
}

extern "C" const char* accel_last_error(const AccelHandle* h) {
  if (h) return h->err.c_str();
  std::lock_guard<std::mutex> l(g_mu);
  static thread_local std::string copy;
  copy = g_create_err;
  return copy.c_str();
}

extern "C" int accel_param_count(const AccelHandle* h) { return h ? (int)h->graph->params().size() : -1; }

extern "C" int accel_param_info(const AccelHandle* h, int index, const char** name, int64_t shape[4], int* ndim) {
  if (!h || index < 0 || index >= (int)h->graph->params().size()) return 1;
  const ParamSpec& p = h->graph->params()[index];
  if (name) *name = p.name.c_str();
  if (ndim) *ndim = (int)p.shape.size();
  if (shape)
    for (size_t i = 0; i < 4; ++i) shape[i] = i < p.shape.size() ? p.shape[i] : 1;
  return 0;
}

extern "C" int accel_set_param(AccelHandle* h, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!h) return 1;
  ACCEL_TRY
  if (!name || !data || !shape) { h->err = "null argument"; return 1; }
  return h->graph->set_param(name, data, shape, ndim, &h->err) ? 0 : 1;
  ACCEL_CATCH(h)
}

extern "C" int accel_finalize(AccelHandle* h) {
  if (!h) return 1;
  ACCEL_TRY
  return h->graph->finalize(&h->err) ? 0 : 1;
  ACCEL_CATCH(h)
}

static int check_stream_error(AccelHandle* h) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    h->err = std::string("CUDA error: ") + cudaGetErrorString(e);
    return 5;
  }
  return 0;
}

extern "C" int accel_key_forward(AccelHandle* h, const float* data, float* feat_out, float* score_out,
                                 uint8_t* label_out, void* stream) {
  if (!h) return 1;
  ACCEL_TRY
  if (!data) { h->err = "accel_key_forward: data is NULL"; return 1; }
  void* ext[X_COUNT] = {nullptr};
  ext[X_DATA] = (void*)data;
  ext[X_FEAT_OUT] = feat_out;
  ext[X_SCORE_OUT] = score_out;
  ext[X_LABEL_OUT] = label_out;
  if (!h->graph->run("key", ext, (cudaStream_t)stream, &h->err)) return 1;
  return check_stream_error(h);
  ACCEL_CATCH(h)
}

extern "C" int accel_cur_forward(AccelHandle* h, const float* data, const float* data_key, const float* feat_key,
                                 float* feat_out, float* score_out, uint8_t* label_out, void* stream) {
  if (!h) return 1;
  ACCEL_TRY
  if (!data || !data_key || !feat_key) { h->err = "accel_cur_forward: data, data_key and feat_key are required"; return 1; }
  void* ext[X_COUNT] = {nullptr};
  ext[X_DATA] = (void*)data;
  ext[X_DATA_KEY] = (void*)data_key;
  ext[X_FEAT_KEY] = (void*)feat_key;
  ext[X_FEAT_OUT] = feat_out;
  ext[X_SCORE_OUT] = score_out;
  ext[X_LABEL_OUT] = label_out;
  if (!h->graph->run("cur", ext, (cudaStream_t)stream, &h->err)) return 1;
  return check_stream_error(h);
  ACCEL_CATCH(h)
}

extern "C" int accel_key_forward_lin(AccelHandle* h, const float* data, float* feat_out, float* g_out, float* score_out,
                                     uint8_t* label_out, void* stream) {
  if (!h) return 1;
  ACCEL_TRY
  if (!data) { h->err = "accel_key_forward_lin: data is NULL"; return 1; }
  void* ext[X_COUNT] = {nullptr};
  ext[X_DATA] = (void*)data;
  ext[X_FEAT_OUT] = feat_out;
  ext[X_G_OUT] = g_out;
  ext[X_SCORE_OUT] = score_out;
  ext[X_LABEL_OUT] = label_out;
  if (!h->graph->run("key", ext, (cudaStream_t)stream, &h->err)) return 1;
  return check_stream_error(h);
  ACCEL_CATCH(h)
}

extern "C" int accel_cur_forward_lin(AccelHandle* h, const float* data, const float* data_key, const float* g_key, float* g_out,
                                     float* score_out, uint8_t* label_out, void* stream) {
  if (!h) return 1;
  ACCEL_TRY
  if (!data || !data_key || !g_key) { h->err = "accel_cur_forward_lin: data, data_key and g_key are required"; return 1; }
  if (h->cfg.version == ACCEL_VERSION_101) { h->err = "accel_cur_forward_lin: Accel-101 fuses at feature level and needs the warped feature itself"; return 1; }
  void* ext[X_COUNT] = {nullptr};
  ext[X_DATA] = (void*)data;
  ext[X_DATA_KEY] = (void*)data_key;
  ext[X_G_KEY] = (void*)g_key;
  ext[X_G_OUT] = g_out;
  ext[X_SCORE_OUT] = score_out;
  ext[X_LABEL_OUT] = label_out;
  if (!h->graph->run("cur_lin", ext, (cudaStream_t)stream, &h->err)) return 1;
  return check_stream_error(h);
  ACCEL_CATCH(h)
}

extern "C" int accel_rbranch_forward(AccelHandle* h, const float* data, float* score_out, uint8_t* label_out, void* stream) {
  if (!h) return 1;
  ACCEL_TRY
  if (!data) { h->err = "accel_rbranch_forward: data is NULL"; return 1; }
  if (h->cfg.version != ACCEL_VERSION_18 && h->cfg.version != ACCEL_VERSION_34 && h->cfg.version != ACCEL_VERSION_50) {
    h->err = "accel_rbranch_forward: only Accel-18/34/50 carry a correction network with its own head";
    return 1;
  }
  void* ext[X_COUNT] = {nullptr};
  ext[X_DATA] = (void*)data;
  ext[X_SCORE_OUT] = score_out;
  ext[X_LABEL_OUT] = label_out;
  if (!h->graph->run("rbranch", ext, (cudaStream_t)stream, &h->err)) return 1;
  return check_stream_error(h);
  ACCEL_CATCH(h)
}

extern "C" int accel_plan_interval(AccelHandle* h, int interval) {
  if (!h) return 1;
  ACCEL_TRY
  if (!build_interval(*h->graph, h->cfg.version, h->cfg.height, h->cfg.width, h->cfg.num_classes, interval, &h->err)) return 1;
  h->interval = interval;
  return 0;
  ACCEL_CATCH(h)
}

extern "C" int accel_interval_forward(AccelHandle* h, const float* const* frames, float* const* score_out,
                                      uint8_t* const* label_out, void* stream) {
  if (!h) return 1;
  ACCEL_TRY
  if (h->interval < 2) { h->err = "accel_interval_forward: call accel_plan_interval before accel_finalize"; return 1; }
  if (!frames || !label_out) { h->err = "accel_interval_forward: frames and label_out are required"; return 1; }
  void* ext[X_COUNT] = {nullptr};
  for (int t = 0; t < h->interval; ++t) {
    if (!frames[t] || !label_out[t]) { h->err = "accel_interval_forward: NULL frame or label pointer"; return 1; }
    ext[X_FRAME0 + t] = (void*)frames[t];
    ext[X_LABEL0 + t] = label_out[t];
    ext[X_SCORE0 + t] = score_out ? score_out[t] : nullptr;
  }
  if (!h->graph->run("interval", ext, (cudaStream_t)stream, &h->err)) return 1;
  return check_stream_error(h);
  ACCEL_CATCH(h)
}

extern "C" int accel_debug_fetch(AccelHandle* h, const char* plan, const char* op_name, float* out, int64_t shape[4], void* stream) {
  if (!h) return 1;
  ACCEL_TRY
  if (!plan || !op_name || !shape) { h->err = "null argument"; return 1; }
  return h->graph->fetch_op_output(plan, op_name, out, shape, (cudaStream_t)stream, &h->err) ? 0 : 1;
  ACCEL_CATCH(h)
}

extern "C" int accel_graph_cache_stats(const AccelHandle* h, uint64_t* hits, uint64_t* misses) {
  if (!h || !hits || !misses) return 1;
  unsigned long long a = 0, b = 0;
  h->graph->cache_stats(&a, &b);
  *hits = a;
  *misses = b;
  return 0;
}

extern "C" int accel_flownet(AccelHandle* h, const float* data, const float* data_key, float* flow_out, void* stream) {
  if (!h) return 1;
  ACCEL_TRY
  if (!data || !data_key || !flow_out) { h->err = "accel_flownet: null argument"; return 1; }
  void* ext[X_COUNT] = {nullptr};
  ext[X_DATA] = (void*)data;
  ext[X_DATA_KEY] = (void*)data_key;
  ext[X_FLOW_OUT] = flow_out;
  if (!h->graph->run("flow", ext, (cudaStream_t)stream, &h->err)) return 1;
  return check_stream_error(h);
  ACCEL_CATCH(h)
}

extern "C" int accel_last_launch_count(const AccelHandle* h) { return h ? h->graph->last_launches() : -1; }

extern "C" int accel_set_profiling(AccelHandle* h, int enabled) {
  if (!h) return 1;
  h->graph->set_profiling(enabled != 0);
  return 0;
}

extern "C" int accel_stage_times(AccelHandle* h, const char** names, float* ms, int cap) {
  if (!h) return -1;
  try {
    h->stages = h->graph->stage_times();
    int n = 0;
    for (auto& kv : h->stages) {
      if (n >= cap) break;
      names[n] = kv.first.c_str();
      ms[n] = kv.second;
      ++n;
    }
    return n;
  } catch (...) {
    return -1;
  }
}

extern "C" int accel_op_times(AccelHandle* h, const char** names, float* ms, double* flops, int cap) {
  if (!h) return -1;
  try {
    h->ops = h->graph->op_times();
    int n = 0;
    for (auto& o : h->ops) {
      if (n >= cap) break;
      names[n] = o.name.c_str();
      ms[n] = o.ms;
      if (flops) flops[n] = o.flops;
      ++n;
    }
    return n;
  } catch (...) {
    return -1;
  }
}

// ---- operator-level entry points -------------------------------------------------------------------

static int no_device(char* err, int errlen) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    if (err && errlen > 0) snprintf(err, errlen, "no CUDA device: accel_b200 has no CPU fallback");
    cudaGetLastError();
    return 1;
  }
  return 0;
}

extern "C" int accel_warp(const float* feat, const float* flow, float* out, int channels, int height, int width,
                          void* stream) {
  if (!feat || !flow || !out || channels <= 0 || height <= 0 || width <= 0 || feat == out) return 1;
  if (no_device(nullptr, 0)) return 6;
  WarpParams P{};
  P.feat = feat; P.flow = flow; P.C = channels; P.H = height; P.W = width; P.out_nchw = out;
  return launch_warp(P, (cudaStream_t)stream) == cudaSuccess ? 0 : 5;
}

extern "C" int accel_warp_split(const float* feat, const float* flow, float* out, void* out_hi, void* out_lo, int channels,
                                int height, int width, void* stream) {
  if (!feat || !flow || !out || !out_hi || !out_lo || channels <= 0 || height <= 0 || width <= 0 || feat == out) return 1;
  if ((channels % 32) || ((uintptr_t)out_hi & 15) || ((uintptr_t)out_lo & 15)) return 1;
  if (no_device(nullptr, 0)) return 6;
  WarpParams P{};
  P.feat = feat; P.flow = flow; P.C = channels; P.H = height; P.W = width; P.out_nchw = out;
  P.out_hi = (__half*)out_hi; P.out_lo = (__half*)out_lo; P.out_ld = channels;
  return launch_warp(P, (cudaStream_t)stream) == cudaSuccess ? 0 : 5;
}

extern "C" int accel_fuse_argmax(const float* score_a, const float* score_b, const float* corr_weight,
                                 const float* corr_bias, int num_classes, int h, int w, uint8_t* label,
                                 float* score_full, void* stream) {
  if (!score_a || !label || num_classes < 1 || num_classes > 32 || h <= 0 || w <= 0) return 1;
  if (no_device(nullptr, 0)) return 6;
  cudaStream_t st = (cudaStream_t)stream;
  const float* src = score_a;
  float* fused = nullptr;
  if (corr_weight) {
    if (!score_b) return 1;
    if (cudaMallocAsync((void**)&fused, (size_t)num_classes * h * w * sizeof(float), st) != cudaSuccess) return 5;
    FuseParams F{};
    F.a = score_a; F.b = score_b; F.w = corr_weight; F.out = fused; F.K = num_classes; F.h = h; F.w_ = w;
    if (launch_fuse_lowres(F, st) != cudaSuccess) return 5;
    src = fused;
  }
  TailParams T{};
  T.score = src; T.bias = corr_bias; T.K = num_classes; T.h = h; T.w = w; T.factor = 16;
  T.label = label; T.score_out = score_full;
  cudaError_t e = launch_tail(T, st);
  if (fused) cudaFreeAsync(fused, st);
  return e == cudaSuccess ? 0 : 5;
}

extern "C" int accel_preprocess(const uint8_t* bgr_hwc, int height, int width, const double pixel_means_bgr[3], float* out,
                                void* stream) {
  if (!bgr_hwc || !out || !pixel_means_bgr || height <= 0 || width <= 0) return 1;
  if (no_device(nullptr, 0)) return 6;
  return launch_preprocess(bgr_hwc, height, width, pixel_means_bgr, out, (cudaStream_t)stream) == cudaSuccess ? 0 : 5;
}

extern "C" int accel_resize_size(int src_height, int src_width, double fx, double fy, int* dst_height, int* dst_width) {
  if (!dst_height || !dst_width || src_height < 1 || src_width < 1 || !(fx > 0.0) || !(fy > 0.0)) return 1;
  resize_linear_size(src_height, src_width, fx, fy, dst_height, dst_width);
  return 0;
}

extern "C" int accel_resize_bgr(const uint8_t* src_hwc, int src_height, int src_width, double fx, double fy, uint8_t* dst_hwc,
                                void* stream) {
  if (!src_hwc || !dst_hwc || src_height < 1 || src_width < 1 || !(fx > 0.0) || !(fy > 0.0)) return 1;
  if (no_device(nullptr, 0)) return 6;
  return launch_resize_linear(src_hwc, src_height, src_width, fx, fy, dst_hwc, (cudaStream_t)stream) == cudaSuccess ? 0 : 5;
}

extern "C" int accel_confusion(const uint8_t* pred, const uint8_t* label, size_t count, int num_classes, int64_t* hist,
                               void* stream) {
  if (!pred || !label || !hist || num_classes < 1 || num_classes > 32) return 1;
  if (no_device(nullptr, 0)) return 6;
  return launch_confusion(pred, label, count, num_classes, reinterpret_cast<unsigned long long*>(hist),
                          (cudaStream_t)stream) == cudaSuccess ? 0 : 5;
}

extern "C" int accel_conv_layer(int kind, const float* in, int cin, int hin, int win, const float* weight, int cout,
                                int ksize, int stride, int pad, int dilate, int deform_groups, const float* offset,
                                const float* scale, const float* shift, int act, const float* residual, int engine,
                                float* out, int device, char* err, int errlen) {
  auto fail = [&](const std::string& m, int code) {
    if (err && errlen > 0) snprintf(err, errlen, "%s", m.c_str());
    return code;
  };
  try {
    if (!in || !weight || !out) return fail("null argument", 1);
    if (no_device(err, errlen)) return 6;
    Graph g(device, engine == 1 ? 1 : 0);
    std::vector<Op>& s = g.seq("layer");
    if (kind == 3 || kind == 4) {
      // 7x7/s2/p3 stem over fp32 NCHW frames: kind 3 = one frame (cin 3); kind 4 = FlowNet's stem: the frame
      // pair (`in`, `offset`), each 2x2 average-pooled and divided by 255 first (cin 6)
      if (ksize != 7 || stride != 2 || pad != 3 || cout != 64 || cin != (kind == 3 ? 3 : 6) || (kind == 4 && !offset))
        return fail("stem test hook: 7x7, stride 2, pad 3, 64 output channels, cin 3 (kind 3) or 6 (kind 4)", 1);
      EpiSpec se;
      se.bn = "post";
      se.eps = 0.f;
      se.act = act;
      const int y = g.stem(s, "layer", X_DATA, kind == 4 ? X_DATA_KEY : X_NONE, hin, win, kind == 4,
                           kind == 4 ? 1.0f / 255.0f : 1.f, "", "w", cin, se);
      g.to_nchw(s, y, X_AUX_OUT);
      std::string msg;
      std::vector<float> ones(cout, 1.f), zeros(cout, 0.f);
      const int64_t c1[1] = {cout};
      const int64_t wshape[4] = {cout, cin, 7, 7};
      bool ok = g.set_param("w_weight", weight, wshape, 4, &msg) &&
                g.set_param("post_gamma", scale ? scale : ones.data(), c1, 1, &msg) &&
                g.set_param("post_beta", shift ? shift : zeros.data(), c1, 1, &msg) &&
                g.set_param("post_moving_mean", zeros.data(), c1, 1, &msg) &&
                g.set_param("post_moving_var", ones.data(), c1, 1, &msg);
      if (!ok) return fail(msg, 1);
      void* ext[X_COUNT] = {nullptr};
      ext[X_DATA] = (void*)in;
      ext[X_DATA_KEY] = (void*)offset;
      ext[X_AUX_OUT] = out;
      if (!g.run("layer", ext, 0, &msg)) return fail(msg, 1);
      if (cudaDeviceSynchronize() != cudaSuccess) return fail(std::string("CUDA error: ") + cudaGetErrorString(cudaGetLastError()), 5);
      return 0;
    }
    const int x = g.new_tensor(cin, hin, win);
    g.to_split(s, X_DATA, x);
    EpiSpec e;
    e.bn = "post";          // scale/shift travel as a BatchNorm with mean 0, var 1-eps
    e.eps = 0.f;
    e.act = act;
    int y;
    int ho, wo;
    if (kind == 0) {
      ho = (hin + 2 * pad - (dilate * (ksize - 1) + 1)) / stride + 1;
      wo = (win + 2 * pad - (dilate * (ksize - 1) + 1)) / stride + 1;
    } else if (kind == 1) {
      ho = 2 * hin; wo = 2 * win;
    } else {
      ho = hin; wo = win;
    }
    if (residual) {
      const int r = g.new_tensor(cout, ho, wo);
      g.to_split(s, X_AUX_IN, r);
      e.res = r;
    }
    if (kind == 0) {
      y = g.conv(s, "layer", x, "w", cout, ksize, stride, pad, dilate, e);
    } else if (kind == 1) {
      y = g.deconv4(s, "layer", x, "w", cout, e);
    } else {
      if (!offset || ksize != 3 || stride != 1 || pad != 2 || dilate != 2) return fail("deformable test hook: 3x3, stride 1, pad 2, dilate 2 and an offset tensor are required", 1);
      const int off = g.new_tensor(deform_groups * 18, hin, win, true);
      g.copy_f32(s, X_FEAT_KEY, off);
      y = g.dcn(s, "layer", x, off, "w", cout, deform_groups, e);
    }
    if (engine != 0)
      for (auto& op : s)
        if (op.type == OP_CONV) op.engine = engine;
    g.to_nchw(s, y, X_AUX_OUT);
    std::string msg;
    std::vector<float> ones(cout, 1.f), zeros(cout, 0.f);
    const int64_t c1[1] = {cout};
    int64_t wshape[4] = {cout, cin, ksize, ksize};
    if (kind == 1) { wshape[0] = cin; wshape[1] = cout; wshape[2] = wshape[3] = 4; }
    bool ok = g.set_param("w_weight", weight, wshape, 4, &msg) &&
              g.set_param("post_gamma", scale ? scale : ones.data(), c1, 1, &msg) &&
              g.set_param("post_beta", shift ? shift : zeros.data(), c1, 1, &msg) &&
              g.set_param("post_moving_mean", zeros.data(), c1, 1, &msg) &&
              g.set_param("post_moving_var", ones.data(), c1, 1, &msg);
    if (!ok) return fail(msg, 1);
    void* ext[X_COUNT] = {nullptr};
    ext[X_DATA] = (void*)in;
    ext[X_AUX_IN] = (void*)residual;
    ext[X_AUX_OUT] = out;
    ext[X_FEAT_KEY] = (void*)offset;
    if (!g.run("layer", ext, 0, &msg)) return fail(msg, 1);
    if (const char* reps_s = getenv("ACCEL_LAYER_REPS")) {   // tuning aid: best-of-n device time of the contraction itself
      const int reps = atoi(reps_s);
      std::vector<OpTime> best;
      g.set_profiling(true);
      for (int i = 0; i < reps; ++i) {
        if (!g.run("layer", ext, 0, &msg)) return fail(msg, 1);
        std::vector<OpTime> cur = g.op_times();
        if (best.empty()) best = cur;
        for (size_t j = 0; j < cur.size() && j < best.size(); ++j)
          if (cur[j].ms < best[j].ms) best[j].ms = cur[j].ms;
      }
      g.set_profiling(false);
      double ms = 0.0, fl = 0.0;
      for (auto& o : best)
        if (o.flops > 0) { ms += o.ms; fl += o.flops; }
      if (getenv("ACCEL_TC_TRACE")) {                       // one more run with a clean trace buffer, then the timeline of CTA 0
        tc_trace_reset();
        if (!g.run("layer", ext, 0, &msg)) return fail(msg, 1);
        cudaDeviceSynchronize();
        tc_trace_dump(stderr);
      }
      fprintf(stderr, "ACCEL_LAYER kind=%d cin=%d cout=%d hw=%dx%d k=%d s=%d: %.2f us, %.2f GF, %.1f TF16/s\n", kind, cin, cout, hin,
              win, ksize, stride, ms * 1e3, fl / 1e9, ms > 0 ? 3.0 * fl / (ms * 1e-3) / 1e12 : 0.0);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(std::string("CUDA error: ") + cudaGetErrorString(cudaGetLastError()), 5);
    return 0;
  } catch (const std::exception& e) {
    return fail(e.what(), 3);
  } catch (...) {
    return fail("unknown C++ exception", 4);
  }
}

// Operator-level DeepLab head (SURVEY.md 8b `accel_head`): fc6 1x1 + bias + ReLU -> score 1x1 + bias at feature
// resolution, through the same plan machinery and kernels the graphs use (two tcgen05 convs).
extern "C" int accel_head(const float* feat, int cin, int height, int width, const float* fc6_weight, const float* fc6_bias,
                          int mid, const float* score_weight, const float* score_bias, int num_classes, float* score_lowres,
                          int device, char* err, int errlen) {
  auto fail = [&](const std::string& m, int code) {
    if (err && errlen > 0) snprintf(err, errlen, "%s", m.c_str());
    return code;
  };
  try {
    if (!feat || !fc6_weight || !fc6_bias || !score_weight || !score_bias || !score_lowres) return fail("null argument", 1);
    if (cin < 1 || mid < 1 || num_classes < 1 || height < 1 || width < 1) return fail("bad shape", 1);
    if (no_device(err, errlen)) return 6;
    Graph g(device, 0);
    std::vector<Op>& s = g.seq("head");
    const int x = g.new_tensor(cin, height, width);
    g.to_split(s, X_DATA, x);
    EpiSpec e1;
    e1.bias = "fc6_bias";
    e1.act = ACT_RELU;
    const int y = g.conv(s, "head", x, "fc6", mid, 1, 1, 0, 1, e1);
    EpiSpec e2;
    e2.bias = "score_bias";
    e2.act = ACT_NONE;
    const int z = g.conv(s, "head", y, "score", num_classes, 1, 1, 0, 1, e2);
    g.to_nchw(s, z, X_AUX_OUT);
    std::string msg;
    const int64_t w1[4] = {mid, cin, 1, 1}, b1[1] = {mid}, w2[4] = {num_classes, mid, 1, 1}, b2[1] = {num_classes};
    const bool ok = g.set_param("fc6_weight", fc6_weight, w1, 4, &msg) && g.set_param("fc6_bias", fc6_bias, b1, 1, &msg) &&
                    g.set_param("score_weight", score_weight, w2, 4, &msg) && g.set_param("score_bias", score_bias, b2, 1, &msg);
    if (!ok) return fail(msg, 1);
    void* ext[X_COUNT] = {nullptr};
    ext[X_DATA] = (void*)feat;
    ext[X_AUX_OUT] = score_lowres;
    if (!g.run("head", ext, 0, &msg)) return fail(msg, 1);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(std::string("CUDA error: ") + cudaGetErrorString(cudaGetLastError()), 5);
    return 0;
  } catch (const std::exception& e) {
    return fail(e.what(), 3);
  } catch (...) {
    return fail("unknown C++ exception", 4);
  }
}
