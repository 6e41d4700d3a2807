// The Accel inference graphs expressed over the plan builder: which layer feeds which, under the
// reference's parameter names.  Structure follows
//   dff_deeplab/symbols/resnet_v1_101_flownet_deeplab.py  (get_flownet :1751-1808, residual_unit/resnet
//     :29-130, get_resnet_dcn_{18,34}_conv5 :132-233, get_resnet_dcn_50 :235-574, get_resnet_dcn :576-1300)
//   dff_deeplab/symbols/accel_{18,34,50,101}.py           (get_key_test_symbol, get_cur_test_symbol)
// Concats are zero-copy channel views, BatchNorm/bias/activation/residual live in conv epilogues,
// Deconvolution(k4,s2)+Crop(1,1) is run as four 2x2-tap phase convolutions, and the x16 score
// upsampling + fusion + argmax is one tail kernel.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "graph.h"

namespace accel {

namespace {

using Seq = std::vector<Op>;

EpiSpec bn_epi(const std::string& bn, int act, float eps = 1e-5f) {
  EpiSpec e;
  e.bn = bn;
  e.eps = eps;
  e.act = act;
  return e;
}

EpiSpec bias_epi(const std::string& conv, int act) {
  EpiSpec e;
  e.bias = conv + "_bias";
  e.act = act;
  return e;
}

// ---- FlowNet-S -----------------------------------------------------------------------------------
// ext_cur / ext_ref: external slots of the frame pair; inner_branches = false: the refinement levels stay one chain
// (whole-interval plan: the chain itself is a branch next to the other frames' chains)
// nb > 1: the pairs (ext_cur + b, ext_ref + b), b < nb, go through the net as ONE batch (one stem launch per pair, every
// other layer one launch over all frames where the batched kernel takes it)
int flownet(Graph& g, Seq& s, int H, int W, int ext_flow_out, int ext_cur = X_DATA, int ext_ref = X_DATA_KEY,
            bool inner_branches = true, int nb = 1) {
  const std::string st = "flownet";
  int r1 = nb > 1 ? g.new_tensor(64, H / 4, W / 4, false, nb) : -1;
  for (int b = 0; b < nb; ++b)
    r1 = g.stem(s, st, ext_cur + b, ext_ref + b, H, W, true, 1.0f / 255.0f, "", "flow_conv1", 6,
                bias_epi("flow_conv1", ACT_LEAKY), r1, b);
  const Tensor t1 = g.tensor(r1);                                   // (64, H/4, W/4)
  const int c5 = g.new_tensor(128 + 64 + 2, t1.H / 2, t1.W / 2, false, nb);     // Concat5
  const int c4 = g.new_tensor(256 + 128 + 2, t1.H / 4, t1.W / 4, false, nb);    // Concat4
  const int c3 = g.new_tensor(512 + 256 + 2, t1.H / 8, t1.W / 8, false, nb);    // Concat3
  const int c2 = g.new_tensor(512 + 512 + 2, t1.H / 16, t1.W / 16, false, nb);  // Concat2
  int r2 = g.conv(s, st, r1, "conv2", 128, 5, 2, 2, 1, bias_epi("conv2", ACT_LEAKY), g.new_view(c5, 0, 128));
  int r3 = g.conv(s, st, r2, "conv3", 256, 5, 2, 2, 1, bias_epi("conv3", ACT_LEAKY));
  int r4 = g.conv(s, st, r3, "conv3_1", 256, 3, 1, 1, 1, bias_epi("conv3_1", ACT_LEAKY), g.new_view(c4, 0, 256));
  int r5 = g.conv(s, st, r4, "conv4", 512, 3, 2, 1, 1, bias_epi("conv4", ACT_LEAKY));
  int r6 = g.conv(s, st, r5, "conv4_1", 512, 3, 1, 1, 1, bias_epi("conv4_1", ACT_LEAKY), g.new_view(c3, 0, 512));
  int r7 = g.conv(s, st, r6, "conv5", 512, 3, 2, 1, 1, bias_epi("conv5", ACT_LEAKY));
  int r8 = g.conv(s, st, r7, "conv5_1", 512, 3, 1, 1, 1, bias_epi("conv5_1", ACT_LEAKY), g.new_view(c2, 0, 512));
  int r9 = g.conv(s, st, r8, "conv6", 1024, 3, 2, 1, 1, bias_epi("conv6", ACT_LEAKY));
  int r10 = g.conv(s, st, r9, "conv6_1", 1024, 3, 1, 1, 1, bias_epi("conv6_1", ACT_LEAKY));
  (void)r8; (void)r6; (void)r4; (void)r2;

  auto refine = [&](int feat, int cat, int skip_c, int deconv_c, const char* flow_name, const char* deconv_name,
                    const char* up_name) {
    const Tensor tf = g.tensor(feat);
    const int f = g.new_tensor(2, tf.H, tf.W, true, nb);
    EpiSpec fe = bias_epi(flow_name, ACT_NONE);
    fe.out_f32 = f;
    fe.no_split_out = true;
    // One refinement level = five independent chains over the same input: the 2-channel flow head followed by its
    // x2 upsampling, and the four output phases of the feature deconvolution.  They write disjoint channel ranges
    // of the next concat buffer, so they run as parallel branches (forked streams -> parallel graph nodes), each
    // deconv phase planned for a quarter of the SMs.
    const int grp = inner_branches ? g.new_par_group() : 0;
    const size_t i0 = s.size();
    g.conv(s, st, feat, flow_name, 2, 3, 1, 1, 1, fe);
    const size_t i1 = s.size();
    g.deconv4(s, st, feat, deconv_name, deconv_c, bias_epi(deconv_name, ACT_LEAKY), g.new_view(cat, skip_c, deconv_c));
    const size_t i2 = s.size();
    g.upflow(s, f, up_name, std::string(up_name) + "_bias", g.new_view(cat, skip_c + deconv_c, 2));
    // The flow head is the longest chain of a level (9 taps over the whole concat against 4 taps per deconv phase,
    // plus the upsampling behind it): planned for a quarter of the SMs like the phases it took 89 us at level 2 while
    // the phases next to it took 41 us.  It is planned for half of the SMs instead (ACCEL_FLOWHEAD_WIDTH overrides).
    static const int head_width = [] { const char* e = getenv("ACCEL_FLOWHEAD_WIDTH"); return e && *e ? std::max(1, atoi(e)) : 2; }();
    for (size_t i = i0; inner_branches && i < s.size(); ++i) {
      s[i].par_group = grp;
      s[i].par_width = (i < i1) ? head_width : 4;
      if (i >= i1 && i < i2) s[i].par_branch = 1 + s[i].phase_y * 2 + s[i].phase_x;     // branches 1..4: deconv phases
      else s[i].par_branch = (i < i1) ? 5 : 5;                                          // branch 5: flow head -> upflow
    }
  };
  refine(r10, c2, 512, 512, "Convolution1", "deconv5", "upsample_flow6to5");
  refine(c2, c3, 512, 256, "Convolution2", "deconv4", "upsample_flow5to4");
  refine(c3, c4, 256, 128, "Convolution3", "deconv3", "upsample_flow4to3");
  refine(c4, c5, 128, 64, "Convolution4", "deconv2", "upsample_flow3to2");
  const int p5 = g.pool(s, st, c5, 2, 2, 0, false, true);           // resize_concat5
  const Tensor tp = g.tensor(p5);
  EpiSpec fe = bias_epi("Convolution5", ACT_NONE);
  fe.mul = 2.5f;                                                    // `Convolution5 * 2.5`, :1808
  fe.no_split_out = true;
  int flow = -1;
  if (ext_flow_out != X_NONE) {
    fe.ext_out = ext_flow_out;
  } else {
    flow = g.new_tensor(2, tp.H, tp.W, true, nb);
    fe.out_f32 = flow;
  }
  g.conv(s, st, p5, "Convolution5", 2, 3, 1, 1, 1, fe);
  return flow;
}

// ---- caffe-style bottleneck nets (R101-DCN, R50-DCN) --------------------------------------------------
struct DeformCfg {
  int off_ch, off_pad, off_dil, dg;
};

// ext_data: external slot of the frame; final_f32 >= 0: the last layer also writes that internal fp32 planar tensor
// nb > 1: frames ext_data .. ext_data + nb - 1 as one batch; final_f32_frames: how many leading frames get the fp32 copy
int bottleneck_net(Graph& g, Seq& s, const std::string& st, int H, int W, const std::string& prefix,
                   const std::vector<std::vector<std::string>>& stage_units, DeformCfg d, int ext_feat_out,
                   int final_out_view, int ext_data = X_DATA, int final_f32 = -1, int nb = 1, int final_f32_frames = 0) {
  int x = nb > 1 ? g.new_tensor(64, H / 2, W / 2, false, nb) : -1;
  for (int b = 0; b < nb; ++b)
    x = g.stem(s, st, ext_data + b, X_NONE, H, W, false, 1.f, "", prefix + "conv1", 3, bn_epi(prefix + "bn_conv1", ACT_RELU), x, b);
  x = g.pool(s, st, x, 3, 2, 0, true, true);                        // pool1: 3x3/s2, pooling_convention='full'
  const int mids[4] = {64, 128, 256, 512};
  for (int si = 0; si < 4; ++si) {
    const int stage = si + 2, mid = mids[si];
    const auto& units = stage_units[si];
    for (size_t n = 0; n < units.size(); ++n) {
      const int stride = (n == 0 && (stage == 3 || stage == 4)) ? 2 : 1;
      const std::string r = prefix + "res" + std::to_string(stage) + units[n];
      const std::string b = prefix + "bn" + std::to_string(stage) + units[n];
      const bool last = si == 3 && n + 1 == units.size();
      int sc = x;
      if (n == 0) sc = g.conv(s, st, x, r + "_branch1", 4 * mid, 1, stride, 0, 1, bn_epi(b + "_branch1", ACT_NONE));
      int a = g.conv(s, st, x, r + "_branch2a", mid, 1, stride, 0, 1, bn_epi(b + "_branch2a", ACT_RELU));
      int m;
      if (stage == 5) {
        const Tensor ta = g.tensor(a);
        const int off = g.new_tensor(d.off_ch, ta.H, ta.W, true, nb);
        EpiSpec oe = bias_epi(r + "_branch2b_offset", ACT_NONE);
        oe.out_f32 = off;
        oe.no_split_out = true;
        g.conv(s, st, a, r + "_branch2b_offset", d.off_ch, 3, 1, d.off_pad, d.off_dil, oe);
        m = g.dcn(s, st, a, off, r + "_branch2b", mid, d.dg, bn_epi(b + "_branch2b", ACT_RELU));
      } else {
        m = g.conv(s, st, a, r + "_branch2b", mid, 3, 1, 1, 1, bn_epi(b + "_branch2b", ACT_RELU));
      }
      EpiSpec ce = bn_epi(b + "_branch2c", ACT_RELU);
      ce.res = sc;
      if (last) ce.ext_out = ext_feat_out;
      if (last && final_f32 >= 0) { ce.out_f32 = final_f32; ce.f32_frames = final_f32_frames; }
      x = g.conv(s, st, m, r + "_branch2c", 4 * mid, 1, 1, 0, 1, ce, last ? final_out_view : -1);
    }
  }
  return x;
}

std::vector<std::vector<std::string>> units_101() {
  std::vector<std::string> r4 = {"a"};
  for (int i = 1; i <= 22; ++i) r4.push_back("b" + std::to_string(i));
  return {{"a", "b", "c"}, {"a", "b1", "b2", "b3"}, r4, {"a", "b", "c"}};
}

std::vector<std::vector<std::string>> units_50() {
  return {{"a", "b", "c"}, {"a", "b", "c", "d"}, {"a", "b", "c", "d", "e", "f"}, {"a", "b", "c"}};
}

// ---- pre-activation basic-block trunk + deformable conv5 (Accel-18 / Accel-34 R branch) ----------------
int preact_branch(Graph& g, Seq& s, const std::string& st, int H, int W, const std::string& pre,
                  const std::vector<int>& units, const std::string& letters, bool fold_fc6, int ext_data = X_DATA, int nb = 1) {
  const float eps = 2e-5f;
  int x = nb > 1 ? g.new_tensor(64, H / 2, W / 2, false, nb) : -1;
  for (int b = 0; b < nb; ++b)
    x = g.stem(s, st, ext_data + b, X_NONE, H, W, false, 1.f, pre + "bn_data", pre + "conv0", 3,
               bn_epi(pre + "bn0", ACT_RELU, eps), x, b);
  // max pool, then stage1_unit1's bn1 + relu on the pooled map (unit 1 never reads the raw input:
  // its shortcut conv also takes act1, :80-81)
  int act = g.pool(s, st, x, 3, 2, 1, true, false, bn_epi(pre + "stage1_unit1_bn1", ACT_RELU, eps));
  int raw = -1;
  const int chans[3] = {64, 128, 256};
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < units[i]; ++j) {
      const std::string name = pre + "stage" + std::to_string(i + 1) + "_unit" + std::to_string(j + 1);
      const int c = chans[i];
      const int stride = (j == 0 && i > 0) ? 2 : 1;
      const bool dim_match = j > 0;
      int c1 = g.conv(s, st, act, name + "_conv1", c, 3, stride, 1, 1, bn_epi(name + "_bn2", ACT_RELU, eps));
      int shortcut = dim_match ? raw : g.conv(s, st, act, name + "_sc", c, 1, stride, 0, 1, EpiSpec());
      // what follows this unit?
      const bool last_unit = i == 2 && j + 1 == units[i];
      std::string next;
      bool next_needs_raw = last_unit;
      if (!last_unit) {
        if (j + 1 < units[i]) {
          next = pre + "stage" + std::to_string(i + 1) + "_unit" + std::to_string(j + 2);
          next_needs_raw = true;                                   // identity shortcut reads the raw sum
        } else {
          next = pre + "stage" + std::to_string(i + 2) + "_unit1";
        }
      }
      const Tensor t1 = g.tensor(c1);
      EpiSpec e;
      e.res = shortcut;
      if (!next.empty()) {
        e.bn2 = next + "_bn1";
        e.eps2 = eps;
        e.act2 = ACT_RELU;
        e.out2 = g.new_tensor(c, t1.H, t1.W, false, nb);
      }
      e.no_split_out = !next_needs_raw;
      raw = g.conv(s, st, c1, name + "_conv2", c, 3, 1, 1, 1, e);
      act = e.out2;
    }
  }
  // conv5: post-activation basic units, deformable second conv (3x3, dil 2, dg 4)
  int xx = raw;
  for (size_t n = 0; n < letters.size(); ++n) {
    const std::string L(1, letters[n]);
    const bool first = n == 0;
    int sc = xx;
    if (first) sc = g.conv(s, st, xx, pre + "res5" + L + "_branch1", 512, 1, 2, 0, 1, bn_epi(pre + "bn5" + L + "_branch1", ACT_NONE));
    int a = g.conv(s, st, xx, pre + "res5" + L + "_branch2a", 512, 3, first ? 2 : 1, 1, 1,
                   bn_epi(pre + "bn5" + L + "_branch2a", ACT_RELU));
    const Tensor ta = g.tensor(a);
    const int off = g.new_tensor(72, ta.H, ta.W, true, nb);
    EpiSpec oe = bias_epi(pre + "res5" + L + "_branch2b_offset", ACT_NONE);
    oe.out_f32 = off;
    oe.no_split_out = true;
    g.conv(s, st, a, pre + "res5" + L + "_branch2b_offset", 72, 3, 1, 2, 2, oe);
    EpiSpec be = bn_epi(pre + "bn5" + L + "_branch2b", ACT_RELU);
    be.res = sc;
    xx = g.dcn(s, st, a, off, pre + "res5" + L + "_branch2b", 512, 4, be);
  }
  // `<pre>fc6` directly follows `<pre>feat_upsampling` (accel_18.py:204-213; no bias, BN or
  // activation in between), so the two linear maps run as ONE 512 -> 1024 transposed conv whose weights are composed
  // at finalize (34.4 instead of 103.1 GFLOP at 1024x2048).  ACCEL_FOLD_FC6=0 keeps the graph as written.
  if (fold_fc6) return g.deconv4(s, st, xx, pre + "feat_upsampling", 1024, bias_epi(pre + "fc6", ACT_RELU), -1, pre + "fc6", 2048);
  return g.deconv4(s, st, xx, pre + "feat_upsampling", 2048, EpiSpec());
}

// ---- DeepLab head: fc6 1x1 + ReLU -> score 1x1; returns the low-res fp32 score map ----------------------
// ext_raw: external slot that receives fc6's LINEAR part W*F (no bias, no ReLU) as fp32 NCHW when the caller passes
// a pointer -- the quantity the commuted L head of the cur frames warps instead of the 2048-channel feature.
// fc6_done >= 0: `feat` already IS relu(fc6(.)) (commuted head): only the score conv runs.
int head(Graph& g, Seq& s, const std::string& st, int feat, const std::string& fc6, const std::string& score,
         const std::string& upsampling, int K, int ext_raw = X_NONE, bool fc6_done = false) {
  int x = feat;
  if (!fc6_done) {
    EpiSpec fe = bias_epi(fc6, ACT_RELU);
    fe.ext_raw = ext_raw;
    x = g.conv(s, st, feat, fc6, 1024, 1, 1, 0, 1, fe);
  }
  const Tensor tx = g.tensor(x);
  const int sc = g.new_tensor(K, tx.H, tx.W, true, tx.nb);
  EpiSpec e = bias_epi(score, ACT_NONE);
  e.out_f32 = sc;
  e.no_split_out = true;
  g.conv(s, st, x, score, K, 1, 1, 0, 1, e);
  g.require_bilinear(upsampling + "_weight", K);
  return sc;
}

// The correction ("R") network of Accel-18/34/50 with its own DeepLab head: returns the low-res fp32 score map
// (accel_18.py:199-221, accel_50.py:195-216).  On its own it is the plain DeepLab-<v> segmentation net (BASELINE config 1).
int rbranch_scores(Graph& g, Seq& s, int version, int H, int W, int K, bool fold_fc6, int ext_data = X_DATA, int nb = 1) {
  if (version == 50) {
    int f = bottleneck_net(g, s, "rbranch", H, W, "50_", units_50(), DeformCfg{72, 2, 2, 4}, X_NONE, -1, ext_data, -1, nb);
    return head(g, s, "rhead", f, "curr_fc6", "curr_score", "curr_upsampling", K);
  }
  const std::string pre = std::to_string(version) + "_";
  int f = preact_branch(g, s, "rbranch", H, W, pre, version == 18 ? std::vector<int>{2, 2, 2} : std::vector<int>{3, 4, 6},
                        version == 18 ? "ab" : "abc", fold_fc6, ext_data, nb);
  return head(g, s, "rhead", f, pre + "fc6", pre + "score", pre + "upsampling", K, X_NONE, fold_fc6);
}

}  // namespace

// ---- whole-interval plan ------------------------------------------------------------------------------------
// One key interval of the chained schedule (dff_deeplab/demo.py:228-250) as ONE launch sequence: frame 0 through the
// key graph, frames 1..I-1 through the cur graph with data_key = the previous frame and feat_key = the previous
// frame's (warped) feature.  Everything that depends on a frame alone -- R101 of the key frame, FlowNet of every
// (frame, previous frame) pair, the correction network of every cur frame -- is an independent chain; the chains
// are parallel branches of one section, each planned for a share of the SMs proportional to its flops, so the GPU
// works on the whole interval's tiles at once instead of one small layer at a time (temporal batching; precedent:
// dff_rfcn/demo_batch.py:76-96).  Only the flow-guided warps are sequential (warp_t reads warp_{t-1}); they and the
// heads / fusion / tails that consume them follow the parallel section.
bool build_interval(Graph& g, int version, int H, int W, int K, int interval, std::string* err) {
  if (interval < 2 || interval > kMaxInterval) {
    *err = "accel_plan_interval: interval must be in [2, " + std::to_string(kMaxInterval) + "]";
    return false;
  }
  if (!g.seq("interval").empty()) { *err = "accel_plan_interval: already planned"; return false; }
  if (g.finalized()) { *err = "accel_plan_interval must precede accel_finalize"; return false; }
  const int h = H / 16, w = W / 16, I = interval;
  const char* ff = getenv("ACCEL_FOLD_FC6");
  const bool fold_fc6 = !(ff && ff[0] == '0');
  Seq& s = g.seq("interval");
  const int grp = g.new_par_group();
  struct Chain { size_t begin, end; };
  std::vector<Chain> chains;
  auto close_chain = [&](size_t b) { chains.push_back({b, s.size()}); };

  // Frames that go through the same network with the same weights share ONE launch per layer (ACCEL_IVL_BATCH=0: one chain
  // per frame instead): R101 over all I frames in Accel-101 (the key net and the correction net are the same network,
  // accel_101.py:161), FlowNet over the I-1 frame pairs, the correction network over the I-1 cur frames.
  const char* be = getenv("ACCEL_IVL_BATCH");
  const bool batched = !(be && be[0] == '0') && !(g.flags() & 1);
  std::vector<int> feat(I);                                  // fp32 planar features: F_0 = res5c_relu, F_t = warp_t(F_{t-1})
  for (int t = 0; t < I; ++t) feat[t] = g.new_tensor(2048, h, w, true);
  std::vector<int> flow(I, -1), cat(I, -1), sr(I, -1), warped(I, -1);
  int cat_all = -1, flow_all = -1, sr_all = -1, warped_all = -1;
  if (batched) {
    {
      const size_t b = s.size();
      int f0;
      if (version == 101) {
        cat_all = g.new_tensor(4096, h, w, false, I);          // frame t: [warp_t(F_{t-1}) | R101(frame t)]
        const int upper = g.new_view(cat_all, 2048, 2048);
        bottleneck_net(g, s, "backbone", H, W, "", units_101(), DeformCfg{18, 1, 1, 1}, X_NONE, upper, X_FRAME0, feat[0], I, 1);
        f0 = g.frame_view(upper, 0, 1);
        for (int t = 1; t < I; ++t) cat[t] = g.frame_view(cat_all, t, 1);
      } else {
        f0 = bottleneck_net(g, s, "backbone", H, W, "", units_101(), DeformCfg{18, 1, 1, 1}, X_NONE, -1, X_FRAME0, feat[0]);
      }
      int sc = head(g, s, "head", f0, "fc6", "score", "upsampling", K);
      g.tail(s, sc, "", X_LABEL0, X_SCORE0);
      close_chain(b);
    }
    {
      const size_t b = s.size();
      flow_all = flownet(g, s, H, W, X_NONE, X_FRAME0 + 1, X_FRAME0, false, I - 1);
      for (int t = 1; t < I; ++t) flow[t] = g.frame_view(flow_all, t - 1, 1);
      close_chain(b);
    }
    if (version != 0 && version != 101) {
      const size_t b = s.size();
      sr_all = rbranch_scores(g, s, version, H, W, K, fold_fc6, X_FRAME0 + 1, I - 1);
      for (int t = 1; t < I; ++t) sr[t] = g.frame_view(sr_all, t - 1, 1);
      close_chain(b);
    }
    if (version != 101) {
      warped_all = g.new_tensor(2048, h, w, false, I - 1);
      for (int t = 1; t < I; ++t) warped[t] = g.frame_view(warped_all, t - 1, 1);
    }
  } else {
  // chain 0: key frame -- R101-DCN, its head and its tail (get_key_test_symbol)
  {
    const size_t b = s.size();
    int f0 = bottleneck_net(g, s, "backbone", H, W, "", units_101(), DeformCfg{18, 1, 1, 1}, X_NONE, -1, X_FRAME0, feat[0]);
    int sc = head(g, s, "head", f0, "fc6", "score", "upsampling", K);
    g.tail(s, sc, "", X_LABEL0, X_SCORE0);
    close_chain(b);
  }
  // FlowNet of every frame pair
  for (int t = 1; t < I; ++t) {
    const size_t b = s.size();
    flow[t] = flownet(g, s, H, W, X_NONE, X_FRAME0 + t, X_FRAME0 + t - 1, false);
    close_chain(b);
  }
  // correction network of every cur frame
  for (int t = 1; t < I && version != 0; ++t) {
    const size_t b = s.size();
    if (version == 101) {
      cat[t] = g.new_tensor(4096, h, w);
      bottleneck_net(g, s, "rbranch", H, W, "", units_101(), DeformCfg{18, 1, 1, 1}, X_NONE, g.new_view(cat[t], 2048, 2048),
                     X_FRAME0 + t);
    } else {
      sr[t] = rbranch_scores(g, s, version, H, W, K, fold_fc6, X_FRAME0 + t);
    }
    close_chain(b);
  }
  }
  // SM shares proportional to the chains' flops
  {
    std::vector<double> fl(chains.size(), 0.0);
    double total = 0.0;
    for (size_t c = 0; c < chains.size(); ++c) {
      for (size_t i = chains[c].begin; i < chains[c].end; ++i) fl[c] += s[i].flops;
      total += fl[c];
    }
    const int sms = g.num_sms_hint();
    const char* mb = getenv("ACCEL_IVL_MIN_SMS");
    const int min_sms = mb && *mb ? std::max(1, atoi(mb)) : 8;
    std::vector<int> share(chains.size());
    int used = 0;
    for (size_t c = 0; c < chains.size(); ++c) {
      share[c] = std::max(min_sms, (int)(sms * fl[c] / total));
      used += share[c];
    }
    // hand out / take back the rounding remainder, largest chains first
    for (int guard = 0; used != sms && guard < 4 * sms; ++guard) {
      size_t best = 0;
      for (size_t c = 1; c < chains.size(); ++c)
        if (fl[c] / share[c] > fl[best] / share[best]) best = c;
      if (used < sms) { ++share[best]; ++used; continue; }
      size_t worst = chains.size();
      for (size_t c = 0; c < chains.size(); ++c)
        if (share[c] > min_sms && (worst == chains.size() || fl[c] / share[c] < fl[worst] / share[worst])) worst = c;
      if (worst == chains.size()) break;
      --share[worst]; --used;
    }
    const char* se = getenv("ACCEL_IVL_SERIAL");           // 1: the chains one after the other, each on all SMs
    const bool serial = se && se[0] == '1';
    for (size_t c = 0; c < chains.size() && !serial; ++c)
      for (size_t i = chains[c].begin; i < chains[c].end; ++i) {
        s[i].par_group = grp;
        s[i].par_branch = (int)c + 1;
        s[i].par_width = 1;
        static const bool full_width = [] { const char* e = getenv("ACCEL_IVL_FULLWIDTH"); return e && e[0] == '1'; }();
        s[i].sm_budget = full_width ? 0 : share[c];       // 1: every chain plans for all SMs and the hardware interleaves them
      }
  }
  // sequential part: the chained warps and what consumes them
  if (batched) {
    for (int t = 1; t < I; ++t)
      g.warp_internal(s, feat[t - 1], flow[t], feat[t], version == 101 ? g.new_view(cat[t], 0, 2048) : warped[t]);
    int sl_all;
    if (version == 101) {
      int fused = g.conv(s, "fusion", g.frame_view(cat_all, 1, I - 1), "corr", 2048, 1, 1, 0, 1, bias_epi("corr", ACT_NONE));
      sl_all = head(g, s, "head", fused, "fc6", "score", "upsampling", K);
    } else {
      sl_all = head(g, s, "head", warped_all, "fc6", "score", "upsampling", K);
    }
    for (int t = 1; t < I; ++t) {
      const int sl = g.frame_view(sl_all, t - 1, 1);
      if (version == 0 || version == 101) {
        g.tail(s, sl, "", X_LABEL0 + t, X_SCORE0 + t);
      } else {
        const int fused = g.new_tensor(K, h, w, true);
        g.fuse(s, sl, sr[t], "corr", fused);
        g.tail(s, fused, "corr_bias", X_LABEL0 + t, X_SCORE0 + t);
      }
    }
    return true;
  }
  for (int t = 1; t < I; ++t) {
    if (version == 101) {
      g.warp_internal(s, feat[t - 1], flow[t], feat[t], g.new_view(cat[t], 0, 2048));
      int fused = g.conv(s, "fusion", cat[t], "corr", 2048, 1, 1, 0, 1, bias_epi("corr", ACT_NONE));
      int sc = head(g, s, "head", fused, "fc6", "score", "upsampling", K);
      g.tail(s, sc, "", X_LABEL0 + t, X_SCORE0 + t);
    } else {
      const int wt = g.new_tensor(2048, h, w);
      g.warp_internal(s, feat[t - 1], flow[t], feat[t], wt);
      const int sl = head(g, s, "head", wt, "fc6", "score", "upsampling", K);
      if (version == 0) {
        g.tail(s, sl, "", X_LABEL0 + t, X_SCORE0 + t);
      } else {
        const int fused = g.new_tensor(K, h, w, true);
        g.fuse(s, sl, sr[t], "corr", fused);
        g.tail(s, fused, "corr_bias", X_LABEL0 + t, X_SCORE0 + t);
      }
    }
  }
  return true;
}

bool build_accel(Graph& g, int version, int H, int W, int K, std::string* err) {
  if (H <= 0 || W <= 0 || H % 128 || W % 128) {
    *err = "frame height and width must be positive multiples of 128";
    return false;
  }
  if (K < 1 || K > 32) { *err = "num_classes must be in [1, 32]"; return false; }
  if (version != 0 && version != 18 && version != 34 && version != 50 && version != 101) {
    *err = "unknown Accel version (expected 0=dff, 18, 34, 50, 101)";
    return false;
  }
  const int h = H / 16, w = W / 16;
  const char* ff = getenv("ACCEL_FOLD_FC6");
  const bool fold_fc6 = !(ff && ff[0] == '0');

  // key frame: R101-DCN + head (get_key_test_symbol)
  {
    Seq& s = g.seq("key");
    int feat = bottleneck_net(g, s, "backbone", H, W, "", units_101(), DeformCfg{18, 1, 1, 1}, X_FEAT_OUT, -1);
    int sc = head(g, s, "head", feat, "fc6", "score", "upsampling", K, X_G_OUT);
    g.tail(s, sc, "", X_LABEL_OUT, X_SCORE_OUT);
  }
  // FlowNet alone (accel_flownet)
  {
    Seq& s = g.seq("flow");
    flownet(g, s, H, W, X_FLOW_OUT);
  }
  // cur frame (get_cur_test_symbol)
  {
    Seq& s = g.seq("cur");
    const int flow = flownet(g, s, H, W, X_NONE);
    if (version == 101) {
      const int cat = g.new_tensor(4096, h, w);
      g.warp(s, X_FEAT_KEY, flow, g.new_view(cat, 0, 2048), X_FEAT_OUT);
      // the correction net only reads the current frame: it runs as its own lane next to FlowNet + warp
      const size_t r0 = s.size();
      bottleneck_net(g, s, "rbranch", H, W, "", units_101(), DeformCfg{18, 1, 1, 1}, X_NONE, g.new_view(cat, 2048, 2048));
      for (size_t i = r0; i < s.size(); ++i) s[i].lane = 1;
      const size_t j0 = s.size();
      int fused = g.conv(s, "fusion", cat, "corr", 2048, 1, 1, 0, 1, bias_epi("corr", ACT_NONE));
      s[j0].join_lane = 1;
      int sc = head(g, s, "head", fused, "fc6", "score", "upsampling", K);
      g.tail(s, sc, "", X_LABEL_OUT, X_SCORE_OUT);
    } else {
      const int warped = g.new_tensor(2048, h, w);
      g.warp(s, X_FEAT_KEY, flow, warped, X_FEAT_OUT);
      const int sl = head(g, s, "head", warped, "fc6", "score", "upsampling", K);
      if (version == 0) {
        g.tail(s, sl, "", X_LABEL_OUT, X_SCORE_OUT);
      } else {
        const size_t r0 = s.size();                       // R branch + R head: a lane of its own (reads only `data`)
        const int sr = rbranch_scores(g, s, version, H, W, K, fold_fc6);
        for (size_t i = r0; i < s.size(); ++i) s[i].lane = 1;
        const int fused = g.new_tensor(K, h, w, true);
        const size_t j0 = s.size();
        g.fuse(s, sl, sr, "corr", fused);
        s[j0].join_lane = 1;
        g.tail(s, fused, "corr_bias", X_LABEL_OUT, X_SCORE_OUT);
      }
    }
  }
  // cur frame with the L head commuted through the warp (DFF, Accel-18/34/50): fc6 is a per-pixel linear map and the
  // bilinear warp a per-channel linear one, so fc6(warp(F)) = warp(W*F) + b.  The key frame hands out G = W*F_key
  // (X_G_OUT above); every cur frame warps the 1024-channel G instead of the 2048-channel feature (half the bytes),
  // adds the bias and the ReLU while converting to the split layout, and never runs the 34 GFLOP fc6 GEMM or the
  // 2048-channel layout conversion.  Chained: G_t = warp_t(G_{t-1}); un-chained: G_key is reused.  Accel-101 needs
  // the warped feature itself for its feature-level fusion and keeps the plan above.
  if (version != 101) {
    Seq& s = g.seq("cur_lin");
    const int flow = flownet(g, s, H, W, X_NONE);
    const int gw = g.new_tensor(1024, h, w);
    g.warp(s, X_G_KEY, flow, gw, X_G_OUT, "fc6_bias", ACT_RELU);
    const int sl = head(g, s, "head", gw, "fc6", "score", "upsampling", K, X_NONE, true);
    if (version == 0) {
      g.tail(s, sl, "", X_LABEL_OUT, X_SCORE_OUT);
    } else {
      const size_t r0 = s.size();
      const int sr = rbranch_scores(g, s, version, H, W, K, fold_fc6);
      for (size_t i = r0; i < s.size(); ++i) s[i].lane = 1;
      const int fused = g.new_tensor(K, h, w, true);
      const size_t j0 = s.size();
      g.fuse(s, sl, sr, "corr", fused);
      s[j0].join_lane = 1;
      g.tail(s, fused, "corr_bias", X_LABEL_OUT, X_SCORE_OUT);
    }
  }
  // the correction network on its own = plain DeepLab-<v> on one frame (accel_rbranch_forward; BASELINE config 1)
  if (version == 18 || version == 34 || version == 50) {
    Seq& s = g.seq("rbranch");
    const int sr = rbranch_scores(g, s, version, H, W, K, fold_fc6);
    g.tail(s, sr, "", X_LABEL_OUT, X_SCORE_OUT);
  }
  return true;
}

}  // namespace accel
