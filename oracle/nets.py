"""Reference graph wiring restated on top of oracle.ops (PyTorch fp32, CPU).  Test infrastructure.

`p` is a flat dict {reference parameter name -> torch.float32 tensor}: `arg:` and `aux:` entries of
the reference checkpoints merged, as `load_param` + `init_params` do (lib/utils/load_model.py:73-93,
dff_deeplab/demo.py:192-195).  Graphs are executed as written: no BN folding, no algebraic
shortcuts, full-resolution 38->19 fusion.
"""
from __future__ import annotations

import torch

from . import ops

SYM = "dff_deeplab/symbols/resnet_v1_101_flownet_deeplab.py"


def _bn(p, x, name, eps, fix_gamma=False):
    return ops.batch_norm(x, p[name + "_gamma"], p[name + "_beta"], p[name + "_moving_mean"],
                          p[name + "_moving_var"], eps, fix_gamma)


def _conv(p, x, name, stride=1, pad=0, dilate=1, bias=False):
    return ops.convolution(x, p[name + "_weight"], p[name + "_bias"] if bias else None, stride, pad, dilate)


# ----------------------------------------------------------------------------- FlowNet-S
def flownet(p, img_cur, img_ref):
    """get_flownet, SYM:1751-1808.  Returns flow*2.5 (N,2,H/16,W/16); the `Convolution5_scale`
    branch is computed by the reference but discarded by every Accel graph (accel_18.py:173)."""
    lk = ops.leaky_relu
    data = torch.cat([img_cur / 255.0, img_ref / 255.0], dim=1)
    x = ops.pooling(data, 2, 2, 0, "avg", full=True)                      # resize_data
    r1 = lk(_conv(p, x, "flow_conv1", 2, 3, bias=True))
    r2 = lk(_conv(p, r1, "conv2", 2, 2, bias=True))
    r3 = lk(_conv(p, r2, "conv3", 2, 2, bias=True))
    r4 = lk(_conv(p, r3, "conv3_1", 1, 1, bias=True))
    r5 = lk(_conv(p, r4, "conv4", 2, 1, bias=True))
    r6 = lk(_conv(p, r5, "conv4_1", 1, 1, bias=True))
    r7 = lk(_conv(p, r6, "conv5", 2, 1, bias=True))
    r8 = lk(_conv(p, r7, "conv5_1", 1, 1, bias=True))
    r9 = lk(_conv(p, r8, "conv6", 2, 1, bias=True))
    r10 = lk(_conv(p, r9, "conv6_1", 1, 1, bias=True))

    def refine(feat, skip, flow_name, deconv_name, up_name):
        flow = _conv(p, feat, flow_name, 1, 1, bias=True)
        d = ops.deconvolution(feat, p[deconv_name + "_weight"], p[deconv_name + "_bias"], 2, 0)
        d = lk(ops.crop(d, skip, (1, 1)))
        u = ops.deconvolution(flow, p[up_name + "_weight"], p[up_name + "_bias"], 2, 0)
        u = ops.crop(u, skip, (1, 1))
        return torch.cat([skip, d, u], dim=1)

    c2 = refine(r10, r8, "Convolution1", "deconv5", "upsample_flow6to5")
    c3 = refine(c2, r6, "Convolution2", "deconv4", "upsample_flow5to4")
    c4 = refine(c3, r4, "Convolution3", "deconv3", "upsample_flow4to3")
    c5 = refine(c4, r2, "Convolution4", "deconv2", "upsample_flow3to2")
    c5 = ops.pooling(c5, 2, 2, 0, "avg", full=True)                       # resize_concat5
    return _conv(p, c5, "Convolution5", 1, 1, bias=True) * 2.5


# ----------------------------------------------------------------------------- pre-activation trunk
def _residual_unit(p, x, name, stride, dim_match):
    """residual_unit(bottle_neck=False), SYM:70-86."""
    act1 = torch.relu(_bn(p, x, name + "_bn1", 2e-5))
    c1 = _conv(p, act1, name + "_conv1", stride, 1)
    act2 = torch.relu(_bn(p, c1, name + "_bn2", 2e-5))
    c2 = _conv(p, act2, name + "_conv2", 1, 1)
    shortcut = x if dim_match else _conv(p, act1, name + "_sc", stride, 0)
    return c2 + shortcut


def resnet_preact(p, data, prefix, units):
    """resnet(data_type='imagenet', bottle_neck=False, num_stages=3), SYM:88-130."""
    x = _bn(p, data, prefix + "bn_data", 2e-5, fix_gamma=True)
    x = _conv(p, x, prefix + "conv0", 2, 3)
    x = torch.relu(_bn(p, x, prefix + "bn0", 2e-5))
    x = ops.pooling(x, 3, 2, 1, "max", full=False)
    for i, n_units in enumerate(units):
        s = 1 if i == 0 else 2
        x = _residual_unit(p, x, "%sstage%d_unit%d" % (prefix, i + 1, 1), s, False)
        for j in range(n_units - 1):
            x = _residual_unit(p, x, "%sstage%d_unit%d" % (prefix, i + 1, j + 2), 1, True)
    return x


def resnet_dcn_basic_conv5(p, feat, prefix, letters):
    """get_resnet_dcn_18_conv5 (SYM:132-170, letters 'ab') / get_resnet_dcn_34_conv5 (SYM:172-233,
    letters 'abc'): post-activation basic units at stride 32 with a deformable 3x3 dil-2 second conv."""
    eps = 1e-5
    x = feat
    for n, L in enumerate(letters):
        first = n == 0
        if first:
            sc = _bn(p, _conv(p, x, "%sres5%s_branch1" % (prefix, L), 2, 0), "%sbn5%s_branch1" % (prefix, L), eps)
        else:
            sc = x
        a = _conv(p, x, "%sres5%s_branch2a" % (prefix, L), 2 if first else 1, 1)
        a = torch.relu(_bn(p, a, "%sbn5%s_branch2a" % (prefix, L), eps))
        off = _conv(p, a, "%sres5%s_branch2b_offset" % (prefix, L), 1, 2, 2, bias=True)
        b = ops.deformable_convolution(a, off, p["%sres5%s_branch2b_weight" % (prefix, L)], 1, 2, 2, 4)
        b = _bn(p, b, "%sbn5%s_branch2b" % (prefix, L), eps)
        x = torch.relu(sc + b)
    return x


# ----------------------------------------------------------------------------- caffe-style bottleneck nets
def _bottleneck(p, x, prefix, unit, stride, project, deform):
    """One res{unit} block of get_resnet_dcn_50 / get_resnet_dcn (post-activation bottleneck; the
    stride sits on the 1x1 branch2a and branch1 convs, e.g. SYM:305-309,647-652).
    deform = None | (offset_channels, offset_pad, offset_dilate, num_deformable_group)."""
    eps = 1e-5
    r, b = prefix + "res" + unit, prefix + "bn" + unit
    sc = _bn(p, _conv(p, x, r + "_branch1", stride, 0), b + "_branch1", eps) if project else x
    a = torch.relu(_bn(p, _conv(p, x, r + "_branch2a", stride, 0), b + "_branch2a", eps))
    if deform is None:
        m = _conv(p, a, r + "_branch2b", 1, 1)
    else:
        _, opad, odil, dg = deform
        off = _conv(p, a, r + "_branch2b_offset", 1, opad, odil, bias=True)
        m = ops.deformable_convolution(a, off, p[r + "_branch2b_weight"], 1, 2, 2, dg)
    m = torch.relu(_bn(p, m, b + "_branch2b", eps))
    c = _bn(p, _conv(p, m, r + "_branch2c", 1, 0), b + "_branch2c", eps)
    return torch.relu(sc + c)


def _bottleneck_net(p, data, prefix, stage_units, deform):
    x = _conv(p, data, prefix + "conv1", 2, 3)
    x = torch.relu(_bn(p, x, prefix + "bn_conv1", 1e-5))
    x = ops.pooling(x, 3, 2, 0, "max", full=True)
    for stage, units in zip((2, 3, 4, 5), stage_units):
        for n, u in enumerate(units):
            stride = 2 if (n == 0 and stage in (3, 4)) else 1
            x = _bottleneck(p, x, prefix, "%d%s" % (stage, u), stride, n == 0, deform if stage == 5 else None)
    return x


def resnet_dcn_50(p, data):
    """get_resnet_dcn_50, SYM:235-574: [3,4,6,3]; res5 offsets 3x3 dil2 pad2 -> 72 ch, dg=4."""
    units = (("a", "b", "c"), ("a", "b", "c", "d"), ("a", "b", "c", "d", "e", "f"), ("a", "b", "c"))
    return _bottleneck_net(p, data, "50_", units, (72, 2, 2, 4))


def resnet_dcn_101(p, data):
    """get_resnet_dcn, SYM:576-1300: [3,4,23,3]; res5 offsets plain 3x3 pad1 -> 18 ch, dg=1
    (SYM:1230-1237)."""
    units = (("a", "b", "c"), ("a", "b1", "b2", "b3"), ("a",) + tuple("b%d" % i for i in range(1, 23)),
             ("a", "b", "c"))
    return _bottleneck_net(p, data, "", units, (18, 1, 1, 1))


def resnet_dcn_101_layers(p, data):
    """resnet_dcn_101 with every activation kept under the name of the layer that produces it (conv + BN [+ ReLU]
    [+ shortcut] fused, i.e. what the CUDA plan's op of that name leaves behind): per-layer parity through
    accel_debug_fetch (tests/test_gpu_layers.py, tools/layer_parity.py).  Same code path as _bottleneck_net."""
    acts = {}
    eps = 1e-5
    x = torch.relu(_bn(p, _conv(p, data, "conv1", 2, 3), "bn_conv1", eps))
    acts["conv1"] = x
    x = ops.pooling(x, 3, 2, 0, "max", full=True)
    acts["maxpool"] = x
    units = (("a", "b", "c"), ("a", "b1", "b2", "b3"), ("a",) + tuple("b%d" % i for i in range(1, 23)), ("a", "b", "c"))
    for stage, us in zip((2, 3, 4, 5), units):
        for n, u in enumerate(us):
            stride = 2 if (n == 0 and stage in (3, 4)) else 1
            r, b = "res%d%s" % (stage, u), "bn%d%s" % (stage, u)
            sc = x
            if n == 0:
                sc = _bn(p, _conv(p, x, r + "_branch1", stride, 0), b + "_branch1", eps)
                acts[r + "_branch1"] = sc
            a = torch.relu(_bn(p, _conv(p, x, r + "_branch2a", stride, 0), b + "_branch2a", eps))
            acts[r + "_branch2a"] = a
            if stage == 5:
                off = _conv(p, a, r + "_branch2b_offset", 1, 1, 1, bias=True)
                m = ops.deformable_convolution(a, off, p[r + "_branch2b_weight"], 1, 2, 2, 1)
            else:
                m = _conv(p, a, r + "_branch2b", 1, 1)
            m = torch.relu(_bn(p, m, b + "_branch2b", eps))
            acts[r + "_branch2b"] = m
            x = torch.relu(sc + _bn(p, _conv(p, m, r + "_branch2c", 1, 0), b + "_branch2c", eps))
            acts[r + "_branch2c"] = x
    return acts


# ----------------------------------------------------------------------------- heads and graphs
def deeplab_head(p, feat, data, fc6, score, upsampling):
    """fc6 1x1 -> ReLU -> score 1x1 -> grouped 32x32/s16 deconv -> Crop(8,8)  (accel_18.py:177-197)."""
    x = torch.relu(_conv(p, feat, fc6, bias=True))
    s = _conv(p, x, score, bias=True)
    up = ops.deconvolution(s, p[upsampling + "_weight"], None, 16, 0, num_group=s.shape[1])
    return ops.crop(up, data, (8, 8)), s


def key_forward(p, data):
    """get_key_test_symbol (accel_18.py:121-159, same in 34/50/101).  Output names as :157."""
    feat = resnet_dcn_101(p, data)
    croped, _ = deeplab_head(p, feat, data, "fc6", "score", "upsampling")
    return {"res5c_relu_output": feat, "croped_score_output": croped}


def warp(p, data_cur, data_key, feat_key):
    """flow -> GridGenerator(warp) -> BilinearSampler  (accel_18.py:172-175)."""
    flow = flownet(p, data_cur, data_key)
    return ops.bilinear_sampler(feat_key, ops.grid_generator_warp(flow)), flow


def cur_forward(p, version, data, data_key, feat_key):
    """get_cur_test_symbol of accel_{18,34,50,101}.py, plus 'dff' = the L branch alone (Deep Feature
    Flow: FlowNet + warp + task head, BASELINE.json configs[1]).  Output names follow
    accel_18.py:237 / accel_101.py:191."""
    version = str(version)
    warped, flow = warp(p, data, data_key, feat_key)
    out = {"warping_feat_output": warped, "flow": flow}
    if version == "101":                                           # accel_101.py:160-189
        feat_cur = resnet_dcn_101(p, data)
        fused = _conv(p, torch.cat([warped, feat_cur], dim=1), "corr", bias=True)
        out["croped_score_output"], _ = deeplab_head(p, fused, data, "fc6", "score", "upsampling")
        return out
    croped, _ = deeplab_head(p, warped, data, "fc6", "score", "upsampling")
    if version == "dff":
        out["croped_score_output"] = croped
        return out
    cur_croped = rbranch_forward(p, version, data)["croped_score_output"]
    out["correction_output"] = _conv(p, torch.cat([croped, cur_croped], dim=1), "corr", bias=True)
    return out


def rbranch_forward(p, version, data):
    """The correction network with its own DeepLab head, alone: accel_18.py:199-227 (R18 / R34 trunk + deformable
    conv5 + `feat_upsampling` + `<v>_fc6/score/upsampling`), accel_50.py:195-216 (R50-DCN + `curr_*` head).  On its
    own this is the plain DeepLab-<v> net of BASELINE config 1 (deeplab/core/tester.py:84-85 argmaxes its softmax)."""
    version = str(version)
    if version in ("18", "34"):                                    # accel_18.py:199-227
        pre = version + "_"
        f = resnet_preact(p, data, pre, [2, 2, 2] if version == "18" else [3, 4, 6])
        f = resnet_dcn_basic_conv5(p, f, pre, "ab" if version == "18" else "abc")
        f = ops.deconvolution(f, p[pre + "feat_upsampling_weight"], None, 2, 1)
        names = (pre + "fc6", pre + "score", pre + "upsampling")
    elif version == "50":                                          # accel_50.py:195-216
        f = resnet_dcn_50(p, data)
        names = ("curr_fc6", "curr_score", "curr_upsampling")
    else:
        raise ValueError("no stand-alone correction network for version %r" % version)
    cur_croped, lowres = deeplab_head(p, f, data, *names)
    return {"croped_score_output": cur_croped, "score_lowres": lowres}


def output_key(version):
    """dff_deeplab/demo.py:244."""
    return "croped_score_output" if str(version) in ("101", "dff") else "correction_output"
