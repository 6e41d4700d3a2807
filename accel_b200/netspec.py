"""Parameter inventory of the Accel inference graphs, keyed by the reference's parameter names.

This is the Python-side statement of which tensors `accel_set_param` must receive for a given Accel
version (the C library enumerates the same list from its own graph builder through
`accel_param_spec`; tests/test_abi_cpu.py checks the two agree name by name).

Names follow /root/reference/dff_deeplab/symbols/resnet_v1_101_flownet_deeplab.py and
accel_{18,34,50,101}.py.  `kind` drives the synthetic initialiser only.
"""
from __future__ import annotations

from collections import OrderedDict

VERSIONS = ("dff", "18", "34", "50", "101")
NUM_CLASSES = 19
FEAT_DIM = 2048          # config.network.DFF_FEAT_DIM (dff_deeplab/config/config.py:35)


class _Spec(OrderedDict):
    def conv(self, name, cout, cin, k, bias=False, kind="conv"):
        self[name + "_weight"] = ((cout, cin, k, k), kind)
        if bias:
            self[name + "_bias"] = ((cout,), "bias")

    def deconv(self, name, cin, cout, k, bias=False, kind="deconv"):
        self[name + "_weight"] = ((cin, cout, k, k), kind)
        if bias:
            self[name + "_bias"] = ((cout,), "bias")

    def bn(self, name, c, fixed_gamma=False):
        self[name + "_gamma"] = ((c,), "gamma_one" if fixed_gamma else "gamma")
        self[name + "_beta"] = ((c,), "beta")
        self[name + "_moving_mean"] = ((c,), "mean")
        self[name + "_moving_var"] = ((c,), "var")


def _flownet(s):
    # get_flownet, ...flownet_deeplab.py:1754-1807 (all convs carry a bias)
    for name, cout, cin, k in (("flow_conv1", 64, 6, 7), ("conv2", 128, 64, 5), ("conv3", 256, 128, 5),
                               ("conv3_1", 256, 256, 3), ("conv4", 512, 256, 3), ("conv4_1", 512, 512, 3),
                               ("conv5", 512, 512, 3), ("conv5_1", 512, 512, 3), ("conv6", 1024, 512, 3),
                               ("conv6_1", 1024, 1024, 3)):
        s.conv(name, cout, cin, k, bias=True)
    for flow, cin, deconv, dout, up in (("Convolution1", 1024, "deconv5", 512, "upsample_flow6to5"),
                                        ("Convolution2", 1026, "deconv4", 256, "upsample_flow5to4"),
                                        ("Convolution3", 770, "deconv3", 128, "upsample_flow4to3"),
                                        ("Convolution4", 386, "deconv2", 64, "upsample_flow3to2")):
        s.conv(flow, 2, cin, 3, bias=True, kind="flow_head")
        s.deconv(deconv, cin, dout, 4, bias=True)
        s.deconv(up, 2, 2, 4, bias=True, kind="flow_up")
    s.conv("Convolution5", 2, 194, 3, bias=True, kind="flow_head")


def _head(s, fc6, score, upsampling, cin=FEAT_DIM):
    s.conv(fc6, 1024, cin, 1, bias=True)
    s.conv(score, NUM_CLASSES, 1024, 1, bias=True, kind="score")
    s[upsampling + "_weight"] = ((NUM_CLASSES, 1, 32, 32), "bilinear")


def _bottleneck_net(s, prefix, stage_units, off_ch, off_k):
    s.conv(prefix + "conv1", 64, 3, 7)
    s.bn(prefix + "bn_conv1", 64)
    cin = 64
    for stage, units, mid in zip((2, 3, 4, 5), stage_units, (64, 128, 256, 512)):
        for n, u in enumerate(units):
            r, b = "%sres%d%s" % (prefix, stage, u), "%sbn%d%s" % (prefix, stage, u)
            if n == 0:
                s.conv(r + "_branch1", 4 * mid, cin, 1)
                s.bn(b + "_branch1", 4 * mid)
            s.conv(r + "_branch2a", mid, cin, 1)
            s.bn(b + "_branch2a", mid)
            if stage == 5:
                s.conv(r + "_branch2b_offset", off_ch, mid, off_k, bias=True, kind="offset")
            s.conv(r + "_branch2b", mid, mid, 3)
            s.bn(b + "_branch2b", mid)
            s.conv(r + "_branch2c", 4 * mid, mid, 1)
            s.bn(b + "_branch2c", 4 * mid)
            cin = 4 * mid


def _r101(s):
    units = (("a", "b", "c"), ("a", "b1", "b2", "b3"), ("a",) + tuple("b%d" % i for i in range(1, 23)), ("a", "b", "c"))
    _bottleneck_net(s, "", units, 18, 3)


def _r50(s):
    units = (("a", "b", "c"), ("a", "b", "c", "d"), ("a", "b", "c", "d", "e", "f"), ("a", "b", "c"))
    _bottleneck_net(s, "50_", units, 72, 3)


def _preact(s, prefix, units, letters):
    # resnet(...) :88-130 and get_resnet_dcn_{18,34}_conv5 :132-233
    s.bn(prefix + "bn_data", 3, fixed_gamma=True)
    s.conv(prefix + "conv0", 64, 3, 7)
    s.bn(prefix + "bn0", 64)
    cin = 64
    for i, (n_units, c) in enumerate(zip(units, (64, 128, 256))):
        for j in range(n_units):
            name = "%sstage%d_unit%d" % (prefix, i + 1, j + 1)
            s.bn(name + "_bn1", cin)
            s.conv(name + "_conv1", c, cin, 3)
            s.bn(name + "_bn2", c)
            s.conv(name + "_conv2", c, c, 3)
            if j == 0:
                s.conv(name + "_sc", c, cin, 1)
            cin = c
    for n, L in enumerate(letters):
        if n == 0:
            s.conv("%sres5%s_branch1" % (prefix, L), 512, cin, 1)
            s.bn("%sbn5%s_branch1" % (prefix, L), 512)
        s.conv("%sres5%s_branch2a" % (prefix, L), 512, cin, 3)
        s.bn("%sbn5%s_branch2a" % (prefix, L), 512)
        s.conv("%sres5%s_branch2b_offset" % (prefix, L), 72, 512, 3, bias=True, kind="offset")
        s.conv("%sres5%s_branch2b" % (prefix, L), 512, 512, 3)
        s.bn("%sbn5%s_branch2b" % (prefix, L), 512)
        cin = 512
    s.deconv(prefix + "feat_upsampling", 512, FEAT_DIM, 4)


def param_spec(version):
    """OrderedDict name -> (shape, kind) for everything the key graph and the `version` cur graph
    read: R101-DCN key net + L head (always), FlowNet, the R branch of `version`, the fusion conv."""
    version = str(version)
    if version not in VERSIONS:
        raise ValueError("unknown Accel version %r (expected one of %s)" % (version, ", ".join(VERSIONS)))
    s = _Spec()
    _r101(s)
    _head(s, "fc6", "score", "upsampling")
    _flownet(s)
    if version in ("18", "34"):
        pre = version + "_"
        _preact(s, pre, [2, 2, 2] if version == "18" else [3, 4, 6], "ab" if version == "18" else "abc")
        _head(s, pre + "fc6", pre + "score", pre + "upsampling")
    elif version == "50":
        _r50(s)
        _head(s, "curr_fc6", "curr_score", "curr_upsampling")
    if version in ("18", "34", "50"):
        s.conv("corr", NUM_CLASSES, 2 * NUM_CLASSES, 1, bias=True, kind="corr_score")
    elif version == "101":
        s.conv("corr", FEAT_DIM, 2 * FEAT_DIM, 1, bias=True, kind="corr_feat")
    return s
