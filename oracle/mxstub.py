"""Test infrastructure: a minimal stand-in for the `mxnet.symbol` API, just wide enough to EXECUTE the reference's own
symbol-building code (dff_deeplab/symbols/accel_{18,34,50,101}.py, resnet_v1_101_flownet_deeplab.py) without MXNet.

Apache MXNet @ 62ecb60 is not in this image, so the reference's graphs cannot be run as they are.  What can be done is
to let the reference's Python build its graph against this stub -- every `mx.sym.Convolution(...)`, `BatchNorm(...)`,
`Crop(...)`, `a + b`, `flow * 2.5` call records a node with exactly the arguments the reference passed -- and then
evaluate the recorded graph with the operator functions of oracle/ops.py.  That pins everything the reference's
files determine: layer order, wiring, parameter NAMES (MXNet's `<name>_weight` / `_bias` / `_gamma` / `_beta` /
`_moving_mean` / `_moving_var` convention), kernel / stride / pad / dilate / eps / fix_gamma / no_bias of every node
and the output names.  What stays outside (and stays tagged [MXNet-ext] in oracle/ops.py) is the arithmetic of each
operator itself.

Operator defaults follow MXNet's documented ones where the reference leaves an argument out: Convolution
`no_bias=False`, Deconvolution `no_bias=True`, BatchNorm `eps=1e-3, fix_gamma=True`, Pooling
`pooling_convention='valid'`, Concat `dim=1`, stride / dilate 1, pad 0.

Only tests/ and tests/golden/make_reference_wired.py import this module (see oracle/__init__.py)."""
from __future__ import annotations

import os
import re
import sys
import types

import torch

from . import ops

_DATA_NAMES = ("data", "data_key", "feat_key", "data_ref", "eq_flag", "label", "im_info")


class Sym:
    _counters = {}

    def __init__(self, op, name=None, inputs=(), attrs=None, outputs=1, index=None):
        if name is None:                                   # MXNet NameManager: lower-cased op name + running index
            hint = op.lower()
            n = Sym._counters.get(hint, 0)
            Sym._counters[hint] = n + 1
            name = "%s%d" % (hint, n)
        self.op, self.name, self.inputs, self.attrs = op, name, list(inputs), dict(attrs or {})
        self.outputs, self.index = outputs, index

    # ---- arithmetic the reference's files use ------------------------------------------------------------------
    def __add__(self, other):
        return Sym("_plus", None, [self, other]) if isinstance(other, Sym) else Sym("_plus_scalar", None, [self], {"scalar": float(other)})

    __radd__ = __add__

    def __mul__(self, other):
        return Sym("_mul", None, [self, other]) if isinstance(other, Sym) else Sym("_mul_scalar", None, [self], {"scalar": float(other)})

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Sym):
            raise NotImplementedError("symbol / symbol")
        return Sym("_div_scalar", None, [self], {"scalar": float(other)})

    __div__ = __truediv__

    def _set_attr(self, **kw):
        pass

    def __getitem__(self, i):
        if self.op == "Group":
            return self.inputs[i]
        return Sym("_item", "%s_item%d" % (self.name, i), [self], {"index": int(i)})

    # ---- introspection (mx.symbol.Symbol surface the reference's Symbol base class uses) ------------------------------
    def _walk(self, seen=None, order=None):
        seen = set() if seen is None else seen
        order = [] if order is None else order
        stack = [(self, False)]
        while stack:
            node, done = stack.pop()
            if done:
                order.append(node)
                continue
            if id(node) in seen:
                continue
            seen.add(id(node))
            stack.append((node, True))
            for inp in reversed(node.inputs):
                if id(inp) not in seen:
                    stack.append((inp, False))
        return order

    def _variables(self):
        return [n for n in self._walk() if n.op == "Variable"]

    def list_arguments(self):
        return [n.name for n in self._variables() if not n.attrs.get("aux")]

    def list_auxiliary_states(self):
        return [n.name for n in self._variables() if n.attrs.get("aux")]

    def list_outputs(self):
        heads = self.inputs if self.op == "Group" else [self]
        return [h.name if h.op == "Variable" else h.name + "_output" for h in heads]

    def get_internals(self):
        return _Internals(self)

    # ---- evaluation with the oracle's operator functions ------------------------------------------------------------
    def eval_dict(self, feed):
        """name -> tensor for every head (MXNet output naming); `feed` maps variable names to torch tensors."""
        memo = {}
        heads = self.inputs if self.op == "Group" else [self]
        return dict(zip(self.list_outputs(), [_eval(h, feed, memo) for h in heads]))


class _Internals:
    def __init__(self, root):
        self.nodes = {}
        for n in root._walk():
            self.nodes[n.name if n.op == "Variable" else n.name + "_output"] = n

    def __getitem__(self, key):
        return self.nodes[key]

    def list_outputs(self):
        return list(self.nodes)


def _pair(v, default):
    if v is None:
        return default
    v = tuple(int(x) for x in v) if isinstance(v, (tuple, list)) else (int(v), int(v))
    assert v[0] == v[1], "anisotropic %s" % (v,)
    return v[0]


def _eval(node, feed, memo):
    order = node._walk()
    for n in order:
        if id(n) in memo:
            continue
        ins = [memo[id(i)] for i in n.inputs]
        memo[id(n)] = _apply(n, ins, feed)
    return memo[id(node)]


def _apply(n, ins, feed):
    a = n.attrs
    op = n.op
    if op == "Variable":
        if n.name not in feed:
            raise KeyError("the reference's graph needs '%s', which was not provided" % n.name)
        return feed[n.name]
    if op == "Convolution":
        x, w = ins[0], ins[1]
        b = ins[2] if len(ins) > 2 else None
        assert w.shape[0] == int(a["num_filter"]) and tuple(w.shape[2:]) == tuple(a["kernel"]), n.name
        assert int(a.get("num_group", 1)) == 1
        return ops.convolution(x, w, b, _pair(a.get("stride"), 1), _pair(a.get("pad"), 0), _pair(a.get("dilate"), 1))
    if op == "Deconvolution":
        x, w = ins[0], ins[1]
        b = ins[2] if len(ins) > 2 else None
        g = int(a.get("num_group", 1))
        assert w.shape[1] * g == int(a["num_filter"]) and tuple(w.shape[2:]) == tuple(a["kernel"]), n.name
        return ops.deconvolution(x, w, b, _pair(a.get("stride"), 1), _pair(a.get("pad"), 0), num_group=g)
    if op == "DeformableConvolution":
        x, off, w = ins[0], ins[1], ins[2]
        assert w.shape[0] == int(a["num_filter"]) and tuple(w.shape[2:]) == tuple(a["kernel"]), n.name
        y = ops.deformable_convolution(x, off, w, _pair(a.get("stride"), 1), _pair(a.get("pad"), 0), _pair(a.get("dilate"), 1),
                                       int(a.get("num_deformable_group", 1)))
        return y + ins[3].view(1, -1, 1, 1) if len(ins) > 3 else y
    if op == "BatchNorm":
        x, gamma, beta, mean, var = ins
        return ops.batch_norm(x, gamma, beta, mean, var, float(a.get("eps", 1e-3)), fix_gamma=bool(a.get("fix_gamma", True)))
    if op == "Activation":
        assert a["act_type"] == "relu", a
        return torch.relu(ins[0])
    if op == "LeakyReLU":
        assert a.get("act_type", "leaky") == "leaky"
        return ops.leaky_relu(ins[0], float(a.get("slope", 0.25)))
    if op == "Pooling":
        assert not a.get("global_pool", False)
        return ops.pooling(ins[0], _pair(a["kernel"], 1), _pair(a.get("stride"), 1), _pair(a.get("pad"), 0), a["pool_type"],
                           a.get("pooling_convention", "valid") == "full")
    if op == "Concat":
        return torch.cat(ins, dim=int(a.get("dim", 1)))
    if op == "Crop":
        assert len(ins) == 2 and not a.get("center_crop", False)
        return ops.crop(ins[0], ins[1], tuple(a.get("offset", (0, 0))))
    if op in ("broadcast_add", "_plus", "ElementWiseSum"):
        out = ins[0]
        for t in ins[1:]:
            out = out + t
        return out
    if op == "_mul":
        return ins[0] * ins[1]
    if op == "_plus_scalar":
        return ins[0] + a["scalar"]
    if op == "_mul_scalar":
        return ins[0] * a["scalar"]
    if op == "_div_scalar":
        return ins[0] / a["scalar"]
    if op == "GridGenerator":
        assert a["transform_type"] == "warp"
        return ops.grid_generator_warp(ins[0])
    if op == "BilinearSampler":
        return ops.bilinear_sampler(ins[0], ins[1])
    raise NotImplementedError("mxstub: operator %s (%s) is not on the inference path" % (op, n.name))


# ---- the `mx.symbol` namespace ---------------------------------------------------------------------------------------
def _sym_inputs(args, kwargs, keys):
    """Positional symbols first (mx.symbol.Crop(*[a, b], ...)), then keyword symbols in MXNet's argument order."""
    ins = [x for x in args if isinstance(x, Sym)]
    for k in keys:
        if isinstance(kwargs.get(k), Sym):
            ins.append(kwargs.pop(k))
    return ins


def Variable(name, **kw):
    return Sym("Variable", name, [], kw)


def _aux(name):
    return Sym("Variable", name, [], {"aux": True})


def _op(opname, params=(), aux=(), no_bias_default=None, data_keys=("data",)):
    def make(*args, **kwargs):
        name = kwargs.pop("name", None)
        node = Sym(opname, name, [], {})
        ins = _sym_inputs(args, kwargs, data_keys)
        for p in params:
            given = kwargs.pop(p, None)
            if p == "bias":
                if bool(kwargs.get("no_bias", no_bias_default)):
                    continue
            ins.append(given if isinstance(given, Sym) else Variable("%s_%s" % (node.name, p)))
        for p in aux:
            ins.append(_aux("%s_%s" % (node.name, p)))
        node.inputs = ins
        node.attrs = kwargs
        return node
    return make


def Group(symbols):
    return Sym("Group", "group", list(symbols), {})


def _unsupported(opname):
    def make(*args, **kwargs):
        return Sym(opname, kwargs.pop("name", None), _sym_inputs(args, kwargs, [k for k, v in list(kwargs.items()) if isinstance(v, Sym)]),
                   kwargs)
    return make


def make_mxnet_module():
    """A module object that can be installed as `mxnet` (sys.modules) while the reference's files are executed."""
    sym = types.ModuleType("mxnet.symbol")
    sym.Variable = Variable
    sym.Group = Group
    sym.Convolution = _op("Convolution", params=("weight", "bias"), no_bias_default=False)
    sym.Deconvolution = _op("Deconvolution", params=("weight", "bias"), no_bias_default=True)
    sym.BatchNorm = _op("BatchNorm", params=("gamma", "beta"), aux=("moving_mean", "moving_var"))
    for plain in ("Activation", "LeakyReLU", "Pooling", "Concat", "Crop", "broadcast_add", "ElementWiseSum"):
        setattr(sym, plain, _op(plain))
    sym.GridGenerator = _op("GridGenerator")
    sym.BilinearSampler = _op("BilinearSampler", data_keys=("data", "grid"))
    for off_path in ("Custom", "Reshape", "SliceChannel", "split", "SoftmaxActivation", "SoftmaxOutput", "MultiProposal", "PSROIPooling",
                     "BlockGrad", "MakeLoss", "ROIPooling", "smooth_l1", "slice_axis", "tile", "mean", "sum", "sqrt", "square"):
        setattr(sym, off_path, _unsupported(off_path))
    contrib_sym = types.ModuleType("mxnet.contrib.symbol")
    contrib_sym.DeformableConvolution = _op("DeformableConvolution", params=("weight", "bias"), no_bias_default=False,
                                            data_keys=("data", "offset"))
    for off_path in ("MultiProposal", "PSROIPooling", "DeformablePSROIPooling", "Proposal"):
        setattr(contrib_sym, off_path, _unsupported(off_path))
    contrib = types.ModuleType("mxnet.contrib")
    contrib.symbol = contrib.sym = contrib_sym
    mx = types.ModuleType("mxnet")
    mx.symbol = mx.sym = sym
    mx.contrib = contrib
    mx.operator = types.SimpleNamespace(CustomOp=object, CustomOpProp=object, register=lambda name: (lambda cls: cls))
    return mx


_PRINT_STMT = re.compile(r"^(\s*)print (?!\()(.*)$", re.M)


def load_reference_symbols(ref_root):
    """Executes the reference's symbol files against the stub and returns {'18': accel_18, ...} (the CLASSES the
    reference defines).  Nothing is copied: the files are read from `ref_root` and executed in memory; Python 2
    `print x` statements (accel_101.py:287, off the inference path) are rewritten to `print(x)` on the fly."""
    Sym._counters.clear()
    saved = {k: sys.modules.get(k) for k in ("mxnet", "cPickle", "utils", "utils.symbol", "operator_py", "operator_py.proposal",
                                             "operator_py.proposal_target", "operator_py.box_annotator_ohem",
                                             "operator_py.rpn_inv_normalize", "operator_py.tile_as", "resnet_v1_101_flownet_deeplab")}

    def run(path, modname):
        src = _PRINT_STMT.sub(r"\1print(\2)", open(os.path.join(ref_root, path)).read())
        mod = types.ModuleType(modname)
        mod.__file__ = os.path.join(ref_root, path)
        sys.modules[modname] = mod
        exec(compile(src, mod.__file__, "exec"), mod.__dict__)
        return mod

    try:
        sys.modules["mxnet"] = make_mxnet_module()
        sys.modules["cPickle"] = types.ModuleType("cPickle")
        sys.modules["utils"] = types.ModuleType("utils")
        run("lib/utils/symbol.py", "utils.symbol")                       # the reference's own Symbol base class
        sys.modules["operator_py"] = types.ModuleType("operator_py")
        for m in ("proposal", "proposal_target", "box_annotator_ohem", "rpn_inv_normalize", "tile_as"):
            sys.modules["operator_py." + m] = types.ModuleType("operator_py." + m)      # CustomOps of the detection fork
        run("dff_deeplab/symbols/resnet_v1_101_flownet_deeplab.py", "resnet_v1_101_flownet_deeplab")
        out = {}
        for v in ("18", "34", "50", "101"):
            mod = run("dff_deeplab/symbols/accel_%s.py" % v, "ref_accel_%s" % v)
            out[v] = getattr(mod, "accel_%s" % v)
            sys.modules.pop("ref_accel_%s" % v, None)
        return out
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def reference_config(num_classes=19):
    """The fields of the reference's `config` object that get_key_test_symbol / get_cur_test_symbol read."""
    ns = types.SimpleNamespace
    return ns(dataset=ns(NUM_CLASSES=num_classes), CLASS_AGNOSTIC=True, network=ns(NUM_ANCHORS=9),
              TEST=ns(BATCH_IMAGES=1, KEY_FRAME_INTERVAL=5), TRAIN=ns(KEY_INTERVAL=5))
