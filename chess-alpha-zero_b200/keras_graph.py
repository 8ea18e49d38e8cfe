"""Keras ``Model.get_config()`` JSON  <->  evaluator architecture, and weight-file I/O.

What ``ChessModel.load`` hands to Keras in the reference
(/root/reference/src/chess_zero/agent/model_chess.py:121-153):
    self.model = Model.from_config(json.load(f)); self.model.load_weights(weight_path)
is replaced by: walk the layer graph of the JSON (config-driven: kernel sizes, filter counts, block
count, BN epsilon, Dense widths all come from the JSON, never hard-coded), check that it
IS the graph ChessModel.build makes (model_chess.py:57-111), and emit
  * a ``caz_arch`` for the C ABI, and
  * the ordered list of Keras weight names the ABI expects (include/caz_b200.h).

``build_config`` is the inverse, used by ``ChessModel.build``/``save``: it emits the JSON in the schema of
the one sample the reference ships (data/model/model_best_config.json).
"""
import json
import os
import zipfile
from collections import OrderedDict

import numpy as np

from ._lib import Arch

BN_PARTS = ("gamma:0", "beta:0", "moving_mean:0", "moving_variance:0")


class GraphError(ValueError):
    pass


def _need(cond, msg):
    if not cond:
        raise GraphError(msg)


class _Graph:
    def __init__(self, cfg):
        if isinstance(cfg, dict) and "layers" not in cfg and isinstance(cfg.get("config"), dict):
            cfg = cfg["config"]  # newer Keras wraps the functional config one level down
        _need(isinstance(cfg, dict) and "layers" in cfg, "not a Keras functional-model config")
        self.cfg = cfg
        self.layers = OrderedDict((l["name"], l) for l in cfg["layers"])
        self.consumers = {}
        for l in cfg["layers"]:
            for src in self.inbound(l["name"]):
                self.consumers.setdefault(src, []).append(l["name"])

    def inbound(self, name):
        nodes = self.layers[name]["inbound_nodes"]
        return [n[0] for n in nodes[0]] if nodes else []

    def cls(self, name):
        return self.layers[name]["class_name"]

    def conf(self, name):
        return self.layers[name]["config"]

    def only_consumer(self, name, cls):
        outs = [c for c in self.consumers.get(name, []) if self.cls(c) == cls]
        _need(len(outs) == 1, f"expected exactly one {cls} after {name}, found {outs}")
        return outs[0]


def _check_conv(g, name, padding):
    c = g.conf(name)
    _need(g.cls(name) == "Conv2D", f"{name}: expected Conv2D")
    _need(c.get("data_format") == "channels_first", f"{name}: data_format must be channels_first")
    _need(not c.get("use_bias"), f"{name}: use_bias must be false (BatchNorm follows)")
    _need(list(c.get("strides", [1, 1])) == [1, 1] and list(c.get("dilation_rate", [1, 1])) == [1, 1],
          f"{name}: strides/dilation must be 1")
    _need(c.get("activation", "linear") == "linear", f"{name}: activation must be linear")
    k = list(c["kernel_size"])
    _need(k[0] == k[1] and k[0] % 2 == 1, f"{name}: kernel must be odd and square")
    _need(c["padding"] == padding or k[0] == 1, f"{name}: padding must be {padding}")
    return int(c["filters"]), int(k[0])


def _check_bn(g, name):
    c = g.conf(name)
    _need(g.cls(name) == "BatchNormalization", f"{name}: expected BatchNormalization")
    _need(c.get("axis", 1) == 1, f"{name}: BatchNormalization axis must be 1 (channels_first)")
    _need(c.get("center", True) and c.get("scale", True), f"{name}: center/scale must be on")
    return float(c.get("epsilon", 1e-3))


def _conv_bn_act(g, src, act, padding="same"):
    conv = g.only_consumer(src, "Conv2D")
    filters, k = _check_conv(g, conv, padding)
    bn = g.only_consumer(conv, "BatchNormalization")
    eps = _check_bn(g, bn)
    return conv, bn, filters, k, eps


def parse_config(cfg):
    """Returns (Arch, ordered Keras weight names) for the graph of ChessModel.build."""
    g = _Graph(cfg)
    cfg = g.cfg
    _need(len(cfg["input_layers"]) == 1 and len(cfg["output_layers"]) == 2, "expected 1 input and 2 outputs")
    inp = cfg["input_layers"][0][0]
    shape = g.conf(inp)["batch_input_shape"]
    _need(list(shape[2:]) == [8, 8], "input must be (planes, 8, 8)")
    names = []
    eps_seen = set()

    def take(conv, bn, eps):
        names.append(f"{conv}/kernel:0")
        names.extend(f"{bn}/{p}" for p in BN_PARTS)
        eps_seen.add(round(eps, 9))

    # stem (model_chess.py:65-69)
    conv, bn, filters, stem_k, eps = _conv_bn_act(g, inp, "relu")
    take(conv, bn, eps)
    x = g.only_consumer(bn, "Activation")
    _need(g.conf(x)["activation"] == "relu", "stem activation must be relu")
    # residual blocks (model_chess.py:96-111): x -> conv,bn,relu,conv,bn -> Add([x, bn2]) -> relu
    blocks, block_k = 0, 3
    while True:
        adds = [c for c in g.consumers.get(x, []) if g.cls(c) == "Add"]
        if not adds:
            break
        _need(len(adds) == 1, f"{x}: more than one Add")
        convs = [c for c in g.consumers[x] if g.cls(c) == "Conv2D"]
        _need(len(convs) == 1, f"{x}: a residual block has exactly one conv on its input")
        c1 = convs[0]
        f1, k1 = _check_conv(g, c1, "same")
        b1 = g.only_consumer(c1, "BatchNormalization")
        e1 = _check_bn(g, b1)
        r1 = g.only_consumer(b1, "Activation")
        _need(g.conf(r1)["activation"] == "relu", f"{r1}: must be relu")
        c2 = g.only_consumer(r1, "Conv2D")
        f2, k2 = _check_conv(g, c2, "same")
        b2 = g.only_consumer(c2, "BatchNormalization")
        e2 = _check_bn(g, b2)
        _need(f1 == filters and f2 == filters and k1 == k2, "residual block filter/kernel mismatch")
        _need(sorted(g.inbound(adds[0])) == sorted([x, b2]), f"{adds[0]}: must add the block input and bn2")
        _need(blocks == 0 or k1 == block_k, "all residual blocks must share one kernel size")
        block_k = k1
        take(c1, b1, e1)
        take(c2, b2, e2)
        x = g.only_consumer(adds[0], "Activation")
        _need(g.conf(x)["activation"] == "relu", f"{x}: must be relu")
        blocks += 1
    # heads (model_chess.py:77-92), identified from the two outputs backwards
    pol_out, val_out = cfg["output_layers"][0][0], cfg["output_layers"][1][0]
    _need(g.cls(pol_out) == "Dense" and g.conf(pol_out)["activation"] == "softmax", "output 0 must be the softmax policy")
    _need(g.cls(val_out) == "Dense" and g.conf(val_out)["activation"] == "tanh" and g.conf(val_out)["units"] == 1,
          "output 1 must be the tanh value")

    def head_chain(dense):
        flat = g.inbound(dense)[0]
        _need(g.cls(flat) == "Flatten", f"{dense}: expected Flatten before it")
        act = g.inbound(flat)[0]
        _need(g.cls(act) == "Activation" and g.conf(act)["activation"] == "relu", f"{flat}: expected relu before it")
        bn_ = g.inbound(act)[0]
        e = _check_bn(g, bn_)
        conv_ = g.inbound(bn_)[0]
        f, k = _check_conv(g, conv_, "valid")
        _need(k == 1, f"{conv_}: head convolutions are 1x1")
        _need(g.inbound(conv_) == [x], f"{conv_}: heads must hang off the last residual activation")
        return conv_, bn_, f, e

    p_conv, p_bn, pf, pe = head_chain(pol_out)
    val_dense = g.inbound(val_out)[0]
    _need(g.cls(val_dense) == "Dense" and g.conf(val_dense)["activation"] == "relu", "value_dense must be Dense(relu)")
    v_conv, v_bn, vf, ve = head_chain(val_dense)
    for d in (pol_out, val_dense, val_out):
        _need(g.conf(d).get("use_bias", True), f"{d}: Dense layers carry a bias")
    take(p_conv, p_bn, pe)
    names.extend([f"{pol_out}/kernel:0", f"{pol_out}/bias:0"])
    take(v_conv, v_bn, ve)
    names.extend([f"{val_dense}/kernel:0", f"{val_dense}/bias:0", f"{val_out}/kernel:0", f"{val_out}/bias:0"])
    _need(len(eps_seen) == 1, f"all BatchNormalization layers must share one epsilon, found {sorted(eps_seen)}")
    arch = Arch(input_planes=int(shape[1]), stem_kernel=stem_k, filters=filters, res_blocks=blocks,
                block_kernel=block_k, policy_filters=pf, value_filters=vf,
                value_fc=int(g.conf(val_dense)["units"]), n_labels=int(g.conf(pol_out)["units"]),
                bn_epsilon=float(eps_seen.pop()))
    return arch, names


def keras_layer_order(cfg):
    """Weight names in the order Keras ``save_weights``/``load_weights`` walks them (layer list order)."""
    g = _Graph(cfg)
    names = []
    for l in g.cfg["layers"]:
        n, c = l["name"], l["class_name"]
        if c == "Conv2D":
            names.append(f"{n}/kernel:0")
        elif c == "BatchNormalization":
            names.extend(f"{n}/{p}" for p in BN_PARTS)
        elif c == "Dense":
            names.extend([f"{n}/kernel:0", f"{n}/bias:0"])
    return names


def expected_shapes(cfg):
    """name -> shape for every weight of the graph (used to validate files and to initialise build())."""
    arch, _ = parse_config(cfg)
    g = _Graph(cfg)
    shapes = OrderedDict()
    chans = {}
    for l in g.cfg["layers"]:
        n, c, conf = l["name"], l["class_name"], l["config"]
        src = g.inbound(n)
        if c == "InputLayer":
            chans[n] = ("chw", arch.input_planes)
        elif c == "Conv2D":
            k = conf["kernel_size"][0]
            shapes[f"{n}/kernel:0"] = (k, k, chans[src[0]][1], conf["filters"])
            chans[n] = ("chw", conf["filters"])
        elif c == "BatchNormalization":
            for p in BN_PARTS:
                shapes[f"{n}/{p}"] = (chans[src[0]][1],)
            chans[n] = chans[src[0]]
        elif c == "Flatten":
            chans[n] = ("flat", chans[src[0]][1] * 64)
        elif c == "Dense":
            shapes[f"{n}/kernel:0"] = (chans[src[0]][1], conf["units"])
            shapes[f"{n}/bias:0"] = (conf["units"],)
            chans[n] = ("flat", conf["units"])
        else:
            chans[n] = chans[src[0]]
    return shapes


# ---------------------------------------------------------------------------------------------------
# JSON emission (ChessModel.build / save)
# ---------------------------------------------------------------------------------------------------
def _init(cls, **kw):
    return {"class_name": cls, "config": kw}


def build_config(stem_k=5, filters=256, blocks=7, block_k=3, value_filters=4, value_fc=256, n_labels=1968,
                 input_depth=18, policy_filters=2, l2=1e-4):
    """The get_config() dict of the graph ChessModel.build creates for these ModelConfig values
    (configs/normal.py:60-68: cnn_first_filter_size, cnn_filter_num, res_layer_num, cnn_filter_size,
    value_fc_size, input_depth, l2_reg)."""
    vs = _init("VarianceScaling", scale=1.0, mode="fan_avg", distribution="uniform", seed=None)
    zeros, ones = _init("Zeros"), _init("Ones")
    reg = _init("L1L2", l1=0.0, l2=l2)

    def node(*srcs):
        return [[[s, 0, 0, {}] for s in srcs]]

    def conv(name, f, k, padding, src):
        return {"name": name, "class_name": "Conv2D", "inbound_nodes": node(src), "config": {
            "name": name, "trainable": True, "filters": f, "kernel_size": [k, k], "strides": [1, 1],
            "padding": padding, "data_format": "channels_first", "dilation_rate": [1, 1], "activation": "linear",
            "use_bias": False, "kernel_initializer": vs, "bias_initializer": zeros, "kernel_regularizer": reg,
            "bias_regularizer": None, "activity_regularizer": None, "kernel_constraint": None,
            "bias_constraint": None}}

    def bn(name, src):
        return {"name": name, "class_name": "BatchNormalization", "inbound_nodes": node(src), "config": {
            "name": name, "trainable": True, "axis": 1, "momentum": 0.99, "epsilon": 0.001, "center": True,
            "scale": True, "beta_initializer": zeros, "gamma_initializer": ones,
            "moving_mean_initializer": zeros, "moving_variance_initializer": ones, "beta_regularizer": None,
            "gamma_regularizer": None, "beta_constraint": None, "gamma_constraint": None}}

    def act(name, fn, src):
        return {"name": name, "class_name": "Activation", "inbound_nodes": node(src),
                "config": {"name": name, "trainable": True, "activation": fn}}

    def flatten(name, src):
        return {"name": name, "class_name": "Flatten", "inbound_nodes": node(src),
                "config": {"name": name, "trainable": True}}

    def dense(name, units, fn, src):
        return {"name": name, "class_name": "Dense", "inbound_nodes": node(src), "config": {
            "name": name, "trainable": True, "units": units, "activation": fn, "use_bias": True,
            "kernel_initializer": vs, "bias_initializer": zeros, "kernel_regularizer": reg,
            "bias_regularizer": None, "activity_regularizer": None, "kernel_constraint": None,
            "bias_constraint": None}}

    layers = [{"name": "input_1", "class_name": "InputLayer", "inbound_nodes": [],
               "config": {"batch_input_shape": [None, input_depth, 8, 8], "dtype": "float32", "sparse": False,
                          "name": "input_1"}}]
    stem = f"input_conv-{stem_k}-{filters}"
    layers += [conv(stem, filters, stem_k, "same", "input_1"), bn("input_batchnorm", stem),
               act("input_relu", "relu", "input_batchnorm")]
    x = "input_relu"
    for i in range(1, blocks + 1):
        r = f"res{i}"
        c1, c2 = f"{r}_conv1-{block_k}-{filters}", f"{r}_conv2-{block_k}-{filters}"
        layers += [conv(c1, filters, block_k, "same", x), bn(f"{r}_batchnorm1", c1),
                   act(f"{r}_relu1", "relu", f"{r}_batchnorm1"),
                   conv(c2, filters, block_k, "same", f"{r}_relu1"), bn(f"{r}_batchnorm2", c2),
                   {"name": f"{r}_add", "class_name": "Add", "inbound_nodes": node(x, f"{r}_batchnorm2"),
                    "config": {"name": f"{r}_add", "trainable": True}},
                   act(f"{r}_relu2", "relu", f"{r}_add")]
        x = f"{r}_relu2"
    vc, pc = f"value_conv-1-{value_filters}", f"policy_conv-1-{policy_filters}"
    layers += [conv(vc, value_filters, 1, "valid", x), conv(pc, policy_filters, 1, "valid", x),
               bn("value_batchnorm", vc), bn("policy_batchnorm", pc),
               act("value_relu", "relu", "value_batchnorm"), act("policy_relu", "relu", "policy_batchnorm"),
               flatten("value_flatten", "value_relu"), flatten("policy_flatten", "policy_relu"),
               dense("value_dense", value_fc, "relu", "value_flatten"),
               dense("policy_out", n_labels, "softmax", "policy_flatten"),
               dense("value_out", 1, "tanh", "value_dense")]
    return {"name": "chess_model", "layers": layers, "input_layers": [["input_1", 0, 0]],
            "output_layers": [["policy_out", 0, 0], ["value_out", 0, 0]]}


def init_weights(cfg, seed=None):
    """Keras default initialisers named in the JSON: glorot-uniform kernels (VarianceScaling fan_avg
    uniform), zero biases, identity BatchNorm -- what ChessModel.build leaves a fresh model with."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in expected_shapes(cfg).items():
        kind = name.rsplit("/", 1)[1]
        if kind == "kernel:0":
            rf = int(np.prod(shape[:-2])) if len(shape) == 4 else 1
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            out[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif kind in ("gamma:0", "moving_variance:0"):
            out[name] = np.ones(shape, np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out


# ---------------------------------------------------------------------------------------------------
# weight files
# ---------------------------------------------------------------------------------------------------
def load_config_file(path):
    with open(path, "rt") as f:
        return json.load(f)


def load_weight_file(path, cfg):
    """Returns {keras weight name: float32 array}.  Accepts
      * ``.npz`` keyed by Keras weight names (this package's native container), and
      * Keras ``save_weights`` / ``model.save`` HDF5 (model_chess.py:164) through :mod:`.hdf5_min`, a
        dependency-free reader of the format subset Keras writes (h5py is not needed)."""
    shapes = expected_shapes(cfg)
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic[:4] == b"PK\x03\x04" or zipfile.is_zipfile(path):
        with np.load(path) as z:
            raw = {k: z[k] for k in z.files}
    elif magic == b"\x89HDF\r\n\x1a\n":
        from . import hdf5_min
        raw = hdf5_min.read_keras_weights(path, keras_layer_order(cfg))
    else:
        raise GraphError(f"{path}: neither an .npz nor an HDF5 weight file")
    out = OrderedDict()
    for name, shape in shapes.items():
        if name not in raw:
            raise GraphError(f"{path}: weight {name} is missing")
        a = np.ascontiguousarray(raw[name], dtype=np.float32)
        if tuple(a.shape) != tuple(shape):
            raise GraphError(f"{path}: weight {name} has shape {a.shape}, expected {shape}")
        out[name] = a
    return out


def save_weight_file(path, weights, cfg=None):
    """``.h5`` / ``.hdf5`` paths get the Keras ``save_weights`` tree (needs the model config for the layer list);
    anything else an ``.npz`` keyed by Keras weight names."""
    tmp = path + ".tmp"
    if path.endswith((".h5", ".hdf5")):
        if cfg is None:
            raise ValueError("writing a Keras HDF5 weight file needs the model config")
        from . import hdf5_min
        layers = []
        for layer in _Graph(cfg).cfg["layers"]:
            prefix = layer["name"] + "/"
            layers.append((layer["name"], [(k, np.asarray(v, np.float32)) for k, v in weights.items() if k.startswith(prefix)]))
        hdf5_min.write_keras_weights(tmp, layers)
    else:
        with open(tmp, "wb") as f:
            np.savez(f, **{k: np.asarray(v, np.float32) for k, v in weights.items()})
    os.replace(tmp, path)
