"""Oracle: Keras-schema model configs and seeded synthetic weights.

Test infrastructure (see oracle/__init__.py).

``keras_config`` emits the dict ``keras.Model.get_config()`` would produce for the graph
built by ChessModel.build (/root/reference/src/chess_zero/agent/model_chess.py:57-111);
field names and layer order follow the only real sample the reference ships,
/root/reference/data/model/model_best_config.json (which is the older 3x3-stem,
1-filter-value variant: ``keras_config(stem_k=3, value_filters=1)`` reproduces its
layer list name for name -- checked in tests/test_oracle_network.py).

``synth_weights`` draws a weight set in Keras naming (``<layer>/kernel:0`` ...) with the
initialisers that JSON spells out: VarianceScaling(scale=1, fan_avg, uniform) = glorot
uniform for kernels, zeros for biases, BN gamma=1 beta=0 mean=0 var=1 (recipe
"keras_default"), or a second recipe "randomized"  that randomises the BN
statistics and biases so that BN folding and bias paths are actually exercised and the
value head does not saturate.
"""
from collections import OrderedDict

import numpy as np

_VS = {"class_name": "VarianceScaling",
       "config": {"scale": 1.0, "mode": "fan_avg", "distribution": "uniform", "seed": None}}
_ZEROS = {"class_name": "Zeros", "config": {}}
_ONES = {"class_name": "Ones", "config": {}}
_L2 = {"class_name": "L1L2", "config": {"l1": 0.0, "l2": 0.0001}}


def _conv(name, filters, k, padding, inbound):
    return {"name": name, "class_name": "Conv2D", "config": {
        "name": name, "trainable": True, "filters": filters, "kernel_size": [k, k], "strides": [1, 1],
        "padding": padding, "data_format": "channels_first", "dilation_rate": [1, 1],
        "activation": "linear", "use_bias": False, "kernel_initializer": _VS, "bias_initializer": _ZEROS,
        "kernel_regularizer": _L2, "bias_regularizer": None, "activity_regularizer": None,
        "kernel_constraint": None, "bias_constraint": None},
        "inbound_nodes": [[[inbound, 0, 0, {}]]]}


def _bn(name, inbound):
    return {"name": name, "class_name": "BatchNormalization", "config": {
        "name": name, "trainable": True, "axis": 1, "momentum": 0.99, "epsilon": 0.001, "center": True,
        "scale": True, "beta_initializer": _ZEROS, "gamma_initializer": _ONES,
        "moving_mean_initializer": _ZEROS, "moving_variance_initializer": _ONES,
        "beta_regularizer": None, "gamma_regularizer": None, "beta_constraint": None,
        "gamma_constraint": None}, "inbound_nodes": [[[inbound, 0, 0, {}]]]}


def _act(name, fn, inbound):
    return {"name": name, "class_name": "Activation",
            "config": {"name": name, "trainable": True, "activation": fn},
            "inbound_nodes": [[[inbound, 0, 0, {}]]]}


def _dense(name, units, fn, inbound):
    return {"name": name, "class_name": "Dense", "config": {
        "name": name, "trainable": True, "units": units, "activation": fn, "use_bias": True,
        "kernel_initializer": _VS, "bias_initializer": _ZEROS, "kernel_regularizer": _L2,
        "bias_regularizer": None, "activity_regularizer": None, "kernel_constraint": None,
        "bias_constraint": None}, "inbound_nodes": [[[inbound, 0, 0, {}]]]}


def keras_config(stem_k=5, filters=256, blocks=7, block_k=3, value_filters=4, value_fc=256,
                 n_labels=1968, input_depth=18):
    L = [{"name": "input_1", "class_name": "InputLayer",
          "config": {"batch_input_shape": [None, input_depth, 8, 8], "dtype": "float32",
                     "sparse": False, "name": "input_1"}, "inbound_nodes": []}]
    stem = f"input_conv-{stem_k}-{filters}"
    L += [_conv(stem, filters, stem_k, "same", "input_1"), _bn("input_batchnorm", stem),
          _act("input_relu", "relu", "input_batchnorm")]
    x = "input_relu"
    for i in range(1, blocks + 1):
        r = f"res{i}"
        c1, c2 = f"{r}_conv1-{block_k}-{filters}", f"{r}_conv2-{block_k}-{filters}"
        L += [_conv(c1, filters, block_k, "same", x), _bn(f"{r}_batchnorm1", c1),
              _act(f"{r}_relu1", "relu", f"{r}_batchnorm1"),
              _conv(c2, filters, block_k, "same", f"{r}_relu1"), _bn(f"{r}_batchnorm2", c2),
              {"name": f"{r}_add", "class_name": "Add", "config": {"name": f"{r}_add", "trainable": True},
               "inbound_nodes": [[[x, 0, 0, {}], [f"{r}_batchnorm2", 0, 0, {}]]]},
              _act(f"{r}_relu2", "relu", f"{r}_add")]
        x = f"{r}_relu2"
    vc, pc = f"value_conv-1-{value_filters}", "policy_conv-1-2"
    L += [_conv(vc, value_filters, 1, "valid", x), _conv(pc, 2, 1, "valid", x),
          _bn("value_batchnorm", vc), _bn("policy_batchnorm", pc),
          _act("value_relu", "relu", "value_batchnorm"), _act("policy_relu", "relu", "policy_batchnorm"),
          {"name": "value_flatten", "class_name": "Flatten",
           "config": {"name": "value_flatten", "trainable": True},
           "inbound_nodes": [[["value_relu", 0, 0, {}]]]},
          {"name": "policy_flatten", "class_name": "Flatten",
           "config": {"name": "policy_flatten", "trainable": True},
           "inbound_nodes": [[["policy_relu", 0, 0, {}]]]},
          _dense("value_dense", value_fc, "relu", "value_flatten"),
          _dense("policy_out", n_labels, "softmax", "policy_flatten"),
          _dense("value_out", 1, "tanh", "value_dense")]
    return {"name": "chess_model", "layers": L, "input_layers": [["input_1", 0, 0]],
            "output_layers": [["policy_out", 0, 0], ["value_out", 0, 0]]}


def _shapes(cfg):
    """Walk the layer list and yield (layer, in_channels_or_features)."""
    chans = {}
    for layer in cfg["layers"]:
        name, cls, c = layer["name"], layer["class_name"], layer["config"]
        inbound = [n[0] for n in layer["inbound_nodes"][0]] if layer["inbound_nodes"] else []
        if cls == "InputLayer":
            chans[name] = ("chw", c["batch_input_shape"][1])
        elif cls == "Conv2D":
            yield layer, chans[inbound[0]][1]
            chans[name] = ("chw", c["filters"])
        elif cls == "Dense":
            yield layer, chans[inbound[0]][1]
            chans[name] = ("flat", c["units"])
        elif cls == "Flatten":
            chans[name] = ("flat", chans[inbound[0]][1] * 64)
        else:
            if cls == "BatchNormalization":
                yield layer, chans[inbound[0]][1]
            chans[name] = chans[inbound[0]]


def synth_weights(cfg, seed=0, recipe="keras_default"):
    rng = np.random.RandomState(seed)
    W = OrderedDict()
    for layer, cin in _shapes(cfg):
        name, cls, c = layer["name"], layer["class_name"], layer["config"]
        if cls == "Conv2D":
            k, co = c["kernel_size"][0], c["filters"]
            lim = np.sqrt(6.0 / (k * k * cin + k * k * co))
            W[f"{name}/kernel:0"] = rng.uniform(-lim, lim, size=(k, k, cin, co)).astype(np.float32)
        elif cls == "Dense":
            u = c["units"]
            lim = np.sqrt(6.0 / (cin + u))
            W[f"{name}/kernel:0"] = rng.uniform(-lim, lim, size=(cin, u)).astype(np.float32)
            if recipe == "randomized":
                W[f"{name}/bias:0"] = rng.normal(0, 0.1, size=(u,)).astype(np.float32)
            else:
                W[f"{name}/bias:0"] = np.zeros((u,), np.float32)
        elif cls == "BatchNormalization":
            if recipe == "randomized":
                W[f"{name}/gamma:0"] = rng.uniform(0.75, 1.25, size=(cin,)).astype(np.float32)
                W[f"{name}/beta:0"] = rng.normal(0, 0.1, size=(cin,)).astype(np.float32)
                W[f"{name}/moving_mean:0"] = rng.normal(0, 0.1, size=(cin,)).astype(np.float32)
                W[f"{name}/moving_variance:0"] = rng.uniform(0.75, 1.25, size=(cin,)).astype(np.float32)
            else:
                W[f"{name}/gamma:0"] = np.ones((cin,), np.float32)
                W[f"{name}/beta:0"] = np.zeros((cin,), np.float32)
                W[f"{name}/moving_mean:0"] = np.zeros((cin,), np.float32)
                W[f"{name}/moving_variance:0"] = np.ones((cin,), np.float32)
    if recipe == "randomized":
        # keep tanh away from saturation and the policy away from uniform: rescale the two output layers
        W["value_out/kernel:0"] *= np.float32(0.25)
        W["policy_out/kernel:0"] *= np.float32(2.0)
        # Tower activations are post-ReLU (all >= 0), so a random 1x1 head filter has an output whose sign is
        # nearly the same on every square and position; when it is negative the head's ReLU kills the feature
        # and the policy degenerates to softmax(bias).  Zero-mean head filters keep about half the features alive.
        for name in list(W):
            if name.startswith(("policy_conv", "value_conv")) and name.endswith("/kernel:0"):
                k = W[name]
                W[name] = (k - k.mean(axis=2, keepdims=True)).astype(np.float32)
    return W
