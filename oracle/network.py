"""Oracle: CPU forward pass of the reference's policy/value network.

Test infrastructure (see oracle/__init__.py).  PARITY UNPINNED against Keras/TF (not
installable here; see oracle/__init__.py) -- restates, from the published Keras semantics,
  ChessModel.build            /root/reference/src/chess_zero/agent/model_chess.py:57-94
  _build_residual_block       /root/reference/src/chess_zero/agent/model_chess.py:96-111
  predict_on_batch call site  /root/reference/src/chess_zero/agent/api_chess.py:66-67
with every hyper-parameter read from the Keras JSON (kernel sizes, filters, padding,
data_format channels_first, use_bias, BN axis/epsilon, Dense units/activations), as
Model.from_config would (model_chess.py:144-146).

Three formulations:
  forward_graph   -- interprets the JSON layer list node by node on NCHW tensors with
                     un-folded BatchNormalization (torch CPU, fp32 or fp64).  THE oracle.
  forward_folded  -- independent numpy statement: BN folded into conv weights, NHWC,
                     explicit im2col GEMM.  Must agree with forward_graph to <= 1e-5.
  forward_bf16    -- forward_graph's arithmetic with the bf16 build's rounding points
                     (folded conv weights and inter-layer activations rounded to bf16,
                     fp32 accumulation, fp32 heads) -- a tight checker for the tensor-core
                     path, not a parity target.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _inbound(layer):
    return [n[0] for n in layer["inbound_nodes"][0]] if layer["inbound_nodes"] else []


def _activation(x, fn):
    if fn == "linear":
        return x
    if fn == "relu":
        return torch.relu(x)
    if fn == "tanh":
        return torch.tanh(x)
    if fn == "softmax":
        return torch.softmax(x, dim=-1)
    raise ValueError(f"activation {fn!r} not used by the reference graph")


def forward_graph(cfg, weights, planes, dtype=torch.float32, threads=None):
    """Returns [policy (B, n_labels), value (B, 1)] as float32 numpy, like predict_on_batch."""
    if threads:
        torch.set_num_threads(threads)
    t = {}
    w = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in weights.items()}
    with torch.no_grad():
        for layer in cfg["layers"]:
            name, cls, c = layer["name"], layer["class_name"], layer["config"]
            src = [t[i] for i in _inbound(layer)]
            if cls == "InputLayer":
                t[name] = torch.from_numpy(np.ascontiguousarray(planes)).to(dtype)
            elif cls == "Conv2D":
                assert c["data_format"] == "channels_first" and not c["use_bias"]
                assert c["strides"] == [1, 1] and c["dilation_rate"] == [1, 1]
                k = w[f"{name}/kernel:0"]                      # (kh, kw, Cin, Cout): Keras HWIO
                kh, kw = k.shape[0], k.shape[1]
                pad = ((kh - 1) // 2, (kw - 1) // 2) if c["padding"] == "same" else (0, 0)
                assert c["padding"] in ("same", "valid") and kh % 2 == 1 and kw % 2 == 1
                y = F.conv2d(src[0], k.permute(3, 2, 0, 1).contiguous(), padding=pad)  # cross-correlation
                t[name] = _activation(y, c["activation"])
            elif cls == "BatchNormalization":
                assert c["axis"] == 1
                g, b = w[f"{name}/gamma:0"], w[f"{name}/beta:0"]
                m, v = w[f"{name}/moving_mean:0"], w[f"{name}/moving_variance:0"]
                inv = g * torch.rsqrt(v + c["epsilon"])
                t[name] = src[0] * inv.view(1, -1, 1, 1) + (b - m * inv).view(1, -1, 1, 1)
            elif cls == "Activation":
                t[name] = _activation(src[0], c["activation"])
            elif cls == "Add":
                t[name] = src[0] + src[1]
            elif cls == "Flatten":
                t[name] = src[0].reshape(src[0].shape[0], -1)   # NCHW: index c*64 + r*8 + f
            elif cls == "Dense":
                y = src[0] @ w[f"{name}/kernel:0"] + w[f"{name}/bias:0"]
                t[name] = _activation(y, c["activation"])
            else:
                raise ValueError(f"layer class {cls!r} not used by the reference graph")
        outs = [t[o[0]] for o in cfg["output_layers"]]
    return [o.to(torch.float32).numpy() for o in outs]


# ---------------------------------------------------------------------------
# Structured plan (stem / blocks / heads) with BN folded -- independent of the interpreter
# ---------------------------------------------------------------------------
def fold_plan(cfg, weights):
    """Pattern-match the layer list into conv(+BN)(+add)+relu steps and the two heads."""
    layers = {l["name"]: l for l in cfg["layers"]}
    eps = {n: l["config"]["epsilon"] for n, l in layers.items() if l["class_name"] == "BatchNormalization"}

    def folded(conv_name, bn_name):
        k = weights[f"{conv_name}/kernel:0"].astype(np.float64)
        inv = weights[f"{bn_name}/gamma:0"].astype(np.float64) / np.sqrt(
            weights[f"{bn_name}/moving_variance:0"].astype(np.float64) + eps[bn_name])
        bias = weights[f"{bn_name}/beta:0"].astype(np.float64) - weights[f"{bn_name}/moving_mean:0"].astype(np.float64) * inv
        return k * inv, bias

    names = [l["name"] for l in cfg["layers"]]
    stem = [n for n in names if n.startswith("input_conv")][0]
    tower = [("stem", *folded(stem, "input_batchnorm"))]
    i = 1
    while f"res{i}_add" in layers:
        c1 = [n for n in names if n.startswith(f"res{i}_conv1")][0]
        c2 = [n for n in names if n.startswith(f"res{i}_conv2")][0]
        tower.append(("conv1", *folded(c1, f"res{i}_batchnorm1")))
        tower.append(("conv2", *folded(c2, f"res{i}_batchnorm2")))
        i += 1
    pconv = [n for n in names if n.startswith("policy_conv")][0]
    vconv = [n for n in names if n.startswith("value_conv")][0]
    heads = {
        "pconv": folded(pconv, "policy_batchnorm"), "vconv": folded(vconv, "value_batchnorm"),
        "pfc": (weights["policy_out/kernel:0"].astype(np.float64), weights["policy_out/bias:0"].astype(np.float64)),
        "vfc1": (weights["value_dense/kernel:0"].astype(np.float64), weights["value_dense/bias:0"].astype(np.float64)),
        "vfc2": (weights["value_out/kernel:0"].astype(np.float64), weights["value_out/bias:0"].astype(np.float64)),
    }
    return tower, heads


def _im2col_nhwc(x, k):
    """x: (B, 8, 8, C) -> (B*64, k*k*C) with zero 'same' padding, tap-major then channel."""
    B, H, Wd, C = x.shape
    p = (k - 1) // 2
    xp = np.zeros((B, H + 2 * p, Wd + 2 * p, C), dtype=x.dtype)
    xp[:, p:p + H, p:p + Wd] = x
    cols = [xp[:, dy:dy + H, dx:dx + Wd] for dy in range(k) for dx in range(k)]
    return np.concatenate(cols, axis=-1).reshape(B * H * Wd, k * k * C)


def forward_folded(cfg, weights, planes, dtype=np.float64):
    tower, heads = fold_plan(cfg, weights)
    x = np.transpose(np.asarray(planes, dtype=dtype), (0, 2, 3, 1))      # NHWC
    B = x.shape[0]
    skip = None
    for kind, k, b in tower:
        kh = k.shape[0]
        y = _im2col_nhwc(x, kh) @ k.reshape(-1, k.shape[3]).astype(dtype) + b.astype(dtype)
        y = y.reshape(B, 8, 8, -1)
        if kind == "conv1":
            skip = x
        if kind == "conv2":
            y = y + skip
        x = np.maximum(y, 0)
    feats = x.reshape(B * 64, -1)
    out = {}
    for head, conv in (("p", "pconv"), ("v", "vconv")):
        k, b = heads[conv]
        h = np.maximum(feats @ k.reshape(k.shape[2], k.shape[3]).astype(dtype) + b.astype(dtype), 0)
        out[head] = np.transpose(h.reshape(B, 64, -1), (0, 2, 1)).reshape(B, -1)   # Flatten on NCHW: c*64 + s
    logits = out["p"] @ heads["pfc"][0].astype(dtype) + heads["pfc"][1].astype(dtype)
    logits = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(logits)
    policy = e / e.sum(axis=1, keepdims=True)
    h = np.maximum(out["v"] @ heads["vfc1"][0].astype(dtype) + heads["vfc1"][1].astype(dtype), 0)
    value = np.tanh(h @ heads["vfc2"][0].astype(dtype) + heads["vfc2"][1].astype(dtype))
    return [policy.astype(np.float32), value.astype(np.float32)]


def _bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _fp16_round(x):
    return x.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32)   # saturating, like the kernels' stores


ROUNDERS = {"bf16": _bf16_round, "fp16": _fp16_round, "fp32": lambda x: x}


def forward_bf16(cfg, weights, planes, fp32_residual=False, return_logits=False, fp32_head_input=False,
                 act_format="bf16", w_format="bf16"):
    """16-bit-operand / fp32-accumulate emulation of the tensor-core builds (see module docstring): conv weights are
    rounded to `w_format`, the operand copies of the activations to `act_format` (the library's bf16 build is
    w_format="bf16", act_format="fp16"; its BF16_PURE build "bf16"/"bf16"; its FP16 build "fp16"/"fp16")."""
    tower, heads = fold_plan(cfg, weights)
    _act, _w = ROUNDERS[act_format], ROUNDERS[w_format]
    with torch.no_grad():
        x = _act(torch.from_numpy(np.ascontiguousarray(planes, dtype=np.float32)))   # NCHW
        skip = None
        for kind, k, b in tower:
            kt = _w(torch.from_numpy(k.astype(np.float32)).permute(3, 2, 0, 1).contiguous())
            bt = torch.from_numpy(b.astype(np.float32)).view(1, -1, 1, 1)
            p = (k.shape[0] - 1) // 2
            y = F.conv2d(_act(x), kt, padding=p) + bt
            if kind == "conv1":
                skip = x
            if kind == "conv2":
                y = y + skip
            x = torch.relu(y)
            if not fp32_residual or kind == "conv1":
                x = _act(x)   # with an fp32 residual stream only conv1's output is ever stored as a 16-bit operand
        if not fp32_head_input:
            x = _act(x)
        B = x.shape[0]
        feats = x.permute(0, 2, 3, 1).reshape(B * 64, -1)
        out = {}
        for head, conv in (("p", "pconv"), ("v", "vconv")):
            k, b = heads[conv]
            kk = torch.from_numpy(k.reshape(k.shape[2], k.shape[3]).astype(np.float32))
            h = torch.relu(feats @ kk + torch.from_numpy(b.astype(np.float32)))
            out[head] = h.reshape(B, 64, -1).permute(0, 2, 1).reshape(B, -1)
        f32 = lambda a: torch.from_numpy(a.astype(np.float32))
        logits = out["p"] @ f32(heads["pfc"][0]) + f32(heads["pfc"][1])
        policy = torch.softmax(logits, dim=-1)
        h = torch.relu(out["v"] @ f32(heads["vfc1"][0]) + f32(heads["vfc1"][1]))
        value = torch.tanh(h @ f32(heads["vfc2"][0]) + f32(heads["vfc2"][1]))
    if return_logits:
        return [policy.numpy(), value.numpy(), logits.numpy()]
    return [policy.numpy(), value.numpy()]


def legal_masked_argmax(policy, masks):
    """argmax of each policy row restricted to mask != 0 (how player_chess.py:269-275 consumes it)."""
    p = np.where(masks != 0, policy, -1.0)
    return p.argmax(axis=1)


def decisive_rows(policy, masks, margin):
    """Rows whose best legal move leads the second best by more than `margin`: where an argmax can be expected to survive
    an error of margin / 2 per probability."""
    p = np.where(masks != 0, policy, -1.0)
    top2 = np.sort(p, axis=1)[:, -2:]
    return (top2[:, 1] - top2[:, 0]) > margin


def synthetic_legal_masks(n, n_labels=1968, seed=0):
    rng = np.random.RandomState(seed)
    masks = np.zeros((n, n_labels), dtype=np.uint8)
    for i in range(n):
        k = int(rng.randint(5, 61))
        masks[i, rng.permutation(n_labels)[:k]] = 1
    return masks
