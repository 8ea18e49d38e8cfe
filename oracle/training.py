"""Oracle: one training step of the reference's `opt` worker on the CPU (torch autograd, float64 by default).

Test infrastructure (see oracle/__init__.py).  PARITY UNPINNED against Keras/TF (not installable here) -- restates
  compile_model / train_epoch   /root/reference/src/chess_zero/worker/optimize.py:79-103
      Adam() with keras 2.0.8 defaults, losses ['categorical_crossentropy', 'mean_squared_error'],
      loss_weights [1.25, 1.0] (configs/normal.py:57), batch_size 384
  the graph                     /root/reference/src/chess_zero/agent/model_chess.py:57-111
      kernel_regularizer l2(1e-4) on every Conv2D and Dense kernel (not on biases or BatchNormalization),
      BatchNormalization(axis=1) in TRAINING mode: batch mean / biased variance over (N, H, W), epsilon from the JSON,
      moving averages  moving = moving * 0.99 + batch * 0.01
from the published Keras semantics (TensorFlow backend):
  categorical_crossentropy(y_true, y_pred): y_pred /= sum(y_pred); clip to [1e-7, 1 - 1e-7]; -sum(y_true * log(y_pred)); mean over the batch
  mean_squared_error: mean over the last axis of (y_pred - y_true)^2; mean over the batch
  total loss = 1.25 * CE + 1.0 * MSE + sum over regularised kernels of 1e-4 * sum(w^2)
  Adam: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; p -= lr_t * m / (sqrt(v) + eps)
The forward graph is interpreted from the Keras JSON exactly like oracle/network.py:forward_graph.
"""
import numpy as np
import torch
import torch.nn.functional as F

ADAM = dict(learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-8)
L2 = 1e-4
LOSS_WEIGHTS = (1.25, 1.0)
BN_MOMENTUM = 0.99


def _inbound(layer):
    return [n[0] for n in layer["inbound_nodes"][0]] if layer["inbound_nodes"] else []


def trainable_names(cfg):
    """(regularised kernels, other trainables, non-trainable moving statistics) as Keras weight names."""
    kernels, others, moving = [], [], []
    for layer in cfg["layers"]:
        n, cls = layer["name"], layer["class_name"]
        if cls == "Conv2D":
            kernels.append(f"{n}/kernel:0")
        elif cls == "Dense":
            kernels.append(f"{n}/kernel:0")
            others.append(f"{n}/bias:0")
        elif cls == "BatchNormalization":
            others += [f"{n}/gamma:0", f"{n}/beta:0"]
            moving += [f"{n}/moving_mean:0", f"{n}/moving_variance:0"]
    return kernels, others, moving


def forward_train(cfg, w, planes):
    """Training-mode forward pass on torch tensors `w` (requires_grad where trainable).  Returns (policy probabilities,
    value, {moving statistic name: batch statistic}) -- BatchNormalization uses the batch statistics."""
    t, batch_stats = {}, {}
    for layer in cfg["layers"]:
        name, cls, c = layer["name"], layer["class_name"], layer["config"]
        src = [t[i] for i in _inbound(layer)]
        if cls == "InputLayer":
            t[name] = planes
        elif cls == "Conv2D":
            k = w[f"{name}/kernel:0"]
            pad = ((k.shape[0] - 1) // 2, (k.shape[1] - 1) // 2) if c["padding"] == "same" else (0, 0)
            t[name] = F.conv2d(src[0], k.permute(3, 2, 0, 1), padding=pad)
        elif cls == "BatchNormalization":
            x = src[0]
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            batch_stats[f"{name}/moving_mean:0"] = mean.detach()
            batch_stats[f"{name}/moving_variance:0"] = var.detach()
            xhat = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + c["epsilon"])
            t[name] = xhat * w[f"{name}/gamma:0"].view(1, -1, 1, 1) + w[f"{name}/beta:0"].view(1, -1, 1, 1)
        elif cls == "Activation":
            t[name] = torch.relu(src[0]) if c["activation"] == "relu" else src[0]
        elif cls == "Add":
            t[name] = src[0] + src[1]
        elif cls == "Flatten":
            t[name] = src[0].reshape(src[0].shape[0], -1)
        elif cls == "Dense":
            y = src[0] @ w[f"{name}/kernel:0"] + w[f"{name}/bias:0"]
            act = c["activation"]
            t[name] = torch.relu(y) if act == "relu" else torch.tanh(y) if act == "tanh" else torch.softmax(y, dim=-1) if act == "softmax" else y
        else:
            raise ValueError(cls)
    outs = [t[o[0]] for o in cfg["output_layers"]]
    return outs[0], outs[1], batch_stats


def losses(cfg, w, policy, value, policy_t, value_t, l2=L2, loss_weights=LOSS_WEIGHTS):
    p = policy / policy.sum(dim=-1, keepdim=True)
    p = p.clamp(1e-7, 1 - 1e-7)
    ce = -(policy_t * torch.log(p)).sum(dim=-1).mean()
    mse = ((value - value_t.view(-1, 1)) ** 2).mean(dim=-1).mean()
    kernels, _, _ = trainable_names(cfg)
    reg = sum(l2 * (w[k] ** 2).sum() for k in kernels)
    return loss_weights[0] * ce + loss_weights[1] * mse + reg, ce, mse, reg


class OracleTrainer:
    """State (weights, Adam moments, step) + train_on_batch, float64 unless told otherwise."""

    def __init__(self, cfg, weights, dtype=torch.float64, adam=None, l2=L2, loss_weights=LOSS_WEIGHTS, bn_momentum=BN_MOMENTUM):
        self.cfg, self.dtype = cfg, dtype
        self.adam = dict(ADAM, **(adam or {}))
        self.l2, self.loss_weights, self.bn_momentum = l2, loss_weights, bn_momentum
        self.kernels, self.others, self.moving = trainable_names(cfg)
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in weights.items()}
        self.m = {k: torch.zeros_like(self.w[k]) for k in self.kernels + self.others}
        self.v = {k: torch.zeros_like(self.w[k]) for k in self.kernels + self.others}
        self.step = 0

    def gradients(self, planes, policy_t, value_t):
        """-> (losses dict, {name: d total_loss / d weight} WITH the l2 term, data-loss gradients without it, outputs)."""
        w = {k: (v.clone().requires_grad_(True) if k in self.m else v) for k, v in self.w.items()}
        x = torch.from_numpy(np.ascontiguousarray(planes)).to(self.dtype)
        pt = torch.from_numpy(np.ascontiguousarray(policy_t)).to(self.dtype)
        vt = torch.from_numpy(np.ascontiguousarray(value_t)).to(self.dtype).reshape(-1)
        policy, value, stats = forward_train(self.cfg, w, x)
        total, ce, mse, reg = losses(self.cfg, w, policy, value, pt, vt, self.l2, self.loss_weights)
        names = list(self.m)
        grads = torch.autograd.grad(total, [w[k] for k in names])
        g = dict(zip(names, grads))
        data = {k: (g[k] - 2 * self.l2 * self.w[k] if k in self.kernels else g[k]) for k in names}
        out = {"total": float(total.detach()), "ce": float(ce.detach()), "mse": float(mse.detach()), "l2": float(reg.detach())}
        return out, g, data, (policy.detach().numpy(), value.detach().numpy()), stats

    def train_on_batch(self, planes, policy_t, value_t):
        out, g, _, _, stats = self.gradients(planes, policy_t, value_t)
        self.step += 1
        a = self.adam
        lr_t = a["learning_rate"] * np.sqrt(1 - a["beta_2"] ** self.step) / (1 - a["beta_1"] ** self.step)
        for k in self.m:
            self.m[k] = a["beta_1"] * self.m[k] + (1 - a["beta_1"]) * g[k]
            self.v[k] = a["beta_2"] * self.v[k] + (1 - a["beta_2"]) * g[k] ** 2
            self.w[k] = self.w[k] - lr_t * self.m[k] / (torch.sqrt(self.v[k]) + a["epsilon"])
        for k, s in stats.items():
            self.w[k] = self.w[k] * self.bn_momentum + s * (1 - self.bn_momentum)
        return out

    def weights(self):
        return {k: v.to(torch.float32).numpy() for k, v in self.w.items()}
