"""CPU-side ablation behind the bf16 build's operand formats (profiles/r2_precision_ablation.md): which rounding points of the
16-bit tower cost how much of the value head's error on the reference's random-init 7x256 net (256 positions), and how
the error splits between weight rounding and activation rounding.  Uses the oracle only (test infrastructure):
    python tools/precision_ablation.py [keras_default|randomized]"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import encoder as oenc, network as onet, weights as oweights
torch.set_num_threads(os.cpu_count() or 1)
cfg = oweights.keras_config()
w = oweights.synth_weights(cfg, seed=0, recipe=sys.argv[1] if len(sys.argv) > 1 else "keras_default")
fens = oenc.KNOWN_ANSWER_FENS + oenc.synthetic_fens(251, 5)
x0 = np.stack([oenc.canon_input_planes(f) for f in fens[:256]])
wp, wv = onet.forward_graph(cfg, w, x0)
tower, heads = onet.fold_plan(cfg, w)
bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
def run(act_round, w_round):
    # act_round[l], w_round[l]: True => bf16-round operand of layer l
    with torch.no_grad():
        x = torch.from_numpy(x0.astype(np.float32)); skip = None
        for l, (kind, k, b) in enumerate(tower):
            kt = torch.from_numpy(k.astype(np.float32)).permute(3, 2, 0, 1).contiguous()
            if w_round[l]: kt = bf(kt)
            bt = torch.from_numpy(b.astype(np.float32)).view(1, -1, 1, 1)
            xin = bf(x) if act_round[l] else x
            y = F.conv2d(xin, kt, padding=(k.shape[0]-1)//2) + bt
            if kind == "conv1": skip = x
            if kind == "conv2": y = y + skip
            x = torch.relu(y)
        B = x.shape[0]
        feats = x.permute(0, 2, 3, 1).reshape(B*64, -1)
        out = {}
        f32 = lambda a: torch.from_numpy(a.astype(np.float32))
        for head, conv in (("p", "pconv"), ("v", "vconv")):
            k, b = heads[conv]
            h = torch.relu(feats @ f32(k.reshape(k.shape[2], k.shape[3])) + f32(b))
            out[head] = h.reshape(B, 64, -1).permute(0, 2, 1).reshape(B, -1)
        logits = out["p"] @ f32(heads["pfc"][0]) + f32(heads["pfc"][1])
        policy = torch.softmax(logits, dim=-1)
        h = torch.relu(out["v"] @ f32(heads["vfc1"][0]) + f32(heads["vfc1"][1]))
        value = torch.tanh(h @ f32(heads["vfc2"][0]) + f32(heads["vfc2"][1]))
    return policy.numpy(), value.numpy()
def rep(tag, a, ww):
    p, v = run(a, ww)
    dv = np.abs(v - wv).reshape(-1)
    print(f"{tag:40s} max|dv|={dv.max():.3e} rms={np.sqrt((dv**2).mean()):.3e} p99={np.sort(dv)[-3]:.3e} max|dp|={np.abs(p-wp).max():.3e}", flush=True)
L = len(tower)
T, Fa = [True]*L, [False]*L
print("value range", wv.min(), wv.max(), "mean|v|", np.abs(wv).mean())
rep("none rounded", Fa, Fa)
rep("all rounded", T, T)
rep("only weights rounded", Fa, T)
rep("only acts rounded", T, Fa)
for l in range(L):
    a = list(Fa); a[l] = True
    rep(f"only act layer {l}", a, Fa)
for l in range(L):
    a = list(Fa); a[l] = True
    rep(f"only w layer {l}", Fa, a)

# ---- operand formats, several summation orders (a 2e-7 relative jitter of every accumulator stands in for a different order) ----
f16 = lambda t: t.clamp(-65504, 65504).to(torch.float16).to(torch.float32)


def run_formats(fa, fw, jitter, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        x = torch.from_numpy(x0.astype(np.float32)); skip = None
        for l, (kind, k, b) in enumerate(tower):
            kt = fw(torch.from_numpy(k.astype(np.float32)).permute(3, 2, 0, 1).contiguous())
            y = F.conv2d(fa(x), kt, padding=(k.shape[0] - 1) // 2)
            if jitter:
                y = y * (1 + jitter * (torch.rand(y.shape, generator=g) - 0.5))
            y = y + torch.from_numpy(b.astype(np.float32)).view(1, -1, 1, 1)
            if kind == "conv1": skip = x
            if kind == "conv2": y = y + skip
            x = torch.relu(y)
        B = x.shape[0]
        feats = x.permute(0, 2, 3, 1).reshape(B * 64, -1)
        f32 = lambda a: torch.from_numpy(a.astype(np.float32))
        k, b = heads["vconv"]
        h = torch.relu(feats @ f32(k.reshape(k.shape[2], k.shape[3])) + f32(b)).reshape(B, 64, -1).permute(0, 2, 1).reshape(B, -1)
        h = torch.relu(h @ f32(heads["vfc1"][0]) + f32(heads["vfc1"][1]))
        return torch.tanh(h @ f32(heads["vfc2"][0]) + f32(heads["vfc2"][1])).numpy()


for tag, fa, fw in (("A bf16 x W bf16", bf, bf), ("A fp16 x W bf16 (the bf16 build)", f16, bf), ("A bf16 x W fp16", bf, f16), ("A fp16 x W fp16", f16, f16)):
    mx = []
    for s in range(6):
        dv = np.abs(run_formats(fa, fw, 2e-7 if s else 0.0, s) - wv).reshape(-1)
        mx.append(dv.max())
    print(f"{tag:34s} max|dv| over 6 summation orders: {' '.join('%.2e' % m for m in mx)}", flush=True)
