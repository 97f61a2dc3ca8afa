"""Design probe (CPU): which operand precision can each contraction of the hot path
tolerate under the north-star tolerance (1e-3 tensor-core mode / 1e-5 fp32 mode)?

Emulates tensor-core operand rounding (fp32 accumulate) stage by stage on the oracle's
restatement and reports rel-L2 / rel-max error of pred_cls / pred_loc against the fp64 run.
Uses oracle/ as a checker only; not part of the product path.
"""
import sys, os, itertools
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dsnet_oracle as orc

def rnd(t, mode):
    if mode == "fp32": return t
    if mode == "fp16": return t.half().float()
    if mode == "bf16": return t.bfloat16().float()
    if mode == "tf32":   # round-to-nearest 10-bit mantissa
        i = t.contiguous().view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32)
    if mode == "tf32t":  # truncation (what the MMA datapath does to raw fp32 bits)
        i = t.contiguous().view(torch.int32) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(mode)

def mm(a, b, mode):
    """a @ b with emulated operand precision; split modes use hi+lo decomposition."""
    if mode.endswith("x3"):
        base = mode[:-2]
        ah, bh = rnd(a, base), rnd(b, base)
        al, bl = rnd(a - ah, base), rnd(b - bh, base)
        return ah @ bh + (ah @ bl + al @ bh)
    if mode.endswith("x2w"):     # only B (the weight) split: A_hi . [B_hi | B_lo], ONE N = 256 tcgen05 instruction
        base = mode[:-3]
        ah, bh = rnd(a, base), rnd(b, base)
        bl = rnd(b - bh, base)
        return ah @ bh + ah @ bl
    if mode.endswith("x2"):      # only A split (B single)
        base = mode[:-2]
        ah, bh = rnd(a, base), rnd(b, base)
        al = rnd(a - ah, base)
        return ah @ bh + al @ bh
    return rnd(a, mode) @ rnd(b, mode)

def forward(x, p, scales, depth, P):
    T, F = x.shape; h, d, m = 8, 64, 64
    pad = (m - T % m) % m; n = T + pad; seg = n // m
    xp = torch.cat([x.new_zeros(pad, F), x]) if pad else x
    qkv = mm(xp, p["base_model.to_qkv.weight"].t(), P["qkv"])
    q, k, v = (t.reshape(n, h, d).permute(1, 0, 2) for t in qkv.chunk(3, -1))
    q = q * 0.125
    ql = q.reshape(h, m, seg, d).sum(2) / seg; kl = k.reshape(h, m, seg, d).sum(2) / seg
    a1 = torch.softmax(mm(q, kl.transpose(1, 2), P["sim"]), -1)
    a2 = torch.softmax(mm(ql, kl.transpose(1, 2), P["sim2"]), -1)
    a3 = torch.softmax(mm(ql, k.transpose(1, 2), P["sim"]), -1)
    mag = a2.abs(); z = a2.transpose(-1, -2) / (mag.sum(-1).max() * mag.sum(-2).max())
    eye = torch.eye(m)
    for _ in range(6):
        az = mm(a2, z, P["pinv"])
        t1 = 7 * eye - az
        t2 = 15 * eye - mm(az, t1, P["pinv"])
        t3 = 13 * eye - mm(az, t2, P["pinv"])
        z = 0.25 * mm(z, t3, P["pinv"])
    a3v = mm(a3, v, P["agg"])
    if P.get("assoc", False):
        o = mm(a1, mm(z, a3v, P["agg"]), P["agg"])
    else:
        o = mm(mm(a1, z, P["agg"]), a3v, P["agg"])
    w = p["base_model.res_conv.weight"].reshape(h, 1, 33, 1)
    o = o + torch.nn.functional.conv2d(v.unsqueeze(0), w, padding=(16, 0), groups=h)[0]
    y = mm(o.permute(1, 0, 2).reshape(n, h * d), p["base_model.to_out.0.weight"].t(), P["out"]) + p["base_model.to_out.0.bias"]
    y = y[pad:] + x
    u = orc.layer_norm(y, p["layer_norm.weight"], p["layer_norm.bias"])
    u = mm(u, p["fc1.weight"].t(), P["fc1"]) + p["fc1.bias"]
    for _ in range(depth):
        u = torch.relu(mm(u, p["fc_block.0.weight"].t(), P["fcb"]) + p["fc_block.0.bias"])
        u = orc.layer_norm(u, p["fc_block.3.weight"], p["fc_block.3.bias"])
    pooled = orc.roi_pool_direct(u, scales)
    cls = torch.sigmoid(pooled @ p["fc_cls.0.weight"].t() + p["fc_cls.0.bias"])
    loc = pooled @ p["fc_loc.0.weight"].t() + p["fc_loc.0.bias"]
    return cls.reshape(T, -1), loc.reshape(T, -1, 2)

def plan(**kw):
    base = dict(qkv="fp32", out="fp32", fc1="fp32", fcb="fp32", sim="fp32", sim2="fp32", pinv="fp32", agg="fp32")
    base.update(kw); return base

PLANS = {
  "all fp32 (assoc Z@A3V first)": plan(assoc=True),
  "big3 tf32 (rn)": plan(qkv="tf32", out="tf32", fc1="tf32"),
  "big3 tf32 (trunc)": plan(qkv="tf32t", out="tf32t", fc1="tf32t"),
  "big3 fp16": plan(qkv="fp16", out="fp16", fc1="fp16"),
  "big3+fcb fp16": plan(qkv="fp16", out="fp16", fc1="fp16", fcb="fp16"),
  "qkv fp16 only": plan(qkv="fp16"),
  "out fp16 only": plan(out="fp16"),
  "fc1 fp16 only": plan(fc1="fp16"),
  "fcb fp16 only": plan(fcb="fp16"),
  "big3 bf16": plan(qkv="bf16", out="bf16", fc1="bf16"),
  "big3+fcb bf16x3": plan(qkv="bf16x3", out="bf16x3", fc1="bf16x3", fcb="bf16x3"),
  "big3+fcb fp16x2(A split)": plan(qkv="fp16x2", out="fp16x2", fc1="fp16x2", fcb="fp16x2"),
  "big3 fp16x2w, fcb fp16x3": plan(qkv="fp16x2w", out="fp16x2w", fc1="fp16x2w", fcb="fp16x3"),
  "qkv fp16x2w, rest fp16x3": plan(qkv="fp16x2w", out="fp16x3", fc1="fp16x3", fcb="fp16x3"),
  "out fp16x2w, rest fp16x3": plan(qkv="fp16x3", out="fp16x2w", fc1="fp16x3", fcb="fp16x3"),
  "fc1 fp16x2w, rest fp16x3": plan(qkv="fp16x3", out="fp16x3", fc1="fp16x2w", fcb="fp16x3"),
  "qkv+out fp16x2w, rest fp16x3": plan(qkv="fp16x2w", out="fp16x2w", fc1="fp16x3", fcb="fp16x3"),
  "big3 fp16x2 (A split), fcb fp16x3": plan(qkv="fp16x2", out="fp16x2", fc1="fp16x2", fcb="fp16x3"),
  "big3+fcb fp16x3": plan(qkv="fp16x3", out="fp16x3", fc1="fp16x3", fcb="fp16x3"),
  "big3+fcb tf32x3": plan(qkv="tf32x3", out="tf32x3", fc1="tf32x3", fcb="tf32x3"),
  "attn core tf32 (sim,agg), pinv fp32": plan(sim="tf32", sim2="tf32", agg="tf32"),
  "pinv tf32 only": plan(pinv="tf32"),
  "pinv bf16x3 only": plan(pinv="bf16x3"),
  "pinv tf32x3 only": plan(pinv="tf32x3"),
  "everything tf32": plan(qkv="tf32", out="tf32", fc1="tf32", fcb="tf32", sim="tf32", sim2="tf32", pinv="tf32", agg="tf32"),
  "everything fp16, pinv tf32": plan(qkv="fp16", out="fp16", fc1="fp16", fcb="fp16", sim="fp16", sim2="fp16", pinv="tf32", agg="fp16"),
  "big3+fcb fp16, core bf16x3": plan(qkv="fp16", out="fp16", fc1="fp16", fcb="fp16", sim="bf16x3", sim2="bf16x3", pinv="bf16x3", agg="bf16x3"),
}

if __name__ == "__main__":
    torch.set_num_threads(8)
    cases = [(320, [12], 5, "default", 12345, 12345), (320, [12], 5, "xavier", 12345, 777),
             (450, [4, 8, 16, 32], 7, "xavier", 3, 4), (800, [12], 5, "default", 9, 10), (2048, [4, 8, 16, 32], 5, "xavier", 13, 14)]
    refs = []
    for T, sc, dep, init, xs, ws in cases:
        x = orc.synth_features(T, xs); p = orc.synth_params(ws, init)
        with torch.no_grad():
            c64, l64 = orc.dsnet_forward(x.double(), {k: v.double() for k, v in p.items()}, sc, dep)
        refs.append((x, p, sc, dep, c64.numpy(), l64.numpy()))
    for name, P in PLANS.items():
        row = []
        for x, p, sc, dep, c64, l64 in refs:
            with torch.no_grad():
                c, l = forward(x, p, sc, dep, P)
            row.append(f"{orc.rel_l2(c.numpy(), c64):.1e}/{orc.rel_l2(l.numpy(), l64):.1e}|{orc.rel_max(l.numpy(), l64):.1e}")
        print(f"{name:42s}", "  ".join(row))
