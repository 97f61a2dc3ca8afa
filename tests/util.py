"""Shared helpers for the parity tests (oracle = checker only)."""
import hashlib
import os

import numpy as np
import torch

from oracle import dsnet_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerance of each arithmetic mode on pred_cls / pred_loc, rel-L2 against the fp32 reference output
# (BASELINE.json north_star: 1e-3 for tensor-core stages, 1e-5 for the FP32 mode).  fp16x3 is the hi/lo split
# tensor-core mode and is held to the fp32-mode bar.
# The single-pass fp16 mode is an opt-in fast mode that does NOT meet 1e-3 on every input (measured 3e-4 .. 1.3e-3 on
# pred_loc: eleven-bit operands through five LayerNorm blocks); it is held to 3e-3 and is never the default.
# fp16x2 (two MMA passes on to_qkv / to_out: activations at 11 bits, weights at 22) is the fast mode that DOES meet
# the 1e-3 bar, gated at half of it (profiles/r02_precision_probe.log: <= 4.2e-4 predicted).
TOL = {"fp32": 1e-5, "fp16x3": 1e-5, "fp16": 3e-3, "fp16x2": 5e-4}


def load_npz(name):
    return np.load(os.path.join(GOLD, name))


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def golden_case(fwd, name):
    g = {k.split("/", 1)[1]: fwd[k] for k in fwd.files if k.startswith(name + "/")}
    x = orc.synth_features(int(g["T"]), int(g["x_seed"]))
    p = orc.synth_params(int(g["w_seed"]), str(g["init"]))
    assert sha(x.numpy()) == str(g["x_sha"])
    return g, x, p


def ref_state_dict(p, fc_depth):
    """Reference DSNet.state_dict() layout incl. the aliased fc.N.* entries (anchor_based/dsnet.py:96)."""
    sd = dict(p)
    for i in range(fc_depth):
        for j in ("0.weight", "0.bias", "3.weight", "3.bias"):
            sd[f"fc.{i}.{j}"] = p[f"fc_block.{j}"]
    return sd


def make_model(p, scales, fc_depth, precision, device=None, base="nystromformer"):
    from edsnet_b200 import DSNet
    m = DSNet(base, 1024, 128, list(scales), 8, fc_depth=fc_depth, orientation=None,
              pooling_type="roi", precision=precision).eval()
    m.load_state_dict(ref_state_dict(p, fc_depth), strict=True)
    if device is not None:
        m = m.to(device)
    return m
