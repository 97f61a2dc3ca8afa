"""CPU-only tests: C-ABI library loads and exports every declared symbol, host planning / chunking / sharding
logic, module surface (state-dict layout, error behaviour).  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest
import torch

import edsnet_b200
from edsnet_b200 import BatchPlan, DSNet, ScoringPipeline, _capi, shard_videos
from oracle import dsnet_oracle as orc
from tests.util import load_npz, ref_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = _capi.lib()
    header = open(os.path.join(ROOT, "include", "edsnet_b200.h")).read()
    declared = set(re.findall(r"\b(edsnet_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.edsnet_abi_version() == _capi.EDSNET_ABI_VERSION


def test_workspace_sizing_and_argument_errors_without_gpu():
    lib = _capi.lib()
    cfg = _capi.make_config([4, 8], 5, _capi.PREC_FP16X3)
    import ctypes as C
    L = _capi.WorkspaceLayout()
    total = lib.edsnet_workspace_bytes(cfg, 1000, 3, C.byref(L))
    assert total == L.total and L.total > 1000 * 1536 * 4
    assert L.qkv == 0 and L.merged > L.qkv and L.x16 >= L.u1
    small = lib.edsnet_workspace_bytes(_capi.make_config([4], 5, _capi.PREC_FP32), 1000, 3, None)
    assert small < total
    # argument validation happens before any CUDA call
    assert lib.edsnet_forward(None, None, None, None, None, None, None, 0, None) == _capi.E_ARG
    bad = _capi.make_config([5], 5, 0)
    assert lib.edsnet_decode_boxes(bad, None, None, None, None, None) == _capi.E_ARG
    assert "odd anchor scale" in _capi.last_error()
    assert lib.edsnet_forward_launches(cfg) == 11
    assert lib.edsnet_forward_launches(_capi.make_config([4, 8], 5, _capi.PREC_FP16)) == 12    # keeps the LayerNorm kernel
    assert lib.edsnet_forward_launches(_capi.make_config([4, 8], 5, _capi.PREC_FP32)) == 11
    # the LayerNorm-fold layout: fc1 operand planes of z in the y region (+ 4 bytes of scale per row), row statistics
    assert L.u0 - L.y >= 1000 * 1024 * 4 + 1000 * 4 and L.zstat > L.zeros and L.xstat >= L.zstat + 1000 * 32 * 4
    # the stage entry point only exposes the general epilogues (the LayerNorm-fold pair is internal to the forward)
    assert lib.edsnet_gemm(_capi.PREC_FP16X3, 5, None, None, None, None, None, 128, 1024, 512, None, None, 0, None) == _capi.E_ARG
    assert "unknown epilogue" in _capi.last_error()
    # a tcgen05 forward without the derived LayerNorm-fold operands is refused before any launch
    plan = BatchPlan.build([100, 200])
    fake = 1 << 20                                              # never dereferenced: validation comes first
    w = _capi.Weights()
    for name in _capi.WEIGHT_FIELDS:
        setattr(w, name, None if name.startswith(("fc1_fold", "to_out_b")) and name != "to_out_b" else fake)
    b = _capi.Batch(2, 300, 200, fake, fake, plan.tiles64.shape[0], fake, plan.tiles128.shape[0], None)
    need = lib.edsnet_workspace_bytes(cfg, 300, 2, None)
    assert lib.edsnet_forward(cfg, C.byref(w), C.byref(b), fake, fake, fake, fake, need, None) == _capi.E_ARG
    assert "LayerNorm-folded" in _capi.last_error()


def test_batch_plan_tables():
    plan = BatchPlan.build([1, 64, 65, 300])
    assert plan.cu_rows.tolist() == [0, 1, 65, 130, 430]
    assert plan.tiles64.tolist() == [[0, 0], [1, 0], [2, 0], [2, 64], [3, 0], [3, 64], [3, 128], [3, 192], [3, 256]]
    assert plan.tiles128.tolist() == [[0, 0], [1, 0], [2, 0], [3, 0], [3, 128], [3, 256]]
    assert plan.n_videos == 4 and plan.total_rows == 430 and plan.max_rows == 300
    off, total = BatchPlan.build([100, 2000, 5000]).nms_scratch(4)
    assert off.tolist() == [0, 0, 8192 * 48] and total == (8192 + 32768) * 48
    with pytest.raises(ValueError):
        BatchPlan.build([])
    with pytest.raises(ValueError):
        BatchPlan.build([10, 0])


def test_chunking_and_sharding():
    lengths = [100, 200, 300, 50, 700, 20]
    chunks = ScoringPipeline.chunk_videos(lengths, 400)
    assert chunks == [(0, 2), (2, 4), (4, 5), (5, 6)]
    assert ScoringPipeline.chunk_videos([1000], 400) == [(0, 1)]
    rng = np.random.default_rng(0)
    lens = rng.integers(100, 801, size=4096)
    for w in (1, 2, 4, 8):
        parts = shard_videos(lens, w)
        allv = sorted(i for p in parts for i in p)
        assert allv == list(range(4096))
        loads = [int(sum(((lens[i] + 63) // 64) * 64 for i in p)) for p in parts]
        assert max(loads) - min(loads) <= 832


def test_module_surface_matches_reference_state_dict():
    g = load_npz("forward_golden.npz")
    p = orc.synth_params(1, "default")
    m = DSNet("nystromformer", 1024, 128, 12, 8, fc_depth=5, orientation=None, pooling_type="roi")
    assert m.anchor_scales == [12] and m.num_scales == 1            # int -> list, dsnet.py:69-70
    sd = m.state_dict()
    want = ref_state_dict(p, 5)
    assert set(sd) == set(want)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(want[k].shape), k
    m.load_state_dict(want, strict=True)
    assert m.fc[0][0].weight is m.fc_block[0].weight                # ONE shared block
    assert sum(q.numel() for q in m.parameters()) == 2248843        # unique parameters of the reference model

    def xavier_init(module):                                        # anchor_based/train.py:19-24
        name = module.__class__.__name__
        if "Linear" in name or "Conv" in name:
            torch.nn.init.xavier_uniform_(module.weight, gain=np.sqrt(2.0))
            if module.bias is not None:
                torch.nn.init.constant_(module.bias, 0.1)
    m.apply(xavier_init)
    assert float(m.fc1.bias[0].detach()) == pytest.approx(0.1)
    assert float(m.base_model.to_out[0].bias[0].detach()) == pytest.approx(0.1)


def test_unsupported_configurations_raise():
    with pytest.raises(ValueError):
        DSNet("lstm", 1024, 128, [4], 8, pooling_type="roi")
    with pytest.raises(ValueError):
        DSNet("nystromformer", 1024, 128, [4], 8)                   # default pooling_type='fft'
    with pytest.raises(ValueError):
        DSNet("nystromformer", 512, 128, [4], 8, pooling_type="roi")
    m = DSNet("nystromformer", 1024, 128, [4], 8, pooling_type="roi")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 16, 1024))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.predict(torch.zeros(1, 16, 1024))


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "edsnet-efficient-dsnet-for-video-summarization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle\.", src, re.M), f


def test_attention_base_state_dict_matches_reference_layout():
    p = orc.synth_params_mha(3, "default")
    m = DSNet("attention", 1024, 128, [4, 8], 8, fc_depth=3, pooling_type="roi")
    want = ref_state_dict(p, 3)
    assert set(m.state_dict()) == set(want)
    m.load_state_dict(want, strict=True)


def test_layernorm_fold_algebra_on_cpu():
    """The operands dsnet.py derives for the LayerNorm-folded fc1 reproduce fc1(LayerNorm(y)) for any row shift."""
    from edsnet_b200.dsnet import layernorm_fold_operands
    g = torch.Generator().manual_seed(3)
    R = 37
    ln_w, ln_b = torch.rand(1024, generator=g) + 0.5, torch.randn(1024, generator=g) * 0.1
    fc1_w, fc1_b = torch.randn(128, 1024, generator=g) / 32, torch.randn(128, generator=g) * 0.1
    to_out_w, to_out_b = torch.randn(1024, 512, generator=g) / 22, torch.full((1024,), 0.1) + torch.randn(1024, generator=g) * 0.01
    x = torch.relu(torch.randn(R, 1024, generator=g)) * 0.03
    merged = torch.randn(R, 512, generator=g) * 0.02
    op = layernorm_fold_operands(ln_w, ln_b, fc1_w, fc1_b, to_out_w, to_out_b)
    y = (merged.double() @ to_out_w.double().t() + to_out_b.double() + x.double())
    want = torch.nn.functional.layer_norm(y, (1024,), ln_w.double(), ln_b.double(), 1e-5) @ fc1_w.double().t() + fc1_b.double()
    # what the two epilogues compute (in float64 here): z with the centred bias and the row mean of x removed
    z = merged.double() @ to_out_w.double().t() + op["to_out_bc"].double() + (x.double() - x.double().mean(1, keepdim=True))
    mean = z.mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt((z * z).mean(1, keepdim=True) - mean * mean + 1e-5)
    got = rstd * (z @ op["fc1_fold_w"].double().t() - mean * op["fc1_fold_wgsum"].double()[None, :]) + op["fc1_fold_b"].double()[None, :]
    assert float((got - want).norm() / want.norm()) < 1e-6           # the operands themselves are fp32
    # the row-scale bound the to_out epilogue uses really bounds |z|
    l1max, bmax = [float(v) for v in op["to_out_bounds"]]
    bound = 2 * x.abs().amax(1) + bmax + merged.abs().amax(1) * l1max
    assert bool((z.abs().amax(1) <= bound.double()).all())
