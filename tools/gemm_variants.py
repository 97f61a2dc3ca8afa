"""Time and check the tile/accumulator variants of the tcgen05 GEMM (edsnet_debug_set_tc_variant) on one B200:
to_qkv / to_out shapes of a 230k-row launch, GEMM error vs fp64, and the forward goldens per variant.
Usage: python tools/gemm_variants.py [variants...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edsnet_b200 import _capi  # noqa: E402
from oracle import dsnet_oracle as orc  # noqa: E402
from tests.util import golden_case, load_npz, make_model  # noqa: E402

DEV = "cuda:0"


def run_gemm(lib, A, B, bias, res, epi, iters=0):
    M, K = A.shape
    N = B.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    A16 = torch.empty(lib.edsnet_split_f16_bytes(M, K), dtype=torch.uint8, device=DEV)
    B16 = torch.empty(lib.edsnet_split_f16_bytes(N, K), dtype=torch.uint8, device=DEV)
    _capi.check(lib.edsnet_split_f16(A.data_ptr(), A16.data_ptr(), M, K, st))
    _capi.check(lib.edsnet_split_f16(B.data_ptr(), B16.data_ptr(), N, K, st))
    Cd = torch.empty((M, N), device=DEV)

    def go():
        _capi.check(lib.edsnet_gemm(_capi.PRECISIONS["fp16x3"], epi, A.data_ptr(), A16.data_ptr(), B.data_ptr(),
                                    B16.data_ptr(), Cd.data_ptr(), M, N, K, bias.data_ptr(), res.data_ptr(), 512, st))
    go()
    torch.cuda.synchronize()
    ms = None
    if iters:
        for _ in range(3):
            go()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            go()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    return Cd, ms


def main():
    variants = [int(v) for v in sys.argv[1:]] or [0, 3, 4]
    lib = _capi.lib()
    fwd = load_npz("forward_golden.npz")
    g = torch.Generator(device=DEV).manual_seed(1)
    M = 230528
    shapes = {"to_qkv": (1536, 1024, 1), "to_out": (1024, 512, 3), "fc1": (128, 1024, 2)}
    for v in variants:
        _capi.check(lib.edsnet_debug_set_tc_variant(v))
        line = [f"variant {v}:"]
        for name, (N, K, epi) in shapes.items():
            A = torch.relu(torch.randn(M, K, generator=g, device=DEV)) * 0.05
            B = (torch.rand(N, K, generator=g, device=DEV) * 2 - 1) / K ** 0.5
            bias = torch.randn(N, generator=g, device=DEV) * 0.1
            res = torch.randn(M, N, generator=g, device=DEV) * 0.1
            Cd, ms = run_gemm(lib, A, B, bias, res, epi, iters=10)
            tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
            # error vs fp64 on the first 2048 rows
            m = 2048
            ref = A[:m].double() @ B.double().t()
            if epi == 1:
                ref[:, :512] *= 0.125
            if epi >= 2:
                ref += bias.double()
            if epi == 3:
                ref += res[:m].double()
            err = float((Cd[:m].double() - ref).norm() / ref.norm())
            line.append(f"{name} {ms:.3f} ms ({tf:.0f} TF/s algorithmic, x3 = {3 * tf:.0f}) err {err:.2e};")
            del A, B, res, Cd
        print(" ".join(line), flush=True)
        assert lib.edsnet_debug_tc_status(1) == 0
        worst = [0.0, 0.0]
        for name in list(fwd["forward_cases"]):
            gg, x, p = golden_case(fwd, name)
            scales = [int(s) for s in gg["scales"]]
            model = make_model(p, scales, int(gg["fc_depth"]), "fp16x3", DEV)
            with torch.no_grad():
                cls, loc = model(x[None].to(DEV))
            e_cls = orc.rel_l2(cls.cpu().numpy(), gg["pred_cls"])
            e_loc = orc.rel_l2(loc.cpu().numpy(), gg["pred_loc"])
            worst = [max(worst[0], e_cls), max(worst[1], e_loc)]
            print(f"   {name}: cls {e_cls:.2e} loc {e_loc:.2e}")
        print(f"   worst over goldens: cls {worst[0]:.2e} loc {worst[1]:.2e}", flush=True)
        assert lib.edsnet_debug_tc_status(1) == 0
    _capi.check(lib.edsnet_debug_set_tc_variant(0))


if __name__ == "__main__":
    main()
