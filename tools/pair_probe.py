"""Two-pass GEMM: one-CTA kernel (variant 0) against the CTA-pair kernel (EDSNET_TC_VARIANT=5), same operands:
bit difference of the outputs and time per launch.  python tools/pair_probe.py [rows...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edsnet_b200 import _capi  # noqa: E402

DEV = "cuda:0"


def run(lib, A16, B16, A, B, Cd, M, N, K, bias, res, epi, iters):
    st = torch.cuda.current_stream().cuda_stream

    def go():
        _capi.check(lib.edsnet_gemm(_capi.PRECISIONS["fp16x2"], epi, A.data_ptr(), A16.data_ptr(), B.data_ptr(),
                                    B16.data_ptr(), Cd.data_ptr(), M, N, K, bias.data_ptr(), res.data_ptr(), 512, st))
    go()
    torch.cuda.synchronize()
    if not iters:
        return None
    for _ in range(3):
        go()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        go()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rows = [int(v) for v in sys.argv[1:]] or [1000, 917968]
    lib = _capi.lib()
    g = torch.Generator(device=DEV).manual_seed(1)
    st = torch.cuda.current_stream().cuda_stream
    for M in rows:
        for name, (N, K, epi) in {"to_qkv": (1536, 1024, 1), "to_out-shape": (1024, 512, 3)}.items():
            A = torch.relu(torch.randn(M, K, generator=g, device=DEV)) * 0.05
            B = (torch.rand(N, K, generator=g, device=DEV) * 2 - 1) / K ** 0.5
            bias = torch.randn(N, generator=g, device=DEV) * 0.1
            res = torch.randn(M, N, generator=g, device=DEV) * 0.1
            A16 = torch.empty(lib.edsnet_split_f16_bytes(M, K), dtype=torch.uint8, device=DEV)
            B16 = torch.empty(lib.edsnet_split_f16_bytes(N, K), dtype=torch.uint8, device=DEV)
            _capi.check(lib.edsnet_split_f16(A.data_ptr(), A16.data_ptr(), M, K, st))
            _capi.check(lib.edsnet_split_f16(B.data_ptr(), B16.data_ptr(), N, K, st))
            out, ms = {}, {}
            for v in (0, 5):
                _capi.check(lib.edsnet_debug_set_tc_variant(v))
                Cd = torch.full((M, N), float("nan"), device=DEV)
                ms[v] = run(lib, A16, B16, A, B, Cd, M, N, K, bias, res, epi, 10 if M > 100000 else 0)
                out[v] = Cd
                assert lib.edsnet_debug_tc_status(1) == 0, f"variant {v}: pipeline time-out"
            m = min(M, 2048)
            ref = A[:m].double() @ B.double().t()
            if epi == 1:
                ref[:, :512] *= 0.125
            if epi >= 2:
                ref += bias.double()
            if epi == 3:
                ref += res[:m].double()
            err = {v: float((out[v][:m].double() - ref).norm() / ref.norm()) for v in out}
            same = bool(torch.equal(out[0], out[5]))
            print(f"M={M} {name}: identical={same} max|diff|={float((out[0] - out[5]).abs().max()):.3e} "
                  f"err0={err[0]:.2e} err5={err[5]:.2e} ms0={ms[0]} ms5={ms[5]}", flush=True)
            del A, B, res, out
    _capi.check(lib.edsnet_debug_set_tc_variant(0))


if __name__ == "__main__":
    main()
