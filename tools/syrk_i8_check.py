#!/usr/bin/env python3
"""Correctness + timing of the int8 tcgen05 SYRK (csrc/ozaki.cu) against the FP64 DMMA wsyrk."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nonstationary_precip_b200 import ops  # noqa: E402


def case(n, M, seed=0, bench=False):
    g = torch.Generator().manual_seed(seed)
    K = (torch.rand(n, M, generator=g, dtype=torch.float64) * torch.exp(
        2 * torch.randn(1, M, generator=g, dtype=torch.float64))).cuda()
    w0 = torch.tensor([-0.37], dtype=torch.float64, device="cuda")
    ref = ops.wsyrk(K, alpha=-0.37 * 2.0)
    got = ops.syrk_i8(K, w0=w0, alpha=2.0)
    torch.cuda.synchronize()
    scale = 0.74 * (K.abs().T @ K.abs())
    out = {"n": n, "M": M, "max|i8-f64|/(|K|^T|K|)": ((got - ref).abs() / scale).max().item(),
           "sym": (got - got.T).abs().max().item()}
    if bench:
        for fn, name in ((lambda: ops.wsyrk(K), "dmma"), (lambda: ops.syrk_i8(K), "i8")):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            out["ms_" + name] = min(ts)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    case(64, 128)
    case(1000, 128, seed=1)
    case(5000, 256, seed=2)
    if len(sys.argv) > 1:
        case(65536, 1024, seed=3, bench=True)
        case(8192, 1024, seed=4, bench=True)
