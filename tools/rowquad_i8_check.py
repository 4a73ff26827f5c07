#!/usr/bin/env python3
"""Correctness + timing of the int8 tcgen05 rowquad (csrc/ozaki.cu) against the FP64 DMMA rowquad and a float128-free
reference (torch fp64 matmul).  Run under `timeout` on the GPU box."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nonstationary_precip_b200 import ops  # noqa: E402


def case(n, M, seed=0, bench=False):
    g = torch.Generator().manual_seed(seed)
    K = torch.rand(n, M, generator=g, dtype=torch.float64).cuda() * torch.exp(
        3 * torch.randn(n, 1, generator=g, dtype=torch.float64)).cuda()
    A = torch.randn(M, M, generator=g, dtype=torch.float64).cuda()
    C = (A @ A.T) / M - torch.eye(M, dtype=torch.float64, device="cuda") * 0.3
    C = 0.5 * (C + C.T)
    T0, q0 = ops.rowquad(K, C)
    T1, q1 = ops.rowquad_i8(K, C)
    torch.cuda.synchronize()
    # error measured against the row/column scale the FP64 bound refers to: |K| |C|
    scale = (K.abs() @ C.abs())
    e_i8 = ((T1 - T0).abs() / scale).max().item()
    eq = ((q1 - q0).abs() / (scale * K.abs()).sum(1)).max().item()
    out = {"n": n, "M": M, "max|T_i8-T_f64|/(|K||C|)": e_i8, "max|q_i8-q_f64|/sum(|K||C||K|)": eq}
    if bench:
        for fn, name in ((lambda: ops.rowquad(K, C), "dmma"), (lambda: ops.rowquad_i8(K, C), "i8")):
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
    case(128, 64)
    case(300, 128, seed=1)
    case(1000, 256, seed=2)
    if len(sys.argv) > 1:
        from nonstationary_precip_b200._lib import lib
        dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
        lib().npgp_rowquad_i8_debug(dbg.data_ptr())
        case(65536, 1024, seed=3, bench=True)
        names = ["producer_wait_empty", "mma_wait_acc_empty", "mma_wait_full", "mma_total", "epilogue_wait_acc_full"]
        print(json.dumps(dict(zip(names, dbg.tolist()[:5]))))
        lib().npgp_rowquad_i8_debug(None)
