#!/usr/bin/env python3
"""Timings of the M x M x M GEMM shapes of the replicated O(M^3) chain (svgp.py::_zz_forward / m3 backward), with the
triangular-operand flags they use."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nonstationary_precip_b200 import ops  # noqa: E402

M = int(os.environ.get("M", 1024))
g = torch.Generator().manual_seed(0)
A = torch.randn(M, M, generator=g, dtype=torch.float64).cuda()
Lo = torch.tril(A).contiguous()
Up = torch.triu(A).contiguous()
cases = {
    "dense A@A": lambda: ops.dgemm(A, A),
    "E@P (tri_b=lower)": lambda: ops.dgemm(A, Lo, tri_b=1),
    "P^T@X (tri_a=upper)": lambda: ops.dgemm(Lo, A, transA=True, tri_a=2),
    "P@dC (tri_a=lower)": lambda: ops.dgemm(Lo, A, tri_a=1),
    "W@P^T (tri_b=upper)": lambda: ops.dgemm(A, Lo, transB=True, tri_b=2),
    "Ls@Ls^T (lower x upper)": lambda: ops.dgemm(Lo, Lo, transB=True, tri_a=1, tri_b=2),
    "P^T@X (upper x lower)": lambda: ops.dgemm(Lo, Lo, transA=True, tri_a=2, tri_b=1),
}
ref = {"dense A@A": A @ A, "E@P (tri_b=lower)": A @ Lo, "P^T@X (tri_a=upper)": Lo.T @ A, "P@dC (tri_a=lower)": Lo @ A,
       "W@P^T (tri_b=upper)": A @ Lo.T, "Ls@Ls^T (lower x upper)": Lo @ Lo.T, "P^T@X (upper x lower)": Lo.T @ Lo}
for name, fn in cases.items():
    out = fn()
    err = ((out - ref[name]).abs().max() / ref[name].abs().max()).item()
    # time 20 back-to-back calls replayed as one CUDA graph (eager launches of ~60 us kernels are CPU bound)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gr.replay()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / 20)
    ts.sort()
    print(json.dumps({"gemm": name, "M": M, "ms_median": round(ts[len(ts) // 2], 4), "ms_best": round(ts[0], 4),
                      "rel_err": err}), flush=True)
