#!/usr/bin/env python3
"""Cholesky + inverse factor timings: dataflow kernel (chol_flow.cu) vs panel-step kernels (chol.cu) vs cuSOLVER potrf."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from nonstationary_precip_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for M in (512, 1024, 2048):
    f64 = dict(dtype=torch.float64, device="cuda")
    X = torch.rand(M, 3, **f64) * 2 - 1
    ell = torch.full((3, M), 0.3, **f64)
    A = ops.gibbs_diag_fwd(X, ell, X, ell) + 1e-6 * torch.eye(M, **f64)
    for impl in ("flow", "steps"):
        ms, best = timeit(lambda: ops.potrf_inv(A, impl=impl))
        print(json.dumps(dict(kernel="potrf_inv[%s] (incl. the input copy)" % impl, M=M, ms_median=round(ms, 4), ms_best=round(best, 4))), flush=True)
    ms, best = timeit(lambda: torch.linalg.cholesky(A))
    print(json.dumps(dict(kernel="cusolver potrf (no inverse; reference point)", M=M, ms_median=round(ms, 4), ms_best=round(best, 4))), flush=True)
    if M == 1024:  # two factorisations at once on two streams, as in the SVGP step
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        def both(impl):
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s1):
                ops.potrf_inv(A, impl=impl)
            with torch.cuda.stream(s2):
                ops.potrf_inv(A, impl=impl)
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        for impl in ("flow", "steps"):
            ms, best = timeit(lambda: both(impl))
            print(json.dumps(dict(kernel="2 x potrf_inv[%s] on two streams" % impl, M=M, ms_median=round(ms, 4), ms_best=round(best, 4))), flush=True)

# batched launch (round 2): n matrices in one interleaved dataflow kernel
if __name__ == "__main__":
    import json as _json
    from nonstationary_precip_b200 import ops as _ops
    for _n in (1, 2, 4):
        _g = torch.Generator().manual_seed(_n)
        _mats = []
        for _ in range(_n):
            _X = torch.rand(1024, 3, generator=_g, dtype=torch.float64) * 2 - 1
            _mats.append((torch.exp(-((_X[:, None] - _X[None]) ** 2).sum(-1) / 0.18) + 1e-6 * torch.eye(1024, dtype=torch.float64)).cuda())
        for _ in range(3):
            _ops.potrf_inv_batch(_mats)
        torch.cuda.synchronize()
        _ts = []
        for _ in range(10):
            _a, _b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            _a.record()
            _ops.potrf_inv_batch(_mats)
            _b.record()
            _b.synchronize()
            _ts.append(_a.elapsed_time(_b))
        print(_json.dumps({"kernel": "potrf_inv_flow_batch (incl. %d input copies and allocations)" % _n, "n": _n, "M": 1024,
                           "ms_median": round(sorted(_ts)[5], 4), "ms_best": round(min(_ts), 4)}), flush=True)
