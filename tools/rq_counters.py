#!/usr/bin/env python3
"""Cycle counters of CTA 0 of o8_rowquad_kernel at the C2 shapes (B=65536, M=1024): how long the producer waits for a free
stage, the MMA issuer for operands / for the epilogue to drain TMEM, the epilogue for the accumulators.
Run on the GPU box:  python tools/rq_counters.py > gpurun_out/rq_counters.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nonstationary_precip_b200 import ops  # noqa: E402
from nonstationary_precip_b200._lib import lib  # noqa: E402

torch.manual_seed(0)
B, M, d = int(os.environ.get("B", 65536)), int(os.environ.get("M", 1024)), 3
f64 = dict(dtype=torch.float64, device="cuda")
x = torch.rand(B, d, **f64) * 2 - 1
z = torch.rand(M, d, **f64) * 2 - 1
Dm = torch.diag(torch.tensor([1.2, -1.3, 1.25], **f64))
Sx, Sz = ops.sigma_from_h_fwd(torch.randn(B, d, **f64), Dm), ops.sigma_from_h_fwd(torch.randn(M, d, **f64), Dm)
s = torch.tensor([0.644], **f64)
u = torch.randn(M, **f64)
digits = torch.empty(ops.digits_bytes(B, M, 128), dtype=torch.uint8, device="cuda")
parts = torch.empty(ops.gibbs_digits_splits(B, M), B, **f64)
ops.gibbs_full_fwd_digits(x, Sx, z, Sz, 1e-5, s, digits, u=u, Ku_part=parts)
A = torch.randn(M, M, **f64)
C = A @ A.T / M - 0.3 * torch.eye(M, **f64)
C = 0.5 * (C + C.T)
Cd = torch.empty(ops.digits_bytes(M, M, 64), dtype=torch.uint8, device="cuda")
cexp = torch.empty(M, dtype=torch.int32, device="cuda")
ops.o8_slice_rows(C, 64, Cd, cexp)
T = torch.empty(B, M, **f64)
q_part = torch.empty(M // 64, B, **f64)
du_part = torch.empty((B + 127) // 128, M, **f64)
gvec = torch.randn(B, **f64)
names = ["producer_wait_empty", "mma_wait_acc_empty", "mma_wait_full", "mma_total", "epilogue_wait_acc_full"]
for what, kw in (("T only", {}), ("T, q, K^T g", dict(q_part=q_part, gvec=gvec, du_part=du_part))):
    for _ in range(2):
        ops.o8_rowquad_digits(B, M, digits, s, Cd, cexp, T, **kw)
    dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
    lib().npgp_rowquad_i8_debug(dbg.data_ptr())
    ops.o8_rowquad_digits(B, M, digits, s, Cd, cexp, T, **kw)
    torch.cuda.synchronize()
    lib().npgp_rowquad_i8_debug(None)
    out = dict(zip(names, dbg.tolist()[:5]))
    tiles = (B // 128) * (M // 64) / 148.0
    out.update(what=what, tiles_per_cta=round(tiles, 2), cycles_per_tile=round(out["mma_total"] / tiles),
               mma_wait_full_frac=round(out["mma_wait_full"] / out["mma_total"], 4),
               mma_wait_acc_empty_frac=round(out["mma_wait_acc_empty"] / out["mma_total"], 4),
               ideal_cycles_per_tile_at_36_per_mma=(M // 32) * 28 * 36)
    print(json.dumps(out), flush=True)
