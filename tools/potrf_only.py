#!/usr/bin/env python3
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nonstationary_precip_b200 import ops
M = int(os.environ.get("M", 1024))
g = torch.Generator().manual_seed(0)
X = torch.rand(M, 3, generator=g, dtype=torch.float64).cuda() * 2 - 1
ell = torch.full((3, M), 0.3, dtype=torch.float64, device="cuda")
A = ops.gibbs_diag_fwd(X, ell, X, ell) + 1e-6 * torch.eye(M, dtype=torch.float64, device="cuda")
for _ in range(2):
    ops.potrf_inv(A)
torch.cuda.synchronize()
torch.cuda.profiler.start()
L, P, info = ops.potrf_inv(A)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("info", int(info), "resid", ((L @ L.T - A).abs().max()).item())
