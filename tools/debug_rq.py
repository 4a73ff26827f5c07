#!/usr/bin/env python3
"""Debug aid: reads back the K tile the row-quadratic kernel rebuilds from the digit planes (through du_part with a one-hot
gvec) and compares it with the FP64 kernel matrix."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from nonstationary_precip_b200 import ops  # noqa: E402
from test_digits_gpu import decode_planes, fwd_both, gibbs_inputs  # noqa: E402

torch.set_default_dtype(torch.float64)
n, M = 300, 128
x, f1, z, f2, s, u = gibbs_inputs("full", n, M, 3, seed=1)
K, Ku, digits, parts = fwd_both("full", x, f1, z, f2, s, u)
Kd = decode_planes(digits, n, M) * 2.0 ** -54
print("decode vs K:", (Kd[:n] - K).abs().max().item())
C = torch.eye(M, device="cuda")
Cd = torch.empty(ops.digits_bytes(M, M, 64), dtype=torch.uint8, device="cuda")
cexp = torch.empty(M, dtype=torch.int32, device="cuda")
ops.o8_slice_rows(C, 64, Cd, cexp)
T = torch.empty(n, M, device="cuda")
q_part = torch.empty(M // 64, n, device="cuda")
du_part = torch.empty((n + 127) // 128, M, device="cuda")
for r in (0, 1, 5, 37, 64, 100, 127, 128, 200, 299):
    g = torch.zeros(n, device="cuda")
    g[r] = 1.0
    ops.o8_rowquad_digits(n, M, digits, s, Cd, cexp, T, q_part=q_part, gvec=g, du_part=du_part)
    got = du_part[r // 128]
    err = (got - K[r]).abs()
    print("row", r, "max err", err.max().item(), "argmax col", int(err.argmax()), "got", got[:4].tolist(), "want", K[r, :4].tolist())
    print("   T row err", (T[r] - K[r]).abs().max().item(), "q", q_part[:, r].sum().item(), "want", (K[r] * K[r]).sum().item())
