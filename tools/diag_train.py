#!/usr/bin/env python3
"""Training-health trace at the C2 shapes: loss, Cholesky info, D, min det of Sigma per step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nonstationary_precip_b200 import ops  # noqa: E402
from nonstationary_precip_b200.svgp import SVGPGibbs  # noqa: E402

variant = os.environ.get("VARIANT", "full")
steps = int(os.environ.get("STEPS", 30))
dev = torch.device("cuda", 0)
x_h, y_h, perm = bench.make_data(bench.N_TOTAL, bench.DIM)
kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in bench.make_params(variant, bench.M_IND, bench.DIM).items()}
model = SVGPGibbs(variant, x_h[perm[:bench.M_IND]].to(dev), bench.N_TOTAL, **kw)
X, Y = x_h.to(dev), y_h.to(dev)
B = bench.B_GLOBAL
for k in range(steps):
    lo = (k % 16) * B
    loss = model.train_step(X[lo:lo + B], Y[lo:lo + B], lr=0.01)
    msg = "step %2d loss %.6f info %d" % (k, loss.item(), int(model.last["info"]))
    if variant == "full":
        from oracle import gibbs_oracle as o
        Sz = ops.sym_unpack(ops.sigma_from_h_fwd(model.p["H"], model.p["D"]), 3)
        msg += " D %s min_eig_Sz %.4f" % (torch.diagonal(model.p["D"]).tolist(), torch.linalg.eigvalsh(Sz).min().item())
    print(msg, flush=True)
