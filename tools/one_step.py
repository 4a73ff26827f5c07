#!/usr/bin/env python3
"""One SVGP-Gibbs ELBO step at the C2 shapes inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nonstationary_precip_b200.svgp import SVGPGibbs  # noqa: E402

variant = os.environ.get("VARIANT", "full")
B = int(os.environ.get("B", bench.B_GLOBAL))
dev = torch.device("cuda", 0)
x_h, y_h, perm = bench.make_data(bench.N_TOTAL, bench.DIM)
kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in bench.make_params(variant, bench.M_IND, bench.DIM).items()}
model = SVGPGibbs(variant, x_h[perm[:bench.M_IND]].to(dev), bench.N_TOTAL, **kw)
if os.environ.get("ENGINE", "c") == "c":
    model.use_c_engine()  # the whole step is one C call (npgp_svgp_step)
X, Y = x_h[:B].to(dev), y_h[:B].to(dev)
for _ in range(2):
    model.train_step(X, Y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.train_step(X, Y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", model.grad[-2].item())
