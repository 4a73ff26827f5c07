#!/usr/bin/env python3
"""Timeline of ONE replayed SVGP-Gibbs step graph (BASELINE config 2 shapes): start / end of every section on the GPU's own
timer, stamped by capturable one-thread kernels (npgp_timestamp) on the stream each section runs on.  `--ranks G` runs the
local work of one rank of a G-rank job (B/G rows, replicated terms weighted 1/G, no collective) on this GPU.
  python tools/step_timeline.py [--ranks 8] [--variant full]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nonstationary_precip_b200.svgp import SVGPGibbs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=1)
ap.add_argument("--variant", default="full")
ap.add_argument("--replays", type=int, default=10)
ap.add_argument("--engine", default="c", choices=["c", "python"])
args = ap.parse_args()
dev = torch.device("cuda", 0)
x_h, y_h, perm = bench.make_data(bench.N_TOTAL, bench.DIM)
kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in bench.make_params(args.variant, bench.M_IND, bench.DIM).items()}
model = SVGPGibbs(args.variant, x_h[perm[:bench.M_IND]].to(dev), bench.N_TOTAL, **kw)
Bl = bench.B_GLOBAL // args.ranks
if args.engine == "c":
    model.use_c_engine()
model.timeline = torch.zeros(128, dtype=torch.int64, device=dev)
model.capture(Bl, args.ranks, bench.B_GLOBAL, lr=0.01)
X, Y = x_h[:Bl].to(dev), y_h[:Bl].to(dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    model.train_step_graph(X, Y)
torch.cuda.synchronize()
ev0.record()
for _ in range(args.replays):
    model.train_step_graph(X, Y)
ev1.record()
torch.cuda.synchronize()
tl = model.timeline_ms()
out = {"engine": args.engine, "ranks": args.ranks, "B_local": Bl, "ms_per_step": round(ev0.elapsed_time(ev1) / args.replays, 4),
       "timeline_ms(start,end)": dict(sorted(tl.items(), key=lambda kv: kv[1][0]))}
print(json.dumps(out))
for k, (a, b) in sorted(tl.items(), key=lambda kv: kv[1][0]):
    print("%-22s %8.3f -> %8.3f  (%6.3f ms)" % (k, a, b, b - a), file=sys.stderr)
