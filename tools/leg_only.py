#!/usr/bin/env python3
"""Run one of bench.py's legs (c5 | c4 | c3) alone on one GPU:  python tools/leg_only.py c4 [USE_C_LAYER=0|1]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bench_legs  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c4"
if len(sys.argv) > 2:
    from nonstationary_precip_b200.models import dgps
    dgps.USE_C_LAYER = bool(int(sys.argv[2]))
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = bench_legs.run_all(0, 1, dev, lambda v: v, None, bench.make_params, which=[which])
print(json.dumps(out))
