#!/usr/bin/env python3
"""int8 tensor ceiling of a CTA pair (tcgen05.mma.cta_group::2) against the single-CTA probe, at the digit engine's tile
shape (N = 64) and at N = 256.   python tools/probe_pair.py > gpurun_out/probe_pair.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nonstationary_precip_b200._lib import check, lib, stream  # noqa: E402


def t(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


reps = 4096
for n_tile in (64, 256):
    for coll in (0, 1):
        ms1 = t(lambda: check(lib().npgp_i8_peak_probe(n_tile, coll, 148, reps, stream()), "probe"))
        ms2 = t(lambda: check(lib().npgp_i8_peak_probe_pair(n_tile, coll, 148, reps, stream()), "probe pair"))
        print(json.dumps({"n_tile": n_tile, "collector": coll,
                          "single_cta_tops": round(148.0 * reps * 8 * 2 * 128 * n_tile * 32 / ms1 / 1e9, 1),
                          "single_cycles_per_mma": round(ms1 * 1e-3 * 1.965e9 / (reps * 8), 1),
                          "pair_tops": round(74.0 * reps * 8 * 2 * 256 * n_tile * 32 / ms2 / 1e9, 1),
                          "pair_cycles_per_mma": round(ms2 * 1e-3 * 1.965e9 / (reps * 8), 1)}), flush=True)
