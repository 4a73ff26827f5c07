#!/usr/bin/env python3
"""Do the int8 tcgen05 tensor pipe and the FP64 (DFMA) pipe run concurrently on one SM?  The int8 peak probe (one 192-thread
CTA per SM, operands resident in shared memory) and the DFMA peak probe (here 1 or 2 CTAs of 256 threads per SM) alone and
together on two streams; also against an FP32 FMA stream (torch elementwise) as a control.
  python tools/pipe_overlap_i8.py > gpurun_out/pipe_overlap_i8.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nonstationary_precip_b200._lib import check, lib, ptr  # noqa: E402

out = torch.zeros(8, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(i8_reps=0, f64_blocks=0, f64_iters=0, i8_first=True):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    if i8_reps:
        check(lib().npgp_i8_peak_probe(64, 1, 148, i8_reps, s1.cuda_stream), "i8")
    if f64_blocks:
        check(lib().npgp_fp64_peak_probe(0, f64_blocks, f64_iters, ptr(out), s2.cuda_stream), "f64")
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


REPS = 8192
for cta_per_sm in (1, 2, 4):
    blocks, iters = 148 * cta_per_sm, 16384 // cta_per_sm
    for _ in range(2):
        run(REPS)
        run(0, blocks, iters)
    ti, tf = run(REPS), run(0, blocks, iters)
    tb = run(REPS, blocks, iters)
    print("DFMA CTAs/SM %d: int8 alone %.3f ms, dfma alone %.3f ms, together %.3f ms (sum %.3f, max %.3f) -> %s" % (
        cta_per_sm, ti, tf, tb, ti + tf, max(ti, tf), "concurrent" if tb < 0.75 * (ti + tf) else "serialised"))
