#!/usr/bin/env python3
"""Are DMMA (tensor pipe) and DFMA (fp64 pipe) separate execution resources?  Run the two peak probes alone and
concurrently on two streams (both co-resident: 148*4 CTAs of 256 threads each)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nonstationary_precip_b200._lib import lib, ptr, check
out = torch.zeros(8, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
blocks = 148 * 4
def run(mode_a, it_a, mode_b=None, it_b=0):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    check(lib().npgp_fp64_peak_probe(mode_a, blocks, it_a, ptr(out), s1.cuda_stream), "a")
    if mode_b is not None:
        check(lib().npgp_fp64_peak_probe(mode_b, blocks, it_b, ptr(out), s2.cuda_stream), "b")
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
IT_F, IT_M = 8192, 1024   # DFMA: 16*2 flop/iter/thread; DMMA: 16*512 flop/iter/warp -> similar durations
for _ in range(2): run(0, IT_F); run(1, IT_M)
tf, tm = run(0, IT_F), run(1, IT_M)
tboth = run(0, IT_F, 1, IT_M)
tff = run(0, IT_F, 0, IT_F)
tmm = run(1, IT_M, 1, IT_M)
print("dfma alone %.3f ms, dmma alone %.3f ms, dfma||dmma %.3f ms, dfma||dfma %.3f, dmma||dmma %.3f" % (tf, tm, tboth, tff, tmm))
print("separate pipes" if tboth < 0.75 * (tf + tm) else "shared pipe")
