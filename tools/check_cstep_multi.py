#!/usr/bin/env python3
"""Multi-GPU check of the C step (run under torchrun, one rank per GPU): G ranks take B/G rows each and run npgp_svgp_step
on an npgp communicator (early all-reduce of (m, Ls) + grouped all-reduce of the rest); every rank also computes the
full-batch step alone.  The all-reduced flat gradient / loss and the updated parameters must agree with the single-rank
values.    torchrun --nproc-per-node 2 tools/check_cstep_multi.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from svgp_cases import make_problem  # noqa: E402
from nonstationary_precip_b200.comm import NpgpComm  # noqa: E402
from nonstationary_precip_b200.svgp import SVGPGibbs  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
comm = NpgpComm(rank, world, dev)
out = {}
for variant in ("full", "diag"):
    B, M = 2048, 256
    x, y, Z, p, N = make_problem(variant, B=B, M=M, d=3, seed=11, device=dev)
    single = SVGPGibbs(variant, Z, N, **p).use_c_engine()
    multi = SVGPGibbs(variant, Z, N, **p).use_c_engine()
    graph = SVGPGibbs(variant, Z, N, **p).use_c_engine()
    Bl = B // world
    xs, ys = x[rank * Bl:(rank + 1) * Bl].contiguous(), y[rank * Bl:(rank + 1) * Bl].contiguous()
    graph.capture(Bl, world, B, lr=0.01, all_reduce=comm)
    worst = 0.0
    for step in range(3):
        l1 = single.train_step(x, y, lr=0.01)
        lm = multi.train_step(xs, ys, lr=0.01, world_size=world, B_global=B, all_reduce=comm)
        lg = graph.train_step_graph(xs, ys)
        torch.cuda.synchronize()
        rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
        worst = max(worst, rel(multi.grad, single.grad), rel(graph.grad, single.grad), abs(lm.item() - l1.item()) / abs(l1.item()))
    worst = max(worst, rel(multi.theta, single.theta), rel(graph.theta, single.theta))
    t = torch.tensor([worst], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[variant] = t.item()
if rank == 0:
    out["world"] = world
    out["ok"] = bool(max(out["full"], out["diag"]) < 1e-6)  # 3 Adam steps amplify the FP64-atomics noise of the gradient sums
    print(json.dumps(out))
dist.barrier()
os._exit(0)
