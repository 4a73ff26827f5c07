#!/usr/bin/env python3
"""Config 4 (2-layer DGP, DSVI) data-parallel over likelihood samples: one process per GPU (torchrun), S/G samples per
rank, one NCCL all-reduce of the flat gradient per step.  Prints the step time (max over ranks, CUDA events) and the
loss of the first step, which must not depend on the number of ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dgp_multi_gpu.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nonstationary_precip_b200.models import dgps
    B, S, M = int(os.environ.get("B", 65536)), int(os.environ.get("S", 32)), int(os.environ.get("M", 512))
    steps = int(os.environ.get("STEPS", 5))
    torch.manual_seed(4)  # same initial parameters on every rank
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(B, 3, generator=g, dtype=torch.float64) * 2 - 1).cuda()
    y = (torch.sin(3 * x[:, 0]) + 0.5 * torch.cos(5 * x[:, 1] * x[:, 2])).contiguous()
    model = dgps.DeepGP(1, x.shape, num_inducing=M).cuda().double()
    mll = dgps.DeepApproximateMLL(dgps.VariationalELBO(model.likelihood, model, 1 << 20))
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    reduce_ = (lambda t: dist.all_reduce(t)) if world > 1 else (lambda t: t)

    def step(k):
        opt.zero_grad(set_to_none=True)
        with dgps.num_likelihood_samples(S), dgps.sample_shard(rank, world):
            loss = -mll(model(x, seed=1000 + k), y)
        loss.backward()
        dgps.allreduce_gradients(model, reduce_)
        opt.step()
        return loss.detach()

    first = step(0).clone()
    if world > 1:
        dist.all_reduce(first)
    step(1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(steps):
        step(2 + k)
    b.record()
    b.synchronize()
    ms = torch.tensor(a.elapsed_time(b) / steps, device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "c4 2-layer DGP DSVI, B=%d, M=%d/layer, S=%d sharded over %d rank(s)" % (B, M, S, world),
                          "n_gpus": world, "ms_per_step": ms.item(), "steps_per_s": 1e3 / ms.item(),
                          "first_step_loss_sum_over_ranks": first.item()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
