"""N > 1 path on CPU: two gloo ranks shard the minibatch rows, all-reduce the flat gradient buffer once per step and
must reproduce the single-rank training trajectory.  Runs the product's svgp.py through the CPU ops
emulator (tests/cpu_ops_emulator.py)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, variant, out_path):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_ops_emulator as emu
    from nonstationary_precip_b200.svgp import SVGPGibbs
    from svgp_cases import make_problem
    x, y, Z, kw, N = make_problem(variant, B=64, M=16, d=3, seed=21)
    model = SVGPGibbs(variant, Z, N, ops=emu, **kw)
    Bl = x.shape[0] // world
    xs, ys = x[rank * Bl:(rank + 1) * Bl].contiguous(), y[rank * Bl:(rank + 1) * Bl].contiguous()
    losses = []
    for _ in range(4):
        loss = model.train_step(xs, ys, lr=0.01, world_size=world, B_global=x.shape[0],
                                all_reduce=lambda t: dist.all_reduce(t))
        losses.append(loss.item())
    if rank == 0:
        torch.save(dict(losses=losses, theta=model.theta.clone()), out_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("variant", ["diag", "full"])
def test_two_ranks_match_one_rank(tmp_path, variant):
    sys.path.insert(0, HERE)
    import cpu_ops_emulator as emu
    from nonstationary_precip_b200.svgp import SVGPGibbs
    from svgp_cases import make_problem
    torch.set_default_dtype(torch.float64)
    out = str(tmp_path / "r0.pt")
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, variant, out), nprocs=2, join=True)
    got = torch.load(out)
    x, y, Z, kw, N = make_problem(variant, B=64, M=16, d=3, seed=21)
    model = SVGPGibbs(variant, Z, N, ops=emu, **kw)
    want = [model.train_step(x, y, lr=0.01).item() for _ in range(4)]
    for a, b in zip(got["losses"], want):
        assert abs(a - b) < 1e-12 * max(1.0, abs(b))
    assert (got["theta"] - model.theta).abs().max() < 1e-10
