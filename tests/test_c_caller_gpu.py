"""A C program (tools/c_demo/svgp_step_demo.c: gcc + libcudart + libnpgp.so, no Python, no torch) drives the whole SVGP-Gibbs
ELBO step through include/npgp.h -- npgp_svgp_plan_create / npgp_svgp_step -- and must reproduce the Python-side engine on
the same problem.  This is the drop-in boundary exercised the way a non-Python host would use it."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from svgp_cases import make_problem

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_demo(tmp_path):
    exe = str(tmp_path / "svgp_step_demo")
    libdir = os.path.join(ROOT, "nonstationary_precip_b200")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tools", "c_demo", "svgp_step_demo.c"), "-o", exe, "-L", libdir, "-lnpgp",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.mark.parametrize("variant", ["full", "diag"])
def test_c_program_runs_the_step_and_matches_the_python_engine(variant, tmp_path):
    from nonstationary_precip_b200.svgp import SVGPGibbs
    exe = build_demo(tmp_path)
    B, M, d, iters = 640, 128, 3, 3
    x, y, Z, p, N = make_problem(variant, B=B, M=M, d=d, seed=21, device="cuda")
    model = SVGPGibbs(variant, Z, N, **p).use_c_engine()
    consts = ([model.row_os.reshape(-1), model.row_lam.reshape(-1)] if variant == "full" else
              [model.prior_c.reshape(-1), model.prior_os.reshape(-1), model.prior_lam.reshape(-1)])
    blob = torch.cat([t.detach().reshape(-1).cpu() for t in [x, y, model.theta, model.mask] + consts]).numpy()
    prob, out = str(tmp_path / "problem.bin"), str(tmp_path / "out.bin")
    with open(prob, "wb") as f:
        f.write(struct.pack("8i", 1 if variant == "full" else 0, d, M, B, N, 1, 1, iters))
        f.write(blob.astype(np.float64).tobytes())
    r = subprocess.run([exe, prob, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    got = np.fromfile(out, dtype=np.float64)
    n_pad = model.theta.numel()
    assert got.size == iters + 2 * n_pad + 2
    losses, grad, theta = got[:iters], got[iters:iters + n_pad + 2], got[iters + n_pad + 2:]
    want = [model.train_step(x, y, lr=0.01).item() for _ in range(iters)]
    for a, b in zip(losses, want):
        assert abs(a - b) < 1e-9 * abs(b)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert rel(grad, model.grad.cpu().numpy()) < 1e-7
    assert rel(theta, model.theta.cpu().numpy()) < 1e-8
