import torch
B, M = 65536, 1024
K = torch.randn(B, M, dtype=torch.float64, device="cuda"); C = torch.randn(M, M, dtype=torch.float64, device="cuda")
T = torch.empty_like(K)
for _ in range(3): torch.matmul(K, C, out=T)
torch.cuda.synchronize()
Kc = torch.ones_like(K); Cc = torch.ones_like(C)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nonstationary_precip_b200 import ops
def t(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); fn(); b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)
print("ours random", t(lambda: ops.rowquad(K, C, T=T, need_q=False)))
print("ours ones  ", t(lambda: ops.rowquad(Kc, Cc, T=T, need_q=False)))
print("ours zeros ", t(lambda: ops.rowquad(torch.zeros_like(K), torch.zeros_like(C), T=T, need_q=False)))
print("cublas rand", t(lambda: torch.matmul(K, C, out=T)))
