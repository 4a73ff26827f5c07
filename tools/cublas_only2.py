import torch
B, M = 65536, 1024
K = torch.randn(B, M, dtype=torch.float64, device="cuda"); C = torch.randn(M, M, dtype=torch.float64, device="cuda")
T = torch.empty_like(K)
torch.cuda.synchronize()
torch.cuda.profiler.start()
torch.matmul(K, C, out=T)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
