"""One plain 8192^3 bf16 GEMM through mmg_gemm and through torch.matmul (cuBLAS) -- for an ncu A/B of the mainloop."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import ops
M = N = K = 8192
A = torch.randn(M, K, device="cuda").bfloat16()
B = torch.randn(N, K, device="cuda").bfloat16()
C = torch.empty(M, N, device="cuda")
for _ in range(2):
    ops.gemm(A, B, M, N, K, prec="bf16", out=C)
    torch.matmul(A, B.t())
torch.cuda.synchronize()
