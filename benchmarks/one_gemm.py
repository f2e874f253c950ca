"""Launch one GEMM shape a few times (for `ncu --set full` captures)."""
import sys
import torch
sys.path.insert(0, ".")
from csts_b200 import kernels as K
M, Kd, N = (int(x) for x in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "fwd"
dev = "cuda"
A = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
B = (torch.randn(N, Kd, device=dev) * 0.05).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
for _ in range(4):
    if mode == "fwd":
        K.gemm(A, B, M=M, N=N, K=Kd, bias=bias, out=out, backend=2)
    else:
        dW = torch.empty(N, Kd, dtype=torch.float32, device=dev)
        K.gemm(out, A, M=N, N=Kd, K=M, a_kmajor=False, b_kmajor=False, out=dW, split_k=32, backend=2)
torch.cuda.synchronize()
