"""Launch one pool shape a few times (for ncu captures): python benchmarks/one_pool.py heads d T H W st sh sw transposed"""
import sys
import torch
sys.path.insert(0, ".")
from csts_b200 import kernels as K
h, d, T, H, W, st, sh, sw, tr = (int(x) for x in sys.argv[1:10])
B, dev = 8, "cuda"
N, Cn = T * H * W, h * d
qkv = torch.randn(B, N, 3, h, d, device=dev).to(torch.bfloat16)
w = torch.randn(d, 1, 3, 3, 3, device=dev) * 0.2
gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
qs = (N * 3 * Cn, d, 3 * Cn)
for _ in range(3):
    out, pre, mean, rstd, thw_o = K.dwconv(qkv, qs, Cn, B, h, d, (T, H, W), (st, sh, sw), w, transposed=bool(tr), norm=(gamma, beta))
torch.cuda.synchronize()
