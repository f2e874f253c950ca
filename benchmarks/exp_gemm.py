import sys, os, json, torch
sys.path.insert(0, ".")
from csts_b200 import kernels as K
from benchmarks.micro_gemm import timeit
dev="cuda"
for (M,Kd,N) in [(8192,384,1536),(32768,768,1536),(131072,192,576),(8192,1536,384)]:
    byts = 2.0*(M*Kd+N*Kd+M*N); sets=max(2,int(300e6//byts)+1)
    As=[torch.randn(M,Kd,device=dev).to(torch.bfloat16) for _ in range(sets)]
    Bs=[(torch.randn(N,Kd,device=dev)*0.05).to(torch.bfloat16) for _ in range(sets)]
    outs=[torch.empty(M,N,dtype=torch.bfloat16,device=dev) for _ in range(sets)]
    bias=torch.randn(N,device=dev)
    row={"shape":(M,Kd,N)}
    for use_bias in (True, False):
        t=timeit(lambda i:(lambda:K.gemm(As[i],Bs[i],M=M,N=N,K=Kd,bias=bias if use_bias else None,out=outs[i],backend=2)),sets)
        row["bias" if use_bias else "nobias"]=round(t*1e3,1)
    print(json.dumps(row), flush=True)
