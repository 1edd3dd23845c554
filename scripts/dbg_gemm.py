import sys, os, torch
sys.path.insert(0, '/root/repo')
from cap4d_b200 import ops
dev=torch.device('cuda:0')
for (M,N,K) in [(65536,320,320),(65536,960,320),(65536,320,2880)]:
    a=torch.randn(M,K,device=dev).to(torch.bfloat16); w=(torch.randn(N,K,device=dev)/K**.5).to(torch.bfloat16)
    bias=torch.randn(N,device=dev); r=torch.randn(M,N,device=dev)
    for name,mode,res in [("f32",0,None),("f32 nostore",16,None),("f32 noepi",32,None),("f32+res",0,r),("f32+res nostore",16,r),("bf16",1,None)]:
        _,ms=ops.gemm(a,w,bias=bias,residual=res,out_mode=mode,time_iters=10)
        print(f"M={M} N={N} K={K} {name:16s} {ms*1e3:7.1f} us")
