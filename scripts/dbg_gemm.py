import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops
dev=torch.device('cuda:0')
for (M,N,K) in [(65536,2560,320),(16384,5120,640),(65536,320,1280),(65536,960,320)]:
    a=torch.randn(M,K,device=dev).to(torch.bfloat16); w=(torch.randn(N,K,device=dev)/K**.5).to(torch.bfloat16)
    bias=torch.randn(N,device=dev)
    for name,mode in [("geglu",2),("geglu noepi",2|32),("bf16",1),("bf16 noepi",1|32),("f32",0)]:
        if N==320 and "geglu" in name: continue
        _,ms=ops.gemm(a,w,bias=bias,out_mode=mode,time_iters=10)
        print(f"M={M} N={N} K={K} {name:14s} {ms*1e3:7.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TF/s")
