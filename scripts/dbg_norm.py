import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import ops
dev=torch.device('cuda:0')
for (n_img,hw,C1,C2) in [(16,4096,320,0),(16,4096,640,320),(16,1024,640,0),(16,256,1280,0),(16,64,1280,1280)]:
    C=C1+C2
    x1=torch.randn(n_img*hw,C1,device=dev); x2=torch.randn(n_img*hw,C2,device=dev) if C2 else None
    g=torch.ones(C,device=dev); b=torch.zeros(C,device=dev)
    for silu in (True,False):
        for raw in (False,True):
            r=ops.groupnorm(x1,x2,n_img,hw,g,b,1e-5,silu,want_raw=raw,time_iters=20)
            ms=r[-1]
            bytes_=n_img*hw*C*(4+2+(2 if raw else 0))
            print(f"GN n={n_img} hw={hw} C={C1}+{C2} silu={int(silu)} raw={int(raw)}: {ms*1e3:7.1f} us  {bytes_/ms/1e6:7.1f} GB/s (algorithmic)")
for (M,C) in [(65536,320),(16384,640),(4096,1280)]:
    x=torch.randn(M,C,device=dev); g=torch.ones(C,device=dev); b=torch.zeros(C,device=dev)
    _,ms=ops.layernorm(x,g,b,1e-5,time_iters=20)
    print(f"LN M={M} C={C}: {ms*1e3:7.1f} us {M*C*6/ms/1e6:7.1f} GB/s")
