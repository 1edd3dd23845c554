import torch, time
dev=torch.device('cuda:0')
n=1<<28  # 1 GiB of fp32
a=torch.empty(n,device=dev); b=torch.empty(n,device=dev)
def t(fn,bytes_,name,it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/it
    print(f"{name:28s} {bytes_/ms/1e6:8.1f} GB/s")
t(lambda: a.fill_(1.0), n*4, "write only (fill)")
t(lambda: a.sum(), n*4, "read only (sum)")
t(lambda: b.copy_(a), n*8, "copy (read+write)")
t(lambda: a.mul_(1.0001), n*8, "in-place scale (r+w same)")
h=torch.empty(n//2,device=dev,dtype=torch.bfloat16)
t(lambda: h.copy_(a[:n//2]), n//2*6, "f32->bf16 cast (4r+2w)")
