import torch,time
x=torch.empty(4*1024**3//8,dtype=torch.float64,device='cuda')
y=torch.empty_like(x)
def t(f,n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n
ms=t(lambda: x.fill_(1.5)); print("fill 4GB", 4.295/ms*1e3,"GB/s")
ms=t(lambda: x.zero_()); print("zero 4GB", 4.295/ms*1e3,"GB/s")
ms=t(lambda: y.copy_(x)); print("copy 4+4GB", 8.59/ms*1e3,"GB/s")
ms=t(lambda: x.sum()); print("read 4GB", 4.295/ms*1e3,"GB/s")
