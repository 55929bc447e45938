import torch, time
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n*1e-3
N=4*1024**3
x=torch.empty(N,dtype=torch.uint8,device='cuda'); y=torch.empty(N,dtype=torch.uint8,device='cuda')
s=t(lambda: x.fill_(1)); print('fill 4GiB  %.1f GB/s'%(N/s/1e9))
s=t(lambda: x.zero_()); print('zero 4GiB  %.1f GB/s'%(N/s/1e9))
s=t(lambda: y.copy_(x)); print('copy 4GiB  %.1f GB/s (r+w)'%(2*N/s/1e9))
xi=x.view(torch.int32)
s=t(lambda: xi.sum()); print('read 4GiB  %.1f GB/s'%(N/s/1e9))
# mixed: read 0.7GB + write 4.8GB pattern emulation: copy 0.7 then fill 4.1
a=x[:700*10**6]; b=y[:700*10**6]; c=y[700*10**6:700*10**6+4100*10**6]
def mix():
    b.copy_(a); c.fill_(3)
s=t(mix); print('mix (0.7r+0.7w+4.1w) %.1f GB/s'%((0.7e9*2+4.1e9)/s/1e9))
