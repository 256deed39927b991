import torch
n=8140687
dev='cuda'
a=torch.randn(n,dtype=torch.float64,device=dev); b=torch.randn_like(a); c=torch.randn_like(a); out=torch.empty_like(a)
out2=torch.empty(2*n,dtype=torch.float64,device=dev)
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
def t(fn,nrep=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot=0
    for _ in range(nrep):
        flush.zero_()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); tot+=e0.elapsed_time(e1)
    return tot/nrep*1e3
print('addcmul 3r1w us', t(lambda: torch.addcmul(a,b,c,out=out)), 'bytes', 4*8*n/1e6)
print('copy 1r1w us', t(lambda: out.copy_(a)))
print('mul 1r 2w us', t(lambda: torch.mul(a.unsqueeze(0).expand(2,n), 0.5, out=out2.view(2,n))))
print('fill 1w us', t(lambda: out.fill_(1.0)))
print('sum 1r us', t(lambda: a.sum()))
