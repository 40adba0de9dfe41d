import sys, os, torch
sys.path.insert(0, os.getcwd())
import vsn_b200
from vsn_b200 import ops
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(4_000_000)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters
for rows,C in [(16128,384),(54432,192),(2016,768),(435456,96)]:
    x=torch.randn(rows,C,device='cuda'); g=torch.ones(C,device='cuda'); b=torch.zeros(C,device='cuda')
    y,m,r=ops.layernorm_fwd(x,g,b)
    dy=torch.randn(rows,C,device='cuda').bfloat16(); rg=torch.randn(rows,C,device='cuda')
    dg=torch.zeros(C,device='cuda'); db=torch.zeros(C,device='cuda')
    ms=timeit(lambda: ops.layernorm_bwd(dy,x,m,r,g,resid_grad=rg,want_bf16=True,dgamma=dg,dbeta=db))
    msf=timeit(lambda: ops.layernorm_fwd(x,g,b))
    print(f"rows{rows} C{C}: bwd {ms*1e3:.1f} us  fwd {msf*1e3:.1f} us",flush=True)
