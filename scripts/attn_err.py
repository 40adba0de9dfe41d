import sys, os, torch
sys.path.insert(0, os.getcwd())
from tests.test_kernels_gpu import _ops, _bf, _rand, _ref_window_attention
from tests.helpers import rel_err
ops=_ops()
for grid,heads,shifted in [((12,14,12),2,False),((12,14,12),2,True),((18,21,12),1,True)]:
    B, window, shift, hd = 2, (6,7,6), (3,3,3), 32
    C=heads*hd; T=B*grid[0]*grid[1]*grid[2]
    qkv=_bf(_rand(T,3*C,seed=1)); table=0.5*_rand(11*13*11,heads,seed=2)
    geom=ops.WindowGeom(B,grid,window,shift if shifted else (0,0,0),shifted)
    out,lse=ops.attn_fwd(qkv,heads,hd,S=geom.S,N=geom.N,scale=hd**-0.5,geom=geom,table=table)
    q32=qkv.float().requires_grad_(True); t32=table.clone().requires_grad_(True)
    ref=_ref_window_attention(q32,t32,B,grid,window,shift,shifted,heads,hd)
    dout=_bf(_rand(T,C,seed=3))
    ref.backward(dout.float())
    dt=torch.zeros_like(table)
    dqkv=ops.attn_bwd(qkv,out,dout,lse,heads,hd,S=geom.S,N=geom.N,scale=hd**-0.5,geom=geom,table=table,dtable=dt)
    e_out=rel_err(out.float(),ref.detach()); e_dq=rel_err(dqkv.float(),q32.grad); e_dt=rel_err(dt,t32.grad)
    print(f"grid {grid} shifted {shifted}: out {e_out:.3e} dqkv {e_dq:.3e} dtable {e_dt:.3e}  max|out-ref| {float((out.float()-ref.detach()).abs().max()):.3e}",flush=True)
