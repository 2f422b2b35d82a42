"""A/B timing of the LayerNorm kernels inside one gpurun call."""
import os, subprocess, sys
code = """
import sys, torch
sys.path.insert(0,'.')
from simvg_b200 import kernels as K
dev='cuda'
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
R=64*1601
out=[]
for C in (768, 3072):
    x32=torch.randn(R,C,device=dev); xb=x32.bfloat16(); dy=torch.randn(R,C,device=dev).bfloat16()
    g=torch.randn(C,device=dev); b=torch.randn(C,device=dev)
    _,mean,rstd=K.ln_fwd(x32,g,b,1e-5)
    dg=torch.zeros(C,device=dev); db=torch.zeros(C,device=dev); dbp=torch.zeros(C,device=dev)
    dres=torch.randn(R,C,device=dev); dyb=torch.empty(R,C,device=dev,dtype=torch.bfloat16); dx=torch.empty(R,C,device=dev,dtype=torch.bfloat16)
    try:
        tg = t(lambda: K.ln_fwd(xb,g,b,1e-5,gelu=True))
    except TypeError:
        tg = float('nan')
    out.append('C=%d fwd f32->bf16 %.3f  fwd bf16->bf16 %.3f  +gelu %.3f' % (C, t(lambda: K.ln_fwd(x32,g,b,1e-5)), t(lambda: K.ln_fwd(xb,g,b,1e-5)), tg))
    out.append('C=%d bwd mode0 %.3f  mode1 %.3f  mode2 %.3f' % (C,
        t(lambda: K.ln_bwd(0,x32,dy,g,mean,rstd,dg,db,dres_in=dres,dres_out=dres,dyb=dyb,dbias_prev=dbp)),
        t(lambda: K.ln_bwd(1,xb,dy,g,mean,rstd,dg,db,dx=dx)),
        t(lambda: K.ln_bwd(2,xb,dy,g,mean,rstd,dg,db,dx=dx,u=xb,dbias_prev=dbp))))
print(' | '.join(out))
"""
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["SIMVGB_LIB"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("%-24s %s" % (lib, r.stdout.strip() or r.stderr[-500:]), flush=True)
