"""A/B timing of GEMM builds on the in-step shapes (with bias / scale / residual), one gpurun call."""
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
def mk(M,N,Kd,amn=False,bmn=False):
    A=(torch.randn(Kd,M,device=dev) if amn else torch.randn(M,Kd,device=dev)).bfloat16()
    B=(torch.randn(Kd,N,device=dev) if bmn else torch.randn(N,Kd,device=dev)).bfloat16()
    return A,B
A,B=mk(R,2304,768); bias=torch.randn(2304,device=dev)
out.append('qkv+bias+scale %.3f' % t(lambda: K.gemm(A,B,R,2304,768,epilogue=K.EPI_BF16,bias=bias,scale=0.125,scale_cols=768)))
out.append('qkv nobias %.3f' % t(lambda: K.gemm(A,B,R,2304,768,epilogue=K.EPI_BF16)))
A,B=mk(R,768,768); bias=torch.randn(768,device=dev); res=torch.randn(R,768,device=dev); o=torch.empty(R,768,device=dev)
out.append('out_proj resid %.3f' % t(lambda: K.gemm(A,B,R,768,768,epilogue=K.EPI_RESID,bias=bias,res=res,out=o)))
A,B=mk(768,3072,R,True,True); g=torch.zeros(768,3072,device=dev)
out.append('wgrad fc2 %.3f' % t(lambda: K.wgrad(A,B,768,3072,R,out=g)))
A,B=mk(768,3072,1280,True,True)
out.append('wgrad fc2 text %.3f' % t(lambda: K.wgrad(A,B,768,3072,1280,out=g)))
print(' | '.join(out))
"""
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["SIMVGB_LIB"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("%-24s %s" % (lib, r.stdout.strip() or r.stderr[-600:]), flush=True)
