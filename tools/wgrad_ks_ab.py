"""Split-K choice for the weight-gradient GEMMs: times each in-step shape with a list of candidate k_splits (one short gpurun call)."""
import sys, torch
sys.path.insert(0, '.')
from simvg_b200 import kernels as K
dev = 'cuda'
def t(fn, n=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
R = 64 * 1601
dY = {n: torch.randn(R, n, device=dev).bfloat16() for n in (768, 2304, 3072)}
for (M, N, cands) in ((768, 3072, (5, 2, 4)), (3072, 768, (5, 2, 4)), (2304, 768, (6, 8, 5)), (768, 768, (17, 8, 16))):
    g = torch.zeros(M, N, device=dev)
    out = []
    for ks in cands:
        out.append("ks=%d %.3f" % (ks, t(lambda: K.gemm(dY[M], dY[N], M, N, R, a_mn=True, b_mn=True, epilogue=K.EPI_ATOMIC, out=g, k_splits=ks))))
    print("wgrad %dx%d: %s" % (M, N, " | ".join(out)), flush=True)
