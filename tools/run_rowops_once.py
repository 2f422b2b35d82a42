"""One launch of each row kernel at the cfg2 problem size (vision expert: 64 x 1601 rows), for ncu captures.  Tooling."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simvg_b200 import kernels as K
R, D, F = 64 * 1601, 768, 3072
dev = "cuda"
torch.manual_seed(0)
u = torch.randn(R, F, device=dev).bfloat16()
g3, b3 = torch.randn(F, device=dev), torch.randn(F, device=dev)
f, mf, rf = K.ln_fwd(u, g3, b3, 1e-5, gelu=True)                                   # ln_fwd_warp_kernel<bf16, bf16, 12, 1>
dg, db, dbias = (torch.zeros(F, device=dev) for _ in range(3))
du = torch.empty(R, F, device=dev, dtype=torch.bfloat16)
K.ln_bwd(2, None, torch.randn(R, F, device=dev).bfloat16(), g3, mf, rf, dg, db, dx=du, u=u, dbias_prev=dbias)   # ln_bwd_wide_kernel<3>
x = torch.randn(R, D, device=dev)
g1, b1 = torch.randn(D, device=dev), torch.randn(D, device=dev)
h, m1, r1 = K.ln_fwd(x, g1, b1, 1e-5)                                              # ln_fwd_warp_kernel<float, bf16, 3, 0>
dres = torch.randn(R, D, device=dev)
dyb = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
dg1, db1, dbp = (torch.zeros(D, device=dev) for _ in range(3))
K.ln_bwd(0, x, torch.randn(R, D, device=dev).bfloat16(), g1, m1, r1, dg1, db1, dres_in=dres, dres_out=dres, dyb=dyb, rows_per_scale=1601,
         dbias_prev=dbp)                                                           # ln_bwd_warp_kernel<0, 0, 3>
torch.cuda.synchronize()
print("row kernels launched")
