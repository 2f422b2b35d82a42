import ctypes, os, sys
import torch
sys.path.insert(0, '.')
from simvg_b200 import _lib as L
from tools.gpu_check_attn import run
lib = L.lib()
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib.simvgb_debug_attn_trace(ctypes.c_void_p(buf.data_ptr()))
run(64, 12, 1601, 20, [i % 14 for i in range(64)], check=False, iters=0, tag="trace")
torch.cuda.synchronize()
t = buf.view(64, 16).cpu()
print("pair | MMA: issueS(i+1)  wait_p_full  issue dK,dP,dV  wait_dq_empty  issue dQ | CMP: wait_s_full  P-phase  wait_dp_full  dS-phase | period")
prev = None
for i in range(13):
    m = t[i, 0:6].tolist(); c = t[i, 8:13].tolist()
    per = (m[0] - prev) if prev else 0
    prev = m[0]
    print("%2d | %6d %6d %6d %6d %6d | %6d %6d %6d %6d | %6d" % (
        i, m[1]-m[0], m[2]-m[1], m[3]-m[2], m[4]-m[3], m[5]-m[4], c[1]-c[0], c[2]-c[1], c[3]-c[2], c[4]-c[3], per))
