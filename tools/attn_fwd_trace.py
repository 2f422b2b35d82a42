import ctypes, sys
import torch
sys.path.insert(0, '.')
from simvg_b200 import _lib as L
from tools.gpu_check_attn import run
lib = L.lib()
buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
lib.simvgb_debug_attn_fwd_trace(ctypes.c_void_p(buf.data_ptr()))
run(64, 12, 1601, 20, [i % 14 for i in range(64)], check=False, iters=0, tag="trace")
torch.cuda.synchronize()
t = buf.view(64, 8).cpu()
print("tile | wait s_empty | issue S(j+1) incl K wait | wait p_full | wait V | issue PV | period")
prev = None
for j in range(13):
    c = t[j].tolist()
    per = c[0] - prev if prev else 0
    prev = c[0]
    print("%2d | %6d %6d %6d %6d %6d | %6d" % (j, c[1]-c[0], c[2]-c[1], c[3]-c[2], c[4]-c[3], c[5]-c[4], per))
