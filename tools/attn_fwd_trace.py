"""clock64 trace of the forward attention kernel's softmax warps (CTA 0, first work item).  Needs a -DSIMVGB_FWD_TRACE build:
    SIMVGB_EXTRA_FLAGS=-DSIMVGB_FWD_TRACE SIMVGB_OUT=simvg_b200/libsimvg_b200_trace.so python -m simvg_b200.build
    SIMVGB_LIB=simvg_b200/libsimvg_b200_trace.so python tools/attn_fwd_trace.py"""
import ctypes, sys
import torch
sys.path.insert(0, '.')
from simvg_b200 import _lib as L
from tools.gpu_check_attn import run
lib = L.lib()
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
lib.simvgb_debug_attn_fwd_trace(ctypes.c_void_p(buf.data_ptr()))
run(64, 12, 1601, 20, [i % 14 for i in range(64)], check=False, iters=0, tag="trace")
torch.cuda.synchronize()
t = buf.view(2, 32, 8).cpu()
for x in range(2):
    print("tile %s: j | wait S | LDTM+release | mask+max+xchg | exp chunk 0 | wait PV(j-1) | rescale+st+exp1+arrive | period (start offset vs tile A)" % "AB"[x])
    prev = None
    for j in range(13):
        c = t[x, j].tolist()
        per = c[0] - prev if prev else 0
        prev = c[0]
        print("%2d | %6d %6d %6d %6d %6d %6d | %6d (%d)" % (j, c[1]-c[0], c[2]-c[1], c[3]-c[2], c[4]-c[3], c[5]-c[4], c[6]-c[5], per, c[0] - t[0, j, 0].item()))
