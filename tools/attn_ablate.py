import os, subprocess, sys
for flag in (0, 1, 2, 4, 8, 16, 6, 14, 31):
    env = dict(os.environ, SIMVGB_ATTN_DEBUG=str(flag), ABL="1")
    out = subprocess.run([sys.executable, "-c", """
import os,sys
sys.path.insert(0,'.')
from tools.gpu_check_attn import run
r = run(32, 12, 1601, 20, [i % 14 for i in range(32)], check=False, iters=5, tag='abl')
"""], env=env, capture_output=True, text=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith("time " + (sys.argv[1] if len(sys.argv) > 1 else "bwd"))]
    print("dbg=%2d  %s" % (flag, lines[0] if lines else out.stderr[-300:]), flush=True)
