"""A/B timing of attention builds inside ONE gpurun call (box-to-box clocks differ by ~15%)."""
import os, subprocess, sys
libs = sys.argv[1:]
for rep in range(2):
    for lib in libs:
        env = dict(os.environ)
        if lib != "default":
            env["SIMVGB_LIB"] = os.path.abspath(lib)
        out = subprocess.run([sys.executable, "-c", """
import sys
sys.path.insert(0,'.')
from tools.gpu_check_attn import run
run(64, 12, 1601, 20, [i % 14 for i in range(64)], check=False, iters=5, tag='ab')
"""], env=env, capture_output=True, text=True)
        lines = [l.replace("B=64 H=12 Lv=1601 Lt=20: ", "") for l in out.stdout.splitlines() if l.startswith("time")]
        print("%-28s %s" % (lib, " | ".join(lines) if lines else out.stderr[-400:]), flush=True)
