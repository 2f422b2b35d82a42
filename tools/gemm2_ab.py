"""2-CTA vs 1-CTA GEMM: parity tests under SIMVGB_GEMM_2CTA=1 (bounded by timeout), then in-step shape timings for both."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
env1 = dict(os.environ, SIMVGB_GEMM_2CTA="1")
r = subprocess.run(["timeout", "300", sys.executable, "-m", "pytest", "tests/test_gpu_kernels.py", "-m", "gpu", "-x", "-q", "-k", "gemm or wgrad or linear"],
                   env=env1, capture_output=True, text=True)
print(r.stdout[-1500:], r.stderr[-800:], flush=True)
for v in ("0", "1"):
    env = dict(os.environ, SIMVGB_GEMM_2CTA=v)
    r = subprocess.run(["timeout", "300", sys.executable, os.path.join(here, "gemm_ab.py"), "default"], env=env, capture_output=True, text=True)
    print("2CTA=%s %s %s" % (v, r.stdout.strip(), r.stderr[-500:]), flush=True)
