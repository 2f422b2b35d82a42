import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_check_attn import run
B = int(os.environ.get("B", 16))
run(B, 12, 1601, 20, [i % 14 for i in range(B)], check=False, iters=0, tag="ncu")
