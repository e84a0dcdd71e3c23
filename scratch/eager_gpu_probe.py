import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, traceback
try:
    print(bench.eager_gpu_arm(4096, "cuda:0"))
except Exception:
    traceback.print_exc()
