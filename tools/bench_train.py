"""The cfg3-shaped training iteration of bench.py alone (A/B: FENERF_B200_BWD_GEMM=cublas|tcgen05)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench

class A: pass
args = A(); args.precision = "guard"
out = bench.measure_train_step(args, 1, 0, 0)
print(os.environ.get("FENERF_B200_BWD_GEMM", "tcgen05"), json.dumps({k: out[k] for k in ("ms_per_iteration", "rendered_faces_per_s")}))
