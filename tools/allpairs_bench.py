"""All-pairs scoring throughput (BASELINE configs[3]: 162541 x 62423, dim 128): tcgen05 path with the fused top-1
consumer, CUDA events.  Usage (GPU box): python tools/allpairs_bench.py [users items dim]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tf_recomm_b200  # noqa: E402,F401
from tf_recomm_b200.engine import SvdEngine  # noqa: E402

U, I, d = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (162541, 62423, 128)
eng = SvdEngine(U, I, d, 1e-3, 0.05, device_init_seed=3)
for tc in (True, False):
    if not tc and U * I > 3e9:
        reps = 1
    else:
        reps = 5
    eng.allpairs(use_tensor_cores=tc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = eng.allpairs(use_tensor_cores=tc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s: %d x %d x %d  %.3f ms  %.1f TFLOP/s (2*U*I*dim), %.2e pairs/s  best_item[:4]=%s" % (
        "tcgen05 tf32" if tc else "cuda-core fp32", U, I, d, ms, 2.0 * U * I * d / ms / 1e9, U * I / ms * 1e3,
        out["best_item"][:4].tolist()))
