"""A few host-fed steps (the driver's loop: three batches handed over ahead) for an ncu launch list of the feed graph's
kernels.  Usage (GPU box): python tools/prof_feed.py [workload] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import tf_recomm_b200  # noqa: E402,F401
from tf_recomm_b200.engine import SvdEngine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ml25m_d128_b65536"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    w = bench.WORKLOADS[name]
    cols = bench.make_columns(w)
    B = w["B"]
    eng = SvdEngine(w["U"], w["I"], w["d"], bench.LR, bench.REG, device_init_seed=1)
    rng = np.random.default_rng(3)
    batches = []
    for _ in range(steps):
        rows = rng.integers(0, len(cols[0]), B)
        batches.append(tuple(c[rows].astype(np.float64) for c in cols))
    ahead = 3
    for j in range(ahead):
        eng.prefetch_host(*batches[j])
    for j in range(steps):
        if j + ahead < steps:
            eng.prefetch_host(*batches[j + ahead])
        for e in eng._host_state[B]["pending"]:   # (under a profiler the worker may lag: keep every step on the graph path)
            if e["fut"] is not None:
                e["fut"].result()
        eng.train_step_host(*batches[j])
    torch.cuda.synchronize()
    print("ok: %d host-fed steps, global_step %d" % (steps, eng.global_step))
    eng.close()


if __name__ == "__main__":
    main()
