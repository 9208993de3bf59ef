"""Host-side breakdown of SvdEngine.train_step_host (the Session.run feed_dict path): where the end-to-end
microseconds go.  Usage (GPU box): python tools/e2e_breakdown.py [workload]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import tf_recomm_b200  # noqa: E402,F401
from tf_recomm_b200._lib import check  # noqa: E402
from tf_recomm_b200.engine import SvdEngine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ml25m_d128_b65536"
    w = bench.WORKLOADS[name]
    cols = bench.make_columns(w)
    B = w["B"]
    eng = SvdEngine(w["U"], w["I"], w["d"], bench.LR, bench.REG, device_init_seed=1)
    rng = np.random.default_rng(3)
    batches = []
    for _ in range(40):
        rows = rng.integers(0, len(cols[0]), B)
        batches.append((cols[0][rows].astype(np.float64), cols[1][rows].astype(np.float64), cols[2][rows].astype(np.float64)))
    for b in batches[:5]:
        eng.train_step_host(*b)
    torch.cuda.synchronize()
    t = dict(total=0.0)
    t0 = time.perf_counter()
    for users, items, rates in batches[5:]:
        eng.train_step_host(users, items, rates)
    t["total"] = time.perf_counter() - t0
    n = len(batches) - 5
    print(name, {k: round(v / n * 1e6, 1) for k, v in t.items()}, "us per train_step_host")


if __name__ == "__main__":
    main()
