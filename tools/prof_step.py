"""A few eager (no graph) train steps of one workload, for `ncu -k regex:<kernel> --launch-skip N -c 1` captures.
Usage (GPU box): python tools/prof_step.py [workload] [steps]"""
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
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    w = bench.WORKLOADS[name]
    cols = bench.make_columns(w)
    eng = SvdEngine(w["U"], w["I"], w["d"], bench.LR, bench.REG, device_init_seed=1)
    eng.set_train_data(*cols)
    np.random.seed(1)
    eng.set_index_stream(np.random.randint(0, len(cols[0]), (steps + 2) * w["B"]), w["B"])
    eng.run_stream_steps(steps, use_graph=False, pipeline=True)
    torch.cuda.synchronize()
    print("ran %d eager steps of %s" % (steps, name))


if __name__ == "__main__":
    main()
