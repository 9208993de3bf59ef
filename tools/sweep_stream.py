"""One configuration of the streaming Adam pass (env TFR_STREAM_CTAS_PER_SM / TFR_STREAM_UNROLL): the kernel alone
on an all-untouched table, and the full graph-replayed step.  Prints one line."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import tf_recomm_b200  # noqa: E402,F401
from tf_recomm_b200._lib import check  # noqa: E402
from tf_recomm_b200.engine import SvdEngine  # noqa: E402

w = bench.WORKLOADS["ml25m_d128_b65536"]
cols = bench.make_columns(w)
eng = SvdEngine(w["U"], w["I"], w["d"], bench.LR, bench.REG, device_init_seed=1)
L, st = eng.L, torch.cuda.current_stream().cuda_stream
U, d, B = w["U"], w["d"], w["B"]


def alone(reps=30):
    from tf_recomm_b200 import _lib
    tabs = bench._adam_tables(eng, eng.step_ws(B), _lib)

    def go():
        check(L.tfr_adam_stream_multi(tabs, 4, eng.opt.data_ptr(), 15, st))
    for _ in range(5):
        go()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    return us, 24.0 * (U + w['I']) * (d + 1) / us / 1e3


us, gbs = alone()
eng.set_train_data(*cols)
np.random.seed(1)
steps = 100
eng.set_index_stream(np.random.randint(0, len(cols[0]), (steps + 10) * B), B)
eng.run_stream_steps(10)
secs = bench.time_stream_steps(eng, steps, torch)
bs = bench.algorithmic_bytes_step(w["U"], w["I"], d, B)
print("ctas/SM=%s unroll=%s : stream alone %.1f us = %.0f GB/s | full step %.1f us = %.0f GB/s (%.3f of 6545)" % (
    os.environ.get("TFR_STREAM_CTAS_PER_SM", "1"), os.environ.get("TFR_STREAM_UNROLL", "4"), us, gbs,
    secs / steps * 1e6, bs / (secs / steps) / 1e9, bs / (secs / steps) / 1e9 / 6545.3))
