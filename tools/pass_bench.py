"""The Adam table pass alone: N back-to-back launches of tfr_adam_stream_multi at one workload's table sizes, CUDA
events around the batch and around single launches.  Env knobs: TFR_STREAM_THREADS / _CTAS_PER_SM / _UNROLL.
Usage (GPU box): python tools/pass_bench.py [workload] [launches]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import tf_recomm_b200  # noqa: E402,F401
from tf_recomm_b200 import _lib  # noqa: E402
from tf_recomm_b200._lib import check  # noqa: E402
from tf_recomm_b200.engine import SvdEngine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ml25m_d128_b65536"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    if "," in name:  # "U,I,d": tables of that size, no training data (the pass alone at scale)
        U, I, d = (int(x) for x in name.split(","))
        w = dict(U=U, I=I, d=d, B=65536)
        eng = SvdEngine(U, I, d, bench.LR, bench.REG, device_init_seed=1)
    else:
        w = bench.WORKLOADS[name]
        cols = bench.make_columns(w)
        eng = SvdEngine(w["U"], w["I"], w["d"], bench.LR, bench.REG, device_init_seed=1)
        eng.set_train_data(*cols)
        np.random.seed(1)
        eng.set_index_stream(np.random.randint(0, len(cols[0]), 6 * w["B"]), w["B"])
        eng.run_stream_steps(4, use_graph=False)
    torch.cuda.synchronize()
    if os.environ.get("TFR_PASS_WARM_STATE", "1") == "1":
        # steady-state optimizer slots: every row has been touched (m, v in the normal range).  With the all-zero
        # slots of a fresh model most rows take the slow paths of the correctly rounded sqrt / divide (0 operands).
        g = torch.Generator(device=eng.device); g.manual_seed(3)
        for n_, t_ in eng.slots.items():
            if n_.startswith("m_"):
                t_.normal_(0.0, 1e-2, generator=g)
            else:
                t_.uniform_(1e-6, 1e-2, generator=g)
    ws = eng.step_ws(w["B"])
    tabs = bench._adam_tables(eng, ws, _lib)
    if os.environ.get("TFR_PASS_NOSLOT"):
        for k_ in range(4):
            tabs[k_].slot = None
    st = torch.cuda.current_stream().cuda_stream
    opt = eng.opt.data_ptr()
    for _ in range(3):
        check(eng.L.tfr_adam_stream_multi(tabs, 4, opt, 15, st))
    singles = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(eng.L.tfr_adam_stream_multi(tabs, 4, opt, 15, st))
        e1.record()
        torch.cuda.synchronize()
        singles.append(e0.elapsed_time(e1) * 1e3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        check(eng.L.tfr_adam_stream_multi(tabs, 4, opt, 15, st))
    e1.record()
    torch.cuda.synchronize()
    if os.environ.get("TFR_PASS_INTERLEAVED"):
        rows = w["U"] + w["I"]
        T = torch.empty(rows, 3, w["d"], device=eng.device)
        T[:, 0].normal_(0, 0.02); T[:, 1].normal_(0, 1e-2); T[:, 2].uniform_(1e-6, 1e-2)
        for co in (1, 0):
            for _ in range(3):
                check(eng.L.tfr_experiment_interleaved_pass(T.data_ptr(), rows, w["d"], opt, co, st))
            e0.record()
            for _ in range(n):
                check(eng.L.tfr_experiment_interleaved_pass(T.data_ptr(), rows, w["d"], opt, co, st))
            e1.record()
            torch.cuda.synchronize()
            t_ = e0.elapsed_time(e1) * 1e3 / n
            print("interleaved [rows][3][%d] %s: %.1f us = %.0f GB/s" % (w["d"], "copy-only" if co else "decay math", t_,
                                                                         24.0 * rows * w["d"] / t_ / 1e3))
    bytes_ = 24.0 * (w["U"] + w["I"]) * (w["d"] + 1)
    b2b = e0.elapsed_time(e1) * 1e3 / n
    print("%s threads=%s ctas=%s unroll=%s: single median %.1f us (min %.1f), back-to-back %.1f us = %.0f GB/s" % (
        name, os.environ.get("TFR_STREAM_THREADS", "default"), os.environ.get("TFR_STREAM_CTAS_PER_SM", "default"),
        os.environ.get("TFR_STREAM_UNROLL", "default"), float(np.median(singles)), min(singles), b2b,
        bytes_ / b2b / 1e3))


if __name__ == "__main__":
    main()
