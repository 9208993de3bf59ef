"""The Adam table pass alone: back-to-back launches of tfr_adam_stream_multi at one workload's table sizes, CUDA
events around the batch and around single launches, for a list of tuning-knob settings (tfr_tune_set), plus a
bit-identity check of every setting against the first one.
Usage (GPU box): python tools/pass_bench.py [workload | U,I,d] [launches] [cfg ...]
  cfg = comma-separated KNOB=VALUE list, e.g.  PASS_RING=0  PASS_RING=1,RING_STAGES=6,RING_STAGE_KB=24"""
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

DEFAULT_CFGS = ["PASS_RING=0", "PASS_RING=1"]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ml25m_d128_b65536"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    cfgs = sys.argv[3:] or DEFAULT_CFGS
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
        eng.run_stream_steps(4, use_graph=False)   # leaves a live slice (slot map + gsum) behind
    torch.cuda.synchronize()
    # steady-state optimizer slots: every row has been touched (m, v in the normal range).  With the all-zero
    # slots of a fresh model most rows take the slow paths of the correctly rounded sqrt / divide (0 operands).
    g = torch.Generator(device=eng.device); g.manual_seed(3)
    for n_, t_ in eng.slots.items():
        if n_.startswith("m_"):
            t_.normal_(0.0, 1e-2, generator=g)
        else:
            t_.uniform_(1e-6, 1e-2, generator=g)
    # make the last step's slice live again for the pass alone: stamps are global_step's; step back by one
    off = _lib.OptScalars.global_step.offset
    gs = eng.global_step
    ws = eng.step_ws(w["B"])
    tabs = bench._adam_tables(eng, ws, _lib)
    st = torch.cuda.current_stream().cuda_stream
    opt = eng.opt.data_ptr()
    snap = {k: v.clone() for k, v in list(eng.t.items()) + list(eng.slots.items())}
    bytes_ = 24.0 * (w["U"] + w["I"]) * (w["d"] + 1)
    ref = None
    for cfg in cfgs:
        for kv in cfg.split(","):
            k, v = kv.split("=")
            _lib.tune_set(k, int(v))
        # bit-identity: one launch from the snapshot, with the previous step's slice live
        for k, v in snap.items():
            (eng.t[k] if k in eng.t else eng.slots[k]).copy_(v)
        if gs > 0:
            eng.opt[off:off + 8].copy_(torch.tensor([gs - 1], dtype=torch.int64).view(torch.uint8))
        check(eng.L.tfr_adam_stream_multi(tabs, 4, opt, 15, st))
        torch.cuda.synchronize()
        got = {k: (eng.t[k] if k in eng.t else eng.slots[k]).clone() for k in snap}
        same = "reference"
        if ref is None:
            ref = got
        else:
            same = "bit-identical" if all(torch.equal(ref[k].view(torch.int32), got[k].view(torch.int32)) for k in ref) \
                else "DIFFERENT"
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
        b2b = e0.elapsed_time(e1) * 1e3 / n
        print("%-60s single median %6.1f us (min %6.1f)  back-to-back %6.1f us = %5.0f GB/s  [%s]" % (
            cfg, float(np.median(singles)), min(singles), b2b, bytes_ / b2b / 1e3, same), flush=True)


if __name__ == "__main__":
    main()
