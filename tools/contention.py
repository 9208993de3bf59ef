"""How much do the chain kernels (forward + segment sums, fix-up) slow down when a table pass streams beside them,
and how much does the pass slow down?  Decides whether the step can be re-ordered so that the chain of a step runs
under the part of the table pass that does not depend on it.
Usage (GPU box): python tools/contention.py [cfg ...]   cfg = KNOB=VALUE[,KNOB=VALUE...] applied to the PASS"""
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
    cfgs = sys.argv[1:] or ["PASS_RING=0", "PASS_RING=1"]
    w = bench.WORKLOADS["ml25m_d128_b65536"]
    cols = bench.make_columns(w)
    eng = SvdEngine(w["U"], w["I"], w["d"], bench.LR, bench.REG, device_init_seed=1)
    eng.set_train_data(*cols)
    B = w["B"]
    np.random.seed(1)
    eng.set_index_stream(np.random.randint(0, len(cols[0]), 8 * B), B)
    eng.run_stream_steps(4, use_graph=False)
    torch.cuda.synchronize()
    dev = eng.device
    # a second table of the same size for the pass to stream (so that the chain's inputs stay put)
    rows = w["U"] + w["I"]
    T2 = torch.empty(rows, 3, w["d"], device=dev)
    T2[:, 0].normal_(0, 0.02); T2[:, 1].normal_(0, 1e-2); T2[:, 2].uniform_(1e-6, 1e-2)
    tabs = (_lib.AdamTable * 1)()
    tabs[0].var, tabs[0].m, tabs[0].v = T2[:, 0].data_ptr(), T2[:, 1].data_ptr(), T2[:, 2].data_ptr()
    tabs[0].rows, tabs[0].width, tabs[0].slot, tabs[0].gsum, tabs[0].stride = rows, w["d"], None, None, 3 * w["d"]
    main_s = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)
    opt = eng.opt.data_ptr()
    # the batch at the cursor, assembled + sorted into set 0
    eng._prefetch(B, 0, True, main_s.cuda_stream)
    bufs, ws = eng.stream_buffers(B, 0), eng.workspace((B, 0))

    def chain():
        check(eng.L.tfr_svd_train_step_presorted(C.byref(eng.tables_struct), opt, bufs["users"].data_ptr(),
                                                 bufs["items"].data_ptr(), bufs["rates"].data_ptr(), B,
                                                 bufs["logits"].data_ptr(), bufs["infer"].data_ptr(), eng.flags,
                                                 eng.var_mask, 1, ws.data_ptr(), ws.numel(), main_s.cuda_stream))

    def the_pass():
        check(eng.L.tfr_adam_stream_multi(tabs, 1, opt, 15, side.cuda_stream))

    def timed(fn, stream, n=20):
        ts = []
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return float(np.median(ts))

    for _ in range(3):
        chain()
    torch.cuda.synchronize()
    print("chain alone (tiles + fix-up): %.1f us" % timed(chain, main_s), flush=True)
    for cfg in cfgs:
        for kv in cfg.split(","):
            k, v = kv.split("=")
            _lib.tune_set(k, int(v))
        for _ in range(3):
            the_pass()
        torch.cuda.synchronize()
        alone = timed(the_pass, side)
        tc, tp = [], []
        for _ in range(20):
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(side); the_pass(); the_pass(); p1.record(side)
            c0.record(main_s); chain(); c1.record(main_s)
            torch.cuda.synchronize()
            tc.append(c0.elapsed_time(c1) * 1e3); tp.append(p0.elapsed_time(p1) * 1e3 / 2)
        print("%-50s pass alone %6.1f us | beside each other: chain %6.1f us, pass %6.1f us (per pass, 2 back to back)"
              % (cfg, alone, float(np.median(tc)), float(np.median(tp))), flush=True)


if __name__ == "__main__":
    main()
