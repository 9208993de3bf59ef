"""Times every kernel of one train step in situ (CUDA events around each C-ABI call, one stream, warm caches).
Usage (GPU box): python tools/step_breakdown.py [workload] [steps]   -> table on stdout."""
import ctypes as C
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


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ml25m_d128_b65536"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    w = bench.WORKLOADS[name]
    cols = bench.make_columns(w)
    eng = SvdEngine(w["U"], w["I"], w["d"], bench.LR, bench.REG, device_init_seed=1)
    L = eng.L
    B, d, U, I = w["B"], w["d"], w["U"], w["I"]
    rng = np.random.default_rng(5)
    tp = C.byref(eng.tables_struct)
    ws = eng.step_ws(B)
    st = torch.cuda.current_stream().cuda_stream
    opt = eng.opt.data_ptr()
    logits = torch.empty(B, device=eng.device); infer = torch.empty(B, device=eng.device)
    T, S = eng.t, eng.slots
    from tf_recomm_b200 import _lib
    tabs = bench._adam_tables(eng, ws, _lib)
    acc = {}

    def timed(label, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        acc.setdefault(label, []).append((e0, e1))

    for s in range(steps + 3):
        rows = rng.integers(0, len(cols[0]), B)
        du = eng._dev_i32(cols[0][rows]); di = eng._dev_i32(cols[1][rows]); dr = eng._dev_f32(cols[2][rows])
        if s == 3:
            acc.clear()
        timed("begin_step", lambda: check(L.tfr_svd_begin_step(opt, st)))
        timed("fwd_err", lambda: check(L.tfr_svd_fwd_err(tp, opt, du.data_ptr(), di.data_ptr(), dr.data_ptr(), B,
                                                         logits.data_ptr(), infer.data_ptr(), C.byref(ws), st)))
        timed("dedup_sort", lambda: check(L.tfr_dedup_sort_pairs(du.data_ptr(), U, ws.su_ids, ws.su_pos, di.data_ptr(), I,
                                                                 ws.si_ids, ws.si_pos, B, ws.sort_ws, ws.sort_ws_bytes, st)))
        timed("segment_grads(tiles+fixup)", lambda: check(L.tfr_svd_segment_grads(tp, opt, du.data_ptr(), di.data_ptr(), B,
                                                                                  C.byref(ws), st)))
        timed("adam_pass (4 tables, 1 launch)", lambda: check(L.tfr_adam_stream_multi(tabs, 4, opt, 15, st)))
        nu = int((eng.user_slot >= 0).sum()); ni = int((eng.item_slot >= 0).sum())
        timed("finish", lambda: check(L.tfr_svd_finish_step(tp, opt, du.data_ptr(), di.data_ptr(), B, C.byref(ws),
                                                            bench._n_partials(d, B), st)))
    torch.cuda.synchronize()
    tot = 0.0
    print("workload %s: unique users %d / %d, unique items %d / %d" % (name, nu, U, ni, I))
    for k, evs in acc.items():
        ms = np.mean([a.elapsed_time(b) for a, b in evs])
        tot += ms
        print("%-30s %9.2f us" % (k, ms * 1e3))
    print("%-30s %9.2f us (serial sum)" % ("total", tot * 1e3))


if __name__ == "__main__":
    main()
