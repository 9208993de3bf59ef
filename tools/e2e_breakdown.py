"""Breakdown of SvdEngine.train_step_host (the Session.run feed_dict path): where the end-to-end microseconds go.
Prints, for 1 .. N_FEED_SETS - 2 batches handed over ahead: the wall time per step (mean and percentiles), the host-side
phases of train_step_host, the step stream's busy / idle time by CUDA events, and the in-kernel %globaltimer timeline of a
steady-state step -- beside the device-resident step of the same box.
Usage (GPU box): [TFR_NB=215] [TFR_FEED_SETS=n] [TFR_FEED_GRAPHS=0] [TFR_FEED_WORKER=0] python tools/e2e_breakdown.py [workload]"""
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
    if os.environ.get("TFR_FEED_SETS"):
        SvdEngine.N_FEED_SETS = int(os.environ["TFR_FEED_SETS"])
    eng = SvdEngine(w["U"], w["I"], w["d"], bench.LR, bench.REG, device_init_seed=1)
    rng = np.random.default_rng(3)
    batches = []
    for _ in range(int(os.environ.get("TFR_NB", "40"))):
        rows = rng.integers(0, len(cols[0]), B)
        batches.append((cols[0][rows].astype(np.float64), cols[1][rows].astype(np.float64), cols[2][rows].astype(np.float64)))
    # the device-resident step of this box, for scale (graph replay, next batch assembled + sorted on the side stream)
    eng.set_train_data(*cols)
    eng.set_index_stream(rng.integers(0, len(cols[0]), 72 * B), B)
    eng.run_stream_steps(8, use_graph=True, pipeline=True)
    eng.set_batch_cursor(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.run_stream_steps(64, use_graph=True, pipeline=True)
    torch.cuda.synchronize()
    print("%s: device-resident step %.1f us (graph replay)" % (name, (time.perf_counter() - t0) / 64 * 1e6))
    for b in batches[:5]:
        eng.train_step_host(*b)
    torch.cuda.synchronize()
    n = len(batches) - 5
    t0 = time.perf_counter()
    for users, items, rates in batches[5:]:
        eng.train_step_host(users, items, rates)
    torch.cuda.synchronize()
    plain = (time.perf_counter() - t0) / n * 1e6
    print("%s feed_worker=%d: %.1f us/step unprefetched" % (name, eng.feed_worker, plain))
    # the driver's loop: batch t+ahead handed over before step t is asked for
    for ahead in range(1, eng.N_FEED_SETS - 1):
        t_pre = t_step = 0.0
        per = []
        nb = len(batches)
        for j in range(5, 5 + ahead):
            eng.prefetch_host(*batches[j])
        t0 = time.perf_counter()
        for j in range(5, nb):
            a = time.perf_counter()
            if j + ahead < nb:
                eng.prefetch_host(*batches[j + ahead])
            b = time.perf_counter()
            eng.train_step_host(*batches[j])
            t_pre += b - a
            t_step += time.perf_counter() - b
            per.append((time.perf_counter() - a) * 1e6)
        torch.cuda.synchronize()
        tot = (time.perf_counter() - t0) / n * 1e6
        print("  prefetched %d ahead: %.1f us/step (prefetch_host call %.1f us, train_step_host call %.1f us)"
              % (ahead, tot, t_pre / n * 1e6, t_step / n * 1e6))
        print("    per-step wall time: p10 %.0f  p50 %.0f  p90 %.0f  max %.0f us; first five: %s"
              % (np.percentile(per, 10), np.percentile(per, 50), np.percentile(per, 90), max(per), [int(x) for x in per[:5]]))
        if len(per) > 60:
            tail = np.array(per[10:])
            slow = [(i + 10, int(x)) for i, x in enumerate(tail) if x > 1.25 * np.median(tail)]
            print("    after the first ten: mean %.1f  p50 %.1f us; %d steps > 1.25 x median: %s"
                  % (tail.mean(), np.median(tail), len(slow), slow[:24]))
        eng.host_prof = {}
        for j in range(5, 5 + ahead):
            eng.prefetch_host(*batches[j])
        for j in range(5, nb):
            if j + ahead < nb:
                eng.prefetch_host(*batches[j + ahead])
            eng.train_step_host(*batches[j])
        torch.cuda.synchronize()
        print("    train_step_host:", {k: round(v / n * 1e6, 1) for k, v in eng.host_prof.items()})
        eng.host_prof = None
        # device side: events (made beforehand) on the step stream right before and after each step's launches
        A = [torch.cuda.Event(enable_timing=True) for _ in range(nb)]
        Z = [torch.cuda.Event(enable_timing=True) for _ in range(nb)]
        for j in range(5, 5 + ahead):
            eng.prefetch_host(*batches[j])
        for j in range(5, nb):
            if j + ahead < nb:
                eng.prefetch_host(*batches[j + ahead])
            A[j].record()
            eng.train_step_host(*batches[j])
            Z[j].record()
        torch.cuda.synchronize()
        js = range(10, nb - 3)
        med = lambda xs: float(np.median(list(xs))) * 1e3
        print("    step stream (events): step %.1f us, idle before the next step %.1f us"
              % (med(A[j].elapsed_time(Z[j]) for j in js), med(Z[j].elapsed_time(A[j + 1]) for j in js)))
    kernel_timeline(eng, batches, eng.N_FEED_SETS - 2)
    eng.close()


def kernel_timeline(eng, batches, ahead):
    """In-kernel %globaltimer stamps of ONE host-fed step in steady state: the stamps are reset on the step stream (no
    synchronisation) before every step, so what is left after the loop belongs to the last step and to the id sort of
    the batch handed over after it."""
    eng.enable_timeline()
    nb = len(batches)
    acc = {}
    for rep in range(7):
        seq = batches[5:] + batches[5:5 + ahead]   # the tail is handed over (its sort is what we want to see), not stepped
        for j in range(ahead):
            eng.prefetch_host(*seq[j])
        for j in range(nb - 5):
            if j + ahead < len(seq):
                eng.prefetch_host(*seq[j + ahead])
            eng.reset_timeline()
            eng.train_step_host(*seq[j])
        torch.cuda.synchronize()
        for k, (a, b) in eng.read_timeline().items():
            acc.setdefault(k, []).append((a, b))
        eng.train_step_host(*batches[0])   # drops what is pending
        torch.cuda.synchronize()
    print("  in-kernel timeline of a steady-state host-fed step (us from its first kernel's entry; median of 7):")
    for k in eng.TL_NAMES:
        if k in acc:
            a = np.median([x[0] for x in acc[k]]); b = np.median([x[1] for x in acc[k]])
            print("    %-12s start %7.1f  end %7.1f  dur %7.1f" % (k, a, b, b - a))


if __name__ == "__main__":
    main()
