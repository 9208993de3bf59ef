"""Full RMSE-curve parity run (BASELINE configs[0]/[1]): README model on the synthetic ML-1M shape for N epochs on the
CUDA path and on the CPU oracle with identical injected tables and the reference's index stream; prints both curves
and the largest gap.  Usage (GPU box): python tools/rmse_curve.py [batch] [epochs] > gpurun_out/rmse_curve.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import tf_recomm_b200  # noqa: E402,F401
from tf_recomm_b200 import dataio, init, synthetic  # noqa: E402
from tf_recomm_b200.engine import SvdEngine  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    U, I, N = synthetic.SHAPES["ml1m"]
    d, lr, reg = 15, 1e-3, 0.05
    users, items, rates = synthetic.make_ratings(U, I, N, seed=13575)
    (tu, ti, tr), (vu, vi, vr) = synthetic.split(users, items, rates)
    nb = len(tu) // B
    tabs = init.init_tables(U, I, d, seed=13575)
    eng = SvdEngine(U, I, d, lr, reg, tables=tabs)
    orc = oracle.SvdOracle(tabs["mu"], tabs["user_bias"], tabs["item_bias"], tabs["user_feat"], tabs["item_feat"], lr, reg)
    np.random.seed(13575)
    it = dataio.ShuffleIterator([tu, ti, tr], batch_size=B)
    eng.set_train_data(tu, ti, tr)
    eng.set_se_ring(nb)
    from collections import deque
    win = deque(maxlen=nb)
    done, gap = 0, 0.0
    t_gpu = t_cpu = 0.0
    print("epoch  gpu_train  gpu_val  cpu_train  cpu_val  |gap|")
    for ep in range(epochs + 1):
        n = 1 if ep == 0 else nb
        stream = it.draw_index_stream(n)
        t0 = time.perf_counter()
        eng.set_index_stream(stream, B)
        eng.run_stream_steps(n, use_graph=True)
        ring = eng.se_ring.cpu().numpy()
        done += n
        filled = min(done, nb)
        g_train = float(np.sqrt(ring[:filled].sum() / (filled * B)))
        g_val = float(np.sqrt(np.mean((vr.astype(np.float64) - eng.forward(vu, vi)[1].cpu().numpy()) ** 2)))
        t_gpu += time.perf_counter() - t0
        t0 = time.perf_counter()
        for k in range(n):
            rows = stream[k * B:(k + 1) * B]
            _, infer = orc.train_step(tu[rows], ti[rows], tr[rows])
            win.append(np.sum((tr[rows].astype(np.float64) - infer.astype(np.float64)) ** 2))
        c_train = float(np.sqrt(np.sum(win) / (len(win) * B)))
        c_val = float(np.sqrt(np.mean((vr.astype(np.float64) - orc.forward(vu, vi)[1]) ** 2)))
        t_cpu += time.perf_counter() - t0
        g = max(abs(g_train - c_train), abs(g_val - c_val))
        gap = max(gap, g)
        if ep < 3 or ep % 10 == 0 or ep == epochs:
            print("%3d  %.6f  %.6f  %.6f  %.6f  %.2e" % (ep, g_train, g_val, c_train, c_val, g))
    print("max |gap| over %d reports: %.3e (bar: 1e-3)   wall: cuda path %.1f s, cpu oracle %.1f s (%d threads)" % (
        epochs + 1, gap, t_gpu, t_cpu, os.cpu_count()))


if __name__ == "__main__":
    main()
