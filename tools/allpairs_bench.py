"""All-pairs scoring throughput (BASELINE configs[3]: 162541 x 62423, dim 128): the tcgen05 sweep with each fused
consumer (top-1; top-50 ranking with float64 rescore + certificate; squared error on observed pairs), CUDA events.
Usage (GPU box): python tools/allpairs_bench.py [users items dim] [--simt]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tf_recomm_b200  # noqa: E402,F401
from tf_recomm_b200.engine import SvdEngine  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def run(U, I, d, simt=False, reps=5, n_obs=None):
    eng = SvdEngine(U, I, d, 1e-3, 0.05, device_init_seed=3)
    flop = 2.0 * U * I * d
    res = {}
    ms, out = timed(lambda: eng.allpairs(use_tensor_cores=True), reps)
    res["top1"] = dict(ms=ms, tflops=flop / ms / 1e9, consumer="bias + running best item per user (top-1) in the epilogue")
    ms, out = timed(lambda: eng.rank_all_users(k=50, n_cand=64), max(1, reps // 2))
    res["top50"] = dict(ms=ms, tflops=flop / ms / 1e9, n_uncertified=out[2],
                        consumer="top-64 candidates per user in the epilogue + float64 rescore, certificate, exact fallback "
                                 "(whole call: GEMM sweep + rescore + fallback rows)")
    rng = np.random.default_rng(0)
    n_obs = n_obs or min(25_000_000, 200 * U)
    users = torch.from_numpy(rng.integers(0, U, n_obs).astype(np.int32)).cuda()
    items = torch.from_numpy(rng.integers(0, I, n_obs).astype(np.int32)).cuda()
    rates = torch.from_numpy(rng.integers(1, 6, n_obs).astype(np.float32)).cuda()
    ms, out = timed(lambda: eng.observed_rmse(users, items, rates), max(1, reps // 2))
    res["observed_se"] = dict(ms=ms, tflops=flop / ms / 1e9, pairs=int(n_obs),
                              consumer="squared error on the observed pairs (CSR walk in the epilogue); whole call incl. "
                                       "the CSR build (torch sort)")
    if simt:
        ms, out = timed(lambda: eng.allpairs(use_tensor_cores=False), 1)
        res["cuda_core_fp32"] = dict(ms=ms, tflops=flop / ms / 1e9)
    return res


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    U, I, d = (int(x) for x in args[:3]) if len(args) >= 3 else (162541, 62423, 128)
    for k, v in run(U, I, d, simt="--simt" in sys.argv).items():
        print("%-16s %9.3f ms  %7.1f TFLOP/s (2*U*I*dim)  %s" % (k, v["ms"], v["tflops"], {a: b for a, b in v.items() if a not in ("ms", "tflops")}))
