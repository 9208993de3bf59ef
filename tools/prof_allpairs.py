"""One all-pairs sweep per consumer at a reduced user count (ncu replays every launch ~40 times): for
`ncu -k regex:allpairs_tc_kernel`.  Usage (GPU box): python tools/prof_allpairs.py [users]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tf_recomm_b200  # noqa: E402,F401
from tf_recomm_b200.engine import SvdEngine  # noqa: E402

U = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 2
I, d = 62423, 128
eng = SvdEngine(U, I, d, 1e-3, 0.05, device_init_seed=3)
eng.allpairs(use_tensor_cores=True)
eng.rank_all_users(k=50, n_cand=64)
rng = np.random.default_rng(0)
n = 150 * U
eng.observed_rmse(rng.integers(0, U, n).astype(np.int32), rng.integers(0, I, n).astype(np.int32),
                  rng.integers(1, 6, n).astype(np.float32))
torch.cuda.synchronize()
print("ran top-1, top-50 and observed-pairs sweeps at %d x %d x %d" % (U, I, d))
