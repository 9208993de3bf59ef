"""Which stage of the fused ranking (tfr_allpairs_consume, k = 50) costs what: the tensor-core sweep with the candidate
buffers, the float64 rescore + certificate, the exact fallback rows.  Usage (GPU box): python tools/ap_stage_timing.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tf_recomm_b200  # noqa: E402,F401
from tf_recomm_b200 import _lib  # noqa: E402
from tf_recomm_b200.engine import SvdEngine  # noqa: E402

U, I, d = 162541, 62423, 128
eng = SvdEngine(U, I, d, 1e-3, 0.05, device_init_seed=3)
eng.rank_all_users(k=50, n_cand=64)
for stages, what in ((1, "sweep + candidate buffers"), (3, "+ rescore"), (7, "+ exact rows (all)")):
    _lib.tune_set("AP_STAGES", stages)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = eng.rank_all_users(k=50, n_cand=64)
    e1.record(); torch.cuda.synchronize()
    print("%-28s %8.2f ms" % (what, e0.elapsed_time(e1) / 3))
_lib.tune_set("AP_STAGES", 7)
