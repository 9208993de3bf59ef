"""Initial values of the five variables (ops.py:8-12,29-32), generated with numpy so that the SAME arrays can
be injected into the CUDA engine and into the CPU oracle (TF's Philox stream is not reproducible outside TF,
SURVEY A.1 -- "identical seeds" means identical injected tables).

  bias_global   []      no initializer -> TF default glorot_uniform, limit sqrt(6/(1+1)) = sqrt(3)   (ops.py:8)
  user/item_bias [n]    fork: truncated_normal(stddev=1) (ops.py:9-12); README era: TF default
                        (glorot_uniform with fans n,n -> limit sqrt(3/n)); selectable with bias_init
  *_features    [n,dim] truncated_normal(stddev=0.02), resampled outside +-2 sigma            (ops.py:29-32)
"""
import numpy as np


def truncated_normal(rng, shape, stddev):
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * stddev).astype(np.float32)


def init_tables(user_num, item_num, dim, seed=13575, bias_init="glorot"):
    """bias_init: 'glorot' (README-era default initializer), 'truncated_normal' (fork, ops.py:9-12) or 'zeros'."""
    rng = np.random.default_rng(seed)
    lim = np.sqrt(3.0)
    mu = rng.uniform(-lim, lim, size=1).astype(np.float32)
    if bias_init == "truncated_normal":
        ub = truncated_normal(rng, (user_num,), 1.0)
        ib = truncated_normal(rng, (item_num,), 1.0)
    elif bias_init == "glorot":
        ub = rng.uniform(-np.sqrt(3.0 / user_num), np.sqrt(3.0 / user_num), size=user_num).astype(np.float32)
        ib = rng.uniform(-np.sqrt(3.0 / item_num), np.sqrt(3.0 / item_num), size=item_num).astype(np.float32)
    elif bias_init == "zeros":
        ub = np.zeros(user_num, np.float32)
        ib = np.zeros(item_num, np.float32)
    else:
        raise ValueError("bias_init must be glorot | truncated_normal | zeros, got %r" % (bias_init,))
    uf = truncated_normal(rng, (user_num, dim), 0.02)
    itf = truncated_normal(rng, (item_num, dim), 0.02)
    return dict(mu=mu, user_bias=ub, item_bias=ib, user_feat=uf, item_feat=itf)
