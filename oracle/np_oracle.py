"""Independent numpy restatement of the same path -- TEST INFRASTRUCTURE, used to pin oracle/tfr_oracle.c.

Written separately from the C file on purpose (vectorised numpy instead of loops) so that a slip in
one is caught by the other.  Cites the same reference lines.  fp32 throughout; reduction ORDER of
np.sum differs from the C file's sequential order, so float comparisons between the two use a
tolerance while all integer outputs must match exactly.
"""
import numpy as np

f32 = np.float32


def forward(p, users, items, abs_item=False):
    """ops.py:13-14,37-38,44-47."""
    u = p["user_feat"][users]
    v = p["item_feat"][items]
    if abs_item:
        v = np.abs(v)  # ops.py:44
    x = np.sum(u * v, axis=1, dtype=f32)
    x = x + p["mu"][0]
    x = x + p["user_bias"][users]
    x = x + p["item_bias"][items]
    return x.astype(f32)


def sigmoid(x):
    """ops.py:94-95."""
    return 1 / (1 + np.exp(-x))


def dloss(x, z, sigmoid_ce=False):
    """README: ops.py:124 (l2_loss(infer-rate)) -> x-z; fork: ops.py:125 -> sigmoid(x)-z."""
    if not sigmoid_ce:
        return (x - z).astype(f32)
    return (sigmoid(x.astype(np.float64)) - z).astype(f32)


def grads(p, users, items, rates, reg, abs_item=False, sigmoid_ce=False, reg_bias=False):
    """SURVEY 8a row a10 (TF autodiff of ops.py:124-126,140 through ops.py:44-47,81-89)."""
    reg = f32(reg)
    x = forward(p, users, items, abs_item)
    e = dloss(x, rates.astype(f32), sigmoid_ce)
    u = p["user_feat"][users]
    v = p["item_feat"][items]
    if abs_item:
        g_u = e[:, None] * np.abs(v) + reg * u
        g_v = (e[:, None] * u) * np.sign(v) + reg * v
    else:
        g_u = e[:, None] * v + reg * u
        g_v = e[:, None] * u + reg * v
    g_ub = e.copy()
    g_ib = e.copy()
    if reg_bias:
        g_ub = g_ub + reg * p["user_bias"][users]
        g_ib = g_ib + reg * p["item_bias"][items]
    return dict(logits=x, err=e, g_uf=g_u.astype(f32), g_if=g_v.astype(f32), g_ub=g_ub.astype(f32),
                g_ib=g_ib.astype(f32), g_mu=np.sum(e, dtype=f32))


def unique_first_occurrence(ids):
    """tf.unique: first-occurrence order."""
    ids = np.asarray(ids)
    srt, first, inv = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # sorted-unique slots ranked by first position
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return srt[order].astype(np.int32), rank[inv].astype(np.int32)


def segment_sum(values, idx, n):
    """tf.unsorted_segment_sum CPU (A.3). np.add.at is unbuffered and applies in index order."""
    out = np.zeros((n,) + values.shape[1:], f32)
    np.add.at(out, idx, values)
    return out


def adam_lr_t(lr, b1p, b2p):
    return f32(f32(lr) * np.sqrt(f32(1) - f32(b2p))) / (f32(1) - f32(b1p))


def adam_sparse(var, m, v, uniq, gsum, lr_t, beta1=0.9, beta2=0.999, eps=1e-8):
    """TF adam.py::_apply_sparse_shared (A.4), returns new (var, m, v)."""
    b1, b2, eps, lr_t = f32(beta1), f32(beta2), f32(eps), f32(lr_t)
    m = m * b1
    m[uniq] += gsum * (f32(1) - b1)
    v = v * b2
    v[uniq] += (gsum * gsum) * (f32(1) - b2)
    var = var - (lr_t * m) / (np.sqrt(v) + eps)
    return var.astype(f32), m.astype(f32), v.astype(f32)


def adam_dense_zero_filled(var, m, v, uniq, gsum, lr, b1p, b2p, beta1=0.9, beta2=0.999, eps=1e-8):
    """Cross-check formulation: dense Adam (TF epsilon-hat form) fed a zero-filled dense gradient."""
    g = np.zeros_like(var)
    g[uniq] = gsum
    b1, b2 = f32(beta1), f32(beta2)
    m = b1 * m + (f32(1) - b1) * g
    v = b2 * v + (f32(1) - b2) * g * g
    lr_t = adam_lr_t(lr, b1p, b2p)
    var = var - lr_t * m / (np.sqrt(v) + f32(eps))
    return var.astype(f32), m.astype(f32), v.astype(f32)


def fm_forward_dense(X, w0, W, V):
    """forward.py:21-22 verbatim semantics for dense/scipy X with 0/1 or real entries (x^2 form)."""
    X = np.asarray(X, np.float64)
    V = np.asarray(V, np.float64)
    return w0 + X @ np.asarray(W, np.float64) + 0.5 * (np.linalg.norm(X @ V, axis=1) ** 2 - ((X ** 2) @ (V ** 2)).sum(axis=1))
