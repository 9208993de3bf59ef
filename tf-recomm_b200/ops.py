"""Model-graph API under the reference's names: inference_svd, optimization, sigmoid (ops.py in
jilljenn/TF-recomm).  The functions keep the reference's call shapes -- both the current one
(ops.py:6,91,118,153) and the README-era one visible at svd_train_val.py:49 -- and return handles that a
session.Session runs through libtfrecomm.so.  Nothing is computed here.

Which model is built is NOT decided by the call shape but by `variant` (default config.MODEL_VARIANT):
"readme" = dot + biases, squared error, L2 on gathered embeddings, Adam (README.md:31-39);
"fork"   = abs item factors, sigmoid cross-entropy, bias L2, SGD (ops.py:44,85-89,125,145).
"""
import numpy as np

from . import _lib, config, session
from .session import (Placeholder, Session, float32, global_variables_initializer, group, int32,  # noqa: F401
                      local_variables_initializer, placeholder, reset_default_graph, train)


def inference_svd(user_batch, item_batch, *args, **kwargs):
    """Current shape: inference_svd(user_batch, item_batch, wins_batch, fails_batch, user_num, item_num, dim=5,
    device="/cpu:0") -> (infer, logits, regularizer, user_bias, user_features, item_bias, item_features).
    README-era shape: inference_svd(user_batch, item_batch, user_num, item_num, dim=5, device="/cpu:0")
    -> (infer, regularizer).
    Declares the five variables of ops.py:8-12,29-32, the gathers (:13-14,37-38), the score (:44-47), the head
    (:76-78) and the regulariser (:81-89).  Extra keywords: variant="readme"|"fork", init_tables=dict, seed=int."""
    variant = kwargs.pop("variant", None) or config.MODEL_VARIANT
    init_tables = kwargs.pop("init_tables", None)
    seed = kwargs.pop("seed", config.SEED)
    if variant not in ("readme", "fork"):
        raise ValueError("variant must be 'readme' or 'fork'")
    new_shape = "wins_batch" in kwargs or (len(args) >= 1 and (args[0] is None or isinstance(args[0], Placeholder)))
    if new_shape:
        names = ("wins_batch", "fails_batch", "user_num", "item_num", "dim", "device")
    else:
        names = ("user_num", "item_num", "dim", "device")
    vals = dict(zip(names, args))
    for k in list(kwargs):
        if k in names:
            if k in vals:
                raise TypeError("inference_svd() got multiple values for %r" % k)
            vals[k] = kwargs.pop(k)
    if kwargs:
        raise TypeError("inference_svd() got unexpected arguments %s" % sorted(kwargs))
    for k in ("user_num", "item_num"):
        if k not in vals:
            raise TypeError("inference_svd() missing %r" % k)
    m = session.Model(user_batch, item_batch, vals.get("wins_batch"), vals.get("fails_batch"), vals["user_num"],
                      vals["item_num"], vals.get("dim", 5), variant)
    m.init_tables, m.seed = init_tables, seed
    session._set_model(m)
    h = m.h
    if new_shape:
        return (h["infer"], h["logits"], h["regularizer"], h["user_bias"], h["user_feat"], h["item_bias"],
                h["item_feat"])
    return h["infer"], h["regularizer"]


def sigmoid(x):
    """numpy helper the driver applies to fetched logits (ops.py:94-95)."""
    return 1 / (1 + np.exp(-x))


def optimization(infer, *args, **kwargs):
    """Current shape: optimization(infer, logits, regularizer, rate_batch, learning_rate, reg, device="/cpu:0",
    var_list=None) -> (cost, train_op).  README-era: optimization(infer, regularizer, rate_batch,
    learning_rate=..., reg=..., device=...) -> (cost_l2, train_op).
    cost = data_loss + reg * regularizer (ops.py:124-126,137-140); train_op = Optimizer(lr).minimize(cost,
    global_step[, var_list]) (ops.py:143-149).  optimizer="adam"|"sgd" overrides the variant's default."""
    if session.train.get_global_step() is None:  # ops.py:119-120
        raise AssertionError("global_step is None")
    m = infer.model
    optimizer = kwargs.pop("optimizer", None) or ("adam" if m.variant == "readme" else "sgd")
    new_shape = len(args) >= 3 and isinstance(args[2], Placeholder) or "logits" in kwargs
    names = (("logits", "regularizer", "rate_batch", "learning_rate", "reg", "device", "var_list") if new_shape
             else ("regularizer", "rate_batch", "learning_rate", "reg", "device", "var_list"))
    vals = dict(zip(names, args))
    for k in list(kwargs):
        if k in names:
            vals[k] = kwargs.pop(k)
    if kwargs:
        raise TypeError("optimization() got unexpected arguments %s" % sorted(kwargs))
    for k in ("rate_batch", "learning_rate", "reg"):
        if k not in vals:
            raise TypeError("optimization() missing %r" % k)
    m.rate_batch = vals["rate_batch"]
    m.lr, m.reg = float(vals["learning_rate"]), float(vals["reg"])
    if optimizer == "sgd":
        m.flags |= _lib.OPT_SGD
    elif optimizer == "adam":
        m.flags &= ~_lib.OPT_SGD
    else:
        raise ValueError("optimizer must be 'adam' or 'sgd'")
    var_list = vals.get("var_list")
    if var_list is not None:
        mask = 0
        for v in var_list:
            if not isinstance(v, session.Handle) or v.kind not in session.VAR_BITS:
                raise TypeError("var_list entries must be variables returned by inference_svd, got %r" % (v,))
            mask |= session.VAR_BITS[v.kind]
        m.var_mask = mask
    return m.h["cost"], m.h["train_op"]
