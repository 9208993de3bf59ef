"""CPU oracle for the TF-recomm train step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``oracle/tfr_oracle.c`` (see that file's header: what it restates, file:line,
and why parity of the train step is UNPINNED -- TensorFlow is absent from /root/reference).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  Nothing under ``tf-recomm_b200/`` imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtfr_oracle.so")

ABS_ITEM, LOSS_SIGMOID_CE, REG_BIAS, OPT_SGD = 1, 2, 4, 8
FORK_FLAGS = ABS_ITEM | LOSS_SIGMOID_CE | REG_BIAS | OPT_SGD
VAR_MU, VAR_UB, VAR_UF, VAR_IB, VAR_IF, VAR_ALL = 1, 2, 4, 8, 16, 31

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def build(force=False):
    """Compile the C restatement (gcc, seconds). Building the checker is not using it."""
    src = os.path.join(_HERE, "tfr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _SvdState(C.Structure):
    _fields_ = [("user_num", C.c_int32), ("item_num", C.c_int32), ("dim", C.c_int32), ("flags", C.c_int32),
                ("mu", _f32p), ("user_bias", _f32p), ("item_bias", _f32p), ("user_feat", _f32p), ("item_feat", _f32p),
                ("m_mu", _f32p), ("v_mu", _f32p), ("m_ub", _f32p), ("v_ub", _f32p), ("m_ib", _f32p), ("v_ib", _f32p),
                ("m_uf", _f32p), ("v_uf", _f32p), ("m_if", _f32p), ("v_if", _f32p),
                ("beta1_power", C.c_float), ("beta2_power", C.c_float), ("global_step", C.c_int64),
                ("lr", C.c_float), ("reg", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("var_mask", C.c_int32)]


class _FmState(C.Structure):
    _fields_ = [("n_feat", C.c_int32), ("dim", C.c_int32), ("flags", C.c_int32),
                ("w0", _f32p), ("W", _f32p), ("V", _f32p),
                ("m_w0", _f32p), ("v_w0", _f32p), ("m_W", _f32p), ("v_W", _f32p), ("m_V", _f32p), ("v_V", _f32p),
                ("beta1_power", C.c_float), ("beta2_power", C.c_float), ("global_step", C.c_int64),
                ("lr", C.c_float), ("reg", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_svd_forward.argtypes = [C.POINTER(_SvdState), _i32p, _i32p, C.c_int64, _f32p, _f32p]
        L.orc_svd_forward.restype = None
        L.orc_svd_data_loss.argtypes = [C.POINTER(_SvdState), _f32p, _f32p, C.c_int64]
        L.orc_svd_data_loss.restype = C.c_double
        L.orc_svd_regularizer.argtypes = [C.POINTER(_SvdState), _i32p, _i32p, C.c_int64]
        L.orc_svd_regularizer.restype = C.c_double
        L.orc_svd_grads.argtypes = [C.POINTER(_SvdState), _i32p, _i32p, _f32p, C.c_int64, _f32p, _f32p,
                                    _f32p, _f32p, _f32p, _f32p, _f32p]
        L.orc_svd_grads.restype = None
        L.orc_unique_first_occurrence.argtypes = [_i32p, C.c_int64, _i32p, _i32p]
        L.orc_unique_first_occurrence.restype = C.c_int64
        L.orc_segment_sum.argtypes = [_f32p, _i32p, C.c_int64, C.c_int32, C.c_int64, _f32p]
        L.orc_segment_sum.restype = None
        L.orc_set_segment_order.argtypes = [C.c_int]
        L.orc_set_segment_order.restype = None
        L.orc_adam_lr_t.argtypes = [C.c_float, C.c_float, C.c_float]
        L.orc_adam_lr_t.restype = C.c_float
        L.orc_adam_sparse.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_int32, _i32p, C.c_int64, _f32p,
                                      C.c_float, C.c_float, C.c_float, C.c_float]
        L.orc_adam_sparse.restype = None
        L.orc_adam_dense.argtypes = [_f32p, _f32p, _f32p, C.c_int64, _f32p, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_float, C.c_float]
        L.orc_adam_dense.restype = None
        L.orc_sgd_scatter.argtypes = [_f32p, C.c_int32, _i32p, C.c_int64, _f32p, C.c_float]
        L.orc_sgd_scatter.restype = None
        L.orc_svd_train_step.argtypes = [C.POINTER(_SvdState), _i32p, _i32p, _f32p, C.c_int64, _f32p, _f32p]
        L.orc_svd_train_step.restype = C.c_int
        L.orc_fm_forward.argtypes = [C.c_int64, _i64p, _i32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, _f32p, _f32p]
        L.orc_fm_forward.restype = None
        L.orc_fm_train_step.argtypes = [C.POINTER(_FmState), C.c_int64, _i64p, _i32p, _f32p, _f32p, _f32p]
        L.orc_fm_train_step.restype = C.c_int
        L.orc_allpairs.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_float, C.c_int64, C.c_int64, C.c_int32, _f32p]
        L.orc_allpairs.restype = None
        _lib = L
    return _lib


def set_segment_order(order):
    """0 = batch order (TF's CPU unsorted_segment_sum, the default), 1 = reverse batch order (another legitimate fp32
    evaluation of the same sums).  Process-wide; tests restore 0."""
    lib().orc_set_segment_order(int(order))


def _p(a, t=_f32p):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class SvdOracle:
    """The five variables of ops.py:8-12,29-32 plus TF Adam slots, stepped on the CPU.

    Initial tables are INJECTED (TF's Philox truncated_normal stream is not reproducible outside
    TF): pass numpy arrays; they are copied.
    """

    def __init__(self, mu, user_bias, item_bias, user_feat, item_feat, lr, reg, flags=0,
                 beta1=0.9, beta2=0.999, eps=1e-8, var_mask=VAR_ALL):
        self.mu = _f32(np.array(mu).reshape(1)).copy()
        self.user_bias = _f32(user_bias).copy()
        self.item_bias = _f32(item_bias).copy()
        self.user_feat = _f32(user_feat).copy()
        self.item_feat = _f32(item_feat).copy()
        self.U, self.d = self.user_feat.shape
        self.I = self.item_feat.shape[0]
        self.slots = {}
        for name in ("mu", "user_bias", "item_bias", "user_feat", "item_feat"):
            self.slots["m_" + name] = np.zeros_like(getattr(self, name))
            self.slots["v_" + name] = np.zeros_like(getattr(self, name))
        s = _SvdState()
        s.user_num, s.item_num, s.dim, s.flags = self.U, self.I, self.d, flags
        s.mu, s.user_bias, s.item_bias = _p(self.mu), _p(self.user_bias), _p(self.item_bias)
        s.user_feat, s.item_feat = _p(self.user_feat), _p(self.item_feat)
        sl = self.slots
        s.m_mu, s.v_mu = _p(sl["m_mu"]), _p(sl["v_mu"])
        s.m_ub, s.v_ub = _p(sl["m_user_bias"]), _p(sl["v_user_bias"])
        s.m_ib, s.v_ib = _p(sl["m_item_bias"]), _p(sl["v_item_bias"])
        s.m_uf, s.v_uf = _p(sl["m_user_feat"]), _p(sl["v_user_feat"])
        s.m_if, s.v_if = _p(sl["m_item_feat"]), _p(sl["v_item_feat"])
        s.beta1_power, s.beta2_power = beta1, beta2
        s.global_step = 0
        s.lr, s.reg, s.beta1, s.beta2, s.eps = lr, reg, beta1, beta2, eps
        s.var_mask = var_mask
        self.s = s
        self.flags = flags

    def forward(self, users, items):
        users, items = _i32(users), _i32(items)
        B = len(users)
        logits = np.empty(B, np.float32)
        infer = np.empty(B, np.float32)
        lib().orc_svd_forward(C.byref(self.s), _p(users, _i32p), _p(items, _i32p), B, _p(logits), _p(infer))
        return logits, infer

    def grads(self, users, items, rates):
        users, items, rates = _i32(users), _i32(items), _f32(rates)
        B, d = len(users), self.d
        out = dict(logits=np.empty(B, np.float32), err=np.empty(B, np.float32),
                   g_uf=np.empty((B, d), np.float32), g_if=np.empty((B, d), np.float32),
                   g_ub=np.empty(B, np.float32), g_ib=np.empty(B, np.float32), g_mu=np.zeros(1, np.float32))
        lib().orc_svd_grads(C.byref(self.s), _p(users, _i32p), _p(items, _i32p), _p(rates), B,
                            _p(out["logits"]), _p(out["err"]), _p(out["g_uf"]), _p(out["g_if"]),
                            _p(out["g_ub"]), _p(out["g_ib"]), _p(out["g_mu"]))
        return out

    def data_loss(self, logits, rates):
        logits, rates = _f32(logits), _f32(rates)
        return lib().orc_svd_data_loss(C.byref(self.s), _p(logits), _p(rates), len(logits))

    def regularizer(self, users, items):
        users, items = _i32(users), _i32(items)
        return lib().orc_svd_regularizer(C.byref(self.s), _p(users, _i32p), _p(items, _i32p), len(users))

    def train_step(self, users, items, rates):
        """svd_train_val.py:70-72: returns (logits, infer) from PRE-update parameters."""
        users, items, rates = _i32(users), _i32(items), _f32(rates)
        B = len(users)
        logits = np.empty(B, np.float32)
        infer = np.empty(B, np.float32)
        rc = lib().orc_svd_train_step(C.byref(self.s), _p(users, _i32p), _p(items, _i32p), _p(rates), B,
                                      _p(logits), _p(infer))
        assert rc == 0
        return logits, infer

    @property
    def global_step(self):
        return self.s.global_step


def unique_first_occurrence(ids):
    """tf.unique (A.3): (unique ids in first-occurrence order, idx into them)."""
    ids = _i32(ids)
    B = len(ids)
    uq = np.empty(max(B, 1), np.int32)
    idx = np.empty(max(B, 1), np.int32)
    n = lib().orc_unique_first_occurrence(_p(ids, _i32p), B, _p(uq, _i32p), _p(idx, _i32p))
    return uq[:n].copy(), idx[:B].copy()


def segment_sum(values, idx, n_uniq):
    values = _f32(values)
    v2 = values.reshape(len(values), -1)
    idx = _i32(idx)
    out = np.empty((n_uniq, v2.shape[1]), np.float32)
    lib().orc_segment_sum(_p(v2), _p(idx, _i32p), len(idx), v2.shape[1], n_uniq, _p(out))
    return out.reshape((n_uniq,) + values.shape[1:])


def adam_lr_t(lr, b1p, b2p):
    return lib().orc_adam_lr_t(lr, b1p, b2p)


def adam_sparse(var, m, v, uniq, gsum, lr_t, beta1=0.9, beta2=0.999, eps=1e-8):
    """In place on float32 C-contiguous var/m/v of shape [rows] or [rows, width]."""
    for a in (var, m, v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    rows = var.shape[0]
    width = 1 if var.ndim == 1 else var.shape[1]
    uniq, gsum = _i32(uniq), _f32(gsum)
    lib().orc_adam_sparse(_p(var), _p(m), _p(v), rows, width, _p(uniq, _i32p), len(uniq), _p(gsum),
                          lr_t, beta1, beta2, eps)


def adam_dense(var, m, v, g, lr, b1p, b2p, beta1=0.9, beta2=0.999, eps=1e-8):
    g = _f32(g)
    lib().orc_adam_dense(_p(var), _p(m), _p(v), var.size, _p(g), lr, b1p, b2p, beta1, beta2, eps)


def sgd_scatter(var, ids, values, lr):
    width = 1 if var.ndim == 1 else var.shape[1]
    ids, values = _i32(ids), _f32(values)
    lib().orc_sgd_scatter(_p(var), width, _p(ids, _i32p), len(ids), _p(values), lr)


def fm_forward(indptr, indices, data, w0, W, V, return_sums=False):
    """forward.py:21-22 on CSR rows."""
    indptr = np.ascontiguousarray(indptr, np.int64)
    indices, data = _i32(indices), _f32(data)
    w0 = _f32(np.array(w0).reshape(1))
    W, V = _f32(W), _f32(V)
    n = len(indptr) - 1
    y = np.empty(n, np.float32)
    sums = np.empty((n, V.shape[1]), np.float32) if return_sums else None
    lib().orc_fm_forward(n, _p(indptr, _i64p), _p(indices, _i32p), _p(data), _p(w0), _p(W), _p(V), V.shape[1],
                         _p(y), _p(sums))
    return (y, sums) if return_sums else y


class FmOracle:
    def __init__(self, w0, W, V, lr, reg, flags=0, beta1=0.9, beta2=0.999, eps=1e-8):
        self.w0 = _f32(np.array(w0).reshape(1)).copy()
        self.W = _f32(W).copy()
        self.V = _f32(V).copy()
        self.F, self.d = self.V.shape
        self.slots = {k + n: np.zeros_like(getattr(self, n)) for n in ("w0", "W", "V") for k in ("m_", "v_")}
        s = _FmState()
        s.n_feat, s.dim, s.flags = self.F, self.d, flags
        s.w0, s.W, s.V = _p(self.w0), _p(self.W), _p(self.V)
        s.m_w0, s.v_w0 = _p(self.slots["m_w0"]), _p(self.slots["v_w0"])
        s.m_W, s.v_W = _p(self.slots["m_W"]), _p(self.slots["v_W"])
        s.m_V, s.v_V = _p(self.slots["m_V"]), _p(self.slots["v_V"])
        s.beta1_power, s.beta2_power, s.global_step = beta1, beta2, 0
        s.lr, s.reg, s.beta1, s.beta2, s.eps = lr, reg, beta1, beta2, eps
        self.s = s

    def forward(self, indptr, indices, data):
        return fm_forward(indptr, indices, data, self.w0, self.W, self.V)

    def train_step(self, indptr, indices, data, y):
        indptr = np.ascontiguousarray(indptr, np.int64)
        indices, data, y = _i32(indices), _f32(data), _f32(y)
        n = len(indptr) - 1
        yhat = np.empty(n, np.float32)
        rc = lib().orc_fm_train_step(C.byref(self.s), n, _p(indptr, _i64p), _p(indices, _i32p), _p(data), _p(y),
                                     _p(yhat))
        assert rc == 0
        return yhat


def allpairs(U, V, wu, wi, mu):
    """als3.py:110-113."""
    U, V, wu, wi = _f32(U), _f32(V), _f32(wu), _f32(wi)
    M = np.empty((U.shape[0], V.shape[0]), np.float32)
    lib().orc_allpairs(_p(U), _p(V), _p(wu), _p(wi), float(mu), U.shape[0], V.shape[0], U.shape[1], _p(M))
    return M
