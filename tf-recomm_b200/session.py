"""A deliberately small stand-in for the slice of the TensorFlow 1.x session API that the reference's driver
uses around the hot path (svd_train_val.py:40-57,70-72,94,121-122,197-198): placeholders, variable initialisation,
global_step, Saver and Session.run with its three fetch patterns

    sess.run([train_op, logits, infer], feed_dict={user, item, rate[, wins, fails]})   -> fused train step,
                                                                  predictions from PRE-update tables
    sess.run([logits, infer], feed_dict={user, item[, wins, fails]})                   -> forward only
    sess.run(cost, feed_dict={rate, logits|infer: array})                              -> loss from fed predictions

There is no graph compiler here: ops.inference_svd / ops.optimization record WHAT to run, the session maps each
run onto calls into libtfrecomm.so.  Anything outside these patterns raises -- loudly -- instead of pretending
to be TensorFlow.
"""
import numpy as np

from . import _lib, init
from .engine import SvdEngine


class Placeholder(object):
    def __init__(self, dtype=None, shape=None, name=None):
        self.dtype, self.shape, self.name = dtype, shape, name

    def __repr__(self):
        return "<placeholder %s>" % (self.name or hex(id(self)))


int32 = "int32"
float32 = "float32"


def placeholder(dtype=None, shape=None, name=None):
    return Placeholder(dtype, shape, name)


class Handle(object):
    """A fetchable node of the model: 'infer', 'logits', 'regularizer', 'cost', 'train_op', 'init',
    'global_step' or a variable ('mu', 'user_bias', 'item_bias', 'user_feat', 'item_feat')."""

    def __init__(self, model, kind):
        self.model, self.kind = model, kind

    def __repr__(self):
        return "<tfrecomm %s>" % self.kind


VAR_BITS = {"mu": _lib.VAR_MU, "user_bias": _lib.VAR_UB, "user_feat": _lib.VAR_UF, "item_bias": _lib.VAR_IB,
            "item_feat": _lib.VAR_IF}


class Model(object):
    """What inference_svd / optimization declared; the engine is created by the initializer run."""

    def __init__(self, user_batch, item_batch, wins_batch, fails_batch, user_num, item_num, dim, variant):
        self.user_batch, self.item_batch = user_batch, item_batch
        self.wins_batch, self.fails_batch = wins_batch, fails_batch
        self.user_num, self.item_num, self.dim = int(user_num), int(item_num), int(dim)
        self.variant = variant
        self.flags = _lib.README_FLAGS if variant == "readme" else (_lib.FORK_FLAGS & ~_lib.OPT_SGD)
        self.rate_batch = None
        self.lr = self.reg = None
        self.var_mask = _lib.VAR_ALL
        self.engine = None
        self.init_tables = None      # inject numpy tables before the initializer runs (parity harness)
        self.seed = 13575
        self.h = {k: Handle(self, k) for k in ("infer", "logits", "regularizer", "cost", "train_op",
                                               "mu", "user_bias", "item_bias", "user_feat", "item_feat")}

    def create_engine(self):
        if self.lr is None:
            raise _lib.TfrError("ops.optimization must be called before variables are initialised")
        tabs = self.init_tables
        if tabs is None:
            bias_init = "truncated_normal" if self.variant == "fork" else "glorot"
            tabs = init.init_tables(self.user_num, self.item_num, self.dim, seed=self.seed, bias_init=bias_init)
        self.engine = SvdEngine(self.user_num, self.item_num, self.dim, self.lr, self.reg, flags=self.flags,
                                var_mask=self.var_mask, tables=tabs)
        return self.engine


_state = {"model": None, "global_step": None}


def reset_default_graph():
    _state["model"] = None
    _state["global_step"] = None


def current_model():
    if _state["model"] is None:
        raise _lib.TfrError("no model: call ops.inference_svd first")
    return _state["model"]


def _set_model(m):
    _state["model"] = m


class _Train(object):
    """tf.train.* used by the driver (svd_train_val.py:48,54)."""

    @staticmethod
    def get_or_create_global_step():
        if _state["global_step"] is None:
            _state["global_step"] = Handle(None, "global_step")
        return _state["global_step"]

    @staticmethod
    def get_global_step():
        return _state["global_step"]

    class Saver(object):
        """saver.save(sess, path) / saver.restore(sess, path): all five variables, Adam slots, beta powers
        and global_step in one .npz (svd_train_val.py:54,197-198; adaptive_test.py:40)."""

        def save(self, sess, path):
            current_model().engine.save(path)
            return path

        def restore(self, sess, path):
            m = current_model()
            if m.engine is None:
                m.create_engine()
            m.engine.restore(path)


train = _Train()


def global_variables_initializer():
    return Handle(None, "init")


local_variables_initializer = global_variables_initializer


def group(*ops):
    return Handle(None, "init")


class Session(object):
    def __init__(self, *a, **k):
        self.graph = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    # -- helpers ------------------------------------------------------------------------------------------------
    @staticmethod
    def _fed(feed, ph, what):
        if ph is None or ph not in feed:
            raise _lib.TfrError("fetch needs %s in feed_dict" % what)
        return feed[ph]

    def prefetch(self, feed_dict):
        """Not in TensorFlow's API: hands a COMING train step's feed_dict over early (up to four may be pending; the driver
        hands over three ahead), so that its host packing runs on the feed worker thread and its copy to the device and
        id sort run under the table pass of the step before it (SvdEngine.prefetch_host).  The handed-over batches must
        then be stepped in that order -- run([train_op, ...], feed_dict) with the same arrays; anything else just drops
        what is pending."""
        m = current_model()
        if m.engine is None:
            raise _lib.TfrError("variables are not initialised: run the initializer op first")
        m.engine.prefetch_host(self._fed(feed_dict, m.user_batch, "user_batch"),
                               self._fed(feed_dict, m.item_batch, "item_batch"),
                               self._fed(feed_dict, m.rate_batch, "rate_batch"))

    def run(self, fetches, feed_dict=None):
        feed = feed_dict or {}
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        for f in fl:
            if not isinstance(f, Handle):
                raise _lib.TfrError("cannot fetch %r: only handles returned by ops.inference_svd / ops.optimization"
                                    % (f,))
        kinds = [f.kind for f in fl]
        out = {}
        if "init" in kinds:
            m = current_model()
            if m.engine is None:
                m.create_engine()
            out["init"] = None
        m = _state["model"]
        compute = [k for k in kinds if k not in ("init", "global_step")]
        if compute:
            m = current_model()
            if m.engine is None:
                raise _lib.TfrError("variables are not initialised: run the initializer op first")
            eng = m.engine
            if "train_op" in kinds:
                users = self._fed(feed, m.user_batch, "user_batch")
                items = self._fed(feed, m.item_batch, "item_batch")
                rates = self._fed(feed, m.rate_batch, "rate_batch")
                fetch_pred = any(k in ("logits", "infer") for k in kinds)
                if any(k in ("cost", "regularizer") for k in kinds):
                    out["regularizer"], out["cost"] = self._scalars(m, users, items, rates)
                res = eng.train_step_host(users, items, rates, fetch=fetch_pred)
                out["train_op"] = None
                if fetch_pred:
                    out["logits"], out["infer"] = res
            else:
                fed_pred = m.h["logits"] in feed or m.h["infer"] in feed
                if fed_pred:
                    # metrics re-evaluation with a fed intermediate tensor (svd_train_val.py:94,100,138)
                    if kinds != ["cost"]:
                        raise _lib.TfrError("with logits/infer fed only `cost` can be fetched")
                    rates = np.asarray(self._fed(feed, m.rate_batch, "rate_batch"), np.float32)
                    pred = np.asarray(feed.get(m.h["logits"], feed.get(m.h["infer"])), np.float32)
                    out["cost"] = data_loss(pred, rates, m.flags)
                else:
                    users = self._fed(feed, m.user_batch, "user_batch")
                    items = self._fed(feed, m.item_batch, "item_batch")
                    if any(k in ("logits", "infer") for k in kinds):
                        lg, inf = eng.forward(users, items)
                        out["logits"], out["infer"] = lg.cpu().numpy(), inf.cpu().numpy()
                    if any(k in ("cost", "regularizer") for k in kinds):
                        rates = feed.get(m.rate_batch)
                        out["regularizer"], out["cost"] = self._scalars(m, users, items, rates)
            for k in kinds:
                if k in VAR_BITS:
                    out[k] = eng.variable(k).detach().cpu().numpy().copy()
                    if k == "mu":
                        out[k] = out[k].reshape(())
        if "global_step" in kinds:
            out["global_step"] = m.engine.global_step if (m is not None and m.engine is not None) else 0
        res = [out[k] for k in kinds]
        return res[0] if single else res

    @staticmethod
    def _scalars(m, users, items, rates):
        """regularizer (ops.py:81-89) and cost (ops.py:124-126,140) scalars, fp32 like TF's, from the current
        tables.  Only evaluated when a driver fetches them; the train step itself never needs them."""
        eng = m.engine
        u = np.asarray(users).astype(np.int64)
        i = np.asarray(items).astype(np.int64)
        import torch
        ut = torch.from_numpy(u).to(eng.device)
        it = torch.from_numpy(i).to(eng.device)
        pu, qi = eng.t["user_feat"][ut], eng.t["item_feat"][it]
        regl = 0.5 * (pu * pu).sum() + 0.5 * (qi * qi).sum()
        if m.flags & _lib.REG_BIAS:
            regl = regl + 0.5 * (eng.t["user_bias"][ut] ** 2).sum() + 0.5 * (eng.t["item_bias"][it] ** 2).sum()
        regl = float(regl)
        cost = None
        if rates is not None:
            lg, _ = eng.forward(users, items)
            cost = data_loss(lg.cpu().numpy(), np.asarray(rates, np.float32), m.flags)
        return np.float32(regl), cost


def data_loss(pred, rates, flags):
    """cost_l2 = l2_loss(infer - rate) (ops.py:124) / cost_nll = sum sigmoid-CE (ops.py:125-126), fp32."""
    pred = np.asarray(pred, np.float32)
    rates = np.asarray(rates, np.float32)
    if flags & _lib.LOSS_SIGMOID_CE:
        x, z = pred, rates
        return np.float32(np.sum(np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x))), dtype=np.float32))
    e = pred - rates
    return np.float32(np.sum(e * e, dtype=np.float32) / np.float32(2))
