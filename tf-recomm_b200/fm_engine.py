"""FmEngine: device-resident 2nd-order factorization machine (forward.py:21-22's model: w0, W [F], V [F, dim]) and
its train step on CSR batches -- the model fm.py:104-110,154-155 hands to libFM, trained here with the SVD path's own
step (squared error or sigmoid cross-entropy + L2 on the gathered rows, tf.unique dedup of the batch's feature ids,
TF-semantics Adam over the whole tables) instead of libFM's MCMC sampler (external binary, out of scope).

Drives libtfrecomm.so through ctypes; torch is plumbing (device memory, streams).  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FM_PRESORTED, OPT_SGD, FmTables, OptScalars, StepWs, TfrError, check

FM_TABLE_NAMES = ("w0", "W", "V")


class FmEngine:
    def __init__(self, n_feat, dim, lr, reg, flags=0, beta1=0.9, beta2=0.999, eps=1e-8, tables=None, device=None,
                 seed=13575, init_stdev=0.1):
        """tables: dict(w0, W, V) of numpy arrays to inject; default: w0 = 0, W = 0, V ~ N(0, init_stdev^2) (libFM's
        init_stdev default 0.1, pywFM.FM(init_stdev=0.1)), drawn with numpy so that an oracle can be given the same."""
        if not torch.cuda.is_available():
            raise TfrError("tf-recomm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback for this path")
        self.L = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.F, self.d = int(n_feat), int(dim)
        self.flags = int(flags)
        self.sgd = bool(self.flags & OPT_SGD)
        if tables is None:
            tables = self.init_tables(self.F, self.d, seed, init_stdev)
        dev = self.device
        with torch.cuda.device(dev):
            self.t = {
                "w0": torch.from_numpy(np.ascontiguousarray(tables["w0"], np.float32).reshape(1)).to(dev),
                "W": torch.from_numpy(np.ascontiguousarray(tables["W"], np.float32).reshape(self.F)).to(dev),
                "V": torch.from_numpy(np.ascontiguousarray(tables["V"], np.float32).reshape(self.F, self.d)).to(dev)}
            self.slots = {}
            if not self.sgd:
                for n in FM_TABLE_NAMES:
                    self.slots["m_" + n] = torch.zeros_like(self.t[n])
                    self.slots["v_" + n] = torch.zeros_like(self.t[n])
            self.slot = torch.full((self.F,), -1, dtype=torch.int64, device=dev)
            self.opt = torch.zeros(C.sizeof(OptScalars), dtype=torch.uint8, device=dev)
            s = FmTables()
            s.n_feat, s.dim = self.F, self.d
            s.w0, s.W, s.V = (self.t[n].data_ptr() for n in FM_TABLE_NAMES)
            if not self.sgd:
                s.m_w0, s.v_w0 = self.slots["m_w0"].data_ptr(), self.slots["v_w0"].data_ptr()
                s.m_W, s.v_W = self.slots["m_W"].data_ptr(), self.slots["v_W"].data_ptr()
                s.m_V, s.v_V = self.slots["m_V"].data_ptr(), self.slots["v_V"].data_ptr()
            s.slot = self.slot.data_ptr()
            self.tables_struct = s
            check(self.L.tfr_opt_init(self.opt.data_ptr(), lr, reg, beta1, beta2, eps, self.flags, _lib.VAR_ALL,
                                      self._stream()))
        self._scratch = {}

    @staticmethod
    def init_tables(n_feat, dim, seed=13575, init_stdev=0.1):
        rng = np.random.default_rng(seed)
        return dict(w0=np.zeros(1, np.float32), W=np.zeros(n_feat, np.float32),
                    V=(rng.standard_normal((n_feat, dim)) * init_stdev).astype(np.float32))

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- CSR batches ------------------------------------------------------------------------------------------
    def upload_csr(self, X, y=None):
        """scipy CSR (or (indptr, indices, data)) -> dict of device tensors in the ABI's dtypes."""
        if isinstance(X, tuple):
            indptr, indices, data = X
        else:
            X = X.tocsr()
            indptr, indices, data = X.indptr, X.indices, X.data
        indptr = np.ascontiguousarray(indptr, np.int64)
        if indptr[0] != 0:
            raise TfrError("a CSR batch must start at indptr[0] == 0")
        dev = self.device
        b = dict(indptr=torch.from_numpy(indptr).to(dev),
                 indices=torch.from_numpy(np.ascontiguousarray(indices, np.int32)).to(dev),
                 data=torch.from_numpy(np.nan_to_num(np.ascontiguousarray(data, np.float32))).to(dev),  # fm.py:125,129
                 n_rows=len(indptr) - 1, nnz=int(indptr[-1]))
        if len(indices) and (int(np.max(indices)) >= self.F or int(np.min(indices)) < 0):
            raise TfrError("feature index out of range [0, %d)" % self.F)
        if y is not None:
            b["y"] = torch.from_numpy(np.ascontiguousarray(y, np.float32)).to(dev)
        return b

    def _bufs(self, n_rows, nnz):
        key = (n_rows, nnz)
        s = self._scratch.get(key)
        if s is None:
            dev = self.device
            nbytes = check(self.L.tfr_svd_step_workspace_bytes(max(nnz, 1), self.d))
            s = dict(yhat=torch.empty(n_rows, dtype=torch.float32, device=dev),
                     sums=torch.empty(n_rows, self.d, dtype=torch.float32, device=dev),
                     err=torch.empty(n_rows, dtype=torch.float32, device=dev),
                     rowof=torch.empty(max(nnz, 1), dtype=torch.int32, device=dev),
                     ws=torch.empty(nbytes, dtype=torch.uint8, device=dev))
            self._scratch[key] = s
        return s

    # ---- forward: forward.py:21-22 ----------------------------------------------------------------------------
    def forward(self, batch):
        if not isinstance(batch, dict):
            batch = self.upload_csr(batch)
        n = batch["n_rows"]
        yhat = torch.empty(n, dtype=torch.float32, device=self.device)
        if n == 0:
            return yhat
        with torch.cuda.device(self.device):
            check(self.L.tfr_fm_forward(n, batch["indptr"].data_ptr(), batch["indices"].data_ptr(),
                                        batch["data"].data_ptr(), self.t["w0"].data_ptr(), self.t["W"].data_ptr(),
                                        self.t["V"].data_ptr(), self.d, yhat.data_ptr(), None, self._stream()))
        return yhat

    # ---- one train step on a CSR batch; returns yhat of the PRE-update tables ------------------------------------
    def _enqueue_step(self, batch, s, presorted=False):
        flags = self.flags | (FM_PRESORTED if presorted else 0)
        check(self.L.tfr_fm_train_step(C.byref(self.tables_struct), self.opt.data_ptr(), batch["n_rows"],
                                       batch["indptr"].data_ptr(), batch["indices"].data_ptr(),
                                       batch["data"].data_ptr(), s["rowof"].data_ptr(), batch["nnz"],
                                       batch["y"].data_ptr(), s["yhat"].data_ptr(), s["sums"].data_ptr(),
                                       s["err"].data_ptr(), flags, s["ws"].data_ptr(), s["ws"].numel(),
                                       self._stream()))

    def _sort_once(self, batch, s):
        """The batch's (feature id, position) pairs, sorted into its own scratch: a fixed CSR batch is sorted once,
        every later step on it (tfr_fm_train_step with TFR_FM_PRESORTED) starts at the forward."""
        ws = StepWs()
        check(self.L.tfr_svd_step_carve(s["ws"].data_ptr(), s["ws"].numel(), batch["nnz"], self.d, C.byref(ws)))
        check(self.L.tfr_dedup_sort_pairs(batch["indices"].data_ptr(), self.F + 1, ws.su_ids, ws.su_pos, None, 1, None,
                                          None, batch["nnz"], ws.sort_ws, ws.sort_ws_bytes, self._stream()))

    def train_step(self, batch, y=None):
        if not isinstance(batch, dict):
            batch = self.upload_csr(batch, y)
        if batch["n_rows"] == 0 or batch["nnz"] == 0:
            raise TfrError("empty FM batch")
        s = self._bufs(batch["n_rows"], batch["nnz"])
        with torch.cuda.device(self.device):
            self._enqueue_step(batch, s)
        return s["yhat"]

    def run_epoch(self, batches, use_graph=True):
        """One sweep over pre-uploaded CSR batches (the same chunks every epoch, like libFM's in-order sweep,
        fm.py:154-155).  With use_graph each batch's step is captured once (with scratch buffers of its own: a
        captured graph bakes addresses) and replayed on later epochs.  Predictions: forward()."""
        with torch.cuda.device(self.device):
            for b in batches:
                if not use_graph:
                    self._enqueue_step(b, self._bufs(b["n_rows"], b["nnz"]))
                    continue
                g = b.get("_graph")
                if g is None:
                    proto = self._bufs(b["n_rows"], b["nnz"])
                    sc = b["_scratch"] = {n: torch.empty_like(v) for n, v in proto.items()}
                    self._sort_once(b, sc)
                    cap = torch.cuda.Stream(device=self.device)
                    cap.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(cap):
                        check(self.L.tfr_graph_begin_capture(cap.cuda_stream))
                        try:
                            self._enqueue_step(b, sc, presorted=True)
                        finally:
                            exe = C.c_void_p()
                            rc = self.L.tfr_graph_end_capture(cap.cuda_stream, C.byref(exe))
                        check(rc)
                    torch.cuda.current_stream(self.device).wait_stream(cap)
                    g = b["_graph"] = exe
                check(self.L.tfr_graph_launch(g, self._stream()))

    # ---- state ---------------------------------------------------------------------------------------------------
    def get_tables(self):
        out = {n: self.t[n].detach().cpu().numpy().copy() for n in FM_TABLE_NAMES}
        for k, v in self.slots.items():
            out[k] = v.detach().cpu().numpy().copy()
        return out

    def opt_scalars(self):
        return OptScalars.from_buffer_copy(self.opt.cpu().numpy().tobytes())

    @property
    def global_step(self):
        return int(self.opt_scalars().global_step)

    def save(self, path):
        """fm_mangaki.py:39-45 pickles {mu, W, V}; np.save('vectors-d.npy', pairwise_interactions) fm.py:159."""
        arrays = self.get_tables()
        s = self.opt_scalars()
        arrays["__opt__"] = np.array([s.beta1_power, s.beta2_power], np.float32)
        arrays["__step__"] = np.array([s.global_step], np.int64)
        with open(path, "wb") as f:
            np.savez(f, **arrays)
