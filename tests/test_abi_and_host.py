"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/tfrecomm.h declares
(no compute calls -- there is no GPU here), the ctypes mirrors match the header's struct layouts, the product path
fails loudly without CUDA and never imports the oracle, and the host-side mirror (ops / session / config) keeps
the reference's call shapes."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from tf_recomm_b200 import _lib, config, init, ops, session

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tfrecomm.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), "libtfrecomm.so does not export %s" % name
    assert sorted(_lib.exported_symbols()) == declared, "ctypes prototypes and header disagree"
    assert L.tfr_abi_version() == 1


def test_struct_layouts_match_header():
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "tfrecomm.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(tfr_opt_scalars), sizeof(tfr_svd_tables), sizeof(tfr_svd_step_ws),
         sizeof(tfr_adam_table), sizeof(tfr_slice_update));
  printf("%zu %zu %zu %zu\n", offsetof(tfr_opt_scalars, global_step), offsetof(tfr_opt_scalars, batch_cursor),
         offsetof(tfr_opt_scalars, se_ring), offsetof(tfr_opt_scalars, timeline));
  printf("%zu %zu %zu\n", offsetof(tfr_svd_tables, mu), offsetof(tfr_svd_tables, user_slot), offsetof(tfr_svd_step_ws, sort_ws));
  printf("%zu %zu %zu\n", sizeof(tfr_fm_tables), offsetof(tfr_fm_tables, w0), offsetof(tfr_fm_tables, slot));
  printf("%zu %zu %zu %zu\n", sizeof(tfr_feed_set), offsetof(tfr_feed_set, workspace_bytes), offsetof(tfr_feed_set, ev_done),
         offsetof(tfr_feed_set, copied));
  printf("%zu\n", offsetof(tfr_opt_scalars, prefetch_cursor));
  printf("%zu %zu %zu %zu\n", offsetof(tfr_feed_set, staged), offsetof(tfr_feed_set, d_sync), offsetof(tfr_feed_set, h_flag),
         offsetof(tfr_feed_set, deliver_seq));
  return 0;
}'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    got = [int(x) for x in out]
    exp = [C.sizeof(_lib.OptScalars), C.sizeof(_lib.SvdTables), C.sizeof(_lib.StepWs), C.sizeof(_lib.AdamTable),
           C.sizeof(_lib.SliceUpdate), _lib.OptScalars.global_step.offset, _lib.OptScalars.batch_cursor.offset,
           _lib.OptScalars.se_ring.offset, _lib.OptScalars.timeline.offset, _lib.SvdTables.mu.offset,
           _lib.SvdTables.user_slot.offset, _lib.StepWs.sort_ws.offset, C.sizeof(_lib.FmTables),
           _lib.FmTables.w0.offset, _lib.FmTables.slot.offset, C.sizeof(_lib.FeedSet), _lib.FeedSet.workspace_bytes.offset,
           _lib.FeedSet.ev_done.offset, _lib.FeedSet.copied.offset, _lib.OptScalars.prefetch_cursor.offset,
           _lib.FeedSet.staged.offset, _lib.FeedSet.d_sync.offset, _lib.FeedSet.h_flag.offset, _lib.FeedSet.deliver_seq.offset]
    assert got == exp


def test_workspace_queries_need_no_gpu():
    L = _lib.load()
    assert L.tfr_svd_step_workspace_bytes(65536, 128) > 2 * 65536 * 128 * 4
    assert L.tfr_svd_step_workspace_bytes(0, 128) < 0
    assert L.tfr_dedup_workspace_bytes(1000) > 4 * 1000 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    from tf_recomm_b200.engine import SvdEngine
    with pytest.raises(_lib.TfrError, match="no CPU fallback"):
        SvdEngine(10, 10, 4, 1e-3, 0.05, tables=init.init_tables(10, 10, 4))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tf-recomm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "tfr_oracle" not in text, f
    for f in ("ops.py", "dataio.py", "config.py", "svd_train_val.py", "fm.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert not re.search(r"^\s*(import|from)\s+oracle\b", open(p).read(), flags=re.M), f


def test_inference_svd_and_optimization_call_shapes():
    """Current shape (ops.py:6,91,118,153) and README-era shape (svd_train_val.py:49) both build the same model."""
    ops.reset_default_graph()
    u, i, r = ops.placeholder(ops.int32, [None], "id_user"), ops.placeholder(ops.int32, [None], "id_item"), ops.placeholder(ops.float32, [None])
    w, f = ops.placeholder(ops.float32, [None]), ops.placeholder(ops.float32, [None])
    out = ops.inference_svd(u, i, w, f, user_num=50, item_num=30, dim=7, device="/cpu:0")
    assert len(out) == 7
    infer, logits, regularizer, user_bias, user_features, item_bias, item_features = out
    with pytest.raises(AssertionError):
        ops.optimization(infer, logits, regularizer, r, learning_rate=1e-3, reg=0.05)   # no global_step yet (ops.py:119-120)
    ops.train.get_or_create_global_step()
    cost, train_op = ops.optimization(infer, logits, regularizer, r, learning_rate=1e-3, reg=0.05, device="/cpu:0",
                                      var_list=[user_bias, user_features])
    m = session.current_model()
    assert (m.user_num, m.item_num, m.dim) == (50, 30, 7)
    assert m.var_mask == (_lib.VAR_UB | _lib.VAR_UF) and m.rate_batch is r and m.flags == _lib.README_FLAGS
    ops.reset_default_graph()
    infer2, reg2 = ops.inference_svd(u, i, 50, 30, 7, "/cpu:0")
    ops.train.get_or_create_global_step()
    cost2, train2 = ops.optimization(infer2, reg2, r, learning_rate=0.01, reg=0.1, device="/cpu:0")
    m2 = session.current_model()
    assert m2.lr == 0.01 and m2.reg == 0.1 and m2.var_mask == _lib.VAR_ALL
    ops.reset_default_graph()
    ops.inference_svd(u, i, w, f, user_num=5, item_num=3, dim=2, variant="fork")
    ops.train.get_or_create_global_step()
    ops.optimization(session.current_model().h["infer"], session.current_model().h["logits"],
                     session.current_model().h["regularizer"], r, 5e-3, 0.01)
    assert session.current_model().flags == _lib.FORK_FLAGS      # abs + sigmoid-CE + bias L2 + SGD (ops.py:44,85-89,125,145)
    with pytest.raises(ValueError):
        ops.inference_svd(u, i, w, f, user_num=5, item_num=3, variant="nope")
    with pytest.raises(TypeError):
        ops.inference_svd(u, i, w, f, item_num=3)


def test_session_rejects_unknown_fetches_and_uninitialised_runs():
    ops.reset_default_graph()
    u, i, r = ops.placeholder(ops.int32), ops.placeholder(ops.int32), ops.placeholder(ops.float32)
    infer, regl = ops.inference_svd(u, i, 5, 4, 3, "/cpu:0")
    ops.train.get_or_create_global_step()
    cost, train_op = ops.optimization(infer, regl, r, learning_rate=1e-3, reg=0.05)
    sess = ops.Session()
    with pytest.raises(_lib.TfrError, match="not initialised"):
        sess.run([train_op, infer], feed_dict={u: [0], i: [0], r: [1.0]})
    with pytest.raises(_lib.TfrError, match="cannot fetch"):
        sess.run("infer")
    # the loss-from-fed-predictions pattern needs no device (svd_train_val.py:94,100)
    m = session.current_model()
    m.engine = object()  # pretend initialised; this pattern never touches it
    got = sess.run(cost, feed_dict={r: np.array([1.0, 3.0], np.float32), infer: np.array([2.0, 1.0], np.float32)})
    assert got == pytest.approx(0.5 * (1 + 4))
    m.engine = None


def test_data_loss_matches_tf_definitions():
    x = np.array([-3.0, 0.0, 2.5], np.float32); z = np.array([0.0, 1.0, 1.0], np.float32)
    nll = session.data_loss(x, z, _lib.LOSS_SIGMOID_CE)
    ref = np.sum(np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x))))
    assert nll == pytest.approx(ref, rel=1e-6)
    assert session.data_loss(x, z, 0) == pytest.approx(0.5 * np.sum((x - z) ** 2), rel=1e-6)


def test_config_has_reference_names_and_the_missing_ones():
    for name in ("DIM", "EPOCH_MAX", "LEARNING_RATE", "LAMBDA_REG", "DISCRETE", "DEVICE", "PREFIX", "BASE_DIR",
                 "ARTICLE_FOLDER", "BATCH_SIZE", "USER_NUM", "ITEM_NUM", "NB_CLASSES", "MODEL_VARIANT"):
        assert hasattr(config, name), name
    import config as root_config
    import dataio as root_dataio
    import ops as root_ops
    assert root_config.DIM == config.DIM and root_ops.inference_svd is ops.inference_svd
    assert root_dataio.ShuffleIterator is __import__("tf_recomm_b200").dataio.ShuffleIterator


def test_init_tables_follow_tf_initialisers():
    t = init.init_tables(2000, 1000, 16, seed=3, bias_init="truncated_normal")
    assert np.all(np.abs(t["user_feat"]) <= 0.04 + 1e-7) and abs(float(t["user_feat"].std()) - 0.0176) < 0.002
    assert np.all(np.abs(t["user_bias"]) <= 2.0) and abs(float(t["mu"][0])) <= np.sqrt(3.0)
    g = init.init_tables(2000, 1000, 16, seed=3, bias_init="glorot")
    assert np.all(np.abs(g["user_bias"]) <= np.sqrt(3.0 / 2000) + 1e-7)
    same = init.init_tables(2000, 1000, 16, seed=3, bias_init="truncated_normal")
    assert all(np.array_equal(t[k], same[k]) for k in t)


def test_host_pack_feed_casts_like_a_tf_feed():
    """tfr_host_pack_feed (host-only, no GPU): the columns a reference iterator yields -- float64 VIEWS of one [B, ncols]
    matrix, ids included (dataio.py:103,116-117) -- land in the staging buffer as int32 ids + float32 rates, by value
    (TF feeding an int32 placeholder), for every dtype / stride the engine accepts."""
    from tf_recomm_b200 import dataio
    L = _lib.load()
    rng = np.random.default_rng(0)
    for B in (0, 1, 7, 20000):          # 20000 >= the threshold where the packing is split over host threads
        users = rng.integers(0, 100000, B).astype(np.int32)
        items = rng.integers(0, 5000, B).astype(np.int32)
        rates = rng.integers(1, 6, B).astype(np.float32) + 0.5
        it = dataio.OneEpochIterator([users, items, rates], batch_size=-1)
        cols = next(it) if B else [np.zeros(0), np.zeros(0), np.zeros(0)]
        if B:
            assert cols[0].dtype == np.float64 and cols[0].strides[0] == 24      # strided float64 views
        variants = [cols, [users, items, rates], [users.astype(np.int64), items.astype(np.int64), rates.astype(np.float64)],
                    [users.astype(np.float32), items.astype(np.float64), rates]]
        codes = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3}
        for u, i, r in variants:
            out = np.full(3 * B + 1, -7, np.int32)
            args = []
            for a in (u, i, r):
                a = np.asarray(a)
                args += [a.ctypes.data, codes[a.dtype], a.strides[0] if a.size else a.itemsize]
            assert L.tfr_host_pack_feed(*args, B, out.ctypes.data) == 0
            assert np.array_equal(out[:B], users) and np.array_equal(out[B:2 * B], items)
            assert np.array_equal(out[2 * B:3 * B].view(np.float32), rates)
            assert out[3 * B] == -7                                              # nothing written past 12 * B bytes
    bad = np.zeros(4)
    assert L.tfr_host_pack_feed(bad.ctypes.data, 9, 8, bad.ctypes.data, 0, 8, bad.ctypes.data, 0, 8, 4,
                                np.zeros(12, np.int32).ctypes.data) < 0           # unknown dtype code -> error, no crash


def test_host_pack_feed_checked_rejects_out_of_range_ids():
    """tfr_host_pack_feed_checked: ids outside [0, num) are TFR_ERR_INVALID whatever the source dtype -- judged on the
    source value, so an id that does not fit int32 cannot alias a valid row after the cast."""
    import ctypes as C
    L = _lib.load()
    n = 20000
    rng = np.random.default_rng(0)
    rates = rng.random(n)
    out = np.zeros(3 * n, np.int32)
    for dt, code in ((np.float64, 0), (np.float32, 1), (np.int32, 2), (np.int64, 3)):
        users = rng.integers(0, 100, n).astype(dt)
        items = rng.integers(0, 50, n).astype(dt)

        def call(u, i):
            return L.tfr_host_pack_feed_checked(u.ctypes.data, code, u.itemsize, i.ctypes.data, code, i.itemsize,
                                                rates.ctypes.data, 0, 8, n, out.ctypes.data, 100, 50)
        assert call(users, items) == 0
        assert np.array_equal(out[:n], users.astype(np.int32)) and np.array_equal(out[n:2 * n], items.astype(np.int32))
        for pos, val in ((0, 100), (n - 1, -1), (n // 2, 1e6)):
            u2 = users.copy(); u2[pos] = val
            assert call(u2, items) == -1 and b"out of range" in L.tfr_last_error()
            i2 = items.copy(); i2[pos] = 50 if val == 100 else val
            assert call(users, i2) == -1
        if dt in (np.int64, np.float64):
            u2 = users.copy(); u2[3] = 2 ** 32 + 7   # would alias row 7 after a cast to int32
            assert call(u2, items) == -1
    # unchecked entry point: same packing, no range test
    users = np.array([5, 1000000, -3], np.int64); items = np.zeros(3, np.int64); r3 = np.zeros(3)
    o3 = np.zeros(9, np.int32)
    assert L.tfr_host_pack_feed(users.ctypes.data, 3, 8, items.ctypes.data, 3, 8, r3.ctypes.data, 0, 8, 3, o3.ctypes.data) == 0
    assert list(o3[:3]) == [5, 1000000, -3]


def test_tensorboard_event_file_round_trip(tmp_path):
    """summary.FileWriter writes what tf.summary.FileWriter writes (svd_train_val.py:20-21,57,189-192): TFRecord framing
    with masked CRC32-C, Event / Summary protos; read back by our reader and, when the package is there, by
    TensorBoard's own EventFileLoader."""
    from tf_recomm_b200 import summary
    w = summary.FileWriter(str(tmp_path))
    vals = [(0, "training_error", 2.63753), (0, "test_error", 2.587753), (900, "training_error", 0.9), (2 ** 33, "test_error", float("nan"))]
    for step, tag, v in vals:
        w.add_summary(summary.make_scalar_summary(tag, v), step)
    w.close()
    got = summary.read_events(w.path)
    assert [(s, t) for s, t, _ in got] == [(s, t) for s, t, _ in vals]
    for (_, _, a), (_, _, b) in zip(got, vals):
        assert (np.isnan(a) and np.isnan(b)) or a == np.float32(b)
    assert summary._crc32c(b"123456789") == 0xE3069283          # the CRC-32C check value
    raw = bytearray(open(w.path, "rb").read())
    raw[40] ^= 1
    bad = tmp_path / "bad"
    bad.write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        summary.read_events(str(bad))
    try:
        from tensorboard.backend.event_processing.event_file_loader import EventFileLoader
    except Exception:
        return
    evs = list(EventFileLoader(w.path).Load())
    assert evs[0].file_version == "brain.Event:2"
    seen = [(e.step, v.tag) for e in evs[1:] for v in e.summary.value]
    assert seen == [(s, t) for s, t, _ in vals]


def test_host_wait_flag_spins_until_the_count_is_reached_and_times_out():
    """tfr_host_wait_flag (the host side of the feed graph's delivery flag) is plain host code: it returns once the
    32-bit counter has reached the target -- wrap-around included -- and reports a timeout instead of hanging."""
    import threading
    import time
    L = _lib.load()
    flag = (C.c_uint32 * 16)()
    flag[0] = 7
    assert L.tfr_host_wait_flag(C.addressof(flag), 7, 1000) == 0        # already there
    assert L.tfr_host_wait_flag(C.addressof(flag), 5, 1000) == 0        # past it
    t0 = time.perf_counter()
    assert L.tfr_host_wait_flag(C.addressof(flag), 8, 20_000) < 0      # never written: times out
    assert 0.015 < time.perf_counter() - t0 < 2.0
    assert b"not reached" in L.tfr_last_error()

    def bump():
        time.sleep(0.01)
        flag[0] = 8
    th = threading.Thread(target=bump)
    th.start()
    assert L.tfr_host_wait_flag(C.addressof(flag), 8, 5_000_000) == 0   # written by another thread while we spin
    th.join()
    flag[0] = 0xFFFFFFFE
    assert L.tfr_host_wait_flag(C.addressof(flag), 0xFFFFFFFD, 1000) == 0
    assert L.tfr_host_wait_flag(C.addressof(flag), 2, 5_000) < 0        # 2 is AHEAD of 0xFFFFFFFE (wrapped): not reached
    flag[0] = 3
    assert L.tfr_host_wait_flag(C.addressof(flag), 2, 1000) == 0


def test_feed_stage_is_host_only():
    """tfr_svd_feed_stage packs a handed-over batch into the set's staging buffer and touches no stream: on a fresh set
    it runs without a GPU.  Ids out of range are refused and leave the set unstaged."""
    L = _lib.load()
    B, U, I = 1000, 50, 40
    rng = np.random.default_rng(3)
    users = rng.integers(0, U, B).astype(np.float64)
    items = rng.integers(0, I, B).astype(np.int64)
    rates = rng.integers(1, 6, B).astype(np.float32)
    t = _lib.SvdTables()
    t.user_num, t.item_num, t.dim = U, I, 8
    staging = np.zeros(3 * B, dtype=np.int32)
    fs = _lib.FeedSet()
    fs.h_feed = staging.ctypes.data
    fs.ev_h2d = 1   # (only dereferenced once the set has a history: used != 0)

    def stage(u, i, r):
        return L.tfr_svd_feed_stage(C.byref(t), C.byref(fs), u.ctypes.data, 0, u.strides[0], i.ctypes.data, 3, i.strides[0],
                                    r.ctypes.data, 1, r.strides[0], B)
    assert stage(users, items, rates) == 0
    assert (fs.used, fs.staged, fs.sorted) == (1, 1, 0)
    assert np.array_equal(staging[:B], users.astype(np.int32))
    assert np.array_equal(staging[B:2 * B], items.astype(np.int32))
    assert np.array_equal(staging[2 * B:].view(np.float32), rates)
    fs.used = 0
    bad = users.copy()
    bad[17] = U
    assert stage(bad, items, rates) < 0
    assert (fs.used, fs.staged) == (0, 0)
