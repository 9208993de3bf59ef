"""Row-sharded path on CPU: world_size-2 and -4 torch.distributed runs over gloo.  Covers the host-side
logic the CUDA engine shares (tf-recomm_b200/sharding.py: ownership, local rows, batch slices) and the exchange
protocol itself (all-gather of the batch slices, all-reduce of the owner-filled row buffers), with the oracle's numpy
mirror standing in for the device kernels: after every step the re-assembled shards must equal the single-table
oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tf_recomm_b200 import init, sharding


def test_sharding_index_math():
    for G in (1, 2, 3, 8):
        for n in (0, 1, 7, 8, 9, 100):
            sizes = [sharding.rows_on_rank(n, G, r) for r in range(G)]
            assert sum(sizes) == n
            ids = np.arange(n)
            for r in range(G):
                mine = ids[sharding.owner(ids, G) == r]
                assert len(mine) == sizes[r]
                assert np.array_equal(sharding.global_row(sharding.local_row(mine, G), G, r), mine)
                keys = sharding.local_keys(ids, G, r, sizes[r])
                assert np.all((keys == sizes[r]) == (ids % G != r))
    t = np.arange(23 * 3).reshape(23, 3)
    assert np.array_equal(sharding.unshard_table([sharding.shard_table(t, 4, r) for r in range(4)]), t)
    cuts = [sharding.batch_slice(10, 4, r) for r in range(4)]
    assert cuts == [(0, 3), (3, 6), (6, 9), (9, 10)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, U, I, d, B, steps, out):
    import oracle
    from oracle import np_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lr, reg = 1e-2, 0.05
    tabs = init.init_tables(U, I, d, seed=5, bias_init="truncated_normal")
    U_loc, I_loc = sharding.rows_on_rank(U, world, rank), sharding.rows_on_rank(I, world, rank)
    loc = {k: (tabs[k].copy() if k == "mu" else sharding.shard_table(tabs[k], world, rank)) for k in tabs}
    slots = {k: [np.zeros_like(v), np.zeros_like(v)] for k, v in loc.items()}
    b1p, b2p = np.float32(0.9), np.float32(0.999)
    rng = np.random.default_rng(100 + rank)        # every rank draws ITS slice of the batch
    lo, hi = sharding.batch_slice(B, world, rank)
    for step in range(steps):
        mine = np.stack([rng.integers(0, U, hi - lo), rng.integers(0, max(I // 4, 1), hi - lo),
                         rng.integers(1, 6, hi - lo)]).astype(np.float32)
        gathered = [torch.zeros(3, hi - lo) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine))                          # step 1
        glob = torch.cat(gathered, 1).numpy()
        users, items, rates = glob[0].astype(np.int32), glob[1].astype(np.int32), glob[2]
        ku = sharding.local_keys(users, world, rank, U_loc)
        ki = sharding.local_keys(items, world, rank, I_loc)
        g_uf = np.zeros((B, d), np.float32); g_if = np.zeros((B, d), np.float32)
        g_ub = np.zeros(B, np.float32); g_ib = np.zeros(B, np.float32)
        mu_, mi_ = ku < U_loc, ki < I_loc
        g_uf[mu_], g_ub[mu_] = loc["user_feat"][ku[mu_]], loc["user_bias"][ku[mu_]]   # step 2
        g_if[mi_], g_ib[mi_] = loc["item_feat"][ki[mi_]], loc["item_bias"][ki[mi_]]
        flat = torch.from_numpy(np.concatenate([g_uf.ravel(), g_if.ravel(), g_ub, g_ib]))
        dist.all_reduce(flat)                                                          # step 3
        flat = flat.numpy()
        g_uf, g_if = flat[:B * d].reshape(B, d), flat[B * d:2 * B * d].reshape(B, d)
        g_ub, g_ib = flat[2 * B * d:2 * B * d + B], flat[2 * B * d + B:]
        # step 4 on local tables: forward from the gathered rows, per-occurrence grads, dedup over OWNED keys
        x = (np.sum(g_uf * g_if, axis=1, dtype=np.float32) + loc["mu"][0] + g_ub + g_ib).astype(np.float32)
        e = (x - rates).astype(np.float32)
        lr_t = np_oracle.adam_lr_t(lr, b1p, b2p)
        for keys, ok, own, partner, feat, bias in ((ku, mu_, g_uf, g_if, "user_feat", "user_bias"),
                                                   (ki, mi_, g_if, g_uf, "item_feat", "item_bias")):
            gf = (e[:, None] * partner + np.float32(reg) * own)[ok].astype(np.float32)
            uq, idx = np_oracle.unique_first_occurrence(keys[ok])
            for name, g in ((feat, gf), (bias, e[ok])):
                gs = np_oracle.segment_sum(g, idx, len(uq))
                loc[name], slots[name][0], slots[name][1] = np_oracle.adam_sparse(loc[name], slots[name][0], slots[name][1],
                                                                                 uq, gs, lr_t)
        gmu = np.sum(e, dtype=np.float32)
        m_, v_ = slots["mu"]
        m_ = m_ + (gmu - m_) * (np.float32(1) - np.float32(0.9))
        v_ = v_ + (gmu * gmu - v_) * (np.float32(1) - np.float32(0.999))
        loc["mu"] = (loc["mu"] - (m_ * lr_t) / (np.sqrt(v_) + np.float32(1e-8))).astype(np.float32)
        slots["mu"] = [m_.astype(np.float32), v_.astype(np.float32)]
        b1p, b2p = b1p * np.float32(0.9), b2p * np.float32(0.999)
        if rank == 0:
            out.setdefault("batches", []).append((users, items, rates))
    # re-assemble on rank 0
    for name in ("user_feat", "item_feat", "user_bias", "item_bias"):
        parts = [None] * world
        dist.all_gather_object(parts, loc[name])
        if rank == 0:
            out[name] = sharding.unshard_table(parts)
    if rank == 0:
        out["mu"] = loc["mu"]
        out["batches"] = out["batches"]
    dist.barrier()
    dist.destroy_process_group()


def _run(rank, world, port, shape, q):
    out = {}
    _rank_main(rank, world, port, *shape, out)
    if rank == 0:
        q.put(out)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_protocol_equals_single_table_oracle(world):
    import oracle
    U, I, d, B, steps = 41, 23, 6, 64, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, (U, I, d, B, steps), q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tabs = init.init_tables(U, I, d, seed=5, bias_init="truncated_normal")
    orc = oracle.SvdOracle(tabs["mu"], tabs["user_bias"], tabs["item_bias"], tabs["user_feat"], tabs["item_feat"], 1e-2, 0.05)
    for users, items, rates in out["batches"]:
        assert len(users) == B
        orc.train_step(users, items, rates)
    for name in ("user_feat", "item_feat", "user_bias", "item_bias", "mu"):
        np.testing.assert_allclose(out[name].reshape(getattr(orc, name).shape), getattr(orc, name), rtol=2e-5, atol=2e-7,
                                   err_msg=name)


# ---- the all-to-all exchange (north_star: ids -> rows back -> gradient records to the owners) over gloo ---------------
def _a2a_rank_main(rank, world, port, U, I, d, B, steps, out):
    from oracle import np_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, lr, reg = world, 1e-2, 0.05
    tabs = init.init_tables(U, I, d, seed=5, bias_init="truncated_normal")
    U_loc, I_loc = sharding.rows_on_rank(U, G, rank), sharding.rows_on_rank(I, G, rank)
    loc = {k: (tabs[k].copy() if k == "mu" else sharding.shard_table(tabs[k], G, rank)) for k in tabs}
    slots = {k: [np.zeros_like(v), np.zeros_like(v)] for k, v in loc.items()}
    b1p, b2p = np.float32(0.9), np.float32(0.999)
    rng = np.random.default_rng(100 + rank)
    lo, hi = sharding.batch_slice(B, G, rank)
    n = hi - lo

    def a2a(send, send_splits, recv_splits):
        recv = torch.empty((int(sum(recv_splits)),) + tuple(send.shape[1:]), dtype=send.dtype)
        dist.all_to_all_single(recv, send, [int(x) for x in recv_splits], [int(x) for x in send_splits])
        return recv
    for step in range(steps):
        users = rng.integers(0, U, n).astype(np.int32)
        items = rng.integers(0, max(I // 4, 1), n).astype(np.int32)
        rates = rng.integers(1, 6, n).astype(np.float32)
        # 1. bucket by owner, stably (what tfr_shard_bucket does): combined layout [dst: users | items]
        send_ids, slot, counts = [], {}, np.zeros(2 * G, np.int64)
        for g in range(G):
            for side, ids in ((0, users), (1, items)):
                pos = np.flatnonzero(ids % G == g)
                for p in pos:
                    slot[(side, p)] = len(send_ids)
                    send_ids.append(ids[p] // G)
                counts[side * G + g] = len(pos)
        allc = [torch.zeros(2 * G, dtype=torch.int64) for _ in range(G)]
        dist.all_gather(allc, torch.from_numpy(counts))
        cnt = torch.stack(allc).numpy()
        send_splits = cnt[rank, :G] + cnt[rank, G:]
        recv_cu, recv_ci = cnt[:, rank], cnt[:, G + rank]
        recv_splits = recv_cu + recv_ci
        # 2. ids to the owners
        recv_ids = a2a(torch.tensor(send_ids, dtype=torch.int64).reshape(-1), send_splits, recv_splits).numpy()
        is_user = np.concatenate([np.r_[np.ones(recv_cu[s], bool), np.zeros(recv_ci[s], bool)] for s in range(G)]) \
            if len(recv_ids) else np.zeros(0, bool)
        # 3. owners pack the rows; 4. back to the requesters
        rec = np.zeros((len(recv_ids), d + 1), np.float32)
        rec[is_user, :d], rec[is_user, d] = loc["user_feat"][recv_ids[is_user]], loc["user_bias"][recv_ids[is_user]]
        rec[~is_user, :d], rec[~is_user, d] = loc["item_feat"][recv_ids[~is_user]], loc["item_bias"][recv_ids[~is_user]]
        rec_in = a2a(torch.from_numpy(rec), recv_splits, send_splits).numpy()
        # 5. forward on MY slice; one record per occurrence and table: [partner row | e]
        su = np.array([slot[(0, p)] for p in range(n)], np.int64)
        si = np.array([slot[(1, p)] for p in range(n)], np.int64)
        ur, vr = rec_in[su], rec_in[si]
        x = (np.sum(ur[:, :d] * vr[:, :d], axis=1, dtype=np.float32) + loc["mu"][0] + ur[:, d] + vr[:, d]).astype(np.float32)
        e = (x - rates).astype(np.float32)
        rec_out = np.zeros((2 * n, d + 1), np.float32)
        rec_out[su, :d], rec_out[su, d] = vr[:, :d], e
        rec_out[si, :d], rec_out[si, d] = ur[:, :d], e
        sums = torch.tensor([float(np.sum(e, dtype=np.float32))], dtype=torch.float64)
        dist.all_reduce(sums)
        # 6. records to the owners; 7. ordered sums in ARRIVAL order (= global batch order) + local Adam
        grads_in = a2a(torch.from_numpy(rec_out), send_splits, recv_splits).numpy()
        lr_t = np_oracle.adam_lr_t(lr, b1p, b2p)
        for mask, feat, bias in ((is_user, "user_feat", "user_bias"), (~is_user, "item_feat", "item_bias")):
            keys = recv_ids[mask]
            part, ee = grads_in[mask, :d], grads_in[mask, d]
            gf = (ee[:, None] * part + np.float32(reg) * loc[feat][keys]).astype(np.float32)
            uq, idx = np_oracle.unique_first_occurrence(keys.astype(np.int32))
            for name, g in ((feat, gf), (bias, ee)):
                gs = np_oracle.segment_sum(g, idx, len(uq))
                loc[name], slots[name][0], slots[name][1] = np_oracle.adam_sparse(loc[name], slots[name][0], slots[name][1],
                                                                                 uq, gs, lr_t)
        gmu = np.float32(sums.item())
        m_, v_ = slots["mu"]
        m_ = m_ + (gmu - m_) * (np.float32(1) - np.float32(0.9))
        v_ = v_ + (gmu * gmu - v_) * (np.float32(1) - np.float32(0.999))
        loc["mu"] = (loc["mu"] - (m_ * lr_t) / (np.sqrt(v_) + np.float32(1e-8))).astype(np.float32)
        slots["mu"] = [m_.astype(np.float32), v_.astype(np.float32)]
        b1p, b2p = b1p * np.float32(0.9), b2p * np.float32(0.999)
        parts = [None] * G
        dist.all_gather_object(parts, (users, items, rates))
        if rank == 0:
            out.setdefault("batches", []).append(tuple(np.concatenate([p[k] for p in parts]) for k in range(3)))
    for name in ("user_feat", "item_feat", "user_bias", "item_bias"):
        parts = [None] * G
        dist.all_gather_object(parts, loc[name])
        if rank == 0:
            out[name] = sharding.unshard_table(parts)
    if rank == 0:
        out["mu"] = loc["mu"]
    dist.barrier()
    dist.destroy_process_group()


def _a2a_run(rank, world, port, shape, q):
    out = {}
    _a2a_rank_main(rank, world, port, *shape, out)
    if rank == 0:
        q.put(out)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_a2a_protocol_equals_single_table_oracle(world):
    """Real multi-process all_to_all_single (gloo) with the exact per-peer counts: ids to the owners, rows back,
    [partner row | e] records to the owners.  Summing the records in ARRIVAL order reproduces the single-table oracle:
    arrival order = (source rank, position in its slice) = global batch order."""
    import oracle
    U, I, d, B, steps = 41, 23, 6, 64, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_a2a_run, args=(r, world, port, (U, I, d, B, steps), q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tabs = init.init_tables(U, I, d, seed=5, bias_init="truncated_normal")
    orc = oracle.SvdOracle(tabs["mu"], tabs["user_bias"], tabs["item_bias"], tabs["user_feat"], tabs["item_feat"], 1e-2, 0.05)
    for users, items, rates in out["batches"]:
        assert len(users) == B
        orc.train_step(users, items, rates)
    for name in ("user_feat", "item_feat", "user_bias", "item_bias", "mu"):
        np.testing.assert_allclose(out[name].reshape(getattr(orc, name).shape), getattr(orc, name), rtol=2e-5, atol=2e-7,
                                   err_msg=name)
