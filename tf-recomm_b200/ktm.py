"""Knowledge-tracing-machine feature encoding: the sparse design matrix fm.py:61-93 (`df_to_sparse`) builds and feeds
to the factorization machine, the per-skill win / fail counters it reads from skill_wins.npz / skill_fails.npz
(computed in doc/Assistments from scratch.ipynb), and a synthetic ASSISTments-shaped dataset in the reference's
on-disk layout (dataio.py:19-28: data/<name>/{all.csv, config.yml, qmatrix.npz, skill_wins.npz, skill_fails.npz}).

Host-side input preparation (scipy / numpy), not the device path.  The encoder is pinned by the known-answer table
typeset in the reference's diagram_pretty.tex:16-22,31 (tests/golden/ktm_encoder_dummy.json).
"""
import os

import numpy as np
from scipy.sparse import coo_matrix, csr_matrix, diags, hstack, load_npz, save_npz

AGENT_ORDER = ("users", "items", "skills", "attempts", "wins", "fails", "item_wins", "item_fails")


def skill_counters(users, items, outcomes, qmatrix):
    """-> (skill_wins, skill_fails) CSR [n_events, n_skills]: for event k of user u on item i, on every skill s of
    the item (qmatrix[i, s] != 0), how many EARLIER events of u on an item carrying s were wins / fails.  Events
    are taken in file order (doc/Assistments from scratch.ipynb builds the same counters with a running dict)."""
    q = csr_matrix(qmatrix)
    n, n_skills = len(users), q.shape[1]
    wins, fails = {}, {}
    rows, cols, w_val, f_val = [], [], [], []
    for k in range(n):
        u, i, won = int(users[k]), int(items[k]), bool(outcomes[k] > 0.5)
        for s in q.indices[q.indptr[i]:q.indptr[i + 1]]:
            key = (u, int(s))
            rows.append(k)
            cols.append(int(s))
            w_val.append(wins.get(key, 0))
            f_val.append(fails.get(key, 0))
            if won:
                wins[key] = wins.get(key, 0) + 1
            else:
                fails[key] = fails.get(key, 0) + 1
    shape = (n, n_skills)
    # explicit zeros are kept (same sparsity pattern as the skills block), like the notebook's lil -> csr matrices
    return (csr_matrix((np.array(w_val, np.float64), (rows, cols)), shape=shape),
            csr_matrix((np.array(f_val, np.float64), (rows, cols)), shape=shape))


def df_to_sparse(df, active_agents, user_num, item_num, qmatrix=None, skill_wins=None, skill_fails=None):
    """fm.py:61-93: one block per active agent, hstacked in `active_agents` order ->  CSR [n_events, sum widths].
    users / items: one-hot; skills: qmatrix[item]; item_wins / item_fails: the item one-hot scaled by the event's
    wins / fails column; attempts / wins / fails: the per-skill counters."""
    n = len(df)
    rows = np.arange(n)
    user = np.asarray(df["user"], dtype=np.int64)
    item = np.asarray(df["item"], dtype=np.int64)
    if qmatrix is None:  # fm.py:45-46: no q-matrix file -> every item is its own skill
        qmatrix = diags([1.0] * item_num).tocsr()
    X = {}
    X["users"] = coo_matrix((np.ones(n), (rows, user)), shape=(n, user_num))
    X["items"] = coo_matrix((np.ones(n), (rows, item)), shape=(n, item_num))
    X["skills"] = csr_matrix(qmatrix)[item]
    X["item_wins"] = coo_matrix((np.asarray(df["wins"], dtype=np.float64), (rows, item)), shape=(n, item_num))
    X["item_fails"] = coo_matrix((np.asarray(df["fails"], dtype=np.float64), (rows, item)), shape=(n, item_num))
    if skill_wins is not None:
        X["attempts"] = skill_wins + skill_fails
        X["wins"] = skill_wins
        X["fails"] = skill_fails
    missing = [a for a in active_agents if a != "extra" and a not in X]
    if missing:
        raise ValueError("agents %s need skill_wins.npz / skill_fails.npz" % missing)
    return hstack([X[a] for a in active_agents if a != "extra"]).tocsr()


def df_to_sparse_device(df, active_agents, user_num, item_num, qmatrix=None, skill_wins=None, skill_fails=None,
                        device=None):
    """df_to_sparse built as CSR directly in HBM (tfr_ktm_csr_indptr / tfr_ktm_csr_fill: count -> scan -> fill): the
    event columns, the q-matrix and the per-event skill counters are uploaded once, the design matrix itself never
    exists on the host.  -> (indptr int64 [n+1], indices int32 [nnz], data float32 [nnz], n_cols), torch CUDA tensors.
    Same matrix as df_to_sparse (explicit zeros of the counter blocks kept, zero sums of `attempts` dropped, as scipy
    does)."""
    import ctypes as C
    import torch
    from . import _lib
    from ._lib import check
    L = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.TfrError("tf-recomm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback for this path")
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    n = len(df)
    if qmatrix is None:
        qmatrix = diags([1.0] * item_num).tocsr()
    q = csr_matrix(qmatrix)
    n_skills = q.shape[1]
    width = dict(users=user_num, items=item_num, skills=n_skills, attempts=n_skills, wins=n_skills, fails=n_skills,
                 item_wins=item_num, item_fails=item_num)
    agents = [a for a in active_agents if a != "extra"]
    need_counters = [a for a in agents if a in ("attempts", "wins", "fails")]
    if need_counters and skill_wins is None:
        raise ValueError("agents %s need skill_wins.npz / skill_fails.npz" % need_counters)
    kinds = (C.c_int32 * 8)(*[AGENT_ORDER.index(a) for a in agents])
    col0, at = [], 0
    for a in agents:
        col0.append(at)
        at += width[a]
    col0_arr = (C.c_int32 * 8)(*col0)

    def up(a, dt):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)

    def csr3(m):
        if m is None:
            return None, None, None
        m = csr_matrix(m, copy=True)
        m.sort_indices()   # column order inside a row (explicit zeros stay): the blocks are emitted in this order
        return up(m.indptr, np.int64), up(m.indices, np.int32), up(m.data, np.float32)
    user, item = up(df["user"], np.int32), up(df["item"], np.int32)
    if n and (int(user.min()) < 0 or int(user.max()) >= user_num or int(item.min()) < 0 or int(item.max()) >= item_num):
        raise _lib.TfrError("indices out of range in the event log")
    wins_col, fails_col = up(df["wins"], np.float32), up(df["fails"], np.float32)
    qp, qi, qd = csr3(q)
    swp, swi, swd = csr3(skill_wins)
    sfp, sfi, sfd = csr3(skill_fails)
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    nbytes = check(L.tfr_ktm_workspace_bytes(n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        common = (user.data_ptr(), item.data_ptr(), wins_col.data_ptr(), fails_col.data_ptr(), ptr(qp), ptr(qi), ptr(qd),
                  ptr(swp), ptr(swi), ptr(swd), ptr(sfp), ptr(sfi), ptr(sfd), n, kinds, col0_arr, len(agents))
        check(L.tfr_ktm_csr_indptr(*common, indptr.data_ptr(), ws.data_ptr(), nbytes, st))
        nnz = int(indptr[-1].item())
        indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
        data = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)[:nnz]
        check(L.tfr_ktm_csr_fill(*common, indptr.data_ptr(), indices.data_ptr() if nnz else None,
                                 data.data_ptr() if nnz else None, st) if nnz else 0)
    return indptr, indices, data, at


def load_dataset(dataset, data_folder="data"):
    """What fm.py:34-51 loads: (df, config, qmatrix, skill_wins | None, skill_fails | None)."""
    from . import dataio
    _, _, config_file, q_npz, sw_npz, sf_npz = dataio.build_new_paths(dataset, data_folder)
    config = dataio.get_config(config_file)
    df = dataio.get_new_data(dataset, data_folder)
    try:
        qmatrix = load_npz(q_npz)
    except FileNotFoundError:
        qmatrix = diags([1.0] * config["ITEM_NUM"]).tocsr()
    try:
        skill_wins, skill_fails = load_npz(sw_npz), load_npz(sf_npz)
    except Exception:  # fm.py:50-52 swallows everything here
        skill_wins = skill_fails = None
    return df, config, qmatrix, skill_wins, skill_fails


# per-item skill-count distribution of ASSISTments (doc/Assistments from scratch.ipynb cell 48)
_SKILLS_PER_ITEM = ((0, 0.335), (1, 0.552), (2, 0.098), (3, 0.012), (4, 0.003))


def make_ktm_events(n_events=346860, user_num=4217, item_num=26688, n_skills=123, seed=13575, dim=5):
    """Synthetic event log of the ASSISTments shape: (users, items, outcomes, qmatrix).  Outcomes ~ Bernoulli(sigmoid)
    of a planted FM over user / item / skill features; events of a user are contiguous (a student's sequence)."""
    rng = np.random.default_rng(seed)
    counts = rng.choice([c for c, _ in _SKILLS_PER_ITEM], size=item_num, p=[p for _, p in _SKILLS_PER_ITEM])
    rows = np.repeat(np.arange(item_num), counts)
    cols = np.concatenate([rng.choice(n_skills, size=c, replace=False) for c in counts if c > 0]) if rows.size else []
    qmatrix = csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(item_num, n_skills))
    act = rng.lognormal(0.0, 1.0, user_num)
    users = np.sort(rng.choice(user_num, size=n_events, p=act / act.sum())).astype(np.int32)
    pop = 1.0 / np.arange(1, item_num + 1) ** 0.8
    items = rng.permutation(item_num)[rng.choice(item_num, size=n_events, p=pop / pop.sum())].astype(np.int32)
    bu, bi, bs = rng.normal(0, 0.8, user_num), rng.normal(0, 0.8, item_num), rng.normal(0, 0.3, n_skills)
    P, Q = rng.normal(0, 0.5, (user_num, dim)), rng.normal(0, 0.5, (item_num, dim))
    skill_term = np.asarray(qmatrix[items] @ bs).ravel()
    logit = 0.5 + bu[users] + bi[items] + skill_term + np.einsum("nk,nk->n", P[users], Q[items])
    outcomes = (rng.random(n_events) < 1.0 / (1.0 + np.exp(-logit))).astype(np.float32)
    return users, items, outcomes, qmatrix


def write_dataset(dataset, users, items, outcomes, qmatrix, user_num, item_num, data_folder="data", batch_size=10000,
                  with_counters=True):
    """Writes the reference's on-disk layout for fm.py (dataio.py:19-28,38-46: header-less all.csv with columns
    user,item,outcome,wins,fails; config.yml with USER_NUM / ITEM_NUM / NB_CLASSES / BATCH_SIZE)."""
    import yaml
    from . import dataio
    folder, all_csv, config_file, q_npz, sw_npz, sf_npz = dataio.build_new_paths(dataset, data_folder)
    os.makedirs(folder, exist_ok=True)
    # the wins / fails columns of all.csv: the user's earlier wins / fails on this very item
    w, f = {}, {}
    wins, fails = np.zeros(len(users), np.int64), np.zeros(len(users), np.int64)
    for k in range(len(users)):
        key = (int(users[k]), int(items[k]))
        wins[k], fails[k] = w.get(key, 0), f.get(key, 0)
        if outcomes[k] > 0.5:
            w[key] = wins[k] + 1
        else:
            f[key] = fails[k] + 1
    with open(all_csv, "w") as fh:
        for k in range(len(users)):
            fh.write("%d,%d,%d,%d,%d\n" % (users[k], items[k], int(outcomes[k]), wins[k], fails[k]))
    with open(config_file, "w") as fh:
        yaml.safe_dump(dict(USER_NUM=int(user_num), ITEM_NUM=int(item_num), NB_CLASSES=2, BATCH_SIZE=int(batch_size)), fh)
    save_npz(q_npz, csr_matrix(qmatrix))
    if with_counters:
        sw, sf = skill_counters(users, items, outcomes, qmatrix)
        save_npz(sw_npz, sw)
        save_npz(sf_npz, sf)
    return folder
