"""Generates tests/golden/*.npz|json from the REAL reference code. Run in the build container only:

    python tests/golden/make_golden.py        (needs /root/reference; the GPU box never runs this)

What can be executed of the reference here: dataio.py imports and runs; ops.py /
svd_train_val.py need TensorFlow (absent), fm.py needs pywFM/libFM (absent).  So the fixtures pin
  * dataio.ShuffleIterator / OneEpochIterator batch composition (dataio.py:94-138) under
    np.random.seed(13575) (svd_train_val.py:15),
  * dataio.read_process dtypes and dataio.get_legend strings (dataio.py:38-46,63-86),
  * the FM forward formula, by exec-ing the `fma` definition lifted from forward.py:21-22 at
    generation time (the script around it has hard-coded paths and cannot be imported),
  * the KTM encoder known-answer table typeset in diagram_pretty.tex:16-22,31 for the dummy
    dataset of doc/"Assistments from scratch.ipynb" cell 64.
"""
import io
import json
import os
import re
import sys
import contextlib

import numpy as np
import scipy.sparse as sp

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import dataio as ref_dataio  # noqa: E402  (the reference's own module)


def shuffle_fixture():
    rng = np.random.RandomState(7)
    N = 1000
    cols5 = [rng.randint(0, 50, N).astype(np.int32), rng.randint(0, 30, N).astype(np.int32),
             rng.randint(1, 6, N).astype(np.float32), rng.randint(0, 4, N).astype(np.float32),
             rng.randint(0, 4, N).astype(np.float32)]
    out = {"N": N, "cols": np.stack([c.astype(np.float64) for c in cols5])}
    for ncols, tag in ((5, "c5"), (3, "c3")):
        np.random.seed(13575)
        it = ref_dataio.ShuffleIterator(cols5[:ncols], batch_size=64)
        batches = [next(it) for _ in range(4)]
        out["shuffle_%s" % tag] = np.stack([np.stack(b) for b in batches])  # [4, ncols, 64] float64
        out["shuffle_%s_len" % tag] = len(it)
    # the ML-1M-sized index stream (first five draws 215620 358102 13815 741476 617586)
    np.random.seed(13575)
    out["randint_900188_first16"] = np.random.randint(0, 900188, (1000,))[:16]
    # sequential iterator: chunking rule np.array_split(arange(N), ceil(N/bs)) and the bs<=0 case
    for bs in (3, 128, 1000, 2048, -1):
        for N2 in (10, 1000):
            it = ref_dataio.OneEpochIterator([c[:N2] for c in cols5[:3]], batch_size=bs)
            sizes = [len(b[0]) for b in it]
            again = [len(b[0]) for b in it]  # StopIteration resets group_id (dataio.py:133-135)
            assert sizes == again
            out["oneepoch_sizes_N%d_bs%d" % (N2, bs)] = np.array(sizes)
    it = ref_dataio.OneEpochIterator(cols5[:3], batch_size=300)
    out["oneepoch_first_batch_bs300"] = np.stack(next(it))
    np.savez_compressed(os.path.join(OUT, "dataio_iterators.npz"), **{k: np.asarray(v) for k, v in out.items()})


def legend_fixture():
    cases = [dict(d=0, users=True, items=True), dict(d=5, users=True, items=True),
             dict(d=0, skills=True, attempts=True), dict(d=0, skills=True, wins=True, fails=True),
             dict(d=20, users=True, items=True, skills=True, wins=True, fails=True),
             dict(d=3, items=True, item_wins=True, item_fails=True, extra=True)]
    res = []
    for c in cases:
        with contextlib.redirect_stdout(io.StringIO()):
            short, full, latex, active = ref_dataio.get_legend(c)
        res.append(dict(args=c, short=short, full=full, latex=latex, active=active))
    csv = "3,1,1,0,0\n3,2,0,1,0\n0,2,1,0,0\n"
    path = os.path.join(OUT, "_tmp.csv")
    with open(path, "w") as f:
        f.write(csv)
    df = ref_dataio.read_process(path, sep=",")
    os.remove(path)
    rp = dict(csv=csv, columns=list(df.columns), dtypes=[str(t) for t in df.dtypes],
              values=df.to_numpy(dtype=np.float64).tolist())
    with open(os.path.join(OUT, "dataio_legend_readprocess.json"), "w") as f:
        json.dump(dict(legend=res, read_process=rp), f, indent=1)


def fm_forward_fixture():
    src = open(os.path.join(REF, "forward.py")).read().splitlines()
    fma_src = "\n".join(src[20:22])  # forward.py:21-22
    assert fma_src.startswith("def fma(x):"), fma_src
    rng = np.random.RandomState(11)
    F, d, n = 40, 6, 25
    W = rng.randn(F) * 0.3
    V = rng.randn(F, d) * 0.2
    mu = 0.37
    # binary multi-hot rows like fm.py:61-93 builds (users | items | skills)
    rows, cols = [], []
    for r in range(n):
        picks = {rng.randint(0, 10), 10 + rng.randint(0, 20)} | set(30 + rng.choice(10, rng.randint(0, 4), replace=False))
        for c in sorted(picks):
            rows.append(r); cols.append(c)
    X = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, F))
    # pywFM hands back pairwise_interactions as np.matrix (fm_mangaki.py:39-42 pickles it as is), which is
    # what makes forward.py:22's `.A1` valid; mirror that type here.
    Vm = np.matrix(V)
    ns = dict(np=np, mu=mu, W=W, V=Vm, V2=np.power(Vm, 2))
    exec(fma_src, ns)
    y = ns["fma"](X)
    np.savez_compressed(os.path.join(OUT, "fm_forward.npz"), indptr=X.indptr.astype(np.int64),
                        indices=X.indices.astype(np.int32), data=X.data.astype(np.float32), W=W, V=V,
                        mu=np.array(mu), y=np.asarray(y, np.float64))


def encoder_fixture():
    tex = open(os.path.join(REF, "diagram_pretty.tex")).read().splitlines()
    body = tex[15:22]  # diagram_pretty.tex:16-22, seven data rows
    mat = []
    for line in body:
        nums = [int(t) for t in re.findall(r"-?\d+", line.replace("\\\\", ""))]
        assert len(nums) == 14, (line, nums)
        mat.append(nums)
    outcome = [int(t) for t in re.findall(r"\d", tex[30])]  # diagram_pretty.tex:31
    assert len(outcome) == 7
    fixture = dict(
        source="diagram_pretty.tex:16-22 (X), :31 (outcome); dataset: doc/Assistments from scratch.ipynb cell 64",
        rows_user_item_outcome=[[1, 1, 1], [1, 1, 0], [1, 1, 1], [1, 2, 0], [1, 2, 1], [0, 1, 1], [0, 0, 0]],
        qmatrix=[[0, 0, 0], [1, 1, 0], [0, 1, 1]],
        blocks=["users", "items", "skills", "wins", "fails"], block_widths=[2, 3, 3, 3, 3],
        X=mat, outcome=outcome)
    with open(os.path.join(OUT, "ktm_encoder_dummy.json"), "w") as f:
        json.dump(fixture, f, indent=1)


if __name__ == "__main__":
    shuffle_fixture()
    legend_fixture()
    fm_forward_fixture()
    encoder_fixture()
    print("golden fixtures written to", OUT)
