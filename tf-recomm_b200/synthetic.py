"""Synthetic rating data of the BASELINE.json shapes (there is no network for MovieLens): users with log-normal
activity, items with Zipf popularity, ratings from a planted low-rank biased model + noise, clipped to 1..5
(SURVEY section 8d).  Deterministic in `seed`.  Host-side numpy: this is input preparation, not the path."""
import numpy as np

SHAPES = {  # name: (users, items, ratings)
    "ml1m": (6040, 3952, 1000209),
    "ml25m": (162541, 62423, 25000095),
}


def make_ratings(user_num, item_num, n, seed=13575, rank=8, noise=0.9, zipf_a=1.0, binary=False):
    rng = np.random.default_rng(seed)
    act = rng.lognormal(0.0, 1.0, user_num)
    pop = 1.0 / np.arange(1, item_num + 1) ** zipf_a
    users = rng.choice(user_num, size=n, p=act / act.sum()).astype(np.int32)
    items = rng.permutation(item_num)[rng.choice(item_num, size=n, p=pop / pop.sum())].astype(np.int32)
    bu = rng.normal(0, 0.4, user_num)
    bi = rng.normal(0, 0.5, item_num)
    P = rng.normal(0, 1.0, (user_num, rank)) / np.sqrt(rank)
    Q = rng.normal(0, 1.0, (item_num, rank)) / np.sqrt(rank)
    score = 3.58 + bu[users] + bi[items] + np.einsum("nk,nk->n", P[users], Q[items]) * 0.7
    score = score + rng.normal(0, noise, n)
    if binary:
        rates = (score > 3.58).astype(np.float32)
    else:
        rates = np.clip(np.rint(score), 1, 5).astype(np.float32)
    return users, items, rates


def split(users, items, rates, val_frac=0.1, seed=13575):
    rng = np.random.default_rng(seed + 1)
    perm = rng.permutation(len(users))
    n_val = int(round(len(users) * val_frac))
    va, tr = perm[:n_val], perm[n_val:]
    return (users[tr], items[tr], rates[tr]), (users[va], items[va], rates[va])
