"""Index math of the row-sharded tables (SURVEY 8e / BASELINE configs[4]): pure functions, no device code, so the
same rules drive the CUDA engine (sharded.py), the CPU gloo tests and the oracle-side emulation.

Row r of a table lives on rank r mod G at local index r // G (mod, not ranges: Zipf-hot ids spread over ranks).
The global batch of B occurrences is cut into G contiguous slices; rank h draws slice h and the ids are
all-gathered, so every rank sees the same B (user, item, rate) triples in the same order -- the order the
single-GPU path (and TF's unsorted_segment_sum) adds duplicate rows in.
"""
import numpy as np


def owner(ids, n_ranks):
    return np.asarray(ids) % n_ranks


def local_row(ids, n_ranks):
    return np.asarray(ids) // n_ranks


def rows_on_rank(n_rows, n_ranks, rank):
    """Number of rows r in [0, n_rows) with r mod n_ranks == rank."""
    return (int(n_rows) - int(rank) + int(n_ranks) - 1) // int(n_ranks) if n_rows > rank else 0


def global_row(local, n_ranks, rank):
    return np.asarray(local) * n_ranks + rank


def shard_table(table, n_ranks, rank):
    """The rows of a [n_rows, ...] host table that rank owns, in local order."""
    return np.ascontiguousarray(np.asarray(table)[rank::n_ranks])


def unshard_table(shards):
    """Inverse of shard_table over all ranks."""
    n_ranks = len(shards)
    n = sum(len(s) for s in shards)
    out = np.empty((n,) + shards[0].shape[1:], shards[0].dtype)
    for r, s in enumerate(shards):
        out[r::n_ranks] = s
    return out


def local_keys(ids, n_ranks, rank, rows_local):
    """What tfr_shard_gather_rows writes: the local row for owned occurrences, rows_local ("not mine") otherwise."""
    ids = np.asarray(ids)
    return np.where(ids % n_ranks == rank, ids // n_ranks, rows_local).astype(np.int32)


def batch_slice(B, n_ranks, rank):
    """Contiguous slice of the global batch that rank draws: [lo, hi)."""
    per = (B + n_ranks - 1) // n_ranks
    lo = min(rank * per, B)
    return lo, min(lo + per, B)
