"""ShardedSvdEngine: the train step with embedding tables + Adam state row-sharded over the ranks of one
NVLink/NVSwitch box (BASELINE configs[4]: 100M users x 10M items, dim 128, 170 GB of state).

Per step, on every rank (one process per GPU, torch.distributed/NCCL for the plumbing):
  1. all-gather of the batch slices' (user, item, rate) -> every rank holds the same global batch   [NCCL, 12 B/rating]
  2. tfr_shard_gather_rows: copy the rows this rank OWNS to their batch positions, zeros elsewhere    [local kernel]
  3. all-reduce(sum) of the [B, dim+1] x 2 gathered buffers: exact (one non-zero term per element) -- every rank
     now has all B user rows and item rows of the batch                                              [NCCL over NVLink]
  4. the single-GPU kernels, unchanged, on LOCAL tables: forward from the gathered rows (identical on all ranks, so
     the error vector and bias_global need no further exchange), sort of the local row keys, ordered segment sums
     (partner rows from the gathered buffers), one Adam pass over the local shard                     [local kernels]
The exchange is ~68 MB per step against a 42.6 GB local table pass at G=8, so scaling is governed by the local pass
(SURVEY 8e).  The reference has no distributed code; this is the B200-native equivalent north_star asks for.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, sharding
from ._lib import README_FLAGS, VAR_ALL, OptScalars, SvdTables, check
from .engine import SvdEngine


class ShardedSvdEngine:
    def __init__(self, user_num, item_num, dim, lr, reg, rank, world, flags=README_FLAGS, tables=None, device=None,
                 group=None, device_init_seed=13575):
        """tables: full (unsharded) numpy tables to inject (tests); None -> each rank draws its shard on device."""
        self.rank, self.world, self.group = int(rank), int(world), group
        self.U, self.I, self.d = int(user_num), int(item_num), int(dim)
        self.U_loc = sharding.rows_on_rank(self.U, world, rank)
        self.I_loc = sharding.rows_on_rank(self.I, world, rank)
        local = None
        if tables is not None:
            local = dict(mu=tables["mu"], user_bias=sharding.shard_table(tables["user_bias"], world, rank),
                         item_bias=sharding.shard_table(tables["item_bias"], world, rank),
                         user_feat=sharding.shard_table(tables["user_feat"], world, rank),
                         item_feat=sharding.shard_table(tables["item_feat"], world, rank))
        # the local shard is an ordinary engine: same tables struct, same kernels
        self.local = SvdEngine(self.U_loc, self.I_loc, dim, lr, reg, flags=flags, var_mask=VAR_ALL, tables=local,
                               device=device, device_init_seed=device_init_seed + 1000 * rank)
        self.device = self.local.device
        self.L = self.local.L
        self._bufs = {}

    def _buffers(self, B):
        b = self._bufs.get(B)
        if b is None:
            dev, d = self.device, self.d
            # one contiguous exchange buffer: [user rows | item rows | user bias | item bias]
            flat = torch.empty(2 * B * d + 2 * B, dtype=torch.float32, device=dev)
            b = dict(flat=flat, g_uf=flat[:B * d], g_if=flat[B * d:2 * B * d], g_ub=flat[2 * B * d:2 * B * d + B],
                     g_ib=flat[2 * B * d + B:], key_u=torch.empty(B, dtype=torch.int32, device=dev),
                     key_i=torch.empty(B, dtype=torch.int32, device=dev),
                     logits=torch.empty(B, dtype=torch.float32, device=dev),
                     infer=torch.empty(B, dtype=torch.float32, device=dev))
            t = SvdTables()
            C.memmove(C.byref(t), C.byref(self.local.tables_struct), C.sizeof(SvdTables))
            t.g_user_feat, t.g_item_feat = b["g_uf"].data_ptr(), b["g_if"].data_ptr()
            t.g_user_bias, t.g_item_bias = b["g_ub"].data_ptr(), b["g_ib"].data_ptr()
            b["tables"] = t
            self._bufs[B] = b
        return b

    def gather_owned(self, users, items, bufs):
        """Step 2: this rank's half of the row exchange."""
        B, e, st = users.numel(), self.local, self.local._stream()
        check(self.L.tfr_shard_gather_rows(e.t["user_feat"].data_ptr(), e.t["user_bias"].data_ptr(), self.U_loc, self.d,
                                           e.feat_stride,
                                           users.data_ptr(), B, self.world, self.rank, bufs["g_uf"].data_ptr(),
                                           bufs["g_ub"].data_ptr(), bufs["key_u"].data_ptr(), st))
        check(self.L.tfr_shard_gather_rows(e.t["item_feat"].data_ptr(), e.t["item_bias"].data_ptr(), self.I_loc, self.d,
                                           e.feat_stride,
                                           items.data_ptr(), B, self.world, self.rank, bufs["g_if"].data_ptr(),
                                           bufs["g_ib"].data_ptr(), bufs["key_i"].data_ptr(), st))

    def local_step(self, bufs, rates):
        """Step 4: everything after the exchange, on local tables."""
        e = self.local
        B = rates.numel()
        ws = e.workspace(B)
        st = e._stream()
        check(self.L.tfr_svd_begin_step(e.opt.data_ptr(), st))
        check(self.L.tfr_svd_train_step(C.byref(bufs["tables"]), e.opt.data_ptr(), bufs["key_u"].data_ptr(),
                                        bufs["key_i"].data_ptr(), rates.data_ptr(), B, bufs["logits"].data_ptr(),
                                        bufs["infer"].data_ptr(), e.flags, e.var_mask, ws.data_ptr(), ws.numel(), st,
                                        e._side_arr, e._n_side(), e._fj_events))
        return bufs["logits"], bufs["infer"]

    def train_step(self, users, items, rates):
        """users/items/rates: the GLOBAL batch (identical on every rank), device int32/int32/float32 tensors."""
        e = self.local
        users, items, rates = e._dev_i32(users), e._dev_i32(items), e._dev_f32(rates)
        bufs = self._buffers(users.numel())
        with torch.cuda.device(self.device):
            self.gather_owned(users, items, bufs)
            if self.world > 1:
                dist.all_reduce(bufs["flat"], op=dist.ReduceOp.SUM, group=self.group)
            return self.local_step(bufs, rates)

    def train_step_from_slices(self, users_slice, items_slice, rates_slice):
        """Each rank passes ITS slice of the global batch (what its data loader drew); step 1 all-gathers them."""
        e = self.local
        parts = torch.stack([e._dev_i32(users_slice).view(torch.float32), e._dev_i32(items_slice).view(torch.float32),
                             e._dev_f32(rates_slice)])
        if self.world > 1:
            out = torch.empty((self.world,) + tuple(parts.shape), dtype=parts.dtype, device=self.device)
            dist.all_gather_into_tensor(out, parts, group=self.group)
            users = out[:, 0].reshape(-1).view(torch.int32)
            items = out[:, 1].reshape(-1).view(torch.int32)
            rates = out[:, 2].reshape(-1)
        else:
            users, items, rates = parts[0].view(torch.int32), parts[1].view(torch.int32), parts[2]
        return self.train_step(users.contiguous(), items.contiguous(), rates.contiguous())

    def exchange_bytes(self, B):
        """Bytes this rank contributes to the per-step collectives (ids all-gather + rows all-reduce)."""
        return 12 * (B // self.world) + 4 * (2 * B * self.d + 2 * B)

    def get_local_tables(self):
        return self.local.get_tables()
