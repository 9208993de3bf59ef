"""bench.py --gpus N (N >= 2): BASELINE configs[4] -- SVD dim 128 on 100M users x 10M items, global batch 65536,
tables + Adam state row-sharded (id mod N) over the N GPUs of one box, one process per GPU (torchrun).

Strong scaling: the workload is the same at every N (340.6 GB of table traffic per step in total, 170 GB of state),
so the baseline is the 2-GPU run -- the state does not fit one 180 GB GPU.  Per step and rank:
all-gather of the batch slices (NCCL), owner-side row gather (kernel), all-reduce of the gathered rows (NCCL over
NVLink), then the single-GPU kernels on the local shard.  Timed with CUDA events between barriers, max over ranks.
"""
import ctypes as C
import json
import os
import time

import numpy as np

SHARDED = dict(U=100_000_000, I=10_000_000, N=2_000_000_000, d=128, B=65536, config="configs[4]")
LR, REG = 1e-3, 0.05


def _draw_slice(torch, gen, n, U, I, device):
    """This rank's slice of the global batch, drawn on the device (no 2B-row host array): users
    uniform, items skewed (cube of a uniform: a few hot items collect most ratings)."""
    users = torch.randint(0, U, (n,), generator=gen, device=device, dtype=torch.int64).to(torch.int32)
    u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
    items = torch.clamp((u * u * u * I).to(torch.int64), max=I - 1).to(torch.int32)
    rates = torch.randint(1, 6, (n,), generator=gen, device=device).to(torch.float32)
    return users, items, rates


REF_SCALE = 64


def sharded_config(w, world, scale_down=1):
    return {"workload": "sharded_100Mx10M_d128_b65536" + ("_tables_scaled_down_%dx" % scale_down if scale_down > 1 else ""),
            "baseline_config": w["config"], "users": w["U"], "items": w["I"], "dim": w["d"], "batch": w["B"],
            "ratings": w["N"], "sharding": "rows id mod %d" % world,
            "model": "README: squared error + L2 + Adam (TF IndexedSlices semantics)"}


def reference_line(args, world, bench):
    """The CPU restatement on what FITS the host: the same step (batch 65536, dim 128) on tables scaled down REF_SCALE x
    (1.56 M x 156 k; the full 170 GB of state does not fit host memory).  The measured number is reported AS MEASURED,
    under a workload name that says so -- no extrapolation to the full tables.  Since the TF path's step cost is the
    table-wide Adam passes, the full-size CPU step would be ~REF_SCALE x slower: any ratio formed with this line
    UNDERSTATES the GPU path by about that factor."""
    w = dict(U=SHARDED["U"] // REF_SCALE, I=SHARDED["I"] // REF_SCALE, d=SHARDED["d"], B=SHARDED["B"], N=SHARDED["N"],
             config=SHARDED["config"])
    steps = min(args.steps or 3, 3)
    warm = min(max(args.warmup, 1), 3)
    tot, n = bench.cpu_reference_run(w, steps, warm)
    val = SHARDED["B"] / (tot / n)
    cores = os.cpu_count()
    return {
        "impl": "reference", "metric": "train ratings/sec", "value": val, "unit": "ratings/s", "n_gpus": world,
        "steps": n, "warmup": warm, "ms_per_step": tot / n * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": sharded_config(w, world, REF_SCALE),
        "cpu_baseline": {"value": val, "unit": "ratings/s", "cores": cores, "kind": "port",
                         "sample": "%d train steps (batch 65536) of oracle/tfr_oracle.c (OpenMP, %d threads) on tables "
                                   "scaled down %dx (%d x %d); reported as measured, NOT extrapolated: the full-size step "
                                   "would be about %dx slower" % (n, cores, REF_SCALE, w["U"], w["I"], REF_SCALE)},
        "e2e": {"value": val, "unit": "ratings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def in_run_parity_check(eng, torch, dist, w, world, rank, dev, steps=3):
    """Small-scale runs only (TFR_SHARDED_SCALE): the REAL NCCL step against the oracle.  Every rank's initial shard and
    slices are gathered on rank 0, the oracle steps the same global batches, the re-assembled shards are compared with
    the bar of tests/test_gpu_parity.py.  Returns a dict for the JSON line (rank 0) / None."""
    import oracle
    from tf_recomm_b200 import sharding
    B, G = w["B"], world
    lo, hi = eng_slice(B, world, rank)
    names = ("user_feat", "item_feat", "user_bias", "item_bias")

    def assembled():
        loc = eng.get_local_tables()
        parts = [None] * G
        dist.all_gather_object(parts, {k: loc[k] for k in names + ("mu",)})
        if rank != 0:
            return None
        out = {k: sharding.unshard_table([p[k] for p in parts]) for k in names}
        out["mu"] = parts[0]["mu"]
        return out
    t0 = assembled()
    gen = torch.Generator(device=dev)
    gen.manual_seed(777 + rank)
    orc = None
    if rank == 0:
        orc = oracle.SvdOracle(t0["mu"], t0["user_bias"], t0["item_bias"], t0["user_feat"], t0["item_feat"], LR, REG)
    worst_logit = 0.0
    for _ in range(steps):
        us, it, rt = _draw_slice(torch, gen, hi - lo, w["U"], w["I"], dev)
        logits, _ = eng.step(us, it, rt)
        if logits.numel() != hi - lo:      # the all-reduce exchange computes every rank's predictions: keep my slice
            logits = logits[lo:hi]
        parts = [None] * G
        dist.all_gather_object(parts, (us.cpu().numpy(), it.cpu().numpy(), rt.cpu().numpy(), logits.cpu().numpy()))
        if rank == 0:
            gu, gi, gr, gl = (np.concatenate([p[k] for p in parts]) for k in range(4))
            ref_logits, _ = orc.train_step(gu, gi, gr)
            worst_logit = max(worst_logit, float(np.max(np.abs(gl - ref_logits) / (np.abs(ref_logits) + 1e-3))))
    t1 = assembled()
    if rank != 0:
        return None
    res = {"steps": steps, "world": G, "max_rel_logit_err": worst_logit, "tables": {}}
    ok = worst_logit <= 1e-4
    for k in names + ("mu",):
        ref = getattr(orc, k).reshape(-1).astype(np.float64)
        got = t1[k].reshape(-1).astype(np.float64)
        nref = np.linalg.norm(ref)
        rms = nref / np.sqrt(max(ref.size, 1))
        rel = float(np.linalg.norm(got - ref) / max(nref, 1e-30))
        outside = int(np.sum(np.abs(got - ref) > 1e-5 * np.abs(ref) + 1e-5 * rms + 1e-30))
        res["tables"][k] = {"rel_l2": rel, "outside_1e-5": outside, "n": int(ref.size)}
        ok = ok and rel <= 1e-5 and outside <= 1e-4 * ref.size
    res["pass"] = bool(ok)
    return res


def main(args):
    import bench
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            bench.emit(reference_line(args, max(world, args.gpus), bench))
        return
    import torch
    import torch.distributed as dist
    from tf_recomm_b200 import _lib
    from tf_recomm_b200._lib import check
    from tf_recomm_b200.sharded import ShardedSvdEngine
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    assert torch.cuda.is_available(), "bench.py needs CUDA devices; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    w = dict(SHARDED)
    if os.environ.get("TFR_SHARDED_SCALE"):       # smaller tables for smoke runs: TFR_SHARDED_SCALE=100
        s = int(os.environ["TFR_SHARDED_SCALE"])
        w["U"], w["I"] = w["U"] // s, w["I"] // s
    B, d = w["B"], w["d"]
    steps = args.steps or 30
    warmup = max(args.warmup, 3)
    eng = ShardedSvdEngine(w["U"], w["I"], d, LR, REG, rank, world, device=dev)
    # default: the all-gather + all-reduce exchange (measured faster at 2 GPUs: 0.59 against 0.74 ms of non-pass time per
    # step); TFR_SHARDED_EXCHANGE=a2a selects the all-to-all exchange north_star names
    exchange = os.environ.get("TFR_SHARDED_EXCHANGE", "allreduce")
    if exchange == "a2a":
        eng.step = lambda u, i, r, nxt=None: eng.train_step_a2a(u, i, r, next_slice=nxt)
    else:
        eng.step = lambda u, i, r, nxt=None: eng.train_step_from_slices(u, i, r)
    parity = None
    if os.environ.get("TFR_SHARDED_SCALE"):
        parity = in_run_parity_check(eng, torch, dist, w, world, rank, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(13575 + rank)
    lo, hi = eng_slice(B, world, rank)
    slices = [_draw_slice(torch, gen, hi - lo, w["U"], w["I"], dev) for _ in range(warmup + steps)]
    def nxt(k):
        return slices[k + 1][:2] if k + 1 < len(slices) else None
    for k in range(warmup):
        eng.step(*slices[k], nxt(k))
    torch.cuda.synchronize()
    dist.barrier()
    sampler = bench.ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(warmup, warmup + steps):
        eng.step(*slices[k], nxt(k))
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    secs = float(ms.item()) / 1e3
    clocks = sampler.stop() if sampler else None

    # roofline of the dominant kernel (the local Adam pass), timed live with events inside complete steps
    pass_ms, n_pass = 0.0, min(steps, 5)
    bufs = eng._buffers(B)
    for k in range(n_pass):
        us, it, rt = slices[warmup + k]
        parts = torch.stack([us.view(torch.float32), it.view(torch.float32), rt])
        out = torch.empty((world,) + tuple(parts.shape), dtype=parts.dtype, device=dev)
        dist.all_gather_into_tensor(out, parts)
        gu = out[:, 0].reshape(-1).view(torch.int32).contiguous()
        gi = out[:, 1].reshape(-1).view(torch.int32).contiguous()
        gr = out[:, 2].reshape(-1).contiguous()
        eng.gather_owned(gu, gi, bufs)
        dist.all_reduce(bufs["flat"])
        pass_ms += timed_local_step(eng, bufs, gr, torch, _lib, check)
    pass_ms /= n_pass
    params_local = (eng.U_loc + eng.I_loc) * (d + 1)
    peak, peak_src = bench.measured_peaks()
    pass_gbs = 24.0 * params_local / (pass_ms / 1e3) / 1e9
    t = torch.tensor([pass_gbs], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    pass_gbs = float(t.item())

    # e2e: each rank's slice comes from pinned host memory, its slice of the predictions goes back
    n_e2e = min(steps, 10)
    host = [tuple(x.cpu().pin_memory() for x in s) for s in slices[:n_e2e]]
    pred_host = torch.empty(hi - lo, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for hs in host:
        ds = [x.to(dev, non_blocking=True) for x in hs]
        _, infer = eng.step(*ds)
        pred_host.copy_(infer if exchange == "a2a" else infer[lo:hi], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_val = B * n_e2e / float(dt.item())

    if rank == 0:
        bytes_step = 24 * (w["U"] + w["I"]) * (d + 1) + 8 * B * (d + 1) + 12 * B
        line = {
            "metric": "train ratings/sec", "value": B * steps / secs, "unit": "ratings/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(sharded_config(w, world), **{
                       "exchange": "NCCL all-to-all of ids, rows back, gradient records to the owners (exact per-peer "
                                   "counts, exchanged a step ahead)" if exchange == "a2a" else
                                   "all_gather(ids) + all_reduce(owner-filled rows)",
                       "scaling_baseline": "n_gpus=2: the 170 GB of tables + Adam state do not fit one 180 GB GPU",
                       "l2_policy": "local shard %.1f GB per step >> 126 MB L2: no flush needed" % (24 * params_local / 1e9),
                       "timing": "CUDA events between barriers, max over ranks"}),
            "hbm": {"algorithmic_bytes_per_step": bytes_step, "achieved_gbs_per_gpu": bytes_step / world / (secs / steps) / 1e9,
                    "frac_of_measured_peak": bytes_step / world / (secs / steps) / 1e9 / peak, "peak_gbs": peak,
                    "peak_source": peak_src},
            "roofline": {"bound": "hbm", "kernel": "adam_stream_multi_kernel (local shard, min over ranks)",
                         "achieved": pass_gbs, "peak": peak, "unit": "GB/s", "frac": pass_gbs / peak,
                         "peak_source": peak_src, "bytes_per_launch": 24.0 * params_local, "launch_ms": pass_ms,
                         "traffic": None},
            "comm": {"per_rank_bytes_per_step": eng.a2a_exchange_bytes(B) if exchange == "a2a" else eng.exchange_bytes(B),
                     "collectives": "3 x all_to_all_single + all_reduce(2 scalars) + all_gather(counts, a step ahead) [NCCL]"
                     if exchange == "a2a" else "all_gather(ids) + all_reduce(rows) [NCCL]",
                     "step_minus_pass_ms": secs / steps * 1e3 - pass_ms},
            "scaling_efficiency_vs_2gpu": scaling_vs_2gpu(B * steps / secs, world),
            "parity_check": parity,
            "cpu_baseline": None,
            "e2e": {"value": e2e_val, "unit": "ratings/s", "h2d_bytes_per_step": 12 * B, "d2h_bytes_per_step": 4 * B,
                    "steps": n_e2e, "path": "ShardedSvdEngine.train_step_from_slices with pinned host slices"},
            "clocks": clocks, "gpu_launches": (2 + 6) * steps,
        }
        bench.emit(line)
    dist.barrier()
    dist.destroy_process_group()


def scaling_vs_2gpu(value, world):
    """value_N / (value_2 * N / 2) with value_2 from the committed 2-GPU measurement of the same exchange
    (profiles/r02_scaling_sharded.json); None until that file exists.  The driver computes its own from SCALE_rNN."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_scaling_sharded.json")
    try:
        v2 = json.load(open(p))["value_2gpu"]
    except Exception:
        return None
    return {"efficiency": value / (v2 * world / 2.0), "value_2gpu": v2, "source": "profiles/r02_scaling_sharded.json"}


def eng_slice(B, world, rank):
    from tf_recomm_b200 import sharding
    return sharding.batch_slice(B, world, rank)


def timed_local_step(eng, bufs, rates, torch, _lib, check):
    """One local step issued piecewise so that CUDA events bracket the Adam pass alone; returns its ms."""
    e, L = eng.local, eng.L
    B, d = rates.numel(), eng.d
    t = bufs["tables"]
    tp = C.byref(t)
    ws_t = e.workspace(B)
    ws = e.step_ws(B)
    st = e._stream()
    opt = e.opt.data_ptr()
    ku, ki = bufs["key_u"].data_ptr(), bufs["key_i"].data_ptr()
    check(L.tfr_svd_begin_step(opt, st))
    check(L.tfr_svd_fwd_err(tp, opt, ku, ki, rates.data_ptr(), B, bufs["logits"].data_ptr(), bufs["infer"].data_ptr(),
                            C.byref(ws), st))
    check(L.tfr_dedup_sort_pairs(ku, eng.U_loc + 1, ws.su_ids, ws.su_pos, ki, eng.I_loc + 1, ws.si_ids, ws.si_pos, B,
                                 ws.sort_ws, ws.sort_ws_bytes, st))
    check(L.tfr_svd_segment_grads(tp, opt, ku, ki, B, C.byref(ws), st))
    T, S = e.t, e.slots
    arr = (_lib.AdamTable * 4)()
    for k, (tab, rows, width, slot, gsum) in enumerate((("user_feat", eng.U_loc, d, e.user_slot, ws.gsum_uf),
                                                        ("item_feat", eng.I_loc, d, e.item_slot, ws.gsum_if),
                                                        ("user_bias", eng.U_loc, 1, e.user_slot, ws.gsum_ub),
                                                        ("item_bias", eng.I_loc, 1, e.item_slot, ws.gsum_ib))):
        arr[k].var, arr[k].m, arr[k].v = T[tab].data_ptr(), S["m_" + tab].data_ptr(), S["v_" + tab].data_ptr()
        arr[k].rows, arr[k].width, arr[k].slot, arr[k].gsum = rows, width, slot.data_ptr(), gsum
        arr[k].stride = (e.feat_stride if width > 1 else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(L.tfr_adam_stream_multi(arr, 4, opt, 15, st))
    e1.record()
    import bench
    check(L.tfr_svd_finish_step(tp, opt, ku, ki, B, C.byref(ws), bench._n_partials(d, B), st))
    torch.cuda.synchronize()
    del ws_t
    return e0.elapsed_time(e1)
