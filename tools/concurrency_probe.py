"""Does work launched eagerly on a second stream start while a long bandwidth-bound kernel runs on the first?
Plain torch ops, nothing of this library: A = in-place scale of a 1 GiB tensor (~350 us), B = a tiny fill / a 256 KiB D2H
copy / a 768 KiB H2D copy, launched ~40 us after A.  Prints when B finished relative to A's start and end.
Usage (GPU box): python tools/concurrency_probe.py"""
import time

import torch


def probe(kind, hold_us=40.0):
    dev = torch.device("cuda:0")
    x = torch.ones(256 << 20, dtype=torch.float32, device=dev)
    y = torch.zeros(1024, dtype=torch.float32, device=dev)
    d = torch.zeros(64 << 10, dtype=torch.float32, device=dev)
    h = torch.zeros(64 << 10, dtype=torch.float32).pin_memory()
    d3 = torch.zeros(192 << 10, dtype=torch.float32, device=dev)
    h3 = torch.zeros(192 << 10, dtype=torch.float32).pin_memory()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    res = []
    for rep in range(8):
        a0, a1, b0, b1 = ev(), ev(), ev(), ev()
        torch.cuda.synchronize()
        with torch.cuda.stream(sa):
            a0.record()
            x.mul_(1.0)
            a1.record()
        t = time.perf_counter()
        while (time.perf_counter() - t) * 1e6 < hold_us:
            pass
        with torch.cuda.stream(sb):
            b0.record()
            if kind == "fill":
                y.fill_(1.0)
            elif kind == "d2h":
                h.copy_(d, non_blocking=True)
            elif kind == "h2d":
                d3.copy_(h3, non_blocking=True)
            b1.record()
        torch.cuda.synchronize()
        res.append((a0.elapsed_time(a1) * 1e3, a0.elapsed_time(b0) * 1e3, a0.elapsed_time(b1) * 1e3))
    res = res[2:]
    med = lambda k: sorted(r[k] for r in res)[len(res) // 2]
    print("%-5s A takes %6.1f us; B issued at %6.1f us, finished at %6.1f us after A's start" % (kind, med(0), med(1), med(2)))


if __name__ == "__main__":
    for kind in ("fill", "d2h", "h2d"):
        probe(kind)


def dep_probe(kind, graph):
    """s0: small kernel K; event e1; long kernel A.   s1: wait e1; B (fill / D2H); event.  When does B finish?
    (the shape of the feed step: forward -> {copy of the predictions, next sort} beside the table pass)"""
    dev = torch.device("cuda:0")
    x = torch.ones(256 << 20, dtype=torch.float32, device=dev)
    k = torch.ones(1 << 20, dtype=torch.float32, device=dev)
    y = torch.zeros(1024, dtype=torch.float32, device=dev)
    d = torch.zeros(64 << 10, dtype=torch.float32, device=dev)
    h = torch.zeros(64 << 10, dtype=torch.float32).pin_memory()
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()

    def body():
        k.mul_(1.0)
        e1 = torch.cuda.Event()
        e1.record()
        s1.wait_event(e1)
        with torch.cuda.stream(s1):
            if kind == "fill":
                y.fill_(1.0)
            else:
                h.copy_(d, non_blocking=True)
            e2 = torch.cuda.Event()
            e2.record()
        x.mul_(1.0)
        torch.cuda.current_stream().wait_event(e2)

    res = []
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s0):
            body()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s0):
                body()
    for rep in range(8):
        torch.cuda.synchronize()
        h.zero_()
        d.fill_(float(rep + 1))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s0):
            if graph:
                g.replay()
            else:
                body()
        t_seen = None
        if kind == "d2h":
            while h[-1].item() != float(rep + 1):   # the host polls the pinned buffer itself
                pass
            t_seen = (time.perf_counter() - t0) * 1e6
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) * 1e6
        res.append((t_seen, t_all))
    res = res[2:]
    seen = sorted(r[0] for r in res)[len(res) // 2] if kind == "d2h" else float("nan")
    print("dep %-4s %-5s: host sees B's data %6.1f us after issue; everything done at %6.1f us"
          % (kind, "graph" if graph else "eager", seen, sorted(r[1] for r in res)[len(res) // 2]))


if __name__ == "__main__":
    for graph in (False, True):
        dep_probe("d2h", graph)
