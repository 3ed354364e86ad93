"""Multi-GPU leg of bench.py: R ranks, rows hash-sharded by farmhash64(sign) % R, data-parallel batches,
forward / backward all-to-all over NCCL (BASELINE configs[2..3]).  Weak scaling: every rank keeps
`--rows` resident rows and a batch of `--batch` samples."""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def run(args, rank, local_rank, world, METRIC, UNIT, workload_config, ClockSampler, cpu_arm):
    from . import native as N
    from . import shard as SH
    from . import workload as W
    from .worker import CudaBackend, ShardedEmbeddingWorker

    dev = torch.device("cuda", local_rank)
    # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the one JSON line
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)
        torch.cuda.synchronize()
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    lib = N.load()
    dim = args.dim or 64
    S, B, K, Wm = args.slots, args.batch, args.steps, max(args.warmup, 3)
    rows_total = int(args.rows) * world
    card = W.scaled_cardinalities(rows_total, S)
    pf = W.index_prefixes(S)
    n_occ = S * B
    cap = int(int(args.rows) * 1.02) + 4096
    be = CudaBackend(dim, cap, dev, dict(kind=N.OPT_ADAGRAD, lr=0.01, initialization=0.01, eps=1e-10), {},
                     max_occurrences=max(4 * n_occ, 1 << 16))
    wk = ShardedEmbeddingWorker(S, dim, pf, be)

    # ---- make this rank's share of every slot resident
    t_fill = time.time()
    chunk = 1 << 21
    buf = torch.empty((chunk, dim), dtype=torch.float32, device=dev)
    for s in range(S):
        for lo in range(0, int(card[s]), chunk):
            hi = min(int(card[s]), lo + chunk)
            ids = torch.arange(lo, hi, dtype=torch.int64, device=dev)
            signs = SH.add_prefix(ids, [0, hi - lo], [pf[s]])
            mine = signs[SH.shard_of(signs, world) == rank].contiguous()
            if mine.numel():
                be.shard.lookup(mine, training=True, out=buf[: mine.numel()])
    torch.cuda.synchronize()
    resident = len(be.shard)
    t_fill = time.time() - t_fill
    del buf
    tot = torch.tensor([resident], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    assert int(tot) == rows_total, (int(tot), rows_total)
    assert be.shard.counters()["capacity_refused"] == 0

    n_sets = max(2, args.sets)
    ids_host = W.make_batches(100 + rank, card, B, n_sets, args.alpha)  # every rank draws its own samples
    ids_pinned = torch.from_numpy(ids_host.view(np.int64)).pin_memory()
    ids_dev = [ids_pinned[k].to(dev) for k in range(n_sets)]
    g = torch.Generator(device=dev)
    g.manual_seed(5 + rank)
    grads = (torch.randn((n_sets, S, B, dim), generator=g, device=dev) * 1e-2).half()

    static = not args.dist_dynamic
    mode = "dynamic"
    gstream = torch.cuda.Stream(device=dev)
    if static:
        # slots per GPU pair: sized from the batches themselves (+10 %); a production loop sizes it on its warm-up
        # batches and re-runs a batch that raises the overflow flag with a larger capacity
        wk.enable_static(B, cap=wk.calibrate_cap(ids_dev, B))
        mode = "nccl-framed"
    if static and not args.dist_nccl:
        try:
            wk.enable_p2p(B)
            mode = "p2p"
            be.ctx.set_async_grouping(True)  # forward and backward of a step share one capture: the grouping may overlap
        except Exception as e:  # noqa: BLE001 — no symmetric memory on this box: NCCL framed path
            print(f"[bench] peer-memory exchange unavailable ({e!r}); using NCCL", file=__import__("sys").stderr)

    def eager_step(k):
        if mode == "p2p":
            wk.forward_p2p(ids_dev[k], B, training=True)
            wk.backward_p2p(grads[k])
        elif mode == "nccl-framed":
            wk.forward_static(ids_dev[k], B, training=True)
            wk.backward_static(grads[k])
        else:
            wk.forward(ids_dev[k], B, training=True)
            wk.backward(grads[k])

    graphs, seg_steps = None, None
    if mode == "p2p" and not args.no_graph:
        # kernels + peer-memory exchanges + flag barriers of a whole step: one CUDA graph per buffer set, no NCCL inside
        with torch.cuda.stream(gstream):
            for i in range(3):
                eager_step(i % n_sets)
            gstream.synchronize()
            dist.barrier()
            graphs = []
            for k in range(n_sets):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=gstream, capture_error_mode="thread_local"):
                    eager_step(k)
                graphs.append(gph)
        torch.cuda.synchronize()
        dist.barrier()
    elif mode == "nccl-framed" and not args.no_graph:
        seg_steps = [wk.make_graphed_step(ids_dev[k], grads[k], B, gstream)[0] for k in range(n_sets)]
        torch.cuda.synchronize()
        dist.barrier()

    def step(k):
        with torch.cuda.stream(gstream):
            if graphs is not None:
                graphs[k].replay()
            elif seg_steps is not None:
                seg_steps[k]()
            else:
                eager_step(k)

    def timed(fn, n):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tstream = gstream if fn is step else torch.cuda.current_stream()
        e0.record(tstream)
        for i in range(n):
            fn(i % n_sets)
        e1.record(tstream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # device time, max over ranks
        dist.barrier()
        return float(ms)

    for i in range(Wm):
        step(i % n_sets)
    torch.cuda.synchronize()
    l0 = lib.pb_launch_count()
    step(0)
    launches_per_step = int(lib.pb_launch_count() - l0)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    t0 = time.time()
    ms = timed(step, K)
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None

    ids_stage = torch.empty(n_occ, dtype=torch.int64, device=dev)
    probe_host = torch.empty(4, dtype=torch.float16).pin_memory()

    def e2e_step(k):
        ids_stage.copy_(ids_pinned[k], non_blocking=True)
        if mode == "p2p":
            out = wk.forward_p2p(ids_stage, B, training=True)
            wk.backward_p2p(grads[k])
        elif mode == "nccl-framed":
            out = wk.forward_static(ids_stage, B, training=True)
            wk.backward_static(grads[k])
        else:
            out = wk.forward(ids_stage, B, training=True)
            wk.backward(grads[k])
        probe_host.copy_(out.view(-1)[:4], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(Wm):
        e2e_step(i % n_sets)
    ms_e2e = timed(e2e_step, K)

    breakdown = None
    if mode == "nccl-framed" and os.environ.get("PB_DIST_BREAKDOWN"):
        # eager framed step with CUDA events between its segments (diagnostic; rank 0's view)
        names = ["prefix+partition+frame", "a2a signs", "owner forward", "a2a rows", "unframe", "frame grads", "a2a grads", "owner backward"]
        acc = [0.0] * len(names)
        S_, R_, cap_ = S, world, wk.cap
        for it in range(20):
            k = it % n_sets
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            slot_off = [s * B for s in range(S_ + 1)]
            ev[0].record()
            signs = be.add_prefix(ids_dev[k], slot_off, pf, 8)
            perm, counts = be.partition(signs, R_)
            send = be.frame_signs(signs, perm, counts, R_, cap_, wk.overflow)
            ev[1].record()
            recv = wk._a2a_equal(send)
            ev[2].record()
            rows = be.serve_lookup(recv, True)
            ev[3].record()
            back = wk._a2a_equal(rows)
            ev[4].record()
            out = be.frame_rows(back, perm, counts, R_, cap_, False, be.empty_rows(n_occ))
            ev[5].record()
            gs = be.frame_rows(grads[k].reshape(n_occ, dim), perm, counts, R_, cap_, True, be.empty_rows(R_ * cap_, grads.dtype))
            ev[6].record()
            gr = wk._a2a_equal(gs)
            ev[7].record()
            be.serve_update(gr, 1.0)
            ev[8].record()
            torch.cuda.synchronize()
            if it >= 4:
                for i in range(len(names)):
                    acc[i] += ev[i].elapsed_time(ev[i + 1]) * 1e3 / 16
        breakdown = {n: round(v, 1) for n, v in zip(names, acc)}
    overflowed = wk.check_overflow() if static else False
    assert not overflowed, "framed exchange overflowed its capacity: raise the slack"
    if mode == "p2p":
        assert not wk.check_p2p(), "a peer-memory barrier timed out"
    if rank == 0:
        ms_per_step = ms / K
        GB = B * world
        state = dim
        bytes_per_id = W.algorithmic_bytes_per_id(dim, state, "total")
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        whole = n_occ * bytes_per_id / (ms_per_step * 1e-3) / 1e9  # per GPU
        line = {
            "metric": METRIC, "value": GB / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, dim, card, world), resident_rows_rank0=resident,
                           table_fill_seconds=round(t_fill, 2),
                           l2="inputs larger than L2: %.1f GB table per GPU + %d rotating id/grad sets" % (
                               resident * 4.0 * (dim + state) / 1e9, n_sets),
                           launch={"p2p": "framed exchange (cap %d slots per GPU pair) stored straight into the peers' buffers over NVLink by "
                                          "libpersia_b200's own kernels + flag barriers; %s" % (
                                              getattr(wk, "cap", 0), "whole step = one CUDA graph per rank" if graphs is not None else "kernel by kernel"),
                                   "nccl-framed": "framed exchange (cap %d) over NCCL all_to_all_single; %s" % (
                                       getattr(wk, "cap", 0), "compute segments replay as CUDA graphs" if seg_steps is not None else "kernel by kernel"),
                                   "dynamic": "NCCL all_to_all_single with split sizes; one host sync per step"}[mode]),
            "clocks": clocks,
            "e2e": {"value": GB / (ms_e2e / K * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_occ * 8 * world,
                    "d2h_bytes_per_step": 8 * world, "ms_per_step": ms_e2e / K,
                    "path": "pinned host ids -> H2D -> ShardedEmbeddingWorker.forward/backward -> D2H of 4 output "
                            "values, host sync every step"},
            "gpu_launches": launches_per_step * K,
            "roofline": {"bound": "hbm", "kernel": "whole step per GPU (multi-GPU runs report no per-kernel split)",
                         "achieved": whole, "peak": peak, "unit": "GB/s", "frac": whole / peak, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
            "cpu_baseline": None,
            "segments_us": breakdown,
        }
        print(json.dumps(line))
    dist.barrier()
    torch.cuda.synchronize()
    import sys

    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)  # CUDA graphs + NCCL teardown order is fragile; everything has been reported
