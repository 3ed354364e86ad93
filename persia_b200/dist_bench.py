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
    if static:
        wk.enable_static(B)

    def step(k):
        if static:
            wk.forward_static(ids_dev[k], B, training=True)
            wk.backward_static(grads[k])
        else:
            wk.forward(ids_dev[k], B, training=True)
            wk.backward(grads[k])

    graphs = None
    graph_error = None
    if static and args.dist_graph:  # static shapes: kernels + NCCL collectives of a step replay as one CUDA graph per buffer set
        for i in range(3):
            step(i % n_sets)
        torch.cuda.synchronize()
        dist.barrier()
        try:
            graphs = []
            gstream = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(gstream):
                for k in range(n_sets):
                    gph = torch.cuda.CUDAGraph()
                    # thread_local: the NCCL watchdog thread keeps polling its events while this thread captures
                    with torch.cuda.graph(gph, stream=gstream, capture_error_mode="thread_local"):
                        step(k)
                    graphs.append(gph)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001 — fall back to eager static steps, say so in the output
            graphs = None
            graph_error = repr(e)[:200]
    seg_steps = None
    if static and graphs is None and not args.no_graph:
        gstream = torch.cuda.Stream(device=dev)
        seg_steps = [wk.make_graphed_step(ids_dev[k], grads[k], B, gstream)[0] for k in range(n_sets)]
        torch.cuda.synchronize()
        dist.barrier()
    eager_step = step

    def step(k):  # noqa: F811
        if graphs is not None:
            graphs[k].replay()
        elif seg_steps is not None:
            with torch.cuda.stream(gstream):
                seg_steps[k]()
        else:
            eager_step(k)

    def timed(fn, n):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tstream = gstream if (seg_steps is not None and fn is step) else torch.cuda.current_stream()
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
        if static:
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

    overflowed = wk.check_overflow() if static else False
    assert not overflowed, "framed exchange overflowed its capacity: raise the slack"
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
                           launch=("fixed-capacity framed exchange (cap %d slots per GPU pair), step = one CUDA graph incl. the NCCL "
                                   "all-to-alls" % wk.cap) if (static and graphs is not None) else
                                  ("fixed-capacity framed exchange (cap %d slots per GPU pair); the compute segments between the 3 NCCL "
                                   "all-to-alls replay as CUDA graphs" % wk.cap) if seg_steps is not None else
                                  ("fixed-capacity framed exchange, kernel by kernel" if static else
                                   "kernel by kernel (NCCL all_to_all_single with split sizes; one host sync per step)")),
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
        }
        print(json.dumps(line))
    dist.barrier()
    torch.cuda.synchronize()
    import sys

    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)  # CUDA graphs + NCCL teardown order is fragile; everything has been reported
