"""bench_dist.py — multi-GPU leg of bench.py (bench infrastructure, not part of the persia_b200 package: it uses the
oracle as a checker).   R ranks (one process per GPU), rows hash-sharded by farmhash64(sign) % R, data-parallel
batches, the exchange fused into the kernels over NVLink peer memory (BASELINE configs[2..3]).  Weak scaling: every rank
keeps `--rows` resident rows and a batch of `--batch` samples (dim 128 by default: the metric's config).

What is timed is one CUDA graph per rank and buffer set holding the whole step (pb_forward_sharded + pb_backward_sharded:
compute kernels, peer stores, flag waits).  After the timed loops the same captured graphs are replayed from the live
tables and checked against the oracle (R parameter servers, the R ranks' requests applied in rank order): outputs and
every row this rank owns among the touched ones, bit for bit — `parity_checked`."""
import json
import os
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist


def _check_parity(rank, world, wk, replay, outs, all_ids, all_grads, pf, S, B, dim, dev, sets):
    """Every rank runs its own oracle worker with R parameter servers over ALL ranks' check batches (inputs are
    regenerated from their seeds), seeds only ITS shard with the live rows and compares only what it owns: rows of one
    shard never depend on another shard's.  Forward outputs are compared where the sign is owned by this rank."""
    import oracle

    owned = []
    for k in sets:
        for q in range(world):
            sg = np.concatenate([oracle.add_prefix(all_ids[q][k][i * B:(i + 1) * B], 8, pf[i]) for i in range(S)])
            owned.append(sg[oracle.shard_of(sg, world) == rank])
    signs = np.unique(np.concatenate(owned))
    d_signs = torch.from_numpy(signs.view(np.int64)).to(dev)
    ent, found = wk.shard.get_entries(d_signs)
    found = found.cpu().numpy()
    w = oracle.Worker([oracle.SlotCfg(dim, prefix=p) for p in pf], n_ps=world, capacity_per_ps=1 << 40)
    w.configure()
    w.set_optimizer(oracle.Optim(oracle.ADAGRAD, lr=0.01, init_acc=0.01, eps=1e-10))
    w.set_embedding(signs[found], ent.cpu().numpy()[found], dim)  # signs not resident yet are admitted by both sides
    oracle.set_rsqrt_exact(True)
    bad_out = 0
    try:
        row_off = np.arange(S * B + 1, dtype=np.uint32)
        for k in sets:
            dist.barrier()
            replay(k)
            torch.cuda.synchronize()
            dist.barrier()
            octx = [w.forward(all_ids[q][k], row_off, B, training=True) for q in range(world)]
            got = outs[k].cpu().numpy()
            mine = np.concatenate([oracle.add_prefix(all_ids[rank][k][i * B:(i + 1) * B], 8, pf[i]) for i in range(S)])
            mask = (oracle.shard_of(mine, world) == rank).reshape(S, B)
            for i in range(S):
                bad_out += int((got[i][mask[i]].view(np.uint16) != octx[rank][0][i][mask[i]].view(np.uint16)).any(axis=1).sum())
            for q in range(world):  # the owner applies the R requests in rank order
                w.backward(octx[q][1], [all_grads[q][k][i] for i in range(S)])
        ent2, found2 = wk.shard.get_entries(d_signs)
        ent2 = ent2.cpu().numpy()
        assert bool(found2.all())
        bad_rows = sum(ent2[j].tobytes() != w.get_entry(int(s)).tobytes() for j, s in enumerate(signs))
    finally:
        oracle.set_rsqrt_exact(False)
    return bad_out, bad_rows, int(signs.size)


def model_leg_dist(args, rank, world, local_rank, B_, steps=30, warmup=8):
    """e2e_model at N > 1: persia_b200.api.TrainCtx on every rank — DDP dense tower (persia/distributed.py:174-191), the
    embeddings behind PersiaCommonContext(replica_size = world) served by a ShardedEmbeddingWorker — driven from host numpy
    batches, wall clock around the loop, max over ranks."""
    from persia_b200 import api
    from persia_b200 import persia_core as PC
    from persia_b200 import workload as W

    S, B, dim, n_dense = args.slots, args.batch, args.dim, 13
    names = [f"C{i + 1}" for i in range(S)]
    dev = torch.device("cuda", local_rank)
    PC.reset()
    PC._S.capacity = int(2e7 / world) + (1 << 20)
    PC.set_embedding_config({"slots_config": {n: {"dim": dim} for n in names}})
    card = W.scaled_cardinalities(int(2e7), S)
    torch.manual_seed(0)
    prev_precision = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("high")
    model = W.make_dlrm_tower(S, dim, n_dense=n_dense).cuda()
    dense_opt = torch.optim.SGD(model.parameters(), lr=0.01)
    loss_fn = torch.nn.BCEWithLogitsLoss()
    n_pool = 8
    ids_pool = W.make_batches(7 + rank, card, B, n_pool, args.alpha).reshape(n_pool, S, B)
    rng = np.random.default_rng(11 + rank)
    dense_pool = rng.standard_normal((n_pool, B, n_dense)).astype(np.float32)
    label_pool = (rng.random((n_pool, B, 1)) < 0.25).astype(np.float32)
    try:
        with api.TrainCtx(model=model, embedding_optimizer=api.Adagrad(lr=0.01, initial_accumulator_value=0.01, eps=1e-10),
                          dense_optimizer=dense_opt, device_id=local_rank, mixed_precision=False) as ctx:
            def step(k):
                pb = api.PersiaBatch([api.IDTypeFeatureWithSingleID(names[i], ids_pool[k, i]) for i in range(S)],
                                     non_id_type_features=[api.NonIDTypeFeature(dense_pool[k], name="dense")],
                                     labels=[api.Label(label_pool[k], name="click")], requires_grad=True)
                out, labels = ctx.forward(ctx.get_embedding_from_data(pb))
                loss = loss_fn(out, labels[0].squeeze(1))
                ctx.backward(loss)
                return loss

            for i in range(warmup):
                step(i % n_pool)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.time()
            first = last = None
            for i in range(steps):
                last = step(i % n_pool)
                first = last if first is None else first
            ctx.backward_engine.flush()
            torch.cuda.synchronize()
            dt = torch.tensor([time.time() - t0], device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.barrier()
    finally:
        PC.reset()
        torch.set_float32_matmul_precision(prev_precision)
    dt = float(dt)
    return {"value": B * world * steps / dt, "unit": B_.UNIT, "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "first_loss_rank0": float(first.detach()), "last_loss_rank0": float(last.detach()),
            "path": "numpy batch -> api.PersiaBatch -> TrainCtx.get_embedding_from_data (ShardedEmbeddingWorker over %d GPUs) -> "
                    "DLRM tower under DistributedDataParallel (TF32 matmuls) -> BCE -> TrainCtx.backward (dense SGD all-reduce + "
                    "sparse Adagrad on the owners); 2e7-id key space, rows admitted on the fly; wall clock, max over ranks" % world}


def run(args, rank, local_rank, world, B_):
    from persia_b200 import native as N
    from persia_b200 import shard as SH
    from persia_b200 import workload as W
    from persia_b200.worker import ShardedEmbeddingWorker

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
    dim = args.dim
    S, B, K, Wm = args.slots, args.batch, args.steps, max(args.warmup, 3)
    rows = int(args.rows)
    ks = B_.keyspace_per_gpu(args)
    bounded = ks > rows
    card = W.scaled_cardinalities(int(ks) * world, S)       # the key space the ids are drawn from
    fill_card = W.scaled_cardinalities(rows * world, S)      # what is made resident before timing
    pf = W.index_prefixes(S)
    n_occ = S * B
    n_sets = max(2, args.sets)
    n_sets += n_sets % 2  # (the e2e staging buffers alternate with the sets)
    all_ids = None
    ids_host = W.make_batches(100 + rank, card, B, n_sets, args.alpha)  # every rank draws its own samples
    cap = ShardedEmbeddingWorker.calibrate_cap([ids_host[k] for k in range(n_sets)], B, pf, world)
    table_cap = int(rows * 1.02) + 4096
    wk = ShardedEmbeddingWorker.distributed(S, dim, pf, table_cap, cap, dict(kind=N.OPT_ADAGRAD, lr=0.01, initialization=0.01, eps=1e-10),
                                            dev, max_batch=B)
    # ---- make this rank's share of every slot resident
    t_fill = time.time()
    B_.fill_table(torch, SH, wk.shard, fill_card, pf, dev, dim, owner=(rank, world))
    resident = len(wk.shard)
    t_fill = time.time() - t_fill
    tot = torch.tensor([resident], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    assert int(tot) == rows * world, (int(tot), rows * world)
    assert wk.shard.counters()["capacity_refused"] == 0
    if bounded:  # configs[3]: ids from a key space far larger than the table; least recently used rows make room
        free_now = table_cap - resident
        wk.shard.set_eviction(check_every=1, low_water=max(1 << 20, free_now // 2), target_free=max(1 << 22, 2 * free_now),
                              keep_batches=2)

    ids_pinned = torch.from_numpy(ids_host.view(np.int64)).pin_memory()
    ids_dev = [ids_pinned[k].to(dev) for k in range(n_sets)]
    g = torch.Generator(device=dev)
    g.manual_seed(5 + rank)
    grads = (torch.randn((n_sets, S, B, dim), generator=g, device=dev) * 1e-2).half()
    outs = [torch.empty((S, B, dim), dtype=torch.float16, device=dev) for _ in range(n_sets)]
    gstream = torch.cuda.Stream(device=dev)
    wk.stream = gstream

    def eager_step(k):
        wk.forward(ids_dev[k], B, training=True, out=outs[k])
        wk.backward(grads[k])

    # kernels + peer stores + flag waits of a whole step: one CUDA graph per rank and buffer set, no NCCL inside
    with torch.cuda.stream(gstream):
        for i in range(3):
            eager_step(i % n_sets)
        gstream.synchronize()
        l0 = lib.pb_launch_count()
        eager_step(0)
        launches_per_step = int(lib.pb_launch_count() - l0)
        gstream.synchronize()
        dist.barrier()
        graphs = None
        if not args.no_graph:
            graphs = []
            for k in range(n_sets):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=gstream, capture_error_mode="thread_local"):
                    eager_step(k)
                graphs.append(gph)
    torch.cuda.synchronize()
    dist.barrier()

    def step(k):
        with torch.cuda.stream(gstream):
            if graphs is not None:
                graphs[k].replay()
            else:
                eager_step(k)

    def timed(fn, n):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(gstream)
        for i in range(n):
            fn(i % n_sets)
        e1.record(gstream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # device time, max over ranks
        dist.barrier()
        return float(ms)

    for i in range(Wm):
        step(i % n_sets)
    torch.cuda.synchronize()
    stats = wk.ctx.batch_stats()
    admitted0 = len(wk.shard)
    sampler = B_.ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    t0 = time.time()
    ms = timed(step, K)
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    reps = [timed(step, K) / K for _ in range(3)]

    # ---- e2e: pinned host ids -> H2D -> sharded forward + backward -> D2H of the slot status, host sync every step;
    # the same captured graph mechanism, with the copies inside
    # the ids of step i + 1 travel on a copy stream while step i computes (the reference's Forward engine prefetches
    # batches the same way); every copy is inside the timed region
    ids_stage = [torch.empty(n_occ, dtype=torch.int64, device=dev) for _ in range(2)]
    status_host = [torch.empty(S, dtype=torch.int32).pin_memory() for _ in range(2)]
    status_dev = torch.empty(S, dtype=torch.int32, device=dev)
    ev_done = [torch.cuda.Event() for _ in range(2)]
    last = [None]
    copy_stream = torch.cuda.Stream(device=dev)
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    assert n_sets % 2 == 0, "the staging buffers alternate with the buffer sets"

    def e2e_body(k):
        wk.forward(ids_stage[k % 2], B, training=True, out=outs[0])
        wk.backward(grads[k], want_status=True, status=status_dev)
        status_host[k % 2].copy_(status_dev, non_blocking=True)

    staged = {}

    def e2e_copy(k):
        b = k % 2
        if staged.get(b) == ("set", k):
            return
        staged[b] = ("set", k)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[b])
            ids_stage[b].copy_(ids_pinned[k], non_blocking=True)
            ev_copied[b].record(copy_stream)

    e2e_graphs = None
    with torch.cuda.stream(gstream):
        for b in range(2):
            ids_stage[b].copy_(ids_pinned[b])
        e2e_body(0)
        gstream.synchronize()
        dist.barrier()
        if not args.no_graph:
            e2e_graphs = []
            for k in range(n_sets):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=gstream, capture_error_mode="thread_local"):
                    e2e_body(k)
                e2e_graphs.append(gph)
        for b in range(2):
            ev_free[b].record(gstream)
    torch.cuda.synchronize()
    dist.barrier()
    e2e_copy(0)

    def e2e_step(k):
        e2e_copy(k)                 # (already under way, except for the first step of a loop)
        with torch.cuda.stream(gstream):
            gstream.wait_event(ev_copied[k % 2])
            if e2e_graphs is not None:
                e2e_graphs[k].replay()
            else:
                e2e_body(k)
            ev_free[k % 2].record(gstream)
            ev_done[k % 2].record(gstream)
        staged.pop(k % 2, None)
        e2e_copy((k + 1) % n_sets)  # the next step's ids travel while this step computes
        if last[0] is not None:     # the host reads every step's status, one step behind the launches
            ev_done[last[0]].synchronize()
            assert int(status_host[last[0]][0]) >= -1
        last[0] = k % 2

    for i in range(Wm):
        e2e_step(i % n_sets)
    ms_e2e = timed(e2e_step, K)
    admitted1 = len(wk.shard)
    counters = wk.shard.counters()

    overflowed, gave_up = wk.status()
    assert not overflowed, "a (source, owner) pair needed more than cap slots: raise the margin of calibrate_cap"
    assert not gave_up and counters["wait_errors"] == 0, "a wait for a peer gave up: the run is void"

    # ---- where a step's time goes on this rank: CUDA events around every launch of ONE kernel family at a time (eager
    # launches; the backward's kernels then run one after another).  Family 8 is the time spent waiting for peers' flags.
    fam_names = {0: "k_owner_lookup", 1: "k_dedup", 2: "k_expand", 3: "k_nan_scan", 4: "k_reduce_hot<send>", 5: "k_reduce_cold<send>",
                 7: "k_reduce_warm<send>", 8: "k_wait (peers)", 9: "k_route_items + k_signal", 10: "k_owner_update_all"}
    kern = {}
    if not args.no_kernel_table:
        import ctypes as C

        n_prof = 10
        for f in sorted(fam_names):
            dist.barrier()
            lib.pb_profile_enable(1 << f)
            with torch.cuda.stream(gstream):
                for i in range(n_prof):
                    eager_step(i % n_sets)
            fam_ms = (C.c_double * 11)()
            fam_cnt = (C.c_uint64 * 11)()
            N.check(lib.pb_profile_read(fam_ms, fam_cnt, 11))
            lib.pb_profile_enable(0)
            if fam_cnt[f]:
                kern[fam_names[f]] = {"us_per_step": round(1e3 * fam_ms[f] / n_prof, 1), "launches_per_step": fam_cnt[f] / n_prof}
        torch.cuda.synchronize()
        dist.barrier()

    parity = None
    if not args.no_parity:
        sets = (0, 1)
        all_ids = [W.make_batches(100 + q, card, B, n_sets, args.alpha) if q != rank else ids_host for q in range(world)]
        all_grads = []
        for q in range(world):
            gq = torch.Generator(device=dev)
            gq.manual_seed(5 + q)
            full = (torch.randn((n_sets, S, B, dim), generator=gq, device=dev) * 1e-2).half()
            all_grads.append({k: full[k].cpu().numpy() for k in sets})
            del full
        replay = step
        if bounded:
            # a sweep between seeding the oracle and the check would re-admit a seeded sign from scratch on the GPU only:
            # the check replays graphs captured with the sweep switched off (same kernels otherwise)
            wk.shard.set_eviction(check_every=0)
            chk = {}
            with torch.cuda.stream(gstream):
                for k in sets:
                    gph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gph, stream=gstream, capture_error_mode="thread_local"):
                        eager_step(k)
                    chk[k] = gph
            torch.cuda.synchronize()
            dist.barrier()

            def replay(k):
                with torch.cuda.stream(gstream):
                    chk[k].replay()
        bad_out, bad_rows, n_rows = _check_parity(rank, world, wk, replay, outs, all_ids, all_grads, pf, S, B, dim, dev, sets)
        res = torch.tensor([bad_out, bad_rows, n_rows], dtype=torch.int64, device=dev)
        dist.all_reduce(res)
        if int(res[0]) or int(res[1]):
            raise AssertionError(f"parity: {int(res[0])} output rows and {int(res[1])} table rows differ from the oracle")
        parity = {"checked": True, "steps_replayed": len(sets), "rows_compared": int(res[2]),
                  "what": ("CUDA graphs of the same step captured with the capacity sweep switched off, " if bounded else "the timed CUDA graphs ") +
                          "replayed from the live tables on every rank; each rank's shard and the "
                          "outputs it serves are bit-identical to the oracle (R parameter servers, the R requests of a "
                          "step applied in rank order)"}

    line = None
    if rank == 0:
        ms_per_step = ms / K
        GB = B * world
        peak, peak_src = B_.peak_hbm()
        by = B_.kernel_bytes(stats, dim, dim)
        whole = by["whole_step"] / (ms_per_step * 1e-3) / 1e9  # per GPU, measured multiplicities of rank 0's batch
        line = {
            "metric": B_.METRIC, "value": GB / (ms_per_step * 1e-3), "unit": B_.UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": B_.workload_config(args),
            "run": {"ms_per_step_repetitions": reps, "resident_rows_rank0": resident, "table_fill_seconds": round(t_fill, 2),
                    "l2": "inputs larger than L2: %.1f GB table per GPU + %d rotating id/grad/output sets" % (
                        resident * 4.0 * 2 * dim / 1e9, n_sets),
                    "kernels_rank0": kern,
                    "exchange": "slots per (source, owner) pair: %d (calibrated on the batches, +15%%); distinct signs per "
                                "batch on rank 0: %d of %d occurrences" % (cap, stats["items"], stats["occurrences"]),
                    "launch": "whole step (compute kernels, peer stores, flag waits) = one CUDA graph per rank" if graphs is not None else "kernel by kernel",
                    "eviction": None if not bounded else {
                        "policy": "sweep when free rows < low water; rows of the last 2 batches are protected",
                        "resident_rows_rank0_before_after_e2e_loop": [int(admitted0), int(admitted1)],
                        "index_miss_count_rank0": int(counters["lookup_miss"]), "capacity_refused_rank0": int(counters["capacity_refused"])}},
            "clocks": clocks,
            "e2e": {"value": GB / (ms_e2e / K * 1e-3), "unit": B_.UNIT, "h2d_bytes_per_step": n_occ * 8 * world,
                    "d2h_bytes_per_step": S * 4 * world, "ms_per_step": ms_e2e / K,
                    "path": "pinned host ids -> H2D (copy stream, one step ahead) -> pb_forward_sharded -> pb_backward_sharded -> D2H slot status (one CUDA "
                            "graph per rank), host sync every step"},
            "gpu_launches": launches_per_step * K,
            "parity_checked": bool(parity and parity["checked"]), "parity": parity,
            "roofline": {"bound": "hbm", "kernel": "whole step per GPU (the per-kernel roofline is the N = 1 run's)",
                         "achieved": whole, "peak": peak, "unit": "GB/s", "frac": whole / peak, "traffic": None,
                         "peak_source": peak_src, "batch_stats_rank0": stats,
                         "bytes_model": "measured multiplicities of rank 0's batch (see bench.py kernel_bytes); NVLink bytes not counted"},
            "cpu_baseline": None,
        }
    # ---- the measured line is complete; what follows (teardown of the 100 GB tables, then the TrainCtx / DDP leg) can
    # only add to it: a watchdog prints the line as it stands and ends the process if any of it stalls
    printed = threading.Event()

    def emit(extra):
        if printed.is_set():
            return
        printed.set()
        if rank == 0:
            line["e2e_model"] = extra
            print(json.dumps(line))
            sys.stdout.flush()

    def line_watchdog():
        if not printed.wait(300):
            sys.stderr.write("[bench] teardown / e2e_model leg stalled: reporting without it\n")
            sys.stderr.flush()
            emit({"error": "the e2e_model leg (or the teardown before it) did not finish within 300 s"})
            os._exit(0)

    threading.Thread(target=line_watchdog, daemon=True).start()
    # orderly teardown of the timed path: graphs first (they reference the exchange areas), then the worker
    dist.barrier()
    torch.cuda.synchronize()
    del graphs, e2e_graphs
    torch.cuda.synchronize()
    wk.close()
    del wk
    torch.cuda.empty_cache()
    dist.barrier()
    if args.no_model_leg:
        emit(None)
    else:
        try:
            emit(model_leg_dist(args, rank, world, local_rank, B_))
        except Exception as e:  # noqa: BLE001 — reported, never fatal for the measured line
            emit({"error": repr(e)[:300]})
    done = threading.Event()

    def watchdog():
        if not done.wait(60):
            sys.stderr.write("[bench] teardown stalled: exiting\n")
            sys.stderr.flush()
            os._exit(0)

    threading.Thread(target=watchdog, daemon=True).start()
    dist.barrier()
    dist.destroy_process_group()
    done.set()
