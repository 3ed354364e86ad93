#!/usr/bin/env python
"""bench.py — samples/s of PERSIA's sparse-embedding hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic Criteo-shaped ids: training forward
(prefix -> find-or-admit -> gather+pool -> f16) and backward (NaN scan -> group -> reduce -> Adagrad update).
N=1 workload = BASELINE configs[1]: 26 slots, 1e8 resident rows, dim 64, batch 4096, Adagrad.

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                     the reference's CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from persia_b200 import workload as W  # noqa: E402

METRIC = "samples/sec (Criteo-1TB-shape DLRM sparse path, 26 slots)"
UNIT = "samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=float, default=float(os.environ.get("PB_BENCH_ROWS", 1e8)),
                    help="resident rows per GPU (weak scaling: the table grows with N)")
    ap.add_argument("--batch", type=int, default=4096, help="samples per GPU per step")
    ap.add_argument("--dim", type=int, default=None, help="default 64 (configs[1]) at every N so that the scaling runs compare like with like; --dim 128 gives configs[2..3]")
    ap.add_argument("--slots", type=int, default=26)
    ap.add_argument("--alpha", type=float, default=1.05)
    ap.add_argument("--sets", type=int, default=16, help="rotating input/grad/output buffer sets (> L2 in total)")
    ap.add_argument("--cpu-seconds", type=float, default=float(os.environ.get("PB_BENCH_CPU_SECONDS", 12)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-grouping", action="store_true", help="run the backward's grouping inside pb_backward")
    ap.add_argument("--dist-graph", action="store_true", help="multi-GPU: capture the framed step incl. NCCL in CUDA graphs")
    ap.add_argument("--dist-nccl", action="store_true", help="multi-GPU: NCCL all-to-all instead of the peer-memory exchange")
    ap.add_argument("--dist-dynamic", action="store_true", help="multi-GPU: split-size all-to-all (host sync per step)")
    ap.add_argument("--equal-card", action="store_true", help="diagnostic: every slot gets rows/slots ids (no tiny slots)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel instead of replaying CUDA graphs")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu_index = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.2] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port): timed on the host cores
# ------------------------------------------------------------------------------------------------------
class CpuArm:
    """The reference's CPU path (oracle port of its EW + PS, in process) on the same workload: a pool of distinct
    batches, admitted once untimed (steady state is a warm table; the GPU arm is timed on a resident table too)."""

    def __init__(self, args, dim, card, n_threads=None):
        import oracle

        self.S, self.B, self.dim = args.slots, args.batch, dim
        self.n_threads = n_threads or os.cpu_count() or 1
        pf = W.index_prefixes(self.S)
        self.w = oracle.Worker([oracle.SlotCfg(dim, prefix=p) for p in pf], n_ps=1, capacity_per_ps=1 << 40,
                               n_internal_shards=max(64, 8 * self.n_threads))
        self.w.configure()
        self.w.set_optimizer(oracle.Optim(oracle.ADAGRAD, lr=0.01, init_acc=0.01, eps=1e-10))
        self.row_off = np.arange(self.S * self.B + 1, dtype=np.uint32)
        rng = np.random.default_rng(123)
        self.g = [(rng.standard_normal((self.B, dim)) * 1e-2).astype(np.float16) for _ in range(self.S)]
        self.n_pool = int(min(256, max(4 * self.n_threads, 16)))
        self.ids = W.make_batches(1001, card, self.B, self.n_pool, args.alpha)
        self.warm_seconds = self.run(self.ids)
        self.cursor = 0

    def run(self, ids):
        return self.w.bench(ids, self.row_off, self.B, self.g, self.n_threads)

    def step(self, n_batches):
        """fwd+bwd of the next n_batches of the pool (wraps around); returns seconds."""
        idx = [(self.cursor + i) % self.n_pool for i in range(n_batches)]
        self.cursor = (self.cursor + n_batches) % self.n_pool
        return self.run(np.ascontiguousarray(self.ids[idx]))

    def describe(self, n_batches):
        return (f"{n_batches} batches ({self.n_pool} distinct, table warmed by one untimed pass) x {self.B} samples x "
                f"{self.S} slots dim {self.dim}, Adagrad, fwd+bwd, {self.n_threads} threads each owning whole batches "
                f"(in-process EW+PS, no RPC/codec/H2D)")


def cpu_arm(args, dim, seconds, card, n_threads=None):
    arm = CpuArm(args, dim, card, n_threads)
    t = arm.step(arm.n_pool)
    reps = int(max(1, min(64, seconds / max(t, 1e-3))))
    n_batches, tot = 0, 0.0
    for _ in range(reps):
        tot += arm.step(arm.n_pool)
        n_batches += arm.n_pool
    return {"value": n_batches * arm.B / tot, "unit": UNIT, "cores": arm.n_threads, "kind": "port",
            "sample": arm.describe(n_batches), "seconds": tot}


def reference_main(args):
    """--impl reference: the reference's own CPU implementation of the path (no Rust toolchain here, so the
    oracle port) with all host threads.  A step = two batches per host thread, fwd+bwd."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dim = args.dim or 64
    card = W.scaled_cardinalities(int(args.rows) * args.gpus, args.slots)
    arm = CpuArm(args, dim, card)
    per_step = 2 * arm.n_threads
    for _ in range(max(args.warmup, 1)):
        arm.step(per_step)
    K = max(1, args.steps)
    t0 = time.time()
    tot = 0.0
    done = 0
    for _ in range(K):
        tot += arm.step(per_step)
        done += 1
        if time.time() - t0 > 150:  # keep the whole run within a few minutes
            break
    v = done * per_step * arm.B / tot
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, dim, card, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": arm.n_threads, "kind": "port",
                         "sample": arm.describe(done * per_step)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, dim, card, n_gpus):
    return {
        "workload": f"configs[{1 if dim == 64 else 2}] shape x {n_gpus} GPU: {args.slots} Criteo-shaped slots, "
                    f"{int(args.rows) * n_gpus:.3g} rows, dim {dim}, batch {args.batch}/GPU, Adagrad, training forward + backward",
        "global_batch": args.batch * n_gpus, "slots": args.slots, "dim": dim, "rows_total": int(args.rows) * n_gpus,
        "zipf_alpha": args.alpha, "optimizer": "adagrad(lr=0.01, init=0.01, eps=1e-10)",
        "parallelism": "single shard" if n_gpus == 1 else f"rows hash-sharded over {n_gpus} GPUs (farmhash64 % {n_gpus}), "
                                                          f"data-parallel batches, all-to-all over NCCL",
        "cardinalities": [int(c) for c in card],
    }


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def b200_main(args):
    import torch

    from persia_b200 import native as N

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback; use --impl reference for the CPU arm)"
    if world != args.gpus:
        assert world == 1 and args.gpus == 1, f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})"
    torch.cuda.set_device(local_rank)
    lib = N.load()
    if world > 1:
        from persia_b200 import dist_bench

        return dist_bench.run(args, rank, local_rank, world, METRIC, UNIT, workload_config, ClockSampler, cpu_arm)
    return single_gpu(args, torch, lib)


def single_gpu(args, torch, lib):
    import ctypes as C

    from persia_b200 import native as N
    from persia_b200 import shard as SH

    dev = torch.device("cuda", torch.cuda.current_device())
    dim = args.dim or 64
    S, B, K, Wm = args.slots, args.batch, args.steps, max(args.warmup, 3)
    rows = int(args.rows)
    card = W.scaled_cardinalities(rows, S)
    if args.equal_card:
        card = np.full(S, rows // S, np.int64)
        card[0] += rows - int(card.sum())
    pf = W.index_prefixes(S)
    slot_off = [s * B for s in range(S + 1)]
    n_occ = S * B

    sh = SH.EmbeddingShard(dim, rows + 1024, dev)
    sh.set_optimizer(N.OPT_ADAGRAD, lr=0.01, initialization=0.01, eps=1e-10)
    sh.configure()
    ctx = SH.BatchContext(n_occ, n_occ, pf, device=dev)

    # ---- make every row resident (the reference's "warm table"): admit all ids of every slot
    t_fill = time.time()
    chunk = 1 << 21
    buf = torch.empty((chunk, dim), dtype=torch.float32, device=dev)
    for s in range(S):
        for lo in range(0, int(card[s]), chunk):
            hi = min(int(card[s]), lo + chunk)
            ids = torch.arange(lo, hi, dtype=torch.int64, device=dev)
            signs = SH.add_prefix(ids, [0, hi - lo], [pf[s]])
            sh.lookup(signs, training=True, out=buf[: hi - lo])
    torch.cuda.synchronize()
    resident = len(sh)
    t_fill = time.time() - t_fill
    del buf
    assert resident == rows, (resident, rows)

    # ---- rotating buffer sets: ids, gradients, outputs (together > L2) ; pinned host ids for e2e
    n_sets = max(2, args.sets)
    ids_host = W.make_batches(2, card, B, n_sets, args.alpha)
    ids_pinned = torch.from_numpy(ids_host.view(np.int64)).pin_memory()
    ids_dev = [ids_pinned[k].to(dev) for k in range(n_sets)]
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    grads_all = (torch.randn((n_sets, S, B, dim), generator=g, device=dev) * 1e-2).half()
    grads = [[grads_all[k, s] for s in range(S)] for k in range(n_sets)]
    outs = [torch.empty((S, B, dim), dtype=torch.float16, device=dev) for _ in range(n_sets)]
    uniq = float(np.mean([np.unique(ids_host[k].reshape(S, B) + (np.arange(S, dtype=np.uint64) << np.uint64(56))[:, None]).size
                          for k in range(min(4, n_sets))])) / n_occ

    def step(k):
        ctx.forward(sh, ids_dev[k], slot_off, B, training=True, out=outs[k])
        ctx.backward(sh, grads[k])

    stream = torch.cuda.Stream(device=dev)
    use_graph = not args.no_graph
    with torch.cuda.stream(stream):
        for i in range(3):
            step(i % n_sets)
        stream.synchronize()
        l0 = lib.pb_launch_count()
        step(0)
        launches_per_step = int(lib.pb_launch_count() - l0)
        stream.synchronize()
        graphs = None
        if use_graph:
            graphs = []
            for k in range(n_sets):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=stream):
                    step(k)
                graphs.append(gph)

        def run(i):
            if graphs is not None:
                graphs[i % n_sets].replay()
            else:
                step(i % n_sets)

        for i in range(Wm):
            run(i)
        stream.synchronize()
        sampler = ClockSampler(dev.index or 0)
        sampler.start()
        time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.time()
        e0.record(stream)
        for i in range(K):
            run(i)
        e1.record(stream)
        torch.cuda.synchronize()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t0, t1)

        # ---- per-kernel durations (CUDA events on the launching stream), separate instrumented pass
        fam_names = ["probe_items", "dedup", "gather_pool", "nan_scan", "reduce_hot", "reduce_items", "other"]
        n_prof = min(K, 50)

        def profiled(mask):
            lib.pb_profile_enable(mask)
            for i in range(n_prof):
                step(i % n_sets)
            fam_ms = (C.c_double * 7)()
            fam_cnt = (C.c_uint64 * 7)()
            N.check(lib.pb_profile_read(fam_ms, fam_cnt, 7))
            lib.pb_profile_enable(0)
            return {fam_names[i]: {"us_per_step": 1e3 * fam_ms[i] / n_prof, "launches_per_step": fam_cnt[i] / n_prof}
                    for i in range(7) if fam_cnt[i]}

        # every launch bracketed: the host (two event records per launch) is slower than most of these kernels, so
        # small kernels read high — a table for orientation; the roofline kernel is then timed alone
        kern = profiled(0x7F)
        kern_alone = profiled(1 << 5)

        # ---- e2e: host ids (pinned) -> H2D -> forward -> backward -> D2H of the per-slot status, every step
        status_host = torch.empty(S, dtype=torch.int32).pin_memory()
        ids_stage = torch.empty(n_occ, dtype=torch.int64, device=dev)
        out_e2e = outs[0]

        def e2e_body(k):
            ids_stage.copy_(ids_pinned[k], non_blocking=True)
            ctx.forward(sh, ids_stage, slot_off, B, training=True, out=out_e2e)
            st = ctx.backward(sh, grads[k], want_status=True)
            status_host.copy_(st, non_blocking=True)

        e2e_graphs = None
        if use_graph:  # the whole call sequence incl. the H2D / D2H copies is one graph per pinned input buffer
            e2e_body(0)
            stream.synchronize()
            e2e_graphs = []
            for k in range(n_sets):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=stream):
                    e2e_body(k)
                e2e_graphs.append(gph)

        def e2e_step(k):
            if e2e_graphs is not None:
                e2e_graphs[k].replay()
            else:
                e2e_body(k)
            stream.synchronize()  # the caller reads the status: one host sync per step, as persia's backward does

        for i in range(Wm):
            e2e_step(i % n_sets)
        e0.record(stream)
        for i in range(K):
            e2e_step(i % n_sets)
        e1.record(stream)
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1)

    ms_per_step = ms / K
    value = B / (ms_per_step * 1e-3)
    state = dim  # Adagrad, elementwise
    bytes_per_id = W.algorithmic_bytes_per_id(dim, state, "total")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ku = kern_alone.get("reduce_items", {}).get("us_per_step")
    upd_bytes = n_occ * W.algorithmic_bytes_per_id(dim, state, "backward")
    achieved = upd_bytes / (ku * 1e-6) / 1e9 if ku else None
    traffic = None
    try:  # DRAM bytes per launch of the same kernel from the round's `ncu --set full` capture (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get("k_reduce_update")
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "k_reduce_update (A8+A9: gradient segment-reduce + Adagrad step + weight bound)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
        "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": upd_bytes,
        "whole_step": {"algorithmic_bytes": n_occ * bytes_per_id,
                       "achieved_gbs": n_occ * bytes_per_id / (ms_per_step * 1e-3) / 1e9,
                       "frac": n_occ * bytes_per_id / (ms_per_step * 1e-3) / 1e9 / peak},
        "kernels_us_per_step": kern, "kernel_timed_alone_us": ku,
    }
    cpu = None
    if not args.no_cpu_baseline:
        c = cpu_arm(args, dim, args.cpu_seconds, card)
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": Wm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, dim, card, 1), unique_id_fraction=uniq, resident_rows=resident,
                       table_fill_seconds=round(t_fill, 2),
                       l2="inputs larger than L2: %.1f GB table + %d rotating id/grad/output sets (%.0f MB)" % (
                           resident * 4.0 * (dim + state) / 1e9, n_sets, n_sets * 2 * n_occ * dim * 2 / 1e6),
                       launch=("CUDA graph replay, one graph per buffer set" if graphs is not None else "kernel by kernel") +
                              ("" if args.sync_grouping else "; grouping forked onto the context's side stream in pb_forward")),
        "clocks": clocks,
        "e2e": {"value": B / (ms_e2e / K * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_occ * 8,
                "d2h_bytes_per_step": S * 4, "ms_per_step": ms_e2e / K,
                "path": "pinned host ids -> H2D -> pb_forward -> pb_backward -> D2H slot status, host sync every step"},
        "gpu_launches": launches_per_step * K,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_main(args)
    else:
        b200_main(args)


if __name__ == "__main__":
    main()
