"""CPU, world_size 2, gloo: the exchange bookkeeping of ShardedEmbeddingWorker (partition order, split sizes,
permutations, reassembly) with an oracle-backed shard standing in for the GPU.  The result must equal the
oracle's own embedding worker with R = 2 parameter servers fed the concatenated (global) batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

S, B, DIM, R = 3, 64, 8, 2
CARD = [5, 200, 100000]
STEPS = 3


class OracleBackend:
    """Test double of worker.CudaBackend: numpy for the id plumbing, the oracle's parameter server for rows."""

    NULL = np.uint64(0xFFFFFFFFFFFFFFFE)  # PB_NULL_SIGN: padding of a framed exchange
    device = torch.device("cpu")

    def __init__(self, oracle, dim, optim, rank):
        self.o, self.dim = oracle, dim
        self.w = oracle.Worker([oracle.SlotCfg(dim)], n_ps=1)  # this rank's PS, addressed directly
        self.w.configure()
        self.w.set_optimizer(optim)
        self.last = None

    def add_prefix(self, ids, slot_occ_off, prefixes, prefix_bit):
        x = ids.numpy().view(np.uint64)
        out = np.concatenate([self.o.add_prefix(x[slot_occ_off[i]:slot_occ_off[i + 1]], prefix_bit, prefixes[i])
                              for i in range(len(prefixes))])
        return torch.from_numpy(out.view(np.int64))

    def partition(self, signs, Rn):
        sh = self.o.shard_of(signs.numpy().view(np.uint64), Rn)
        perm = np.argsort(sh, kind="stable").astype(np.int32)
        return torch.from_numpy(perm), torch.from_numpy(np.bincount(sh, minlength=Rn).astype(np.int32))

    def take(self, src, perm):
        return src[perm.long()]

    def take_rows(self, src, perm):
        return src[perm.long()].contiguous()

    def put_rows(self, src, perm):
        out = torch.empty_like(src)
        out[perm.long()] = src
        return out

    def serve_lookup(self, signs, training):
        s = signs.numpy().view(np.uint64)
        self.last = s
        real = s != self.NULL  # an owner-mode context skips the padding: no lookup, zero rows
        rows = np.zeros((s.size, self.dim), np.float32)
        rows[real] = self.w.ps_lookup(0, s[real], np.full(int(real.sum()), self.dim, np.uint32), training).reshape(-1, self.dim)
        return torch.from_numpy(self.o.f32_to_f16(rows).view(np.float16).reshape(-1, self.dim))

    # fixed-capacity framing (pb_frame_signs / pb_frame_rows restated in numpy)
    def empty_rows(self, n, dtype=torch.float16):
        return torch.zeros((n, self.dim), dtype=dtype)

    def frame_signs(self, signs, perm, counts, Rn, cap, overflow):
        x, pm, ct = signs.numpy().view(np.uint64), perm.numpy(), counts.numpy()
        out = np.full(Rn * cap, self.NULL, np.uint64)
        off = 0
        for r in range(Rn):
            k = min(int(ct[r]), cap)
            out[r * cap:r * cap + k] = x[pm[off:off + k]]
            if ct[r] > cap:
                overflow[0] = 1
            off += int(ct[r])
        return torch.from_numpy(out.view(np.int64))

    def frame_rows(self, src, perm, counts, Rn, cap, pack, out):
        pm, ct = perm.numpy().astype(np.int64), counts.numpy()
        off = 0
        out.zero_()
        for r in range(Rn):
            k = min(int(ct[r]), cap)
            idx = torch.from_numpy(pm[off:off + k])
            if pack:
                out[r * cap:r * cap + k] = src[idx]
            else:
                out[idx] = src[r * cap:r * cap + k]
            off += int(ct[r])
        return out

    def serve_update(self, grads, scale):
        # what the owner-side context does: one segment per sign, gradients summed in arrival order
        g = grads.numpy().astype(np.float32)
        if abs(scale - 1.0) > 1.1920929e-07:
            g = g * np.float32(1.0 / scale)
        uniq, first = {}, []
        for k, s in enumerate(self.last.tolist()):
            if s == int(self.NULL):
                continue
            if s not in uniq:
                uniq[s] = np.zeros(self.dim, np.float32)
                first.append(s)
            uniq[s] = uniq[s] + g[k]
        signs = np.array(first, np.uint64)
        self.w.ps_update(0, signs, np.full(signs.size, self.dim, np.uint32), np.stack([uniq[s] for s in first]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batches(R=R):
    rng = np.random.default_rng(77)
    out = []
    for _ in range(STEPS):
        ids = np.stack([np.stack([rng.integers(0, CARD[s], size=B, dtype=np.uint64) for s in range(S)]) for _ in range(R)])
        g = (rng.standard_normal((R, S, B, DIM)) * 1e-2).astype(np.float16)
        out.append((ids, g))
    return out


def _run(rank, port, q, static=False, R=R):
    import oracle
    from persia_b200.worker import ShardedEmbeddingWorker

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=R)
    try:
        pf = [oracle.index_prefix(i) for i in range(S)]
        be = OracleBackend(oracle, DIM, oracle.Optim(oracle.SGD, lr=0.1, wd=0.0), rank)
        wk = ShardedEmbeddingWorker(S, DIM, pf, be)
        if static:  # fixed-capacity frames, capacity calibrated on the batches (a collective)
            sample = [torch.from_numpy(ids[rank].reshape(-1).view(np.int64)) for ids, _ in _batches(R)]
            cap = wk.calibrate_cap(sample, B, margin=1.05, extra=2)
            assert cap < S * B
            wk.enable_static(B, cap=cap)
        fwd, bwd = (wk.forward_static, wk.backward_static) if static else (wk.forward, wk.backward)
        outs = []
        for ids, g in _batches(R):
            out = fwd(torch.from_numpy(ids[rank].reshape(-1).view(np.int64)), B, training=True)
            outs.append(out.numpy().copy())
            assert bwd(torch.from_numpy(g[rank]), scale=1.0)
        if static:
            assert not wk.check_overflow()
        # dump this rank's shard
        probe = np.concatenate([oracle.add_prefix(np.arange(min(c, 400), dtype=np.uint64), 8, pf[i]) for i, c in enumerate(CARD)])
        mine = probe[oracle.shard_of(probe, R) == rank]
        ent = {int(s): be.w.get_entry(int(s)) for s in mine}
        q.put((rank, outs, {k: v for k, v in ent.items() if v is not None}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("static,R", [(False, 2), (True, 2), (True, 3)])
def test_sharded_worker_equals_global_batch_oracle(static, R):
    import oracle

    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_run, args=(r, port, q, static, R)) for r in range(R)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(R):
        rank, outs, ent = q.get(timeout=120)
        res[rank] = (outs, ent)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # the reference: one embedding worker, R parameter servers, the concatenated batch
    pf = [oracle.index_prefix(i) for i in range(S)]
    w = oracle.Worker([oracle.SlotCfg(DIM, prefix=p) for p in pf], n_ps=R)
    w.configure()
    w.set_optimizer(oracle.Optim(oracle.SGD, lr=0.1, wd=0.0))
    GB = R * B
    for step, (ids, g) in enumerate(_batches(R)):
        gid = np.concatenate([ids[:, s, :].reshape(-1) for s in range(S)])  # slot-major, rank-major samples
        want, octx = w.forward(gid, np.arange(S * GB + 1, dtype=np.uint32), GB, training=True)
        for r in range(R):
            got = res[r][0][step]
            for s in range(S):
                np.testing.assert_array_equal(got[s].view(np.uint16), want[s][r * B:(r + 1) * B].view(np.uint16))
        gg = [np.concatenate([g[r, s] for r in range(R)]) for s in range(S)]
        w.backward(octx, gg)
    n = 0
    for r in range(R):
        for sign, e in res[r][1].items():
            ref = w.get_entry(sign)
            assert ref is not None and e.tobytes() == ref.tobytes()
            n += 1
    assert n > 100
