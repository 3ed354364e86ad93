"""GPU: ShardedEmbeddingWorker with the CUDA backend.  R = 1 in process (every GPU box), R = 2 over NCCL when
two GPUs are visible (gpurun --gpus 2): results equal the oracle's embedding worker with R parameter servers
fed the concatenated batch, bit for bit (SGD; strict reduce order on the owners)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu

S, B, DIM = 4, 512, 64
CARD = [3, 50, 3000, 200000]
STEPS = 3


def _batches(R):
    rng = np.random.default_rng(5)
    out = []
    for _ in range(STEPS):
        ids = np.stack([np.stack([rng.integers(0, CARD[s], size=B, dtype=np.uint64) for s in range(S)]) for _ in range(R)])
        g = (rng.standard_normal((R, S, B, DIM)) * 1e-2).astype(np.float16)
        out.append((ids, g))
    return out


def _reference(R):
    import oracle

    pf = [oracle.index_prefix(i) for i in range(S)]
    w = oracle.Worker([oracle.SlotCfg(DIM, prefix=p) for p in pf], n_ps=R)
    w.configure()
    w.set_optimizer(oracle.Optim(oracle.SGD, lr=0.05, wd=0.001))
    GB = R * B
    outs = []
    for ids, g in _batches(R):
        gid = np.concatenate([ids[:, s, :].reshape(-1) for s in range(S)])
        want, octx = w.forward(gid, np.arange(S * GB + 1, dtype=np.uint32), GB, training=True)
        outs.append(want)
        w.backward(octx, [np.concatenate([g[r, s] for r in range(R)]) for s in range(S)])
    return w, outs, pf


def _make_worker(torch, rank_dev):
    import oracle
    from persia_b200 import native as N
    from persia_b200.worker import CudaBackend, ShardedEmbeddingWorker

    pf = [oracle.index_prefix(i) for i in range(S)]
    be = CudaBackend(DIM, 1 << 18, rank_dev, dict(kind=N.OPT_SGD, lr=0.05, wd=0.001), {}, max_occurrences=1 << 16)
    be.ctx.set_strict_reduce(True)
    return ShardedEmbeddingWorker(S, DIM, pf, be), be, pf


def _check_shard(torch, be, w, pf, R, rank, dev):
    import oracle
    from util import to_dev_ids

    probe = np.concatenate([oracle.add_prefix(np.arange(min(c, 3000), dtype=np.uint64), 8, pf[i]) for i, c in enumerate(CARD)])
    mine = probe[oracle.shard_of(probe, R) == rank]
    ent, found = be.shard.get_entries(to_dev_ids(mine, dev))
    ent, found = ent.cpu().numpy(), found.cpu().numpy()
    n = 0
    for k, s in enumerate(mine):
        ref = w.get_entry(int(s))
        assert (ref is not None) == bool(found[k])
        if ref is not None:
            assert ent[k].tobytes() == ref.tobytes()
            n += 1
    return n


def test_worker_single_rank_matches_oracle():
    import torch

    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    wk, be, pf = _make_worker(torch, dev)
    w, outs, _ = _reference(1)
    for step, (ids, g) in enumerate(_batches(1)):
        out = wk.forward(torch.from_numpy(ids[0].reshape(-1).view(np.int64)).to(dev), B, training=True).cpu().numpy()
        for s in range(S):
            np.testing.assert_array_equal(out[s].view(np.uint16), outs[step][s].view(np.uint16))
        wk.backward(torch.from_numpy(g[0]).to(dev))
    assert _check_shard(torch, be, w, pf, 1, 0, dev) > 1000


@pytest.mark.parametrize("p2p", [False, True])
def test_worker_single_rank_static_frames(p2p):
    import torch

    dev = torch.device("cuda", 0)
    wk, be, pf = _make_worker(torch, dev)
    wk.enable_static(B)
    if p2p:
        wk.enable_p2p(B)
    w, outs, _ = _reference(1)
    for step, (ids, g) in enumerate(_batches(1)):
        out = (wk.forward_p2p if p2p else wk.forward_static)(torch.from_numpy(ids[0].reshape(-1).view(np.int64)).to(dev), B, training=True).cpu().numpy()
        for s in range(S):
            np.testing.assert_array_equal(out[s].view(np.uint16), outs[step][s].view(np.uint16))
        (wk.backward_p2p if p2p else wk.backward_static)(torch.from_numpy(g[0]).to(dev))
    assert not wk.check_overflow()
    assert _check_shard(torch, be, w, pf, 1, 0, dev) > 1000


def _rank_main(rank, R, port, q, static=False):
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=R, device_id=dev)
    try:
        wk, be, pf = _make_worker(torch, dev)
        w, outs, _ = _reference(R)
        if static == "p2p":  # capacity sized from the batches (collective), as bench.py does
            sample = [torch.from_numpy(ids[rank].reshape(-1).view(np.int64)).to(dev) for ids, _ in _batches(R)]
            cap = wk.calibrate_cap(sample, B)
            assert cap < S * B
            wk.enable_static(B, cap=cap)
        elif static:
            wk.enable_static(B)
        if static == "p2p":
            wk.enable_p2p(B)
        fwd = {False: wk.forward, True: wk.forward_static, "p2p": wk.forward_p2p}[static]
        bwd = {False: wk.backward, True: wk.backward_static, "p2p": wk.backward_p2p}[static]
        for step, (ids, g) in enumerate(_batches(R)):
            d_ids = torch.from_numpy(ids[rank].reshape(-1).view(np.int64)).to(dev)
            out = fwd(d_ids, B, training=True).cpu().numpy()
            for s in range(S):
                np.testing.assert_array_equal(out[s].view(np.uint16), outs[step][s][rank * B:(rank + 1) * B].view(np.uint16))
            bwd(torch.from_numpy(g[rank]).to(dev))
        if static:
            assert not wk.check_overflow()
        if static == "p2p":
            assert not wk.check_p2p()
        n = _check_shard(torch, be, w, pf, R, rank, dev)
        q.put((rank, n))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("static", [False, True, "p2p"])
def test_worker_two_ranks_nccl_matches_oracle(static):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q, static)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(got.values()) > 1000
