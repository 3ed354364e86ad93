"""GPU parity of raw (embedding_summation: false) slots against the oracle's restatement of
embedding_worker_service/mod.rs:498-512, :540-545, :593-623, :790-798 and persia-core forward.rs:336-347.

Distinct signs are numbered by first occurrence on both sides (the reference: hashbrown order, unpinned),
so table, index, non_empty_index and sample_id_num compare bit for bit; so do the rows after the update.
"""
import numpy as np
import pytest

from util import to_dev_i32, to_dev_ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    return torch


@pytest.fixture(scope="module")
def pb(torch_cuda):
    from persia_b200 import shard

    return shard


def _pair(pb, oracle, dim, kind, fixed, group=0, prefix_bit=8, optim_kw=None, cap=1 << 16, max_occ=1 << 16):
    optim_kw = optim_kw or {}
    pf = oracle.index_prefix(group, prefix_bit) if group is not None else 0
    s = pb.EmbeddingShard(dim, cap, 0)
    s.set_optimizer(kind, **{{"mom": "g_square_momentum", "init_acc": "initialization", "b1": "beta1",
                              "b2": "beta2"}.get(k, k): v for k, v in optim_kw.items()})
    s.configure()
    ctx = pb.BatchContext(max_occ, max_occ, [pf], prefix_bit=prefix_bit)
    w = oracle.Worker([oracle.SlotCfg(dim, summation=False, sample_fixed_size=fixed, prefix=pf)], n_ps=1,
                      prefix_bit=prefix_bit)
    w.configure()
    w.set_optimizer(oracle.Optim(kind, **optim_kw))
    return s, ctx, w


def _ragged(rng, B, card, max_ids, allow_empty=True):
    counts = rng.integers(0 if allow_empty else 1, max_ids + 1, size=B)
    row_off = np.zeros(B + 1, np.uint32)
    row_off[1:] = np.cumsum(counts)
    ids = rng.integers(0, card, size=int(row_off[-1]), dtype=np.uint64)
    return ids, row_off


def _fwd_both(torch, s, ctx, w, ids, row_off, B, fixed, training=True, single=False):
    ro_dev = None if single else to_dev_i32(row_off, DEV)
    table, index, non_empty, num, counts = ctx.forward_raw(s, to_dev_ids(ids, DEV), B, fixed, row_off=ro_dev,
                                                           training=training)
    U, ne = counts.cpu().numpy().tolist()
    wt, wi, wne, wnum, octx = w.forward_raw(0, ids, row_off, B, training=training)
    assert U == wt.shape[0] - 1
    got_t = table[:U + 1].cpu().numpy()
    assert got_t.tobytes() == wt.tobytes()
    np.testing.assert_array_equal(index.cpu().numpy(), wi)
    assert ne == wne.size
    np.testing.assert_array_equal(non_empty[:ne].cpu().numpy(), wne)
    np.testing.assert_array_equal(num.cpu().numpy().view(np.uint32), wnum)
    return U, octx


def _entries_equal(s, w, signs):
    ent, found = s.get_entries(to_dev_ids(signs, DEV))
    ent = ent.cpu().numpy()
    assert found.all()
    for k, sign in enumerate(signs):
        ref = w.get_entry(int(sign))
        assert ref is not None and ent[k].tobytes() == ref.tobytes(), (k, sign, ent[k], ref)


@pytest.mark.parametrize("dim,fixed,max_ids", [(16, 4, 7), (64, 10, 3), (12, 1, 5)])
def test_raw_forward_ragged_matches_oracle(torch_cuda, pb, oracle, dim, fixed, max_ids):
    torch = torch_cuda
    s, ctx, w = _pair(pb, oracle, dim, oracle.SGD, fixed, optim_kw={"lr": 0.05})
    rng = np.random.default_rng(dim + fixed)
    for B, card in ((257, 50), (1000, 100000), (33, 5)):
        ids, row_off = _ragged(rng, B, card, max_ids)
        _fwd_both(torch, s, ctx, w, ids, row_off, B, fixed, training=True)
    # inference: signs never seen read as zeros but still take a number (lookup miss, mod.rs:308-342)
    ids, row_off = _ragged(rng, 64, 10**9, max_ids)
    _fwd_both(torch, s, ctx, w, ids, row_off, 64, fixed, training=False)
    # empty batch rows only / nothing at all
    ids, row_off = np.zeros(0, np.uint64), np.zeros(9, np.uint32)
    assert _fwd_both(torch, s, ctx, w, ids, row_off, 8, fixed, training=True)[0] == 0


@pytest.mark.parametrize("kind_name,dim,kw", [("SGD", 16, {"lr": 0.1, "wd": 0.01}),
                                               ("ADAGRAD", 64, {"lr": 0.02, "init_acc": 0.01}),
                                               ("ADAGRAD_VW", 24, {"lr": 0.02}),
                                               ("ADAM", 32, {"lr": 0.001})])
def test_raw_training_bit_exact(torch_cuda, pb, oracle, kind_name, dim, kw):
    torch = torch_cuda
    kind = getattr(oracle, kind_name)
    oracle.set_rsqrt_exact(True)
    try:
        fixed = 5
        s, ctx, w = _pair(pb, oracle, dim, kind, fixed, optim_kw=kw)
        rng = np.random.default_rng(5)
        seen = set()
        for it in range(6):
            B = 300
            ids, row_off = _ragged(rng, B, 400, 8)
            U, octx = _fwd_both(torch, s, ctx, w, ids, row_off, B, fixed, training=True)
            seen.update(oracle.add_prefix(ids, 8, oracle.index_prefix(0, 8)).tolist())
            f16 = it % 2 == 1
            g = (rng.standard_normal((U, dim)) * 1e-2).astype(np.float16 if f16 else np.float32)
            scale = 128.0 if it >= 2 else 1.0
            if f16 and U:
                g[0, 0] = np.inf  # clamps to 65504 (persia-common lib.rs:163-180)
            st = ctx.backward_raw(s, torch.from_numpy(g).to(DEV), scale=scale, want_status=True).item()
            assert st == w.backward_raw(0, octx, g, scale=scale) == 0
        _entries_equal(s, w, np.array(sorted(seen), np.uint64))
    finally:
        oracle.set_rsqrt_exact(False)


def test_raw_single_id_layout_nan_and_skip(torch_cuda, pb, oracle):
    torch = torch_cuda
    dim, fixed, B = 16, 2, 512
    s, ctx, w = _pair(pb, oracle, dim, oracle.SGD, fixed, optim_kw={"lr": 0.1})
    rng = np.random.default_rng(9)
    seen = set()
    for it in range(4):
        ids = rng.integers(0, 300, size=B, dtype=np.uint64)
        row_off = np.arange(B + 1, dtype=np.uint32)
        U, octx = _fwd_both(torch, s, ctx, w, ids, row_off, B, fixed, training=True, single=True)
        seen.update(oracle.add_prefix(ids, 8, oracle.index_prefix(0, 8)).tolist())
        g = (rng.standard_normal((U, dim)) * 1e-2).astype(np.float32)
        if it == 1:
            g[U // 2, 3] = np.nan  # the whole gradient is dropped (mod.rs:731-746)
        if it == 2:
            st = ctx.backward_raw(s, None, want_status=True).item()
            assert st == w.backward_raw(0, octx, None, skip=True) == 1
            continue
        st = ctx.backward_raw(s, torch.from_numpy(g).to(DEV), want_status=True).item()
        assert st == w.backward_raw(0, octx, g) == (2 if it == 1 else 0)
    _entries_equal(s, w, np.array(sorted(seen), np.uint64))
    # a second backward without a forward is the reference's "backward_ref_id not found"
    from persia_b200.native import PersiaB200Error
    with pytest.raises(PersiaB200Error):
        ctx.backward_raw(s, torch.zeros((1, dim), device=DEV))


def test_raw_no_prefix_marker_signs(torch_cuda, pb, oracle):
    """index_prefix 0 (indices_add_prefix is a no-op, mod.rs:411): ids pass through unchanged, including the three
    values that collide with the index's cell markers and the scratch set's empty key."""
    torch = torch_cuda
    dim, fixed = 8, 3
    s, ctx, w = _pair(pb, oracle, dim, oracle.SGD, fixed, group=None, optim_kw={"lr": 0.5})
    M = 2**64
    ids = np.array([M - 1, 7, M - 2, M - 1, M - 3, 7, M - 3, 0], np.uint64)
    row_off = np.array([0, 3, 3, 7, 8], np.uint32)
    for it in range(2):
        U, octx = _fwd_both(torch, s, ctx, w, ids, row_off, 4, fixed, training=True)
        assert U == 5
        g = np.full((U, dim), 0.25 * (it + 1), np.float32)
        assert ctx.backward_raw(s, torch.from_numpy(g).to(DEV), want_status=True).item() == 0
        assert w.backward_raw(0, octx, g) == 0
    _entries_equal(s, w, np.unique(ids))
