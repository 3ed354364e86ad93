"""CPU: properties of the oracle beyond the reference's golden vectors — two independent restatements agreeing
(C++ vs numpy farmhash, the LRU map vs an OrderedDict model, Adam vs a float64 rendering of adam_avx2), and
invariants of the reference's algorithm that hold for any input (dedup bookkeeping, hash-stack key ranges, results
independent of the number of parameter servers, raw vs summed views of the same rows)."""
from collections import OrderedDict

import numpy as np
import pytest


def test_farmhash_cxx_equals_numpy_restatement(oracle):
    from persia_b200.persia_core import farmhash64_np

    rng = np.random.default_rng(0)
    x = np.concatenate([rng.integers(0, 2**63, size=20000, dtype=np.uint64) * np.uint64(2) + np.uint64(1),
                        np.array([0, 1, 2**32, 2**64 - 1, 2**63], np.uint64)])
    np.testing.assert_array_equal(oracle.farmhash64(x), farmhash64_np(x))
    for R in (1, 2, 7, 8, 64):
        np.testing.assert_array_equal(oracle.shard_of(x, R), (farmhash64_np(x) % np.uint64(R)).astype(np.uint32))


@pytest.mark.parametrize("cap,keys", [(1, 5), (8, 40), (64, 64), (100, 1000)])
def test_eviction_map_matches_lru_model(oracle, cap, keys):
    """eviction_map.rs:34-111: insert appends (evicting the head beyond capacity), get_refresh moves to the tail."""
    rng = np.random.default_rng(cap + keys)
    m, model = oracle.EvictionMap(cap), OrderedDict()
    for _ in range(5000):
        k = int(rng.integers(0, keys))
        if rng.random() < 0.5:
            hit = m.get_refresh(k)
            assert hit == (k in model)
            if hit:
                model.move_to_end(k)
        else:
            m.insert(k)
            model[k] = True
            model.move_to_end(k)
            if len(model) > cap:
                model.popitem(last=False)
        assert len(m) == len(model)
    for k in range(keys):  # final membership
        assert m.get_refresh(k) == (k in model)


def test_feature_batch_bookkeeping(oracle):
    """FeatureBatch::new (persia-common/src/lib.rs:45-82): distinct signs, every occurrence listed exactly once
    under its sign with its (sample, column), sample_num_signs = ids per sample."""
    rng = np.random.default_rng(3)
    for B, card, mx in ((1, 3, 4), (64, 10, 6), (300, 5000, 3), (7, 2, 0)):
        batch = [rng.integers(0, card, size=rng.integers(0, mx + 1)).tolist() for _ in range(B)]
        ids, off = oracle.lil_to_csr(batch)
        signs, seg, occ_s, occ_c, sns = oracle.FeatureBatch(ids, off).export()
        assert len(set(signs.tolist())) == signs.size
        assert sns.tolist() == [len(x) for x in batch]
        first = list(OrderedDict.fromkeys(ids.tolist()))
        assert signs.tolist() == first  # first-occurrence order (the reference: hashbrown order, unpinned)
        seen = set()
        for u, sign in enumerate(signs.tolist()):
            occ = list(zip(occ_s[seg[u]:seg[u + 1]].tolist(), occ_c[seg[u]:seg[u + 1]].tolist()))
            assert occ == sorted(occ)  # samples ascending, columns ascending inside a sample
            for b, c in occ:
                assert batch[b][c] == sign
                seen.add((b, c))
        assert len(seen) == ids.size == seg[-1]


def test_hashstack_key_ranges_and_counts(oracle):
    """indices_to_hashstack_indices (mod.rs:347-400): round r's keys live in [r*size, (r+1)*size), every id yields
    one key per round, sample_num_signs is multiplied by the number of rounds."""
    from persia_b200.persia_core import farmhash64_np

    rng = np.random.default_rng(5)
    ids = rng.integers(0, 2**40, size=500, dtype=np.uint64)
    off = np.arange(0, 501, 5, dtype=np.uint32)  # 100 samples x 5 ids
    rounds, size = 3, 97
    fb = oracle.FeatureBatch(ids, off)
    fb.hashstack(rounds, size)
    signs, seg, occ_s, occ_c, sns = fb.export()
    assert (sns == 5 * rounds).all()
    assert seg[-1] == ids.size * rounds
    want = set()
    h = ids.copy()
    for r in range(rounds):
        h = farmhash64_np(h)
        want.update((h % np.uint64(size) + np.uint64(r * size)).tolist())
    assert set(signs.tolist()) == want
    assert all(0 <= s < rounds * size for s in signs.tolist())


def test_adam_matches_float64_rendering(oracle):
    """adam_avx2 (persia-simd/src/lib.rs:147-228) + the beta powers of optim.rs:155-197, against plain float64
    arithmetic: agreement to float32 rounding over several steps."""
    rng = np.random.default_rng(9)
    dim, lr, b1, b2, eps = 20, 0.01, 0.9, 0.999, 1e-8
    opt = oracle.Optim(oracle.ADAM, lr=lr, b1=b1, b2=b2, eps=eps)
    w0 = rng.standard_normal(dim).astype(np.float32)
    e = opt.new_entry(w0)
    w, m, v = w0.astype(np.float64), np.zeros(dim), np.zeros(dim)
    p1, p2 = b1, b2
    for _ in range(5):
        g = (rng.standard_normal(dim) * 0.1).astype(np.float32)
        p1, p2 = p1 * b1, p2 * b2  # the accumulated powers are advanced before use
        opt.update(e, g, dim, b1p=np.float32(p1), b2p=np.float32(p2))
        m = b1 * m + (1 - b1) * g.astype(np.float64)
        v = b2 * v + (1 - b2) * g.astype(np.float64) ** 2
        w = w - lr * (m / (1 - p1)) / (np.sqrt(v / (1 - p2)) + eps)
        np.testing.assert_allclose(e[:dim], w, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(e[dim:2 * dim], m, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(e[2 * dim:3 * dim], v, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("kind_name", ["SGD", "ADAGRAD"])
def test_results_do_not_depend_on_the_number_of_parameter_servers(oracle, kind_name):
    """Sharding by farmhash64 % R only decides where a row lives: outputs and updated rows are the same for any R
    when every sample holds one id (multi-id sums follow the shard order, mod.rs:547-561)."""
    kind = getattr(oracle, kind_name)
    S, B, dim = 3, 128, 12
    pf = [oracle.index_prefix(i) for i in range(S)]
    rng = np.random.default_rng(11)
    batches = [(np.concatenate([rng.integers(0, c, size=B, dtype=np.uint64) for c in (7, 300, 10**6)]),
                (rng.standard_normal((S, B, dim)) * 1e-2).astype(np.float16)) for _ in range(3)]
    outs, rows = [], []
    for R in (1, 2, 5):
        w = oracle.Worker([oracle.SlotCfg(dim, prefix=p) for p in pf], n_ps=R)
        w.configure()
        w.set_optimizer(oracle.Optim(kind, lr=0.05))
        got, touched = [], set()
        for ids, g in batches:
            o, ctx = w.forward(ids, np.arange(S * B + 1, dtype=np.uint32), B, training=True)
            got.append(np.stack(o))
            for s in range(S):
                touched.update(w.ctx_signs(ctx, s).tolist())
            w.backward(ctx, [g[s] for s in range(S)])
        outs.append(got)
        rows.append({t: w.get_entry(t) for t in sorted(touched)})
        assert sum(w.ps_len(r) for r in range(R)) == len(touched)
    for k in (1, 2):
        for a, b in zip(outs[0], outs[k]):
            np.testing.assert_array_equal(a.view(np.uint16), b.view(np.uint16))
        assert rows[0].keys() == rows[k].keys()
        for t in rows[0]:
            assert rows[0][t].tobytes() == rows[k][t].tobytes()


def test_raw_and_summed_views_of_one_id_samples_agree(oracle):
    """With one id per sample a raw slot's table row index[b] is the summed slot's output row b (same lookup, no
    pooling, same RNE to f16)."""
    dim, B = 8, 200
    pf = oracle.index_prefix(0)
    rng = np.random.default_rng(13)
    ids = rng.integers(0, 60, size=B, dtype=np.uint64)
    off = np.arange(B + 1, dtype=np.uint32)
    ws = oracle.Worker([oracle.SlotCfg(dim, prefix=pf)], n_ps=2)
    wr = oracle.Worker([oracle.SlotCfg(dim, summation=False, sample_fixed_size=1, prefix=pf)], n_ps=2)
    for w in (ws, wr):
        w.configure()
        w.set_optimizer(oracle.Optim(oracle.SGD, lr=0.1))
    (summed,), _ = ws.forward(ids, off, B, training=True)
    table, index, non_empty, num, _ = wr.forward_raw(0, ids, off, B, training=True)
    assert (num == 1).all() and non_empty.tolist() == list(range(B))
    np.testing.assert_array_equal(table[index].view(np.uint16), summed.view(np.uint16))
