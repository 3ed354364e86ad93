"""The host tier's store (persia_b200/tier.py: sorted runs, vectorised lookups) against a dictionary."""
import numpy as np


def test_runs_store_matches_a_dictionary():
    from persia_b200.tier import _Runs

    r, rng, ref = _Runs(max_runs=3), np.random.default_rng(0), {}
    for it in range(40):
        k = rng.choice(np.arange(1, 100000, dtype=np.uint64), 500, replace=False)
        k = np.array([x for x in k if int(x) not in ref], np.uint64)
        v = rng.standard_normal((k.size, 4)).astype(np.float32)
        r.add(k, v)
        ref.update({int(a): b for a, b in zip(k, v)})
        q = rng.choice(np.arange(1, 100000, dtype=np.uint64), 300, replace=False)
        fk, fv = r.take(q)
        assert set(fk.tolist()) == {int(x) for x in q if int(x) in ref}
        for a, b in zip(fk, fv if fv is not None else []):
            assert (ref.pop(int(a)) == b).all()
        assert len(r) == len(ref)
        some = next(iter(ref)) if ref else None
        if some is not None:
            assert (r.get(some) == ref[some]).all()
    assert r.get(10 ** 9) is None and len(r.runs) <= 4
