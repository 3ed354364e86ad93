"""Host-DRAM tier (BASELINE.json configs[4]): a shard whose HBM holds a working set and a host store that
holds the rest behave like ONE table of unbounded capacity — compared with an oracle that never evicts, bit for bit."""
import numpy as np
import pytest

from util import full_row_off, make_batch, to_dev_ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("kind,dim", [(1, 16), (0, 32)])
def test_shard_plus_host_tier_equals_an_unbounded_table(oracle, kind, dim):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    from persia_b200 import shard as pb
    from persia_b200.tier import HostTier

    rng = np.random.default_rng(17 + dim)
    S, B, card, steps = 3, 200, [40, 3000, 40000], 40
    pf = [oracle.index_prefix(i) for i in range(S)]
    cap = 2500  # rows in HBM: far fewer than the signs the run touches
    s = pb.EmbeddingShard(dim, cap, 0)
    okw = dict(lr=0.05, wd=0.001) if kind == 0 else dict(lr=0.02, init_acc=0.01, eps=1e-10)
    s.set_optimizer(kind, **{{"init_acc": "initialization"}.get(k, k): v for k, v in okw.items()})
    s.configure()
    tier = HostTier(s, reserve=S * B)
    ctx = pb.BatchContext(4 * S * B, 4 * S * B, pf)
    w = oracle.Worker([oracle.SlotCfg(dim, prefix=pf[i]) for i in range(S)], n_ps=1)  # never evicts
    w.configure()
    w.set_optimizer(oracle.Optim(kind, **okw))
    oracle.set_rsqrt_exact(True)
    try:
        seen = set()
        for step in range(steps):
            ids, _, slot_off = make_batch(rng, S, B, card)
            got = ctx.forward(s, to_dev_ids(ids, DEV), slot_off, B, training=True).cpu().numpy()
            want, octx = w.forward(ids, full_row_off(S, B), B, training=True)
            for i in range(S):
                assert got[i].tobytes() == want[i].tobytes(), (step, i)
                seen.update(w.ctx_signs(octx, i).tolist())
            g = (rng.standard_normal((S, B, dim)) * 1e-2).astype(np.float16)
            ctx.backward(s, [torch.from_numpy(g[i]).to(DEV) for i in range(S)])
            w.backward(octx, [g[i] for i in range(S)])
        torch.cuda.synchronize()
        st = tier.stats()
        assert len(seen) > 2 * cap and st["spilled"] > 0 and st["restored"] > 0, (len(seen), st)
        assert len(s) + len(tier) == len(seen)  # every sign lives in exactly one place
        assert s.counters()["capacity_refused"] == 0
        for sign in sorted(seen):
            e = tier.get_entry(sign)
            assert e is not None and e.tobytes() == w.get_entry(int(sign)).tobytes(), sign
    finally:
        oracle.set_rsqrt_exact(False)
        ctx.close()
        s.close()
