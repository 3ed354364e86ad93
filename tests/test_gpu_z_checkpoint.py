"""GPU: embedding checkpoints in the reference's `.emb` layout (persia_b200/checkpoint.py) — dump a trained shard,
read the file back with the format decoder, restore it into a fresh shard, and the same through the `persia_core`
surface (`PersiaCommonContext.dump` / `load`, lib.rs:356-378)."""
import os

import numpy as np
import pytest

from util import make_batch, to_dev_ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    return torch


def test_dump_decode_load_round_trip(torch_cuda, oracle, tmp_path):
    torch = torch_cuda
    from persia_b200 import checkpoint as CK
    from persia_b200 import native as N
    from persia_b200 import shard as pb

    dim, S, B = 16, 3, 256
    pf = [oracle.index_prefix(i) for i in range(S)]
    s = pb.EmbeddingShard(dim, 1 << 14, 0)
    s.set_optimizer(N.OPT_ADAGRAD, lr=0.05, initialization=0.01)
    s.configure()
    ctx = pb.BatchContext(S * B, S * B, pf)
    rng = np.random.default_rng(4)
    seen = []
    for it in range(3):  # three training requests: later ones are more recently used
        ids, _, slot_off = make_batch(rng, S, B, [50, 400, 3000])
        ctx.forward(s, to_dev_ids(ids, DEV), slot_off, B, training=True)
        g = (rng.standard_normal((S, B, dim)) * 1e-2).astype(np.float16)
        ctx.backward(s, [torch.from_numpy(g[i]).to(DEV) for i in range(S)])
        seen.append(np.concatenate([oracle.add_prefix(ids[i * B:(i + 1) * B], 8, pf[i]) for i in range(S)]))
    # the three signs that collide with the index's cell markers live in reserved cells: they must travel too
    odd = np.array([2**64 - 1, 2**64 - 2, 2**64 - 3], np.uint64)
    s.set_entries(to_dev_ids(odd, DEV), torch.arange(3 * 2 * dim, dtype=torch.float32, device=DEV).view(3, 2 * dim))
    resident = np.unique(np.concatenate(seen + [odd]))
    assert len(s) == resident.size

    signs, rec = s.export_signs()
    got = signs.cpu().numpy().view(np.uint64)
    assert sorted(got.tolist()) == resident.tolist()
    r = rec.cpu().numpy().astype(np.int64)
    assert (np.diff(r) >= 0).all()  # least recently used first, like the reference's list
    last_used = {int(x): k for k, b in enumerate(seen) for x in b}  # request of the last use
    order = [last_used[x] for x in got.tolist() if x in last_used]
    assert order == sorted(order)

    CK.dump_shards(str(tmp_path), [s])
    files = sorted(os.listdir(tmp_path / "s0"))
    assert files == ["embedding_dump_done", "replica_0_shard_0.emb"] and (tmp_path / "embedding_dump_done").is_file()
    fs, fd, fe = CK.decode_list((tmp_path / "s0" / "replica_0_shard_0.emb").read_bytes())
    np.testing.assert_array_equal(fs, got)  # file order == export order
    assert (fd == dim).all()
    ent, found = s.get_entries(signs)
    assert found.all()
    np.testing.assert_array_equal(np.stack(fe), ent.cpu().numpy())

    fresh = pb.EmbeddingShard(dim, 1 << 14, 0)
    fresh.set_optimizer(N.OPT_ADAGRAD, lr=0.05, initialization=0.01)
    fresh.configure()
    assert CK.load_shards(str(tmp_path), {dim: fresh}) == resident.size
    assert len(fresh) == resident.size
    ent2, found2 = fresh.get_entries(signs)
    assert found2.all() and torch.equal(ent, ent2)
    # a checkpoint of another optimizer's entry length is refused, an unfinished directory too
    sgd = pb.EmbeddingShard(dim, 1 << 14, 0)
    sgd.set_optimizer(N.OPT_SGD, lr=0.1)
    sgd.configure()
    with pytest.raises(RuntimeError):
        CK.load_shards(str(tmp_path), {dim: sgd})
    os.remove(tmp_path / "s0" / "embedding_dump_done")
    with pytest.raises(RuntimeError):
        CK.load_shards(str(tmp_path), {dim: fresh})


def test_surface_dump_and_load(torch_cuda, tmp_path):
    from persia_b200 import persia_core as impl

    impl.reset()
    pc = impl.install()
    try:
        pc.set_embedding_config({"feature_index_prefix_bit": 8,
                                 "slots_config": {"a": {"dim": 8}, "b": {"dim": 8}, "c": {"dim": 32}}})
        ctx = pc.PersiaCommonContext(10, 0, 1, 0)
        opt = pc.optim.OptimizerBase()
        opt.init_adagrad(0.01, 0.0, 1.0, 0.01, 1e-10, False)
        opt.apply()
        ctx.configure_embedding_parameter_servers(-0.01, 0.01, 1.0, True, 10.0)
        rng = np.random.default_rng(8)
        rows = [(int(sign), rng.standard_normal(d).astype(np.float32), rng.random(d).astype(np.float32))
                for d, n in ((8, 300), (32, 120)) for sign in rng.integers(1, 2**60, size=n, dtype=np.uint64)]
        ctx.set_embedding(rows)
        assert ctx.get_embedding_size() == [300, 120]
        ctx.dump(str(tmp_path))
        ctx.wait_for_emb_dumping()
        assert sorted(os.listdir(tmp_path / "s0")) == ["embedding_dump_done", "replica_0_shard_0.emb", "replica_0_shard_1.emb"]
        ctx.clear_embeddings()
        assert ctx.get_embedding_size() == [0, 0]
        ctx.load(str(tmp_path))
        ctx.wait_for_emb_loading()
        assert ctx.get_embedding_size() == [300, 120]
        from persia_b200.persia_core import _S

        for sign, emb, st in rows[:40] + rows[-40:]:
            sh = _S.groups[emb.size]["shard"]
            ent, found = sh.get_entries(to_dev_ids(np.array([sign], np.uint64), DEV))
            assert bool(found[0])
            np.testing.assert_array_equal(ent[0].cpu().numpy(), np.concatenate([emb, st]))
    finally:
        impl.reset()
