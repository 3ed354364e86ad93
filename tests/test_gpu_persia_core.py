"""GPU: the persia_core surface end to end, driven the way persia/ctx.py drives it (forward: ctx.py:75-199,
backward: ctx.py:926-1005), against the oracle's embedding worker."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu

CFG = {
    "feature_index_prefix_bit": 8,
    "slots_config": {
        "user": {"dim": 16},
        "item": {"dim": 16, "sqrt_scaling": True},
        "tags": {"dim": 32},
        "shop": {"dim": 16},
        "hs": {"dim": 16, "hash_stack_config": {"hash_stack_rounds": 2, "embedding_size": 1000}},
    },
}
CARD = {"user": 500, "item": 40, "tags": 3000, "shop": 7, "hs": 100000}
KEYS = dict(CARD, hs=2000)  # resident key space per slot (hash stack: rounds x embedding_size)


def _batch(pc, rng, B):
    lil = {"user": rng.integers(0, CARD["user"], size=B, dtype=np.uint64),
           "item": [rng.integers(0, CARD["item"], size=rng.integers(0, 5), dtype=np.uint64) for _ in range(B)],
           "tags": [rng.integers(0, CARD["tags"], size=rng.integers(1, 4), dtype=np.uint64) for _ in range(B)],
           "shop": rng.integers(0, CARD["shop"], size=B, dtype=np.uint64),
           "hs": rng.integers(0, CARD["hs"], size=B, dtype=np.uint64)}
    b = pc.data.PersiaBatch()
    for name, v in lil.items():
        if isinstance(v, np.ndarray):
            b.add_id_type_feature_with_single_id(v, name)
        else:
            b.add_id_type_feature(v, name)
    b.add_non_id_type_feature(rng.standard_normal((B, 3)).astype(np.float32), np.dtype(np.float32), "dense")
    b.add_label(rng.integers(0, 2, size=(B, 1)).astype(np.float32), np.dtype(np.float32), "y")
    b.converted_id_type_features2embedding_tensor(True)
    return b, lil


def _oracle_worker(oracle, pc):
    from persia_b200.persia_core import parse_embedding_config

    _, slots = parse_embedding_config(CFG)
    ws = {}
    for dim in sorted({s.dim for s in slots}):
        part = [s for s in slots if s.dim == dim]
        w = oracle.Worker([oracle.SlotCfg(dim, sqrt_scaling=s.sqrt_scaling, prefix=s.index_prefix,
                                          hs_rounds=s.hash_stack_rounds, hs_size=s.hash_stack_embedding_size) for s in part], n_ps=1)
        w.configure(-0.01, 0.01, 1.0, True, 10.0)
        w.set_optimizer(oracle.Optim(oracle.SGD, lr=0.05, wd=0.0))
        ws[dim] = (w, [s.name for s in part])
    return ws


def _csr(lils, B):
    rows = []
    for v in lils:
        rows.extend([np.array([x], np.uint64) for x in v] if isinstance(v, np.ndarray) else v)
    off = np.zeros(len(rows) + 1, np.uint32)
    off[1:] = np.cumsum([len(r) for r in rows])
    return (np.concatenate(rows) if rows else np.zeros(0, np.uint64)), off


def test_surface_forward_backward_matches_oracle(oracle):
    import torch
    import torch.utils.dlpack as dl

    from persia_b200 import persia_core as impl

    impl.reset()
    pc = impl.install()
    try:
        pc.set_embedding_config(CFG)
        ctx = pc.PersiaCommonContext(10, 0, 1, 0)
        opt = pc.optim.OptimizerBase()
        opt.init_sgd(0.05, 0.0)
        opt.apply()
        ctx.configure_embedding_parameter_servers(-0.01, 0.01, 1.0, True, 10.0)
        bwd = pc.backward.Backward(8)
        bwd.launch(2)
        ws = _oracle_worker(oracle, pc)
        rng = np.random.default_rng(3)
        B = 96
        # a batch through the dataflow path (send ids -> remote ref; send dense -> channel -> Forward.get_batch)
        ch = pc.utils.PersiaBatchDataChannel(4)
        pc.nats.initialize_dataflow(1, ch.get_sender())
        fwd = pc.forward.Forward(8, True, 1)
        fwd.set_input_channel(ch.get_receiver())
        fwd.launch(2)
        for step in range(4):
            b, lil = _batch(pc, rng, B)
            if step % 2 == 0:
                tb = ctx.get_embedding_from_data(b, 0)
            else:
                ctx.send_id_type_features_to_embedding_worker(b)
                assert b.batch_id() >= 0
                ctx.send_non_id_type_features_to_nn_worker(b)
                tb = fwd.get_batch(1000)
            dense = [dl.from_dlpack(t.dlpack) for t in tb.consume_all_non_id_type_feature_tensors()]
            labels = [dl.from_dlpack(t.dlpack) for t in tb.consume_all_label_tensors()]
            assert dense[0].shape == (B, 3) and labels[0].is_cuda
            embs = tb.consume_all_id_type_feature_embedding_tensors()
            keep, torch_embs = [], {}
            for e in embs:
                assert not e.is_raw_embedding()
                t = e.get_sum_embedding()
                keep.append(t)
                x = dl.from_dlpack(t.dlpack)
                assert x.dtype == torch.float16
                x.requires_grad = True
                torch_embs[t.name] = x
            # oracle forward, per dim group
            octx = {}
            for dim, (w, names) in ws.items():
                ids, off = _csr([lil[n] for n in names], B)
                want, octx[dim] = w.forward(ids, off, B, training=True)
                for n, wv in zip(names, want):
                    got = torch_embs[n].detach().cpu().numpy()
                    assert np.abs(got.astype(np.float32) - wv.astype(np.float32)).max() <= 2e-5  # <= 1 f16 ulp at |x|<=0.04
                    if isinstance(lil[n], np.ndarray) and n != "hs":
                        np.testing.assert_array_equal(got.view(np.uint16), wv.view(np.uint16))
            # a loss and its gradients, then the GradientBatch protocol
            scale = 128.0
            loss = sum((x.float() * (i + 1)).sum() for i, x in enumerate(torch_embs.values())) * scale * 1e-3
            loss.backward()
            gb = tb.create_gradient_batch()
            for n, x in torch_embs.items():
                if step == 3 and n == "shop":
                    gb.add_skipped_gradient(n)
                else:
                    gb.add_gradient(n, x.grad.data_ptr(), list(x.grad.shape), True, scale)
            torch.cuda.synchronize()
            bwd.update_id_type_feature_gradient_batched(gb)
            for dim, (w, names) in ws.items():
                g = [torch_embs[n].grad.cpu().numpy() for n in names]
                w.backward(octx[dim], g, scale=[scale] * len(names),
                           skip=[int(step == 3 and n == "shop") for n in names])
        sizes = ctx.get_embedding_size()
        assert sizes == [ws[d][0].ps_len(0) for d in sorted(ws)]
        # rows equal the oracle's (single-id slots bit for bit; multi-id slots see 1-ulp different f16 inputs only
        # through the forward, the update itself reads f16 gradients produced by torch: identical on both sides)
        from persia_b200.persia_core import _S
        from util import to_dev_ids

        for dim, (w, names) in ws.items():
            sh = _S.groups[dim]["shard"]
            for n in names:
                pf = _S.by_name[n].index_prefix
                signs = oracle.add_prefix(np.arange(KEYS[n], dtype=np.uint64), 8, pf)
                ent, found = sh.get_entries(to_dev_ids(signs, "cuda:0"))
                ent, found = ent.cpu().numpy(), found.cpu().numpy()
                for k, s in enumerate(signs):
                    ref = w.get_entry(int(s))
                    assert (ref is not None) == bool(found[k])
                    if ref is not None:
                        assert ent[k].tobytes() == ref.tobytes(), (n, k)
        ctx.clear_embeddings()
        assert ctx.get_embedding_size() == [0, 0]
    finally:
        impl.reset()


RAW_CFG = {
    "feature_index_prefix_bit": 8,
    "slots_config": {
        "seq": {"dim": 16, "embedding_summation": False, "sample_fixed_size": 4},
        "user": {"dim": 16},
    },
}


def test_surface_raw_slot_matches_oracle(oracle):
    """A raw (embedding_summation: false) slot through the surface, consumed the way persia/ctx.py:121-165 builds
    the [B, fixed, dim+1] tensor and :966-981 builds the [U, dim] f32 gradient."""
    import torch
    import torch.utils.dlpack as dl

    from persia_b200 import persia_core as impl
    from persia_b200.persia_core import parse_embedding_config

    impl.reset()
    pc = impl.install()
    try:
        pc.set_embedding_config(RAW_CFG)
        ctx = pc.PersiaCommonContext(10, 0, 1, 0)
        opt = pc.optim.OptimizerBase()
        opt.init_sgd(0.05, 0.0)
        opt.apply()
        ctx.configure_embedding_parameter_servers(-0.01, 0.01, 1.0, True, 10.0)
        bwd = pc.backward.Backward(8)
        _, slots = parse_embedding_config(RAW_CFG)
        by = {s.name: s for s in slots}
        # the oracle serves the two slots as two one-slot requests against the same parameter server
        w = oracle.Worker([oracle.SlotCfg(16, summation=False, sample_fixed_size=4, prefix=by["seq"].index_prefix),
                           oracle.SlotCfg(16, prefix=by["user"].index_prefix)], n_ps=1)
        w.configure(-0.01, 0.01, 1.0, True, 10.0)
        w.set_optimizer(oracle.Optim(oracle.SGD, lr=0.05, wd=0.0))
        rng = np.random.default_rng(21)
        B, fixed, dim = 64, 4, 16
        seen = set()
        for step in range(3):
            seq = [rng.integers(0, 50, size=rng.integers(0, 7), dtype=np.uint64) for _ in range(B)]
            b = pc.data.PersiaBatch()
            b.add_id_type_feature(seq, "seq")
            b.add_label(np.zeros((B, 1), np.float32), np.dtype(np.float32), "y")
            b.converted_id_type_features2embedding_tensor(True)
            tb = ctx.get_embedding_from_data(b, 0)
            (e,) = tb.consume_all_id_type_feature_embedding_tensors()
            assert e.is_raw_embedding()
            with pytest.raises(RuntimeError):
                e.get_sum_embedding()
            raw, index, non_empty, sample_id_num = e.get_raw_embedding()
            distinct = dl.from_dlpack(raw.dlpack)
            index_t = dl.from_dlpack(index.dlpack)
            non_empty_t = dl.from_dlpack(non_empty.dlpack)
            ids = np.concatenate(seq) if sum(map(len, seq)) else np.zeros(0, np.uint64)
            off = np.zeros(B + 1, np.uint32)
            off[1:] = np.cumsum([len(r) for r in seq])
            wt, wi, wne, wnum, octx = w.forward_raw(0, ids, off, B, training=True)
            assert distinct.cpu().numpy().tobytes() == wt.tobytes()
            np.testing.assert_array_equal(index_t.cpu().numpy(), wi)
            np.testing.assert_array_equal(non_empty_t.cpu().numpy(), wne)
            assert sample_id_num == wnum.tolist()
            seen.update(oracle.add_prefix(ids, 8, by["seq"].index_prefix).tolist())
            # persia/ctx.py:131-165
            assert index_t.max() < distinct.shape[0]
            sel = distinct.index_select(0, index_t.view(-1))
            sel.requires_grad = True
            x = sel.view(-1, fixed, dim)
            mask = (index_t.view(B, fixed, 1) != 0).half()
            out = torch.cat([x, mask], dim=2)
            scale = 64.0
            (out.float() * torch.arange(1, dim + 2, device=out.device)).sum().mul(scale * 1e-3).backward()
            # persia/ctx.py:966-981
            gb = tb.create_gradient_batch()
            U = distinct.shape[0] - 1
            if distinct.shape[0] > 1:
                grad = torch.zeros_like(distinct, dtype=torch.float32)
                nz = sel.grad.index_select(0, non_empty_t.view(-1)).float()
                grad.index_add_(0, index_t.view(-1)[non_empty_t.view(-1)], nz)
                grad = grad[1:, :].contiguous()
                gb.add_gradient("seq", grad.data_ptr(), list(grad.shape), False, scale)
            torch.cuda.synchronize()
            bwd.update_id_type_feature_gradient_batched(gb)
            if U:
                assert w.backward_raw(0, octx, grad.cpu().numpy(), scale=scale) == 0
        from persia_b200.persia_core import _S
        from util import to_dev_ids

        signs = np.array(sorted(seen), np.uint64)
        ent, found = _S.groups[16]["shard"].get_entries(to_dev_ids(signs, "cuda:0"))
        assert found.all()
        ent = ent.cpu().numpy()
        for k, s in enumerate(signs):
            assert ent[k].tobytes() == w.get_entry(int(s)).tobytes()
    finally:
        impl.reset()
