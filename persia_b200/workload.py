"""Criteo-1TB-shaped synthetic workload (BASELINE.json configs 2-4): 26 categorical slots, one id per
sample per slot, per-slot cardinalities shaped like the public Criteo-Terabyte vocabulary sizes rescaled
to a target total, ids Zipf(alpha)-distributed per slot.  numpy only (host side; seeds fixed by callers).
"""
import numpy as np

# per-field vocabulary sizes of the Criteo Terabyte click logs as used by the MLPerf DLRM benchmark
CRITEO_1TB_CARDINALITIES = [
    39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546, 403346, 10, 2208, 11938, 155, 4,
    976, 14, 39979771, 25641295, 39664984, 585935, 12972, 108, 36,
]


def scaled_cardinalities(total_rows, n_slots=26):
    """Rescale the large fields so the table holds `total_rows` rows in all; small fields keep their size."""
    base = np.array((CRITEO_1TB_CARDINALITIES * ((n_slots + 25) // 26))[:n_slots], dtype=np.float64)
    small = base <= 1000
    budget = float(total_rows) - base[small].sum()
    if budget <= 0:  # tiny test tables: shrink everything
        card = np.maximum(1, np.floor(base * (total_rows / base.sum()))).astype(np.int64)
    else:
        card = base.copy()
        card[~small] = np.maximum(1001, np.floor(base[~small] * (budget / base[~small].sum())))
        card = card.astype(np.int64)
    card[int(np.argmax(card))] += int(total_rows) - int(card.sum())
    assert card.sum() == int(total_rows) and card.min() >= 1
    return card


def zipf_ids(rng, cardinality, size, alpha=1.05):
    """Bounded Zipf over [0, cardinality) by inverting the continuous power-law CDF (rank 0 = hottest)."""
    u = rng.random(size)
    n = float(cardinality)
    if abs(alpha - 1.0) < 1e-9:
        x = np.exp(u * np.log(n + 1.0))
    else:
        a = 1.0 - alpha
        x = ((np.power(n + 1.0, a) - 1.0) * u + 1.0) ** (1.0 / a)
    return np.minimum(np.floor(x).astype(np.int64) - 1, cardinality - 1).clip(0).astype(np.uint64)


def make_batches(seed, cardinalities, batch, n_batches, alpha=1.05, scramble=True):
    """Returns ids u64 [n_batches, n_slots*batch] (slot-major inside a batch).
    scramble: map the Zipf rank through a fixed bijection-ish multiplicative hash inside the slot so hot rows
    are not the first rows of the table."""
    rng = np.random.default_rng(seed)
    S = len(cardinalities)
    out = np.empty((n_batches, S * batch), np.uint64)
    for s, card in enumerate(cardinalities):
        card = int(card)
        ids = zipf_ids(rng, card, n_batches * batch, alpha).reshape(n_batches, batch)
        if scramble and card > 1:
            ids = (ids * np.uint64(2654435761) + np.uint64(s * 97 + 13)) % np.uint64(card)
        out[:, s * batch:(s + 1) * batch] = ids
    return out


def index_prefixes(n_slots, prefix_bit=8):
    """parse_embedding_config: every slot is its own feature group (persia-embedding-config lib.rs:600-650)."""
    return [(g + 1) << (64 - prefix_bit) for g in range(n_slots)]


def algorithmic_bytes_per_id(dim, state_floats, phase="total"):
    """SURVEY.md §8(d) per-occurrence figures at U=N, single-id slots, fp32 table, f16 activations/grads."""
    D, S = dim, state_floats
    parts = {
        "partition": 20,
        "lookup": 16 + 4 * D,
        "pool_out": 2 * D,
        "grad_in": 2 * D,
        "update": 16 + 8 * (D + S),
    }
    parts["forward"] = parts["lookup"] + parts["pool_out"]
    parts["backward"] = parts["grad_in"] + parts["update"]
    parts["total"] = parts["partition"] + parts["forward"] + parts["backward"]
    return parts[phase]


def make_dlrm_tower(n_slots, dim, n_dense=13, bottom=(512, 256), top=(1024, 1024, 512, 256)):
    """A DLRM-style dense tower (the caller of the path, not part of it): bottom MLP over the dense features, pairwise
    dot-product interaction between its output and the n_slots pooled embeddings, top MLP to one logit.  Called as
    `model(non_id_type_tensors, embedding_tensors)` like every PERSIA model (persia/ctx.py:446-448; shape after
    examples/src/adult-income/model.py and the MLPerf DLRM)."""
    import torch
    from torch import nn

    class DLRMTower(nn.Module):
        def __init__(self):
            super().__init__()
            layers, d = [], n_dense
            for h in tuple(bottom) + (dim,):
                layers += [nn.Linear(d, h), nn.ReLU()]
                d = h
            self.bottom = nn.Sequential(*layers)
            n_vec = n_slots + 1
            tri = torch.triu_indices(n_vec, n_vec, offset=1)
            self.register_buffer("tri_flat", tri[0] * n_vec + tri[1], persistent=False)  # upper triangle of the Gram matrix
            layers, d = [], dim + n_vec * (n_vec - 1) // 2
            for h in top:
                layers += [nn.Linear(d, h), nn.ReLU()]
                d = h
            layers.append(nn.Linear(d, 1))
            self.top = nn.Sequential(*layers)

        def forward(self, non_id_type_tensors, embedding_tensors):
            dense = non_id_type_tensors[0].float()
            x = self.bottom(dense)                                                    # [B, dim]
            e = torch.stack(list(embedding_tensors), dim=1).float()                   # one stack of the f16 slots, one cast
            vecs = torch.cat([x.unsqueeze(1), e], dim=1)                              # [B, n_slots + 1, dim]
            inter = torch.bmm(vecs, vecs.transpose(1, 2))                             # pooled-embedding x dense-bottom interaction
            z = torch.cat([x, inter.flatten(1).index_select(1, self.tri_flat)], dim=1)
            return self.top(z).squeeze(-1)

    return DLRMTower()
