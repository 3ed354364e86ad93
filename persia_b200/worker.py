"""ShardedEmbeddingWorker — the embedding worker's fan-out over R GPUs of one box.

Reference: rust/persia-embedding-server/src/embedding_worker_service/mod.rs.  There the EW shards a batch's
signs by farmhash64(sign) % R and issues one HTTP `lookup_mixed` / `update_gradient_mixed` per parameter
server (:886-919, :835-859).  Here every rank is at once a data-parallel trainer (its own slice of the batch),
an embedding worker (prefix, partition, exchange) and parameter server r (its pb_table); the R requests become
one all-to-all of signs, one of rows back, and in backward one of gradients — torch.distributed (NCCL over
NVLink) carries them, the CUDA library does everything else.

Semantics (documented in DESIGN.md §Multi-GPU): a step is synchronous over the GLOBAL batch — every shard
sees all ranks' occurrences of its signs at once, exactly what the reference computes when one embedding
worker is handed the concatenated batch; duplicates across ranks are reduced in (rank, sample) order.

`backend` isolates what touches a device so that the exchange bookkeeping (partition order, split sizes,
permutations) can be exercised on CPU with gloo (tests/test_worker_gloo.py supplies an oracle-backed one).
"""
import torch
import torch.distributed as dist


class CudaBackend:
    """The product backend: libpersia_b200.so on the current CUDA device."""

    def __init__(self, dim, capacity, device, optimizer, hyper, max_occurrences):
        from . import native as N
        from . import shard as SH

        self.SH, self.N = SH, N
        self.device = device
        self.dim = dim
        self.shard = SH.EmbeddingShard(dim, capacity, device)
        self.shard.set_optimizer(**optimizer)
        self.shard.configure(**hyper)
        # owner-side context: one logical slot of already prefixed signs
        self.ctx = SH.BatchContext(max_occurrences, max_occurrences, [0], device=device)
        self.ctx.set_owner_mode(True)

    def add_prefix(self, ids, slot_occ_off, prefixes, prefix_bit):
        return self.SH.add_prefix(ids, slot_occ_off, prefixes, prefix_bit)

    def partition(self, signs, R):
        return self.SH.partition_by_shard(signs, R)

    def take(self, src, perm):
        return self.SH.permute_u64(src, perm)

    def take_rows(self, src, perm):
        return self.SH.permute_rows(src, perm, scatter=False)

    def put_rows(self, src, perm):
        return self.SH.permute_rows(src, perm, scatter=True)

    def serve_lookup(self, signs, training):
        m = signs.numel()
        if m == 0:
            return torch.empty((0, self.dim), dtype=torch.float16, device=self.device)
        return self.ctx.forward(self.shard, signs, [0, m], m, training=training).view(m, self.dim)

    def serve_update(self, grads, scale):
        if grads.shape[0]:
            self.ctx.backward(self.shard, [grads], scales=[scale])

    def empty_rows(self, n, dtype=torch.float16):
        return torch.empty((n, self.dim), dtype=dtype, device=self.device)


class ShardedEmbeddingWorker:
    def __init__(self, n_slots, dim, prefixes, backend, group=None, prefix_bit=8):
        self.S, self.dim, self.prefixes, self.prefix_bit = n_slots, dim, list(prefixes), prefix_bit
        self.be = backend
        self.group = group
        self.R = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._pending = None

    # ---- exchange primitives -------------------------------------------------------------------------
    def _exchange_counts(self, counts):
        """counts: int32[R] on device -> (send list, recv list).  One host sync: torch's all_to_all_single needs
        the split sizes on the host (NCCL's API does)."""
        if self.R == 1:
            c = int(counts[0])
            return [c], [c]
        recv = torch.empty_like(counts)
        dist.all_to_all_single(recv, counts, group=self.group)
        both = torch.stack([counts, recv]).tolist()
        return both[0], both[1]

    def _a2a(self, send, send_splits, recv_splits, out=None):
        if self.R == 1:
            return send
        shape = (sum(recv_splits),) + tuple(send.shape[1:])
        if out is None:
            out = torch.empty(shape, dtype=send.dtype, device=send.device)
        dist.all_to_all_single(out, send, output_split_sizes=recv_splits, input_split_sizes=send_splits, group=self.group)
        return out

    # ---- forward_batched_direct (mod.rs:1076-1107 -> :874-942) -----------------------------------------
    def forward(self, ids, batch, training=True):
        """ids: int64-bit raw ids [n_slots * batch] (slot-major, one id per sample per slot) on the device.
        Returns f16 [n_slots, batch, dim]."""
        S, B = self.S, batch
        n = S * B
        assert ids.numel() == n, "the multi-GPU path takes one id per sample per slot (Criteo layout)"
        slot_off = [s * B for s in range(S + 1)]
        signs = self.be.add_prefix(ids, slot_off, self.prefixes, self.prefix_bit)
        perm, counts = self.be.partition(signs, self.R)  # stable: (slot, sample) order kept inside a shard
        send_signs = self.be.take(signs, perm)
        send_splits, recv_splits = self._exchange_counts(counts)
        recv_signs = self._a2a(send_signs, send_splits, recv_splits)
        rows = self.be.serve_lookup(recv_signs, training)                      # [m, dim] f16, this shard's rows
        back = self._a2a(rows, recv_splits, send_splits)                       # [n, dim] in partition order
        out = self.be.put_rows(back, perm)                                     # batch order: row s*B+b
        if training:
            self._pending = (perm, send_splits, recv_splits, n)
        return out.view(S, B, self.dim)

    # ---- update_gradient_batched (mod.rs:1109-1129 -> :703-872) ------------------------------------------
    def backward(self, grads, scale=1.0):
        """grads: f16 [n_slots, batch, dim] (the loss-scaled gradient of forward()'s output)."""
        assert self._pending is not None, "no forward batch is pending"
        perm, send_splits, recv_splits, n = self._pending
        self._pending = None
        g = grads.reshape(n, self.dim)
        # NaN rule: the reference skips a slot whose gradient holds a NaN (mod.rs:731-746).  Here the scan runs on
        # the owner over what it received (device side, no host sync): a shard that is handed any NaN skips this
        # step's update of its rows; see DESIGN.md §Multi-GPU for the difference.
        send = self.be.take_rows(g, perm)
        recv = self._a2a(send, send_splits, recv_splits)
        self.be.serve_update(recv, scale)
        return True
