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

    # framed (fixed-capacity) exchange helpers
    def frame_signs(self, signs, perm, counts, R, cap, overflow):
        return self.SH.frame_signs(signs, perm, counts, R, cap, overflow)

    def frame_rows(self, src, perm, counts, R, cap, pack, out):
        return self.SH.frame_rows(src, perm, counts, R, cap, pack, out)


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

    # ---- the same two calls with static shapes --------------------------------------------------------
    # Every (source, destination) pair exchanges exactly `cap` slots (padding = PB_NULL_SIGN / zero rows), so no
    # split sizes travel to the host and the whole step — kernels and NCCL collectives — can be captured in a
    # CUDA graph.  `overflow` (device int32) is raised when a pair needs more than cap slots; check_overflow()
    # reads it (a host sync: call it outside the hot loop).
    def enable_static(self, batch, slack=1.3, extra=4096, cap=None):
        n = self.S * batch
        if cap is not None:
            self.cap = min(int(cap), n) if self.R > 1 else n
        else:
            self.cap = int(n / self.R * slack) + extra if self.R > 1 else n
        self.cap = (self.cap + 7) // 8 * 8
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.be.device)
        return self.cap

    def check_overflow(self):
        return bool(int(self.overflow))

    def calibrate_cap(self, id_batches, batch, margin=1.10, extra=256):
        """Capacity per (source, destination) pair from sample batches: the largest per-owner count seen on any rank
        (hot signs make it a property of the id distribution, not of chance), plus a margin.  Collective: every
        rank must call it.  A later batch that still overflows raises the overflow flag (check_overflow)."""
        if self.R == 1:
            return self.S * batch
        slot_off = [s * batch for s in range(self.S + 1)]
        worst, where = 0, None
        for ids in id_batches:
            signs = self.be.add_prefix(ids, slot_off, self.prefixes, self.prefix_bit)
            _, counts = self.be.partition(signs, self.R)
            worst, where = max(worst, int(counts.max())), counts.device
        t = torch.tensor([worst], dtype=torch.int64, device=where)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(int(t) * margin) + extra

    def _a2a_equal(self, send):
        if self.R == 1:
            return send
        out = torch.empty_like(send)
        dist.all_to_all_single(out, send, group=self.group)
        return out

    def forward_static(self, ids, batch, training=True):
        S, B, R, cap = self.S, batch, self.R, self.cap
        n = S * B
        slot_off = [s * B for s in range(S + 1)]
        signs = self.be.add_prefix(ids, slot_off, self.prefixes, self.prefix_bit)
        perm, counts = self.be.partition(signs, R)
        send = self.be.frame_signs(signs, perm, counts, R, cap, self.overflow)       # [R*cap]
        recv = self._a2a_equal(send)
        rows = self.be.serve_lookup(recv, training)                                   # [R*cap, dim] f16
        back = self._a2a_equal(rows)
        out = self.be.frame_rows(back, perm, counts, R, cap, False, self.be.empty_rows(n))
        if training:
            self._pending = (perm, counts, None, n)
        return out.view(S, B, self.dim)

    def backward_static(self, grads, scale=1.0):
        assert self._pending is not None, "no forward batch is pending"
        perm, counts, _, n = self._pending
        self._pending = None
        g = grads.reshape(n, self.dim)
        send = self.be.frame_rows(g, perm, counts, self.R, self.cap, True, self.be.empty_rows(self.R * self.cap, g.dtype))
        recv = self._a2a_equal(send)
        self.be.serve_update(recv, scale)
        return True

    # ---- the framed exchange over NVLink peer memory (no NCCL on the data path) -----------------------------
    def enable_p2p(self, batch):
        """Receive buffers and barrier flags in symmetric memory (torch maps every peer's buffer into this process);
        the exchange is then libpersia_b200's own store-to-peer kernel + flag barrier: stream-ordered, no host
        involvement, capturable in one CUDA graph together with the compute."""
        import torch.distributed._symmetric_memory as symm

        if not hasattr(self, "cap"):
            self.enable_static(batch)
        R, cap, dev = self.R, self.cap, self.be.device
        grp = self.group if self.group is not None else dist.group.WORLD
        self._symm = {}

        def mapped(name, shape, dtype):
            t = symm.empty(shape, dtype=dtype, device=dev)
            t.zero_()
            h = symm.rendezvous(t, grp)
            self._symm[name] = (t, h, [int(p) for p in h.buffer_ptrs])
            return t

        if R > 1:
            self.p2p_recv = mapped("recv", (R * cap,), torch.int64)
            self.p2p_back = mapped("back", (R * cap, self.dim), torch.float16)
            self.p2p_grecv = mapped("grecv", (R * cap, self.dim), torch.float16)
            mapped("flags", (16,), torch.int32)
        else:
            self.p2p_recv = torch.zeros(cap, dtype=torch.int64, device=dev)
            self.p2p_back = torch.zeros((cap, self.dim), dtype=torch.float16, device=dev)
            self.p2p_grecv = torch.zeros((cap, self.dim), dtype=torch.float16, device=dev)
            flags = torch.zeros(16, dtype=torch.int32, device=dev)
            self._symm = {"recv": (self.p2p_recv, None, [self.p2p_recv.data_ptr()]),
                          "back": (self.p2p_back, None, [self.p2p_back.data_ptr()]),
                          "grecv": (self.p2p_grecv, None, [self.p2p_grecv.data_ptr()]),
                          "flags": (flags, None, [flags.data_ptr()])}
        self.p2p_epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.p2p_err = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        if R > 1:
            dist.barrier(group=self.group)
        return self

    def _p2p_send(self, framed, name):
        self.be.SH.p2p_exchange(framed, self._symm[name][2], self.rank, self.cap)
        self.be.SH.p2p_barrier(self._symm["flags"][2], self.p2p_epoch, self.rank, self.p2p_err)

    def forward_p2p(self, ids, batch, training=True):
        S, B, R, cap = self.S, batch, self.R, self.cap
        n = S * B
        slot_off = [s * B for s in range(S + 1)]
        signs = self.be.add_prefix(ids, slot_off, self.prefixes, self.prefix_bit)
        perm, counts = self.be.partition(signs, R)
        send = self.be.frame_signs(signs, perm, counts, R, cap, self.overflow)
        self._p2p_send(send, "recv")
        rows = self.be.serve_lookup(self.p2p_recv, training)
        self._p2p_send(rows, "back")
        out = self.be.frame_rows(self.p2p_back, perm, counts, R, cap, False, self.be.empty_rows(n))
        if training:
            self._pending = (perm, counts, None, n)
        return out.view(S, B, self.dim)

    def backward_p2p(self, grads, scale=1.0):
        assert self._pending is not None, "no forward batch is pending"
        perm, counts, _, n = self._pending
        self._pending = None
        g = grads.reshape(n, self.dim)
        send = self.be.frame_rows(g, perm, counts, self.R, self.cap, True, self.be.empty_rows(self.R * self.cap, g.dtype))
        self._p2p_send(send, "grecv")
        self.be.serve_update(self.p2p_grecv, scale)
        return True

    def check_p2p(self):
        """True if a barrier gave up waiting for a peer (host sync)."""
        return bool(int(self.p2p_err))

    def make_graphed_step(self, ids, grads, batch, stream, scale=1.0):
        """The framed step on fixed buffers (`ids` int64 [S*B], `grads` f16 [S,B,dim]) with every compute segment
        between two collectives replayed as a CUDA graph: 5 graph launches + 3 NCCL calls per step instead of ~25
        kernel launches.  The collectives themselves stay outside the graphs.  Returns (step_fn, out) where out is
        the f16 [S,B,dim] forward result buffer; refresh `ids` / `grads` in place between calls."""
        S, B, R, cap, n = self.S, batch, self.R, self.cap, self.S * batch
        slot_off = [s * B for s in range(S + 1)]
        st = {}

        def seg_a():
            signs = self.be.add_prefix(ids, slot_off, self.prefixes, self.prefix_bit)
            st["perm"], st["counts"] = self.be.partition(signs, R)
            st["send"] = self.be.frame_signs(signs, st["perm"], st["counts"], R, cap, self.overflow)

        def seg_b():
            st["rows"] = self.be.serve_lookup(st["recv"], True)

        def seg_c():
            st["out"] = self.be.frame_rows(st["back"], st["perm"], st["counts"], R, cap, False, self.be.empty_rows(n))

        def seg_d():
            st["gsend"] = self.be.frame_rows(grads.reshape(n, self.dim), st["perm"], st["counts"], R, cap, True,
                                             self.be.empty_rows(R * cap, grads.dtype))

        def seg_e():
            self.be.serve_update(st["grecv"], scale)

        def a2a(key_in, key_out):
            if key_out not in st:
                st[key_out] = torch.empty_like(st[key_in])
            if R == 1:
                st[key_out].copy_(st[key_in])
            else:
                dist.all_to_all_single(st[key_out], st[key_in], group=self.group)

        graphs = {}
        with torch.cuda.stream(stream):
            for _ in range(2):  # warm-up: allocations, NCCL channels
                seg_a(); a2a("send", "recv"); seg_b(); a2a("rows", "back"); seg_c(); seg_d(); a2a("gsend", "grecv"); seg_e()  # noqa: E702
            stream.synchronize()
            for name, fn, nxt in (("a", seg_a, ("send", "recv")), ("b", seg_b, ("rows", "back")), ("c", seg_c, None),
                                  ("d", seg_d, ("gsend", "grecv")), ("e", seg_e, None)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                    fn()
                graphs[name] = g
                if nxt:  # run the collective once so that the next segment captures against a live buffer
                    a2a(*nxt)
            stream.synchronize()

        def step():
            graphs["a"].replay(); a2a("send", "recv")    # noqa: E702
            graphs["b"].replay(); a2a("rows", "back")    # noqa: E702
            graphs["c"].replay()
            graphs["d"].replay(); a2a("gsend", "grecv")  # noqa: E702
            graphs["e"].replay()

        return step, st["out"].view(S, B, self.dim)

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
