"""TEST INFRASTRUCTURE: a CPU model of the sharded protocol (csrc/pb_shard.cu) — numpy for the id plumbing, the oracle's
parameter server for the rows, torch.distributed (gloo) for the "stores into the peer's area".  It fixes what the GPU
kernels must compute: dedup per slot BEFORE the exchange, one request per (rank, owner), fixed `cap` slots per pair, the
NaN rule per slot on the requesting rank, gradients reduced per distinct sign in ascending sample order, and an owner
that applies the R gradient requests of a step one after another in rank order."""
import numpy as np
import torch
import torch.distributed as dist


class ExchangeModel:
    def __init__(self, oracle, prefixes, dim, optim, rank, world, cap):
        self.o, self.pf, self.dim, self.rank, self.R, self.cap = oracle, prefixes, dim, rank, world, cap
        self.ps = oracle.Worker([oracle.SlotCfg(dim)], n_ps=1)  # parameter server `rank`, addressed directly
        self.ps.configure()
        self.ps.set_optimizer(optim)
        self.overflow = False

    def _a2a(self, t):
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t)
        return out

    def forward(self, ids, B):
        """ids: uint64 [S*B] slot-major, one id per sample and slot.  Returns f16 [S, B, dim]."""
        S, R, cap = len(self.pf), self.R, self.cap
        signs = np.concatenate([self.o.add_prefix(ids[i * B:(i + 1) * B], 8, self.pf[i]) for i in range(S)])
        items, where = [], {}          # distinct (slot, sign) in first-occurrence order; occurrence lists
        occ_item = np.empty(S * B, np.int64)
        for i in range(S * B):
            key = (i // B, int(signs[i]))
            if key not in where:
                where[key] = len(items)
                items.append([key[1], key[0], []])
            occ_item[i] = where[key]
            items[where[key]][2].append(i)
        item_sign = np.array([it[0] for it in items], np.uint64)
        owner = self.o.shard_of(item_sign, R).astype(np.int64)
        send = torch.zeros((R, cap), dtype=torch.int64)
        counts = torch.zeros(R, dtype=torch.int64)
        target = np.full(len(items), -1, np.int64)
        for u in range(len(items)):
            q = int(owner[u])
            k = int(counts[q])
            counts[q] += 1
            if k < cap:
                send[q, k] = int(item_sign[u].view(np.int64)) if hasattr(item_sign[u], "view") else int(item_sign[u])
                target[u] = q * cap + k
            else:
                self.overflow = True
        counts = torch.clamp(counts, max=cap)
        recv_cnt = self._a2a(counts)                                  # ctrl words: signs per source
        recv = self._a2a(send.view(-1)).view(R, cap)                  # this owner's sign area, by source
        rows = torch.zeros((R, cap, self.dim), dtype=torch.float16)
        self.own_signs = []
        for src in range(R):                                          # lookup_mixed of the R requests
            n = int(recv_cnt[src])
            sg = recv[src, :n].numpy().view(np.uint64).copy()
            self.own_signs.append(sg)
            if n:
                f32 = self.ps.ps_lookup(0, sg, np.full(n, self.dim, np.uint32), True).reshape(n, self.dim)
                rows[src, :n] = torch.from_numpy(self.o.f32_to_f16(f32 + np.float32(0)).view(np.float16).reshape(n, self.dim))
        back = self._a2a(rows.view(-1)).view(R * cap, self.dim)       # rows returned to this requester, by owner
        out = torch.zeros((S * B, self.dim), dtype=torch.float16)
        for i in range(S * B):
            t = target[occ_item[i]]
            if t >= 0:
                out[i] = back[t]
        self._pending = (items, target, B)
        return out.view(S, B, self.dim).numpy()

    def backward(self, grads, skip=None):
        """grads: f16 [S, B, dim].  Returns the slot status list (0 applied / 1 skipped / 2 NaN)."""
        items, target, B = self._pending
        S, R, cap = len(self.pf), self.R, self.cap
        status = []
        for s in range(S):
            if skip is not None and skip[s]:
                status.append(1)
            elif np.isnan(grads[s].astype(np.float32)).any():
                status.append(2)
            else:
                status.append(0)
        gsend = torch.zeros((R, cap, self.dim), dtype=torch.float32)
        ok = torch.zeros((R, cap), dtype=torch.int32)
        for u, (sign, slot, occ) in enumerate(items):
            if target[u] < 0 or status[slot]:
                continue
            acc = np.zeros(self.dim, np.float32)
            for i in occ:                                             # ascending sample order, f32, +-inf clamped
                acc = acc + np.clip(grads[slot, i - slot * B].astype(np.float32), -65504.0, 65504.0)
            q, k = divmod(int(target[u]), cap)
            gsend[q, k] = torch.from_numpy(acc)
            ok[q, k] = 1
        grecv = self._a2a(gsend.view(-1)).view(R, cap, self.dim)
        okrecv = self._a2a(ok.view(-1)).view(R, cap)
        for src in range(R):                                          # update_gradient_mixed, one request after another
            sg = self.own_signs[src]
            m = okrecv[src, :sg.size].numpy().astype(bool)
            if m.any():
                self.ps.ps_update(0, sg[m], np.full(int(m.sum()), self.dim, np.uint32), grecv[src, :sg.size].numpy()[m])
        return status
