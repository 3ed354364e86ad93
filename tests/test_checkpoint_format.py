"""CPU: the `.emb` checkpoint format (persia_b200/checkpoint.py) against bytes laid out by hand from the reference's
(de)serialisation code — persia-speedy (little endian, usize = u64, u32 length prefixes), ArrayLinkedList
(array_linked_list.rs:137-272) and HashMapEmbeddingEntry (emb_entry.rs:17-25).  PARITY UNPINNED: no reference
fixture exists and the reference cannot run here; these vectors are this repo's reading of that code."""
import struct

import numpy as np
import pytest

from persia_b200 import checkpoint as CK


def _entry(inner, dim, sign):
    return struct.pack("<I", len(inner)) + struct.pack(f"<{len(inner)}f", *inner) + struct.pack("<QQ", dim, sign)


def test_two_entries_hand_layout():
    # push_back(A); push_back(B) on an empty list (array_linked_list.rs:215-272):
    # nodes [A: next 2, prev 0] [B: next 0, prev 1]; first 1, last 2, free 0, end 0, count 2
    a = _entry([0.5, -1.25, 2.0], 2, 0x0100000000000007)   # dim 2 + 1 float of optimizer state
    b = _entry([1.0, 3.0, 0.125], 2, 0x0100000000000009)
    want = (struct.pack("<QIIII", 2, 1, 2, 0, 0) + struct.pack("<I", 2) +
            struct.pack("<IIB", 2, 0, 1) + a + struct.pack("<IIB", 0, 1, 1) + b)
    signs = np.array([0x0100000000000007, 0x0100000000000009], np.uint64)
    ent = np.array([[0.5, -1.25, 2.0], [1.0, 3.0, 0.125]], np.float32)
    assert CK.encode_list(signs, ent, 2) == want                      # vectorised writer
    assert CK.encode_list(signs, [ent[0], ent[1]], [2, 2]) == want    # per-entry writer
    s, d, e = CK.decode_list(want)
    assert s.tolist() == signs.tolist() and d.tolist() == [2, 2]
    np.testing.assert_array_equal(np.stack(e), ent)


def test_reference_shaped_file_with_free_nodes_and_scattered_order():
    # with_capacity(4): four deleted nodes chained 2,3,4,0 (fill_elements, :229-241); then push_back(X) takes node 1,
    # push_back(Y) node 2, push_front(Z) node 3 -> list order Z, X, Y; node 4 stays free (flag 0, next 0)
    X, Y, Z = _entry([1.0], 1, 11), _entry([2.0, 2.5], 1, 22), _entry([3.0], 1, 33)
    buf = (struct.pack("<QIIII", 3, 3, 2, 4, 4) + struct.pack("<I", 4) +
           struct.pack("<IIB", 2, 3, 1) + X +      # node 1 = X: next Y(2), prev Z(3)
           struct.pack("<IIB", 0, 1, 1) + Y +      # node 2 = Y: last
           struct.pack("<IIB", 1, 0, 1) + Z +      # node 3 = Z: first
           struct.pack("<IIB", 0, 0, 0))           # node 4: unused
    s, d, e = CK.decode_list(buf)
    assert s.tolist() == [33, 11, 22]
    assert [x.tolist() for x in e] == [[3.0], [1.0], [2.0, 2.5]]  # entry lengths may differ (optimizer state)


def test_empty_list_and_round_trip_large():
    empty = CK.encode_list(np.zeros(0, np.uint64), np.zeros((0, 8), np.float32), 4)
    assert empty == struct.pack("<QIIII", 0, 0, 0, 0, 0) + struct.pack("<I", 0)
    s, d, e = CK.decode_list(empty)
    assert s.size == 0 and e == []
    rng = np.random.default_rng(3)
    signs = rng.integers(0, 2**63, size=5000, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    ent = rng.standard_normal((5000, 128)).astype(np.float32)
    buf = CK.encode_list(signs, ent, 64)
    assert len(buf) == 28 + 5000 * (9 + 4 + 128 * 4 + 16)
    s, d, e = CK.decode_list(buf)
    np.testing.assert_array_equal(s, signs)
    assert (d == 64).all()
    np.testing.assert_array_equal(np.stack(e), ent)


@pytest.mark.parametrize("cut", [10, 27, 40, -1])
def test_damaged_files_are_rejected(cut):
    signs = np.array([1, 2, 3], np.uint64)
    buf = CK.encode_list(signs, np.ones((3, 2), np.float32), 2)
    with pytest.raises(ValueError):
        CK.decode_list(buf[:cut] if cut > 0 else buf + b"\x00")


def test_broken_links_are_rejected():
    buf = bytearray(CK.encode_list(np.array([1, 2], np.uint64), [np.ones(1, np.float32), np.ones(2, np.float32)], 1))
    struct.pack_into("<I", buf, 28, 1)  # node 1 now points at itself
    with pytest.raises(ValueError):
        CK.decode_list(bytes(buf))


def test_done_marker_yaml():
    import yaml

    info = yaml.safe_load(CK.done_yaml(2, 128, now=1700000000.25))
    assert info == {"num_shards": 2, "num_internal_shards": 128,
                    "datetime": {"secs_since_epoch": 1700000000, "nanos_since_epoch": 250000000}}


class _ListModel:
    """ArrayLinkedList as the reference implements it (array_linked_list.rs:215-330, 422-520): 1-based node indices,
    0 = none, a free list threaded through `next_index` of unused nodes."""

    def __init__(self, capacity):
        self.count, self.first, self.last, self.free, self.end = 0, 0, 0, 0, 0
        self.nodes = []  # [next, prev, data]
        if capacity:  # fill_elements
            for i in range(1, capacity):
                self.nodes.append([i + 1, 0, None])
            self.nodes.append([0, 0, None])
            self.free, self.end = 1, capacity

    def _take(self, node):
        if self.free == 0:
            self.nodes.append(node)
            return len(self.nodes)
        idx = self.free
        self.free = self.nodes[idx - 1][0]
        self.nodes[idx - 1] = node
        return idx

    def push_back(self, data):
        idx = self._take([0, self.last, data])
        if self.last:
            self.nodes[self.last - 1][0] = idx
        else:
            self.first = idx
        self.last = idx
        self.count += 1
        return idx

    def push_front(self, data):
        idx = self._take([self.first, 0, data])
        if self.first:
            self.nodes[self.first - 1][1] = idx
        else:
            self.last = idx
        self.first = idx
        self.count += 1
        return idx

    def remove(self, idx):
        nxt, prv, data = self.nodes[idx - 1]
        if prv:
            self.nodes[prv - 1][0] = nxt
        else:
            self.first = nxt
        if nxt:
            self.nodes[nxt - 1][1] = prv
        else:
            self.last = prv
        self.nodes[idx - 1] = [self.free, 0, None]  # back on the free list
        self.free = idx
        self.count -= 1
        return data

    def order(self):
        out, cur = [], self.first
        while cur:
            out.append(self.nodes[cur - 1][2])
            cur = self.nodes[cur - 1][0]
        return out

    def to_bytes(self):
        buf = struct.pack("<QIIII", self.count, self.first, self.last, self.free, self.end) + struct.pack("<I", len(self.nodes))
        for nxt, prv, data in self.nodes:
            buf += struct.pack("<IIB", nxt, prv, 1 if data is not None else 0)
            if data is not None:
                buf += _entry(*data)
        return buf


@pytest.mark.parametrize("seed,capacity", [(0, 0), (1, 8), (2, 50), (3, 200)])
def test_decoder_follows_any_list_the_reference_can_produce(seed, capacity):
    """An LRU holder's life — inserts at the back, refreshes (remove + push_back), evictions from the front, the odd
    push_front — replayed on a model of the reference's list, serialised node by node, decoded, compared."""
    rng = np.random.default_rng(seed)
    m, where, next_sign = _ListModel(capacity), {}, 1
    for _ in range(600):
        r = rng.random()
        if r < 0.45 or not where:
            inner = rng.standard_normal(int(rng.integers(1, 6))).astype(np.float32).tolist()
            data = (inner, int(rng.integers(1, 4)), next_sign)
            where[next_sign] = m.push_back(data) if rng.random() < 0.9 else m.push_front(data)
            next_sign += 1
        elif r < 0.8:  # get_refresh: move to the back (eviction_map.rs:48-60)
            sign = int(rng.choice(list(where)))
            where[sign] = m.push_back(m.remove(where[sign]))
        else:  # evict the least recently used (eviction_map.rs:76-97)
            data = m.remove(m.first)
            del where[data[2]]
    want = m.order()
    signs, dims, entries = CK.decode_list(m.to_bytes())
    assert signs.tolist() == [d[2] for d in want]
    assert dims.tolist() == [d[1] for d in want]
    for e, d in zip(entries, want):
        assert e.tolist() == d[0]


class _FakeShard:
    """Stands in for EmbeddingShard on the CPU: records what load_shards hands to set_entries."""

    def __init__(self, dim, entry_len):
        import torch

        self.dim, self.entry_len, self.device = dim, entry_len, torch.device("cpu")
        self.got = {}

    def set_entries(self, signs, ent):
        for s, e in zip(signs.numpy().view(np.uint64).tolist(), ent.numpy()):
            self.got[s] = e.copy()


def test_load_reshards_when_the_shard_count_changed(tmp_path, oracle):
    """A checkpoint written by 3 parameter servers loaded by 2 (and by 1): every entry lands on the replica that
    farmhash64(sign) % replica_size names (embedding_worker_service/mod.rs:1150-1259), nothing is lost or duplicated;
    the same count loads shard by shard."""
    from persia_b200 import checkpoint as CK

    rng = np.random.default_rng(5)
    dim, L, n_old = 4, 8, 3
    signs = rng.integers(1, 1 << 60, size=500, dtype=np.uint64)
    ent = rng.standard_normal((signs.size, L)).astype(np.float32)
    old_owner = oracle.shard_of(signs, n_old)
    for r in range(n_old):
        d = tmp_path / f"s{r}"
        d.mkdir()
        m = old_owner == r
        (d / f"replica_{r}_shard_0.emb").write_bytes(CK.encode_list(signs[m], ent[m], dim))
        (d / CK.DONE_FILE).write_text(CK.done_yaml(n_old, 1))
    (tmp_path / CK.DONE_FILE).write_text(CK.done_yaml(n_old, 1))
    assert CK.checkpoint_info(str(tmp_path)) == (n_old, 1)
    for new_size in (2, 1):
        seen = {}
        for r in range(new_size):
            sh = _FakeShard(dim, L)
            n = CK.load_shards(str(tmp_path), {dim: sh}, replica_index=r, replica_size=new_size)
            assert n == len(sh.got)
            want_owner = oracle.shard_of(np.array(sorted(sh.got), np.uint64), new_size)
            assert (want_owner == r).all()
            seen.update(sh.got)
        assert len(seen) == signs.size
        for s, e in zip(signs.tolist(), ent):
            assert np.array_equal(seen[s], e)
    sh = _FakeShard(dim, L)  # same shard count: replica 1 reads s1 only
    assert CK.load_shards(str(tmp_path), {dim: sh}, replica_index=1, replica_size=3) == int((old_owner == 1).sum())
    assert set(sh.got) == set(signs[old_owner == 1].tolist())
