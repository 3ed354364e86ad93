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
