"""Host-side helpers that carry data across the C-ABI: the 16-byte record packing and the order-independent
table digest bench.py uses for its N>1 parity leg (CPU only)."""
import numpy as np
import pytest

from limitador_b200.engine import RECORD16_DTYPE, RECORD_DTYPE, pack_records16


def test_pack_records16_layout_matches_the_header():
    """include/rl_engine.h rl_record16: word0 = ns_id (0..23) | hits_addend (24..31) | key_hi (32..63), word1 = key_lo."""
    r = np.zeros(3, dtype=RECORD_DTYPE)
    r["ns_id"] = [1, 0xABCDEF, 3]
    r["hits_addend"] = [1, 2, 255]
    r["key_lo"] = [5, 6, 2**64 - 1]
    r["key_hi"] = [9, 0xFFFFFFFF, (0x7F << 56) | 4]  # the lane byte (bits 56..63) is not part of the key
    r["now_us"] = 123
    p = pack_records16(r)
    assert p.dtype == RECORD16_DTYPE and p.itemsize == 16
    assert [hex(int(x)) for x in p["ns_hits_keyhi"]] == ["0x901000001", "0xffffffff02abcdef", "0x4ff000003"]
    assert p["key_lo"].tolist() == [5, 6, 2**64 - 1]


@pytest.mark.parametrize("field,value", [("ns_id", 1 << 24), ("hits_addend", 256), ("key_hi", 1 << 32)])
def test_pack_records16_refuses_what_does_not_fit(field, value):
    r = np.zeros(2, dtype=RECORD_DTYPE)
    r[field][1] = value
    with pytest.raises(ValueError):
        pack_records16(r)


def test_table_digest_is_order_independent_and_content_sensitive():
    import bench
    rng = np.random.default_rng(1)
    n = 1000
    lid = rng.integers(0, 50, n).astype(np.uint32)
    lo, hi = rng.integers(0, 1 << 60, n).astype(np.uint64), rng.integers(0, 1 << 32, n).astype(np.uint64)
    val, exp = rng.integers(0, 100, n).astype(np.uint64), rng.integers(1, 1 << 50, n).astype(np.uint64)
    a = bench.table_digest(lid, lo, hi, val, exp)
    perm = rng.permutation(n)
    assert bench.table_digest(lid[perm], lo[perm], hi[perm], val[perm], exp[perm]) == a
    val2 = val.copy()
    val2[17] += 1
    assert bench.table_digest(lid, lo, hi, val2, exp) != a
    assert bench.table_digest(lid[:0], lo[:0], hi[:0], val[:0], exp[:0])[0] == 0
    exp2 = exp.copy()
    exp2[3], exp2[4] = exp[4], exp[3]  # swapping a field between two rows is seen
    assert (exp[3] == exp[4]) or bench.table_digest(lid, lo, hi, val, exp2) != a
