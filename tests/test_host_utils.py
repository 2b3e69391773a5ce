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


def test_balanced_namespace_ids_place_namespaces_through_the_ids_they_get():
    """exchange.balanced_namespace_ids: the static namespace -> GPU placement is an id assignment (owner =
    rl_owner_of(ns_id)), heaviest namespaces first on the least loaded rank."""
    import torch
    from limitador_b200 import exchange
    rng = np.random.default_rng(3)
    for world in (2, 4, 8):
        n = 64 * world
        # the weak-scaled C2 stream (bench.py): key rank ~ Zipf(1.1) over 1 M x world rows, namespace = rank % n
        pmf = 1.0 / (1 + np.arange(1_000_000 * world, dtype=np.float64)) ** 1.1
        load = np.bincount(np.arange(len(pmf)) % n, weights=pmf / pmf.sum(), minlength=n)
        ids, owner_load = exchange.balanced_namespace_ids(load, world)
        assert len(set(ids.tolist())) == n and ids.min() >= 0 and ids.max() < 16 * n
        got = np.zeros(world)
        for j, i in enumerate(ids):
            got[exchange.owner_of(int(i), world)] += load[j]
        assert np.allclose(got, owner_load)
        hashed = np.zeros(world)
        for j in range(n):
            hashed[exchange.owner_of(j, world)] += load[j]
        assert got.max() / got.mean() < 1.01 < 1.05 < hashed.max() / hashed.mean(), (world, got, hashed)
        if world == 8:
            assert hashed.max() / hashed.mean() > 1.5  # one owner holds the Zipf head AND its hash share of the rest
        ids2, _ = exchange.balanced_namespace_ids(load, world)
        assert ids.tolist() == ids2.tolist()
        # the records follow: ns_id replaced, hits_addend (the high half of word 0) untouched
        recs = torch.zeros((3, 50, 4), dtype=torch.int64)
        ns = torch.from_numpy(rng.integers(0, n, size=(3, 50)))
        hits = torch.from_numpy(rng.integers(1, 1000, size=(3, 50)))
        recs[:, :, 0] = ns | (hits << 32)
        recs[:, :, 1] = 7
        exchange.remap_namespace_ids(recs, torch.from_numpy(ids))
        assert torch.equal(recs[:, :, 0] & 0xFFFFFFFF, torch.from_numpy(ids)[ns]) and torch.equal(recs[:, :, 0] >> 32, hits)
        assert int(recs[:, :, 1].min()) == 7


def test_place_namespaces_rewrites_records_and_limits_consistently():
    """bench.place_namespaces on a weak-scaled C2 stream (world 4): the placement is balanced, every rank would compute the
    same ids, and nothing observable changes — the oracle gives the same verdicts on (remapped records, remapped limits)."""
    import torch
    import bench
    from limitador_b200 import exchange, streams
    from oracle import binding as ob
    world, batch, steps = 4, 4096, 6
    n_ns, n_rows = 64 * world, 200_000 * world
    limits = bench.c2_limits(n_ns)
    recs = streams.c2_device_stream(steps, batch, "cpu", n_rows=n_rows, n_ns=n_ns)
    before = recs.clone()

    class OneRank:  # the all-reduce of a single rank
        @staticmethod
        def all_reduce(t):
            return t

    new_limits, summary = bench.place_namespaces(OneRank, world, torch.device("cpu"), recs, limits, steps, lambda m: None)
    assert summary["owner_load_max_over_mean"] < 1.02 < summary["owner_load_max_over_mean_if_ids_were_hashed_as_generated"]
    assert torch.equal(recs[:, :, 1:], before[:, :, 1:]) and torch.equal(recs[:, :, 0] >> 32, before[:, :, 0] >> 32)
    ids = np.zeros(n_ns, dtype=np.int64)
    ids[limits["ns_id"]] = new_limits["ns_id"]
    assert torch.equal(recs[:, :, 0] & 0xFFFFFFFF, torch.from_numpy(ids)[before[:, :, 0] & 0xFFFFFFFF])
    load = np.zeros(world)
    for ns, c in zip(*np.unique((recs[:, :, 0] & 0xFFFFFFFF).numpy(), return_counts=True)):
        load[exchange.owner_of(int(ns), world)] += c
    assert load.max() / load.mean() < 1.03

    def verdicts(lim, r):
        o = ob.Oracle(1 << 16)
        for d in lim:
            o.limit_set(int(d["limit_id"]), int(d["ns_id"]), int(d["max_value"]), int(d["window_us"]), bool(d["qualified"]))
        return np.concatenate([o.batch_records(0, r[s].numpy().view(RECORD_DTYPE).reshape(-1))[0] for s in range(steps)])

    a, b = verdicts(limits, before), verdicts(new_limits, recs)
    assert a.tolist() == b.tolist() and 0 < int(a.sum()) < len(a)
