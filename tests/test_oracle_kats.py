"""Pins the CPU oracle against the reference's own known-answer tests.

The reference ships no golden-vector files; its unit tests with explicit timestamps
(limitador/src/storage/atomic_expiring_value.rs:175-245) and its behaviour tests are the
pins.  Each test cites the reference test it restates (paths under /root/reference/).
SURVEY.md Appendix A.1 traces (hand-derived from the same code) are included as KAT-n.
"""
import numpy as np
import pytest

from oracle import binding as ob

S = 1_000_000
T = 1_700_000_000_000_000


def one(limit_id=0, lo=7, hi=0):
    return ob.counters([(limit_id, lo, hi)])


def fresh(max_value=10, seconds=10, qualified=True):
    o = ob.Oracle()
    o.limit_set(0, 0, max_value, seconds * S, qualified)
    return o


def entry(o, limit_id=0):
    return [(v, e) for (l, _, _, v, e) in o.dump() if l == limit_id]


# --- atomic_expiring_value.rs unit tests ------------------------------------------------
def _seed_value(o, value, expiry):
    """AtomicExpiringValue::new(value, expiry): build it through the public ops."""
    # update at (expiry - window) on a missing counter creates (value, expiry)
    o.update_counters(one(), value, expiry - 10 * S)
    assert entry(o) == [(value, expiry)]


def test_returns_value_when_valid():
    """atomic_expiring_value.rs:181-186 — new(42, now).value_at(now - 1s) == 42."""
    o = fresh(max_value=100)
    _seed_value(o, 42, T)
    # value_at(T - 1s) == 42  <=>  within limits for delta 58, over for 59
    assert o.is_within_limits(one(), 58, T - S)
    assert not o.is_within_limits(one(), 59, T - S)


def test_returns_default_when_expired():
    """:188-193 — new(42, now - 1s).value_at(now) == 0."""
    o = fresh(max_value=100)
    _seed_value(o, 42, T - S)
    assert o.is_within_limits(one(), 100, T)
    assert not o.is_within_limits(one(), 101, T)


def test_returns_default_on_expiry():
    """:195-200 — expiry == now reads as 0 (inclusive bound, :76-79)."""
    o = fresh(max_value=100)
    _seed_value(o, 42, T)
    assert o.is_within_limits(one(), 100, T)
    assert not o.is_within_limits(one(), 59, T - 1)  # one µs earlier the 42 still counts


def test_updates_when_valid():
    """:202-208 — new(42, now+1s).update(3, 10s, now) -> 45, expiry kept."""
    o = fresh(max_value=100)
    _seed_value(o, 42, T + S)
    o.update_counters(one(), 3, T)
    assert entry(o) == [(45, T + S)]


def test_updates_when_expired():
    """:210-217 — new(42, now): ttl 0; update(3, 10s, now) -> value 3, expiry now+10s."""
    o = fresh(max_value=100)
    _seed_value(o, 42, T)
    _, _, _, ttl = o.check_and_update(one(), 0, True, T)  # reported ttl is the pre-update one
    assert ttl[0] == 0
    o2 = fresh(max_value=100)
    _seed_value(o2, 42, T)
    o2.update_counters(one(), 3, T)
    assert entry(o2) == [(3, T + 10 * S)]


def test_overlapping_updates_sequential_orders():
    """:219-237 — two racing updates end in {2, 3}; both sequential orders land in that set."""
    for order in ((0, 1), (1, 0)):
        o = ob.Oracle()
        o.limit_set(0, 0, 100, 1 * S, True)
        o.update_counters(one(), 42, T - 1 * S + 10 * S)  # (42, T + 10s)
        assert entry(o) == [(42, T + 10 * S)]
        ops = [(1, T), (2, T + 11 * S)]
        for k in order:
            o.update_counters(one(), ops[k][0], ops[k][1])
        assert entry(o)[0][0] in (2, 3)


def test_size_of_struct():
    """:239-244 — 16 bytes of state per counter (value u64 + expiry u64): the GPU cell size."""
    assert np.dtype([("value", "<u8"), ("expiry", "<u8")]).itemsize == 16


# --- in_memory.rs / lib.rs / integration tests ------------------------------------------
def test_counters_for_multiple_limit_per_ns():
    """in_memory.rs:277-310 — same ns/conditions/variables, different seconds => 2 counters."""
    o = ob.Oracle()
    o.limit_set(0, 0, 1, 1 * S, True)
    o.limit_set(1, 0, 1, 10 * S, True)
    o.update_counters(one(0), 1, T)
    o.update_counters(one(1), 1, T)
    assert len(o.dump()) == 2


def test_rate_limited():
    """tests/integration_tests.rs:493-532 — max 3: three (check, update) rounds then limited."""
    o = fresh(max_value=3, seconds=60)
    for i in range(3):
        assert o.is_rate_limited(one(), 1, T + i)[0] is False
        o.update_counters(one(), 1, T + i)
    assert o.is_rate_limited(one(), 1, T + 3)[0] is True


def test_rate_limited_with_delta_higher_than_one():
    """:656-695 — 5 + 5 of 10, then delta 1 is limited."""
    o = fresh(max_value=10, seconds=60)
    for i in range(2):
        assert not o.is_rate_limited(one(), 5, T + i)[0]
        o.update_counters(one(), 5, T + i)
    assert o.is_rate_limited(one(), 1, T + 2)[0]


def test_rate_limited_with_delta_higher_than_max():
    """:697-722 — delta 11 > max 10 on an empty counter => limited; no state created."""
    o = fresh(max_value=10, seconds=60)
    assert o.is_rate_limited(one(), 11, T)[0]
    assert o.dump() == []


def test_kat3_check_and_update_creates_entry_even_when_limited():
    """SURVEY A.1 KAT-3 / in_memory.rs:122-127 — the lookup inserts (0, now+W) before the verdict."""
    o = fresh(max_value=10, seconds=60)
    limited, idx, _, _ = o.check_and_update(one(), 11, False, T)
    assert limited and idx == 0
    assert entry(o) == [(0, T + 60 * S)]


def test_check_rate_limited_and_update():
    """:843-879 — max 3: three allowed, fourth limited."""
    o = fresh(max_value=3, seconds=60)
    for i in range(3):
        assert not o.check_and_update(one(), 1, False, T + i)[0]
    assert o.check_and_update(one(), 1, False, T + 3)[0]
    assert entry(o) == [(3, T + 60 * S)]


def test_check_rate_limited_and_update_load_counters():
    """:881-929 + KAT-1 — remaining 2,1,0 then limited with remaining 0; ttl pre-update."""
    o = fresh(max_value=3, seconds=60)
    expect = [(False, 2, 60 * S), (False, 1, 59 * S), (False, 0, 58 * S), (True, 0, 57 * S)]
    for i, (lim, rem, ttl) in enumerate(expect):
        limited, _, r, t = o.check_and_update(one(), 1, True, T + i * S)
        assert (limited, int(r[0]), int(t[0])) == (lim, rem, ttl)
    assert entry(o) == [(3, T + 60 * S)]
    # KAT-1 rows 5-6: expiry == now reads as 0, window re-anchored at now
    limited, _, r, t = o.check_and_update(one(), 1, True, T + 60 * S)
    assert (limited, int(r[0]), int(t[0])) == (False, 2, 0)
    assert entry(o) == [(1, T + 120 * S)]
    limited, _, r, t = o.check_and_update(one(), 1, True, T + 60 * S + 1)
    assert (limited, int(r[0]), int(t[0])) == (False, 1, 60 * S - 1)


def test_kat2_unqualified_counter():
    """SURVEY A.1 KAT-2 — unqualified limit pre-created (0, EPOCH) by add_counter
    (in_memory.rs:38-44, atomic_expiring_value.rs:151-158)."""
    o = ob.Oracle()
    o.limit_set(0, 0, 2, 10 * S, False)
    assert o.dump() == [(0, 0, 0, 0, 0)]
    c = ob.counters([(0, 0, 0)])
    limited, _, r, t = o.check_and_update(c, 1, True, T)
    assert (limited, int(r[0]), int(t[0])) == (False, 1, 0)
    limited, _, r, t = o.check_and_update(c, 1, True, T + S)
    assert (limited, int(r[0]), int(t[0])) == (False, 0, 9 * S)
    limited, _, r, t = o.check_and_update(c, 1, True, T + 2 * S)
    assert (limited, int(r[0]), int(t[0])) == (True, 0, 8 * S)
    assert o.dump() == [(0, 0, 0, 2, T + 10 * S)]


def test_kat4_all_or_nothing_across_limits():
    """KAT-4 / in_memory.rs:141-153 — a limited request increments nothing."""
    o = ob.Oracle()
    o.limit_set(0, 0, 1, 60 * S, True)
    o.limit_set(1, 0, 5, 60 * S, True)
    c = ob.counters([(0, 7, 0), (1, 7, 0)])
    assert not o.check_and_update(c, 1, False, T)[0]
    limited, idx, _, _ = o.check_and_update(c, 1, False, T + S)
    assert limited and idx == 0
    assert o.dump() == [(0, 7, 0, 1, T + 60 * S), (1, 7, 0, 1, T + 60 * S)]


def test_kat5_denied_check_anchors_window_order_dependent():
    """KAT-5 / in_memory.rs:110-112,122-127,130-132 — which counters a denied request creates
    depends on the counter order and on load_counters."""
    def run(order, load_counters):
        o = ob.Oracle()
        o.limit_set(0, 0, 0, 60 * S, True)   # A: max 0
        o.limit_set(1, 0, 10, 60 * S, True)  # B: max 10
        both = ob.counters([(l, 7, 0) for l in order])
        assert o.check_and_update(both, 1, load_counters, T)[0]            # r1 denied by A
        assert not o.check_and_update(ob.counters([(1, 7, 0)]), 1, False, T + 30 * S)[0]  # r2 only B
        return o.check_and_update(ob.counters([(1, 7, 0)]), 10, False, T + 70 * S)[0], o

    lim, o = run((1, 0), False)   # B looked up (created) before A limits
    assert lim is False and (1, 7, 0, 10, T + 130 * S) in o.dump()
    lim, o = run((0, 1), True)    # load_counters: no early return, B created
    assert lim is False
    lim, o = run((0, 1), False)   # early return at A: B NOT created, anchored by r2 at T+30s
    assert lim is True and (1, 7, 0, 1, T + 90 * S) in o.dump()


def test_unqualified_processed_before_qualified():
    """in_memory.rs:105 then :121 — simple counters are examined first whatever the order."""
    o = ob.Oracle()
    o.limit_set(0, 0, 0, 60 * S, True)    # qualified, max 0 (limited)
    o.limit_set(1, 0, 0, 60 * S, False)   # unqualified, max 0 (limited)
    c = ob.counters([(0, 7, 0), (1, 0, 0)])
    limited, idx, _, _ = o.check_and_update(c, 1, False, T)
    assert limited and idx == 1           # the unqualified one is named
    assert [d for d in o.dump() if d[0] == 0] == []  # qualified counter never reached


def test_kat6_max_value_change_under_live_counter():
    """lib.rs:760-790 / storage/mod.rs:67-83 — max_value 42 -> 50 keeps the counter."""
    o = fresh(max_value=42, seconds=60)
    _, _, r, _ = o.check_and_update(one(), 1, True, T)
    assert int(r[0]) == 41
    o.limit_set(0, 0, 50, 60 * S, True)
    _, _, r, _ = o.check_and_update(one(), 1, True, T + 1)
    assert int(r[0]) == 48


def test_delete_and_readd_resets_qualified_counter():
    """lib.rs:792-817 — delete + re-add => remaining == 41 again."""
    o = fresh(max_value=42, seconds=60)
    o.check_and_update(one(), 1, True, T)
    o.limit_delete(0)
    o.limit_set(0, 0, 42, 60 * S, True)
    _, _, r, _ = o.check_and_update(one(), 1, True, T + 1)
    assert int(r[0]) == 41


def test_kat7_intra_batch_duplicates():
    """KAT-7 — a batch is the sequential composition of its requests."""
    o = fresh(max_value=3, seconds=60)
    off = np.arange(5, dtype=np.uint32)
    ctrs = ob.counters([(0, 7, 0)] * 4)
    lim, _, _, _ = o.batch_csr(0, off, ctrs, [1] * 4, [T] * 4)
    assert lim.tolist() == [0, 0, 0, 1] and entry(o) == [(3, T + 60 * S)]
    o = fresh(max_value=4, seconds=60)
    lim, _, _, _ = o.batch_csr(0, off[:4], ctrs[:3], [3, 2, 1], [T] * 3)
    assert lim.tolist() == [0, 1, 0] and entry(o) == [(4, T + 60 * S)]


def test_get_counters():
    """tests/integration_tests.rs:989-1039 — remaining 9 and 5 after hits; :1073-1100 expired
    counters are not returned (ttl > 0 filter, in_memory.rs:167-169)."""
    o = ob.Oracle()
    o.limit_set(0, 0, 10, 60 * S, True)
    o.limit_set(1, 0, 10, 1 * S, True)
    o.update_counters(ob.counters([(0, 1, 0)]), 1, T)
    o.update_counters(ob.counters([(0, 2, 0)]), 5, T)
    o.update_counters(ob.counters([(1, 1, 0)]), 1, T)
    got = o.get_counters([0, 1], T + 100)
    assert [(g[0], g[1], g[3]) for g in got] == [(0, 1, 9), (0, 2, 5), (1, 1, 9)]
    got = o.get_counters([0], T + 2 * S)  # limit 1's counter expired
    assert [(g[0], g[1], g[3], g[4]) for g in got] == [(0, 1, 9, 58 * S), (0, 2, 5, 58 * S)]


def test_delete_counters_and_clear():
    """in_memory.rs:189-201,241-257 — delete by limit; clear() drops only unqualified."""
    o = ob.Oracle()
    o.limit_set(0, 0, 10, 60 * S, True)
    o.limit_set(1, 0, 10, 60 * S, False)
    o.update_counters(ob.counters([(0, 1, 0), (1, 0, 0)]), 1, T)
    assert len(o.dump()) == 2
    o.clear()
    assert o.dump() == [(0, 1, 0, 1, T + 60 * S)]
    o.delete_counters([0])
    assert o.dump() == []


def test_invalidate_expired_event():
    """Oracle-only mirror of the sweep kernel: drops qualified entries with expiry <= now;
    a swept entry and an expired-but-present one differ for a later denied check."""
    o = fresh(max_value=1, seconds=1)
    o.update_counters(one(), 1, T)
    assert o.invalidate_expired(T + S - 1) == 0
    assert o.invalidate_expired(T + S) == 1
    assert o.dump() == []


def test_empty_counter_list_is_not_limited():
    """lib.rs:434-440."""
    o = fresh()
    lim, fl, _, _ = o.batch_csr(0, np.array([0, 0], dtype=np.uint32), ob.counters([]), [1], [T])
    assert lim.tolist() == [0] and fl.tolist() == [ob.NONE]


def test_records_match_csr():
    """lo_batch_records is lo_batch_csr with the namespace's limits in registration order."""
    descs = [(0, 0, 5, 1 * S, True), (1, 0, 100, 60 * S, True), (2, 1, 2, 10 * S, False), (3, 1, 3, 10 * S, True)]
    rng = np.random.default_rng(1)
    n = 400
    recs = np.zeros(n, dtype=ob.RECORD_DTYPE)
    recs["ns_id"] = rng.integers(0, 3, size=n)
    recs["hits_addend"] = rng.integers(1, 3, size=n)
    recs["key_lo"] = rng.integers(1, 4, size=n)
    recs["now_us"] = T + np.cumsum(rng.integers(0, 300_000, size=n))
    a, b = ob.Oracle(), ob.Oracle()
    for d in descs:
        a.limit_set(*d)
        b.limit_set(*d)
    lim_a, fl_a, rem_a, ttl_a = a.batch_records(0, recs, True, 2)
    by_ns = {0: [0, 1], 1: [2, 3], 2: []}
    off, ctrs = [0], []
    for r in recs:
        for l in by_ns[int(r["ns_id"])]:
            q = descs[l][4]
            ctrs.append((l, int(r["key_lo"]) if q else 0, 0))
        off.append(len(ctrs))
    lim_b, fl_b, rem_b, ttl_b = b.batch_csr(0, off, ob.counters(ctrs), recs["hits_addend"], recs["now_us"], True)
    assert lim_a.tolist() == lim_b.tolist() and fl_a.tolist() == fl_b.tolist()
    assert a.dump() == b.dump()
    for i in range(n):
        m = off[i + 1] - off[i]
        assert rem_a[2 * i:2 * i + m].tolist() == rem_b[off[i]:off[i + 1]].tolist()
        assert ttl_a[2 * i:2 * i + m].tolist() == ttl_b[off[i]:off[i + 1]].tolist()


def test_mt_baseline_matches_single_thread():
    """The multi-threaded CPU baseline shards by namespace and must give the same verdicts."""
    descs = np.array([(k, k // 2, 3 + k, (1 + k) * S, 1, 0) for k in range(8)], dtype=ob.LIMIT_DESC_DTYPE)
    rng = np.random.default_rng(3)
    n = 20000
    recs = np.zeros(n, dtype=ob.RECORD_DTYPE)
    recs["ns_id"] = rng.integers(0, 4, size=n)
    recs["hits_addend"] = 1
    recs["key_lo"] = rng.integers(1, 50, size=n)
    recs["now_us"] = T + np.arange(n) * 500
    o = ob.Oracle()
    for d in descs:
        o.limit_set(int(d["limit_id"]), int(d["ns_id"]), int(d["max_value"]), int(d["window_us"]), True)
    want, _, _, _ = o.batch_records(0, recs)
    for threads in (1, 3):
        _, got = ob.bench_records_mt(descs, recs, threads, 1024)
        assert got.tolist() == want.tolist()


def test_results_do_not_depend_on_table_growth():
    """The oracle's own hash table is an implementation detail: a run that starts from a tiny table
    (and rehashes many times, also in the middle of a request that already holds counters) must give
    the outputs and the final state of a run that never grows.  (Regression: entry pointers kept
    across a rehash lost updates.)"""
    from tests import helpers as H
    from tests.test_gpu_parity import single_row_limits
    for seed, cells in ((9, 3), (3, 7), (5, 1)):
        descs = single_row_limits(cells, n_ns=9, seed=seed)
        runs = []
        for cap in (16, 1 << 17):
            o = H.oracle_with_limits(descs, cap)
            outs = []
            for b in range(3):
                recs = H.random_records(descs, 4000, seed * 100 + b, n_keys=400, monotone=bool(b & 1))
                outs.append([x.copy() for x in o.batch_records(0, recs, True, cells)])
            runs.append((outs, H.normalise_dump(o.dump(), descs)))
        assert runs[0][1] == runs[1][1]
        for a, b in zip(runs[0][0], runs[1][0]):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
