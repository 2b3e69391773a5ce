"""The replicated counter value (include/rl_crdt.h).

1. The CPU restatement (oracle/crdt_oracle.c) against the reference's own eleven unit tests
   (limitador/src/storage/distributed/cr_counter_value.rs:177-300), ported with explicit clocks.
2. The kernels of rl_crdt.cuh run on the host under tests/emu/cuda_shim.h (the same source the GPU compiles) against that
   oracle on random inc / merge / read / export streams — so the kernel logic is checked in the GPU-less container.
The GPU run of the same streams through the C-ABI is tests/test_zz2_crdt_gpu.py."""
import numpy as np
import pytest

from limitador_b200 import crdt as CR
from oracle.crdt_binding import CrdtOracle
from tests import helpers as H

T0 = 1_700_000_000_000_000
SEC = 1_000_000
K = (7, 1 << 32)  # one counter key
A, B = 0, 1       # the tests' actors 'A' and 'B'
U64 = 2 ** 64


def new(actor, window_us, when=T0, actors=2):
    """CrCounterValue::new(actor, u64::MAX, window) at `when`: a replica holding one counter created at that instant (the
    store creates a counter and increments it under one clock reading, distributed/mod.rs:65-91)."""
    o = CrdtOracle(actors, actor)
    o.created = (window_us, when)
    return o


def other_set(o, when):
    """other.into_inner() (:125-137): (expiry, {actor: value}) including the replica's own value under its own name."""
    d = o.dump()
    if not d:  # never incremented: the reference's new() holds expiry = created + window and a zero value
        return o.created[1] + o.created[0], {o.self_actor: 0}
    _, _, exp, vals = d[0]
    return exp, {a: v for a, v in enumerate(vals) if v or a == o.self_actor}


def test_local_increments_are_readable():  # :181-188
    a = new(A, SEC)
    a.inc_at(K, 3, SEC, T0)
    assert a.read_at(K, T0) == 3
    a.inc_at(K, 2, SEC, T0)
    assert a.read_at(K, T0) == 5


def test_local_increments_expire():  # :190-199
    a = new(A, SEC)
    a.inc_at(K, 3, SEC, T0)
    assert a.read_at(K, T0) == 3
    a.inc_at(K, 2, SEC, T0 + SEC)  # expiry == when: expired (inclusive), the window restarts
    assert a.read_at(K, T0 + SEC) == 2


def test_other_increments_are_readable():  # :201-208
    a = new(A, SEC)
    a.inc_actor_at(K, B, 3, SEC, T0)
    assert a.read_at(K, T0) == 3
    a.inc_actor_at(K, B, 2, SEC, T0)
    assert a.read_at(K, T0) == 5


def test_other_increments_expire():  # :210-219
    a = new(A, SEC)
    a.inc_actor_at(K, B, 3, SEC, T0)
    assert a.read_at(K, T0) == 3
    a.inc_actor_at(K, B, 2, SEC, T0 + SEC)
    assert a.read_at(K, T0 + SEC) == 2


def _merge(dst, src, when):
    exp, vals = other_set(src, when)
    dst.merge_at(K, exp, vals, when)


def test_merges():  # :221-230
    a, b = new(A, SEC), new(B, SEC)
    a.inc_at(K, 3, SEC, T0)
    b.inc_at(K, 2, SEC, T0)
    _merge(a, b, T0)
    assert a.read_at(K, T0) == 5


def test_merges_symetric():  # :232-241
    a, b = new(A, SEC), new(B, SEC)
    a.inc_at(K, 3, SEC, T0)
    b.inc_at(K, 2, SEC, T0)
    _merge(b, a, T0)
    assert b.read_at(K, T0) == 5


def test_merges_overrides_with_larger_value():  # :243-253
    a, b = new(A, SEC), new(B, SEC)
    a.inc_at(K, 3, SEC, T0)
    b.inc_at(K, 2, SEC, T0)
    b.inc_actor_at(K, A, 2, SEC, T0)  # older value!
    _merge(b, a, T0)                  # merges the 3
    assert b.read_at(K, T0) == 5


def test_merges_ignore_lesser_values():  # :255-265
    a, b = new(A, SEC), new(B, SEC)
    a.inc_at(K, 3, SEC, T0)
    b.inc_at(K, 2, SEC, T0)
    b.inc_actor_at(K, A, 5, SEC, T0)  # newer value!
    _merge(b, a, T0)                  # ignores the 3 and keeps its own 5 for a
    assert b.read_at(K, T0) == 7


def test_merge_ignores_expired_sets():  # :267-276
    a = new(A, 0)
    a.inc_at(K, 3, 0, T0)
    b = new(B, SEC)
    b.inc_at(K, 2, SEC, T0)
    _merge(b, a, T0)
    assert b.read_at(K, T0) == 2


def test_merge_ignores_expired_sets_symmetric():  # :278-287
    a = new(A, 0)
    a.inc_at(K, 3, 0, T0)
    b = new(B, SEC)
    b.inc_at(K, 2, SEC, T0)
    _merge(a, b, T0)
    assert a.read_at(K, T0) == 2


def test_merge_uses_earliest_expiry():  # :289-299
    later, sooner = SEC, 200_000
    a, b = new(A, later), new(B, sooner)
    a.inc_at(K, 3, later, T0)   # a's window was created with `later`
    b.inc_at(K, 2, sooner, T0)  # b's with `sooner` (the reference passes `later` to inc, which an unexpired value ignores)
    _merge(a, b, T0)
    assert a.expiry(K) - T0 <= sooner and a.expiry(K) > T0
    assert a.read_at(K, T0) == 5


# ---- the kernels under the host shim against the oracle -------------------------------------------------------------------
def _keys(rng, n, space):
    ids = rng.choice(space, size=n, replace=False)
    return [(int(i) * 2654435761 % (1 << 40) + 1, (int(i) % 5 + 1) << 32 | int(i)) for i in ids]


def _random_session(table, oracle, seed, actors, steps=40, space=300, dup_merges=True):
    """Random inc / merge / read / export / dump rounds; every inc batch holds a key once (exact), merge batches repeat keys."""
    rng = np.random.default_rng(seed)
    now = T0
    for step in range(steps):
        now += int(rng.choice([0, 1, 400_000, 1_100_000, 5_000_000]))
        kind = rng.random()
        if kind < 0.45:
            ks = _keys(rng, int(rng.integers(1, 60)), space)
            actor = rng.integers(0, actors, size=len(ks)).astype(np.uint32)
            inc = rng.choice([1, 1, 2, 7, 2 ** 63], size=len(ks)).astype(np.uint64)
            win = rng.choice([0, SEC, 10 * SEC, 60 * SEC], size=len(ks)).astype(np.uint64)
            assert table.inc(CR.keys_array(ks), actor, inc, win, now) in (0, None)
            for k, a, i, w in zip(ks, actor, inc, win):
                oracle.inc_actor_at(k, int(a), int(i), int(w), now)
        elif kind < 0.85:
            n = int(rng.integers(1, 50))
            ks = _keys(rng, min(n, space), space)
            if dup_merges:
                ks = [ks[int(j)] for j in rng.integers(0, len(ks), size=n)]  # keys repeat inside the batch
            ups = []
            for k in ks:
                exp = now + int(rng.choice([-SEC, 0, 1, SEC // 2, SEC, 30 * SEC]))
                vals = {int(a): int(rng.choice([0, 1, 3, 50, 2 ** 40])) for a in rng.choice(actors, size=int(rng.integers(0, actors + 1)), replace=False)}
                ups.append((k, exp, vals))
            packed = CR.pack_updates(ups)
            assert table.merge(*packed, now) in (0, None)
            for k, exp, vals in ups:
                oracle.merge_at(k, exp, vals, now)
        ks = _keys(rng, 40, space)
        val, exp = table.read(CR.keys_array(ks), now)
        assert val.tolist() == [oracle.read_at(k, now) % U64 for k in ks]
        assert exp.tolist() == [oracle.expiry(k) for k in ks]
        if step % 7 == 0:
            assert table.export(now) == oracle.export(now)
            assert table.dump() == oracle.dump()
    assert table.dump() == oracle.dump()
    assert len(oracle.dump()) > 50


@pytest.mark.parametrize("seed,actors,self_actor", [(1, 2, 0), (2, 3, 2), (3, 5, 1), (4, 16, 15), (5, 1, 0)])
def test_kernels_under_the_host_shim_match_the_oracle(seed, actors, self_actor):
    H.emu_maint_lib().emu_seed(seed * 1000)
    _random_session(H.EmuCrdt(1024, actors, self_actor), CrdtOracle(actors, self_actor), seed, actors)


def test_kernels_under_the_fiber_emulator_match_the_oracle():
    """The same kernels with the threads of a block interleaved as fibers (tests/emu/cuda_simt.h) instead of run one after
    the other: another thread order, the device flavour of the memory helpers' callers."""
    _random_session(H.EmuCrdt(1024, 3, 1, simt=True), CrdtOracle(3, 1), 6, 3, steps=25)


def test_two_replicas_converge_through_export_and_merge():
    """The gossip loop: each replica increments its own values, exports the re-sync stream (distributed/mod.rs:294-332)
    and the peer merges it as CounterUpdate{key, {actor: value}, expires_at} (:236-246): both read the same totals."""
    a, b = H.EmuCrdt(512, 2, 0), H.EmuCrdt(512, 2, 1)
    oa, ob = CrdtOracle(2, 0), CrdtOracle(2, 1)
    rng = np.random.default_rng(8)
    keys = _keys(rng, 80, 500)
    now = T0
    for rnd in range(6):
        now += 300_000
        for t, o, me in ((a, oa, 0), (b, ob, 1)):
            ks = [keys[int(j)] for j in rng.choice(len(keys), size=30, replace=False)]
            inc = rng.integers(1, 5, size=len(ks)).astype(np.uint64)
            t.inc(CR.keys_array(ks), me, inc, 60 * SEC, now)
            for k, i in zip(ks, inc):
                o.inc_at(k, int(i), 60 * SEC, now)
        for src, dst, osrc, odst, me in ((a, b, oa, ob, 0), (b, a, ob, oa, 1)):
            stream = src.export(now)
            assert stream == osrc.export(now)
            ups = [((lo, hi), exp, {me: val}) for lo, hi, val, exp in stream]
            dst.merge(*CR.pack_updates(ups), now)
            for k, exp, vals in ups:
                odst.merge_at(k, exp, vals, now)
        va, _ = a.read(CR.keys_array(keys), now)
        vb, _ = b.read(CR.keys_array(keys), now)
        assert va.tolist() == vb.tolist() == [oa.read_at(k, now) for k in keys]
    assert int(va.sum()) > 300


def test_shim_kernels_report_bad_input_and_a_full_table():
    t = H.EmuCrdt(8, 2, 0)
    assert t.inc(CR.keys_array([(1, 1)]), 2, 1, SEC, T0) == 2            # actor out of range
    assert t.inc(CR.keys_array([(0, 0)]), 0, 1, SEC, T0) == 3            # bad key
    assert t.inc(CR.keys_array([(i + 1, 5) for i in range(9)]), 0, 1, SEC, T0) == 1  # 9 keys, 8 rows
    ups = np.zeros(1, dtype=CR.UPDATE_DTYPE)
    ups[0] = (1, 5, T0 + SEC, 0, 3)
    assert t.merge(ups, np.zeros(1, np.uint32), np.ones(1, np.uint64), T0) == 4   # value range outside the arrays
