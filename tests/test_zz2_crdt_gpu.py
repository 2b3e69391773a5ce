"""The replicated counter value on the GPU (include/rl_crdt.h) through the C-ABI, against oracle/crdt_oracle.c: the same
random sessions tests/test_crdt.py runs through the kernels under the host shim.  Sorted last on purpose."""
import numpy as np
import pytest

from limitador_b200 import crdt as CR
from oracle.crdt_binding import CrdtOracle
from tests.test_crdt import SEC, T0, _keys, _random_session

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,actors,self_actor", [(11, 2, 0), (12, 3, 2), (13, 16, 7), (14, 1, 0)])
def test_gpu_sessions_match_the_oracle(seed, actors, self_actor):
    t = CR.CrdtTable(2048, actors, self_actor)
    _random_session(t, CrdtOracle(actors, self_actor), seed, actors, steps=60, space=600)
    assert t.kernel_launches() > 60


def test_large_batches_match_the_oracle():
    """65 536 increments and 65 536 gossiped updates per call (keys repeat inside the merge batches)."""
    rng = np.random.default_rng(21)
    actors = 4
    t, o = CR.CrdtTable(1 << 18, actors, 1), CrdtOracle(actors, 1)
    space = 100_000
    now = T0
    for rnd in range(3):
        now += 700_000
        ks = _keys(rng, 65536, space)
        actor = rng.integers(0, actors, size=len(ks)).astype(np.uint32)
        inc = rng.integers(1, 9, size=len(ks)).astype(np.uint64)
        t.inc(CR.keys_array(ks), actor, inc, 2 * SEC, now)
        for k, a, i in zip(ks, actor, inc):
            o.inc_actor_at(k, int(a), int(i), 2 * SEC, now)
        base = _keys(rng, 20000, space)
        ups = []
        for j in rng.integers(0, len(base), size=65536):
            ups.append((base[int(j)], now + int(rng.choice([-1, SEC, 3 * SEC])), {int(rng.integers(0, actors)): int(rng.integers(0, 50))}))
        t.merge(*CR.pack_updates(ups), now)
        for k, exp, vals in ups:
            o.merge_at(k, exp, vals, now)
        probe = _keys(rng, 5000, space)
        val, exp = t.read(CR.keys_array(probe), now)
        assert val.tolist() == [o.read_at(k, now) for k in probe]
        assert exp.tolist() == [o.expiry(k) for k in probe]
    assert t.dump(cap=1 << 18) == o.dump(cap=1 << 18)
    assert t.export(now, cap=1 << 18) == o.export(now, cap=1 << 18)


def test_two_gpu_replicas_converge_through_export_and_merge():
    a, b = CR.CrdtTable(4096, 2, 0), CR.CrdtTable(4096, 2, 1)
    rng = np.random.default_rng(8)
    keys = _keys(rng, 800, 5000)
    now = T0
    for rnd in range(5):
        now += 300_000
        for t, me in ((a, 0), (b, 1)):
            ks = [keys[int(j)] for j in rng.choice(len(keys), size=300, replace=False)]
            t.inc(CR.keys_array(ks), me, rng.integers(1, 5, size=len(ks)).astype(np.uint64), 60 * SEC, now)
        for src, dst, me in ((a, b, 0), (b, a, 1)):
            ups = [((lo, hi), exp, {me: val}) for lo, hi, val, exp in src.export(now)]
            dst.merge(*CR.pack_updates(ups), now)
        va, _ = a.read(CR.keys_array(keys), now)
        vb, _ = b.read(CR.keys_array(keys), now)
        assert va.tolist() == vb.tolist()
    assert int(va.sum()) > 3000


def test_errors_are_loud():
    t = CR.CrdtTable(8, 2, 0)
    with pytest.raises(CR.CrdtError, match="actor"):
        t.inc(CR.keys_array([(1, 1)]), 2, 1, SEC, T0)
    with pytest.raises(CR.CrdtError, match="key"):
        t.inc(CR.keys_array([(0, 0)]), 0, 1, SEC, T0)
    with pytest.raises(CR.CrdtError) as ei:
        t.inc(CR.keys_array([(i + 1, 5) for i in range(9)]), 0, 1, SEC, T0)
    assert ei.value.status == 1  # a full table is TRANSIENT, never a dropped update
    t.clear()  # CounterStorage::clear: the table is empty again and takes new counters
    assert t.dump() == []
    t.inc(CR.keys_array([(i + 1, 5) for i in range(8)]), 1, 3, SEC, T0)
    assert t.read(CR.keys_array([(8, 5)]), T0)[0].tolist() == [3] and len(t.dump()) == 8
