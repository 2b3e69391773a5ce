"""Known-answer tests of the counter arithmetic on the GPU, through the C-ABI — the reference's own
`AtomicExpiringValue` unit tests (limitador/src/storage/atomic_expiring_value.rs:175-245) and the edge cases
of SURVEY.md §8 rows a8/a9, with the expected numbers written out (not taken from the oracle):
inclusive expiry bound (expiry == now reads as 0), pre-update ttl, reset on expiry, u64 wrap of value + delta,
intra-batch duplicates (Appendix A.1 KAT-1, KAT-2, KAT-7).  Every case runs as separate calls AND as one batch
(the batch must equal request-by-request execution), on 32-B and 128-B rows."""
import numpy as np
import pytest

from limitador_b200 import Engine
from limitador_b200.engine import COUNTER_DTYPE, LIMIT_DESC_DTYPE, NONE

pytestmark = pytest.mark.gpu
S = 1_000_000
T = 1_700_000_000_000_000
U64 = (1 << 64) - 1


def eng(cells, max_value=100, seconds=10, qualified=1):
    e = Engine(capacity_rows=1 << 10, cells_per_row=cells, max_batch=1 << 12, regions=1)
    e.limits_set(np.array([(0, 0, 1 if qualified else 0, qualified, max_value, seconds * S)], dtype=LIMIT_DESC_DTYPE))
    return e


def ctr(n=1):
    return np.array([(0, 0, 7, 0)] * n, dtype=COUNTER_DTYPE)


def off(n):
    return np.arange(n + 1, dtype=np.uint32)


def entry(e):
    return [(v, x) for (l, _, _, v, x) in e.dump() if l == 0]


def seed(e, value, expiry):
    """AtomicExpiringValue::new(value, expiry): an update at (expiry - window) on a missing counter."""
    e.update_batch(off(1), ctr(), [value], [expiry - 10 * S])
    assert entry(e) == [(value, expiry)]


@pytest.mark.parametrize("cells", [1, 7])
def test_returns_value_when_valid(cells):
    """atomic_expiring_value.rs:181-186 — new(42, now).value_at(now - 1s) == 42."""
    e = eng(cells)
    seed(e, 42, T)
    lim, first = e.is_within_limits_batch(off(2), ctr(2), [58, 59], [T - S, T - S])
    assert lim.tolist() == [0, 1] and first.tolist() == [NONE, 0]


@pytest.mark.parametrize("cells", [1, 7])
def test_returns_default_when_expired_and_on_expiry(cells):
    """:188-200 — new(42, now - 1s).value_at(now) == 0; expiry == now reads as 0 too (inclusive bound, :76-79);
    one µs earlier the 42 still counts."""
    e = eng(cells)
    seed(e, 42, T - S)
    assert e.is_within_limits_batch(off(2), ctr(2), [100, 101], [T, T])[0].tolist() == [0, 1]
    e = eng(cells)
    seed(e, 42, T)
    assert e.is_within_limits_batch(off(3), ctr(3), [100, 59, 58], [T, T - 1, T - 1])[0].tolist() == [0, 1, 0]


@pytest.mark.parametrize("cells", [1, 7])
def test_updates_when_valid_and_when_expired(cells):
    """:202-217 — new(42, now+1s).update(3, 10s, now) -> 45, expiry kept; new(42, now): ttl 0 before,
    update(3, 10s, now) -> value 3, expiry now + 10 s."""
    e = eng(cells)
    seed(e, 42, T + S)
    e.update_batch(off(1), ctr(), [3], [T])
    assert entry(e) == [(45, T + S)]
    e = eng(cells)
    seed(e, 42, T)
    lim, first, rem, ttl = e.check_and_update_batch(off(1), ctr(), [0], [T], load_counters=True)
    assert lim.tolist() == [0] and ttl.tolist() == [0] and rem.tolist() == [100]  # pre-update ttl, value read as 0
    e = eng(cells)
    seed(e, 42, T)
    e.update_batch(off(1), ctr(), [3], [T])
    assert entry(e) == [(3, T + 10 * S)]


@pytest.mark.parametrize("cells", [1, 7])
def test_overlapping_updates_in_one_batch(cells):
    """:219-237 — the two racing updates of the reference end in {2, 3}; a batch applies them in array order,
    i.e. one of the two sequential orders, for both orders."""
    for order, want in (((0, 1), 2), ((1, 0), 3)):
        e = eng(cells, seconds=1)
        e.update_batch(off(1), ctr(), [42], [T + 9 * S])  # (42, T + 10 s)
        assert entry(e) == [(42, T + 10 * S)]
        ops = [(1, T), (2, T + 11 * S)]
        e.update_batch(off(2), ctr(2), [ops[k][0] for k in order], [ops[k][1] for k in order])
        assert entry(e)[0][0] == want


@pytest.mark.parametrize("cells", [1, 7])
def test_value_plus_delta_wraps_like_a_release_build(cells):
    """The reference adds with wrapping u64 arithmetic in release builds: the check
    `value + delta > max` is done on the wrapped sum, and update_counters stores the wrapped sum."""
    e = eng(cells, max_value=U64)
    e.update_batch(off(1), ctr(), [U64 - 2], [T])
    assert entry(e) == [(U64 - 2, T + 10 * S)]
    lim, _, rem, _ = e.check_and_update_batch(off(3), ctr(3), [1, 1, 5], [T + 1, T + 2, T + 3], load_counters=True)
    # U64-2 + 1 = U64-1 (allowed, remaining 1); + 1 = U64 (allowed, remaining 0); + 5 wraps to 4 <= max: allowed
    assert lim.tolist() == [0, 0, 0] and rem.tolist() == [1, 0, U64 - 4]
    assert entry(e) == [(4, T + 10 * S)]


@pytest.mark.parametrize("cells", [1, 7])
@pytest.mark.parametrize("batched", [False, True])
def test_kat1_single_qualified_limit(cells, batched):
    """SURVEY Appendix A.1 KAT-1 (agrees with tests/integration_tests.rs:881-929): max 3 / 60 s, delta 1."""
    e = eng(cells, max_value=3, seconds=60)
    nows = [T, T + S, T + 2 * S, T + 3 * S, T + 60 * S, T + 60 * S + 1]
    want = [(0, 2, 60 * S), (0, 1, 59 * S), (0, 0, 58 * S), (1, 0, 57 * S), (0, 2, 0), (0, 1, 60 * S - 1)]
    if batched:
        lim, first, rem, ttl = e.check_and_update_batch(off(6), ctr(6), [1] * 6, nows, load_counters=True)
        got = list(zip(lim.tolist(), rem.tolist(), ttl.tolist()))
        assert first.tolist() == [NONE, NONE, NONE, 0, NONE, NONE]
    else:
        got = []
        for t in nows:
            lim, _, rem, ttl = e.check_and_update_batch(off(1), ctr(), [1], [t], load_counters=True)
            got.append((int(lim[0]), int(rem[0]), int(ttl[0])))
    assert got == want
    assert entry(e) == [(2, T + 120 * S)]


@pytest.mark.parametrize("cells", [1, 7])
def test_kat2_unqualified_counter_is_precreated_at_epoch(cells):
    """KAT-2: an unqualified limit exists as (0, EPOCH) from add_limit on (in_memory.rs:38-44): ttl 0 on first use."""
    e = eng(cells, max_value=2, seconds=10, qualified=0)
    lim, _, rem, ttl = e.check_and_update_batch(off(3), ctr(3), [1] * 3, [T, T + S, T + 2 * S], load_counters=True)
    assert lim.tolist() == [0, 0, 1] and rem.tolist() == [1, 0, 0] and ttl.tolist() == [0, 9 * S, 8 * S]
    assert entry(e) == [(2, T + 10 * S)]


@pytest.mark.parametrize("cells", [1, 7])
def test_kat7_intra_batch_duplicates(cells):
    """KAT-7: a batch is its requests one at a time, in order — [k,k,k,k] with max 3: allow x3, deny; deltas
    [3,2,1] with max 4: allow, deny, allow (greedy, not a prefix sum)."""
    e = eng(cells, max_value=3, seconds=60)
    assert e.check_and_update_batch(off(4), ctr(4), [1] * 4, [T] * 4)[0].tolist() == [0, 0, 0, 1]
    assert entry(e) == [(3, T + 60 * S)]
    e = eng(cells, max_value=4, seconds=60)
    assert e.check_and_update_batch(off(3), ctr(3), [3, 2, 1], [T] * 3)[0].tolist() == [0, 1, 0]
    assert entry(e) == [(4, T + 60 * S)]
    # a delta above max on an empty counter: limited, and the entry is created (0, now + W) (KAT-3)
    e = eng(cells, max_value=10, seconds=60)
    assert e.check_and_update_batch(off(1), ctr(), [11], [T])[0].tolist() == [1]
    assert entry(e) == [(0, T + 60 * S)]
