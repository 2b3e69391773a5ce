"""The reference's behaviour tests (limitador/tests/integration_tests.rs, lib.rs tests,
envoy_rls/server.rs header tests) restated over the Python mirror of RateLimiter, run
against two CounterStorage implementors: the CPU oracle (pins the mirror + oracle) and
the GPU engine through the C-ABI (`-m gpu`).  File:line citations are under
/root/reference/."""
import pytest

from limitador_b200 import Context, Limit, RateLimiter
from tests import helpers as H

S = 1_000_000


class Clock:
    def __init__(self):
        self.t = H.T0

    def __call__(self):
        return self.t

    def advance(self, us):
        self.t += us


def make_oracle():
    return H.OracleStorage()


def make_gpu():
    from limitador_b200 import Engine, GpuCounterStorage
    return GpuCounterStorage(Engine(capacity_rows=4096, cells_per_row=3, max_batch=1024))


BACKENDS = [
    pytest.param(make_oracle, id="oracle"),
    pytest.param(make_gpu, id="gpu", marks=pytest.mark.gpu),
]


@pytest.fixture(params=BACKENDS)
def rl(request):
    limiter = RateLimiter(request.param(), clock=Clock())
    return limiter


def ctx(**kw):
    return Context(kw)


GET = dict(req_method="GET", app_id="test_app_id")


def test_rate_limited(rl):
    """integration_tests.rs:493-532."""
    limit = Limit("test_namespace", 3, 60, ["req_method == 'GET'"], ["app_id"])
    rl.add_limit(limit)
    for i in range(3):
        assert not rl.is_rate_limited("test_namespace", ctx(**GET), 1).limited, f"Must not be limited after {i}"
        rl.update_counters("test_namespace", ctx(**GET), 1)
    assert rl.is_rate_limited("test_namespace", ctx(**GET), 1).limited


def test_rate_limited_id_counter(rl):
    """integration_tests.rs:534-574 — same flow with a limit carrying an id."""
    limit = Limit("test_namespace", 3, 60, ["req_method == 'GET'"], ["app_id"], id="test-rate_limited_id_counter")
    rl.add_limit(limit)
    for _ in range(3):
        assert not rl.is_rate_limited("test_namespace", ctx(**GET), 1).limited
        rl.update_counters("test_namespace", ctx(**GET), 1)
    assert rl.is_rate_limited("test_namespace", ctx(**GET), 1).limited


def test_multiple_limits_rate_limited(rl):
    """integration_tests.rs:576-654 — per-limit isolation (GET limit 3, POST limit 4)."""
    ns = "test_namespace"
    rl.add_limit(Limit(ns, 3, 60, ["req_method == 'GET'"], ["app_id"]))
    rl.add_limit(Limit(ns, 4, 60, ["req_method == 'POST'"], ["app_id"]))
    get, post = ctx(**GET), ctx(req_method="POST", app_id="test_app_id")
    for i in range(3):
        assert not rl.is_rate_limited(ns, get, 1).limited
        assert not rl.is_rate_limited(ns, post, 1).limited
        rl.check_rate_limited_and_update(ns, get, 1, False)
        rl.check_rate_limited_and_update(ns, post, 1, False)
    assert rl.is_rate_limited(ns, get, 1).limited
    assert not rl.is_rate_limited(ns, post, 1).limited


def test_rate_limited_with_delta_higher_than_one(rl):
    """integration_tests.rs:656-695."""
    rl.add_limit(Limit("test_namespace", 10, 60, ["req_method == 'GET'"], ["app_id"]))
    for _ in range(2):
        assert not rl.is_rate_limited("test_namespace", ctx(**GET), 5).limited
        rl.update_counters("test_namespace", ctx(**GET), 5)
    assert rl.is_rate_limited("test_namespace", ctx(**GET), 1).limited


def test_rate_limited_with_delta_higher_than_max(rl):
    """integration_tests.rs:697-722."""
    rl.add_limit(Limit("test_namespace", 10, 60, ["req_method == 'GET'"], ["app_id"]))
    assert rl.is_rate_limited("test_namespace", ctx(**GET), 11).limited


def test_takes_into_account_only_vars_of_the_limits(rl):
    """integration_tests.rs:724-769 — extra context values do not create new counters."""
    rl.add_limit(Limit("test_namespace", 3, 60, ["req_method == 'GET'"], ["app_id"]))
    for i in range(3):
        c = ctx(does_not_apply=str(i), **GET)
        assert not rl.is_rate_limited("test_namespace", c, 1).limited
        rl.update_counters("test_namespace", c, 1)
    assert rl.is_rate_limited("test_namespace", ctx(does_not_apply="3", **GET), 1).limited


def test_is_rate_limited_returns_false_when_no_limits_in_namespace(rl):
    """integration_tests.rs:771-778."""
    assert not rl.is_rate_limited("test_namespace", ctx(**GET), 1).limited


def test_is_rate_limited_returns_false_when_no_matching_limits(rl):
    """integration_tests.rs:780-815 — the condition does not match."""
    rl.add_limit(Limit("test_namespace", 0, 60, ["req_method == 'GET'"], ["app_id"]))
    assert not rl.is_rate_limited("test_namespace", ctx(req_method="POST", app_id="x"), 1).limited


def test_is_rate_limited_applies_limit_if_its_unconditional(rl):
    """integration_tests.rs:817-841 — max 0, no conditions => limited."""
    rl.add_limit(Limit("test_namespace", 0, 60, [], ["app_id"]))
    assert rl.is_rate_limited("test_namespace", ctx(app_id="test_app_id"), 1).limited


def test_check_rate_limited_and_update(rl):
    """integration_tests.rs:843-879."""
    rl.add_limit(Limit("test_namespace", 3, 60, ["req_method == 'GET'"], ["app_id"]))
    for _ in range(3):
        assert not rl.check_rate_limited_and_update("test_namespace", ctx(**GET), 1, False).limited
    assert rl.check_rate_limited_and_update("test_namespace", ctx(**GET), 1, False).limited


def test_check_rate_limited_and_update_load_counters(rl):
    """integration_tests.rs:881-929 — remaining 2,1,0 then limited with remaining 0; len 1; ttl <= 60."""
    rl.add_limit(Limit("test_namespace", 3, 60, ["req_method == 'GET'"], ["app_id"]))
    for hit in range(3):
        res = rl.check_rate_limited_and_update("test_namespace", ctx(**GET), 1, True)
        assert not res.limited and len(res.counters) == 1
        for c in res.counters:
            assert c.expires_in_secs() <= 60
            assert c.remaining == 3 - (hit + 1)
        rl.clock.advance(S)
    res = rl.check_rate_limited_and_update("test_namespace", ctx(**GET), 1, True)
    assert res.limited and len(res.counters) == 1
    assert res.counters[0].remaining == 0 and res.counters[0].expires_in_secs() <= 60


def test_check_rate_limited_and_update_returns_true_if_no_limits_apply(rl):
    """integration_tests.rs:931-959 — name says true, body asserts NOT limited, no state."""
    rl.add_limit(Limit("test_namespace", 10, 60, ["req_method == 'POST'"], ["app_id"]))
    res = rl.check_rate_limited_and_update("test_namespace", ctx(**GET), 1, False)
    assert not res.limited
    assert rl.get_counters("test_namespace") == set()


def test_check_rate_limited_and_update_applies_limit_if_its_unconditional(rl):
    """integration_tests.rs:961-987."""
    rl.add_limit(Limit("test_namespace", 0, 60, [], ["app_id"]))
    assert rl.check_rate_limited_and_update("test_namespace", ctx(app_id="test_app_id"), 1, False).limited


def test_get_counters(rl):
    """integration_tests.rs:989-1039 — remaining 9 (1 hit) and 5 (5 hits) on two limits."""
    ns = "test_namespace"
    l1 = Limit(ns, 10, 60, ["req_method == 'GET'"], ["app_id"])
    l2 = Limit(ns, 10, 60, ["req_method == 'POST'"], ["app_id"])
    rl.add_limit(l1)
    rl.add_limit(l2)
    rl.update_counters(ns, ctx(**GET), 1)
    rl.update_counters(ns, ctx(req_method="POST", app_id="test_app_id"), 5)
    rl.clock.advance(1000)
    counters = rl.get_counters(ns)
    assert len(counters) == 2
    for c in counters:
        assert c.expires_in_secs() <= 60
        assert c.set_variables == {"app_id": "test_app_id"}
        assert c.remaining == (9 if c.limit == l1 else 5)


def test_get_counters_returns_empty_when_no_limits_or_counters(rl):
    """integration_tests.rs:1041-1071."""
    assert rl.get_counters("test_namespace") == set()
    rl.add_limit(Limit("test_namespace", 10, 60, ["req_method == 'GET'"], ["app_id"]))
    assert rl.get_counters("test_namespace") == set()


def test_get_counters_does_not_return_expired_ones(rl):
    """integration_tests.rs:1073-1100 — after limit_time + 1 s the counter is gone."""
    rl.add_limit(Limit("test_namespace", 10, 1, ["req_method == 'GET'"], ["app_id"]))
    rl.update_counters("test_namespace", ctx(**GET), 1)
    rl.clock.advance(2 * S)
    assert rl.get_counters("test_namespace") == set()


def test_delete_limit_also_deletes_associated_counters(rl):
    """integration_tests.rs:367-394."""
    limit = Limit("test_namespace", 10, 60, ["req_method == 'GET'"], ["app_id"])
    rl.add_limit(limit)
    rl.update_counters("test_namespace", ctx(**GET), 1)
    rl.delete_limit(limit)
    assert rl.get_counters("test_namespace") == set()
    assert rl.get_limits("test_namespace") == set()


def test_delete_limits_of_a_namespace_also_deletes_counters(rl):
    """integration_tests.rs:434-462."""
    rl.add_limit(Limit("test_namespace", 5, 60, ["req_method == 'GET'"], ["app_id"]))
    rl.update_counters("test_namespace", ctx(**GET), 1)
    rl.delete_limits("test_namespace")
    assert rl.get_counters("test_namespace") == set()
    assert "test_namespace" not in rl.get_namespaces()


def test_add_limit_only_adds_if_not_present(rl):
    """integration_tests.rs:1250-1284 — same identity, different max: the first one stays."""
    ns = "test_namespace"
    assert rl.add_limit(Limit(ns, 10, 60, ["req_method == 'GET'"], ["app_id"]))
    assert not rl.add_limit(Limit(ns, 20, 60, ["req_method == 'GET'"], ["app_id"]))
    assert [l.max_value for l in rl.get_limits(ns)] == [10]


def test_configure_with_keeps_the_given_limits_and_counters_if_they_exist(rl):
    """integration_tests.rs:1135-1177 — counters of kept limits survive configure_with."""
    ns = "test_namespace"
    limit = Limit(ns, 10, 60, ["req_method == 'GET'"], ["app_id"])
    rl.add_limit(limit)
    rl.update_counters(ns, ctx(**GET), 1)
    rl.configure_with([limit, Limit(ns, 5, 60, ["req_method == 'POST'"], ["app_id"])])
    assert len(rl.get_limits(ns)) == 2
    (c,) = rl.get_counters(ns)
    assert c.remaining == 9


def test_configure_with_deletes_all_except_the_limits_given(rl):
    """integration_tests.rs:1179-1212."""
    ns = "test_namespace"
    a = Limit(ns, 10, 60, ["req_method == 'GET'"], ["app_id"])
    b = Limit(ns, 20, 60, ["req_method == 'POST'"], ["app_id"])
    rl.add_limit(a)
    rl.add_limit(b)
    rl.update_counters(ns, ctx(req_method="POST", app_id="x"), 1)
    rl.configure_with([a])
    assert rl.get_limits(ns) == {a}
    assert rl.get_counters(ns) == set()


def test_configure_with_updates_the_limits_max_value_under_live_counter(rl):
    """lib.rs:760-790 + integration_tests.rs:1214-1248 — 42 -> 50 keeps the counter."""
    ns = "test_namespace"
    limit = Limit(ns, 42, 60, ["req_method == 'GET'"], ["app_id"])
    rl.add_limit(limit)
    res = rl.check_rate_limited_and_update(ns, ctx(**GET), 1, True)
    assert res.counters[0].remaining == 41
    rl.configure_with([limit.with_max_value(50)])
    res = rl.check_rate_limited_and_update(ns, ctx(**GET), 1, True)
    assert res.counters[0].remaining == 48 and res.counters[0].max_value() == 50


def test_delete_and_readd_limit_resets_counter(rl):
    """lib.rs:792-817 — remaining == 41 again after delete + add."""
    ns = "test_namespace"
    limit = Limit(ns, 42, 60, ["req_method == 'GET'"], ["app_id"])
    rl.add_limit(limit)
    rl.check_rate_limited_and_update(ns, ctx(**GET), 1, True)
    rl.delete_limit(limit)
    rl.add_limit(limit)
    res = rl.check_rate_limited_and_update(ns, ctx(**GET), 1, True)
    assert res.counters[0].remaining == 41


def test_unqualified_limit_add_limit_without_vars(rl):
    """integration_tests.rs:281-297 + KAT-2 — a limit without variables counts all requests."""
    ns = "test_namespace"
    rl.add_limit(Limit(ns, 2, 10, ["req_method == 'GET'"], []))
    r = rl.check_rate_limited_and_update(ns, ctx(req_method="GET"), 1, True)
    assert (r.limited, r.counters[0].remaining, r.counters[0].expires_in_us) == (False, 1, 0)
    rl.clock.advance(S)
    r = rl.check_rate_limited_and_update(ns, ctx(req_method="GET"), 1, True)
    assert (r.limited, r.counters[0].remaining, r.counters[0].expires_in_us) == (False, 0, 9 * S)
    rl.clock.advance(S)
    r = rl.check_rate_limited_and_update(ns, ctx(req_method="GET"), 1, True)
    assert (r.limited, r.counters[0].remaining, r.counters[0].expires_in_us) == (True, 0, 8 * S)


# --- Envoy RLS header strings (limitador-server/src/envoy_rls/server.rs tests) ------------
def rls_ctx(**entries):
    return Context({}, descriptors=[dict(entries)])


def test_rls_headers_single_limit(rl):
    """envoy_rls/server.rs:337-426 — "1, 1;w=60", Remaining "0", on allow and on deny."""
    ns = "test_namespace"
    rl.add_limit(Limit(ns, 1, 60, ["descriptors[0].req_method == 'GET'"], ["descriptors[0].app_id"]))
    c = rls_ctx(req_method="GET", app_id="1")
    res = rl.check_rate_limited_and_update(ns, c, 1, True)
    assert not res.limited
    h = res.response_header()
    assert h["X-RateLimit-Limit"] == "1, 1;w=60" and h["X-RateLimit-Remaining"] == "0"
    res = rl.check_rate_limited_and_update(ns, c, 1, True)
    assert res.limited
    h = res.response_header()
    assert h["X-RateLimit-Limit"] == "1, 1;w=60" and h["X-RateLimit-Remaining"] == "0"
    assert int(h["X-RateLimit-Reset"]) <= 60


def test_rls_headers_two_limits_one_zero(rl):
    """envoy_rls/server.rs:496-591 — limits max 10 and max 0 => OverLimit, "0, 0;w=60, 10;w=60"."""
    ns = "test_namespace"
    rl.add_limit(Limit(ns, 10, 60, ["descriptors[0].x == '1'"], ["descriptors[0].z"]))
    rl.add_limit(Limit(ns, 0, 60, ["descriptors[0].x == '1'", "descriptors[0].y == '2'"], ["descriptors[0].z"]))
    res = rl.check_rate_limited_and_update(ns, rls_ctx(x="1", y="2", z="1"), 1, True)
    assert res.limited
    h = res.response_header()
    assert h["X-RateLimit-Limit"] == "0, 0;w=60, 10;w=60" and h["X-RateLimit-Remaining"] == "0"


def test_rls_hits_addend(rl):
    """envoy_rls/server.rs:593-680 — addend 6 of max 10 => Remaining "4", then OverLimit "0"."""
    ns = "test_namespace"
    rl.add_limit(Limit(ns, 10, 60, ["descriptors[0].req_method == 'GET'"], ["descriptors[0].app_id"]))
    c = rls_ctx(req_method="GET", app_id="1")
    res = rl.check_rate_limited_and_update(ns, c, 6, True)
    assert not res.limited and res.response_header()["X-RateLimit-Remaining"] == "4"
    res = rl.check_rate_limited_and_update(ns, c, 6, True)
    assert res.limited and res.response_header()["X-RateLimit-Remaining"] == "0"


def test_limit_name_is_reported(rl):
    """in_memory.rs:91-94,99-101 — Authorization::Limited carries the limit's name."""
    ns = "test_namespace"
    rl.add_limit(Limit(ns, 0, 60, [], ["app_id"], name="zero"))
    res = rl.check_rate_limited_and_update(ns, ctx(app_id="a"), 1, False)
    assert res.limited and res.limit_name == "zero"


def test_batched_front_equals_one_by_one(rl):
    """The batching front (one kernel pipeline) gives the sequential answers (KAT-7)."""
    ns = "test_namespace"
    rl.add_limit(Limit(ns, 3, 60, ["req_method == 'GET'"], ["app_id"]))
    rl.add_limit(Limit(ns, 5, 60, [], []))
    reqs = [(ns, ctx(**GET), 1)] * 4 + [(ns, ctx(req_method="POST", app_id="b"), 2)] * 2
    res = rl.check_rate_limited_and_update_batch(reqs, True)
    assert [r.limited for r in res] == [False, False, False, True, False, True]
    assert [c.remaining for c in res[2].counters] == [0, 2] or sorted(c.remaining for c in res[2].counters) == [0, 2]
