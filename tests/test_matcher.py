"""The native CPU front (include/rl_match.h): limits -> counters.

Pinned by the reference's own `Limit::applies` tests (limitador/src/limit.rs:239-348), restated for both
the Python mirror (limitador_b200/limiter.py) and the compiled matcher, and by a randomised differential
test of the two over the accepted expression subset (conditions on root bindings and on descriptors[i],
== and !=, variables, unset names, shared variable sets, deletes and re-adds).  No GPU involved."""
import threading

import numpy as np
import pytest

from limitador_b200 import limiter as LM
from limitador_b200 import matcher as MT
from limitador_b200.engine import COUNTER_DTYPE


def both_apply(conditions, variables, values):
    """-> (python mirror says the limit applies, native matcher yields a counter)"""
    lim = LM.Limit("test_namespace", 10, 60, conditions, variables)
    py = lim.applies(LM.Context(values))
    m = MT.Matcher()
    d = m.add_limit("test_namespace", 10, 60, conditions, variables)
    got = m.counters(int(d["ns_id"]), values)
    assert len(got) in (0, 1)
    return py, len(got) == 1


# limit.rs:239-254, 256-271, 273-289, 291-306, 308-327, 329-348
REFERENCE_KATS = [
    ("limit_applies", ['x == "5"'], ["y"], {"x": "5", "y": "1"}, True),
    ("limit_does_not_apply_when_cond_is_false", ['x == "5"'], ["y"], {"x": "1", "y": "1"}, False),
    ("limit_does_not_apply_when_cond_var_is_not_set", ['x == "5"'], ["y"], {"a": "1", "y": "1"}, False),
    ("limit_does_not_apply_when_var_not_set", ['x == "5"'], ["y"], {"x": "5"}, False),
    ("limit_applies_when_all_its_conditions_apply", ['x == "5"', 'y == "2"'], ["z"], {"x": "5", "y": "2", "z": "1"}, True),
    ("limit_does_not_apply_if_one_cond_doesnt", ['x == "5"', 'y == "2"'], ["z"], {"x": "3", "y": "2", "z": "1"}, False),
]


@pytest.mark.parametrize("name,conds,vars_,values,want", REFERENCE_KATS, ids=[k[0] for k in REFERENCE_KATS])
def test_reference_applies_kats(name, conds, vars_, values, want):
    py, native = both_apply(conds, vars_, values)
    assert py == want and native == want


def test_counter_key_is_blake2b_96_of_the_sorted_pairs():
    for sv in ({"a": "1"}, {"descriptors[0].user": "alice", "z": ""}, {"k" * 200: "v" * 300, "b": "x"}, {"é": "ü"}):
        c = LM.Counter(LM.Limit("ns", 1, 1), dict(sv))  # the digest covers the resolved (source, value) pairs only
        assert MT.counter_key(sv) == c.key()
    assert MT.counter_key({}) == (0, 0)


def test_expressions_outside_the_subset_are_refused_and_change_nothing():
    m = MT.Matcher()
    d0 = m.add_limit("ns", 5, 60, ["a == 'x'"], ["u"])
    for conds, vars_ in ((["foo.contains('bar')"], []), ([], ["int(x) * 3"]), (["a == b"], []), (["a == 'x' && b == 'y'"], []),
                         (["descriptors[0] == 'x'"], []), (["limit.name == null"], [])):
        with pytest.raises(MT.MatcherError) as ei:
            m.add_limit("ns2", 5, 60, conds, vars_)
        assert "unsupported" in str(ei.value)
    assert m.namespace_id("ns2") is None
    d1 = m.add_limit("ns", 7, 60, ["a == 'x'"], ["u"], name="renamed")  # same identity: update_limit
    assert int(d1["limit_id"]) == int(d0["limit_id"]) and int(d1["max_value"]) == 7
    assert m.limit_name(int(d0["limit_id"])) == "renamed"


class _CaptureStorage:
    """what RateLimiter._push_limit hands to the storage (the rl_limit_desc rows)"""

    def __init__(self):
        self.descs = {}

    def set_limit(self, limit_id, ns_id, varset_id, qualified, max_value, window_us):
        self.descs[limit_id] = (limit_id, ns_id, varset_id, int(qualified), max_value, window_us)

    def forget_limit(self, limit_id):
        pass

    def delete_counters(self, ids):
        pass


def random_limit(rng, namespaces, keys, values):
    def operand():
        k = str(rng.choice(keys))
        style = int(rng.integers(0, 4))
        if style == 0:
            return k
        if style == 1:
            return f"descriptors[{int(rng.integers(0, 2))}].{k}"
        if style == 2:
            return f"descriptors[{int(rng.integers(0, 2))}]['{k}']"
        return f"descriptors[{int(rng.integers(0, 2))}]['req.{k}']"  # a dotted key is only reachable through the bracket form
    conds = []
    for _ in range(int(rng.integers(0, 4))):
        op = "==" if rng.random() < 0.7 else "!="
        q = "'" if rng.random() < 0.5 else '"'
        pad = " " * int(rng.integers(0, 3))
        conds.append(f"{pad}{operand()} {op}{pad}{q}{rng.choice(values)}{q}{pad}")
    vars_ = [operand() for _ in range(int(rng.integers(0, 3)))]
    name = None if rng.random() < 0.5 else f"lim{int(rng.integers(0, 1000))}"
    return LM.Limit(str(rng.choice(namespaces)), int(rng.integers(0, 100)), int(rng.choice([1, 60, 3600])), conds, vars_, name)


def random_context(rng, keys, values):
    root = {}
    for k in keys:
        if rng.random() < 0.5:
            root[k] = str(rng.choice(values))
        if rng.random() < 0.3:
            root[f"req.{k}"] = str(rng.choice(values))  # never matched: root operands have no dots
    descriptors = []
    for _ in range(int(rng.integers(0, 3))):
        d = {k: str(rng.choice(values)) for k in keys if rng.random() < 0.5}
        d.update({f"req.{k}": str(rng.choice(values)) for k in keys if rng.random() < 0.3})
        descriptors.append(d)
    return root, descriptors


@pytest.mark.parametrize("seed", range(6))
def test_matcher_equals_the_python_mirror_on_random_limits_and_contexts(seed):
    rng = np.random.default_rng(seed)
    namespaces, keys, values = ["ns_a", "ns_b", "ns_c"], ["m", "u", "path", "k9"], ["GET", "POST", "alice", "", "x y"]
    cap = _CaptureStorage()
    rl = LM.RateLimiter(cap)
    m = MT.Matcher()
    live = []
    for step in range(60):
        r = rng.random()
        if r < 0.75 or not live:
            lim = random_limit(rng, namespaces, keys, values)
            is_new = rl.add_limit(lim)
            if not is_new:
                rl.update_limit(lim)
            d = m.add_limit(lim.namespace, lim.max_value, lim.seconds, lim.conditions, lim.variables, lim.name)
            lid = int(d["limit_id"])
            assert tuple(int(d[f]) for f in ("limit_id", "ns_id", "varset_id", "qualified", "max_value", "window_us")) == cap.descs[lid]
            if is_new:
                live.append(lim)
        else:
            lim = live.pop(int(rng.integers(0, len(live))))
            lid = rl._limit_ids[lim.identity()]
            rl.delete_limit(lim)
            m.delete_limit(lid)
        # a batch of requests against the current limits
        reqs = [(str(rng.choice(namespaces + ["unknown_ns"])),) + random_context(rng, keys, values) for _ in range(25)]
        want_off, want = [0], []
        for ns, root, descs in reqs:
            for c in rl.counters_that_apply(ns, LM.Context(root, descs)):
                want.append((c.limit_id,) + c.key())
            want_off.append(len(want))
        ns_ids = [m.namespace_id(ns) if m.namespace_id(ns) is not None else 0xFFFFFF for ns, _, _ in reqs]
        off, ctrs = m.counters_batch(ns_ids, [(root, descs) for _, root, descs in reqs])
        assert off.tolist() == want_off, f"step {step}"
        assert [(int(c["limit_id"]), int(c["key_lo"]), int(c["key_hi"])) for c in ctrs] == want, f"step {step}"
        # the single-request call agrees with the batch
        j = int(rng.integers(0, len(reqs)))
        one = m.counters(ns_ids[j], reqs[j][1], reqs[j][2])
        assert one.tobytes() == ctrs[off[j]:off[j + 1]].tobytes()
    for lim in live:  # names follow update_limit on both sides
        lid = rl._limit_ids[lim.identity()]
        assert m.limit_name(lid) == rl._limits[lim.namespace][lim].name


def test_limits_sharing_a_variable_set_share_the_counter_key_and_the_varset_id():
    m = MT.Matcher()
    a = m.add_limit("ns", 5, 60, [], ["descriptors[0].user"])
    b = m.add_limit("ns", 50, 3600, ["descriptors[0].method == 'GET'"], ["descriptors[0].user"])
    c = m.add_limit("ns", 9, 60, [], [])
    assert int(a["varset_id"]) == int(b["varset_id"]) != 0 and int(c["varset_id"]) == 0 and int(c["qualified"]) == 0
    got = m.counters(int(a["ns_id"]), None, [{"user": "bob", "method": "GET"}])
    assert got["limit_id"].tolist() == [0, 1, 2]
    assert got["key_lo"][0] == got["key_lo"][1] != 0 and got["key_lo"][2] == 0
    assert (int(got["key_lo"][0]), int(got["key_hi"][0])) == MT.counter_key({"descriptors[0].user": "bob"})
    got = m.counters(int(a["ns_id"]), None, [{"user": "bob", "method": "PUT"}])
    assert got["limit_id"].tolist() == [0, 2]


def test_cel_shapes_with_other_semantics_in_the_reference_are_refused():
    """expressions the table-driven subset would silently give a different meaning than CEL.
    * a dotted ROOT operand: CEL parses `req.method` as member access on the variable `req`; with the binding
      {"req.method": "GET"} the reference's Predicate::test fails on the unbound `req` (limit/cel.rs:314-322) and
      the limit never applies — so it must not be matched against a root key "req.method" here;
    * escapes: `x == "a\\nb"` compares against a<LF>b in CEL, not against backslash-n;
    * `descriptors[0]['k"]` is a CEL parse error;  * identifiers are ASCII."""
    m = MT.Matcher()
    for conds, vars_ in ((["req.method == 'GET'"], []), ([], ["req.method"]), (["x == 'a\\nb'"], []), (['x == "a\\"b"'], []),
                         (["descriptors[0]['k\"] == 'v'"], []), ([], ["descriptors[0][\"k']"]), ([], ["descriptors[0]['a\\nb']"]),
                         (["n\u00e9 == 'v'"], []), ([], ["descriptors[0].k\u00e9"])):
        with pytest.raises(MT.MatcherError):
            m.add_limit("ns", 5, 60, conds, vars_)
        with pytest.raises(ValueError):
            LM.Limit("ns", 5, 60, conds, vars_)
    assert m.namespace_id("ns") is None
    # the bracket form is how a dotted key IS reached, and non-ASCII bytes are fine inside literals and bracket keys
    d = m.add_limit("ns", 5, 60, ["descriptors[0]['req.method'] == 'G\u00c9T'"], ["descriptors[0]['cl\u00e9']"])
    got = m.counters(int(d["ns_id"]), None, [{"req.method": "G\u00c9T", "cl\u00e9": "7"}])
    assert got["limit_id"].tolist() == [int(d["limit_id"])]
    lim = LM.Limit("ns", 5, 60, ["descriptors[0]['req.method'] == 'G\u00c9T'"], ["descriptors[0]['cl\u00e9']"])
    assert lim.applies(LM.Context({}, [{"req.method": "G\u00c9T", "cl\u00e9": "7"}]))
    assert not lim.applies(LM.Context({"req.method": "G\u00c9T"}, [{"cl\u00e9": "7"}]))


def test_add_limit_keeps_an_equal_live_limit_update_limit_replaces_it():
    """ADVICE r1 / storage/mod.rs:60-83: Storage::add_limit is HashSet::insert (no-op on an equal element: old
    max_value and name stay); update_limit swaps them.  The C API tells the two apart (rl_matcher_add_limit_ex)."""
    m = MT.Matcher()
    d0 = m.add_limit("ns", 5, 60, ["a == 'x'"], ["u"], name="first")
    d1, existed = m.add_limit_keep("ns", 9, 60, ["a == 'x'"], ["u"], name="second")
    assert existed and int(d1["limit_id"]) == int(d0["limit_id"]) and int(d1["max_value"]) == 5
    assert m.limit_name(int(d0["limit_id"])) == "first"
    d2 = m.add_limit("ns", 9, 60, ["a == 'x'"], ["u"], name="second")  # update_limit
    assert int(d2["max_value"]) == 9 and m.limit_name(int(d0["limit_id"])) == "second"
    d3, existed = m.add_limit_keep("ns", 1, 61, ["a == 'x'"], ["u"])
    assert not existed and int(d3["limit_id"]) != int(d0["limit_id"]) and m.limit_name(int(d3["limit_id"])) is None
    m.delete_limit(int(d0["limit_id"]))
    d4, existed = m.add_limit_keep("ns", 3, 60, ["a == 'x'"], ["u"], name="again")  # deleted: a fresh insert
    assert not existed and int(d4["max_value"]) == 3 and m.limit_name(int(d0["limit_id"])) == "again"


def test_more_counters_than_the_engine_takes_per_request_is_refused_by_the_matcher():
    """RL_MAX_COUNTERS_PER_REQUEST = 16 is an engine limit; 17 applicable limits are refused by the
    matcher before anything is enqueued (the device-side resolve would fail the whole batch half-applied)."""
    m = MT.Matcher()
    for i in range(17):
        d = m.add_limit("big", 5, 60 + i, [], [])
    with pytest.raises(MT.MatcherError):
        m.counters(int(d["ns_id"]), {}, None, cap=64)
    m2 = MT.Matcher()
    for i in range(16):
        d = m2.add_limit("ok", 5, 60 + i, [], [])
    assert len(m2.counters(int(d["ns_id"]), {}, None, cap=64)) == 16


def test_too_many_counters_is_an_error_not_a_truncation():
    m = MT.Matcher()
    for i in range(5):
        d = m.add_limit("ns", 5, 60 + i, [], [])
    with pytest.raises(MT.MatcherError):
        m.counters(int(d["ns_id"]), {}, None, cap=4)
    assert len(m.counters(int(d["ns_id"]), {}, None, cap=5)) == 5


def test_concurrent_matching_threads_agree():
    rng = np.random.default_rng(3)
    keys, values = ["m", "u", "path"], ["GET", "POST", "alice"]
    m = MT.Matcher()
    for _ in range(40):
        lim = random_limit(rng, ["ns"], keys, values)
        m.add_limit(lim.namespace, lim.max_value, lim.seconds, lim.conditions, lim.variables, lim.name)
    ns = m.namespace_id("ns")
    ctxs = [random_context(rng, keys, values) for _ in range(300)]
    want = m.counters_batch([ns] * len(ctxs), ctxs)
    results, errors = [None] * 4, []

    def worker(i):
        try:
            for _ in range(5):
                results[i] = m.counters_batch([ns] * len(ctxs), ctxs)
        except Exception as ex:  # pragma: no cover
            errors.append(ex)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors
    for off, ctrs in results:
        assert off.tolist() == want[0].tolist() and ctrs.tobytes() == want[1].tobytes()


def _serving_scenario(seed, n_req=400):
    """limits of an API gateway shape + a request stream of Envoy-like descriptors"""
    rng = np.random.default_rng(seed)
    limits = [
        LM.Limit("api", 5, 60, ["descriptors[0].method == 'GET'"], ["descriptors[0].user"], name="get-per-user"),
        LM.Limit("api", 3, 60, ["descriptors[0].method == 'POST'"], ["descriptors[0].user"], name="post-per-user"),
        LM.Limit("api", 40, 3600, [], ["descriptors[0].user"], name="hourly-per-user"),
        LM.Limit("api", 120, 60, ["descriptors[0].method != 'OPTIONS'"], [], name="global"),
        LM.Limit("admin", 2, 10, [], ["descriptors[0].user", "descriptors[0].path"]),
    ]
    reqs = []
    for i in range(n_req):
        ns = "api" if rng.random() < 0.85 else ("admin" if rng.random() < 0.8 else "nobody")
        d = {"method": str(rng.choice(["GET", "POST", "OPTIONS"])), "user": f"u{int(rng.integers(0, 6))}"}
        if rng.random() < 0.7:
            d["path"] = str(rng.choice(["/a", "/b"]))
        reqs.append((ns, d, int(rng.choice([1, 1, 2])), 1_700_000_000_000_000 + i * 150_000))
    return limits, reqs


@pytest.mark.parametrize("load_counters", [False, True])
def test_matcher_plus_batched_storage_equals_the_mirror_request_by_request(load_counters):
    """End to end on the CPU: native matcher -> CSR -> the oracle's batched check_and_update gives the
    verdicts, limit names, remaining and ttl of the RateLimiter mirror called one request at a time."""
    from oracle import binding as ob
    from tests import helpers as H
    limits, reqs = _serving_scenario(1)
    clock = {"t": 0}
    rl = LM.RateLimiter(H.OracleStorage(), clock=lambda: clock["t"])
    m = MT.Matcher()
    o = ob.Oracle(64)
    for lim in limits:
        rl.add_limit(lim)
        d = m.add_limit(lim.namespace, lim.max_value, lim.seconds, lim.conditions, lim.variables, lim.name)
        o.limit_set(int(d["limit_id"]), int(d["ns_id"]), int(d["max_value"]), int(d["window_us"]), bool(d["qualified"]))
    want = []
    for ns, d, delta, now in reqs:
        clock["t"] = now
        want.append(rl.check_rate_limited_and_update(ns, LM.Context({}, [d]), delta, load_counters))
    ns_ids = [m.namespace_id(ns) if m.namespace_id(ns) is not None else 0xFFFFFF for ns, _, _, _ in reqs]
    off, ctrs = m.counters_batch(ns_ids, [({}, [d]) for _, d, _, _ in reqs])
    lim, fl, rem, ttl = o.batch_csr(0, off, ctrs, [r[2] for r in reqs], [r[3] for r in reqs], load_counters)
    assert lim.tolist() == [int(w.limited) for w in want]
    assert 0 < int(lim.sum()) < len(lim)
    for i, w in enumerate(want):
        if w.limited:
            assert m.limit_name(int(fl[i])) == w.limit_name
        if load_counters:
            got = sorted((int(c["limit_id"]), int(rem[off[i] + j]), int(ttl[off[i] + j])) for j, c in enumerate(ctrs[off[i]:off[i + 1]]))
            assert got == sorted((c.limit_id, c.remaining, c.expires_in_us) for c in w.counters)


@pytest.mark.gpu
def test_matcher_feeds_the_engine():
    """The same stream through the native matcher and ONE rl_check_and_update_batch call on the GPU."""
    from limitador_b200 import Engine
    from tests import helpers as H
    limits, reqs = _serving_scenario(2, n_req=1500)
    clock = {"t": 0}
    rl = LM.RateLimiter(H.OracleStorage(), clock=lambda: clock["t"])
    m = MT.Matcher()
    e = Engine(capacity_rows=1 << 12, cells_per_row=3, max_batch=4096)
    descs = []
    for lim in limits:
        rl.add_limit(lim)
        descs.append(m.add_limit(lim.namespace, lim.max_value, lim.seconds, lim.conditions, lim.variables, lim.name))
    e.limits_set(np.array(descs))
    want = []
    for ns, d, delta, now in reqs:
        clock["t"] = now
        want.append(rl.check_rate_limited_and_update(ns, LM.Context({}, [d]), delta, False))
    ns_ids = [m.namespace_id(ns) if m.namespace_id(ns) is not None else 0xFFFFFF for ns, _, _, _ in reqs]
    off, ctrs = m.counters_batch(ns_ids, [({}, [d]) for _, d, _, _ in reqs])
    lim, fl, _, _ = e.check_and_update_batch(off, ctrs, [r[2] for r in reqs], [r[3] for r in reqs], False)
    assert lim.tolist() == [int(w.limited) for w in want]
    assert [m.limit_name(int(f)) for f, w in zip(fl, want) if w.limited] == [w.limit_name for w in want if w.limited]


def test_expression_shapes_of_the_reference_docs_and_example_configs_are_accepted():
    """The condition / variable shapes that appear in the reference's documentation and example limit files
    (bracketed keys with dots, dotted member access, != on a path) all fall inside the subset."""
    m = MT.Matcher()
    lims = [
        (["descriptors[0]['req.method'] == 'GET'", "descriptors[0]['req.path'] != '/json'"], []),
        (["descriptors[0]['req.method'] == 'POST'", "descriptors[0]['req.path'] != '/json'"], ["descriptors[0]['user_id']"]),
        (["descriptors[0].req_method == 'GET'"], ["descriptors[0].user_id"]),
        (["descriptors[0].KEY_A == 'VALUE_A'", "descriptors[0].OTHER_KEY == 'WRONG_VALUE'"], []),
    ]
    ids = [int(m.add_limit("test_namespace", 5, 60 + i, c, v)["limit_id"]) for i, (c, v) in enumerate(lims)]
    ns = m.namespace_id("test_namespace")
    got = m.counters(ns, None, [{"req.method": "GET", "req.path": "/", "user_id": "7", "req_method": "GET", "KEY_A": "VALUE_A"}])
    assert got["limit_id"].tolist() == [ids[0], ids[2]]
    got = m.counters(ns, None, [{"req.method": "POST", "req.path": "/x", "user_id": "7"}])
    assert got["limit_id"].tolist() == [ids[1]]
    # descriptors[0].user_id and descriptors[0]['user_id'] are different identities reading the same entry
    assert (int(got["key_lo"][0]), int(got["key_hi"][0])) == MT.counter_key({"descriptors[0]['user_id']": "7"})
    got = m.counters(ns, None, [{"req.method": "GET", "req.path": "/json"}])
    assert len(got) == 0


def test_parser_accepts_exactly_what_the_python_mirror_accepts():
    """Random strings over the grammar's alphabet: the native parser and the mirror's regular expressions
    agree on accept / refuse for conditions and for variables (and nothing crashes)."""
    rng = np.random.default_rng(11)
    pieces = ["descriptors", "[", "]", "0", "12", ".", "'", '"', "==", "!=", " ", "a", "b_1", "req.path", "=", "!", "x y",
              "descriptors[0]", "descriptors[1].k", "descriptors[0]['k.v']", "'lit'", '"lit"', "\t", "é", "9z", "_"]
    m = MT.Matcher()
    n_acc_c = n_acc_v = 0
    operands = ["a", "b_1", "req.path", "descriptors[0].k", "descriptors[12]['k.v']", 'descriptors[1]["x"]', "descriptors[0]",
                "descriptors[0].9", "descriptors[x].k", "descriptors[0]['']", "9z", "_u", "é", "descriptors", "a.b.c", "a b",
                "descriptors[0]['k\"]", "descriptors[0][\"k']", "descriptors[0]['a\\nb']", "aé", "descriptors[0].ké",
                'descriptors[0]["it\'s"]', "req.method"]
    ops = ["==", "!=", " == ", "\t!= ", "=", "!==", "<", ""]
    lits = ["'lit'", '"lit"', "''", "'x y'", "'unterminated", "bare", "'a'b'", '"q\'q"', "'tail' x", "",
            "'a\\nb'", '"a\\"b"', "'\\'", '"tab\there"', "'é'"]
    for i in range(4000):
        if rng.random() < 0.6:  # near-valid: operand op literal with optional padding
            pad = [" ", "", "  ", "\t"]
            s = str(rng.choice(pad)) + str(rng.choice(operands)) + str(rng.choice(ops)) + str(rng.choice(lits)) + str(rng.choice(pad))
            if rng.random() < 0.3:
                s = str(rng.choice(pad)) + str(rng.choice(operands)) + str(rng.choice(pad))
        else:  # noise
            s = "".join(str(rng.choice(pieces)) for _ in range(int(rng.integers(1, 7))))
        want_c, want_v = LM._PRED.match(s) is not None, LM._VAR.match(s) is not None
        try:
            m.add_limit("fuzz", 1, 1 + i, [s], [])
            got_c = True
        except MT.MatcherError:
            got_c = False
        try:
            m.add_limit("fuzz", 1, 100000 + i, [], [s])
            got_v = True
        except MT.MatcherError:
            got_v = False
        assert got_c == want_c, f"condition {s!r}: native {got_c}, mirror {want_c}"
        assert got_v == want_v, f"variable {s!r}: native {got_v}, mirror {want_v}"
        n_acc_c += got_c
        n_acc_v += got_v
    assert n_acc_c > 20 and n_acc_v > 50  # the generator does produce valid expressions


def test_matcher_csr_through_the_kernel_algorithm_emulation():
    """The stream of the GPU test above through tests/emu (the kernels' batching algorithm on the host,
    3 cells per row, coupled requests spanning the per-user row and the namespace's unqualified row)."""
    from tests import helpers as H
    limits, reqs = _serving_scenario(2, n_req=1500)
    clock = {"t": 0}
    rl = LM.RateLimiter(H.OracleStorage(), clock=lambda: clock["t"])
    m = MT.Matcher()
    descs = []
    for lim in limits:
        rl.add_limit(lim)
        descs.append(m.add_limit(lim.namespace, lim.max_value, lim.seconds, lim.conditions, lim.variables, lim.name))
    emu = H.Emu(np.array(descs), 3)
    want = []
    for ns, d, delta, now in reqs:
        clock["t"] = now
        want.append(rl.check_rate_limited_and_update(ns, LM.Context({}, [d]), delta, False))
    ns_ids = [m.namespace_id(ns) if m.namespace_id(ns) is not None else 0xFFFFFF for ns, _, _, _ in reqs]
    off, ctrs = m.counters_batch(ns_ids, [({}, [d]) for _, d, _, _ in reqs])
    lim, fl, _, _ = emu.batch_csr(0, off, ctrs, [r[2] for r in reqs], [r[3] for r in reqs], False)
    assert lim.tolist() == [int(w.limited) for w in want]
    assert [m.limit_name(int(f)) for f, w in zip(fl, want) if w.limited] == [w.limit_name for w in want if w.limited]
    assert emu.rounds >= 1  # coupled requests did go through the fixed-point rounds


def test_response_headers_render_the_reference_strings():
    """CheckResult::response_header (lib.rs:235-275): the exact header strings of the reference's server tests
    (envoy_rls/server.rs:337-426, 496-591, 593-680) and agreement with the Python mirror on random counters,
    names with quotes included."""
    m = MT.Matcher()
    a = m.add_limit("ns", 1, 60, [], ["u"])
    ctr = np.array([(int(a["limit_id"]), 0, 1, 0)], dtype=COUNTER_DTYPE)
    assert m.response_headers(ctr, [0], [59_500_000]) == {"X-RateLimit-Limit": "1, 1;w=60", "X-RateLimit-Remaining": "0",
                                                           "X-RateLimit-Reset": "59"}
    b = m.add_limit("ns", 0, 60, [], ["v"])
    c = m.add_limit("ns", 10, 60, [], ["w"])
    two = np.array([(int(c["limit_id"]), 0, 1, 0), (int(b["limit_id"]), 0, 1, 0)], dtype=COUNTER_DTYPE)
    assert m.response_headers(two, [10, 0], [60_000_000, 60_000_000])["X-RateLimit-Limit"] == "0, 0;w=60, 10;w=60"
    assert m.response_headers(ctr[:0], [], []) == {}
    assert m.response_headers(np.array([(int(c["limit_id"]), 0, 1, 0)], dtype=COUNTER_DTYPE), [4], [1])["X-RateLimit-Remaining"] == "4"
    # random agreement with the mirror
    rng = np.random.default_rng(5)
    lims = []
    for i in range(12):
        name = None if i % 3 == 0 else f'lim "{i}" x'
        L = LM.Limit("rnd", int(rng.integers(0, 1000)), int(rng.choice([1, 60, 3600])) + i, [], ["u"], name)
        lims.append((L, int(m.add_limit(L.namespace, L.max_value, L.seconds, L.conditions, L.variables, L.name)["limit_id"])))
    for _ in range(200):
        pick = [lims[int(j)] for j in rng.choice(len(lims), size=int(rng.integers(1, 6)), replace=False)]
        rem = [int(rng.integers(0, 5)) for _ in pick]  # ties exercise the stable order
        ttl = [int(rng.integers(0, 4_000_000_000)) for _ in pick]
        cs = [LM.Counter(L, {"u": "x"}, remaining=r, expires_in_us=t, limit_id=lid) for (L, lid), r, t in zip(pick, rem, ttl)]
        want = LM.CheckResult(False, cs).response_header()
        got = m.response_headers(np.array([(lid, 0, 1, 0) for _, lid in pick], dtype=COUNTER_DTYPE), rem, ttl)
        assert got == want


def test_counter_cap_defaults_to_what_the_engine_takes_and_can_be_raised_for_matching_only():
    """The reference's bench scenarios hold 50 limits per namespace (benches/bench.rs:65-90): 50 counters per request.
    The engine takes 16 per request, so the matcher refuses such a request before anything is enqueued — unless the
    caller only matches and raises the cap."""
    m = MT.Matcher()
    for l in range(50):
        m.add_limit("ns", 10, 10 + l, ["cond == '1'"], ["var"])
    with pytest.raises(MT.MatcherError, match="counters"):
        m.counters(m.namespace_id("ns"), {"cond": "1", "var": "v"}, cap=64)
    m.set_counter_cap(64)
    got = m.counters(m.namespace_id("ns"), {"cond": "1", "var": "v"}, cap=64)
    assert len(got) == 50 and len(set(got["limit_id"].tolist())) == 50
    assert len({(int(c["key_lo"]), int(c["key_hi"])) for c in got}) == 1  # one variable set: one key digest


def test_batch_matching_by_namespace_string_equals_the_per_request_calls():
    """rl_matcher_counters_batch_ns (one reader section per batch — what the RLS stage uses) against rl_matcher_counters
    request by request, including namespaces without limits and a request with more counters than the engine takes."""
    limits, reqs = _serving_scenario(7, n_req=300)
    m = MT.Matcher()
    for lim in limits:
        m.add_limit(lim.namespace, lim.max_value, lim.seconds, lim.conditions, lim.variables, lim.name)
    for l in range(20):  # a namespace in which 20 limits apply at once
        m.add_limit("wide", 10, 10 + l, [], ["descriptors[0].user"])
    names = [ns for ns, _, _, _ in reqs] + ["wide", "nobody", "api"]
    ctxs = [({}, [d]) for _, d, _, _ in reqs] + [({}, [{"user": "u"}]), ({}, [{"user": "u"}]), ({}, [{"method": "GET", "user": "z"}])]
    off, ctrs, status = m.counters_batch_ns(names, ctxs)
    for i, (ns, ctx) in enumerate(zip(names, ctxs)):
        nid = m.namespace_id(ns)
        got = ctrs[off[i]:off[i + 1]]
        if nid is None:
            assert status[i] == 1 and len(got) == 0
        elif ns == "wide":
            assert status[i] == 2 and len(got) == 0
        else:
            assert status[i] == 0 and got.tobytes() == m.counters(nid, *ctx).tobytes()
    assert status[-1] == 0 and off[-1] - off[-2] >= 2  # the request behind the refused one is matched as usual


def test_batch_header_rendering_equals_the_per_request_rendering():
    rng = np.random.default_rng(4)
    m = MT.Matcher()
    ids = [int(m.add_limit("ns", 10 * (k + 1), 60 * (k + 1), [], ["descriptors[0].u"], name=('a "q" name' if k % 2 else None))["limit_id"])
           for k in range(6)]
    sizes = [0, 1, 3, 6, 2, 0, 5]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    ctrs = np.zeros(int(off[-1]), dtype=MT._eng.COUNTER_DTYPE)
    ctrs["limit_id"] = rng.choice(ids, size=len(ctrs))
    rem = rng.integers(0, 50, size=len(ctrs)).astype(np.uint64)
    ttl = rng.integers(0, 120_000_000, size=len(ctrs)).astype(np.uint64)
    got = m.response_headers_batch(off, ctrs, rem, ttl)  # starts with no room at all: the call says what it needs
    for i, n in enumerate(sizes):
        sl = slice(int(off[i]), int(off[i + 1]))
        assert got[i] == m.response_headers(ctrs[sl], rem[sl], ttl[sl])
    assert got[3]["X-RateLimit-Limit"].count(";w=") == 6 and "name=\"a 'q' name\"" in got[3]["X-RateLimit-Limit"]
