"""Envoy RLS v3 wire surface (include/rl_rls.h): the hand-written codec against the protobuf runtime,
and the service's plan -> store -> finish stages against the reference's own server tests
(limitador-server/src/envoy_rls/server.rs:337-771, kuadrant_service.rs tests), with the CPU oracle as the store
between the two CPU stages (tests/test_zz1_rls_gpu.py runs the same scenarios through rl_rls_serve on the GPU)."""
import numpy as np
import pytest

from limitador_b200 import matcher as MT
from limitador_b200 import rls as R

T0 = 1_700_000_000_000_000


# ---- an independent codec: the protobuf runtime over descriptors built from the .proto field numbers ------------------
def _proto_classes():
    from google.protobuf import descriptor_pb2 as dp
    from google.protobuf import descriptor_pool, message_factory
    f = dp.FileDescriptorProto(name="rls_under_test.proto", package="rlt", syntax="proto3")
    T = dp.FieldDescriptorProto

    def msg(name, fields):
        m = f.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            fd = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                fd.type_name = ".rlt." + tname

    opt, rep = T.LABEL_OPTIONAL, T.LABEL_REPEATED
    # envoy/extensions/common/ratelimit/v3/ratelimit.proto: RateLimitDescriptor.Entry {key = 1, value = 2},
    # RateLimitOverride {requests_per_unit = 1, unit = 2}, RateLimitDescriptor {entries = 1, limit = 2}
    msg("Entry", [("key", 1, T.TYPE_STRING, opt, None), ("value", 2, T.TYPE_STRING, opt, None)])
    msg("Override", [("requests_per_unit", 1, T.TYPE_UINT32, opt, None), ("unit", 2, T.TYPE_INT32, opt, None)])
    msg("Descriptor", [("entries", 1, T.TYPE_MESSAGE, rep, "Entry"), ("limit", 2, T.TYPE_MESSAGE, opt, "Override")])
    # envoy/service/ratelimit/v3/rls.proto: RateLimitRequest {domain = 1, descriptors = 2, hits_addend = 3};
    # RequestExt writes three fields that are not in the .proto (what a newer Envoy may send): they must be skipped
    real = [("domain", 1, T.TYPE_STRING, opt, None), ("descriptors", 2, T.TYPE_MESSAGE, rep, "Descriptor"),
            ("hits_addend", 3, T.TYPE_UINT32, opt, None)]
    msg("Request", real)
    msg("RequestExt", real + [("future", 15, T.TYPE_STRING, opt, None), ("future_fixed", 16, T.TYPE_FIXED64, opt, None),
                              ("future_f32", 17, T.TYPE_FIXED32, opt, None)])
    # envoy/config/core/v3/base.proto HeaderValue {key = 1, value = 2}; RateLimitResponse {overall_code = 1,
    # response_headers_to_add = 3, request_headers_to_add = 4, raw_body = 5}
    msg("HeaderValue", [("key", 1, T.TYPE_STRING, opt, None), ("value", 2, T.TYPE_STRING, opt, None)])
    msg("Response", [("overall_code", 1, T.TYPE_INT32, opt, None),
                     ("response_headers_to_add", 3, T.TYPE_MESSAGE, rep, "HeaderValue"),
                     ("request_headers_to_add", 4, T.TYPE_MESSAGE, rep, "HeaderValue"),
                     ("raw_body", 5, T.TYPE_BYTES, opt, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("rlt." + n))  # noqa: E731
    return get("Request"), get("Response"), get("RequestExt")


@pytest.fixture(scope="module")
def pb():
    return _proto_classes()


def _pb_request(pb, domain, descriptors, hits=0, future=None, override=None):
    q = (pb[2] if future is not None else pb[0])(domain=domain, hits_addend=hits)
    for d in descriptors:
        dd = q.descriptors.add()
        for k, v in d:
            dd.entries.add(key=k, value=v)
        if override:
            dd.limit.requests_per_unit, dd.limit.unit = override
    if future is not None:
        q.future = future
        q.future_fixed = 7
        q.future_f32 = 9
    return q


def test_request_decoder_reads_what_the_protobuf_runtime_writes(pb):
    Request = pb[0]
    rng = np.random.default_rng(3)
    words = ["", "a", "GET", "req.method", "app.id", "ü-ñ", "x" * 200, "1", "日本"]
    for it in range(300):
        descs = [[(str(rng.choice(words)), str(rng.choice(words))) for _ in range(int(rng.integers(0, 4)))]
                 for _ in range(int(rng.integers(0, 4)))]
        domain = str(rng.choice(["", "ns", "test_namespace", "ü"]))
        hits = int(rng.choice([0, 1, 6, 300, 2**32 - 1]))
        q = _pb_request(pb, domain, descs, hits, future="zz" if it % 3 == 0 else None,
                        override=(5, 2) if it % 5 == 0 else None)
        got = R.decode_request(q.SerializeToString())
        assert got == (domain, descs, hits)
        # and the pure-Python encoder of the binding writes the same message
        assert Request.FromString(R.encode_request(domain, descs, hits)) == _pb_request(pb, domain, descs, hits)


def test_request_decoder_and_the_protobuf_runtime_refuse_the_same_mutations(pb):
    """Differential fuzz: truncations and byte flips of valid messages — both decoders accept or both refuse, and an
    accepted message decodes to the same content."""
    from google.protobuf.message import DecodeError
    Request = pb[0]
    rng = np.random.default_rng(11)
    base = [_pb_request(pb, "test_namespace", [[("req.method", "GET"), ("app.id", "1")], [("y", "2")]], 6, "f").SerializeToString(),
            _pb_request(pb, "d", [[("k", "v")]], 0, override=(9, 1)).SerializeToString()]
    checked = refused = stricter = 0
    for it in range(4000):
        b = bytearray(base[it % 2])
        if it % 3 == 0:
            b = b[:int(rng.integers(0, len(b)))]
        else:
            for _ in range(int(rng.integers(1, 3))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        b = bytes(b)
        try:
            q = Request.FromString(b)
            want = (q.domain, [[(e.key, e.value) for e in d.entries] for d in q.descriptors], q.hits_addend)
        except DecodeError:
            want = None
        try:
            got = R.decode_request(b)
        except R.RlsError:
            got = None
        if got is None:
            # prost refuses a known field that arrives with another wire type ("invalid wire type"), upb keeps it as
            # an unknown field: the native decoder follows prost, so it may be stricter than the Python runtime —
            # never more lenient
            refused += 1
            stricter += want is not None
        else:
            assert want is not None, (b.hex(), got)
            assert got == want, b.hex()
            checked += 1
    assert checked > 200 and refused > 200 and stricter < refused // 4


def test_malformed_messages_are_refused():
    for bad in (b"\x0a", b"\x0a\x05ab", b"\x00\x01", b"\x18", b"\x18\xff\xff\xff\xff\xff\xff\xff\xff\xff\x02",
                b"\x0a\x02\xc3\x28", b"\x0a\x03\xed\xa0\x80", b"\x0a\x02\xc0\x80", b"\x0d\x01\x02\x03\x04", b"\x1a\x00",
                b"\x12\x03\x0a\x05a", b"\x0f"):
        with pytest.raises(R.RlsError):
            R.decode_request(bad)
    # unknown fields of every wire type are skipped, groups included
    assert R.decode_request(b"\x7b\x08\x01\x7c" + R.encode_request("ns", [[("k", "v")]], 2)) == ("ns", [[("k", "v")]], 2)
    # a repeated scalar field: the last occurrence wins
    assert R.decode_request(b"\x0a\x01a\x0a\x01b\x18\x05\x18\x07")[0::2] == ("b", 7)


def test_response_encoder_writes_what_the_protobuf_runtime_reads(pb):
    Response = pb[1]
    cases = [(R.CODE_UNKNOWN, []), (R.CODE_OK, []), (R.CODE_OVER_LIMIT, []),
             (R.CODE_OK, [("X-RateLimit-Limit", "1, 1;w=60"), ("X-RateLimit-Remaining", "0"), ("X-RateLimit-Reset", "59")]),
             (R.CODE_OVER_LIMIT, [("X-RateLimit-Limit", "10, 10;w=60;name=\"a 'quoted' one\", " + "5;w=1, " * 40), ("k", "")])]
    for code, headers in cases:
        b = R.encode_response(code, headers)
        r = Response.FromString(b)
        assert r.overall_code == code
        assert [(h.key, h.value) for h in r.response_headers_to_add] == headers
        assert not r.request_headers_to_add and not r.raw_body
        want = Response(overall_code=code)
        for k, v in headers:
            want.response_headers_to_add.add(key=k, value=v)
        assert b == want.SerializeToString()  # byte-identical to the canonical encoding
        assert R.decode_response(b) == (code, headers)
    assert R.encode_response(R.CODE_UNKNOWN) == b""  # proto3: a default message is empty


# ---- the service, with the oracle as the store between the CPU stages --------------------------------------------
class CpuHarness:
    """plan -> (the CPU oracle decides) -> finish.  Test infrastructure only: the product's decide stage is the engine."""

    def __init__(self, limits, headers=R.HEADERS_DRAFT_VERSION_03, threads=1, use_limit_name_label=False):
        from oracle import binding as ob
        self.m = MT.Matcher()
        self.o = ob.Oracle(64)
        self.descs = []
        for (ns, mx, secs, conds, vars_, name) in limits:
            d = self.m.add_limit(ns, mx, secs, conds, vars_, name)
            self.descs.append(d)
            self.o.limit_set(int(d["limit_id"]), int(d["ns_id"]), int(d["max_value"]), int(d["window_us"]), bool(d["qualified"]))
        self.svc = R.RlsService(self.m, None, headers, threads, use_limit_name_label)

    def call(self, method, msgs, now_us=T0):
        buf, off = R.pack_requests(msgs)
        p = self.svc.plan(method, buf, off, now_us)
        self.last_plan = p
        mode = {R.SHOULD_RATE_LIMIT: 0, R.CHECK_RATE_LIMIT: 1, R.REPORT: 2}[method]
        lim = fl = rem = ttl = None
        if p["n_store"]:
            lim, fl, rem, ttl = self.o.batch_csr(mode, p["ctr_off"], p["ctrs"], p["delta"], p["now_us"], p["load_counters"])
        out = self.svc.finish(lim, fl, rem, ttl)
        return [(g, R.decode_response(b) if g == 0 else None) for g, b in out]


def _req(ns, descriptors, hits=1):
    return R.encode_request(ns, descriptors, hits)


def _hdr(resp):
    return dict(resp[1][1])


def test_returns_ok_and_overlimit_correctly():
    """server.rs:337-426: limit 1/60s on descriptors[0]['req.method'] == 'GET' keyed by descriptors[0]['app.id']."""
    h = CpuHarness([("test_namespace", 1, 60, ["descriptors[0]['req.method'] == 'GET'"], ["descriptors[0]['app.id']"], None)])
    req = _req("test_namespace", [[("req.method", "GET"), ("app.id", "1")]])
    r1, = h.call(R.SHOULD_RATE_LIMIT, [req])
    assert r1[0] == R.GRPC_OK and r1[1][0] == R.CODE_OK
    assert [k for k, _ in r1[1][1]] == ["X-RateLimit-Limit", "X-RateLimit-Remaining", "X-RateLimit-Reset"]
    assert _hdr(r1)["X-RateLimit-Limit"] == "1, 1;w=60" and _hdr(r1)["X-RateLimit-Remaining"] == "0"
    assert int(_hdr(r1)["X-RateLimit-Reset"]) <= 60
    r2, = h.call(R.SHOULD_RATE_LIMIT, [req], T0 + 1_000_000)
    assert r2[1][0] == R.CODE_OVER_LIMIT and len(r2[1][1]) == 3
    assert _hdr(r2)["X-RateLimit-Limit"] == "1, 1;w=60" and _hdr(r2)["X-RateLimit-Remaining"] == "0"
    assert int(_hdr(r2)["X-RateLimit-Reset"]) <= 60
    # the same two requests in ONE batch: array order is the stream order
    h2 = CpuHarness([("test_namespace", 1, 60, ["descriptors[0]['req.method'] == 'GET'"], ["descriptors[0]['app.id']"], None)])
    a, b = h2.call(R.SHOULD_RATE_LIMIT, [req, req])
    assert (a[1][0], b[1][0]) == (R.CODE_OK, R.CODE_OVER_LIMIT)


def test_returns_ok_when_no_limits_apply():
    """server.rs:428-461"""
    h = CpuHarness([])
    r, = h.call(R.SHOULD_RATE_LIMIT, [_req("test_namespace", [[("req.method", "GET")]])])
    assert r == (R.GRPC_OK, (R.CODE_OK, []))
    assert h.last_plan["n_store"] == 0  # nothing reaches the store (lib.rs:434-440)


def test_returns_unknown_when_domain_is_empty():
    """server.rs:463-494"""
    h = CpuHarness([("test_namespace", 1, 60, [], [], None)])
    r, = h.call(R.SHOULD_RATE_LIMIT, [_req("", [[("req.method", "GET")]])])
    assert r == (R.GRPC_OK, (R.CODE_UNKNOWN, []))
    assert "authorized_calls{" not in h.svc.metrics() and "limited_calls{" not in h.svc.metrics()


def test_takes_into_account_all_the_descriptors():
    """server.rs:496-591: the second limit (max 0) needs descriptors[1].y == '2'."""
    h = CpuHarness([("test_namespace", 10, 60, ["descriptors[0].x == '1'"], ["descriptors[0].z"], None),
                    ("test_namespace", 0, 60, ["descriptors[0].x == '1'", "descriptors[1].y == '2'"], ["descriptors[0].z"], None)])
    r, = h.call(R.SHOULD_RATE_LIMIT, [_req("test_namespace", [[("x", "1"), ("z", "1")], [("y", "2")]])])
    assert r[1][0] == R.CODE_OVER_LIMIT and len(r[1][1]) == 3
    assert _hdr(r)["X-RateLimit-Limit"] == "0, 0;w=60, 10;w=60" and _hdr(r)["X-RateLimit-Remaining"] == "0"


def test_takes_into_account_the_hits_addend_param():
    """server.rs:593-680: limit 10, addend 6: Ok with Remaining 4, then OverLimit with Remaining 0."""
    h = CpuHarness([("test_namespace", 10, 60, ["descriptors[0].x == '1'"], ["descriptors[0].y"], None)])
    req = _req("test_namespace", [[("x", "1"), ("y", "1")]], 6)
    a, b = h.call(R.SHOULD_RATE_LIMIT, [req, req])
    assert a[1][0] == R.CODE_OK and _hdr(a)["X-RateLimit-Limit"] == "10, 10;w=60" and _hdr(a)["X-RateLimit-Remaining"] == "4"
    assert b[1][0] == R.CODE_OVER_LIMIT and _hdr(b)["X-RateLimit-Remaining"] == "0"
    assert h.last_plan["delta"].tolist() == [6, 6]


def test_0_hits_addend_is_converted_to_1():
    """server.rs:682-771"""
    h = CpuHarness([("test_namespace", 1, 60, ["descriptors[0].x == '1'"], ["descriptors[0].y"], None)])
    req = _req("test_namespace", [[("x", "1"), ("y", "2")]], 0)
    a, b = h.call(R.SHOULD_RATE_LIMIT, [req, req])
    assert (a[1][0], b[1][0]) == (R.CODE_OK, R.CODE_OVER_LIMIT)
    assert h.last_plan["delta"].tolist() == [1, 1]
    assert _hdr(a)["X-RateLimit-Limit"] == "1, 1;w=60" and _hdr(a)["X-RateLimit-Remaining"] == "0"


def test_headers_none_sends_no_headers_and_does_not_load_counters():
    h = CpuHarness([("ns", 1, 60, [], ["descriptors[0].u"], None)], headers=R.HEADERS_NONE)
    a, b = h.call(R.SHOULD_RATE_LIMIT, [_req("ns", [[("u", "1")]])] * 2)
    assert a == (0, (R.CODE_OK, [])) and b == (0, (R.CODE_OVER_LIMIT, []))
    assert h.last_plan["load_counters"] is False


def test_check_rate_limit_and_report_split_path():
    """kuadrant_service.rs: CheckRateLimit = is_rate_limited(ns, ctx, 1) (read-only, no headers, hits_addend ignored);
    Report = update_counters(ns, ctx, hits_addend) (always OK)."""
    h = CpuHarness([("ns", 3, 60, ["descriptors[0].x == '1'"], ["descriptors[0].y"], "L")])
    req = _req("ns", [[("x", "1"), ("y", "k")]], 2)
    for _ in range(3):  # checks never count
        r, = h.call(R.CHECK_RATE_LIMIT, [req])
        assert r == (0, (R.CODE_OK, []))
    assert h.last_plan["delta"].tolist() == [1] and h.last_plan["load_counters"] is False
    r, = h.call(R.REPORT, [req])  # +2
    assert r == (0, (R.CODE_OK, []))
    r, = h.call(R.CHECK_RATE_LIMIT, [req])  # 2 + 1 <= 3
    assert r[1][0] == R.CODE_OK
    h.call(R.REPORT, [req])  # 4: update_counters may exceed the limit
    r, = h.call(R.CHECK_RATE_LIMIT, [req])
    assert r[1][0] == R.CODE_OVER_LIMIT
    assert h.call(R.CHECK_RATE_LIMIT, [_req("", [])])[0] == (0, (R.CODE_UNKNOWN, []))
    assert h.call(R.REPORT, [_req("", [])])[0] == (0, (R.CODE_UNKNOWN, []))
    text = h.svc.metrics()
    assert 'authorized_calls{limitador_namespace="ns"} 4' in text  # 4 allowed checks
    assert 'authorized_hits{limitador_namespace="ns"} 4' in text   # two reports of 2
    assert 'limited_calls{limitador_namespace="ns"} 1' in text


def test_metrics_by_namespace_and_by_limit_name():
    """prometheus_metrics.rs:93-125 and its tests: one increment per request after the decision; limited_calls carries
    limit_name (or "") when the label is on."""
    h = CpuHarness([("a", 2, 60, [], ["descriptors[0].u"], "Some limit"), ("b", 0, 60, [], [], None)], use_limit_name_label=True)
    msgs = [_req("a", [[("u", "1")]], 3)] + [_req("a", [[("u", "2")]], 1)] * 3 + [_req("b", [[("u", "1")]])] * 2 + [_req("nolimits", [])]
    out = h.call(R.SHOULD_RATE_LIMIT, msgs)
    assert [r[1][0] for r in out] == [2, 1, 1, 2, 2, 2, 1]
    t = h.svc.metrics()
    assert 'authorized_calls{limitador_namespace="a"} 2' in t and 'authorized_hits{limitador_namespace="a"} 2' in t
    assert 'authorized_calls{limitador_namespace="nolimits"} 1' in t and 'authorized_hits{limitador_namespace="nolimits"} 1' in t
    assert 'limited_calls{limitador_namespace="a",limit_name="Some limit"} 2' in t
    assert 'limited_calls{limitador_namespace="b",limit_name=""} 2' in t
    assert "limitador_up 1" in t
    h2 = CpuHarness([("b", 0, 60, [], [], None)])
    h2.call(R.SHOULD_RATE_LIMIT, [_req("b", [])] * 3)
    assert 'limited_calls{limitador_namespace="b"} 3' in h2.svc.metrics()


def test_undecodable_and_unsupported_requests_get_a_grpc_error_not_a_verdict():
    h = CpuHarness([("ns", 1, 60, [], ["descriptors[0].u"], None)])
    good = _req("ns", [[("u", "1")]])
    out = h.call(R.SHOULD_RATE_LIMIT, [good, b"\x0a\x05ab", _req("ns", [[("u", "a\x00b")]]), good])
    assert [g for g, _ in out] == [0, R.GRPC_INTERNAL, R.GRPC_UNAVAILABLE, 0]
    assert out[0][1][0] == R.CODE_OK and out[3][1][0] == R.CODE_OVER_LIMIT  # the two good ones are one stream
    assert h.last_plan["store_index"].tolist() == [0, R.NO_STORE, R.NO_STORE, 1]
    # a failing store call: every request that needed it is answered UNAVAILABLE (server.rs:160-172), the others as usual
    buf, off = R.pack_requests([good, _req("", []), _req("other", [])])
    h.svc.plan(R.SHOULD_RATE_LIMIT, buf, off, T0)
    out = h.svc.finish(store_status=1)
    assert [g for g, _ in out] == [R.GRPC_UNAVAILABLE, 0, 0]


def test_serve_without_an_engine_fails_loudly():
    h = CpuHarness([("ns", 1, 60, [], [], None)])
    buf, off = R.pack_requests([_req("ns", [])])
    with pytest.raises(R.RlsError, match="no CPU store"):
        h.svc.serve(R.SHOULD_RATE_LIMIT, buf, off, T0)


def _gateway(seed, n):
    rng = np.random.default_rng(seed)
    limits = [("api", 5, 60, ["descriptors[0].method == 'GET'"], ["descriptors[0].user"], "get-per-user"),
              ("api", 3, 60, ["descriptors[0].method == 'POST'"], ["descriptors[0].user"], "post-per-user"),
              ("api", 40, 3600, [], ["descriptors[0].user"], "hourly-per-user"),
              ("api", 120, 60, ["descriptors[0].method != 'OPTIONS'"], [], "global"),
              ("admin", 2, 10, [], ["descriptors[0].user", "descriptors[1].path"], None)]
    reqs = []
    for _ in range(n):
        ns = "api" if rng.random() < 0.8 else str(rng.choice(["admin", "nobody", ""]))
        d0 = [("method", str(rng.choice(["GET", "POST", "OPTIONS"]))), ("user", f"u{int(rng.integers(0, 7))}")]
        if rng.random() < 0.2:
            d0.append(("user", f"u{int(rng.integers(0, 7))}"))  # a duplicate key: the last one wins (HashMap::insert)
        descs = [d0] + ([[("path", str(rng.choice(["/a", "/b"])))]] if rng.random() < 0.7 else [])
        reqs.append((ns, descs, int(rng.choice([0, 1, 1, 2]))))
    return limits, reqs


@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("headers", [R.HEADERS_NONE, R.HEADERS_DRAFT_VERSION_03])
def test_a_batch_equals_the_python_mirror_called_request_by_request(threads, headers):
    """A batch of wire requests through plan -> oracle -> finish gives, request by request, the response the Python
    mirror of RateLimiter (limiter.py, pinned by the reference's behaviour tests) gives when called one at a time."""
    from limitador_b200 import limiter as LM
    from tests import helpers as H
    limits, reqs = _gateway(5, 700)
    clock = {"t": T0}
    rl = LM.RateLimiter(H.OracleStorage(), clock=lambda: clock["t"])
    for ns, mx, secs, conds, vars_, name in limits:
        rl.add_limit(LM.Limit(ns, mx, secs, conds, vars_, name=name))
    h = CpuHarness(limits, headers=headers, threads=threads)
    got = h.call(R.SHOULD_RATE_LIMIT, [_req(ns, descs, hits) for ns, descs, hits in reqs])
    n_over = 0
    for (ns, descs, hits), (grpc, resp) in zip(reqs, got):
        assert grpc == 0
        if ns == "":
            assert resp == (R.CODE_UNKNOWN, [])
            continue
        w = rl.check_rate_limited_and_update(ns, LM.Context({}, [dict(d) for d in descs]), hits or 1, headers != R.HEADERS_NONE)
        assert resp[0] == (R.CODE_OVER_LIMIT if w.limited else R.CODE_OK)
        n_over += w.limited
        want_h = sorted(w.response_header().items()) if headers != R.HEADERS_NONE else []
        assert resp[1] == want_h
    assert 50 < n_over < 650


# ---- the Kuadrant service's own tests (limitador-server/src/envoy_rls/kuadrant_service.rs:189-650), one by one ---------
_KUADRANT_LIMIT = ("test_namespace", 1, 60, ["descriptors[0]['req.method'] == 'GET'"], ["descriptors[0]['app.id']"], None)
_KUADRANT_REQ = [[("req.method", "GET"), ("app.id", "1")]]


def test_kuadrant_check_returns_ok_correctly():
    """:209-268 — checks never count: a limit of 1 answers OK twice."""
    h = CpuHarness([_KUADRANT_LIMIT])
    for _ in range(2):
        assert h.call(R.CHECK_RATE_LIMIT, [_req("test_namespace", _KUADRANT_REQ, 1)])[0] == (0, (R.CODE_OK, []))


def test_kuadrant_check_returns_overlimit_correctly():
    """:270-323 — max 0: the first check is already over the limit."""
    h = CpuHarness([("test_namespace", 0, 60) + _KUADRANT_LIMIT[3:]])
    assert h.call(R.CHECK_RATE_LIMIT, [_req("test_namespace", _KUADRANT_REQ, 1)])[0] == (0, (R.CODE_OVER_LIMIT, []))


def test_kuadrant_check_returns_ok_when_no_limits_apply():
    """:325-357"""
    h = CpuHarness([])
    assert h.call(R.CHECK_RATE_LIMIT, [_req("test_namespace", [[("req.method", "GET")]], 1)])[0] == (0, (R.CODE_OK, []))


def test_kuadrant_check_returns_unknown_when_domain_is_empty():
    """:359-389"""
    h = CpuHarness([])
    assert h.call(R.CHECK_RATE_LIMIT, [_req("", [[("req.method", "GET")]], 1)])[0] == (0, (R.CODE_UNKNOWN, []))


def test_kuadrant_check_takes_into_account_all_the_descriptors():
    """:391-471 — the max-0 limit needs descriptors[1].y == '2'."""
    h = CpuHarness([("test_namespace", 10, 60, ["descriptors[0].x == '1'"], ["descriptors[0].z"], None),
                    ("test_namespace", 0, 60, ["descriptors[0].x == '1'", "descriptors[1].y == '2'"], ["descriptors[0].z"], None)])
    r, = h.call(R.CHECK_RATE_LIMIT, [_req("test_namespace", [[("x", "1"), ("z", "1")], [("y", "2")]], 1)])
    assert r == (0, (R.CODE_OVER_LIMIT, []))


def test_kuadrant_report_returns_ok_correctly():
    """:487-537"""
    h = CpuHarness([_KUADRANT_LIMIT])
    assert h.call(R.REPORT, [_req("test_namespace", _KUADRANT_REQ, 1)])[0] == (0, (R.CODE_OK, []))


def test_kuadrant_report_going_overlimit_is_ok():
    """:539-589 — Report 20 hits against a limit of 5: still OK (update_counters never refuses), and the counter holds 20."""
    h = CpuHarness([("test_namespace", 5, 60) + _KUADRANT_LIMIT[3:]])
    assert h.call(R.REPORT, [_req("test_namespace", _KUADRANT_REQ, 20)])[0] == (0, (R.CODE_OK, []))
    assert h.last_plan["delta"].tolist() == [20]
    assert [row[3] for row in h.o.dump()] == [20]
    assert h.call(R.CHECK_RATE_LIMIT, [_req("test_namespace", _KUADRANT_REQ, 1)])[0] == (0, (R.CODE_OVER_LIMIT, []))


def test_kuadrant_report_returns_ok_when_no_limits_apply():
    """:591-619"""
    h = CpuHarness([])
    assert h.call(R.REPORT, [_req("test_namespace", [[("req.method", "GET")]], 1)])[0] == (0, (R.CODE_OK, []))


def test_kuadrant_report_returns_unknown_when_domain_is_empty():
    """:621-647"""
    h = CpuHarness([])
    assert h.call(R.REPORT, [_req("", [[("req.method", "GET")]], 1)])[0] == (0, (R.CODE_UNKNOWN, []))


def test_a_batch_is_finished_once():
    h = CpuHarness([("ns", 5, 60, [], [], None)])
    h.call(R.SHOULD_RATE_LIMIT, [_req("ns", [])] * 3)
    with pytest.raises(R.RlsError, match="no planned batch"):
        h.svc.finish(np.zeros(3, np.uint8), np.zeros(3, np.uint32), np.zeros(3, np.uint64), np.zeros(3, np.uint64))
    assert 'authorized_calls{limitador_namespace="ns"} 3' in h.svc.metrics()
