"""rl_rls_serve on the GPU (include/rl_rls.h): the wire surface end to end through ONE engine call per batch, against the
same CPU stages wrapped around the oracle (tests/test_rls.py).  Sorted last on purpose: these entry points are new."""
import numpy as np
import pytest

from limitador_b200 import matcher as MT
from limitador_b200 import rls as R
from tests.test_rls import T0, CpuHarness, _gateway, _req


@pytest.mark.gpu
def test_serve_through_the_engine_equals_the_cpu_stages_around_the_oracle():
    """rl_rls_serve (plan -> ONE rl_check_and_update_batch on the GPU -> finish) against plan -> oracle -> finish:
    same response bytes, same metrics; then CheckRateLimit and Report through the engine."""
    from limitador_b200 import Engine
    limits, reqs = _gateway(9, 3000)
    msgs = [_req(ns, descs, hits) for ns, descs, hits in reqs]
    buf, off = R.pack_requests(msgs)
    for headers in (R.HEADERS_DRAFT_VERSION_03, R.HEADERS_NONE):
        h = CpuHarness(limits, headers=headers, threads=2)
        m = MT.Matcher()
        e = Engine(capacity_rows=1 << 12, cells_per_row=3, max_batch=4096)
        e.limits_set(np.array([m.add_limit(*l) for l in limits]))
        svc = R.RlsService(m, e, headers, 2)
        for step in range(3):
            now = T0 + step * 7_000_000
            want = h.call(R.SHOULD_RATE_LIMIT, msgs, now)
            svc.serve(R.SHOULD_RATE_LIMIT, buf, off, now)
            got = [(g, R.decode_response(b) if g == 0 else None) for g, b in svc.responses()]
            assert got == want
        assert svc.metrics() == h.svc.metrics()
        assert svc.timings()["store_us"] > 0
        want = h.call(R.CHECK_RATE_LIMIT, msgs[:500], T0 + 30_000_000)
        svc.serve(R.CHECK_RATE_LIMIT, *R.pack_requests(msgs[:500]), T0 + 30_000_000)
        assert [(g, R.decode_response(b)) for g, b in svc.responses()] == want
        want = h.call(R.REPORT, msgs[:500], T0 + 31_000_000)
        svc.serve(R.REPORT, *R.pack_requests(msgs[:500]), T0 + 31_000_000)
        assert [(g, R.decode_response(b)) for g, b in svc.responses()] == want
        from tests import helpers as H
        assert H.normalise_dump(e.dump(), np.array(h.descs)) == H.normalise_dump(h.o.dump(), np.array(h.descs))


@pytest.mark.gpu
def test_front_with_the_matcher_inside_concurrent_callers_linearise():
    """rl_front_check_and_update_bindings: 6 threads submit (namespace, context) requests; each runs the native matcher
    on its own thread and queues the counters; replaying the requests through the Python mirror of RateLimiter in the
    front's drain order (out_seq) reproduces every verdict, limit name, remaining and ttl."""
    import threading
    from limitador_b200 import Engine, Front
    from limitador_b200 import limiter as LM
    from limitador_b200.matcher import front_check_and_update
    from tests import helpers as H
    limits, reqs = _gateway(21, 1800)
    reqs = [r for r in reqs if r[0] != ""]
    m = MT.Matcher()
    e = Engine(capacity_rows=1 << 12, cells_per_row=3, max_batch=4096)
    e.limits_set(np.array([m.add_limit(*l) for l in limits]))
    front = Front(e, max_batch=256, max_delay_us=200)
    n_threads = 6
    results = [[] for _ in range(n_threads)]
    now = T0 + 5

    def worker(t):
        for ns, descs, hits in reqs[t::n_threads]:
            got = front_check_and_update(front, m, ns, None, [dict(d) for d in descs], hits or 1, now, True)
            results[t].append((got[2], ns, descs, hits or 1, got))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    st = front.stats()
    front.close()
    rl = LM.RateLimiter(H.OracleStorage(), clock=lambda: now)
    for ns, mx, secs, conds, vars_, name in limits:
        rl.add_limit(LM.Limit(ns, mx, secs, conds, vars_, name=name))
    queued = sorted([r for rs in results for r in rs if len(r[4][3])], key=lambda r: r[0])
    assert [r[0] for r in queued] == list(range(len(queued))) and st["requests"] == len(queued)
    assert st["batches"] < st["requests"], "requests were never coalesced"
    n_lim = 0
    for seq, ns, descs, hits, (lim, first, _, ctrs, rem, ttl) in queued:
        w = rl.check_rate_limited_and_update(ns, LM.Context({}, [dict(d) for d in descs]), hits, True)
        assert lim == w.limited, seq
        n_lim += lim
        if lim:
            assert m.limit_name(first) == w.limit_name
        got = sorted((int(c["limit_id"]), int(r), int(t)) for c, r, t in zip(ctrs, rem, ttl))
        assert got == sorted((c.limit_id, c.remaining, c.expires_in_us) for c in w.counters)
    for rs in results:  # requests nothing applied to never reached the store and are allowed
        for r in rs:
            if len(r[4][3]) == 0:
                assert r[4][0] is False and r[4][1] is None
    assert 100 < n_lim < len(queued)
