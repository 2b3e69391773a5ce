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
