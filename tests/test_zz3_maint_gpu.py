"""rl_compact and rl_ns_metrics_* on the GPU through the C-ABI (the kernels' logic is also run on the host under
tests/emu/cuda_shim.h: tests/test_maint_emu.py).  Sorted last on purpose: these entry points are new."""
import numpy as np
import pytest

from limitador_b200 import Engine
from limitador_b200.engine import LIMIT_DESC_DTYPE, NONE, RECORD_DTYPE, pack_records16
from tests import helpers as H
from tests.test_maint_emu import metrics_by_numpy

pytestmark = pytest.mark.gpu
S = 1_000_000
RL_FLAG_PIPELINE = 2


def _limits(cells):
    # short windows so that a sweep empties most rows; one unqualified namespace; one namespace with a long window
    rows = []
    lid = 0
    for ns in range(6):
        q = 0 if ns == 4 else 1
        for c in range(cells if ns % 2 == 0 else 1):
            win = (3600 if ns == 5 else [1, 2, 5][c % 3]) * S
            rows.append((lid, ns, 1 if q else 0, q, [3, 10, 50][c % 3], win))
            lid += 1
    return np.array(rows, dtype=LIMIT_DESC_DTYPE)


def _records(descs, n, seed, t0, n_keys):
    rng = np.random.default_rng(seed)
    r = np.zeros(n, dtype=RECORD_DTYPE)
    r["ns_id"] = rng.choice(sorted({int(d["ns_id"]) for d in descs}) + [17], size=n)
    r["hits_addend"] = rng.choice([1, 1, 2, 5], size=n)
    r["key_lo"] = rng.integers(1, n_keys + 1, size=n)
    r["now_us"] = t0 + np.arange(n) * 3
    return r


@pytest.mark.parametrize("cells,regions", [(1, 8), (3, 4), (7, 16)])
def test_compact_after_sweep_is_invisible_and_the_hot_path_finds_every_row_again(cells, regions):
    descs = _limits(cells)
    e = Engine(capacity_rows=1 << 13, cells_per_row=cells, max_batch=1 << 14, regions=regions)
    e.limits_set(descs)
    o = H.oracle_with_limits(descs)
    t = H.T0
    for b in range(4):
        recs = _records(descs, 6000, 100 * cells + b, t, n_keys=700)
        got = e.check_and_update_records(recs, False, stride=cells)
        want = o.batch_records(0, recs, False, cells)
        assert got[0].tolist() == want[0].tolist() and got[1].tolist() == want[1].tolist()
        t += 400_000
    t += 6 * S  # every short window is over: the sweep tombstones most rows
    assert e.sweep(t) == o.invalidate_expired(t) > 1000
    before = e.dump()
    assert H.normalise_dump(before, descs) == H.normalise_dump(o.dump(), descs)
    st = e.compact(5)
    assert st["regions"] == regions and st["regions_rebuilt"] > 0 and st["rows_tombstoned"] > 1000
    assert st["rows_reclaimed"] >= st["rows_tombstoned"] * st["regions_rebuilt"] // regions // 2
    assert e.dump() == before, "rl_compact changed the observable state"
    st2 = e.compact(5)
    assert st2["regions_rebuilt"] == 0 and st2["rows_tombstoned"] < st["rows_tombstoned"]
    # the hot path over the rebuilt table: old keys (still live ones and reclaimed ones) and new keys
    for b in range(4):
        recs = _records(descs, 6000, 900 * cells + b, t, n_keys=1000)
        got = e.check_and_update_records(recs, True, stride=cells)
        want = o.batch_records(0, recs, True, cells)
        for k in range(4):
            assert got[k].tolist() == want[k].tolist()
        assert H.normalise_dump(e.dump(), descs) == H.normalise_dump(o.dump(), descs)
        t += 900_000
    e.sweep(t + 10 * S)
    o.invalidate_expired(t + 10 * S)
    e.compact(0)  # rebuild every region that holds a tombstone
    assert e.compact(0)["rows_tombstoned"] == 0
    assert H.normalise_dump(e.dump(), descs) == H.normalise_dump(o.dump(), descs)
    recs = _records(descs, 6000, 7, t + 11 * S, n_keys=1000)
    assert e.check_and_update_records(recs, False, stride=cells)[0].tolist() == o.batch_records(0, recs, False, cells)[0].tolist()


def test_compact_on_a_table_without_tombstones_does_nothing():
    descs = _limits(1)
    e = Engine(capacity_rows=1 << 10, cells_per_row=1, max_batch=4096, regions=4)
    e.limits_set(descs)
    e.check_and_update_records(_records(descs, 2000, 1, H.T0, 100), False, stride=1)  # <= 500 rows in a table of 1024
    before = e.dump()
    st = e.compact(0)
    assert st["regions_rebuilt"] == 0 and st["rows_tombstoned"] == 0 and st["rows_live"] > 100 and st["rows_moved"] == 0
    assert e.dump() == before


def _expected(recs, lim, fl, ns_cap, limits_cap):
    return metrics_by_numpy(recs["ns_id"].astype(np.int64), recs["hits_addend"], lim, fl.astype(np.int64), ns_cap, limits_cap)


@pytest.mark.parametrize("flags", [0, RL_FLAG_PIPELINE])
def test_ns_metrics_accumulate_behind_every_record_call(flags):
    cells = 3
    descs = _limits(cells)
    n_lim = int(descs["limit_id"].max()) + 1
    e = Engine(capacity_rows=1 << 12, cells_per_row=cells, max_batch=1 << 14, regions=8, flags=flags)
    e.limits_set(descs)
    e.ns_metrics_enable(True)
    ns_cap = 32
    tot = [np.zeros(ns_cap, dtype=np.uint64) for _ in range(3)] + [np.zeros(n_lim, dtype=np.uint64)]
    t = H.T0
    for b in range(5):
        recs = _records(descs, 5000 + 37 * b, 50 + b, t, n_keys=200)
        lim, fl, _, _ = e.check_and_update_records(recs, False, stride=cells)
        ac, ah, lc, bl, dropped = _expected(recs, lim, fl, ns_cap, n_lim)
        assert dropped == 0
        for acc, x in zip(tot, (ac, ah, lc, bl)):
            acc += x
        t += 300_000
    # the 16-byte wire form goes through the same hook
    recs = _records(descs, 4000, 99, t, n_keys=200)
    recs["now_us"] = t
    lim, fl = e.check_and_update_compact(pack_records16(recs), t)
    for acc, x in zip(tot, _expected(recs, lim, fl, ns_cap, n_lim)[:4]):
        acc += x
    m = e.ns_metrics_read(ns_cap, n_lim)
    assert m["authorized_calls"].tolist() == tot[0].tolist()
    assert m["authorized_hits"].tolist() == tot[1].tolist()
    assert m["limited_calls"].tolist() == tot[2].tolist()
    assert m["limited_by_limit"].tolist() == tot[3].tolist()
    assert m["dropped"] == 0 and int(tot[2].sum()) > 1000 and int(tot[0][17]) > 100  # namespace 17 has no limits: allowed
    # read with reset, then an explicit accumulate of an already decided batch (what a sharded source rank does)
    assert e.ns_metrics_read(ns_cap, n_lim, reset=True)["limited_calls"].tolist() == tot[2].tolist()
    e.ns_metrics_enable(False)
    recs = _records(descs, 3000, 5, t + S, n_keys=200)
    lim, fl, _, _ = e.check_and_update_records(recs, False, stride=cells)
    assert int(e.ns_metrics_read(ns_cap)["authorized_calls"].sum()) == 0  # disabled: nothing was added
    lim2 = lim.copy()
    lim2[::50] = 0xFF  # error verdicts are not counted
    e.ns_metrics_accumulate(recs, lim2, fl)
    e.ns_metrics_accumulate(pack_records16(recs), lim2, None)
    ac, ah, lc, bl, dropped = _expected(recs, lim2, fl, ns_cap, n_lim)
    m = e.ns_metrics_read(ns_cap, n_lim)
    assert m["authorized_calls"].tolist() == (2 * ac).tolist() and m["authorized_hits"].tolist() == (2 * ah).tolist()
    assert m["limited_calls"].tolist() == (2 * lc).tolist() and m["limited_by_limit"].tolist() == bl.tolist()
    assert m["dropped"] == 2 * dropped == 2 * len(lim2[::50])


def test_ns_metrics_behind_pipelined_device_calls():
    """RL_FLAG_PIPELINE + device-resident records: the reduction kernel rides the replay stream of every call."""
    import torch
    cells = 3
    descs = _limits(cells)
    n_lim = int(descs["limit_id"].max()) + 1
    e = Engine(capacity_rows=1 << 12, cells_per_row=cells, max_batch=1 << 14, regions=8, flags=RL_FLAG_PIPELINE)
    e.limits_set(descs)
    e.ns_metrics_enable(True)
    n, nb, ns_cap = 8192, 6, 32
    recs = [_records(descs, n, 300 + b, H.T0 + b * 250_000, n_keys=150) for b in range(nb)]
    d_recs = [torch.from_numpy(r.view(np.int64).reshape(-1, 4).copy()).cuda() for r in recs]
    d_lim = [torch.full((n,), 9, dtype=torch.uint8, device="cuda") for _ in recs]
    d_first = [torch.zeros((n,), dtype=torch.int32, device="cuda") for _ in recs]
    torch.cuda.synchronize()
    for i in range(nb):
        e.check_and_update_records_ptr(n, d_recs[i].data_ptr(), d_lim[i].data_ptr(), 1, out_first_ptr=d_first[i].data_ptr(), stride=cells)
    e.fence()
    e.sync()
    o = H.oracle_with_limits(descs)
    tot = [np.zeros(ns_cap, dtype=np.uint64) for _ in range(3)] + [np.zeros(n_lim, dtype=np.uint64)]
    for i in range(nb):
        lim = d_lim[i].cpu().numpy()
        fl = d_first[i].cpu().numpy().view(np.uint32)
        want = o.batch_records(0, recs[i], False, cells)
        assert lim.tolist() == want[0].tolist() and fl.tolist() == want[1].tolist()
        for acc, x in zip(tot, _expected(recs[i], lim, fl, ns_cap, n_lim)[:4]):
            acc += x
    m = e.ns_metrics_read(ns_cap, n_lim)
    assert m["authorized_calls"].tolist() == tot[0].tolist() and m["authorized_hits"].tolist() == tot[1].tolist()
    assert m["limited_calls"].tolist() == tot[2].tolist() and m["limited_by_limit"].tolist() == tot[3].tolist()
    assert int(tot[2].sum()) > 5000 and m["dropped"] == 0
