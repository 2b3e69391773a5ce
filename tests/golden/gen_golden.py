"""Generates tests/golden/*.npz: request streams with the oracle's verdicts, named limits,
remaining/ttl and final counter table.  The reference ships no golden vectors and cannot be
built here (Rust), so these are produced by the CPU oracle, which is itself pinned by the
reference's known-answer tests (tests/test_oracle_kats.py).  Run from the repo root:
    python tests/golden/gen_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from limitador_b200 import streams  # noqa: E402
from oracle import binding as ob  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def records_case(name, w, n_batches, load_counters):
    o = H.oracle_with_limits(w.limits, 1 << 16)
    recs, lim, fl, rem, ttl = [], [], [], [], []
    for b in range(n_batches):
        r = w.batch_records(b)
        a = o.batch_records(0, r, load_counters, w.cells_per_row)
        recs.append(r)
        lim.append(a[0])
        fl.append(a[1])
        rem.append(a[2])
        ttl.append(a[3])
    d = o.dump_arrays()
    order = np.lexsort((d[2], d[1], d[0]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), limits=w.limits, cells=w.cells_per_row,
                        capacity=w.capacity_rows, recs=np.stack(recs), limited=np.stack(lim), first=np.stack(fl),
                        remaining=np.stack(rem), ttl=np.stack(ttl), load_counters=load_counters,
                        dump_limit=d[0][order], dump_lo=d[1][order], dump_hi=d[2][order], dump_value=d[3][order],
                        dump_expiry=d[4][order])


def csr_case(name, seed, cells):
    descs = H.mixed_limits(n_ns=12, seed=seed)
    o = H.oracle_with_limits(descs)
    batches = []
    for b in range(4):
        off, ctrs, delta, now = H.random_csr_stream(descs, 1500, 7000 + 10 * seed + b, n_keys=4, monotone=(b != 2))
        lim, fl, rem, ttl = o.batch_csr(0, off, ctrs, delta, now, True)
        batches.append(dict(off=off, ctrs=ctrs, delta=delta, now=now, limited=lim, first=fl, remaining=rem, ttl=ttl))
    d = o.dump_arrays()
    order = np.lexsort((d[2], d[1], d[0]))
    flat = {f"b{i}_{k}": v for i, bt in enumerate(batches) for k, v in bt.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), limits=descs, cells=cells, n_batches=len(batches),
                        dump_limit=d[0][order], dump_lo=d[1][order], dump_hi=d[2][order], dump_value=d[3][order],
                        dump_expiry=d[4][order], **flat)


if __name__ == "__main__":
    records_case("c2_small", streams.c2_zipf_4limits(batch=4096, n_rows=3000, n_ns=16), 6, True)
    records_case("c1_small", streams.c1_bench_like(batch=4096, n_keys=200), 4, False)
    csr_case("csr_mixed_cells3", 3, 3)
    print("golden files written to", OUT)
