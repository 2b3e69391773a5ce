"""Randomised parity fuzzing on a GPU box (not collected by pytest; run as
`python tests/fuzz_gpu.py --seconds 120`).  Every round draws a limits table, an engine
geometry, the chained-commit knobs and a stream with a tiny key space, runs it through the
C-ABI in one of the call styles (records / CSR, host / pipelined device / async host,
check / update) and compares verdicts, outputs and the full table with the oracle."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from limitador_b200 import Engine  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_gpu_parity import single_row_limits  # noqa: E402


def explain(tag, b, k, got, want, recs):
    """Mismatch diagnostics: where, how many, and the request's place among those of its key."""
    bad = np.flatnonzero(got != want)
    lines = [f"{tag} batch {b} output {k}: {len(bad)} of {len(got)} differ"]
    per = len(got) // len(recs)
    for j in bad[:8]:
        i = int(j) // per
        same = np.flatnonzero((recs["ns_id"][:i + 1] == recs["ns_id"][i]) & (recs["key_lo"][:i + 1] == recs["key_lo"][i]) &
                              (recs["key_hi"][:i + 1] == recs["key_hi"][i]))
        lines.append(f"  idx {int(j)} (request {i}, slot {int(j) % per}): got {int(got[j])} want {int(want[j])}; ns {int(recs['ns_id'][i])} "
                     f"delta {int(recs['hits_addend'][i])} now {int(recs['now_us'][i])}; {len(same)}-th request of its key "
                     f"(previous at {same[-2] if len(same) > 1 else None})")
    return "\n".join(lines)


def one_round(seed: int, override=None) -> str:
    import torch
    rng = np.random.default_rng(seed)
    cells = int(rng.choice([1, 3, 7]))
    chunk = int(rng.choice([128, 256]))
    mult = int(rng.choice([1, 1, 2, 4]))
    pt = int(rng.choice([32, 128, 512]))
    style = str(rng.choice(["rec_host", "rec_dev_pipe", "rec_host_async", "csr_host", "rec_update"]))
    regions = int(rng.choice([1, 2, 4, 16]))
    n_keys = int(rng.choice([3, 10, 40, 400]))
    n = int(rng.choice([700, 3000, 9000]))
    lc = bool(rng.integers(0, 2))
    ov = override or {}
    chunk, mult, pt = int(ov.get("chunk", chunk)), int(ov.get("mult", mult)), int(ov.get("pt", pt))
    regions, lc = int(ov.get("regions", regions)), bool(int(ov.get("lc", lc)))
    os.environ["RL_CHUNK"], os.environ["RL_HEAVY_MULT"], os.environ["RL_PART_TARGET"] = str(chunk), str(mult), str(pt)
    flags = 2 if style in ("rec_dev_pipe", "rec_host_async") else 0
    if style == "csr_host":
        descs = H.mixed_limits(n_ns=12, seed=seed)
    else:
        descs = single_row_limits(cells, n_ns=int(rng.choice([2, 9])), seed=seed)
    e = Engine(capacity_rows=1 << 14, cells_per_row=cells, max_batch=1 << 14, regions=regions, flags=flags)
    e.limits_set(descs)
    o = H.oracle_with_limits(descs)
    tag = f"seed={seed} style={style} cells={cells} chunk={chunk} mult={mult} pt={pt} regions={regions} keys={n_keys} n={n} lc={lc}"
    nb = 5
    if style == "csr_host":
        for b in range(nb):
            off, ctrs, delta, now = H.random_csr_stream(descs, n // 3, seed * 100 + b, n_keys=max(2, n_keys // 4),
                                                        monotone=bool(b & 1))
            got = e.check_and_update_batch(off, ctrs, delta, now, lc)
            want = o.batch_csr(0, off, ctrs, delta, now, lc)
            for k in range(4 if lc else 2):
                assert np.array_equal(got[k], want[k]), f"{tag} batch {b} output {k}"
    elif style == "rec_update":
        for b in range(nb):
            recs = H.random_records(descs, n, seed * 100 + b, n_keys=n_keys, monotone=bool(b & 1))
            e.update_records(recs)
            o.batch_records(2, recs)
    else:
        batches = [H.random_records(descs, n, seed * 100 + b, n_keys=n_keys, monotone=bool(b & 1)) for b in range(nb)]
        if style == "rec_host":
            for b, recs in enumerate(batches):
                got = e.check_and_update_records(recs, lc, stride=cells)
                want = o.batch_records(0, recs, lc, cells)
                for k in range(4 if lc else 2):
                    assert np.array_equal(got[k], want[k]), explain(tag, b, k, got[k], want[k], recs)
        else:
            host = torch.stack([torch.from_numpy(r.view(np.int64).reshape(-1, 4).copy()) for r in batches])
            if style == "rec_dev_pipe":
                src = host.cuda()
                lim = torch.zeros((nb, n), dtype=torch.uint8, device="cuda")
                first = torch.zeros((nb, n), dtype=torch.int32, device="cuda")
                mem = 1
            else:
                src = host.pin_memory()
                lim = torch.zeros((nb, n), dtype=torch.uint8).pin_memory()
                first = torch.zeros((nb, n), dtype=torch.int32).pin_memory()
                mem = 2
            torch.cuda.synchronize()
            for b in range(nb):
                e.check_and_update_records_ptr(n, src[b].data_ptr(), lim[b].data_ptr(), mem,
                                               out_first_ptr=first[b].data_ptr(), stride=cells)
            e.sync()
            for b, recs in enumerate(batches):
                want = o.batch_records(0, recs)
                assert np.array_equal(lim[b].cpu().numpy(), want[0]), f"{tag} batch {b} verdicts"
                assert np.array_equal(first[b].cpu().numpy().astype(np.uint32), want[1]), f"{tag} batch {b} first"
    assert H.normalise_dump(e.dump(), descs) == H.normalise_dump(o.dump(), descs), f"{tag} table"
    e.close()
    return tag


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--first-seed", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0, help="run this one seed only")
    ap.add_argument("--set", action="append", default=[], help="override a drawn knob: chunk|mult|pt|regions|lc=VALUE")
    ap.add_argument("--keep-going", action="store_true")
    args = ap.parse_args()
    override = dict(kv.split("=", 1) for kv in args.set)
    if args.seed:
        try:
            print("ok:", one_round(args.seed, override))
        except AssertionError as ex:
            print("FUZZ MISMATCH:", ex)
            sys.exit(1)
        return
    t0, seed, done, bad = time.time(), args.first_seed, 0, 0
    while time.time() - t0 < args.seconds:
        try:
            one_round(seed, override)
        except AssertionError as ex:
            print("FUZZ MISMATCH:", ex)
            bad += 1
            if not args.keep_going:
                sys.exit(1)
        seed += 1
        done += 1
    print(f"fuzz {'ok' if not bad else 'FAILED'}: {done} rounds, {bad} mismatching, seeds {args.first_seed}..{seed - 1}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
