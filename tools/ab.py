#!/usr/bin/env python
"""A/B of build-time variants of librl_engine.so on a GPU box.

    python tools/ab.py build  name1:RL_EXP_PAD_AGG=1  name2:RL_MID_CTAS=5,RL_EXP_ROW_PREFETCH=1 ...
        builds limitador_b200/variants/librl_engine_<name>.so (they travel with the gpurun snapshot) and prints
        the shell snippet to run under gpurun: the default library first and last, every variant in between,
        each through `bench.py --no-cpu-baseline` with RL_ENGINE_LIB pointing at it.
    python tools/ab.py table [gpurun_out]
        one line per gpurun_out/ab_<name>.json: value, us/step, k_main us, e2e.

Parity first: run `python tests/fuzz_gpu.py --seconds 60` with RL_ENGINE_LIB=<variant> before believing a number.
"""
import concurrent.futures as cf
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(specs):
    from limitador_b200 import build as B
    jobs = {}
    for s in specs:
        name, _, defs = s.partition(":")
        jobs[name] = [d for d in defs.split(",") if d]
    with cf.ThreadPoolExecutor(max(1, min(6, len(jobs)))) as ex:
        for path in ex.map(lambda kv: B.build_variant(*kv), jobs.items()):
            print("built", os.path.relpath(path, ROOT), file=sys.stderr)
    B.build_engine()
    steps = os.environ.get("AB_STEPS", "500")
    run = ('B() { timeout 200 python bench.py --steps %s --warmup 5 --no-cpu-baseline > gpurun_out/ab_$1.json '
           '2> gpurun_out/ab_$1.err; }; ' % steps)
    body = "mkdir -p gpurun_out; " + run + "B base; "
    for name in jobs:
        body += f"RL_ENGINE_LIB=$PWD/limitador_b200/variants/librl_engine_{name}.so B {name}; "
    body += "B base2; python tools/ab.py table"
    print(f"/usr/local/graft/bin/gpurun --timeout 900 -- '{body}'")


def table(d):
    rows = []
    for f in sorted(glob.glob(os.path.join(d, "ab_*.json")), key=os.path.getmtime):
        try:
            j = json.load(open(f))
        except Exception as ex:
            rows.append((os.path.basename(f), f"unreadable: {ex}"))
            continue
        r, e = j.get("roofline") or {}, j.get("e2e") or {}
        rows.append((os.path.basename(f)[3:-5],
                     f"{j['value'] / 1e6:8.0f} M dec/s  {j['ms_per_step'] * 1e3:6.1f} us/step  "
                     f"k_main {r.get('avg_launch_ms', 0) * 1e3:5.1f} us  e2e {e.get('value', 0) / 1e6:6.0f} M "
                     f"({e.get('h2d_frac_of_copy_rate', 0):.2f} of the H2D copy rate)"))
    for name, line in rows:
        print(f"{name:24s} {line}")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "build":
        build(sys.argv[2:])
    elif len(sys.argv) >= 2 and sys.argv[1] == "table":
        table(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out"))
    else:
        print(__doc__)
        sys.exit(2)
