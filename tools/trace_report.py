#!/usr/bin/env python
"""Timeline of pipelined / sharded steps from a device-side event trace (bench.py --trace, Engine.trace_dump).
Prints, per kernel, its mean duration, and the mean offset of its start / end from the step's first event,
over the steps of the middle half of the trace."""
import json
import sys
from collections import defaultdict


def report(path):
    ev = json.load(open(path))
    steps = defaultdict(dict)
    for name, end, seq, ns in ev:
        steps[seq][(name, end)] = ns
    seqs = sorted(steps)
    seqs = seqs[len(seqs) // 4: 3 * len(seqs) // 4] or seqs
    order = ["xcount", "xscatter", "xwait", "front", "main", "xreturn", "xwaitv", "xgather"]
    print(f"== {path}: {len(ev)} events, {len(steps)} steps, using {len(seqs)}")
    t_first = {s: min(steps[s].values()) for s in seqs}
    for name in order:
        d, so, eo = [], [], []
        for s in seqs:
            a, b = steps[s].get((name, 0)), steps[s].get((name, 1))
            if a is not None and b is not None:
                d.append(b - a)
            if a is not None:
                so.append(a - t_first[s])
            if b is not None:
                eo.append(b - t_first[s])
        if so or eo:
            f = lambda v: f"{sum(v) / len(v) / 1e3:8.1f}" if v else "       -"
            print(f"  {name:9s} dur {f(d)} us   start +{f(so)} us   end +{f(eo)} us")
    # period: start-to-start of consecutive steps' first events, and end of main
    per = [t_first[b] - t_first[a] for a, b in zip(seqs, seqs[1:]) if b == a + 1]
    if per:
        print(f"  step period (first event to first event): {sum(per) / len(per) / 1e3:.1f} us")
    me = [steps[b][("main", 1)] - steps[a][("main", 1)] for a, b in zip(seqs, seqs[1:])
          if b == a + 1 and ("main", 1) in steps[a] and ("main", 1) in steps[b]]
    if me:
        print(f"  main end to main end: {sum(me) / len(me) / 1e3:.1f} us")
    gaps = [steps[b][("main", 0)] - steps[a][("main", 1)] for a, b in zip(seqs, seqs[1:])
            if b == a + 1 and ("main", 1) in steps[a] and ("main", 0) in steps[b]]
    if gaps:
        print(f"  idle between main(s) end and main(s+1) start: {sum(gaps) / len(gaps) / 1e3:.1f} us")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        report(p)
