import json, sys
from collections import defaultdict
ev = json.load(open(sys.argv[1]))
# hot events: (name, end, seq=list length, ns); pair start/end by order per length
starts = defaultdict(list); durs = []
for name, end, seq, ns in ev:
    if name != "hot": continue
    if not end: starts[seq].append(ns)
    elif starts[seq]:
        durs.append((seq, ns - starts[seq].pop(0)))
durs.sort()
print("hot rows traced:", len(durs))
import itertools
for lo, hi in ((0, 64), (64, 256), (256, 1024), (1024, 4096), (4096, 1 << 30)):
    d = [x[1] for x in durs if lo <= x[0] < hi]
    if d: print(f"  list length [{lo},{hi}): n={len(d)} mean {sum(d)/len(d)/1e3:.1f} us  max {max(d)/1e3:.1f} us")
