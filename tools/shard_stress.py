"""Stress of the peer exchange on ONE GPU: `world` engines in one process, rank-at-a-time issue order
(send+decide+collect per rank, as the one-process-per-GPU deployment issues them), many steps."""
import os
import sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limitador_b200 import Engine, streams
from limitador_b200.engine import Shard

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
lag = int(sys.argv[4]) if len(sys.argv) > 4 else 2
mode = sys.argv[5] if len(sys.argv) > 5 else "rank"
w = streams.WORKLOADS["C2"](batch=batch, n_rows=1_000_000, n_ns=64 * world)
engs = [Engine(capacity_rows=1 << 22, cells_per_row=7, max_batch=world * batch, max_counters=world * batch, flags=2) for _ in range(world)]
for e in engs:
    e.limits_set(w.limits)
sh = [Shard(engs[r], r, world, batch, lag) for r in range(world)]
for s in sh:
    s.connect_ptrs([x.slab for x in sh])
pool = 32
recs = [streams.c2_device_stream(pool, batch, "cuda", n_rows=1_000_000, n_ns=64 * world, seed=42 + r) for r in range(world)]
outs = [torch.zeros((pool, batch), dtype=torch.uint8, device="cuda") for _ in range(world)]
torch.cuda.synchronize()
try:
    for st in range(steps):
        if mode == "rank":
            for r in range(world):
                sh[r].step(batch, recs[r][st % pool].data_ptr(), outs[r][st % pool].data_ptr())
        else:
            for r in range(world):
                sh[r].send(batch, recs[r][st % pool].data_ptr(), outs[r][st % pool].data_ptr())
            for r in range(world):
                sh[r].decide()
            for r in range(world):
                sh[r].collect()
    for s in sh:
        s.flush()
    for e in engs:
        e.sync()
    print(f"shard_stress world={world} steps={steps} batch={batch} lag={lag} mode={mode}: OK")
except Exception as ex:
    print(f"shard_stress world={world} steps={steps} batch={batch} lag={lag} mode={mode}: FAILED at host step {st}: {ex}")
    for r, x in enumerate(sh):
        print("    rank", r, x.debug())
    for e in engs:
        try:
            e.sync()
        except Exception as ex2:
            print("   ", ex2)
