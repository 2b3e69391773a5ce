"""The N>1 path on CPU: two gloo ranks run the namespace-sharded step (exchange.py) with numpy
bucketing and the oracle as the per-shard decider; the reassembled verdicts must equal the
single-process oracle run over the global stream in (rank, index) order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from limitador_b200 import exchange, streams  # noqa: E402
from limitador_b200.engine import RECORD_DTYPE, load_library  # noqa: E402
from oracle import binding as ob  # noqa: E402

WORLD = 2
N_BATCHES = 3
BATCH = 4096


def make_stream(rank):
    w = streams.c2_zipf_4limits(batch=BATCH, n_rows=3000, n_ns=16)
    rng = np.random.default_rng(100 + rank)
    out = []
    for b in range(N_BATCHES):
        r = w.batch_records(b * WORLD + rank)
        r["hits_addend"] = rng.choice([1, 1, 2], size=BATCH)
        out.append(r)
    return w, out


def _worker(rank, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    lib = load_library()
    w, batches = make_stream(rank)
    orc = ob.Oracle(1 << 14)
    for d in w.limits:
        orc.limit_set(int(d["limit_id"]), int(d["ns_id"]), int(d["max_value"]), int(d["window_us"]), bool(d["qualified"]))
    verdicts = []
    for recs_np in batches:
        recs = torch.from_numpy(recs_np.view(np.int64).reshape(-1, 4).copy())

        def bucket(t):
            a = t.numpy().view(RECORD_DTYPE).reshape(-1)
            owners = np.array([lib.rl_owner_of(int(ns), WORLD) for ns in a["ns_id"]], dtype=np.int64)
            perm, src, counts = exchange.stable_bucket_numpy(t.numpy(), owners, WORLD)
            return torch.from_numpy(perm.copy()), torch.from_numpy(src.copy()), counts

        def decide(buf, m, verdict):
            a = buf[:m].numpy().view(RECORD_DTYPE).reshape(-1)
            assert all(lib.rl_owner_of(int(ns), WORLD) == rank for ns in np.unique(a["ns_id"]))
            lim, _, _, _ = orc.batch_records(0, a)
            verdict[:m] = torch.from_numpy(lim)

        def unpermute(vb, src, out):
            out[src.long()] = vb[:len(src)]

        out = torch.zeros(BATCH, dtype=torch.uint8)
        exchange.sharded_step(recs, WORLD, dist, bucket, decide, unpermute, out)
        verdicts.append(out.numpy().copy())
    ret[rank] = np.concatenate(verdicts)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_step_matches_global_oracle():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(port, ret), nprocs=WORLD, join=True)
    # global reference: per step, the stream is rank 0's slice followed by rank 1's
    w, b0 = make_stream(0)
    _, b1 = make_stream(1)
    orc = ob.Oracle(1 << 14)
    for d in w.limits:
        orc.limit_set(int(d["limit_id"]), int(d["ns_id"]), int(d["max_value"]), int(d["window_us"]), bool(d["qualified"]))
    want0, want1 = [], []
    for s0, s1 in zip(b0, b1):
        lim, _, _, _ = orc.batch_records(0, np.concatenate([s0, s1]))
        want0.append(lim[:BATCH])
        want1.append(lim[BATCH:])
    assert np.array_equal(ret[0], np.concatenate(want0))
    assert np.array_equal(ret[1], np.concatenate(want1))
    assert 0 < int(ret[0].sum()) < len(ret[0])
