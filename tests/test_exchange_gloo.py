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


def make_stream(rank, n_batches=N_BATCHES):
    w = streams.c2_zipf_4limits(batch=BATCH, n_rows=3000, n_ns=16)
    rng = np.random.default_rng(100 + rank)
    out = []
    for b in range(n_batches):
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


class _NumpyOps:
    """LanePipelinedExchange's device work on CPU tensors: numpy bucketing, the oracle as decider."""

    def __init__(self, lib, orc, rank, slot_cap):
        self.lib, self.orc, self.rank, self.slot_cap = lib, orc, rank, slot_cap
        self.decided = 0

    def fence(self, age):
        pass

    def bucket(self, recs, send, pos):
        a = recs.numpy().view(RECORD_DTYPE).reshape(-1)
        owners = np.array([self.lib.rl_owner_of(int(ns), WORLD) for ns in a["ns_id"]], dtype=np.int64)
        out = send.numpy()
        out[:] = -1
        p = pos.numpy()
        for o in range(WORLD):
            idx = np.flatnonzero(owners == o)
            assert len(idx) <= self.slot_cap
            out[o * self.slot_cap:o * self.slot_cap + len(idx)] = recs.numpy()[idx]
            p[idx] = o * self.slot_cap + np.arange(len(idx))

    def lane_put(self, send, lane):
        send.numpy().view(np.uint8).reshape(-1, 32)[:, 23] = lane.numpy()

    def decide(self, recv, verdict):
        a = recv.numpy().view(RECORD_DTYPE).reshape(-1)
        real = a["ns_id"] != 0xFFFFFFFF
        assert all(self.lib.rl_owner_of(int(ns), WORLD) == self.rank for ns in np.unique(a["ns_id"][real]))
        lim, _, _, _ = self.orc.batch_records(0, a)  # padding = a namespace without limits: never limited
        verdict[:] = torch.from_numpy(lim)
        self.decided += int(real.sum())

    def lane_gather(self, recv, pos, out):
        out[:] = torch.from_numpy(recv.numpy().view(np.uint8).reshape(-1, 32)[pos.numpy().astype(np.int64), 23].copy())


def _lane_worker(rank, port, ret, lag):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    lib = load_library()
    w, batches = make_stream(rank, n_batches=6)
    orc = ob.Oracle(1 << 14)
    for d in w.limits:
        orc.limit_set(int(d["limit_id"]), int(d["ns_id"]), int(d["max_value"]), int(d["window_us"]), bool(d["qualified"]))
    slot_cap = BATCH  # worst case: every record of a rank has the same owner
    ops = _NumpyOps(lib, orc, rank, slot_cap)
    ex = exchange.LanePipelinedExchange(WORLD, BATCH, slot_cap, dist, ops, "cpu", lag=lag)
    outs = [torch.full((BATCH,), 7, dtype=torch.uint8) for _ in batches]
    delivered = []
    for recs_np, out in zip(batches[:4], outs[:4]):
        d = ex.step(torch.from_numpy(recs_np.view(np.int64).reshape(-1, 4).copy()), out)
        delivered.append(d is not None)
    assert delivered == [i >= lag for i in range(4)]
    assert len(ex.flush()) == min(lag, 4)  # mid-stream flush, then the pipeline refills
    for i, (recs_np, out) in enumerate(zip(batches[4:], outs[4:])):
        d = ex.step(torch.from_numpy(recs_np.view(np.int64).reshape(-1, 4).copy()), out)
        assert (d is not None) == (i >= lag)
    assert len(ex.flush()) == min(lag, 2) and ex.flush() == []
    ret[rank] = np.concatenate([o.numpy() for o in outs])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lag", [1, 2, 3])
def test_two_rank_lane_pipelined_exchange_matches_global_oracle(lag):
    """One all-to-all per step, verdicts returned `lag` steps later in the records' lane byte."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_lane_worker, args=(port, ret, lag), nprocs=WORLD, join=True)
    w, b0 = make_stream(0, n_batches=6)
    _, b1 = make_stream(1, n_batches=6)
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


def test_exchange_blocks_are_sized_from_the_observed_traffic():
    """bench.py sizes the fixed exchange blocks from a sample of the stream: largest (rank -> owner) share x 1.2."""
    lib = load_library()
    world, batch, steps = 4, 8192, 6
    recs = streams.c2_device_stream(steps, batch, "cpu", n_rows=50000, n_ns=64 * world)
    a = recs.numpy().view(RECORD_DTYPE).reshape(steps, batch)
    owner = np.array([lib.rl_owner_of(ns, world) for ns in range(64 * world)])
    want = max(int(np.bincount(owner[a["ns_id"][s]], minlength=world).max()) for s in range(steps))
    lut = torch.tensor([exchange.owner_of(ns, world) for ns in range(64 * world)], dtype=torch.int64)
    assert lut.tolist() == owner.tolist()
    got = exchange.observed_block_max(recs, lut, world)
    assert got == want and batch // world < got < batch
    cap = exchange.slot_cap_for(got, batch)
    assert cap % 256 == 0 and got * 1.2 <= cap + 255 and cap >= got and cap <= batch
    assert exchange.slot_cap_for(batch, batch) == batch


def _placement_worker(rank, port, ret):
    """Two gloo ranks: bench.place_namespaces over each rank's own stream (the all-reduce makes the observed load, hence
    the ids, equal on both), then the sharded step over the re-identified records with the oracle as the decider."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    import bench
    lib = load_library()
    n_ns = 64 * WORLD
    limits = bench.c2_limits(n_ns)
    recs = streams.c2_device_stream(N_BATCHES, BATCH, "cpu", n_rows=200_000, n_ns=n_ns, seed=streams.SEED + 1000 * rank)
    limits, summary = bench.place_namespaces(dist, WORLD, torch.device("cpu"), recs, limits, N_BATCHES, lambda m: None)
    orc = ob.Oracle(1 << 16)
    for d in limits:
        orc.limit_set(int(d["limit_id"]), int(d["ns_id"]), int(d["max_value"]), int(d["window_us"]), bool(d["qualified"]))
    verdicts, inbox = [], 0
    for b in range(N_BATCHES):
        def bucket(t):
            a = t.numpy().view(RECORD_DTYPE).reshape(-1)
            owners = np.array([lib.rl_owner_of(int(ns), WORLD) for ns in a["ns_id"]], dtype=np.int64)
            perm, src, counts = exchange.stable_bucket_numpy(t.numpy(), owners, WORLD)
            return torch.from_numpy(perm.copy()), torch.from_numpy(src.copy()), counts

        def decide(buf, m, verdict):
            nonlocal inbox
            a = buf[:m].numpy().view(RECORD_DTYPE).reshape(-1)
            inbox += m
            lim, _, _, _ = orc.batch_records(0, a)
            verdict[:m] = torch.from_numpy(lim)

        def unpermute(vb, src, out):
            out[src.long()] = vb[:len(src)]

        out = torch.zeros(BATCH, dtype=torch.uint8)
        exchange.sharded_step(recs[b].clone(), WORLD, dist, bucket, decide, unpermute, out)
        verdicts.append(out.numpy().copy())
    ret[rank] = (np.concatenate(verdicts), limits["ns_id"].copy(), recs.numpy().copy(), summary, inbox)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_placement_gives_both_ranks_the_same_ids_and_keeps_the_sharded_step_exact():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_placement_worker, args=(port, ret), nprocs=WORLD, join=True)
    (v0, ids0, recs0, sum0, in0), (v1, ids1, recs1, sum1, in1) = ret[0], ret[1]
    assert ids0.tolist() == ids1.tolist() and sum0 == sum1  # the same placement on every rank
    assert sum0["owner_load_max_over_mean"] < 1.02 <= sum0["owner_load_max_over_mean_if_ids_were_hashed_as_generated"]
    assert abs(in0 - in1) < 0.04 * (in0 + in1)  # the two owners decide (almost) the same number of requests
    # the global reference: ONE oracle over the re-identified records in (step, rank, index) order
    import bench
    limits = bench.c2_limits(64 * WORLD)
    limits["ns_id"] = ids0
    orc = ob.Oracle(1 << 16)
    for d in limits:
        orc.limit_set(int(d["limit_id"]), int(d["ns_id"]), int(d["max_value"]), int(d["window_us"]), bool(d["qualified"]))
    want0, want1 = [], []
    for b in range(N_BATCHES):
        glob = np.concatenate([recs0[b].view(RECORD_DTYPE).reshape(-1), recs1[b].view(RECORD_DTYPE).reshape(-1)])
        lim = orc.batch_records(0, glob)[0]
        want0.append(lim[:BATCH])
        want1.append(lim[BATCH:])
    assert np.array_equal(v0, np.concatenate(want0)) and np.array_equal(v1, np.concatenate(want1))
    assert 0 < int(v0.sum()) < len(v0)
