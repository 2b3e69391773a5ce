"""The N>1 path on GPUs: two NCCL ranks run exchange.LanePipelinedExchange with the engine's kernels
(bucket, lane put/gather, check_and_update through the C-ABI); every rank's reassembled verdicts must
equal the single-process oracle run over the global stream in (step, rank, index) order.  Needs two
GPUs on the box (skipped otherwise); tests/test_exchange_gloo.py covers the same protocol on CPU."""
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
from oracle import binding as ob  # noqa: E402

WORLD = 2
BATCH = 8192
N_STEPS = 7


def make_stream(rank):
    w = streams.c2_zipf_4limits(batch=BATCH, n_rows=5000, n_ns=24)
    return w, [w.batch_records(b * WORLD + rank) for b in range(N_STEPS)]


def _worker(rank, port, ret, lag):
    from limitador_b200 import Engine
    from limitador_b200.engine import MEM_DEVICE
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    w, batches = make_stream(rank)
    slot_cap = BATCH
    eng = Engine(capacity_rows=1 << 16, cells_per_row=w.cells_per_row, max_batch=WORLD * slot_cap, device=rank, flags=2)
    eng.limits_set(w.limits)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)

    class Ops:
        @staticmethod
        def fence(age):
            eng.fence() if age == 0 else eng.fence_call(age)

        @staticmethod
        def bucket(r, send, pos):
            eng.bucket_by_owner_padded_ptr(BATCH, r.data_ptr(), WORLD, slot_cap, send.data_ptr(), pos.data_ptr(),
                                           overflow.data_ptr())

        @staticmethod
        def lane_put(send, lane):
            eng.record_lane_put_ptr(WORLD * slot_cap, send.data_ptr(), lane.data_ptr())

        @staticmethod
        def decide(recv, verdict):
            eng.check_and_update_records_ptr(WORLD * slot_cap, recv.data_ptr(), verdict.data_ptr(), MEM_DEVICE,
                                             stride=w.cells_per_row)

        @staticmethod
        def lane_gather(recv, pos, out):
            eng.record_lane_gather_ptr(BATCH, recv.data_ptr(), pos.data_ptr(), out.data_ptr())

    ex = exchange.LanePipelinedExchange(WORLD, BATCH, slot_cap, dist, Ops, dev, lag=lag)
    d_recs = [torch.from_numpy(r.view(np.int64).reshape(-1, 4).copy()).to(dev) for r in batches]
    outs = [torch.full((BATCH,), 7, dtype=torch.uint8, device=dev) for _ in batches]
    torch.cuda.synchronize()
    for r, o in zip(d_recs[:4], outs[:4]):
        ex.step(r, o)
    ex.flush()  # mid-stream flush, then the pipeline refills
    for r, o in zip(d_recs[4:], outs[4:]):
        ex.step(r, o)
    ex.flush()
    eng.sync()
    torch.cuda.synchronize()
    assert int(overflow.item()) == 0
    ret[rank] = np.concatenate([o.cpu().numpy() for o in outs])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("lag", [2, 3])
def test_two_gpu_lane_pipelined_exchange_matches_global_oracle(lag):
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(port, ret, lag), nprocs=WORLD, join=True)
    w, b0 = make_stream(0)
    _, b1 = make_stream(1)
    orc = ob.Oracle(1 << 16)
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
