"""Namespace-sharded execution of one batch over `world` ranks.

Every counter a request touches belongs to the request's namespace
(limitador/src/lib.rs:512), so a request has exactly one owner rank,
`rl_owner_of(ns_id, world)`, and counters are never replicated.  One step =
  bucket my slice of the global batch by owner (stable)  ->  all-to-all of the 32-B records
  ->  decide locally (stream order = source rank, source index)  ->  all-to-all of the verdict
  bytes back  ->  restore request order.
The only collective on the path is that personalised all-to-all (NCCL over NVLink on GPUs).
The orchestration is backend-agnostic: bench.py plugs in the engine's kernels on GPU tensors,
tests/test_exchange_gloo.py plugs in numpy + the oracle on CPU tensors over gloo.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def sharded_step(recs, world: int, dist, bucket: Callable, decide: Callable, unpermute: Callable, out_limited,
                 recv_buf=None, verdict_recv=None, verdict_back=None):
    """Run one sharded step.

    recs         : [n, 4] int64 tensor (bytes = rl_record[n]) — this rank's slice of the global batch
    bucket(recs) : -> (send [n,4] tensor grouped by owner, stably; src_idx [n] tensor; counts: list[int] per owner)
    decide(recv, m, verdict) : run check_and_update on the first m received records, writing verdict[:m] (uint8)
    unpermute(verdict_back, src_idx, out_limited): out_limited[src_idx[i]] = verdict_back[i]
    Returns the number of requests this rank decided.
    """
    import torch

    n = recs.shape[0]
    send, src_idx, send_counts = bucket(recs)
    sc = torch.tensor(send_counts, dtype=torch.int64, device=recs.device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    recv_counts = rc.cpu().tolist()
    m = int(sum(recv_counts))
    if recv_buf is None or recv_buf.shape[0] < m:
        recv_buf = torch.empty((max(m, 1), 4), dtype=torch.int64, device=recs.device)
    if verdict_recv is None or verdict_recv.shape[0] < m:
        verdict_recv = torch.empty(max(m, 1), dtype=torch.uint8, device=recs.device)
    if verdict_back is None or verdict_back.shape[0] < n:
        verdict_back = torch.empty(max(n, 1), dtype=torch.uint8, device=recs.device)
    dist.all_to_all_single(recv_buf[:m], send, recv_counts, list(send_counts))
    if m:
        decide(recv_buf, m, verdict_recv)
    dist.all_to_all_single(verdict_back[:n], verdict_recv[:m], list(send_counts), recv_counts)
    unpermute(verdict_back, src_idx, out_limited)
    return m


def stable_bucket_numpy(recs_np: np.ndarray, owners: np.ndarray, world: int):
    """Reference (CPU) bucketing: stable sort by owner.  Returns (permuted records, src index, counts)."""
    order = np.argsort(owners, kind="stable")
    counts = np.bincount(owners, minlength=world).astype(np.int64)
    return recs_np[order], order.astype(np.int32), counts.tolist()
