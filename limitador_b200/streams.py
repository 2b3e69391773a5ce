"""Deterministic synthetic request streams for BASELINE.json's configs.

A stream is a limits table plus batches of 32-byte `rl_record`s — requests AFTER limit
matching, i.e. exactly what crosses the CounterStorage boundary
(limitador/src/storage/mod.rs:279-292).  Seed 42 mirrors the reference bench
(limitador/benches/bench.rs:21).  Pure numpy; used by bench.py, the tests and smoke().
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List

import numpy as np

from .engine import LIMIT_DESC_DTYPE, RECORD_DTYPE

T0_US = 1_700_000_000_000_000  # stream epoch (µs)
SEED = 42


def _zipf_cdf(n_items: int, alpha: float) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), alpha)
    c = np.cumsum(w)
    return c / c[-1]


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser: spreads dense ranks over the 64-bit key space (bijective)."""
    x = x.astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


@dataclass
class Workload:
    name: str
    description: str
    limits: np.ndarray          # LIMIT_DESC_DTYPE
    cells_per_row: int
    capacity_rows: int          # >= 2x live rows (no eviction regime)
    batch: int
    limits_per_request: int     # L, for the algorithmic-bytes formula
    gen: Callable[[int], np.ndarray]  # batch index -> records

    def batch_records(self, b: int) -> np.ndarray:
        return self.gen(b)


def _now_column(b: int, batch: int, req_per_us: int = 1000, jump_every: int = 64) -> np.ndarray:
    """now_us[i] = t0 + i/R (monotone, many equal stamps) + a 1 s jump every 64 batches."""
    i = np.arange(batch, dtype=np.uint64) + np.uint64(b) * np.uint64(batch)
    return np.uint64(T0_US) + i // np.uint64(req_per_us) + np.uint64(1_000_000) * np.uint64(b // jump_every)


def c1_bench_like(batch: int = 65536, n_keys: int = 1000) -> Workload:
    """C1: limitador/benches analogue — 1 namespace, 1 limit 10/60 s, 1k uniform keys."""
    limits = np.array([(0, 0, 1, 1, 10, 60_000_000)], dtype=LIMIT_DESC_DTYPE)

    def gen(b: int) -> np.ndarray:
        rng = np.random.default_rng([SEED, 1, b])
        r = np.zeros(batch, dtype=RECORD_DTYPE)
        r["ns_id"] = 0
        r["hits_addend"] = 1
        r["key_lo"] = rng.integers(1, n_keys + 1, size=batch, dtype=np.uint64)
        r["key_hi"] = 0
        i = np.arange(batch, dtype=np.uint64) + np.uint64(b) * np.uint64(batch)
        r["now_us"] = np.uint64(T0_US) + np.uint64(10) * i  # 10 µs apart: one 60 s rollover per 6M requests
        return r

    return Workload("C1", "1 namespace, 1 limit (10/60s), 1k keys uniform, delta=1", limits, 1, 4096, batch, 1, gen)


def c2_zipf_4limits(batch: int = 65536, n_rows: int = 1_000_000, n_ns: int = 64, alpha: float = 1.1) -> Workload:
    """C2 (the headline single-GPU config): 64 namespaces x 4 limits sharing one variable,
    1M distinct (namespace, key) rows => 4M counters, key rank ~ Zipf(1.1), batch 65536."""
    windows = [1, 60, 3600, 86400]
    maxes = [5, 100, 2000, 20000]
    descs = []
    for ns in range(n_ns):
        for k in range(4):
            descs.append((ns * 4 + k, ns, 1, 1, maxes[k], windows[k] * 1_000_000))
    limits = np.array(descs, dtype=LIMIT_DESC_DTYPE)
    cdf = _zipf_cdf(n_rows, alpha)

    def gen(b: int) -> np.ndarray:
        rng = np.random.default_rng([SEED, 2, b])
        rank = np.searchsorted(cdf, rng.random(batch), side="left").astype(np.uint64)
        np.minimum(rank, np.uint64(n_rows - 1), out=rank)
        r = np.zeros(batch, dtype=RECORD_DTYPE)
        r["ns_id"] = (rank % np.uint64(n_ns)).astype(np.uint32)
        r["hits_addend"] = 1
        r["key_lo"] = _mix(rank + np.uint64(1))
        r["key_hi"] = 0
        r["now_us"] = _now_column(b, batch)
        return r

    cap = 1 << int(np.ceil(np.log2(2 * n_rows)))
    return Workload("C2", f"{n_ns} namespaces x 4 limits, {n_rows} keys Zipf({alpha}), batch={batch}",
                    limits, 7, cap, batch, 4, gen)


def c3_uniform_1limit(batch: int = 1 << 20, n_keys: int = 16_000_000) -> Workload:
    """C3: 1 limit (100/60 s), 16M uniform keys, batch 1M.  Reference fixed-window semantics
    (the reference has no sliding window: atomic_expiring_value.rs:36-42 is the only rule)."""
    limits = np.array([(0, 0, 1, 1, 100, 60_000_000)], dtype=LIMIT_DESC_DTYPE)

    def gen(b: int) -> np.ndarray:
        rng = np.random.default_rng([SEED, 3, b])
        r = np.zeros(batch, dtype=RECORD_DTYPE)
        r["ns_id"] = 0
        r["hits_addend"] = 1
        r["key_lo"] = _mix(rng.integers(1, n_keys + 1, size=batch, dtype=np.uint64))
        r["key_hi"] = 0
        r["now_us"] = _now_column(b, batch)
        return r

    cap = 1 << int(np.ceil(np.log2(2 * n_keys)))
    return Workload("C3", f"1 limit (100/60s), {n_keys} keys uniform, batch={batch}", limits, 1, cap, batch, 1, gen)


def c4_namespace_sharded(batch: int = 1 << 20, n_keys: int = 128_000_000, n_ns: int = 10_000,
                         hot: bool = False) -> Workload:
    """C4 / C5 (8-GPU configs): 10k namespaces with 1-4 limits each, namespace popularity
    Zipf(1.0), keys uniform inside a namespace (C4) or Zipf(0.7) with 50% of the traffic
    forced onto 100 fixed keys (C5, hot=True)."""
    rng0 = np.random.default_rng([SEED, 4])
    n_lim = rng0.integers(1, 5, size=n_ns)
    windows = [1, 60, 3600, 86400]
    descs = []
    lid = 0
    for ns in range(n_ns):
        for k in range(int(n_lim[ns])):
            mx = (1 << 32) if hot else [50, 1000, 20000, 200000][k]
            descs.append((lid, ns, 1, 1, mx, windows[k] * 1_000_000))
            lid += 1
    limits = np.array(descs, dtype=LIMIT_DESC_DTYPE)
    ns_cdf = _zipf_cdf(n_ns, 1.0)
    keys_per_ns = max(1, n_keys // n_ns)
    key_cdf = _zipf_cdf(min(keys_per_ns, 1 << 20), 0.7) if hot else None

    def gen(b: int) -> np.ndarray:
        rng = np.random.default_rng([SEED, 5 if hot else 4, b])
        ns = np.searchsorted(ns_cdf, rng.random(batch), side="left").astype(np.uint64)
        np.minimum(ns, np.uint64(n_ns - 1), out=ns)
        if hot:
            k = np.searchsorted(key_cdf, rng.random(batch), side="left").astype(np.uint64)
            hot_sel = rng.random(batch) < 0.5
            hot_idx = rng.integers(0, 100, size=batch, dtype=np.uint64)
            ns = np.where(hot_sel, hot_idx * np.uint64(97) % np.uint64(n_ns), ns)
            k = np.where(hot_sel, np.uint64(0), k)
        else:
            k = rng.integers(0, keys_per_ns, size=batch, dtype=np.uint64)
        r = np.zeros(batch, dtype=RECORD_DTYPE)
        r["ns_id"] = ns.astype(np.uint32)
        r["hits_addend"] = 1
        r["key_lo"] = _mix(ns * np.uint64(keys_per_ns) + k + np.uint64(1))
        r["key_hi"] = 0
        r["now_us"] = _now_column(b, batch)
        return r

    cap = 1 << int(np.ceil(np.log2(2 * n_keys)))
    name = "C5" if hot else "C4"
    return Workload(name, f"{n_ns} namespaces x 1-4 limits, {n_keys} keys, {'hot-key' if hot else 'uniform'}",
                    limits, 7, cap, batch, 2, gen)


WORKLOADS: Dict[str, Callable[..., Workload]] = {
    "C1": c1_bench_like,
    "C2": c2_zipf_4limits,
    "C3": c3_uniform_1limit,
    "C4": c4_namespace_sharded,
    "C5": lambda **kw: c4_namespace_sharded(hot=True, **kw),
}


def algorithmic_bytes(n_decisions: int, n_counters_read: int, n_counters_written: int,
                      load_counters: bool = False) -> int:
    """33 B per decision (32 B record in + 1 B verdict out), 32 B per counter
    examined (16 B key + value + expiry), 16 B per counter written back."""
    b = 33 * n_decisions + 32 * n_counters_read + 16 * n_counters_written
    if load_counters:
        b += 16 * n_counters_read
    return b


# ---------------------------------------------------------------------------------------
# Device-side generation (torch is plumbing here: it only fills HBM with synthetic records
# so that bench.py's timed region starts with inputs resident on the GPU).
def _mix_torch(x):
    """splitmix64 finaliser on an int64 tensor (two's-complement wrap == uint64 arithmetic)."""
    import torch

    def lsr(v, k):  # logical shift right on int64
        return (v >> k) & ((1 << (64 - k)) - 1)

    def s64(c):  # uint64 constant -> signed
        return c - (1 << 64) if c >= (1 << 63) else c

    x = (x ^ lsr(x, 30)) * s64(0xBF58476D1CE4E5B9)
    x = (x ^ lsr(x, 27)) * s64(0x94D049BB133111EB)
    return x ^ lsr(x, 31)


def c2_device_stream(n_batches: int, batch: int, device, n_rows: int = 1_000_000, n_ns: int = 64,
                     alpha: float = 1.1, first_batch: int = 0, ns_base: int = 0, seed: int = SEED,
                     chunk_batches: int = 64):
    """C2 records generated on `device`: int64 tensor [n_batches, batch, 4] whose bytes are
    rl_record[batch] per batch (word0 = ns_id | hits_addend<<32, key_lo, key_hi, now_us).
    Same distribution as c2_zipf_4limits (different RNG stream: torch's, seeded)."""
    import torch

    cdf = torch.from_numpy(_zipf_cdf(n_rows, alpha)).to(device)
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + first_batch + 7919 * ns_base)
    out = torch.empty((n_batches, batch, 4), dtype=torch.int64, device=device)
    for b0 in range(0, n_batches, chunk_batches):
        nb = min(chunk_batches, n_batches - b0)
        u = torch.rand(nb * batch, dtype=torch.float64, device=device, generator=g)
        rank = torch.searchsorted(cdf, u).clamp_(max=n_rows - 1)
        ns = rank % n_ns + ns_base
        gb = torch.arange(first_batch + b0, first_batch + b0 + nb, device=device, dtype=torch.int64)
        i = gb.repeat_interleave(batch) * batch + torch.arange(batch, device=device, dtype=torch.int64).repeat(nb)
        now = T0_US + i // 1000 + 1_000_000 * (i // batch // 64)
        o = out[b0:b0 + nb].view(nb * batch, 4)
        o[:, 0] = ns | (1 << 32)
        o[:, 1] = _mix_torch(rank + 1 + ns_base * n_rows)
        o[:, 2] = 0
        o[:, 3] = now
    return out


def c3_device_stream(n_batches: int, batch: int, device, n_keys: int = 16_000_000, first_batch: int = 0,
                     seed: int = SEED, chunk_batches: int = 8):
    """C3 records on `device` (1 limit, uniform keys)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + first_batch + 13)
    out = torch.empty((n_batches, batch, 4), dtype=torch.int64, device=device)
    for b0 in range(0, n_batches, chunk_batches):
        nb = min(chunk_batches, n_batches - b0)
        k = torch.randint(1, n_keys + 1, (nb * batch,), dtype=torch.int64, device=device, generator=g)
        gb = torch.arange(first_batch + b0, first_batch + b0 + nb, device=device, dtype=torch.int64)
        i = gb.repeat_interleave(batch) * batch + torch.arange(batch, device=device, dtype=torch.int64).repeat(nb)
        o = out[b0:b0 + nb].view(nb * batch, 4)
        o[:, 0] = 1 << 32
        o[:, 1] = _mix_torch(k)
        o[:, 2] = 0
        o[:, 3] = T0_US + i // 1000 + 1_000_000 * (i // batch // 64)
    return out


def c4_device_stream(n_batches: int, batch: int, device, n_keys: int = 128_000_000, n_ns: int = 10_000, hot: bool = False,
                     first_batch: int = 0, seed: int = SEED, chunk_batches: int = 4):
    """C4 / C5 records on `device`: namespace popularity Zipf(1.0) over n_ns namespaces, keys uniform inside a
    namespace (C4) or Zipf(0.7) with 50 % of the traffic forced onto 100 fixed keys (C5, hot=True) — the same
    distributions as c4_namespace_sharded (torch's RNG, seeded), limits from c4_namespace_sharded(...).limits."""
    import torch

    ns_cdf = torch.from_numpy(_zipf_cdf(n_ns, 1.0)).to(device)
    keys_per_ns = max(1, n_keys // n_ns)
    key_cdf = torch.from_numpy(_zipf_cdf(min(keys_per_ns, 1 << 20), 0.7)).to(device) if hot else None
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + first_batch + (29 if hot else 17))
    out = torch.empty((n_batches, batch, 4), dtype=torch.int64, device=device)
    for b0 in range(0, n_batches, chunk_batches):
        nb = min(chunk_batches, n_batches - b0)
        m = nb * batch
        ns = torch.searchsorted(ns_cdf, torch.rand(m, dtype=torch.float64, device=device, generator=g)).clamp_(max=n_ns - 1)
        if hot:
            k = torch.searchsorted(key_cdf, torch.rand(m, dtype=torch.float64, device=device, generator=g)).clamp_(max=key_cdf.numel() - 1)
            hot_sel = torch.rand(m, device=device, generator=g) < 0.5
            hot_idx = torch.randint(0, 100, (m,), dtype=torch.int64, device=device, generator=g)
            ns = torch.where(hot_sel, hot_idx * 97 % n_ns, ns)
            k = torch.where(hot_sel, torch.zeros_like(k), k)
        else:
            k = torch.randint(0, keys_per_ns, (m,), dtype=torch.int64, device=device, generator=g)
        gb = torch.arange(first_batch + b0, first_batch + b0 + nb, device=device, dtype=torch.int64)
        i = gb.repeat_interleave(batch) * batch + torch.arange(batch, device=device, dtype=torch.int64).repeat(nb)
        o = out[b0:b0 + nb].view(m, 4)
        o[:, 0] = ns | (1 << 32)
        o[:, 1] = _mix_torch(ns * keys_per_ns + k + 1)
        o[:, 2] = 0
        o[:, 3] = T0_US + i // 1000 + 1_000_000 * (i // batch // 64)
    return out
