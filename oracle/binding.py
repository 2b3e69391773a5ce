"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE, NOT THE PRODUCT).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  limitador_b200 (the product) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblimitador_oracle.so")

NONE = 0xFFFFFFFF

COUNTER_DTYPE = np.dtype(
    [("limit_id", "<u4"), ("_pad", "<u4"), ("key_lo", "<u8"), ("key_hi", "<u8")]
)
RECORD_DTYPE = np.dtype(
    [("ns_id", "<u4"), ("hits_addend", "<u4"), ("key_lo", "<u8"), ("key_hi", "<u8"), ("now_us", "<u8")]
)
LIMIT_DESC_DTYPE = np.dtype(
    [("limit_id", "<u4"), ("ns_id", "<u4"), ("max_value", "<u8"), ("window_us", "<u8"),
     ("qualified", "<u4"), ("_pad", "<u4")]
)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "limitador_oracle.c")
    hdr = os.path.join(_HERE, "limitador_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "liblimitador_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.lo_create.restype = vp
        L.lo_create.argtypes = [u64]
        L.lo_destroy.argtypes = [vp]
        L.lo_limit_set.argtypes = [vp, u32, u32, u64, u64, i32]
        L.lo_limit_delete.argtypes = [vp, u32]
        L.lo_check_and_update.argtypes = [vp, vp, u32, u64, i32, u64, vp, vp, vp]
        L.lo_is_within_limits.argtypes = [vp, vp, u64, u64]
        L.lo_is_rate_limited.argtypes = [vp, vp, u32, u64, u64, vp]
        L.lo_update_counter.argtypes = [vp, vp, u64, u64]
        L.lo_update_counters.argtypes = [vp, vp, u32, u64, u64]
        L.lo_batch_csr.argtypes = [vp, i32, u64, vp, vp, vp, vp, i32, vp, vp, vp, vp]
        L.lo_batch_records.argtypes = [vp, i32, u64, vp, i32, u32, vp, vp, vp, vp]
        L.lo_get_counters.restype = u64
        L.lo_get_counters.argtypes = [vp, vp, u32, u64, u64, vp, vp, vp, vp, vp]
        L.lo_delete_counters.argtypes = [vp, vp, u32]
        L.lo_clear.argtypes = [vp]
        L.lo_invalidate_expired.restype = u64
        L.lo_invalidate_expired.argtypes = [vp, u64]
        L.lo_dump.restype = u64
        L.lo_dump.argtypes = [vp, u64, vp, vp, vp, vp, vp]
        L.lo_size.restype = u64
        L.lo_size.argtypes = [vp]
        L.lo_mt_create.restype = vp
        L.lo_mt_create.argtypes = [vp, u32, u32, u64]
        L.lo_mt_run.restype = C.c_double
        L.lo_mt_run.argtypes = [vp, u64, vp, vp]
        L.lo_mt_destroy.argtypes = [vp]
        L.lo_mt_destroy.restype = None
        L.lo_mt_pinned.argtypes = [vp]
        L.lo_mt_pinned.restype = u32
        L.lo_bench_records_mt.restype = C.c_double
        L.lo_bench_records_mt.argtypes = [vp, u32, u64, vp, u32, u64, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def counters(items) -> np.ndarray:
    """[(limit_id, key_lo, key_hi), ...] -> structured array."""
    a = np.zeros(len(items), dtype=COUNTER_DTYPE)
    for i, (lid, lo, hi) in enumerate(items):
        a[i] = (lid, 0, lo, hi)
    return a


class Oracle:
    MODE_CHECK_AND_UPDATE, MODE_IS_RATE_LIMITED, MODE_UPDATE = 0, 1, 2

    def __init__(self, capacity_hint: int = 1024):
        self._h = lib().lo_create(capacity_hint)
        if not self._h:
            raise MemoryError("lo_create failed")

    def close(self):
        if self._h:
            lib().lo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- limits --
    def limit_set(self, limit_id, ns_id, max_value, window_us, qualified):
        r = lib().lo_limit_set(self._h, limit_id, ns_id, max_value, window_us, int(bool(qualified)))
        if r != 0:
            raise ValueError("lo_limit_set: identity of a live limit id changed")

    def limit_delete(self, limit_id):
        lib().lo_limit_delete(self._h, limit_id)

    # -- single calls --
    def check_and_update(self, ctrs, delta, load_counters, now_us):
        ctrs = np.ascontiguousarray(ctrs, dtype=COUNTER_DTYPE)
        m = len(ctrs)
        fl = C.c_uint32(NONE)
        rem = np.zeros(m, dtype=np.uint64)
        ttl = np.zeros(m, dtype=np.uint64)
        r = lib().lo_check_and_update(self._h, _p(ctrs), m, delta, int(load_counters), now_us,
                                      C.addressof(fl), _p(rem), _p(ttl))
        if r < 0:
            raise RuntimeError(f"lo_check_and_update error {r}")
        return bool(r), (None if fl.value == NONE else int(fl.value)), rem, ttl

    def is_within_limits(self, ctr, delta, now_us):
        ctr = np.ascontiguousarray(ctr, dtype=COUNTER_DTYPE)
        r = lib().lo_is_within_limits(self._h, _p(ctr), delta, now_us)
        if r < 0:
            raise RuntimeError(f"lo_is_within_limits error {r}")
        return bool(r)

    def is_rate_limited(self, ctrs, delta, now_us):
        ctrs = np.ascontiguousarray(ctrs, dtype=COUNTER_DTYPE)
        fl = C.c_uint32(NONE)
        r = lib().lo_is_rate_limited(self._h, _p(ctrs), len(ctrs), delta, now_us, C.addressof(fl))
        if r < 0:
            raise RuntimeError(f"lo_is_rate_limited error {r}")
        return bool(r), (None if fl.value == NONE else int(fl.value))

    def update_counters(self, ctrs, delta, now_us):
        ctrs = np.ascontiguousarray(ctrs, dtype=COUNTER_DTYPE)
        r = lib().lo_update_counters(self._h, _p(ctrs), len(ctrs), delta, now_us)
        if r < 0:
            raise RuntimeError(f"lo_update_counters error {r}")

    # -- batches --
    def batch_csr(self, mode, off, ctrs, delta, now_us, load_counters=False):
        off = np.ascontiguousarray(off, dtype=np.uint32)
        ctrs = np.ascontiguousarray(ctrs, dtype=COUNTER_DTYPE)
        delta = np.ascontiguousarray(delta, dtype=np.uint64)
        now_us = np.ascontiguousarray(now_us, dtype=np.uint64)
        n = len(delta)
        lim = np.zeros(n, dtype=np.uint8)
        fl = np.full(n, NONE, dtype=np.uint32)
        rem = np.zeros(len(ctrs), dtype=np.uint64)
        ttl = np.zeros(len(ctrs), dtype=np.uint64)
        r = lib().lo_batch_csr(self._h, mode, n, _p(off), _p(ctrs), _p(delta), _p(now_us),
                               int(load_counters), _p(lim), _p(fl), _p(rem), _p(ttl))
        if r < 0:
            raise RuntimeError(f"lo_batch_csr error {r}")
        return lim, fl, rem, ttl

    def batch_records(self, mode, recs, load_counters=False, stride=1):
        recs = np.ascontiguousarray(recs, dtype=RECORD_DTYPE)
        n = len(recs)
        lim = np.zeros(n, dtype=np.uint8)
        fl = np.full(n, NONE, dtype=np.uint32)
        rem = np.zeros(n * stride, dtype=np.uint64)
        ttl = np.zeros(n * stride, dtype=np.uint64)
        r = lib().lo_batch_records(self._h, mode, n, _p(recs), int(load_counters), stride,
                                   _p(lim), _p(fl), _p(rem), _p(ttl))
        if r < 0:
            raise RuntimeError(f"lo_batch_records error {r}")
        return lim, fl, rem, ttl

    # -- maintenance --
    def get_counters(self, limit_ids, now_us):
        ids = np.ascontiguousarray(limit_ids, dtype=np.uint32)
        cap = int(lib().lo_size(self._h)) + 1
        lid = np.zeros(cap, dtype=np.uint32)
        lo = np.zeros(cap, dtype=np.uint64)
        hi = np.zeros(cap, dtype=np.uint64)
        rem = np.zeros(cap, dtype=np.uint64)
        ttl = np.zeros(cap, dtype=np.uint64)
        cnt = lib().lo_get_counters(self._h, _p(ids), len(ids), now_us, cap, _p(lid), _p(lo),
                                    _p(hi), _p(rem), _p(ttl))
        return sorted(zip(lid[:cnt].tolist(), lo[:cnt].tolist(), hi[:cnt].tolist(),
                          rem[:cnt].tolist(), ttl[:cnt].tolist()))

    def delete_counters(self, limit_ids):
        ids = np.ascontiguousarray(limit_ids, dtype=np.uint32)
        lib().lo_delete_counters(self._h, _p(ids), len(ids))

    def clear(self):
        lib().lo_clear(self._h)

    def invalidate_expired(self, now_us):
        return int(lib().lo_invalidate_expired(self._h, now_us))

    def dump(self):
        """Sorted list of (limit_id, key_lo, key_hi, value, expiry_us) for every entry."""
        cap = int(lib().lo_size(self._h)) + 1
        lid = np.zeros(cap, dtype=np.uint32)
        lo = np.zeros(cap, dtype=np.uint64)
        hi = np.zeros(cap, dtype=np.uint64)
        val = np.zeros(cap, dtype=np.uint64)
        exp = np.zeros(cap, dtype=np.uint64)
        cnt = lib().lo_dump(self._h, cap, _p(lid), _p(lo), _p(hi), _p(val), _p(exp))
        return sorted(zip(lid[:cnt].tolist(), lo[:cnt].tolist(), hi[:cnt].tolist(),
                          val[:cnt].tolist(), exp[:cnt].tolist()))

    def dump_arrays(self):
        cap = int(lib().lo_size(self._h)) + 1
        lid = np.zeros(cap, dtype=np.uint32)
        lo = np.zeros(cap, dtype=np.uint64)
        hi = np.zeros(cap, dtype=np.uint64)
        val = np.zeros(cap, dtype=np.uint64)
        exp = np.zeros(cap, dtype=np.uint64)
        cnt = int(lib().lo_dump(self._h, cap, _p(lid), _p(lo), _p(hi), _p(val), _p(exp)))
        return lid[:cnt], lo[:cnt], hi[:cnt], val[:cnt], exp[:cnt]

    def size(self):
        return int(lib().lo_size(self._h))


def bench_records_mt(limit_descs, recs, threads, capacity_hint):
    """Returns (elapsed_seconds, verdicts) of the multi-threaded CPU baseline."""
    limit_descs = np.ascontiguousarray(limit_descs, dtype=LIMIT_DESC_DTYPE)
    recs = np.ascontiguousarray(recs, dtype=RECORD_DTYPE)
    out = np.zeros(len(recs), dtype=np.uint8)
    t = lib().lo_bench_records_mt(_p(limit_descs), len(limit_descs), len(recs), _p(recs), threads,
                                  capacity_hint, _p(out))
    return float(t), out


class OracleMT:
    """T persistent oracles sharded by namespace: the multi-core CPU baseline."""

    def __init__(self, limit_descs, threads, capacity_hint):
        limit_descs = np.ascontiguousarray(limit_descs, dtype=LIMIT_DESC_DTYPE)
        self.threads = threads
        self._h = lib().lo_mt_create(_p(limit_descs), len(limit_descs), threads, capacity_hint)

    def run(self, recs):
        """Process recs in stream order per namespace; returns (seconds, verdicts)."""
        recs = np.ascontiguousarray(recs, dtype=RECORD_DTYPE)
        out = np.zeros(len(recs), dtype=np.uint8)
        t = lib().lo_mt_run(self._h, len(recs), _p(recs), _p(out))
        return float(t), out

    @property
    def pinned(self) -> int:
        """worker threads pinned to a CPU of their own"""
        return int(lib().lo_mt_pinned(self._h))

    def close(self):
        if self._h:
            lib().lo_mt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
