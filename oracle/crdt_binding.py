"""ctypes binding of the CRDT oracle (oracle/crdt_oracle.c).  TEST INFRASTRUCTURE, NOT THE PRODUCT: only tests/ import it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcrdt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "crdt_oracle.c")
    if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libcrdt_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.cro_create.restype = vp
        L.cro_create.argtypes = [u32, u32]
        L.cro_destroy.argtypes = [vp]
        L.cro_inc_actor_at.argtypes = [vp, u64, u64, u32, u64, u64, u64]
        L.cro_inc_actor_at.restype = None
        L.cro_merge_at.argtypes = [vp, u64, u64, u64, vp, vp, u32, u64]
        L.cro_merge_at.restype = None
        L.cro_read_at.argtypes = [vp, u64, u64, u64, vp]
        L.cro_read_at.restype = u64
        L.cro_export.argtypes = [vp, u64, u64, vp, vp, vp, vp]
        L.cro_export.restype = u64
        L.cro_dump.argtypes = [vp, u64, vp, vp, vp, vp]
        L.cro_dump.restype = u64
        _lib = L
    return _lib


class CrdtOracle:
    """One replica: every counter a CrCounterValue<actor index> with explicit clocks."""

    def __init__(self, actors: int, self_actor: int):
        self.actors, self.self_actor = actors, self_actor
        self._h = lib().cro_create(actors, self_actor)
        assert self._h

    def __del__(self):
        try:
            if self._h:
                lib().cro_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def inc_actor_at(self, key, actor, inc, window_us, when):
        lib().cro_inc_actor_at(self._h, key[0], key[1], actor, inc, window_us, when)

    def inc_at(self, key, inc, window_us, when):
        self.inc_actor_at(key, self.self_actor, inc, window_us, when)

    def merge_at(self, key, expiry, values: dict, when):
        a = np.array(list(values.keys()), dtype=np.uint32)
        v = np.array(list(values.values()), dtype=np.uint64)
        lib().cro_merge_at(self._h, key[0], key[1], expiry, a.ctypes.data, v.ctypes.data, len(a), when)

    def read_at(self, key, when):
        return int(lib().cro_read_at(self._h, key[0], key[1], when, None))

    def expiry(self, key):
        e = C.c_uint64()
        lib().cro_read_at(self._h, key[0], key[1], 0, C.byref(e))
        return e.value

    def export(self, when, cap=1 << 20):
        lo, hi, val, exp = (np.zeros(cap, dtype=np.uint64) for _ in range(4))
        n = lib().cro_export(self._h, when, cap, lo.ctypes.data, hi.ctypes.data, val.ctypes.data, exp.ctypes.data)
        return sorted(zip(lo[:n].tolist(), hi[:n].tolist(), val[:n].tolist(), exp[:n].tolist()))

    def dump(self, cap=1 << 20):
        lo, hi, exp = (np.zeros(cap, dtype=np.uint64) for _ in range(3))
        vals = np.zeros(cap * self.actors, dtype=np.uint64)
        n = lib().cro_dump(self._h, cap, lo.ctypes.data, hi.ctypes.data, exp.ctypes.data, vals.ctypes.data)
        v = vals[:n * self.actors].reshape(n, self.actors)
        return sorted((int(lo[i]), int(hi[i]), int(exp[i]), tuple(v[i].tolist())) for i in range(n))
