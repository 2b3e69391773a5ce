"""ctypes binding of the replicated counter value (include/rl_crdt.h, csrc/rl_crdt.cu): the reference's
CrCounterValue (storage/distributed/cr_counter_value.rs) as a GPU table of per-actor values, with batched
inc / merge / read and the re-sync export.  No CPU fallback: constructing a CrdtTable without a CUDA device raises."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from . import engine as _eng

CRDT_SYMBOLS = (
    "rl_crdt_create", "rl_crdt_destroy", "rl_crdt_last_error", "rl_crdt_inc", "rl_crdt_merge", "rl_crdt_read",
    "rl_crdt_export", "rl_crdt_dump", "rl_crdt_kernel_launches", "rl_crdt_clear",
)
KEY_DTYPE = np.dtype([("lo", "<u8"), ("hi", "<u8")])
UPDATE_DTYPE = np.dtype([("key_lo", "<u8"), ("key_hi", "<u8"), ("expires_at_us", "<u8"), ("val_off", "<u4"), ("n_vals", "<u4")])


class CrdtConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("capacity_rows", C.c_uint64),
                ("actors", C.c_uint32), ("self_actor", C.c_uint32)]


class CrdtError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"[{'TRANSIENT' if status == 1 else 'FATAL'}] {msg}")
        self.status = status


def _lib():
    L = _eng.load_library()
    if getattr(L, "_rl_crdt_ready", False):
        return L
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.rl_crdt_create.argtypes = [C.POINTER(CrdtConfig), C.POINTER(vp)]
    L.rl_crdt_destroy.argtypes = [vp]
    L.rl_crdt_destroy.restype = None
    L.rl_crdt_last_error.argtypes = [vp]
    L.rl_crdt_last_error.restype = C.c_char_p
    L.rl_crdt_inc.argtypes = [vp, u64, vp, vp, vp, vp, u64, i32]
    L.rl_crdt_merge.argtypes = [vp, u64, vp, vp, vp, u64, u64, i32]
    L.rl_crdt_read.argtypes = [vp, u64, vp, u64, i32, vp, vp]
    L.rl_crdt_export.argtypes = [vp, u64, u64, vp, vp, vp, C.POINTER(u64)]
    L.rl_crdt_dump.argtypes = [vp, u64, vp, vp, vp, C.POINTER(u64)]
    L.rl_crdt_clear.argtypes = [vp]
    L.rl_crdt_kernel_launches.argtypes = [vp]
    L.rl_crdt_kernel_launches.restype = u64
    L._rl_crdt_ready = True
    return L


def keys_array(keys: Iterable[Tuple[int, int]]) -> np.ndarray:
    keys = list(keys)
    a = np.zeros(len(keys), dtype=KEY_DTYPE)
    for i, (lo, hi) in enumerate(keys):
        a[i] = (lo, hi)
    return a


def pack_updates(updates: Sequence[Tuple[Tuple[int, int], int, Dict[int, int]]]):
    """[(key, expires_at_us, {actor: value})] -> (UPDATE_DTYPE array, actors uint32, values uint64)."""
    ups = np.zeros(len(updates), dtype=UPDATE_DTYPE)
    actors: List[int] = []
    values: List[int] = []
    for i, (key, exp, vals) in enumerate(updates):
        ups[i] = (key[0], key[1], exp, len(actors), len(vals))
        actors += list(vals.keys())
        values += list(vals.values())
    return ups, np.array(actors, dtype=np.uint32), np.array(values, dtype=np.uint64)


class CrdtTable:
    def __init__(self, capacity_rows: int, actors: int, self_actor: int, device: int = 0):
        self._lib = _lib()
        self.actors, self.self_actor = actors, self_actor
        cfg = CrdtConfig(C.sizeof(CrdtConfig), device, capacity_rows, actors, self_actor)
        self._h = C.c_void_p()
        st = self._lib.rl_crdt_create(C.byref(cfg), C.byref(self._h))
        if st != 0:
            msg = "rl_crdt_create failed (no CUDA device, or a bad configuration): there is no CPU implementation"
            if self._h:
                msg = self._lib.rl_crdt_last_error(self._h).decode() or msg
                self._lib.rl_crdt_destroy(self._h)
                self._h = C.c_void_p()
            raise CrdtError(st, msg)

    def close(self):
        if self._h:
            self._lib.rl_crdt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != 0:
            raise CrdtError(st, self._lib.rl_crdt_last_error(self._h).decode())

    def inc(self, keys, actor, increment, window_us, now_us: int):
        keys = np.ascontiguousarray(keys, dtype=KEY_DTYPE)
        n = len(keys)
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(actor, dtype=np.uint32), (n,)))
        inc = np.ascontiguousarray(np.broadcast_to(np.asarray(increment, dtype=np.uint64), (n,)))
        win = np.ascontiguousarray(np.broadcast_to(np.asarray(window_us, dtype=np.uint64), (n,)))
        self._check(self._lib.rl_crdt_inc(self._h, n, keys.ctypes.data, a.ctypes.data, inc.ctypes.data, win.ctypes.data, now_us, 0))

    def merge(self, ups: np.ndarray, actors: np.ndarray, values: np.ndarray, now_us: int):
        ups = np.ascontiguousarray(ups, dtype=UPDATE_DTYPE)
        actors = np.ascontiguousarray(actors, dtype=np.uint32)
        values = np.ascontiguousarray(values, dtype=np.uint64)
        self._check(self._lib.rl_crdt_merge(self._h, len(ups), ups.ctypes.data, actors.ctypes.data if len(actors) else None,
                                            values.ctypes.data if len(values) else None, len(values), now_us, 0))

    def read(self, keys, now_us: int):
        keys = np.ascontiguousarray(keys, dtype=KEY_DTYPE)
        val = np.zeros(len(keys), dtype=np.uint64)
        exp = np.zeros(len(keys), dtype=np.uint64)
        self._check(self._lib.rl_crdt_read(self._h, len(keys), keys.ctypes.data, now_us, 0, val.ctypes.data, exp.ctypes.data))
        return val, exp

    def export(self, now_us: int, cap: int = 1 << 20):
        k = np.zeros(cap, dtype=KEY_DTYPE)
        val = np.zeros(cap, dtype=np.uint64)
        exp = np.zeros(cap, dtype=np.uint64)
        n = C.c_uint64()
        self._check(self._lib.rl_crdt_export(self._h, now_us, cap, k.ctypes.data, val.ctypes.data, exp.ctypes.data, C.byref(n)))
        m = min(n.value, cap)
        return sorted(zip(k["lo"][:m].tolist(), k["hi"][:m].tolist(), val[:m].tolist(), exp[:m].tolist()))

    def dump(self, cap: int = 1 << 20):
        k = np.zeros(cap, dtype=KEY_DTYPE)
        exp = np.zeros(cap, dtype=np.uint64)
        vals = np.zeros(cap * self.actors, dtype=np.uint64)
        n = C.c_uint64()
        self._check(self._lib.rl_crdt_dump(self._h, cap, k.ctypes.data, exp.ctypes.data, vals.ctypes.data, C.byref(n)))
        m = min(n.value, cap)
        v = vals[:m * self.actors].reshape(m, self.actors)
        return sorted((int(k["lo"][i]), int(k["hi"][i]), int(exp[i]), tuple(v[i].tolist())) for i in range(m))

    def clear(self):
        """CounterStorage::clear: every counter is forgotten."""
        self._check(self._lib.rl_crdt_clear(self._h))

    def kernel_launches(self) -> int:
        return int(self._lib.rl_crdt_kernel_launches(self._h))
