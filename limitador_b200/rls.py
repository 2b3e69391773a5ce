"""ctypes binding of the Envoy RLS v3 wire surface (include/rl_rls.h, csrc/rl_rls.cpp).

`RlsService` serves batches of `envoy.service.ratelimit.v3.RateLimitRequest` wire messages:
decode + counters_that_apply on a pool of CPU workers, ONE engine call for the whole batch, then
`RateLimitResponse` bytes (envoy_rls/server.rs:91-208, kuadrant_service.rs:27-186).  `plan`/`finish` are
the CPU stages on their own (drivable without a GPU); `serve` runs all three through the engine.
`encode_request` / `decode_response` are small pure-Python helpers for callers and tests (the tests
cross-check them against the protobuf runtime).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine as _eng
from . import matcher as _m

CODE_UNKNOWN, CODE_OK, CODE_OVER_LIMIT = 0, 1, 2
HEADERS_NONE, HEADERS_DRAFT_VERSION_03 = 0, 1
SHOULD_RATE_LIMIT, CHECK_RATE_LIMIT, REPORT = 0, 1, 2
GRPC_OK, GRPC_INTERNAL, GRPC_UNAVAILABLE = 0, 13, 14
NO_STORE = 0xFFFFFFFF

RLS_SYMBOLS = (
    "rl_rls_decode_request", "rl_rls_encode_response", "rl_rls_create", "rl_rls_destroy", "rl_rls_last_error",
    "rl_rls_plan", "rl_rls_plan_view", "rl_rls_finish", "rl_rls_responses", "rl_rls_serve", "rl_rls_metrics_render",
    "rl_rls_last_timings",
)

ENTRY_DTYPE = np.dtype([("descriptor", "<u4"), ("key_off", "<u4"), ("key_len", "<u4"), ("val_off", "<u4"), ("val_len", "<u4")])


class RlsRequest(C.Structure):
    _fields_ = [("domain_off", C.c_uint32), ("domain_len", C.c_uint32), ("hits_addend", C.c_uint32),
                ("n_descriptors", C.c_uint32), ("n_entries", C.c_uint32)]


class RlsError(RuntimeError):
    pass


def _lib():
    L = _m._lib()
    if getattr(L, "_rl_rls_ready", False):
        return L
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.rl_rls_decode_request.argtypes = [vp, u64, C.POINTER(RlsRequest), vp, u32]
    L.rl_rls_encode_response.argtypes = [u32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), u32, vp, u64, C.POINTER(u64)]
    L.rl_rls_create.argtypes = [vp, vp, i32, u32, i32, C.POINTER(vp)]
    L.rl_rls_destroy.argtypes = [vp]
    L.rl_rls_destroy.restype = None
    L.rl_rls_last_error.argtypes = [vp]
    L.rl_rls_last_error.restype = C.c_char_p
    L.rl_rls_plan.argtypes = [vp, i32, u64, vp, vp, u64]
    L.rl_rls_plan_view.argtypes = [vp, C.POINTER(u64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                   C.POINTER(i32), C.POINTER(vp)]
    L.rl_rls_finish.argtypes = [vp, i32, vp, vp, vp, vp]
    L.rl_rls_responses.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.rl_rls_serve.argtypes = [vp, i32, u64, vp, vp, u64]
    L.rl_rls_metrics_render.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.rl_rls_last_timings.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L._rl_rls_ready = True
    return L


# ---- pure-Python wire helpers ------------------------------------------------------------------------------------
def _varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _len_field(tag: int, payload: bytes) -> bytes:
    return _varint((tag << 3) | 2) + _varint(len(payload)) + payload


def encode_request(domain: str, descriptors: Sequence[Sequence[Tuple[str, str]]], hits_addend: int = 0) -> bytes:
    """RateLimitRequest{domain = 1, descriptors = 2 (entries = 1 {key = 1, value = 2}), hits_addend = 3}; proto3:
    empty strings and a zero hits_addend are not written."""
    out = bytearray()
    if domain:
        out += _len_field(1, domain.encode())
    for d in descriptors:
        body = bytearray()
        for k, v in d:
            e = (_len_field(1, k.encode()) if k else b"") + (_len_field(2, v.encode()) if v else b"")
            body += _len_field(1, e)
        out += _len_field(2, bytes(body))
    if hits_addend:
        out += _varint(3 << 3) + _varint(hits_addend)
    return bytes(out)


def pack_requests(msgs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    """-> (buf uint8, off uint64[n+1]): the batch layout rl_rls_plan / rl_rls_serve take."""
    off = np.zeros(len(msgs) + 1, dtype=np.uint64)
    if msgs:
        off[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(msgs), dtype=np.uint8) if msgs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(buf), off


def _read_varint(b: bytes, p: int) -> Tuple[int, int]:
    v = s = 0
    while True:
        c = b[p]
        p += 1
        v |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return v, p


def decode_response(b: bytes) -> Tuple[int, List[Tuple[str, str]]]:
    """RateLimitResponse bytes -> (overall_code, response_headers_to_add as (key, value) pairs)."""
    code, headers, p = 0, [], 0
    while p < len(b):
        key, p = _read_varint(b, p)
        tag, wt = key >> 3, key & 7
        if wt == 0:
            v, p = _read_varint(b, p)
            if tag == 1:
                code = v
        elif wt == 2:
            n, p = _read_varint(b, p)
            body, p = b[p:p + n], p + n
            if tag == 3:
                k = v = ""
                q = 0
                while q < len(body):
                    kk, q = _read_varint(body, q)
                    ln, q = _read_varint(body, q)
                    s, q = body[q:q + ln].decode(), q + ln
                    if kk >> 3 == 1:
                        k = s
                    elif kk >> 3 == 2:
                        v = s
                headers.append((k, v))
        else:
            raise ValueError(f"unexpected wire type {wt}")
    return code, headers


def decode_request(msg: bytes, cap_entries: int = 64):
    """The native decoder on one message -> (domain, descriptors as lists of (key, value), hits_addend as on the wire);
    raises RlsError for a message prost would refuse."""
    L = _lib()
    arr = np.frombuffer(msg, dtype=np.uint8) if msg else np.zeros(1, dtype=np.uint8)
    q = RlsRequest()
    ent = np.zeros(max(cap_entries, 1), dtype=ENTRY_DTYPE)
    if L.rl_rls_decode_request(arr.ctypes.data, len(msg), C.byref(q), ent.ctypes.data, cap_entries) != 0:
        raise RlsError("malformed RateLimitRequest")
    if q.n_entries > cap_entries:
        return decode_request(msg, q.n_entries)
    descs: List[List[Tuple[str, str]]] = [[] for _ in range(q.n_descriptors)]
    for e in ent[:q.n_entries]:
        descs[int(e["descriptor"])].append((msg[int(e["key_off"]):int(e["key_off"]) + int(e["key_len"])].decode(),
                                            msg[int(e["val_off"]):int(e["val_off"]) + int(e["val_len"])].decode()))
    return msg[q.domain_off:q.domain_off + q.domain_len].decode(), descs, q.hits_addend


def encode_response(code: int, headers: Sequence[Tuple[str, str]] = ()) -> bytes:
    L = _lib()
    ks = _m._strs([k for k, _ in headers])
    vs = _m._strs([v for _, v in headers])
    need = C.c_uint64()
    buf = np.zeros(64 + sum(len(k) + len(v) + 16 for k, v in headers), dtype=np.uint8)
    if L.rl_rls_encode_response(code, ks, vs, len(headers), buf.ctypes.data, len(buf), C.byref(need)) != 0:
        raise RlsError("rl_rls_encode_response failed")
    return buf[:need.value].tobytes()


def _view(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype)


class RlsService:
    """One RLS front over a Matcher and (optionally) an Engine.  Not thread-safe: one batch at a time."""

    def __init__(self, matcher: _m.Matcher, engine: Optional[_eng.Engine] = None, headers: int = HEADERS_NONE,
                 threads: int = 0, use_limit_name_label: bool = False):
        self._lib = _lib()
        self._matcher, self._engine = matcher, engine  # keep them alive
        h = C.c_void_p()
        eh = engine._h if engine is not None else None
        if self._lib.rl_rls_create(matcher._h, eh, headers, threads, int(use_limit_name_label), C.byref(h)) != 0:
            raise RlsError("rl_rls_create failed")
        self._h = h

    def close(self):
        if self._h:
            self._lib.rl_rls_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status):
        if status != 0:
            raise RlsError(self._lib.rl_rls_last_error(self._h).decode())

    def plan(self, method: int, buf: np.ndarray, off: np.ndarray, now_us: int = 0):
        """Stage 1 -> dict(n_store, ctr_off, ctrs, delta, now_us, load_counters, store_index): copies of the store call."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = len(off) - 1
        self._keep = (buf, off)
        self._check(self._lib.rl_rls_plan(self._h, method, n, buf.ctypes.data if len(buf) else None, off.ctypes.data, now_us))
        ns = C.c_uint64()
        p_off, p_ctr, p_delta, p_now, p_idx = (C.c_void_p() for _ in range(5))
        lc = C.c_int()
        self._check(self._lib.rl_rls_plan_view(self._h, C.byref(ns), C.byref(p_off), C.byref(p_ctr), C.byref(p_delta),
                                               C.byref(p_now), C.byref(lc), C.byref(p_idx)))
        m = ns.value
        ctr_off = _view(p_off.value, m + 1, np.uint32).copy()
        return {
            "n_store": m, "ctr_off": ctr_off,
            "ctrs": _view(p_ctr.value, int(ctr_off[-1]) if m else 0, _eng.COUNTER_DTYPE).copy(),
            "delta": _view(p_delta.value, m, np.uint64).copy(), "now_us": _view(p_now.value, m, np.uint64).copy(),
            "load_counters": bool(lc.value), "store_index": _view(p_idx.value, n, np.uint32).copy(),
        }

    def finish(self, limited=None, first_limited=None, remaining=None, ttl_us=None, store_status: int = 0):
        arrs = []

        def ptr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            arrs.append(a)
            return a.ctypes.data if len(a) else None

        self._check(self._lib.rl_rls_finish(self._h, store_status, ptr(limited, np.uint8), ptr(first_limited, np.uint32),
                                            ptr(remaining, np.uint64), ptr(ttl_us, np.uint64)))
        return self.responses()

    def responses(self) -> List[Tuple[int, bytes]]:
        """-> [(grpc_status, response bytes)] of the last finished batch."""
        pb, po, pg, pc = (C.c_void_p() for _ in range(4))
        self._check(self._lib.rl_rls_responses(self._h, C.byref(pb), C.byref(po), C.byref(pg), C.byref(pc)))
        n = len(self._keep[1]) - 1
        off = _view(po.value, n + 1, np.uint64)
        raw = _view(pb.value, int(off[-1]) if n else 0, np.uint8).tobytes()
        grpc = _view(pg.value, n, np.uint8)
        return [(int(grpc[i]), raw[int(off[i]):int(off[i + 1])]) for i in range(n)]

    def codes(self) -> np.ndarray:
        """overall_code of every response of the last finished batch (0 where the gRPC status is not OK)."""
        pc = C.c_void_p()
        self._check(self._lib.rl_rls_responses(self._h, None, None, None, C.byref(pc)))
        return _view(pc.value, len(self._keep[1]) - 1, np.uint8).copy()

    def grpc_status(self) -> np.ndarray:
        pg = C.c_void_p()
        self._check(self._lib.rl_rls_responses(self._h, None, None, C.byref(pg), None))
        return _view(pg.value, len(self._keep[1]) - 1, np.uint8).copy()

    def serve(self, method: int, buf: np.ndarray, off: np.ndarray, now_us: int = 0):
        """plan -> ONE engine call -> finish.  Needs an engine (no CPU store exists in the product)."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        self._keep = (buf, off)
        self._check(self._lib.rl_rls_serve(self._h, method, len(off) - 1, buf.ctypes.data if len(buf) else None,
                                           off.ctypes.data, now_us))

    def timings(self) -> Dict[str, float]:
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._lib.rl_rls_last_timings(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"plan_us": a.value, "store_us": b.value, "finish_us": c.value}

    def metrics(self) -> str:
        need = C.c_uint64()
        self._lib.rl_rls_metrics_render(self._h, None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value)
        self._check(self._lib.rl_rls_metrics_render(self._h, buf, need.value, C.byref(need)))
        return buf.value.decode()
