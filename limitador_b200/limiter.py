"""Host-side mirror of the reference's public API for the hot path.

Mirrors (names, argument meaning, error behaviour) of /root/reference/limitador/src:
  lib.rs:214-220,228-232,362-464,475-522  RateLimiter, CheckResult, is_rate_limited,
                                          update_counters, check_rate_limited_and_update,
                                          configure_with, counters_that_apply
  limit.rs:133-214                        Limit identity / applies / resolve_variables
  counter.rs:10-138                       Counter
  storage/mod.rs:26-141,279-292           Authorization, Storage facade, trait CounterStorage

The reference's toolchain (Rust) is absent from this image, so the host side above the
C-ABI is restated here; limit matching is a small table-driven subset of the CEL
expressions the reference accepts (`a == 'x'`, `a != 'x'`, `descriptors[0].a == 'x'`), and
time is injectable (`clock`) so tests are deterministic.  Counter state lives in the GPU
engine (GpuCounterStorage); nothing here computes a rate-limit decision.
"""
from __future__ import annotations

import hashlib
import re
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

from . import engine as _eng

# The table-driven subset (rl_match.cpp mirrors these rules): ASCII identifiers; a ROOT operand has no dots
# (`req.method` is CEL member access on the unbound variable `req`: the reference never applies such a limit,
# limit/cel.rs:314-322); bracket keys close with the quote they opened with; no backslash in keys or literals
# (CEL escape processing is not done here, so such expressions are refused instead of reinterpreted).
_OPERAND = (r"(?:descriptors\[(\d+)\](?:\.([A-Za-z_][A-Za-z0-9_]*)|\[(?:'([^'\\]+)'|\"([^\"\\]+)\")\])"
            r"|([A-Za-z_][A-Za-z0-9_]*))")
_PRED = re.compile(r"^\s*" + _OPERAND + r"\s*(==|!=)\s*(?:'([^'\\]*)'|\"([^\"\\]*)\")\s*$", re.ASCII)
_VAR = re.compile(r"^\s*" + _OPERAND + r"\s*$", re.ASCII)


class Context(dict):
    """limit/cel.rs:76-145 — root bindings plus an optional `descriptors` list binding."""

    def __init__(self, values=None, descriptors: Optional[List[Dict[str, str]]] = None):
        super().__init__(values or {})
        self.descriptors = descriptors or []

    def _lookup(self, m_groups) -> Optional[str]:
        idx, attr, key1, key2, ident = m_groups
        key = key1 if key1 is not None else key2
        if ident is not None:
            return self.get(ident)
        i = int(idx)
        if i >= len(self.descriptors):
            return None
        return self.descriptors[i].get(attr if attr is not None else key)


def _operand_value(ctx: Context, groups) -> Optional[str]:
    return ctx._lookup(groups)


class Limit:
    """limit.rs:31-48; identity excludes max_value, name, id (limit.rs:177-214)."""

    __slots__ = ("namespace", "max_value", "seconds", "conditions", "variables", "name", "id")

    def __init__(self, namespace, max_value, seconds, conditions=(), variables=(), name=None, id=None):
        self.namespace = namespace
        self.max_value = int(max_value)
        self.seconds = int(seconds)
        self.conditions = tuple(sorted(set(conditions)))
        self.variables = tuple(sorted(set(variables)))
        self.name = name
        self.id = id
        for c in self.conditions:
            if not _PRED.match(c):
                raise ValueError(f"unsupported condition expression: {c!r}")
        for v in self.variables:
            if not _VAR.match(v):
                raise ValueError(f"unsupported variable expression: {v!r}")

    def identity(self):
        return (self.namespace, self.seconds, self.conditions, self.variables)

    def __eq__(self, other):
        return isinstance(other, Limit) and self.identity() == other.identity()

    def __hash__(self):
        return hash(self.identity())

    def with_max_value(self, max_value, name=None):
        return Limit(self.namespace, max_value, self.seconds, self.conditions, self.variables,
                     name if name is not None else self.name, self.id)

    def applies(self, ctx: Context) -> bool:
        """limit.rs:157-174."""
        for c in self.conditions:
            m = _PRED.match(c)
            g = m.groups()
            val = _operand_value(ctx, g[0:5])
            lit = g[6] if g[6] is not None else g[7]
            ok = (val == lit) if g[5] == "==" else (val is not None and val != lit)
            if not ok:
                return False
        return self.resolve_variables(ctx) is not None

    def resolve_variables(self, ctx: Context) -> Optional[Dict[str, str]]:
        """limit.rs:133-148 — None if any variable is unset."""
        out = {}
        for v in self.variables:
            val = _operand_value(ctx, _VAR.match(v).groups())
            if val is None:
                return None
            out[v] = val
        return out


@dataclass
class Counter:
    """counter.rs:10-17; Hash/Eq over limit + set_variables only (:123-138)."""
    limit: Limit
    set_variables: Dict[str, str]
    remaining: Optional[int] = None
    expires_in_us: Optional[int] = None
    limit_id: int = -1

    @staticmethod
    def new(limit: Limit, ctx: Context) -> Optional["Counter"]:
        v = limit.resolve_variables(ctx)
        return None if v is None else Counter(limit, v)

    def is_qualified(self) -> bool:
        return bool(self.set_variables)

    def max_value(self) -> int:
        return self.limit.max_value

    def window_us(self) -> int:
        return self.limit.seconds * 1_000_000

    def expires_in_secs(self) -> Optional[int]:
        return None if self.expires_in_us is None else self.expires_in_us // 1_000_000

    def key(self) -> Tuple[int, int]:
        """96-bit digest of the resolved variable values: (key_lo:64, key_hi:32)."""
        if not self.set_variables:
            return (0, 0)
        h = hashlib.blake2b(digest_size=12)
        for k in sorted(self.set_variables):
            kb, vb = k.encode(), self.set_variables[k].encode()
            h.update(len(kb).to_bytes(4, "little") + kb + len(vb).to_bytes(4, "little") + vb)
        d = h.digest()
        return (int.from_bytes(d[:8], "little"), int.from_bytes(d[8:], "little"))

    def _ident(self):
        return (self.limit.identity(), tuple(sorted(self.set_variables.items())))

    def __eq__(self, other):
        return isinstance(other, Counter) and self._ident() == other._ident()

    def __hash__(self):
        return hash(self._ident())


@dataclass
class Authorization:
    """storage/mod.rs:26-29."""
    limited: bool
    limit_name: Optional[str] = None


@dataclass
class CheckResult:
    """lib.rs:228-232."""
    limited: bool
    counters: List[Counter] = field(default_factory=list)
    limit_name: Optional[str] = None

    def response_header(self) -> Dict[str, str]:
        """lib.rs:235-275 — draft-03 RateLimit headers from the load_counters outputs."""
        ctrs = sorted(self.counters, key=lambda c: c.remaining if c.remaining is not None else c.max_value())
        self.counters = ctrs
        headers: Dict[str, str] = {}
        text = ""
        for c in ctrs:
            text += f", {c.max_value()};w={c.limit.seconds}"
            if c.limit.name is not None:
                text += ';name="{}"'.format(c.limit.name.replace('"', "'"))
        if ctrs:
            first = ctrs[0]
            rem = first.remaining if first.remaining is not None else first.max_value()
            headers["X-RateLimit-Limit"] = f"{first.max_value()}{text}"
            headers["X-RateLimit-Remaining"] = str(rem)
            if first.expires_in_us is not None:
                headers["X-RateLimit-Reset"] = str(first.expires_in_us // 1_000_000)
        return headers


class GpuCounterStorage:
    """`impl CounterStorage` (storage/mod.rs:279-292) over the C-ABI engine.

    Single calls are batches of one; `check_and_update_many` ships many requests in one
    kernel pipeline with identical results to calling them one by one.
    """

    def __init__(self, engine: "_eng.Engine"):
        self.engine = engine

    # registry hook (Storage::add_limit / update_limit, storage/mod.rs:60-83)
    def set_limit(self, limit_id, ns_id, varset_id, qualified, max_value, window_us):
        self.engine.limits_set([(limit_id, ns_id, varset_id, int(qualified), max_value, window_us)])

    def forget_limit(self, limit_id):
        self.engine.limits_delete([limit_id])

    @staticmethod
    def _csr(counter_lists: Sequence[Sequence[Counter]]):
        off = np.zeros(len(counter_lists) + 1, dtype=np.uint32)
        flat = [c for cl in counter_lists for c in cl]
        ctrs = np.zeros(len(flat), dtype=_eng.COUNTER_DTYPE)
        for i, cl in enumerate(counter_lists):
            off[i + 1] = off[i] + len(cl)
        for j, c in enumerate(flat):
            lo, hi = c.key()
            ctrs[j] = (c.limit_id, 0, lo, hi)
        return off, ctrs, flat

    def is_within_limits(self, counter: Counter, delta: int, now_us: int) -> bool:
        off, ctrs, _ = self._csr([[counter]])
        lim, _ = self.engine.is_within_limits_batch(off, ctrs, [delta], [now_us])
        return not bool(lim[0])

    def first_limited(self, counters: Sequence[Counter], delta: int, now_us: int) -> Optional[Counter]:
        """find_first_limited_counter (lib.rs:387-409) in one call."""
        if not counters:
            return None
        off, ctrs, flat = self._csr([counters])
        lim, fl = self.engine.is_within_limits_batch(off, ctrs, [delta], [now_us])
        if not lim[0]:
            return None
        return next(c for c in flat if c.limit_id == int(fl[0]))

    def update_counter(self, counter: Counter, delta: int, now_us: int) -> None:
        off, ctrs, _ = self._csr([[counter]])
        self.engine.update_batch(off, ctrs, [delta], [now_us])

    def check_and_update(self, counters: List[Counter], delta: int, load_counters: bool, now_us: int) -> Authorization:
        return self.check_and_update_many([counters], [delta], [now_us], load_counters)[0]

    def check_and_update_many(self, counter_lists, deltas, nows, load_counters) -> List[Authorization]:
        off, ctrs, flat = self._csr(counter_lists)
        lim, fl, rem, ttl = self.engine.check_and_update_batch(off, ctrs, deltas, nows, load_counters)
        out = []
        for i, cl in enumerate(counter_lists):
            if load_counters:
                for j, c in enumerate(cl):
                    c.remaining = int(rem[off[i] + j])
                    c.expires_in_us = int(ttl[off[i] + j])
            name = None
            if lim[i]:
                named = next(c for c in cl if c.limit_id == int(fl[i]))
                name = named.limit.name
            out.append(Authorization(bool(lim[i]), name))
        return out

    def get_counters(self, limit_ids: Sequence[int], now_us: int):
        return self.engine.get_counters(limit_ids, now_us)

    def delete_counters(self, limit_ids: Sequence[int]) -> None:
        self.engine.delete_counters(limit_ids)

    def clear(self) -> None:
        self.engine.clear()


class RateLimiter:
    """lib.rs:214-216,306-522 over a CounterStorage implementor."""

    def __init__(self, storage, clock: Optional[Callable[[], int]] = None):
        self.storage = storage
        self.clock = clock or (lambda: time.time_ns() // 1000)
        self._limits: Dict[str, Dict[Limit, Limit]] = {}  # namespace -> {identity: current Limit}
        self._limit_ids: Dict[tuple, int] = {}             # identity -> dense id (never reused)
        self._ns_ids: Dict[str, int] = {}
        self._varsets: Dict[Tuple[str, Tuple[str, ...]], int] = {}
        self._ctr_vars: Dict[Tuple[int, int, int], Dict[str, str]] = {}

    # -- interning --
    def _intern(self, limit: Limit) -> int:
        ident = limit.identity()
        if ident not in self._limit_ids:
            self._limit_ids[ident] = len(self._limit_ids)
        return self._limit_ids[ident]

    def _push_limit(self, limit: Limit) -> int:
        lid = self._intern(limit)
        ns_id = self._ns_ids.setdefault(limit.namespace, len(self._ns_ids))
        varset = 0
        if limit.variables:
            varset = self._varsets.setdefault((limit.namespace, limit.variables), len(self._varsets) + 1)
        self.storage.set_limit(lid, ns_id, varset, bool(limit.variables), limit.max_value,
                               limit.seconds * 1_000_000)
        return lid

    # -- limits CRUD (storage/mod.rs:56-124) --
    def get_namespaces(self) -> Set[str]:
        return set(self._limits.keys())

    def add_limit(self, limit: Limit) -> bool:
        ns = self._limits.setdefault(limit.namespace, {})
        if limit in ns:
            self._push_limit(ns[limit])  # add_counter: entry().or_default(); the set keeps the old Limit
            return False
        ns[limit] = limit
        self._push_limit(limit)
        return True

    def update_limit(self, update: Limit) -> bool:
        ns = self._limits.get(update.namespace)
        if ns is None or update not in ns:
            return False
        cur = ns[update]
        if cur.max_value != update.max_value or cur.name != update.name:
            ns[update] = update
            self._push_limit(update)
            return True
        return False

    def delete_limit(self, limit: Limit) -> None:
        lid = self._limit_ids.get(limit.identity())
        ns = self._limits.get(limit.namespace)
        if lid is not None and ns is not None and limit in ns:
            self.storage.delete_counters([lid])
            self.storage.forget_limit(lid)
            del ns[limit]
            if not ns:
                del self._limits[limit.namespace]

    def get_limits(self, namespace: str) -> Set[Limit]:
        return set(self._limits.get(namespace, {}).values())

    def delete_limits(self, namespace: str) -> None:
        for limit in list(self._limits.get(namespace, {}).values()):
            self.delete_limit(limit)

    def configure_with(self, limits: Iterable[Limit]) -> None:
        """lib.rs:475-505."""
        keep: Dict[str, Dict[Limit, Limit]] = {}
        for l in limits:
            keep.setdefault(l.namespace, {})[l] = l
        for namespace in set(self.get_namespaces()) | set(keep.keys()):
            current = dict(self._limits.get(namespace, {}))
            wanted = keep.get(namespace, {})
            for l in current:
                if l not in wanted:
                    self.delete_limit(current[l])
            for l in wanted:
                if l not in current:
                    self.add_limit(wanted[l])
            for l in wanted:
                self.update_limit(wanted[l])

    # -- hot path --
    def counters_that_apply(self, namespace: str, ctx: Context) -> List[Counter]:
        """lib.rs:507-522 (iteration order = insertion order here; HashSet order there)."""
        out = []
        for limit in self._limits.get(namespace, {}).values():
            if not limit.applies(ctx):
                continue
            c = Counter.new(limit, ctx)
            if c is None:
                continue
            c.limit_id = self._limit_ids[limit.identity()]
            lo, hi = c.key()
            self._ctr_vars[(c.limit_id, lo, hi)] = c.set_variables
            out.append(c)
        return out

    def is_rate_limited(self, namespace: str, ctx: Context, delta: int) -> CheckResult:
        counters = self.counters_that_apply(namespace, ctx)
        limited = self.storage.first_limited(counters, delta, self.clock())
        if limited is None:
            return CheckResult(False)
        return CheckResult(True, [], limited.limit.name)

    def update_counters(self, namespace: str, ctx: Context, delta: int) -> None:
        now = self.clock()
        for c in self.counters_that_apply(namespace, ctx):
            self.storage.update_counter(c, delta, now)

    def check_rate_limited_and_update(self, namespace: str, ctx: Context, delta: int,
                                      load_counters: bool) -> CheckResult:
        counters = self.counters_that_apply(namespace, ctx)
        if not counters:
            return CheckResult(False)
        auth = self.storage.check_and_update(counters, delta, load_counters, self.clock())
        return CheckResult(auth.limited, counters if load_counters else [], auth.limit_name)

    def check_rate_limited_and_update_batch(self, requests: Sequence[Tuple[str, Context, int]],
                                            load_counters: bool, nows: Optional[Sequence[int]] = None
                                            ) -> List[CheckResult]:
        """The batching front: many requests, one kernel pipeline, sequential semantics."""
        lists = [self.counters_that_apply(ns, ctx) for ns, ctx, _ in requests]
        deltas = [d for _, _, d in requests]
        if nows is None:
            t = self.clock()
            nows = [t] * len(requests)
        auths = self.storage.check_and_update_many(lists, deltas, list(nows), load_counters)
        return [CheckResult(a.limited, cl if load_counters else [], a.limit_name) for a, cl in zip(auths, lists)]

    def get_counters(self, namespace: str) -> Set[Counter]:
        """lib.rs:466-470 → storage get_counters (in_memory.rs:158-187)."""
        limits = self._limits.get(namespace, {})
        if not limits:
            return set()
        by_id = {self._limit_ids[l.identity()]: l for l in limits.values()}
        out = set()
        for lid, lo, hi, rem, ttl in self.storage.get_counters(list(by_id.keys()), self.clock()):
            if lid not in by_id:
                continue
            vars_ = self._ctr_vars.get((lid, lo, hi), {})
            out.add(Counter(by_id[lid], dict(vars_), remaining=rem, expires_in_us=ttl, limit_id=lid))
        return out
