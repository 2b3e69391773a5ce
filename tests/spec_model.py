"""A second, independent restatement of the reference's in-memory store: plain Python dicts, written from
the reference sources (not from oracle/limitador_oracle.c) to pin the C oracle on random streams — the C
oracle is the checker for everything else, so it gets a checker of its own.  Test infrastructure only.

  value_at / ttl / update      limitador/src/storage/atomic_expiring_value.rs:19-47,62-99
  check_and_update             limitador/src/storage/in_memory.rs:72-156
  is_within_limits             in_memory.rs:20-35      (folded like RateLimiter::is_rate_limited, lib.rs:362-409)
  update_counter               in_memory.rs:47-69
  add_counter                  in_memory.rs:38-44
  get_counters                 in_memory.rs:158-187
  delete_counters / clear      in_memory.rs:189-201,241-257
"""
M64 = (1 << 64) - 1
NONE = 0xFFFFFFFF


class SpecStore:
    def __init__(self):
        self.limits = {}  # limit_id -> [ns_id, max_value, window_us, qualified]
        self.simple = {}  # limit_id -> [value, expiry_us]          (limits without variables)
        self.qualified = {}  # (limit_id, key_lo, key_hi) -> [value, expiry_us]

    # -- limits ---------------------------------------------------------------------------
    def limit_set(self, limit_id, ns_id, max_value, window_us, qualified):
        if limit_id in self.limits:
            self.limits[limit_id][1] = max_value  # update_limit: only max_value moves (storage/mod.rs:67-83)
        else:
            self.limits[limit_id] = [ns_id, max_value, window_us, bool(qualified)]
        if not self.limits[limit_id][3]:
            self.simple.setdefault(limit_id, [0, 0])  # entry().or_default() = (0, UNIX_EPOCH)

    def limit_delete(self, limit_id):
        self.delete_counters([limit_id])
        del self.limits[limit_id]

    # -- AtomicExpiringValue ------------------------------------------------------------------
    @staticmethod
    def value_at(e, now):
        return 0 if e[1] <= now else e[0]  # expired_at: expiry <= when (inclusive)

    @staticmethod
    def ttl(e, now):
        return e[1] - now if e[1] > now else 0

    @staticmethod
    def update(e, delta, window, now):
        if e[1] <= now:  # update_if_expired: re-anchor the window, value = delta
            e[1] = now + window
            e[0] = delta
        else:
            e[0] = (e[0] + delta) & M64

    # -- CounterStorage -----------------------------------------------------------------------
    def check_and_update(self, ctrs, delta, load_counters, now):
        """ctrs: [(limit_id, key_lo, key_hi)].  -> (limited, first limited limit_id or NONE, remaining[], ttl[])"""
        remaining, ttls = [0] * len(ctrs), [0] * len(ctrs)
        first, touched = NONE, []
        for want_qualified in (False, True):  # simple counters first (:105), then qualified (:121)
            for i, (lid, lo, hi) in enumerate(ctrs):
                _, mx, window, q = self.limits[lid]
                if q != want_qualified:
                    continue
                if q:
                    e = self.qualified.setdefault((lid, lo, hi), [0, now + window])  # insert on lookup (:122-127)
                else:
                    e = self.simple[lid]  # must exist: the reference unwrap()s (:107)
                total = (self.value_at(e, now) + delta) & M64
                over = total > mx
                if load_counters:
                    remaining[i] = 0 if over else mx - total  # checked_sub(..).unwrap_or_default (:88-89)
                    if over and first == NONE:
                        first = lid
                    ttls[i] = self.ttl(e, now)  # before any update (:114-116,:134-136)
                elif over:
                    return True, lid, remaining, ttls  # early return, nothing incremented (:110-112,:130-132)
                touched.append((e, window))
        if first != NONE:
            return True, first, remaining, ttls
        for e, window in touched:
            self.update(e, delta, window, now)
        return False, NONE, remaining, ttls

    def is_rate_limited(self, ctrs, delta, now):
        for lid, lo, hi in ctrs:  # given order, no insert
            _, mx, _, q = self.limits[lid]
            e = self.qualified.get((lid, lo, hi)) if q else self.simple.get(lid)
            v = self.value_at(e, now) if e is not None else 0
            if not mx >= ((v + delta) & M64):
                return True, lid
        return False, NONE

    def update_counters(self, ctrs, delta, now):
        for lid, lo, hi in ctrs:
            _, _, window, q = self.limits[lid]
            if q:
                e = self.qualified.setdefault((lid, lo, hi), [0, now + window])
                self.update(e, delta, window, now)
            elif lid not in self.simple:
                self.simple[lid] = [delta, now + window]  # Entry::Vacant
            else:
                self.update(self.simple[lid], delta, window, now)

    def get_counters(self, limit_ids, now):
        """counters_in_namespace(limit.namespace()) for every given limit (:161-172, :214-238: every simple AND
        every qualified counter of that namespace), plus the qualified counters of the given limits (:174-184,
        a subset of the former); kept iff ttl > 0.  -> sorted [(limit_id, lo, hi, remaining, ttl)]"""
        nss = {self.limits[l][0] for l in limit_ids if l in self.limits}
        out = []
        for lid, e in self.simple.items():
            if lid in self.limits and self.limits[lid][0] in nss and self.ttl(e, now) > 0:
                out.append((lid, 0, 0, (self.limits[lid][1] - self.value_at(e, now)) & M64, self.ttl(e, now)))
        for (lid, lo, hi), e in self.qualified.items():
            if self.limits[lid][0] in nss and self.ttl(e, now) > 0:
                out.append((lid, lo, hi, (self.limits[lid][1] - self.value_at(e, now)) & M64, self.ttl(e, now)))
        return sorted(out)

    def delete_counters(self, limit_ids):
        for lid in limit_ids:
            self.simple.pop(lid, None)
            for k in [k for k in self.qualified if k[0] == lid]:
                del self.qualified[k]

    def clear(self):
        self.simple.clear()  # the qualified cache is left alone (:197-201)

    def dump(self):
        out = [(lid, 0, 0, e[0], e[1]) for lid, e in self.simple.items()]
        out += [(lid, lo, hi, e[0], e[1]) for (lid, lo, hi), e in self.qualified.items()]
        return sorted(out)
