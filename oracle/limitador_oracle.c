#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
/*
 * limitador_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT).
 * See limitador_oracle.h for the scope and the parity pin.  Every function cites the
 * reference lines (relative to /root/reference/) it restates.  Time is an explicit
 * `now_us` (µs since the UNIX epoch, atomic_expiring_value.rs:62-66); all arithmetic is
 * u64 and wraps like a Rust release build.
 */
#include "limitador_oracle.h"

#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- AtomicExpiringValue (atomic_expiring_value.rs:6-47) --------------------------- */
typedef struct {
    uint64_t value;
    uint64_t expiry; /* µs since epoch */
} lo_entry;

/* atomic_expiring_value.rs:76-79 — expired_at: expiry <= when (inclusive bound). */
static inline int entry_expired_at(const lo_entry *e, uint64_t when) { return e->expiry <= when; }
/* atomic_expiring_value.rs:19-24 — value_at. */
static inline uint64_t entry_value_at(const lo_entry *e, uint64_t when) {
    return entry_expired_at(e, when) ? 0 : e->value;
}
/* atomic_expiring_value.rs:68-74 — ttl = max(0, expiry - now). */
static inline uint64_t entry_ttl(const lo_entry *e, uint64_t now) {
    return e->expiry > now ? e->expiry - now : 0;
}
/* atomic_expiring_value.rs:36-42 + update_if_expired :87-99 (single-threaded: the CAS
 * always wins). */
static inline uint64_t entry_update(lo_entry *e, uint64_t delta, uint64_t ttl_us, uint64_t when) {
    if (e->expiry <= when) {
        e->expiry = when + ttl_us;
        e->value = delta;
        return delta;
    }
    e->value += delta;
    return e->value;
}

/* ---- limits registry (storage/mod.rs:31-34) ---------------------------------------- */
typedef struct {
    uint8_t defined;
    uint8_t qualified; /* !variables.is_empty() — counter.rs:108-110 */
    uint32_t ns_id;
    uint64_t max_value;
    uint64_t window_us; /* seconds * 1e6 — counter.rs:76-78 */
    /* simple_limits entry (in_memory.rs:14): only for unqualified limits */
    uint8_t simple_present;
    lo_entry simple;
} lo_limit;

/* ---- qualified_counters (in_memory.rs:15; moka used as a plain concurrent map) ----- */
typedef struct {
    uint32_t limit_id;
    uint32_t state; /* 0 empty, 1 full */
    uint64_t key_lo, key_hi;
    lo_entry e;
} lo_slot;

struct lo_oracle {
    lo_limit *limits;
    uint32_t limits_cap;
    /* namespace -> ordered limit ids (registration order), for the record format */
    uint32_t **ns_limits;
    uint32_t *ns_count;
    uint32_t ns_cap;
    lo_slot *slots;
    uint64_t nslots; /* power of two */
    uint64_t nfull;
};

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
static inline uint64_t slot_hash(uint32_t limit_id, uint64_t lo, uint64_t hi) {
    return mix64(lo ^ mix64(hi + 0x9e3779b97f4a7c15ULL * (limit_id + 1ULL)));
}

static uint64_t pow2_at_least(uint64_t x) {
    uint64_t p = 16;
    while (p < x) p <<= 1;
    return p;
}

lo_oracle *lo_create(uint64_t capacity_hint) {
    lo_oracle *o = (lo_oracle *)calloc(1, sizeof(*o));
    if (!o) return NULL;
    o->nslots = pow2_at_least(capacity_hint * 2);
    o->slots = (lo_slot *)calloc(o->nslots, sizeof(lo_slot));
    if (!o->slots) {
        free(o);
        return NULL;
    }
    return o;
}

void lo_destroy(lo_oracle *o) {
    if (!o) return;
    for (uint32_t i = 0; i < o->ns_cap; i++) free(o->ns_limits[i]);
    free(o->ns_limits);
    free(o->ns_count);
    free(o->limits);
    free(o->slots);
    free(o);
}

static void grow_table(lo_oracle *o) {
    uint64_t old_n = o->nslots;
    lo_slot *old = o->slots;
    o->nslots = old_n * 2;
    o->slots = (lo_slot *)calloc(o->nslots, sizeof(lo_slot));
    for (uint64_t i = 0; i < old_n; i++) {
        if (!old[i].state) continue;
        uint64_t h = slot_hash(old[i].limit_id, old[i].key_lo, old[i].key_hi) & (o->nslots - 1);
        while (o->slots[h].state) h = (h + 1) & (o->nslots - 1);
        o->slots[h] = old[i];
    }
    free(old);
}

static lo_slot *q_find(lo_oracle *o, uint32_t limit_id, uint64_t lo, uint64_t hi) {
    uint64_t mask = o->nslots - 1;
    uint64_t h = slot_hash(limit_id, lo, hi) & mask;
    for (;;) {
        lo_slot *s = &o->slots[h];
        if (!s->state) return NULL;
        if (s->limit_id == limit_id && s->key_lo == lo && s->key_hi == hi) return s;
        h = (h + 1) & mask;
    }
}

/* Room for `extra` more entries without growing: callers that keep entry pointers across several
 * q_get_or_insert calls (one request's counters) reserve first, so no pointer outlives a rehash. */
static void q_reserve(lo_oracle *o, uint64_t extra) {
    while ((o->nfull + extra) * 2 > o->nslots) grow_table(o);
}

/* moka get_with / get_with_by_ref (in_memory.rs:51-56,122-127): insert-if-missing. */
static lo_slot *q_get_or_insert(lo_oracle *o, uint32_t limit_id, uint64_t lo, uint64_t hi,
                                lo_entry init) {
    if ((o->nfull + 1) * 2 > o->nslots) grow_table(o);
    uint64_t mask = o->nslots - 1;
    uint64_t h = slot_hash(limit_id, lo, hi) & mask;
    for (;;) {
        lo_slot *s = &o->slots[h];
        if (!s->state) {
            s->state = 1;
            s->limit_id = limit_id;
            s->key_lo = lo;
            s->key_hi = hi;
            s->e = init;
            o->nfull++;
            return s;
        }
        if (s->limit_id == limit_id && s->key_lo == lo && s->key_hi == hi) return s;
        h = (h + 1) & mask;
    }
}

/* Rebuild keeping only the entries for which keep(slot) != 0 (models
 * invalidate_entries_if, in_memory.rs:246-253). */
typedef int (*keep_fn)(const lo_slot *, void *);
static uint64_t q_filter(lo_oracle *o, keep_fn keep, void *arg) {
    lo_slot *old = o->slots;
    uint64_t old_n = o->nslots, dropped = 0;
    o->slots = (lo_slot *)calloc(o->nslots, sizeof(lo_slot));
    o->nfull = 0;
    uint64_t mask = o->nslots - 1;
    for (uint64_t i = 0; i < old_n; i++) {
        if (!old[i].state) continue;
        if (!keep(&old[i], arg)) {
            dropped++;
            continue;
        }
        uint64_t h = slot_hash(old[i].limit_id, old[i].key_lo, old[i].key_hi) & mask;
        while (o->slots[h].state) h = (h + 1) & mask;
        o->slots[h] = old[i];
        o->nfull++;
    }
    free(old);
    return dropped;
}

static lo_limit *get_limit(lo_oracle *o, uint32_t id) {
    if (id >= o->limits_cap || !o->limits[id].defined) return NULL;
    return &o->limits[id];
}

int lo_limit_set(lo_oracle *o, uint32_t limit_id, uint32_t ns_id, uint64_t max_value,
                 uint64_t window_us, int qualified) {
    if (limit_id == LO_NONE) return -1;
    if (limit_id >= o->limits_cap) {
        uint32_t ncap = o->limits_cap ? o->limits_cap : 16;
        while (ncap <= limit_id) ncap *= 2;
        o->limits = (lo_limit *)realloc(o->limits, ncap * sizeof(lo_limit));
        memset(o->limits + o->limits_cap, 0, (ncap - o->limits_cap) * sizeof(lo_limit));
        o->limits_cap = ncap;
    }
    lo_limit *l = &o->limits[limit_id];
    if (l->defined) {
        /* update_limit (storage/mod.rs:67-83): identity fixed, max_value swapped. */
        if (l->ns_id != ns_id || l->window_us != window_us || l->qualified != (qualified != 0))
            return -1;
        l->max_value = max_value;
        if (!l->qualified && !l->simple_present) { /* add_counter again: entry().or_default() */
            l->simple_present = 1;
            l->simple.value = 0;
            l->simple.expiry = 0;
        }
        return 0;
    }
    l->defined = 1;
    l->qualified = qualified != 0;
    l->ns_id = ns_id;
    l->max_value = max_value;
    l->window_us = window_us;
    /* add_counter (in_memory.rs:38-44): entry(limit).or_default() = (0, UNIX_EPOCH)
     * (atomic_expiring_value.rs:151-158), unqualified limits only. */
    if (!l->qualified && !l->simple_present) {
        l->simple_present = 1;
        l->simple.value = 0;
        l->simple.expiry = 0;
    }
    if (ns_id >= o->ns_cap) {
        uint32_t ncap = o->ns_cap ? o->ns_cap : 16;
        while (ncap <= ns_id) ncap *= 2;
        o->ns_limits = (uint32_t **)realloc(o->ns_limits, ncap * sizeof(uint32_t *));
        o->ns_count = (uint32_t *)realloc(o->ns_count, ncap * sizeof(uint32_t));
        for (uint32_t i = o->ns_cap; i < ncap; i++) {
            o->ns_limits[i] = NULL;
            o->ns_count[i] = 0;
        }
        o->ns_cap = ncap;
    }
    uint32_t c = o->ns_count[ns_id];
    o->ns_limits[ns_id] = (uint32_t *)realloc(o->ns_limits[ns_id], (c + 1) * sizeof(uint32_t));
    o->ns_limits[ns_id][c] = limit_id;
    o->ns_count[ns_id] = c + 1;
    return 0;
}

static int keep_not_limit(const lo_slot *s, void *arg) { return s->limit_id != *(uint32_t *)arg; }

/* delete_counters_of_limit (in_memory.rs:241-257). */
static void delete_counters_of_limit(lo_oracle *o, uint32_t limit_id) {
    lo_limit *l = get_limit(o, limit_id);
    if (!l) return;
    if (!l->qualified) {
        l->simple_present = 0;
        l->simple.value = 0;
        l->simple.expiry = 0;
    } else {
        q_filter(o, keep_not_limit, &limit_id);
    }
}

int lo_limit_delete(lo_oracle *o, uint32_t limit_id) {
    lo_limit *l = get_limit(o, limit_id);
    if (!l) return 0;
    delete_counters_of_limit(o, limit_id);
    uint32_t ns = l->ns_id, c = o->ns_count[ns], w = 0;
    for (uint32_t i = 0; i < c; i++)
        if (o->ns_limits[ns][i] != limit_id) o->ns_limits[ns][w++] = o->ns_limits[ns][i];
    o->ns_count[ns] = w;
    memset(l, 0, sizeof(*l));
    return 0;
}

int lo_check_and_update(lo_oracle *o, const lo_counter *ctrs, uint32_t m, uint64_t delta,
                        int load_counters, uint64_t now_us, uint32_t *first_limited_out,
                        uint64_t *remaining, uint64_t *ttl_us) {
    enum { MAXC = 64 };
    lo_entry *touched[MAXC];
    uint64_t touched_w[MAXC];
    uint32_t nt = 0;
    uint32_t first_limited = LO_NONE;
    if (first_limited_out) *first_limited_out = LO_NONE;
    if (m > MAXC) return -2;
    for (uint32_t i = 0; i < m; i++)
        if (!get_limit(o, ctrs[i].limit_id)) return -1;
    q_reserve(o, (uint64_t)m + 1); /* touched[] points into the table until the updates below */

    /* in_memory.rs:105 (simple counters first) then :121 (qualified counters). */
    for (int pass = 0; pass < 2; pass++) {
        for (uint32_t i = 0; i < m; i++) {
            lo_limit *l = &o->limits[ctrs[i].limit_id];
            if ((int)l->qualified != pass) continue;
            lo_entry *e;
            if (!l->qualified) {
                /* :106-107 — limits_by_namespace.get(limit).unwrap(): panics if absent */
                if (!l->simple_present) return -3;
                e = &l->simple;
            } else {
                /* :122-127 — get or insert (0, now + window) BEFORE the verdict */
                lo_entry init = {0, now_us + l->window_us};
                e = &q_get_or_insert(o, ctrs[i].limit_id, ctrs[i].key_lo, ctrs[i].key_hi, init)->e;
            }
            uint64_t v = entry_value_at(e, now_us);
            uint64_t sum = v + delta; /* wrapping */
            int over = sum > l->max_value; /* counter_is_within_limits :259-264 */
            if (load_counters) {
                /* process_counter :85-96 — remaining = max.checked_sub(v+delta) or 0 */
                if (remaining) remaining[i] = over ? 0 : l->max_value - sum;
                if (over && first_limited == LO_NONE) first_limited = i;
            }
            if (over && !load_counters) {
                /* early return :110-112 / :130-132 */
                if (first_limited_out) *first_limited_out = i;
                return 1;
            }
            if (load_counters && ttl_us) ttl_us[i] = entry_ttl(e, now_us); /* :114-116,:134-136 */
            touched[nt] = e;
            touched_w[nt] = l->window_us;
            nt++;
        }
    }
    if (first_limited != LO_NONE) { /* :141-143 */
        if (first_limited_out) *first_limited_out = first_limited;
        return 1;
    }
    /* :146-153 — all-or-nothing update */
    for (uint32_t k = 0; k < nt; k++) entry_update(touched[k], delta, touched_w[k], now_us);
    return 0;
}

int lo_is_within_limits(lo_oracle *o, const lo_counter *c, uint64_t delta, uint64_t now_us) {
    lo_limit *l = get_limit(o, c->limit_id);
    if (!l) return -1;
    uint64_t v = 0;
    if (l->qualified) {
        lo_slot *s = q_find(o, c->limit_id, c->key_lo, c->key_hi);
        if (s) v = entry_value_at(&s->e, now_us);
    } else if (l->simple_present) {
        v = entry_value_at(&l->simple, now_us);
    }
    return l->max_value >= v + delta; /* in_memory.rs:34 */
}

int lo_is_rate_limited(lo_oracle *o, const lo_counter *ctrs, uint32_t m, uint64_t delta,
                       uint64_t now_us, uint32_t *first_limited) {
    if (first_limited) *first_limited = LO_NONE;
    for (uint32_t i = 0; i < m; i++) { /* lib.rs:387-409 — given order, first over wins */
        int r = lo_is_within_limits(o, &ctrs[i], delta, now_us);
        if (r < 0) return r;
        if (!r) {
            if (first_limited) *first_limited = i;
            return 1;
        }
    }
    return 0;
}

int lo_update_counter(lo_oracle *o, const lo_counter *c, uint64_t delta, uint64_t now_us) {
    lo_limit *l = get_limit(o, c->limit_id);
    if (!l) return -1;
    if (l->qualified) {
        /* in_memory.rs:50-57 */
        lo_entry init = {0, now_us + l->window_us};
        lo_slot *s = q_get_or_insert(o, c->limit_id, c->key_lo, c->key_hi, init);
        entry_update(&s->e, delta, l->window_us, now_us);
    } else if (!l->simple_present) {
        /* :60-62 — Vacant: insert (delta, now + window) */
        l->simple_present = 1;
        l->simple.value = delta;
        l->simple.expiry = now_us + l->window_us;
    } else {
        entry_update(&l->simple, delta, l->window_us, now_us); /* :63-65 */
    }
    return 0;
}

int lo_update_counters(lo_oracle *o, const lo_counter *ctrs, uint32_t m, uint64_t delta,
                       uint64_t now_us) {
    for (uint32_t i = 0; i < m; i++) {
        int r = lo_update_counter(o, &ctrs[i], delta, now_us);
        if (r < 0) return r;
    }
    return 0;
}

int lo_batch_csr(lo_oracle *o, int mode, uint64_t n, const uint32_t *off, const lo_counter *ctrs,
                 const uint64_t *delta, const uint64_t *now_us, int load_counters,
                 uint8_t *out_limited, uint32_t *out_first_limited, uint64_t *out_remaining,
                 uint64_t *out_ttl_us) {
    for (uint64_t i = 0; i < n; i++) {
        const lo_counter *c = ctrs + off[i];
        uint32_t m = off[i + 1] - off[i];
        uint32_t fl = LO_NONE;
        int r = 0;
        if (m == 0) {
            r = 0; /* lib.rs:434-440 — no counters apply ⇒ not limited, no state */
        } else if (mode == 0) {
            r = lo_check_and_update(o, c, m, delta[i], load_counters, now_us[i], &fl,
                                    out_remaining ? out_remaining + off[i] : NULL,
                                    out_ttl_us ? out_ttl_us + off[i] : NULL);
        } else if (mode == 1) {
            r = lo_is_rate_limited(o, c, m, delta[i], now_us[i], &fl);
        } else {
            r = lo_update_counters(o, c, m, delta[i], now_us[i]);
        }
        if (r < 0) return r;
        if (out_limited) out_limited[i] = (uint8_t)r;
        if (out_first_limited) out_first_limited[i] = (fl == LO_NONE) ? LO_NONE : c[fl].limit_id;
    }
    return 0;
}

int lo_batch_records(lo_oracle *o, int mode, uint64_t n, const lo_record *recs, int load_counters,
                     uint32_t stride, uint8_t *out_limited, uint32_t *out_first_limited,
                     uint64_t *out_remaining, uint64_t *out_ttl_us) {
    enum { MAXC = 64 };
    lo_counter c[MAXC];
    uint64_t rem[MAXC], ttl[MAXC];
    for (uint64_t i = 0; i < n; i++) {
        const lo_record *r = &recs[i];
        uint32_t m = (r->ns_id < o->ns_cap) ? o->ns_count[r->ns_id] : 0;
        if (m > MAXC) return -2;
        for (uint32_t k = 0; k < m; k++) {
            c[k].limit_id = o->ns_limits[r->ns_id][k];
            int q = o->limits[c[k].limit_id].qualified;
            c[k].key_lo = q ? r->key_lo : 0;
            c[k].key_hi = q ? (r->key_hi & LO_RECORD_KEY_HI_MASK) : 0; /* top byte: opaque lane */
        }
        uint32_t fl = LO_NONE;
        int res = 0;
        uint64_t d = r->hits_addend;
        if (m == 0) {
            res = 0;
        } else if (mode == 0) {
            res = lo_check_and_update(o, c, m, d, load_counters, r->now_us, &fl, rem, ttl);
            if (load_counters)
                for (uint32_t k = 0; k < m && k < stride; k++) {
                    if (out_remaining) out_remaining[i * stride + k] = rem[k];
                    if (out_ttl_us) out_ttl_us[i * stride + k] = ttl[k];
                }
        } else if (mode == 1) {
            res = lo_is_rate_limited(o, c, m, d, r->now_us, &fl);
        } else {
            res = lo_update_counters(o, c, m, d, r->now_us);
        }
        if (res < 0) return res;
        if (out_limited) out_limited[i] = (uint8_t)res;
        if (out_first_limited) out_first_limited[i] = (fl == LO_NONE) ? LO_NONE : c[fl].limit_id;
    }
    return 0;
}

static int limit_in_ns_set(lo_oracle *o, uint32_t limit_id, const uint32_t *ids, uint32_t n) {
    lo_limit *l = get_limit(o, limit_id);
    if (!l) return 0;
    for (uint32_t i = 0; i < n; i++) {
        lo_limit *g = get_limit(o, ids[i]);
        if (g && g->ns_id == l->ns_id) return 1;
    }
    return 0;
}

uint64_t lo_get_counters(lo_oracle *o, const uint32_t *limit_ids, uint32_t n, uint64_t now_us,
                         uint64_t cap, uint32_t *out_limit_id, uint64_t *out_key_lo,
                         uint64_t *out_key_hi, uint64_t *out_remaining, uint64_t *out_ttl_us) {
    /* in_memory.rs:158-187.  The first loop walks counters_in_namespace(limit.namespace())
     * (:214-238) — every simple and qualified counter of that namespace; the second adds
     * qualified counters of the given limits (a subset of the first).  The result is a
     * HashSet<Counter>, so each counter appears once.  Kept iff ttl > 0. */
    uint64_t cnt = 0;
    for (uint32_t id = 0; id < o->limits_cap; id++) {
        lo_limit *l = &o->limits[id];
        if (!l->defined || l->qualified || !l->simple_present) continue;
        if (!limit_in_ns_set(o, id, limit_ids, n)) continue;
        uint64_t ttl = entry_ttl(&l->simple, now_us);
        if (ttl == 0) continue;
        if (cnt < cap) {
            out_limit_id[cnt] = id;
            out_key_lo[cnt] = 0;
            out_key_hi[cnt] = 0;
            out_remaining[cnt] = l->max_value - entry_value_at(&l->simple, now_us);
            out_ttl_us[cnt] = ttl;
        }
        cnt++;
    }
    for (uint64_t i = 0; i < o->nslots; i++) {
        lo_slot *s = &o->slots[i];
        if (!s->state) continue;
        if (!limit_in_ns_set(o, s->limit_id, limit_ids, n)) continue;
        uint64_t ttl = entry_ttl(&s->e, now_us);
        if (ttl == 0) continue;
        if (cnt < cap) {
            out_limit_id[cnt] = s->limit_id;
            out_key_lo[cnt] = s->key_lo;
            out_key_hi[cnt] = s->key_hi;
            out_remaining[cnt] = o->limits[s->limit_id].max_value - entry_value_at(&s->e, now_us);
            out_ttl_us[cnt] = ttl;
        }
        cnt++;
    }
    return cnt;
}

int lo_delete_counters(lo_oracle *o, const uint32_t *limit_ids, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) delete_counters_of_limit(o, limit_ids[i]);
    return 0;
}

int lo_clear(lo_oracle *o) {
    /* in_memory.rs:197-201 — simple_limits.clear(); the qualified cache is untouched. */
    for (uint32_t id = 0; id < o->limits_cap; id++) {
        lo_limit *l = &o->limits[id];
        if (l->defined && !l->qualified) {
            l->simple_present = 0;
            l->simple.value = 0;
            l->simple.expiry = 0;
        }
    }
    return 0;
}

static int keep_unexpired(const lo_slot *s, void *arg) { return s->e.expiry > *(uint64_t *)arg; }

uint64_t lo_invalidate_expired(lo_oracle *o, uint64_t now_us) {
    return q_filter(o, keep_unexpired, &now_us);
}

uint64_t lo_dump(lo_oracle *o, uint64_t cap, uint32_t *out_limit_id, uint64_t *out_key_lo,
                 uint64_t *out_key_hi, uint64_t *out_value, uint64_t *out_expiry_us) {
    uint64_t cnt = 0;
    for (uint32_t id = 0; id < o->limits_cap; id++) {
        lo_limit *l = &o->limits[id];
        if (!l->defined || l->qualified || !l->simple_present) continue;
        if (cnt < cap) {
            out_limit_id[cnt] = id;
            out_key_lo[cnt] = 0;
            out_key_hi[cnt] = 0;
            out_value[cnt] = l->simple.value;
            out_expiry_us[cnt] = l->simple.expiry;
        }
        cnt++;
    }
    for (uint64_t i = 0; i < o->nslots; i++) {
        lo_slot *s = &o->slots[i];
        if (!s->state) continue;
        if (cnt < cap) {
            out_limit_id[cnt] = s->limit_id;
            out_key_lo[cnt] = s->key_lo;
            out_key_hi[cnt] = s->key_hi;
            out_value[cnt] = s->e.value;
            out_expiry_us[cnt] = s->e.expiry;
        }
        cnt++;
    }
    return cnt;
}

uint64_t lo_size(lo_oracle *o) {
    uint64_t cnt = o->nfull;
    for (uint32_t id = 0; id < o->limits_cap; id++)
        if (o->limits[id].defined && !o->limits[id].qualified && o->limits[id].simple_present) cnt++;
    return cnt;
}

/* ---- multi-threaded CPU baseline --------------------------------------------------- */
/* T persistent oracles; every namespace is owned by one of them (all counters
 * of a request belong to its namespace), chosen once by longest-processing-time-first on
 * the first run's per-namespace request counts so the threads are balanced under Zipf. */
struct lo_mt;
typedef struct {
    struct lo_mt *m;
    uint32_t t;
    lo_oracle *o;
    const lo_record *recs;
    const uint32_t *idx;
    uint64_t n;
    uint8_t *out_limited;
} mt_arg;

/* The workers are PERSISTENT and PINNED (one per CPU of the process's affinity mask, round robin): a run
 * only hands them their index lists between two barriers, so the timed region holds no thread creation, no
 * join and no migration — the baseline does not swing with the scheduler's mood. */
struct lo_mt {
    uint32_t threads;
    lo_oracle **o;
    uint32_t *ns_owner; /* [ns_cap] */
    uint32_t ns_cap;
    int assigned;
    pthread_t *tids;
    mt_arg *args;
    pthread_barrier_t start, end;
    volatile int stop;
    uint32_t pinned; /* workers whose pthread_setaffinity_np succeeded */
};

static void *mt_worker(void *p) {
    mt_arg *a = (mt_arg *)p;
    struct lo_mt *m = a->m;
    enum { MAXC = 64 };
    lo_counter c[MAXC];
    for (;;) {
        pthread_barrier_wait(&m->start);
        if (m->stop) break;
        lo_oracle *o = a->o;
        for (uint64_t j = 0; j < a->n; j++) {
            uint32_t i = a->idx[j];
            const lo_record *r = &a->recs[i];
            uint32_t k_n = (r->ns_id < o->ns_cap) ? o->ns_count[r->ns_id] : 0;
            for (uint32_t k = 0; k < k_n; k++) {
                c[k].limit_id = o->ns_limits[r->ns_id][k];
                int q = o->limits[c[k].limit_id].qualified;
                c[k].key_lo = q ? r->key_lo : 0;
                c[k].key_hi = q ? (r->key_hi & LO_RECORD_KEY_HI_MASK) : 0; /* top byte: opaque lane */
            }
            int res = k_n ? lo_check_and_update(o, c, k_n, r->hits_addend, 0, r->now_us, NULL, NULL, NULL) : 0;
            a->out_limited[i] = (uint8_t)(res > 0);
        }
        pthread_barrier_wait(&m->end);
    }
    return NULL;
}

lo_mt *lo_mt_create(const lo_limit_desc *limits, uint32_t n_limits, uint32_t threads, uint64_t capacity_hint) {
    if (threads == 0) threads = 1;
    lo_mt *m = (lo_mt *)calloc(1, sizeof(*m));
    m->threads = threads;
    m->o = (lo_oracle **)calloc(threads, sizeof(lo_oracle *));
    uint32_t ns_cap = 1;
    for (uint32_t k = 0; k < n_limits; k++)
        if (limits[k].ns_id + 1 > ns_cap) ns_cap = limits[k].ns_id + 1;
    m->ns_cap = ns_cap;
    m->ns_owner = (uint32_t *)calloc(ns_cap, sizeof(uint32_t));
    for (uint32_t t = 0; t < threads; t++) {
        lo_oracle *o = lo_create(capacity_hint / threads + 1024);
        /* fault the table in now: the timed runs measure the data path, not page faults */
        memset(o->slots, 0xff, o->nslots * sizeof(lo_slot));
        memset(o->slots, 0, o->nslots * sizeof(lo_slot));
        for (uint32_t k = 0; k < n_limits; k++)
            lo_limit_set(o, limits[k].limit_id, limits[k].ns_id, limits[k].max_value, limits[k].window_us,
                         (int)limits[k].qualified);
        m->o[t] = o;
    }
    /* persistent workers, pinned round robin over the CPUs this process may run on */
    cpu_set_t allowed;
    int cpus[CPU_SETSIZE], ncpu = 0;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; c++)
            if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
    m->tids = (pthread_t *)calloc(threads, sizeof(pthread_t));
    m->args = (mt_arg *)calloc(threads, sizeof(mt_arg));
    pthread_barrier_init(&m->start, NULL, threads + 1);
    pthread_barrier_init(&m->end, NULL, threads + 1);
    for (uint32_t t = 0; t < threads; t++) {
        m->args[t].m = m;
        m->args[t].t = t;
        m->args[t].o = m->o[t];
        pthread_create(&m->tids[t], NULL, mt_worker, &m->args[t]);
        if (ncpu > 0) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[t % (uint32_t)ncpu], &one);
            if (pthread_setaffinity_np(m->tids[t], sizeof one, &one) == 0) m->pinned++;
        }
    }
    return m;
}

uint32_t lo_mt_pinned(lo_mt *m) { return m ? m->pinned : 0; }

void lo_mt_destroy(lo_mt *m) {
    if (!m) return;
    m->stop = 1;
    pthread_barrier_wait(&m->start);
    for (uint32_t t = 0; t < m->threads; t++) pthread_join(m->tids[t], NULL);
    pthread_barrier_destroy(&m->start);
    pthread_barrier_destroy(&m->end);
    for (uint32_t t = 0; t < m->threads; t++) lo_destroy(m->o[t]);
    free(m->tids);
    free(m->args);
    free(m->o);
    free(m->ns_owner);
    free(m);
}

double lo_mt_run(lo_mt *m, uint64_t n, const lo_record *recs, uint8_t *out_limited) {
    const uint32_t T = m->threads;
    if (!m->assigned) {
        uint64_t *load = (uint64_t *)calloc(m->ns_cap, sizeof(uint64_t));
        uint64_t *tl = (uint64_t *)calloc(T, sizeof(uint64_t));
        uint8_t *done = (uint8_t *)calloc(m->ns_cap, 1);
        for (uint64_t i = 0; i < n; i++)
            if (recs[i].ns_id < m->ns_cap) load[recs[i].ns_id]++;
        for (uint32_t round = 0; round < m->ns_cap; round++) {
            uint32_t best = 0;
            int found = 0;
            for (uint32_t ns = 0; ns < m->ns_cap; ns++)
                if (!done[ns] && (!found || load[ns] > load[best])) {
                    best = ns;
                    found = 1;
                }
            uint32_t tmin = 0;
            for (uint32_t t = 1; t < T; t++)
                if (tl[t] < tl[tmin]) tmin = t;
            m->ns_owner[best] = tmin;
            tl[tmin] += load[best] + 1;
            done[best] = 1;
        }
        free(load);
        free(tl);
        free(done);
        m->assigned = 1;
    }
    uint64_t *cnt = (uint64_t *)calloc(T, sizeof(uint64_t));
    uint32_t **idx = (uint32_t **)calloc(T, sizeof(uint32_t *));
#define OWNER(i) (recs[i].ns_id < m->ns_cap ? m->ns_owner[recs[i].ns_id] : 0)
    for (uint64_t i = 0; i < n; i++) cnt[OWNER(i)]++;
    for (uint32_t t = 0; t < T; t++) {
        idx[t] = (uint32_t *)malloc((cnt[t] + 1) * sizeof(uint32_t));
        cnt[t] = 0;
    }
    for (uint64_t i = 0; i < n; i++) {
        uint32_t t = OWNER(i);
        idx[t][cnt[t]++] = (uint32_t)i;
    }
#undef OWNER
    for (uint32_t t = 0; t < T; t++) {
        m->args[t].recs = recs;
        m->args[t].idx = idx[t];
        m->args[t].n = cnt[t];
        m->args[t].out_limited = out_limited;
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&m->start); /* the workers are parked on this barrier */
    pthread_barrier_wait(&m->end);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (uint32_t t = 0; t < T; t++) free(idx[t]);
    free(idx);
    free(cnt);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

double lo_bench_records_mt(const lo_limit_desc *limits, uint32_t n_limits, uint64_t n,
                           const lo_record *recs, uint32_t threads, uint64_t capacity_hint,
                           uint8_t *out_limited) {
    lo_mt *m = lo_mt_create(limits, n_limits, threads, capacity_hint);
    double t = lo_mt_run(m, n, recs, out_limited);
    lo_mt_destroy(m);
    return t;
}
