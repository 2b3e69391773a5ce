/* crdt_oracle.c — CPU restatement of the reference's replicated counter value, explicit clock.
 * TEST INFRASTRUCTURE, NOT THE PRODUCT: only tests/ load it (through oracle/crdt_binding.py).
 *
 * Restates, with `when` passed in instead of SystemTime::now() (paths relative to /root/reference/limitador/src/storage/):
 *   distributed/cr_counter_value.rs:20-29    CrCounterValue::new    expiry = now + window, every value 0
 *   distributed/cr_counter_value.rs:38-46    read_at                expired_at(when) ? 0 : others.sum() + value
 *   distributed/cr_counter_value.rs:54-60    inc_at                 update_if_expired ? value = inc : value += inc
 *   distributed/cr_counter_value.rs:66-75    inc_actor_at           ours -> inc_at; another actor: the same rule on its entry
 *   distributed/cr_counter_value.rs:81-115   merge_at               see cro_merge_at below, line by line
 *   distributed/cr_counter_value.rs:142-147  reset
 *   atomic_expiring_value.rs:76-79           expired_at             expiry <= when (inclusive)
 *   atomic_expiring_value.rs:87-99           update_if_expired      expiry <= when ? (expiry = when + ttl, true) : false
 *   atomic_expiring_value.rs:113-130         AtomicExpiryTime::merge_at   other < ours && other > when ? ours = other
 *   distributed/mod.rs:65-91                 update_counter         vacant -> new(window) then inc_at, same `now`
 *   distributed/mod.rs:294-332               process_re_sync        our own value, if non-zero and unexpired
 * Pinned by the reference's eleven unit tests (cr_counter_value.rs:177-300), ported with explicit clocks in
 * tests/test_crdt.py.  Unpinned by any reference test (the oracle's choice, shared with the GPU path): a gossiped key
 * without a local counter — the reference panics (distributed/mod.rs:242) — is created as a locally expired one.
 *
 * A counter's actors are indices 0..actors-1 (the reference keys them by the peer's identifier string); a vacant
 * entry of the `others` map and an entry holding 0 read the same, so a dense array is an exact model.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CRO_MAX_ACTORS 16

typedef struct {
    uint64_t lo, hi;
    uint64_t expiry;
    uint64_t val[CRO_MAX_ACTORS];
    int used;
} cro_entry;

typedef struct cro {
    uint32_t actors, self_actor;
    uint64_t cap, n;
    cro_entry *e;
} cro;

static uint64_t mix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

cro *cro_create(uint32_t actors, uint32_t self_actor) {
    if (actors < 1 || actors > CRO_MAX_ACTORS || self_actor >= actors) return NULL;
    cro *o = calloc(1, sizeof *o);
    o->actors = actors;
    o->self_actor = self_actor;
    o->cap = 1024;
    o->e = calloc(o->cap, sizeof(cro_entry));
    return o;
}
void cro_destroy(cro *o) {
    if (!o) return;
    free(o->e);
    free(o);
}

static cro_entry *slot(cro_entry *tab, uint64_t cap, uint64_t lo, uint64_t hi) {
    uint64_t p = mix(lo ^ mix(hi)) & (cap - 1);
    while (tab[p].used && !(tab[p].lo == lo && tab[p].hi == hi)) p = (p + 1) & (cap - 1);
    return &tab[p];
}
static cro_entry *find(cro *o, uint64_t lo, uint64_t hi, int create) {
    cro_entry *s = slot(o->e, o->cap, lo, hi);
    if (s->used || !create) return s->used ? s : NULL;
    if ((o->n + 1) * 2 > o->cap) {
        const uint64_t nc = o->cap * 2;
        cro_entry *nt = calloc(nc, sizeof(cro_entry));
        for (uint64_t i = 0; i < o->cap; i++)
            if (o->e[i].used) *slot(nt, nc, o->e[i].lo, o->e[i].hi) = o->e[i];
        free(o->e);
        o->e = nt;
        o->cap = nc;
        s = slot(o->e, o->cap, lo, hi);
    }
    memset(s, 0, sizeof *s);
    s->lo = lo;
    s->hi = hi;
    s->used = 1;
    o->n++;
    return s;
}

/* CrCounterValue::new (:20-29): only through the callers below */
static cro_entry *new_at(cro *o, uint64_t lo, uint64_t hi, uint64_t window_us, uint64_t when) {
    cro_entry *c = find(o, lo, hi, 1);
    c->expiry = when + window_us;
    return c;
}

/* update_counter (distributed/mod.rs:65-91) + inc_actor_at (:66-75) */
void cro_inc_actor_at(cro *o, uint64_t lo, uint64_t hi, uint32_t actor, uint64_t inc, uint64_t window_us, uint64_t when) {
    cro_entry *c = find(o, lo, hi, 0);
    if (!c) c = new_at(o, lo, hi, window_us, when); /* Entry::Vacant: new(.., duration), then the increment at `now` */
    if (c->expiry <= when) {                        /* update_if_expired (atomic_expiring_value.rs:87-99) */
        c->expiry = when + window_us;
        c->val[actor] = inc;                        /* value.store(increment) / guard.insert(actor, increment) */
    } else {
        c->val[actor] += inc;                       /* fetch_add / *entry.or_insert(0) += increment */
    }
}

/* merge_at (:81-115) of the remote set (expiry, {actors[j]: values[j]}) */
void cro_merge_at(cro *o, uint64_t lo, uint64_t hi, uint64_t other_expiry, const uint32_t *actors, const uint64_t *values,
                  uint32_t n, uint64_t when) {
    if (!(other_expiry > when)) return; /* :83 — an expired set is ignored */
    cro_entry *c = find(o, lo, hi, 0);
    if (!c) c = find(o, lo, hi, 1); /* reference: unwrap() panics; here: a locally expired counter (expiry 0) */
    /* :84 — AtomicExpiryTime::merge_at (atomic_expiring_value.rs:116): the earlier unexpired expiry wins */
    if (other_expiry < c->expiry && other_expiry > when) c->expiry = other_expiry;
    if (c->expiry <= when) { /* :85-87 — reset(expiry): expiry = theirs, value = 0, others.clear() */
        c->expiry = other_expiry;
        memset(c->val, 0, sizeof c->val);
    }
    for (uint32_t j = 0; j < n; j++) { /* :90-112 */
        const uint32_t a = actors[j];
        if (values[j] > c->val[a]) c->val[a] = values[j]; /* ours: fetch_add(other - ours); others: insert / max */
    }
}

/* read_at (:38-46); *out_expiry (nullable) = the counter's expiry, 0 when absent */
uint64_t cro_read_at(cro *o, uint64_t lo, uint64_t hi, uint64_t when, uint64_t *out_expiry) {
    cro_entry *c = find(o, lo, hi, 0);
    if (out_expiry) *out_expiry = c ? c->expiry : 0;
    if (!c || c->expiry <= when) return 0;
    uint64_t s = 0;
    for (uint32_t a = 0; a < o->actors; a++) s += c->val[a];
    return s;
}

/* process_re_sync (distributed/mod.rs:302-318) */
uint64_t cro_export(cro *o, uint64_t when, uint64_t cap, uint64_t *lo, uint64_t *hi, uint64_t *value, uint64_t *expiry) {
    uint64_t n = 0;
    for (uint64_t i = 0; i < o->cap; i++) {
        const cro_entry *c = &o->e[i];
        if (!c->used || c->val[o->self_actor] == 0 || c->expiry <= when) continue;
        if (n < cap) {
            lo[n] = c->lo;
            hi[n] = c->hi;
            value[n] = c->val[o->self_actor];
            expiry[n] = c->expiry;
        }
        n++;
    }
    return n;
}

uint64_t cro_dump(cro *o, uint64_t cap, uint64_t *lo, uint64_t *hi, uint64_t *expiry, uint64_t *values) {
    uint64_t n = 0;
    for (uint64_t i = 0; i < o->cap; i++) {
        const cro_entry *c = &o->e[i];
        if (!c->used) continue;
        if (n < cap) {
            lo[n] = c->lo;
            hi[n] = c->hi;
            expiry[n] = c->expiry;
            for (uint32_t a = 0; a < o->actors; a++) values[n * o->actors + a] = c->val[a];
        }
        n++;
    }
    return n;
}
