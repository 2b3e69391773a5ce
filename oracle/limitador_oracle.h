/*
 * limitador_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT).
 *
 * A plain-C, single-threaded restatement of the reference's in-memory rate-limit
 * hot path with an explicit `now_us` instead of the wall clock:
 *
 *   limitador/src/storage/atomic_expiring_value.rs:19-47,62-99,151-158   (arithmetic)
 *   limitador/src/storage/in_memory.rs:20-35,38-44,47-69,72-156,158-201,241-264
 *   limitador/src/lib.rs:362-464                                          (orchestration)
 *   limitador/src/storage/mod.rs:60-83                                    (add/update limit)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (limitador_b200) never does.
 *
 * Parity pin: the reference cannot be built here (no rustc/cargo) and ships no golden
 * vector files; this oracle is pinned by porting the reference's own known-answer unit
 * and integration tests (tests/test_oracle_kats.py cites each one by file:line), and
 * cross-checked on random streams against an independent Python restatement of the same
 * reference files (tests/spec_model.py, tests/test_oracle_vs_spec.py).
 */
#ifndef LIMITADOR_ORACLE_H
#define LIMITADOR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lo_oracle lo_oracle;

/* One counter a request touches: the limit (by caller-interned id) plus the digest of
 * its resolved variable values (ignored for unqualified limits).
 * counter.rs:10-17,123-138 — identity = limit + set_variables. */
typedef struct {
    uint32_t limit_id;
    uint32_t _pad;
    uint64_t key_lo;
    uint64_t key_hi;
} lo_counter;

/* 32-byte synthetic request record (SURVEY §8(d) "common synthetic encoding"): the limit
 * set is implied by ns_id = every limit registered for that namespace, in registration
 * order, all qualified ones sharing (key_lo,key_hi). */
typedef struct {
    uint32_t ns_id;
    uint32_t hits_addend;
    uint64_t key_lo;
    uint64_t key_hi;
    uint64_t now_us;
} lo_record;
/* top byte of lo_record.key_hi: opaque "lane" byte, not part of the counter identity
 * (include/rl_engine.h, rl_record) */
#define LO_RECORD_KEY_HI_MASK 0x00FFFFFFFFFFFFFFull

#define LO_NONE 0xFFFFFFFFu

lo_oracle *lo_create(uint64_t capacity_hint);
void lo_destroy(lo_oracle *o);

/* Storage::add_limit (storage/mod.rs:60-65) + update_limit (:67-83).
 * New id: registers the limit; unqualified ⇒ add_counter pre-creates (0, EPOCH)
 * (in_memory.rs:38-44).  Existing id: only max_value may change.  Returns 0, or -1 on
 * a change of ns/window/qualified for a live id. */
int lo_limit_set(lo_oracle *o, uint32_t limit_id, uint32_t ns_id, uint64_t max_value,
                 uint64_t window_us, int qualified);
/* Storage::delete_limit (storage/mod.rs:93-117): delete_counters + forget the limit. */
int lo_limit_delete(lo_oracle *o, uint32_t limit_id);

/* InMemoryStorage::check_and_update (in_memory.rs:72-156).  Counters are processed
 * unqualified-first (stable), then qualified, in the given order.  Returns 0 = Ok,
 * 1 = Limited, <0 = error (unknown limit / unqualified counter never added — the
 * reference panics at in_memory.rs:107).  first_limited = index INTO ctrs of the
 * limited counter the reference would name, or LO_NONE.  remaining/ttl_us (nullable)
 * are written per counter (indexed like ctrs) only when load_counters. */
int lo_check_and_update(lo_oracle *o, const lo_counter *ctrs, uint32_t m, uint64_t delta,
                        int load_counters, uint64_t now_us, uint32_t *first_limited,
                        uint64_t *remaining, uint64_t *ttl_us);

/* InMemoryStorage::is_within_limits (in_memory.rs:20-35): 1 within, 0 over, <0 error. */
int lo_is_within_limits(lo_oracle *o, const lo_counter *c, uint64_t delta, uint64_t now_us);
/* RateLimiter::is_rate_limited (lib.rs:362-409): given order, stop at first over.
 * Returns 0 = not limited, 1 = limited (first_limited = index), <0 error. */
int lo_is_rate_limited(lo_oracle *o, const lo_counter *ctrs, uint32_t m, uint64_t delta,
                       uint64_t now_us, uint32_t *first_limited);
/* InMemoryStorage::update_counter (in_memory.rs:47-69). */
int lo_update_counter(lo_oracle *o, const lo_counter *c, uint64_t delta, uint64_t now_us);
/* RateLimiter::update_counters (lib.rs:411-423). */
int lo_update_counters(lo_oracle *o, const lo_counter *ctrs, uint32_t m, uint64_t delta,
                       uint64_t now_us);

/* Sequential batch drivers (one call per request, in stream order).  CSR layout:
 * request i owns ctrs[off[i] .. off[i+1]).  out_first_limited holds the LIMIT ID (not
 * index) of the named counter or LO_NONE; remaining/ttl indexed like ctrs.
 * mode: 0 = check_and_update, 1 = is_rate_limited, 2 = update_counters. */
int lo_batch_csr(lo_oracle *o, int mode, uint64_t n, const uint32_t *off, const lo_counter *ctrs,
                 const uint64_t *delta, const uint64_t *now_us, int load_counters,
                 uint8_t *out_limited, uint32_t *out_first_limited, uint64_t *out_remaining,
                 uint64_t *out_ttl_us);
/* Record format: out_remaining/out_ttl_us are [n * stride]; slot k of request i is the
 * k-th limit of the namespace in registration order. */
int lo_batch_records(lo_oracle *o, int mode, uint64_t n, const lo_record *recs, int load_counters,
                     uint32_t stride, uint8_t *out_limited, uint32_t *out_first_limited,
                     uint64_t *out_remaining, uint64_t *out_ttl_us);

/* InMemoryStorage::get_counters (in_memory.rs:158-187): every counter (simple and
 * qualified) of the namespaces of the given limits whose ttl(now) > 0;
 * remaining = max - value_at(now) (wrapping).  Returns the count (may exceed cap; only
 * the first cap are written). */
uint64_t lo_get_counters(lo_oracle *o, const uint32_t *limit_ids, uint32_t n, uint64_t now_us,
                         uint64_t cap, uint32_t *out_limit_id, uint64_t *out_key_lo,
                         uint64_t *out_key_hi, uint64_t *out_remaining, uint64_t *out_ttl_us);
/* InMemoryStorage::delete_counters (in_memory.rs:189-195,241-257). */
int lo_delete_counters(lo_oracle *o, const uint32_t *limit_ids, uint32_t n);
/* InMemoryStorage::clear (in_memory.rs:197-201): drops ONLY the unqualified map. */
int lo_clear(lo_oracle *o);
/* Oracle-only event mirroring moka eviction / the GPU sweep kernel (SURVEY §7 hard
 * part 3c): drop every qualified entry with expiry <= now_us. Returns #dropped. */
uint64_t lo_invalidate_expired(lo_oracle *o, uint64_t now_us);

/* Full state dump, unordered: every present entry (unqualified and qualified). */
uint64_t lo_dump(lo_oracle *o, uint64_t cap, uint32_t *out_limit_id, uint64_t *out_key_lo,
                 uint64_t *out_key_hi, uint64_t *out_value, uint64_t *out_expiry_us);
uint64_t lo_size(lo_oracle *o);

/* CPU baseline: T independent oracles, requests sharded by ns_id % T (each namespace
 * single-owner), processed concurrently on T pthreads.  Limits are
 * given once and replicated.  Returns elapsed seconds of the processing phase only
 * (sharding/bucketing excluded), writes verdicts to out_limited. */
typedef struct {
    uint32_t limit_id, ns_id;
    uint64_t max_value, window_us;
    uint32_t qualified, _pad;
} lo_limit_desc;
/* Persistent form: the T oracles stay warm across runs (as the GPU table does). */
typedef struct lo_mt lo_mt;
lo_mt *lo_mt_create(const lo_limit_desc *limits, uint32_t n_limits, uint32_t threads, uint64_t capacity_hint);
double lo_mt_run(lo_mt *m, uint64_t n, const lo_record *recs, uint8_t *out_limited);
void lo_mt_destroy(lo_mt *m);
/* workers pinned to a CPU of their own (pthread_setaffinity_np succeeded) */
uint32_t lo_mt_pinned(lo_mt *m);
double lo_bench_records_mt(const lo_limit_desc *limits, uint32_t n_limits, uint64_t n,
                           const lo_record *recs, uint32_t threads, uint64_t capacity_hint,
                           uint8_t *out_limited);

#ifdef __cplusplus
}
#endif
#endif
