/* rl_rls.h — the Envoy RLS v3 wire surface in front of the engine (with the f1 matcher inside).
 *
 * Replaces, batched and without a protobuf runtime,
 *   MyRateLimiter::should_rate_limit        limitador-server/src/envoy_rls/server.rs:91-208
 *   KuadrantService::check_rate_limit       limitador-server/src/envoy_rls/kuadrant_service.rs:27-107
 *   KuadrantService::report                 limitador-server/src/envoy_rls/kuadrant_service.rs:109-186
 *   PrometheusMetrics::incr_*               limitador-server/src/prometheus_metrics.rs:93-125 (namespace and
 *                                           limit-name labels; CEL custom labels are out of scope)
 * over the messages of
 *   envoy.service.ratelimit.v3.RateLimitRequest / RateLimitResponse
 *       limitador-server/vendor/protobufs/data-plane-api/envoy/service/ratelimit/v3/rls.proto
 *   envoy.extensions.common.ratelimit.v3.RateLimitDescriptor (entries key/value)
 *   envoy.config.core.v3.HeaderValue (key, value)
 *
 * A batch of n wire requests is served in three stages:
 *   plan    (CPU, `threads` workers): decode every request, build its CEL context
 *           (`descriptors[i]` = the i-th descriptor's entries as a map, last duplicate key wins — server.rs:121-127),
 *           run counters_that_apply (include/rl_match.h) and lay the counters out as the CSR that
 *           rl_check_and_update_batch / rl_is_within_limits_batch / rl_update_batch take.  Requests that never reach
 *           the store are answered here: domain "" -> overall_code UNKNOWN (server.rs:106-116); no limit applies ->
 *           OK (lib.rs:434-440); hits_addend 0 -> 1 (server.rs:131-135).
 *   decide  (GPU): ONE engine call for the whole batch; array order is the stream order that defines the result.
 *   finish  (CPU): verdicts (+ remaining / ttl with draft-03 headers) -> RateLimitResponse bytes, the three
 *           X-RateLimit-* headers sorted by key (server.rs:45-57, lib.rs:235-275), per-namespace metrics.
 * rl_rls_serve runs the three stages through the engine given at creation.  plan / finish are exported on their own so
 * that the CPU stages can be driven (and tested) without a GPU; the product never decides on the CPU.
 *
 * gRPC status per request (what tonic would put in grpc-status): 0 OK with a response body; 13 INTERNAL for a
 * message that does not decode (prost DecodeError); 14 UNAVAILABLE "Service unavailable" when the store call fails
 * (server.rs:160-172) — those two have no body.
 */
#ifndef RL_RLS_H
#define RL_RLS_H

#include <stdint.h>

#include "rl_engine.h"
#include "rl_match.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rl_rls rl_rls;

enum { RL_RLS_CODE_UNKNOWN = 0, RL_RLS_CODE_OK = 1, RL_RLS_CODE_OVER_LIMIT = 2 }; /* RateLimitResponse.Code */
enum { RL_RLS_HEADERS_NONE = 0, RL_RLS_HEADERS_DRAFT_VERSION_03 = 1 };            /* server.rs:38-42 */
enum {
    RL_RLS_SHOULD_RATE_LIMIT = 0, /* check_rate_limited_and_update(ns, ctx, hits_addend, headers != NONE) */
    RL_RLS_CHECK_RATE_LIMIT = 1,  /* is_rate_limited(ns, ctx, 1): read-only, no headers */
    RL_RLS_REPORT = 2             /* update_counters(ns, ctx, hits_addend): always OK */
};
enum { RL_GRPC_OK = 0, RL_GRPC_INTERNAL = 13, RL_GRPC_UNAVAILABLE = 14 };
#define RL_RLS_NO_STORE 0xFFFFFFFFu

/* ---- wire codec (usable on its own) ------------------------------------------------------------------------ */
typedef struct rl_rls_entry {
    uint32_t descriptor;        /* index of the descriptor the entry belongs to */
    uint32_t key_off, key_len;  /* byte ranges inside the request buffer (not NUL-terminated) */
    uint32_t val_off, val_len;
} rl_rls_entry;
typedef struct rl_rls_request {
    uint32_t domain_off, domain_len;
    uint32_t hits_addend;   /* as on the wire: 0 when absent (the service turns it into 1) */
    uint32_t n_descriptors; /* descriptors without entries count too: they shift the indices of the later ones */
    uint32_t n_entries;     /* entries found (may exceed cap_entries: then only the first cap_entries are written) */
} rl_rls_request;
/* Decode one RateLimitRequest.  RL_OK, or RL_FATAL for a malformed message (truncated varint / length, wire type
 * that does not fit the field, field number 0, group nesting, invalid UTF-8 in a string field — what prost refuses). */
int rl_rls_decode_request(const uint8_t *buf, uint64_t len, rl_rls_request *out, rl_rls_entry *entries,
                          uint32_t cap_entries);
/* Encode a RateLimitResponse {overall_code, response_headers_to_add = n_headers x HeaderValue{key, value}} (proto3:
 * zero / empty fields are not written).  *out_len = bytes needed; RL_FATAL if cap is too small. */
int rl_rls_encode_response(uint32_t overall_code, const char *const *keys, const char *const *values,
                           uint32_t n_headers, uint8_t *out, uint64_t cap, uint64_t *out_len);

/* ---- the service ---------------------------------------------------------------------------------------------- */
/* engine may be NULL (plan / finish only).  threads = workers of the plan and finish stages (0 = one per online CPU,
 * at most 64).  use_limit_name_label: limited_calls carries limit_name too (prometheus_metrics.rs:109-117). */
int rl_rls_create(rl_matcher *m, rl_engine *engine, int header_mode, uint32_t threads, int use_limit_name_label,
                  rl_rls **out);
void rl_rls_destroy(rl_rls *s);
const char *rl_rls_last_error(rl_rls *s);

/* Stage 1.  Request i = buf[off[i] .. off[i+1]).  now_us = the batch's clock reading (0 = wall clock now). */
int rl_rls_plan(rl_rls *s, int method, uint64_t n, const uint8_t *buf, const uint64_t *off, uint64_t now_us);
/* The store call of the planned batch: n_store requests (a subset of the batch, in batch order) as CSR arrays owned by
 * the service, valid until the next plan.  store_index[i] (nullable out, n entries) = position of request i in the
 * store call or RL_RLS_NO_STORE. */
int rl_rls_plan_view(rl_rls *s, uint64_t *out_n_store, const uint32_t **out_ctr_off, const rl_counter **out_ctrs,
                     const uint64_t **out_delta, const uint64_t **out_now_us, int *out_load_counters,
                     const uint32_t **out_store_index);
/* Stage 3 (once per planned batch).  store_status = status of the store call (non-OK: every store request is answered
 * UNAVAILABLE); limited / first_limited: n_store entries; remaining / ttl_us: one per counter (only read with headers). */
int rl_rls_finish(rl_rls *s, int store_status, const uint8_t *limited, const uint32_t *first_limited,
                  const uint64_t *remaining, const uint64_t *ttl_us);
/* Responses of the last finished batch: response i = (*out_buf)[(*out_off)[i] .. (*out_off)[i+1]) (empty for a
 * non-OK gRPC status and for overall_code UNKNOWN, which encodes to zero bytes), (*out_grpc)[i] its status,
 * (*out_code)[i] the overall_code.  Valid until the next plan. */
int rl_rls_responses(rl_rls *s, const uint8_t **out_buf, const uint64_t **out_off, const uint8_t **out_grpc,
                     const uint8_t **out_code);
/* plan -> the engine -> finish. */
int rl_rls_serve(rl_rls *s, int method, uint64_t n, const uint8_t *buf, const uint64_t *off, uint64_t now_us);

/* Prometheus text exposition of authorized_calls / authorized_hits / limited_calls (sorted by label values) plus
 * `limitador_up 1`: lines `name{limitador_namespace="ns"[,limit_name="x"]} value` as
 * metrics_exporter_prometheus renders them (prometheus_metrics.rs:415-447).  *out_len = bytes needed incl. NUL. */
int rl_rls_metrics_render(rl_rls *s, char *out, uint64_t cap, uint64_t *out_len);
/* Stage timings of the last serve call in microseconds: plan, store call, finish. */
int rl_rls_last_timings(rl_rls *s, double *out_plan_us, double *out_store_us, double *out_finish_us);

#ifdef __cplusplus
}
#endif
#endif /* RL_RLS_H */
