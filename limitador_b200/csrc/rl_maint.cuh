// rl_maint.cuh — maintenance kernels beside the hot path (SURVEY.md §8 f3 and VERDICT r1 item 10):
//   k_ns_metrics        per-namespace authorized_calls / authorized_hits / limited_calls of a decided batch as ONE
//                       segmented reduction (limitador-server/src/prometheus_metrics.rs:93-125 increments them
//                       once per request on the host, after the decision: envoy_rls/server.rs:183-195)
//   k_region_census     live rows and tombstones per table region
//   k_compact_move /    tombstone reclamation: a region whose probe chains are lengthened by tombstones (rl_sweep
//   k_compact_reinsert  leaves them) is rebuilt in place — its rows go to a scratch slab, the region is cleared, the
//                       rows that still hold a counter are inserted again by the rule the hot path probes by
//                       (rl_kernels.cuh rl_probe: home = low hash bits, linear probing inside the region, 128-bit
//                       CAS on the header).  No reference analog (moka evicts, in_memory.rs:205-212); observable
//                       state is unchanged: rl_dump_table before == after.
//
// Written so that the SAME source runs under tests/emu/cuda_shim.h (one CUDA thread after the other on the host):
// grid-stride / one-item-per-thread kernels, global atomics only; warp-aggregated fast paths sit inside
// `#ifndef RL_SHIM` and produce the same sums.
#pragma once
#include <stdint.h>

#include "rl_core.h"
#include "rl_devmem.cuh"

// ---------------------------------------------------------------------------------------------------------------------
// Per-namespace metrics.
struct RlNsMetricsDev {
    unsigned long long* authorized_calls;  // [ns_cap]   requests allowed
    unsigned long long* authorized_hits;   // [ns_cap]   sum of their hits_addend
    unsigned long long* limited_calls;     // [ns_cap]   requests limited
    unsigned long long* limited_by_limit;  // [limits_cap] limited requests by the limit named (limit_name label), nullable
    unsigned long long* dropped;           // [1] requests not counted: error verdict, or a namespace id out of range
    uint32_t ns_cap, limits_cap;
};

// rec_words = 4: 32-byte rl_record (word 0 = ns_id | hits_addend << 32); 2: 16-byte rl_record16 (word 0 = ns_id:24 |
// hits:8 | key_hi:32).  One request per thread and loop trip; every lane of a warp makes the same number of trips.
__global__ void k_ns_metrics(const unsigned long long* __restrict__ recs, uint32_t rec_words, uint32_t n,
                             const uint8_t* __restrict__ limited, const uint32_t* __restrict__ first_limited,
                             RlNsMetricsDev M) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_up = ((uint64_t)n + 31u) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_up; i += stride) {
        bool valid = i < n;
        uint32_t ns = 0, hits = 0, v = 0, fl = RL_NONE_U32;
        if (valid) {
            const unsigned long long w0 = recs[i * rec_words];
            if (rec_words == 4) {
                ns = (uint32_t)w0;
                hits = (uint32_t)(w0 >> 32);
            } else {
                ns = (uint32_t)w0 & 0x00FFFFFFu;
                hits = (uint32_t)(w0 >> 24) & 0xFFu;
            }
            v = limited[i];
            if (v == 0xFFu || ns >= M.ns_cap) {  // RL_VERDICT_ERROR: the request was not decided
                valid = false;
                atomicAdd(M.dropped, 1ull);
            } else if (v && first_limited) {
                fl = first_limited[i];
            }
        }
#ifndef RL_SHIM
        // warp-aggregated: the lanes that hold the same (namespace, verdict) add once
        const unsigned lane = threadIdx.x & 31u;
        const unsigned key = valid ? ((ns << 1) | (v ? 1u : 0u)) : 0xFFFFFFFFu;
        const unsigned grp = __match_any_sync(0xFFFFFFFFu, key);
        const unsigned lo = __reduce_add_sync(grp, hits & 0xFFFFu), hi = __reduce_add_sync(grp, hits >> 16);
        if (valid && (unsigned)(__ffs(grp) - 1) == lane) {
            const unsigned long long cnt = (unsigned long long)__popc(grp);
            if (v) {
                atomicAdd(&M.limited_calls[ns], cnt);
            } else {
                atomicAdd(&M.authorized_calls[ns], cnt);
                atomicAdd(&M.authorized_hits[ns], (unsigned long long)lo + ((unsigned long long)hi << 16));
            }
        }
        if (M.limited_by_limit) {
            const bool named = valid && v && fl < M.limits_cap;
            const unsigned g2 = __match_any_sync(0xFFFFFFFFu, named ? fl : 0xFFFFFFFFu);
            if (named && (unsigned)(__ffs(g2) - 1) == lane) atomicAdd(&M.limited_by_limit[fl], (unsigned long long)__popc(g2));
        }
#else
        if (valid) {
            if (v) {
                atomicAdd(&M.limited_calls[ns], 1ull);
                if (M.limited_by_limit && fl < M.limits_cap) atomicAdd(&M.limited_by_limit[fl], 1ull);
            } else {
                atomicAdd(&M.authorized_calls[ns], 1ull);
                atomicAdd(&M.authorized_hits[ns], (unsigned long long)hits);
            }
        }
#endif
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tombstone reclamation.
#define RLM_TOMB_HI 0xFFFFFFFFFFFFFFFFull

// live[g] / tomb[g] += rows of region g that hold a key / a tombstone.  One row per thread; threads past the table
// keep going to the warp-wide step.
__global__ void k_region_census(const uint8_t* __restrict__ rows, uint32_t row_bytes, uint32_t log2R, uint64_t nrows,
                                uint32_t* live, uint32_t* tomb) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = r < nrows;
    bool is_tomb = false, is_live = false;
    uint32_t region = 0;
    if (in) {
        const ulonglong2 hdr = rlm_ld(rows + r * row_bytes);
        is_tomb = hdr.y == RLM_TOMB_HI;
        is_live = !is_tomb && hdr.y != 0;
        region = (uint32_t)(r >> log2R);
    }
#ifndef RL_SHIM
    const unsigned lane = threadIdx.x & 31u;
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, in ? region : 0xFFFFFFFFu);
    const unsigned t = __popc(__ballot_sync(0xFFFFFFFFu, is_tomb) & grp), l = __popc(__ballot_sync(0xFFFFFFFFu, is_live) & grp);
    if (in && (unsigned)(__ffs(grp) - 1) == lane) {
        if (t) atomicAdd(&tomb[region], t);
        if (l) atomicAdd(&live[region], l);
    }
#else
    if (is_tomb) atomicAdd(&tomb[region], 1u);
    if (is_live) atomicAdd(&live[region], 1u);
#endif
}

// Rows of the selected regions -> scratch (same row index), region cleared.
__global__ void k_compact_move(uint8_t* rows, uint8_t* scratch, uint32_t row_bytes, uint32_t log2R, uint64_t nrows,
                               const uint8_t* __restrict__ sel) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows || !sel[r >> log2R]) return;
    uint8_t* src = rows + r * row_bytes;
    uint8_t* dst = scratch + r * row_bytes;
    for (uint32_t q = 0; q < row_bytes; q += 16) {
        const ulonglong2 v = rlm_ld(src + q);
        rlm_st(dst + q, v.x, v.y);
        rlm_st(src + q, 0ull, 0ull);
    }
}

// counts[0] rows inserted again, [1] rows dropped because every cell was (0, 0) (they hold no counter: a qualified
// cell with expiry 0 is absent, an unqualified (0, EPOCH) is the default the next access recreates), [2] failures.
__global__ void k_compact_reinsert(uint8_t* rows, const uint8_t* __restrict__ scratch, uint32_t row_bytes, uint32_t log2P,
                                   uint32_t log2R, uint64_t nrows, const uint8_t* __restrict__ sel,
                                   unsigned long long* counts) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows || !sel[r >> log2R]) return;
    const uint8_t* src = scratch + r * row_bytes;
    const ulonglong2 hdr = rlm_ld(src);
    if (hdr.y == 0 || hdr.y == RLM_TOMB_HI) return;
    bool any = false;
    for (uint32_t q = 16; q < row_bytes; q += 16) {
        const ulonglong2 c = rlm_ld(src + q);
        any = any || c.x != 0 || c.y != 0;
    }
    if (!any) {
        atomicAdd(&counts[1], 1ull);
        return;
    }
    const uint64_t h = rl_row_hash(hdr.x, hdr.y);
    const uint64_t region = log2P ? (h >> (64 - log2P)) : 0ull;
    const uint32_t R = 1u << log2R;
    const uint64_t base = region << log2R;
    const uint32_t idx = (uint32_t)h & (R - 1);
    if (region != (r >> log2R)) {  // the row was not where its hash puts it: never rebuilt wrongly, reported
        atomicAdd(&counts[2], 1ull);
        return;
    }
    for (uint32_t i = 0; i < R; i++) {
        uint8_t* dst = rows + (base + ((idx + i) & (R - 1))) * row_bytes;
        const ulonglong2 cur = rlm_ld(dst);
        if (cur.x != 0 || cur.y != 0) continue;  // taken by another row of the region (keys are distinct)
        const ulonglong2 old = rlm_cas128(dst, make_ulonglong2(0ull, 0ull), hdr);
        if (old.x != 0 || old.y != 0) continue;  // lost the race for this slot: the next one
        for (uint32_t q = 16; q < row_bytes; q += 16) {
            const ulonglong2 c = rlm_ld(src + q);
            rlm_st(dst + q, c.x, c.y);
        }
        atomicAdd(&counts[0], 1ull);
        return;
    }
    atomicAdd(&counts[2], 1ull);  // cannot happen: the region held this row before
}
