// rl_maint.cu — host side of the maintenance kernels (rl_maint.cuh): tombstone reclamation (rl_compact) and the
// per-namespace metrics reduction (rl_ns_metrics_*).  A translation unit of its own: it sees an engine only through
// rl_internal.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "rl_internal.h"
#include "rl_maint.cuh"

namespace {

struct MaintState {
    int device = 0;
    // metrics accumulators (device): authorized_calls | authorized_hits | limited_calls [ns_cap], limited_by_limit
    // [limits_cap], dropped [1]
    unsigned long long* d_metrics = nullptr;
    uint32_t ns_cap = 0, limits_cap = 0;
    bool metrics_on = false;
};

void maint_free(void* p) {
    MaintState* s = static_cast<MaintState*>(p);
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->d_metrics) cudaFree(s->d_metrics);
    delete s;
}

MaintState* state_of(rl_engine* e, int device) {
    void** slot = rl_internal_ext(e, maint_free);
    if (!*slot) {
        MaintState* s = new MaintState();
        s->device = device;
        *slot = s;
    }
    return static_cast<MaintState*>(*slot);
}

#define RLM_CUDA(e, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t _r = (call);                                                                            \
        if (_r != cudaSuccess) {                                                                            \
            char _b[256];                                                                                   \
            snprintf(_b, sizeof _b, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(_r), __FILE__, __LINE__, \
                     cudaGetErrorString(_r));                                                               \
            return rl_internal_fail((e), _r == cudaErrorMemoryAllocation ? RL_TRANSIENT : RL_FATAL, _b);    \
        }                                                                                                   \
    } while (0)

// A device scratch array freed at scope exit (every exit path of a call, the error ones included).
template <class E>
struct Scratch {
    E* p = nullptr;
    ~Scratch() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, (n ? n : 1) * sizeof(E)); }
};

size_t metrics_words(uint32_t ns_cap, uint32_t limits_cap) { return (size_t)3 * ns_cap + limits_cap + 1; }

RlNsMetricsDev metrics_dev(const MaintState* s) {
    RlNsMetricsDev M;
    M.authorized_calls = s->d_metrics;
    M.authorized_hits = s->d_metrics + s->ns_cap;
    M.limited_calls = s->d_metrics + 2 * (size_t)s->ns_cap;
    M.limited_by_limit = s->d_metrics + 3 * (size_t)s->ns_cap;
    M.dropped = s->d_metrics + 3 * (size_t)s->ns_cap + s->limits_cap;
    M.ns_cap = s->ns_cap;
    M.limits_cap = s->limits_cap;
    return M;
}

// Make the accumulators cover ns_cap namespaces and limits_cap limits; growing keeps the counts (rare: a full
// device synchronisation, then a copy).
int metrics_reserve(rl_engine* e, MaintState* s, uint32_t ns_cap, uint32_t limits_cap) {
    if (s->d_metrics && ns_cap <= s->ns_cap && limits_cap <= s->limits_cap) return RL_OK;
    const uint32_t new_ns = std::max<uint32_t>({ns_cap, s->ns_cap, 1024u}), new_lim = std::max<uint32_t>({limits_cap, s->limits_cap, 1024u});
    const uint32_t grow_ns = s->d_metrics && new_ns > s->ns_cap ? std::max(new_ns, 2 * s->ns_cap) : new_ns;
    const uint32_t grow_lim = s->d_metrics && new_lim > s->limits_cap ? std::max(new_lim, 2 * s->limits_cap) : new_lim;
    if (grow_ns >= (1u << 31)) return rl_internal_fail(e, RL_FATAL, "namespace ids must stay below 2^31 for the metrics reduction");
    unsigned long long* d_new = nullptr;
    RLM_CUDA(e, cudaDeviceSynchronize());
    RLM_CUDA(e, cudaMalloc((void**)&d_new, metrics_words(grow_ns, grow_lim) * sizeof(unsigned long long)));
    RLM_CUDA(e, cudaMemset(d_new, 0, metrics_words(grow_ns, grow_lim) * sizeof(unsigned long long)));
    if (s->d_metrics) {
        const size_t w = sizeof(unsigned long long);
        for (int k = 0; k < 3; k++)
            RLM_CUDA(e, cudaMemcpy(d_new + (size_t)k * grow_ns, s->d_metrics + (size_t)k * s->ns_cap, s->ns_cap * w, cudaMemcpyDeviceToDevice));
        RLM_CUDA(e, cudaMemcpy(d_new + 3 * (size_t)grow_ns, s->d_metrics + 3 * (size_t)s->ns_cap, s->limits_cap * w, cudaMemcpyDeviceToDevice));
        RLM_CUDA(e, cudaMemcpy(d_new + 3 * (size_t)grow_ns + grow_lim, s->d_metrics + 3 * (size_t)s->ns_cap + s->limits_cap, w, cudaMemcpyDeviceToDevice));
        cudaFree(s->d_metrics);
    }
    s->d_metrics = d_new;
    s->ns_cap = grow_ns;
    s->limits_cap = grow_lim;
    return RL_OK;
}

int launch_metrics(rl_engine* e, MaintState* s, cudaStream_t st, uint32_t n, const void* d_recs, int record_bytes,
                   const uint8_t* d_limited, const uint32_t* d_first) {
    if (n == 0) return RL_OK;
    const uint32_t threads = 256;
    const uint32_t blocks = std::min<uint32_t>((n + threads - 1) / threads, 148u * 8u);  // grid-stride beyond 8 CTAs per SM
    k_ns_metrics<<<blocks, threads, 0, st>>>(static_cast<const unsigned long long*>(d_recs), record_bytes == 16 ? 2u : 4u, n,
                                             d_limited, d_first, metrics_dev(s));
    RLM_CUDA(e, cudaGetLastError());
    rl_internal_launched(e, 1);
    return RL_OK;
}

// the hook rl_engine.cu calls behind the replay of a record call (rl_ns_metrics_enable)
int ns_hook(rl_engine* e, cudaStream_t st, uint32_t n, const void* d_recs, int record_bytes, const uint8_t* d_limited,
            const uint32_t* d_first) {
    void** slot = rl_internal_ext(e, maint_free);
    MaintState* s = static_cast<MaintState*>(*slot);
    if (!s || !s->metrics_on) return RL_OK;
    return launch_metrics(e, s, st, n, d_recs, record_bytes, d_limited, d_first);
}

}  // namespace

extern "C" {

int rl_ns_metrics_enable(rl_engine* e, int on) {
    RlTableView v;
    int r = rl_internal_view(e, &v);
    if (r) return r;
    MaintState* s = state_of(e, v.device);
    if (on) {
        r = metrics_reserve(e, s, std::max<uint32_t>(v.ns_cap, 1u << 16), std::max<uint32_t>(v.limits_cap, 1u << 16));
        if (r) return r;
    }
    s->metrics_on = on != 0;
    rl_internal_set_ns_hook(e, on ? ns_hook : nullptr);
    return RL_OK;
}

int rl_ns_metrics_accumulate(rl_engine* e, uint64_t n, const void* recs, uint32_t record_bytes, const uint8_t* limited,
                             const uint32_t* first_limited, int mem) {
    RlTableView v;
    int r = rl_internal_view(e, &v);
    if (r) return r;
    if (record_bytes != 32 && record_bytes != 16) return rl_internal_fail(e, RL_FATAL, "record_bytes must be 32 (rl_record) or 16 (rl_record16)");
    if (n > 0xFFFFFFFFull || (n && (!recs || !limited))) return rl_internal_fail(e, RL_FATAL, "rl_ns_metrics_accumulate: bad arguments");
    MaintState* s = state_of(e, v.device);
    r = metrics_reserve(e, s, std::max<uint32_t>(v.ns_cap, 1u << 16), std::max<uint32_t>(v.limits_cap, 1u << 16));
    if (r || n == 0) return r;
    if (mem == RL_MEM_DEVICE) return launch_metrics(e, s, v.stream, (uint32_t)n, recs, (int)record_bytes, limited, first_limited);
    // host arrays: staged for the call
    Scratch<uint8_t> d_recs, d_lim;
    Scratch<uint32_t> d_first;
    RLM_CUDA(e, d_recs.alloc(n * record_bytes));
    RLM_CUDA(e, d_lim.alloc(n));
    if (first_limited) RLM_CUDA(e, d_first.alloc(n));
    RLM_CUDA(e, cudaMemcpyAsync(d_recs.p, recs, n * record_bytes, cudaMemcpyHostToDevice, v.stream));
    RLM_CUDA(e, cudaMemcpyAsync(d_lim.p, limited, n, cudaMemcpyHostToDevice, v.stream));
    if (first_limited) RLM_CUDA(e, cudaMemcpyAsync(d_first.p, first_limited, n * sizeof(uint32_t), cudaMemcpyHostToDevice, v.stream));
    r = launch_metrics(e, s, v.stream, (uint32_t)n, d_recs.p, (int)record_bytes, d_lim.p, d_first.p);
    RLM_CUDA(e, cudaStreamSynchronize(v.stream));  // before the staged copies are freed
    return r;
}

int rl_ns_metrics_read(rl_engine* e, uint32_t ns_cap, uint64_t* out_authorized_calls, uint64_t* out_authorized_hits,
                       uint64_t* out_limited_calls, uint32_t limits_cap, uint64_t* out_limited_by_limit,
                       uint64_t* out_dropped, int reset) {
    RlTableView v;
    int r = rl_internal_view(e, &v);
    if (r) return r;
    MaintState* s = state_of(e, v.device);
    if (out_dropped) *out_dropped = 0;
    for (uint32_t i = 0; i < ns_cap; i++) {
        if (out_authorized_calls) out_authorized_calls[i] = 0;
        if (out_authorized_hits) out_authorized_hits[i] = 0;
        if (out_limited_calls) out_limited_calls[i] = 0;
    }
    if (out_limited_by_limit)
        for (uint32_t i = 0; i < limits_cap; i++) out_limited_by_limit[i] = 0;
    if (!s->d_metrics) return RL_OK;
    RLM_CUDA(e, cudaStreamSynchronize(v.stream));  // the view fenced the pipeline onto this stream
    std::vector<unsigned long long> h(metrics_words(s->ns_cap, s->limits_cap));
    RLM_CUDA(e, cudaMemcpy(h.data(), s->d_metrics, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const uint32_t nn = std::min(ns_cap, s->ns_cap), nl = std::min(limits_cap, s->limits_cap);
    for (uint32_t i = 0; i < nn; i++) {
        if (out_authorized_calls) out_authorized_calls[i] = h[i];
        if (out_authorized_hits) out_authorized_hits[i] = h[(size_t)s->ns_cap + i];
        if (out_limited_calls) out_limited_calls[i] = h[2 * (size_t)s->ns_cap + i];
    }
    if (out_limited_by_limit)
        for (uint32_t i = 0; i < nl; i++) out_limited_by_limit[i] = h[3 * (size_t)s->ns_cap + i];
    if (out_dropped) *out_dropped = h[3 * (size_t)s->ns_cap + s->limits_cap];
    if (reset) RLM_CUDA(e, cudaMemset(s->d_metrics, 0, h.size() * sizeof(unsigned long long)));
    return RL_OK;
}

int rl_compact(rl_engine* e, uint32_t min_tombstone_pct, rl_compact_stats* out) {
    RlTableView v;
    int r = rl_internal_view(e, &v);
    if (r) return r;
    if (out) memset(out, 0, sizeof *out);
    const uint32_t P = 1u << v.log2P;
    const uint64_t R = 1ull << v.log2R;
    Scratch<uint32_t> d_census;  // live[P] | tomb[P]
    RLM_CUDA(e, d_census.alloc(2 * (size_t)P));
    RLM_CUDA(e, cudaMemsetAsync(d_census.p, 0, 2 * (size_t)P * sizeof(uint32_t), v.stream));
    const uint32_t threads = 256;
    const uint32_t blocks = (uint32_t)((v.capacity + threads - 1) / threads);
    k_region_census<<<blocks, threads, 0, v.stream>>>(v.rows, v.row_bytes, v.log2R, v.capacity, d_census.p, d_census.p + P);
    RLM_CUDA(e, cudaGetLastError());
    rl_internal_launched(e, 1);
    std::vector<uint32_t> census(2 * (size_t)P);
    RLM_CUDA(e, cudaMemcpyAsync(census.data(), d_census.p, census.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, v.stream));
    RLM_CUDA(e, cudaStreamSynchronize(v.stream));
    std::vector<uint8_t> sel(P, 0);
    uint64_t live = 0, tomb = 0, chosen = 0, tomb_chosen = 0;
    for (uint32_t g = 0; g < P; g++) {
        live += census[g];
        tomb += census[P + g];
        // a region is rebuilt when its tombstones reach min_tombstone_pct % of its rows (0 = any tombstone at all)
        if (census[P + g] && (uint64_t)census[P + g] * 100 >= (uint64_t)min_tombstone_pct * R) {
            sel[g] = 1;
            chosen++;
            tomb_chosen += census[P + g];
        }
    }
    if (out) {
        out->regions = P;
        out->rows_live = live;
        out->rows_tombstoned = tomb;
        out->regions_rebuilt = chosen;
    }
    if (!chosen) return RL_OK;
    Scratch<uint8_t> d_sel, d_scratch;
    Scratch<unsigned long long> d_counts;
    RLM_CUDA(e, d_sel.alloc(P));
    if (d_scratch.alloc((size_t)v.capacity * v.row_bytes) != cudaSuccess) {
        cudaGetLastError();
        return rl_internal_fail(e, RL_TRANSIENT, "rl_compact: no device memory for the scratch slab (one copy of the table)");
    }
    RLM_CUDA(e, d_counts.alloc(3));
    RLM_CUDA(e, cudaMemcpyAsync(d_sel.p, sel.data(), P, cudaMemcpyHostToDevice, v.stream));
    RLM_CUDA(e, cudaMemsetAsync(d_counts.p, 0, 3 * sizeof(unsigned long long), v.stream));
    k_compact_move<<<blocks, threads, 0, v.stream>>>(v.rows, d_scratch.p, v.row_bytes, v.log2R, v.capacity, d_sel.p);
    RLM_CUDA(e, cudaGetLastError());
    k_compact_reinsert<<<blocks, threads, 0, v.stream>>>(v.rows, d_scratch.p, v.row_bytes, v.log2P, v.log2R, v.capacity, d_sel.p, d_counts.p);
    RLM_CUDA(e, cudaGetLastError());
    rl_internal_launched(e, 2);
    r = rl_internal_reset_hot_rows(e);  // table row indices changed
    unsigned long long counts[3] = {0, 0, 0};
    RLM_CUDA(e, cudaMemcpyAsync(counts, d_counts.p, sizeof counts, cudaMemcpyDeviceToHost, v.stream));
    RLM_CUDA(e, cudaStreamSynchronize(v.stream));  // before the scratch slab is freed
    if (r) return r;
    if (out) {
        out->rows_moved = counts[0];
        out->rows_reclaimed = tomb_chosen + counts[1];
    }
    if (counts[2]) return rl_internal_fail(e, RL_FATAL, "rl_compact: a row could not be placed again (table corrupt?)");
    return RL_OK;
}

}  // extern "C"
