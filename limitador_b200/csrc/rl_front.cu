// rl_front.cu — the batching front (SURVEY.md §8b "Threading"): a thread-safe, blocking,
// one-request-at-a-time entry point over the single-threaded batch engine.
//
// The reference's CounterStorage is `Sync + Send` and is called concurrently from tokio /
// actix workers (limitador-server/src/envoy_rls/server.rs:141-158); its only batcher is the
// write-behind one of the cached Redis store (limitador/src/storage/redis/counters_cache.rs:143-247).
// Here concurrent callers enqueue their request and sleep; one dispatcher thread drains the
// queue (up to max_batch requests, waiting at most max_delay_us for more), ships the batch
// through rl_check_and_update_batch and wakes the callers.  The drain order IS the stream order
// that defines the result — a valid linearisation, exactly what concurrent callers of
// InMemoryStorage get — and is returned to each caller as a sequence number.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/rl_engine.h"
#include "../../include/rl_match.h"

namespace {
struct Pending {
    const rl_counter* ctrs;
    uint32_t m;
    uint64_t delta;
    uint64_t now_us;
    int load_counters;
    uint8_t* out_limited;
    uint32_t* out_first;
    uint64_t* out_rem;
    uint64_t* out_ttl;
    uint64_t seq = 0;
    int status = RL_OK;
    bool done = false;
};
}  // namespace

struct rl_front {
    rl_engine* engine = nullptr;
    uint32_t max_batch = 1024;
    uint32_t max_delay_us = 50;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Pending*> queue;
    bool stop = false;
    uint64_t next_seq = 0;
    uint64_t batches = 0, requests = 0;
    std::thread worker;
    // reused host arrays
    std::vector<uint32_t> off;
    std::vector<rl_counter> ctrs;
    std::vector<uint64_t> delta, now, rem, ttl;
    std::vector<uint8_t> limited;
    std::vector<uint32_t> first;
};

static uint64_t wall_us() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(
               std::chrono::system_clock::now().time_since_epoch())
        .count();
}

static void front_loop(rl_front* f) {
    std::vector<Pending*> batch;
    for (;;) {
        batch.clear();
        {
            std::unique_lock<std::mutex> lk(f->mu);
            f->cv_work.wait(lk, [&] { return f->stop || !f->queue.empty(); });
            if (f->stop && f->queue.empty()) return;
            // give concurrent callers a moment to join the batch
            if (f->queue.size() < f->max_batch && f->max_delay_us)
                f->cv_work.wait_for(lk, std::chrono::microseconds(f->max_delay_us),
                                    [&] { return f->stop || f->queue.size() >= f->max_batch; });
            // one engine call has one load_counters flag: cut the batch where it changes
            const int lc = f->queue.front()->load_counters;
            while (!f->queue.empty() && batch.size() < f->max_batch && f->queue.front()->load_counters == lc) {
                batch.push_back(f->queue.front());
                f->queue.pop_front();
            }
            for (Pending* p : batch) p->seq = f->next_seq++;
        }
        const size_t n = batch.size();
        const uint64_t stamp = wall_us();
        f->off.assign(n + 1, 0);
        f->ctrs.clear();
        f->delta.resize(n);
        f->now.resize(n);
        for (size_t i = 0; i < n; i++) {
            const Pending* p = batch[i];
            f->ctrs.insert(f->ctrs.end(), p->ctrs, p->ctrs + p->m);
            f->off[i + 1] = (uint32_t)f->ctrs.size();
            f->delta[i] = p->delta;
            f->now[i] = p->now_us ? p->now_us : stamp;
        }
        const int lc = batch[0]->load_counters;
        f->limited.assign(n, 0);
        f->first.assign(n, RL_NONE);
        f->rem.assign(f->ctrs.size() + 1, 0);
        f->ttl.assign(f->ctrs.size() + 1, 0);
        const int st = rl_check_and_update_batch(f->engine, n, f->off.data(), f->ctrs.data(), f->delta.data(),
                                                 f->now.data(), lc, RL_MEM_HOST, f->limited.data(), f->first.data(),
                                                 lc ? f->rem.data() : nullptr, lc ? f->ttl.data() : nullptr);
        {
            std::lock_guard<std::mutex> lk(f->mu);
            for (size_t i = 0; i < n; i++) {
                Pending* p = batch[i];
                p->status = st;
                if (st == RL_OK) {
                    if (p->out_limited) *p->out_limited = f->limited[i];
                    if (p->out_first) *p->out_first = f->first[i];
                    if (lc && p->out_rem) memcpy(p->out_rem, f->rem.data() + f->off[i], p->m * sizeof(uint64_t));
                    if (lc && p->out_ttl) memcpy(p->out_ttl, f->ttl.data() + f->off[i], p->m * sizeof(uint64_t));
                }
                p->done = true;
            }
            f->batches++;
            f->requests += n;
        }
        f->cv_done.notify_all();
    }
}

extern "C" {

int rl_front_create(rl_engine* e, uint32_t max_batch, uint32_t max_delay_us, rl_front** out) {
    if (!e || !out) return RL_FATAL;
    rl_front* f = new rl_front();
    f->engine = e;
    f->max_batch = max_batch ? max_batch : 1024;
    f->max_delay_us = max_delay_us;
    f->worker = std::thread(front_loop, f);
    *out = f;
    return RL_OK;
}

void rl_front_destroy(rl_front* f) {
    if (!f) return;
    {
        std::lock_guard<std::mutex> lk(f->mu);
        f->stop = true;
    }
    f->cv_work.notify_all();
    if (f->worker.joinable()) f->worker.join();
    delete f;
}

int rl_front_check_and_update(rl_front* f, const rl_counter* ctrs, uint32_t m, uint64_t delta, uint64_t now_us,
                              int load_counters, uint8_t* out_limited, uint32_t* out_first_limited,
                              uint64_t* out_remaining, uint64_t* out_ttl_us, uint64_t* out_seq) {
    if (!f || (!ctrs && m)) return RL_FATAL;
    if (m == 0) {  // lib.rs:434-440 — nothing applies: not limited, no state
        if (out_limited) *out_limited = 0;
        if (out_first_limited) *out_first_limited = RL_NONE;
        return RL_OK;
    }
    Pending p;
    p.ctrs = ctrs;
    p.m = m;
    p.delta = delta;
    p.now_us = now_us;
    p.load_counters = load_counters ? 1 : 0;
    p.out_limited = out_limited;
    p.out_first = out_first_limited;
    p.out_rem = out_remaining;
    p.out_ttl = out_ttl_us;
    std::unique_lock<std::mutex> lk(f->mu);
    if (f->stop) return RL_FATAL;
    f->queue.push_back(&p);
    f->cv_work.notify_one();
    f->cv_done.wait(lk, [&] { return p.done; });
    if (out_seq) *out_seq = p.seq;
    return p.status;
}

int rl_front_check_and_update_bindings(rl_front* f, rl_matcher* m, const char* ns, const rl_binding* binds, uint32_t n_binds,
                                       uint64_t delta, uint64_t now_us, int load_counters, uint8_t* out_limited,
                                       uint32_t* out_first_limited, rl_counter* out_ctrs, uint32_t* out_n_ctrs,
                                       uint64_t* out_remaining, uint64_t* out_ttl_us, uint64_t* out_seq) {
    if (!f || !m || !ns) return RL_FATAL;
    rl_counter local[RL_MAX_COUNTERS_PER_REQUEST];
    rl_counter* ctrs = out_ctrs ? out_ctrs : local;
    uint32_t n = 0, ns_id = 0;
    // counters_that_apply on the caller's thread (lib.rs:507-522); a namespace no limit was ever added for has none
    if (rl_matcher_namespace_id(m, ns, &ns_id) == RL_OK) {
        const int r = rl_matcher_counters(m, ns_id, binds, n_binds, ctrs, RL_MAX_COUNTERS_PER_REQUEST, &n);
        if (r != RL_OK) return r;
    }
    if (out_n_ctrs) *out_n_ctrs = n;
    return rl_front_check_and_update(f, ctrs, n, delta, now_us, load_counters, out_limited, out_first_limited, out_remaining,
                                     out_ttl_us, out_seq);
}

int rl_front_stats(rl_front* f, uint64_t* out_batches, uint64_t* out_requests) {
    if (!f) return RL_FATAL;
    std::lock_guard<std::mutex> lk(f->mu);
    if (out_batches) *out_batches = f->batches;
    if (out_requests) *out_requests = f->requests;
    return RL_OK;
}

}  // extern "C"
