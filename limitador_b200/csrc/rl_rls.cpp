// rl_rls.cpp — Envoy RLS v3 wire surface: RateLimitRequest bytes -> counters -> ONE engine call -> RateLimitResponse
// bytes (include/rl_rls.h; SURVEY.md §8 f2 with the f1 matcher inside the batching stage).
//
// Host-only code.  The reference serves one request per tonic task: prost decodes the message, a HashMap per
// descriptor is built and bound as `descriptors` (envoy_rls/server.rs:121-139), counters_that_apply walks the CEL
// ASTs, the store is called, the response is built (:183-205).  Here a batch of wire messages is decoded and
// matched by a pool of workers (each request is independent), the counters are laid out as one CSR, the store is
// called once for the whole batch, and the responses are encoded by the same pool.  No protobuf runtime: the four
// message types on the path have a handful of fields, decoded and encoded by hand below.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "rl_rls.h"

namespace {

// ---- protobuf wire format (proto3) -----------------------------------------------------------------------------
struct Rd {
    const uint8_t* p;
    const uint8_t* end;
};

// prost::encoding::decode_varint: at most 10 bytes, the 10th may only carry bit 63
bool rd_varint(Rd& r, uint64_t& v) {
    v = 0;
    for (int i = 0; i < 10; i++) {
        if (r.p >= r.end) return false;
        const uint8_t b = *r.p++;
        if (i == 9 && b > 1) return false;
        v |= (uint64_t)(b & 0x7F) << (7 * i);
        if (!(b & 0x80)) return true;
    }
    return false;
}
bool rd_key(Rd& r, uint32_t& tag, uint32_t& wt) {
    uint64_t k;
    if (!rd_varint(r, k) || k > 0xFFFFFFFFull) return false;  // "invalid key value"
    wt = (uint32_t)k & 7u;
    tag = (uint32_t)k >> 3;
    return tag != 0 && wt <= 5;  // "invalid tag value: 0", "invalid wire type value"
}
bool rd_len(Rd& r, Rd& sub) {
    uint64_t n;
    if (!rd_varint(r, n) || n > (uint64_t)(r.end - r.p)) return false;
    sub.p = r.p;
    sub.end = r.p + n;
    r.p += n;
    return true;
}
bool rd_skip(Rd& r, uint32_t tag, uint32_t wt, int depth) {
    uint64_t v;
    Rd sub;
    switch (wt) {
        case 0: return rd_varint(r, v);
        case 1:
            if (r.end - r.p < 8) return false;
            r.p += 8;
            return true;
        case 2: return rd_len(r, sub);
        case 5:
            if (r.end - r.p < 4) return false;
            r.p += 4;
            return true;
        case 3:  // start group: skip to the matching end group (prost::encoding::skip_field)
            if (depth >= 100) return false;
            for (;;) {
                uint32_t t2, w2;
                if (!rd_key(r, t2, w2)) return false;
                if (w2 == 4) return t2 == tag;
                if (!rd_skip(r, t2, w2, depth + 1)) return false;
            }
        default: return false;  // a stray end group
    }
}

// str::from_utf8: no overlong forms, no surrogates, nothing above U+10FFFF
bool utf8_ok(const uint8_t* p, const uint8_t* end) {
    while (p < end) {
        const uint8_t c = *p;
        if (c < 0x80) {
            p++;
            continue;
        }
        int n;
        uint32_t cp;
        if (c >= 0xC2 && c <= 0xDF) n = 1, cp = c & 0x1F;
        else if (c >= 0xE0 && c <= 0xEF) n = 2, cp = c & 0x0F;
        else if (c >= 0xF0 && c <= 0xF4) n = 3, cp = c & 0x07;
        else return false;
        if (end - p <= n) return false;
        for (int i = 1; i <= n; i++) {
            if ((p[i] & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (p[i] & 0x3F);
        }
        if (n == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
        if (n == 3 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
        p += n + 1;
    }
    return true;
}

bool rd_string(Rd& r, uint32_t wt, const uint8_t* base, uint32_t& off, uint32_t& len) {
    Rd s;
    if (wt != 2 || !rd_len(r, s) || !utf8_ok(s.p, s.end)) return false;
    off = (uint32_t)(s.p - base);
    len = (uint32_t)(s.end - s.p);
    return true;
}

// RateLimitDescriptor.RateLimitOverride {1: uint32, 2: enum}: only validated (the path ignores it)
bool decode_override(Rd r) {
    while (r.p < r.end) {
        uint32_t tag, wt;
        uint64_t v;
        if (!rd_key(r, tag, wt)) return false;
        if (tag == 1 || tag == 2) {
            if (wt != 0 || !rd_varint(r, v)) return false;
        } else if (!rd_skip(r, tag, wt, 0)) {
            return false;
        }
    }
    return true;
}

struct EntrySink {
    rl_rls_entry* out;
    uint32_t cap;
    uint32_t n = 0;
};

bool decode_entry(Rd r, const uint8_t* base, uint32_t descriptor, EntrySink& sink) {
    rl_rls_entry e{descriptor, 0, 0, 0, 0};
    while (r.p < r.end) {
        uint32_t tag, wt;
        if (!rd_key(r, tag, wt)) return false;
        if (tag == 1) {
            if (!rd_string(r, wt, base, e.key_off, e.key_len)) return false;
        } else if (tag == 2) {
            if (!rd_string(r, wt, base, e.val_off, e.val_len)) return false;
        } else if (!rd_skip(r, tag, wt, 0)) {
            return false;
        }
    }
    if (sink.n < sink.cap) sink.out[sink.n] = e;
    sink.n++;
    return true;
}

bool decode_descriptor(Rd r, const uint8_t* base, uint32_t descriptor, EntrySink& sink) {
    while (r.p < r.end) {
        uint32_t tag, wt;
        Rd sub;
        if (!rd_key(r, tag, wt)) return false;
        if (tag == 1) {
            if (wt != 2 || !rd_len(r, sub) || !decode_entry(sub, base, descriptor, sink)) return false;
        } else if (tag == 2) {
            if (wt != 2 || !rd_len(r, sub) || !decode_override(sub)) return false;
        } else if (!rd_skip(r, tag, wt, 0)) {
            return false;
        }
    }
    return true;
}

bool decode_request(const uint8_t* buf, uint64_t len, rl_rls_request& q, EntrySink& sink) {
    if (len > 0xFFFFFFFFull) return false;
    Rd r{buf, buf + len};
    q = rl_rls_request{0, 0, 0, 0, 0};
    while (r.p < r.end) {
        uint32_t tag, wt;
        Rd sub;
        uint64_t v;
        if (!rd_key(r, tag, wt)) return false;
        if (tag == 1) {  // string domain = 1 (a repeated occurrence replaces the earlier one)
            if (!rd_string(r, wt, buf, q.domain_off, q.domain_len)) return false;
        } else if (tag == 2) {  // repeated RateLimitDescriptor descriptors = 2
            if (wt != 2 || !rd_len(r, sub) || !decode_descriptor(sub, buf, q.n_descriptors, sink)) return false;
            q.n_descriptors++;
        } else if (tag == 3) {  // uint32 hits_addend = 3
            if (wt != 0 || !rd_varint(r, v)) return false;
            q.hits_addend = (uint32_t)v;
        } else if (!rd_skip(r, tag, wt, 0)) {
            return false;
        }
    }
    q.n_entries = sink.n;
    return true;
}

// ---- encoding ----------------------------------------------------------------------------------------------------
void put_varint(std::vector<uint8_t>& o, uint64_t v) {
    while (v >= 0x80) {
        o.push_back((uint8_t)(v | 0x80));
        v >>= 7;
    }
    o.push_back((uint8_t)v);
}
size_t varint_size(uint64_t v) {
    size_t n = 1;
    while (v >= 0x80) {
        v >>= 7;
        n++;
    }
    return n;
}
void put_string_field(std::vector<uint8_t>& o, uint32_t tag, const char* s, size_t n) {
    if (n == 0) return;  // proto3: an empty string is not written
    put_varint(o, ((uint64_t)tag << 3) | 2);
    put_varint(o, n);
    o.insert(o.end(), (const uint8_t*)s, (const uint8_t*)s + n);
}
void encode_response(std::vector<uint8_t>& o, uint32_t code, const char* const* keys, const char* const* values, uint32_t nh) {
    if (code) {  // Code overall_code = 1
        o.push_back(0x08);
        put_varint(o, code);
    }
    for (uint32_t h = 0; h < nh; h++) {  // repeated HeaderValue response_headers_to_add = 3 {string key = 1; string value = 2}
        const size_t kn = strlen(keys[h]), vn = strlen(values[h]);
        size_t body = 0;
        if (kn) body += 1 + varint_size(kn) + kn;
        if (vn) body += 1 + varint_size(vn) + vn;
        o.push_back(0x1A);
        put_varint(o, body);
        put_string_field(o, 1, keys[h], kn);
        put_string_field(o, 2, values[h], vn);
    }
}

uint64_t wall_us() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}
double mono_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- a small fork-join pool ----------------------------------------------------------------------------------------
struct Pool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<void(uint32_t)> job;
    uint64_t gen = 0;
    uint32_t pending = 0;
    bool stop = false;
    uint32_t n = 1;

    explicit Pool(uint32_t threads) : n(std::max<uint32_t>(threads, 1)) {
        for (uint32_t w = 1; w < n; w++) workers.emplace_back([this, w] { loop(w); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(mu);
            stop = true;
        }
        cv_go.notify_all();
        for (auto& t : workers) t.join();
    }
    void loop(uint32_t w) {
        uint64_t seen = 0;
        for (;;) {
            std::function<void(uint32_t)> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                f = job;
            }
            f(w);
            {
                std::lock_guard<std::mutex> g(mu);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // f(w) runs once for every w in [0, n): worker 0 is the calling thread
    void run(const std::function<void(uint32_t)>& f) {
        if (n == 1) {
            f(0);
            return;
        }
        {
            std::lock_guard<std::mutex> g(mu);
            job = f;
            pending = n - 1;
            gen++;
        }
        cv_go.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

// A per-batch buffer that is NOT value-initialised when it is sized (std::vector::resize would zero megabytes per batch);
// its contents do not survive ensure().
template <class T>
struct RawBuf {
    std::unique_ptr<T[]> p;
    size_t cap = 0, n = 0;
    void ensure(size_t want) {
        if (want > cap) {
            cap = want + want / 2 + 16;
            p.reset(new T[cap]);
        }
        n = want;
    }
    T* data() { return p.get(); }
    const T* data() const { return p.get(); }
    size_t size() const { return n; }
};

struct NsCounts {
    uint64_t authorized_calls = 0, authorized_hits = 0, limited_calls = 0;
};

enum : uint8_t { REQ_BAD_WIRE = 1, REQ_UNKNOWN_DOMAIN = 2, REQ_NO_LIMITS = 3, REQ_STORE = 4, REQ_UNSUPPORTED = 5 };

struct ReqPlan {
    uint8_t kind = 0;
    uint32_t hits = 1;
    uint32_t n_ctr = 0;
    uint32_t store = RL_RLS_NO_STORE;
};

struct WorkerOut {
    // plan: the requests of the range that go to the matcher, as rl_matcher_counters_batch_ns takes them
    RawBuf<rl_counter> ctrs;            // the range's counters, request after request (ctr_off)
    std::vector<rl_rls_entry> entries;
    RawBuf<char> arena;                 // NUL-terminated copies of the keys and values (the bindings point into it)
    uint64_t n_store = 0;               // requests of the range that reach the store
    std::vector<rl_binding> binds;
    std::vector<uint32_t> bind_off, ctr_off;
    std::vector<const char*> ns;
    std::vector<uint64_t> req_of;       // matcher request k = batch request req_of[k]
    std::vector<uint8_t> status;
    // finish
    std::vector<uint8_t> resp;          // this worker's responses, concatenated
    std::vector<uint64_t> resp_len;     // one per request of the worker's range
    std::vector<char> hdr;              // header values of the range's store requests (rl_matcher_response_headers_batch)
    std::vector<uint64_t> hdr_off;
    std::unordered_map<std::string, NsCounts> by_ns;  // the range's metrics, merged into the service's after the workers
    std::map<std::pair<std::string, std::string>, uint64_t> limited_by_name;
};

}  // namespace

struct rl_rls {
    rl_matcher* m = nullptr;
    rl_engine* engine = nullptr;
    int header_mode = RL_RLS_HEADERS_NONE;
    bool use_limit_name = false;
    Pool* pool = nullptr;
    std::string last_error;

    // the batch
    int method = 0;
    uint64_t n = 0;
    const uint8_t* buf = nullptr;  // only dereferenced during plan
    bool planned = false, finished = false;
    std::vector<ReqPlan> plan;
    std::vector<std::string> domains;  // copy of every request's domain (the metrics outlive the input buffer)
    std::vector<WorkerOut> wout;
    std::vector<uint32_t> store_index;
    // store call
    uint64_t n_store = 0;
    std::vector<uint32_t> ctr_off;
    RawBuf<rl_counter> ctrs;
    std::vector<uint64_t> delta, now;
    int load_counters = 0;
    // engine outputs (serve)
    std::vector<uint8_t> o_limited;
    std::vector<uint32_t> o_first;
    std::vector<uint64_t> o_rem, o_ttl;
    // responses
    RawBuf<uint8_t> resp;
    std::vector<uint64_t> resp_off;
    std::vector<uint8_t> grpc, code;
    // metrics
    std::map<std::string, NsCounts> by_ns;
    std::map<std::pair<std::string, std::string>, uint64_t> limited_by_name;
    double t_plan = 0, t_store = 0, t_finish = 0;
};

namespace {

int sfail(rl_rls* s, const char* fmt, ...) {
    char b[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(b, sizeof b, fmt, ap);
    va_end(ap);
    s->last_error = b;
    return RL_FATAL;
}

// requests [lo, hi) of worker w
void range_of(uint64_t n, uint32_t workers, uint32_t w, uint64_t& lo, uint64_t& hi) {
    const uint64_t per = (n + workers - 1) / workers;
    lo = std::min<uint64_t>(n, (uint64_t)w * per);
    hi = std::min<uint64_t>(n, lo + per);
}

// Requests [lo, hi) of worker w: decode every message and lay its CEL context out (pass 1), then ONE matcher call for
// the whole range (pass 2) — a reader section per range, not per request: eight workers taking the matcher's
// reader/writer lock twice per request spent their time passing its cache line around and did not scale at all.
void plan_range(rl_rls* s, const uint64_t* off, uint32_t w) {
    uint64_t lo, hi;
    range_of(s->n, s->pool->n, w, lo, hi);
    WorkerOut& W = s->wout[w];
    W.n_store = 0;
    W.binds.clear();
    W.bind_off.assign(1, 0);
    W.ns.clear();
    W.req_of.clear();
    // every string gets a NUL behind it: an entry is at least two bytes on the wire, so twice the range's bytes is enough
    // — sized up front, because the bindings point into it
    W.arena.ensure(2 * (size_t)(off[hi] - off[lo]) + 16);
    char* a = W.arena.data();
    for (uint64_t i = lo; i < hi; i++) {
        ReqPlan& P = s->plan[i];
        P = ReqPlan();
        const uint8_t* msg = s->buf + off[i];
        const uint64_t len = off[i + 1] - off[i];
        rl_rls_request q;
        if (W.entries.size() < 16) W.entries.resize(16);
        EntrySink sink{W.entries.data(), (uint32_t)W.entries.size()};
        if (!decode_request(msg, len, q, sink)) {
            P.kind = REQ_BAD_WIRE;
            continue;
        }
        if (sink.n > sink.cap) {  // rare: more entries than the scratch holds — decode again into a larger one
            W.entries.resize(sink.n);
            sink = EntrySink{W.entries.data(), (uint32_t)W.entries.size()};
            decode_request(msg, len, q, sink);
        }
        P.hits = q.hits_addend ? q.hits_addend : 1;  // server.rs:131-135
        s->domains[i].assign((const char*)msg + q.domain_off, q.domain_len);
        if (q.domain_len == 0) {  // server.rs:106-116
            P.kind = REQ_UNKNOWN_DOMAIN;
            continue;
        }
        if (memchr(msg + q.domain_off, 0, q.domain_len) != nullptr) {
            P.kind = REQ_NO_LIMITS;  // no namespace the matcher knows holds a NUL: nothing applies (lib.rs:434-440)
            continue;
        }
        // the CEL context: descriptors[d] = map of the d-th descriptor's entries (server.rs:121-127, 137-139)
        bool nul = false;
        const size_t first_bind = W.binds.size();
        for (uint32_t k = 0; k < sink.n; k++) {
            const rl_rls_entry& e = W.entries[k];
            nul = nul || memchr(msg + e.key_off, 0, e.key_len) || memchr(msg + e.val_off, 0, e.val_len);
            rl_binding b;
            b.descriptor = e.descriptor;
            b._pad = 0;
            b.key = a;
            memcpy(a, msg + e.key_off, e.key_len);
            a[e.key_len] = 0;
            a += e.key_len + 1;
            b.value = a;
            memcpy(a, msg + e.val_off, e.val_len);
            a[e.val_len] = 0;
            a += e.val_len + 1;
            W.binds.push_back(b);
        }
        if (nul) {  // the matcher compares NUL-terminated strings: an embedded NUL would be cut, not compared
            W.binds.resize(first_bind);
            P.kind = REQ_UNSUPPORTED;
            continue;
        }
        W.bind_off.push_back((uint32_t)W.binds.size());
        W.ns.push_back(s->domains[i].c_str());  // (s->domains was sized before the workers started: the pointer stays)
        W.req_of.push_back(i);
    }
    const uint64_t k_req = W.req_of.size();
    W.ctr_off.assign(k_req + 1, 0);
    W.status.assign(k_req, 0);
    W.ctrs.ensure((k_req + 1) * (size_t)RL_MAX_COUNTERS_PER_REQUEST);
    if (rl_matcher_counters_batch_ns(s->m, k_req, W.ns.data(), W.bind_off.data(), W.binds.data(), W.ctr_off.data(), W.ctrs.data(),
                                     W.ctrs.size(), W.status.data()) != RL_OK) {
        for (const uint64_t i : W.req_of) s->plan[i].kind = REQ_UNSUPPORTED;  // (the matcher's cap was raised past the engine's)
        W.ctr_off.assign(k_req + 1, 0);
        return;
    }
    for (uint64_t k = 0; k < k_req; k++) {
        ReqPlan& P = s->plan[W.req_of[k]];
        P.n_ctr = W.ctr_off[k + 1] - W.ctr_off[k];
        // 2 = more counters than the engine takes per request; 1 = a namespace without limits (lib.rs:434-440)
        P.kind = W.status[k] == 2 ? REQ_UNSUPPORTED : (P.n_ctr ? REQ_STORE : REQ_NO_LIMITS);
        W.n_store += P.kind == REQ_STORE;
    }
}

// Second pass of the plan: worker w copies its counters into the batch's CSR at the offsets the prefix over the workers
// gave it (store requests in batch order: worker ranges are consecutive).
void plan_scatter(rl_rls* s, uint32_t w, uint64_t store_base, uint64_t ctr_base) {
    WorkerOut& W = s->wout[w];
    uint64_t j = store_base, c = ctr_base;
    for (uint64_t k = 0; k < W.req_of.size(); k++) {
        const uint64_t i = W.req_of[k];
        ReqPlan& P = s->plan[i];
        if (P.kind != REQ_STORE) continue;
        P.store = (uint32_t)j;
        s->store_index[i] = (uint32_t)j;
        s->ctr_off[j] = (uint32_t)c;
        memcpy(s->ctrs.data() + c, W.ctrs.data() + W.ctr_off[k], (size_t)P.n_ctr * sizeof(rl_counter));
        // CheckRateLimit asks with delta 1 whatever hits_addend says (kuadrant_service.rs:62-65)
        s->delta[j] = s->method == RL_RLS_CHECK_RATE_LIMIT ? 1 : P.hits;
        c += P.n_ctr;
        j++;
    }
}

// Second pass of the finish: worker w copies its responses behind those of the workers before it.
void finish_scatter(rl_rls* s, uint32_t w, uint64_t byte_base) {
    uint64_t lo, hi;
    range_of(s->n, s->pool->n, w, lo, hi);
    const WorkerOut& W = s->wout[w];
    if (!W.resp.empty()) memcpy(s->resp.data() + byte_base, W.resp.data(), W.resp.size());
    uint64_t at = byte_base;
    for (uint64_t i = lo; i < hi; i++) {
        at += W.resp_len[i - lo];
        s->resp_off[i + 1] = at;
    }
}

void finish_range(rl_rls* s, int store_status, const uint8_t* limited, const uint32_t* first, const uint64_t* rem, const uint64_t* ttl,
                  uint32_t w) {
    uint64_t lo, hi;
    range_of(s->n, s->pool->n, w, lo, hi);
    WorkerOut& W = s->wout[w];
    W.resp.clear();
    W.resp_len.assign(hi - lo, 0);
    W.by_ns.clear();
    W.limited_by_name.clear();
    static const char* const kKeys[3] = {"X-RateLimit-Limit", "X-RateLimit-Remaining", "X-RateLimit-Reset"};  // sorted by key (server.rs:55)
    // the store requests of the range are one run of store indices: their header values in ONE matcher call
    const bool with_headers = store_status == RL_OK && s->method == RL_RLS_SHOULD_RATE_LIMIT && s->load_counters;
    uint64_t j0 = RL_RLS_NO_STORE, j1 = 0;
    for (uint64_t i = lo; i < hi; i++)
        if (s->plan[i].kind == REQ_STORE) {
            if (j0 == RL_RLS_NO_STORE) j0 = s->plan[i].store;
            j1 = (uint64_t)s->plan[i].store + 1;
        }
    bool headers_ok = false;
    if (with_headers && j0 != RL_RLS_NO_STORE) {
        W.hdr_off.assign(j1 - j0 + 1, 0);
        uint64_t need = 0;
        W.hdr.resize(std::max<size_t>(W.hdr.size(), (size_t)(j1 - j0) * 96));
        int r = rl_matcher_response_headers_batch(s->m, j1 - j0, s->ctr_off.data() + j0, s->ctrs.data(), rem, ttl, W.hdr.data(), W.hdr.size(),
                                                  W.hdr_off.data(), &need);
        if (r != RL_OK && need > W.hdr.size()) {
            W.hdr.resize(need);
            r = rl_matcher_response_headers_batch(s->m, j1 - j0, s->ctr_off.data() + j0, s->ctrs.data(), rem, ttl, W.hdr.data(), W.hdr.size(),
                                                  W.hdr_off.data(), &need);
        }
        headers_ok = r == RL_OK;
    }
    const std::string* last_ns = nullptr;  // consecutive requests of one namespace share the metrics entry
    NsCounts* last_counts = nullptr;
    for (uint64_t i = lo; i < hi; i++) {
        const ReqPlan& P = s->plan[i];
        uint8_t grpc = RL_GRPC_OK, code = RL_RLS_CODE_UNKNOWN;
        uint32_t nh = 0;
        const char* vals[3] = {"", "", ""};
        switch (P.kind) {
            case REQ_BAD_WIRE: grpc = RL_GRPC_INTERNAL; break;
            case REQ_UNSUPPORTED: grpc = RL_GRPC_UNAVAILABLE; break;
            case REQ_UNKNOWN_DOMAIN: code = RL_RLS_CODE_UNKNOWN; break;
            case REQ_NO_LIMITS: code = RL_RLS_CODE_OK; break;
            case REQ_STORE: {
                if (store_status != RL_OK) {  // server.rs:160-172
                    grpc = RL_GRPC_UNAVAILABLE;
                    break;
                }
                const uint32_t j = P.store;
                if (s->method == RL_RLS_REPORT) {
                    code = RL_RLS_CODE_OK;  // kuadrant_service.rs:176-178
                    break;
                }
                if (limited[j] == RL_VERDICT_ERROR) {
                    grpc = RL_GRPC_UNAVAILABLE;
                    break;
                }
                code = limited[j] ? RL_RLS_CODE_OVER_LIMIT : RL_RLS_CODE_OK;
                if (with_headers) {
                    if (!headers_ok) {
                        grpc = RL_GRPC_UNAVAILABLE;
                        break;
                    }
                    vals[0] = W.hdr.data() + W.hdr_off[j - j0];
                    vals[1] = vals[0] + strlen(vals[0]) + 1;
                    vals[2] = vals[1] + strlen(vals[1]) + 1;
                    nh = 3;
                }
                break;
            }
            default: grpc = RL_GRPC_INTERNAL; break;
        }
        s->grpc[i] = grpc;
        s->code[i] = grpc == RL_GRPC_OK ? code : 0;
        if (grpc != RL_GRPC_OK) continue;
        const size_t before = W.resp.size();
        encode_response(W.resp, code, kKeys, vals, nh);
        W.resp_len[i - lo] = W.resp.size() - before;
        // metrics, once per request after the decision (server.rs:183-195, kuadrant_service.rs:81-92,173-174), into the
        // worker's own table (merged after the workers are done)
        if (P.kind == REQ_UNKNOWN_DOMAIN) continue;
        if (!last_ns || *last_ns != s->domains[i]) {
            last_ns = &s->domains[i];
            last_counts = &W.by_ns[s->domains[i]];
        }
        NsCounts& c = *last_counts;
        if (s->method == RL_RLS_REPORT) {
            c.authorized_hits += P.hits;
        } else if (code == RL_RLS_CODE_OVER_LIMIT) {
            c.limited_calls++;
            if (s->use_limit_name) {
                std::string name;
                const uint32_t lid = first ? first[P.store] : RL_NONE;
                if (lid != RL_NONE) {
                    char nb[512];
                    int has = 0;
                    if (rl_matcher_limit_name_copy(s->m, lid, nb, sizeof nb, &has) == RL_OK && has) name = nb;
                }
                W.limited_by_name[{s->domains[i], name}]++;
            }
        } else {
            c.authorized_calls++;
            if (s->method == RL_RLS_SHOULD_RATE_LIMIT) c.authorized_hits += P.hits;
        }
    }
}

}  // namespace

extern "C" {

int rl_rls_decode_request(const uint8_t* buf, uint64_t len, rl_rls_request* out, rl_rls_entry* entries, uint32_t cap_entries) {
    if (!out || (len && !buf) || (cap_entries && !entries)) return RL_FATAL;
    EntrySink sink{entries, cap_entries};
    return decode_request(buf, len, *out, sink) ? RL_OK : RL_FATAL;
}

int rl_rls_encode_response(uint32_t overall_code, const char* const* keys, const char* const* values, uint32_t n_headers,
                           uint8_t* out, uint64_t cap, uint64_t* out_len) {
    if (!out_len || (n_headers && (!keys || !values))) return RL_FATAL;
    std::vector<uint8_t> o;
    encode_response(o, overall_code, keys, values, n_headers);
    *out_len = o.size();
    if (o.size() > cap || (!out && !o.empty())) return RL_FATAL;
    if (!o.empty()) memcpy(out, o.data(), o.size());
    return RL_OK;
}

int rl_rls_create(rl_matcher* m, rl_engine* engine, int header_mode, uint32_t threads, int use_limit_name_label, rl_rls** out) {
    if (!m || !out) return RL_FATAL;
    if (header_mode != RL_RLS_HEADERS_NONE && header_mode != RL_RLS_HEADERS_DRAFT_VERSION_03) return RL_FATAL;
    rl_rls* s = new rl_rls();
    s->m = m;
    s->engine = engine;
    s->header_mode = header_mode;
    s->use_limit_name = use_limit_name_label != 0;
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    threads = std::min<uint32_t>(threads, 64);
    s->pool = new Pool(threads);
    s->wout.resize(threads);
    *out = s;
    return RL_OK;
}

void rl_rls_destroy(rl_rls* s) {
    if (!s) return;
    delete s->pool;
    delete s;
}

const char* rl_rls_last_error(rl_rls* s) { return s ? s->last_error.c_str() : "null service"; }

int rl_rls_plan(rl_rls* s, int method, uint64_t n, const uint8_t* buf, const uint64_t* off, uint64_t now_us) {
    if (!s || (n && (!off || !buf))) return RL_FATAL;
    if (method != RL_RLS_SHOULD_RATE_LIMIT && method != RL_RLS_CHECK_RATE_LIMIT && method != RL_RLS_REPORT)
        return sfail(s, "unknown method %d", method);
    for (uint64_t i = 0; i < n; i++)
        if (off[i + 1] < off[i]) return sfail(s, "request offsets must be non-decreasing (request %llu)", (unsigned long long)i);
    s->planned = s->finished = false;
    s->method = method;
    s->n = n;
    s->buf = buf;
    s->plan.resize(n);
    s->domains.resize(n);
    s->pool->run([&](uint32_t w) { plan_range(s, off, w); });
    s->buf = nullptr;
    // lay the workers' counters out as one CSR, requests in batch order: a prefix over the workers, then every worker
    // copies its own slice
    std::vector<uint64_t> store_base(s->pool->n + 1, 0), ctr_base(s->pool->n + 1, 0);
    for (uint32_t w = 0; w < s->pool->n; w++) {
        const WorkerOut& W = s->wout[w];
        store_base[w + 1] = store_base[w] + W.n_store;
        ctr_base[w + 1] = ctr_base[w] + (W.ctr_off.empty() ? 0 : W.ctr_off.back());
    }
    const uint64_t n_store = store_base[s->pool->n], n_ctr = ctr_base[s->pool->n];
    if (n_ctr > 0xFFFFFFFFull) return sfail(s, "more than 2^32 counters in one batch");
    s->store_index.assign(n, RL_RLS_NO_STORE);
    s->ctr_off.assign(n_store + 1, 0);
    s->ctr_off[n_store] = (uint32_t)n_ctr;
    s->ctrs.ensure(n_ctr);
    s->delta.assign(n_store, 0);
    s->pool->run([&](uint32_t w) { plan_scatter(s, w, store_base[w], ctr_base[w]); });
    s->n_store = n_store;
    s->now.assign(s->n_store, now_us ? now_us : wall_us());
    s->load_counters = (method == RL_RLS_SHOULD_RATE_LIMIT && s->header_mode != RL_RLS_HEADERS_NONE) ? 1 : 0;  // server.rs:146
    s->planned = true;
    return RL_OK;
}

int rl_rls_plan_view(rl_rls* s, uint64_t* out_n_store, const uint32_t** out_ctr_off, const rl_counter** out_ctrs,
                     const uint64_t** out_delta, const uint64_t** out_now_us, int* out_load_counters,
                     const uint32_t** out_store_index) {
    if (!s) return RL_FATAL;
    if (!s->planned) return sfail(s, "no planned batch");
    if (out_n_store) *out_n_store = s->n_store;
    if (out_ctr_off) *out_ctr_off = s->ctr_off.data();
    if (out_ctrs) *out_ctrs = s->ctrs.data();
    if (out_delta) *out_delta = s->delta.data();
    if (out_now_us) *out_now_us = s->now.data();
    if (out_load_counters) *out_load_counters = s->load_counters;
    if (out_store_index) *out_store_index = s->store_index.data();
    return RL_OK;
}

int rl_rls_finish(rl_rls* s, int store_status, const uint8_t* limited, const uint32_t* first_limited,
                  const uint64_t* remaining, const uint64_t* ttl_us) {
    if (!s) return RL_FATAL;
    if (!s->planned) return sfail(s, "no planned batch");
    const bool need_verdicts = s->n_store && store_status == RL_OK && s->method != RL_RLS_REPORT;
    if (need_verdicts && !limited) return sfail(s, "finish needs the verdicts of the store call");
    if (need_verdicts && s->load_counters && (!remaining || !ttl_us))
        return sfail(s, "finish needs remaining / ttl of the store call (draft-03 headers)");
    s->grpc.assign(s->n, 0);
    s->code.assign(s->n, 0);
    s->pool->run([&](uint32_t w) { finish_range(s, store_status, limited, first_limited, remaining, ttl_us, w); });
    // the workers' responses behind one another (a prefix over the workers, then every worker copies its own), their
    // metrics merged
    std::vector<uint64_t> byte_base(s->pool->n + 1, 0);
    for (uint32_t w = 0; w < s->pool->n; w++) byte_base[w + 1] = byte_base[w] + s->wout[w].resp.size();
    s->resp.ensure(byte_base[s->pool->n]);
    s->resp_off.assign(s->n + 1, 0);
    s->pool->run([&](uint32_t w) { finish_scatter(s, w, byte_base[w]); });
    for (uint32_t w = 0; w < s->pool->n; w++) {
        const WorkerOut& W = s->wout[w];
        for (const auto& kv : W.by_ns) {
            NsCounts& c = s->by_ns[kv.first];
            c.authorized_calls += kv.second.authorized_calls;
            c.authorized_hits += kv.second.authorized_hits;
            c.limited_calls += kv.second.limited_calls;
        }
        for (const auto& kv : W.limited_by_name) s->limited_by_name[kv.first] += kv.second;
    }
    s->finished = true;
    s->planned = false;  // a batch is finished once: its metrics are counted once
    return RL_OK;
}

int rl_rls_responses(rl_rls* s, const uint8_t** out_buf, const uint64_t** out_off, const uint8_t** out_grpc, const uint8_t** out_code) {
    if (!s) return RL_FATAL;
    if (!s->finished) return sfail(s, "no finished batch");
    static const uint8_t kEmpty = 0;
    if (out_buf) *out_buf = s->resp.size() ? s->resp.data() : &kEmpty;
    if (out_off) *out_off = s->resp_off.data();
    if (out_grpc) *out_grpc = s->grpc.empty() ? &kEmpty : s->grpc.data();
    if (out_code) *out_code = s->code.empty() ? &kEmpty : s->code.data();
    return RL_OK;
}

int rl_rls_serve(rl_rls* s, int method, uint64_t n, const uint8_t* buf, const uint64_t* off, uint64_t now_us) {
    if (!s) return RL_FATAL;
    if (!s->engine) return sfail(s, "the service was created without an engine: there is no CPU store to fall back to");
    const double t0 = mono_us();
    int r = rl_rls_plan(s, method, n, buf, off, now_us);
    if (r) return r;
    const double t1 = mono_us();
    int st = RL_OK;
    if (s->n_store) {
        const uint64_t m = s->n_store;
        s->o_limited.assign(m, 0);
        s->o_first.assign(m, RL_NONE);
        if (s->load_counters) {
            s->o_rem.assign(s->ctrs.size(), 0);  // (slots of a refused call stay 0)
            s->o_ttl.assign(s->ctrs.size(), 0);
        }
        if (method == RL_RLS_SHOULD_RATE_LIMIT)
            st = rl_check_and_update_batch(s->engine, m, s->ctr_off.data(), s->ctrs.data(), s->delta.data(), s->now.data(),
                                           s->load_counters, RL_MEM_HOST, s->o_limited.data(), s->o_first.data(),
                                           s->load_counters ? s->o_rem.data() : nullptr, s->load_counters ? s->o_ttl.data() : nullptr);
        else if (method == RL_RLS_CHECK_RATE_LIMIT)
            st = rl_is_within_limits_batch(s->engine, m, s->ctr_off.data(), s->ctrs.data(), s->delta.data(), s->now.data(),
                                           RL_MEM_HOST, s->o_limited.data(), s->o_first.data());
        else
            st = rl_update_batch(s->engine, m, s->ctr_off.data(), s->ctrs.data(), s->delta.data(), s->now.data(), RL_MEM_HOST);
        if (st != RL_OK) s->last_error = std::string("store call failed: ") + rl_last_error(s->engine);
    }
    const double t2 = mono_us();
    r = rl_rls_finish(s, st, s->o_limited.data(), s->o_first.data(), s->o_rem.data(), s->o_ttl.data());
    const double t3 = mono_us();
    s->t_plan = t1 - t0;
    s->t_store = t2 - t1;
    s->t_finish = t3 - t2;
    return r;
}

int rl_rls_metrics_render(rl_rls* s, char* out, uint64_t cap, uint64_t* out_len) {
    if (!s || !out_len) return RL_FATAL;
    std::string t;
    auto esc = [](const std::string& v) {  // label values: backslash, quote and newline are escaped
        std::string o;
        for (const char c : v) {
            if (c == '\\') o += "\\\\";
            else if (c == '"') o += "\\\"";
            else if (c == '\n') o += "\\n";
            else o.push_back(c);
        }
        return o;
    };
    t += "# TYPE authorized_calls counter\n";
    for (const auto& kv : s->by_ns)
        if (kv.second.authorized_calls)
            t += "authorized_calls{limitador_namespace=\"" + esc(kv.first) + "\"} " + std::to_string(kv.second.authorized_calls) + "\n";
    t += "# TYPE authorized_hits counter\n";
    for (const auto& kv : s->by_ns)
        if (kv.second.authorized_hits)
            t += "authorized_hits{limitador_namespace=\"" + esc(kv.first) + "\"} " + std::to_string(kv.second.authorized_hits) + "\n";
    t += "# TYPE limited_calls counter\n";
    if (s->use_limit_name) {
        for (const auto& kv : s->limited_by_name)
            t += "limited_calls{limitador_namespace=\"" + esc(kv.first.first) + "\",limit_name=\"" + esc(kv.first.second) + "\"} " +
                 std::to_string(kv.second) + "\n";
    } else {
        for (const auto& kv : s->by_ns)
            if (kv.second.limited_calls)
                t += "limited_calls{limitador_namespace=\"" + esc(kv.first) + "\"} " + std::to_string(kv.second.limited_calls) + "\n";
    }
    t += "# TYPE limitador_up gauge\nlimitador_up 1\n";
    *out_len = t.size() + 1;
    if (!out || cap < t.size() + 1) return RL_FATAL;
    memcpy(out, t.c_str(), t.size() + 1);
    return RL_OK;
}

int rl_rls_last_timings(rl_rls* s, double* out_plan_us, double* out_store_us, double* out_finish_us) {
    if (!s) return RL_FATAL;
    if (out_plan_us) *out_plan_us = s->t_plan;
    if (out_store_us) *out_store_us = s->t_store;
    if (out_finish_us) *out_finish_us = s->t_finish;
    return RL_OK;
}

}  // extern "C"
