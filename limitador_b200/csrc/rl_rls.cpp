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
    std::vector<rl_counter> ctrs;
    std::vector<rl_rls_entry> entries;
    std::vector<char> arena;
    std::vector<rl_binding> binds;
    std::vector<uint8_t> resp;          // finish: this worker's responses, concatenated
    std::vector<uint64_t> resp_len;     // one per request of the worker's range
    std::string error;
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
    std::vector<rl_counter> ctrs;
    std::vector<uint64_t> delta, now;
    int load_counters = 0;
    // engine outputs (serve)
    std::vector<uint8_t> o_limited;
    std::vector<uint32_t> o_first;
    std::vector<uint64_t> o_rem, o_ttl;
    // responses
    std::vector<uint8_t> resp;
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

void plan_range(rl_rls* s, const uint64_t* off, uint32_t w) {
    uint64_t lo, hi;
    range_of(s->n, s->pool->n, w, lo, hi);
    WorkerOut& W = s->wout[w];
    W.ctrs.clear();
    W.error.clear();
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
        uint32_t ns_id;
        if (memchr(msg + q.domain_off, 0, q.domain_len) != nullptr ||
            rl_matcher_namespace_id(s->m, s->domains[i].c_str(), &ns_id) != RL_OK) {
            P.kind = REQ_NO_LIMITS;  // no limit was ever added for the namespace: nothing applies (lib.rs:434-440)
            continue;
        }
        // the CEL context: descriptors[d] = map of the d-th descriptor's entries (server.rs:121-127, 137-139)
        size_t bytes = 0;
        for (uint32_t k = 0; k < sink.n; k++) bytes += (size_t)W.entries[k].key_len + W.entries[k].val_len + 2;
        W.arena.resize(bytes + 1);
        W.binds.resize(sink.n);
        char* a = W.arena.data();
        bool nul = false;
        for (uint32_t k = 0; k < sink.n; k++) {
            const rl_rls_entry& e = W.entries[k];
            nul = nul || memchr(msg + e.key_off, 0, e.key_len) || memchr(msg + e.val_off, 0, e.val_len);
            rl_binding& b = W.binds[k];
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
        }
        if (nul) {  // the matcher compares NUL-terminated strings: an embedded NUL would be cut, not compared
            P.kind = REQ_UNSUPPORTED;
            continue;
        }
        const size_t base = W.ctrs.size();
        W.ctrs.resize(base + RL_MAX_COUNTERS_PER_REQUEST);
        uint32_t got = 0;
        if (rl_matcher_counters(s->m, ns_id, W.binds.data(), sink.n, W.ctrs.data() + base, RL_MAX_COUNTERS_PER_REQUEST, &got) != RL_OK) {
            W.ctrs.resize(base);
            P.kind = REQ_UNSUPPORTED;  // more counters than the engine takes per request
            continue;
        }
        W.ctrs.resize(base + got);
        P.n_ctr = got;
        P.kind = got ? REQ_STORE : REQ_NO_LIMITS;
    }
}

void finish_range(rl_rls* s, int store_status, const uint8_t* limited, const uint64_t* rem, const uint64_t* ttl, uint32_t w) {
    uint64_t lo, hi;
    range_of(s->n, s->pool->n, w, lo, hi);
    WorkerOut& W = s->wout[w];
    W.resp.clear();
    W.resp_len.assign(hi - lo, 0);
    char h_lim[1024], h_rem[32], h_rst[32];
    static const char* const kKeys[3] = {"X-RateLimit-Limit", "X-RateLimit-Remaining", "X-RateLimit-Reset"};  // sorted by key (server.rs:55)
    for (uint64_t i = lo; i < hi; i++) {
        const ReqPlan& P = s->plan[i];
        uint8_t grpc = RL_GRPC_OK, code = RL_RLS_CODE_UNKNOWN;
        uint32_t nh = 0;
        const char* vals[3] = {h_lim, h_rem, h_rst};
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
                if (s->method == RL_RLS_SHOULD_RATE_LIMIT && s->load_counters) {
                    const uint32_t o = s->ctr_off[j];
                    if (rl_matcher_response_headers(s->m, s->ctrs.data() + o, rem + o, ttl + o, P.n_ctr, h_lim, sizeof h_lim,
                                                    h_rem, sizeof h_rem, h_rst, sizeof h_rst) == RL_OK)
                        nh = 3;
                    else
                        grpc = RL_GRPC_UNAVAILABLE;
                }
                break;
            }
            default: grpc = RL_GRPC_INTERNAL; break;
        }
        s->grpc[i] = grpc;
        s->code[i] = grpc == RL_GRPC_OK ? code : 0;
        if (grpc == RL_GRPC_OK) {
            const size_t before = W.resp.size();
            encode_response(W.resp, code, kKeys, vals, nh);
            W.resp_len[i - lo] = W.resp.size() - before;
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
    // lay the workers' counters out as one CSR, requests in batch order
    s->store_index.assign(n, RL_RLS_NO_STORE);
    s->ctr_off.assign(1, 0);
    s->ctrs.clear();
    s->delta.clear();
    uint64_t total = 0;
    for (const auto& W : s->wout) total += W.ctrs.size();
    s->ctrs.reserve(total);
    for (uint32_t w = 0; w < s->pool->n; w++) {
        uint64_t lo, hi;
        range_of(n, s->pool->n, w, lo, hi);
        const WorkerOut& W = s->wout[w];
        size_t at = 0;
        for (uint64_t i = lo; i < hi; i++) {
            ReqPlan& P = s->plan[i];
            if (P.kind != REQ_STORE) continue;
            P.store = (uint32_t)s->delta.size();
            s->store_index[i] = P.store;
            s->ctrs.insert(s->ctrs.end(), W.ctrs.begin() + at, W.ctrs.begin() + at + P.n_ctr);
            at += P.n_ctr;
            s->ctr_off.push_back((uint32_t)s->ctrs.size());
            // CheckRateLimit asks with delta 1 whatever hits_addend says (kuadrant_service.rs:62-65)
            s->delta.push_back(method == RL_RLS_CHECK_RATE_LIMIT ? 1 : P.hits);
        }
    }
    s->n_store = s->delta.size();
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
    s->pool->run([&](uint32_t w) { finish_range(s, store_status, limited, remaining, ttl_us, w); });
    // concatenate the workers' responses
    s->resp.clear();
    s->resp_off.assign(1, 0);
    s->resp_off.reserve(s->n + 1);
    for (uint32_t w = 0; w < s->pool->n; w++) {
        const WorkerOut& W = s->wout[w];
        s->resp.insert(s->resp.end(), W.resp.begin(), W.resp.end());
        for (const uint64_t l : W.resp_len) s->resp_off.push_back(s->resp_off.back() + l);
    }
    // metrics, once per request after the decision (server.rs:183-195, kuadrant_service.rs:81-92,173-174)
    for (uint64_t i = 0; i < s->n; i++) {
        if (s->grpc[i] != RL_GRPC_OK) continue;
        const ReqPlan& P = s->plan[i];
        if (P.kind == REQ_UNKNOWN_DOMAIN) continue;
        NsCounts& c = s->by_ns[s->domains[i]];
        if (s->method == RL_RLS_REPORT) {
            c.authorized_hits += P.hits;
        } else if (s->code[i] == RL_RLS_CODE_OVER_LIMIT) {
            c.limited_calls++;
            if (s->use_limit_name) {
                std::string name;
                const uint32_t lid = first_limited ? first_limited[P.store] : RL_NONE;
                if (lid != RL_NONE) {
                    char nb[512];
                    int has = 0;
                    if (rl_matcher_limit_name_copy(s->m, lid, nb, sizeof nb, &has) == RL_OK && has) name = nb;
                }
                s->limited_by_name[{s->domains[i], name}]++;
            }
        } else {
            c.authorized_calls++;
            if (s->method == RL_RLS_SHOULD_RATE_LIMIT) c.authorized_hits += P.hits;
        }
    }
    s->finished = true;
    return RL_OK;
}

int rl_rls_responses(rl_rls* s, const uint8_t** out_buf, const uint64_t** out_off, const uint8_t** out_grpc, const uint8_t** out_code) {
    if (!s) return RL_FATAL;
    if (!s->finished) return sfail(s, "no finished batch");
    static const uint8_t kEmpty = 0;
    if (out_buf) *out_buf = s->resp.empty() ? &kEmpty : s->resp.data();
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
            s->o_rem.assign(s->ctrs.size(), 0);
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
