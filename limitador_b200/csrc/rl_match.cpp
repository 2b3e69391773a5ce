// rl_match.cpp — CPU front: limits -> counters (include/rl_match.h).
//
// Host-only code.  The reference walks a CEL AST per condition and per variable for every limit of the
// namespace (limit.rs:157-174, cel.rs:185-227,314-334), building string maps on the way; here every
// operand of every accepted expression is interned into a SLOT when the limit is added, a request's
// bindings are dropped into the slot array once, and a limit is a list of (slot, ==|!=, literal) tests plus
// a list of slots that must be present.  The counter key of a variable set is digested once per request
// and shared by the limits that use the same variables.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "rl_match.h"

namespace {

// ---- BLAKE2b (RFC 7693), unkeyed, streaming ----------------------------------------------
struct Blake2b {
    uint64_t h[8];
    uint64_t t = 0;
    uint8_t buf[128];
    size_t fill = 0;
    size_t outlen;

    static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    static inline uint64_t load64(const uint8_t* p) {
        uint64_t v;
        memcpy(&v, p, 8);  // little-endian hosts only (x86-64, aarch64)
        return v;
    }
    explicit Blake2b(size_t out) : outlen(out) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                       0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                       0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        memcpy(h, IV, sizeof h);
        h[0] ^= 0x01010000ULL ^ (uint64_t)out;
    }
    void compress(const uint8_t* block, bool last) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                       0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                       0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        static const uint8_t S[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        for (int i = 0; i < 16; i++) m[i] = load64(block + 8 * i);
        for (int i = 0; i < 8; i++) {
            v[i] = h[i];
            v[i + 8] = IV[i];
        }
        v[12] ^= t;  // message lengths stay far below 2^64: the high counter word is 0
        if (last) v[14] = ~v[14];
#define RL_B2_G(a, b, c, d, x, y)          \
    v[a] = v[a] + v[b] + (x);              \
    v[d] = rotr(v[d] ^ v[a], 32);          \
    v[c] = v[c] + v[d];                    \
    v[b] = rotr(v[b] ^ v[c], 24);          \
    v[a] = v[a] + v[b] + (y);              \
    v[d] = rotr(v[d] ^ v[a], 16);          \
    v[c] = v[c] + v[d];                    \
    v[b] = rotr(v[b] ^ v[c], 63);
        for (int r = 0; r < 12; r++) {
            const uint8_t* s = S[r];
            RL_B2_G(0, 4, 8, 12, m[s[0]], m[s[1]])
            RL_B2_G(1, 5, 9, 13, m[s[2]], m[s[3]])
            RL_B2_G(2, 6, 10, 14, m[s[4]], m[s[5]])
            RL_B2_G(3, 7, 11, 15, m[s[6]], m[s[7]])
            RL_B2_G(0, 5, 10, 15, m[s[8]], m[s[9]])
            RL_B2_G(1, 6, 11, 12, m[s[10]], m[s[11]])
            RL_B2_G(2, 7, 8, 13, m[s[12]], m[s[13]])
            RL_B2_G(3, 4, 9, 14, m[s[14]], m[s[15]])
        }
#undef RL_B2_G
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const void* data, size_t n) {
        const uint8_t* p = (const uint8_t*)data;
        while (n) {
            if (fill == 128) {  // a full buffer is only compressed once more input follows it
                t += 128;
                compress(buf, false);
                fill = 0;
            }
            const size_t k = std::min(n, (size_t)128 - fill);
            memcpy(buf + fill, p, k);
            fill += k;
            p += k;
            n -= k;
        }
    }
    void final(uint8_t* out) {
        t += fill;
        memset(buf + fill, 0, 128 - fill);
        compress(buf, true);
        uint8_t full[64];
        memcpy(full, h, 64);
        memcpy(out, full, outlen);
    }
};

struct KeyDigest {
    Blake2b b{12};
    void str(const char* s, size_t n) {
        const uint32_t len = (uint32_t)n;
        b.update(&len, 4);
        b.update(s, n);
    }
    void finish(uint64_t& lo, uint64_t& hi) {
        uint8_t d[12];
        b.final(d);
        uint32_t hi32;
        memcpy(&lo, d, 8);
        memcpy(&hi32, d + 8, 4);
        hi = hi32;
    }
};

// ---- operand slots ------------------------------------------------------------------------
struct SlotKey {
    uint32_t desc;
    std::string key;
};

struct SlotTable {
    std::vector<int32_t> idx;  // open addressing, -1 = empty
    std::vector<SlotKey> keys;
    SlotTable() : idx(64, -1) {}
    static uint64_t hash(uint32_t desc, const char* k, size_t n) {
        uint64_t h = 0xcbf29ce484222325ULL ^ desc;
        h *= 0x100000001b3ULL;
        for (size_t i = 0; i < n; i++) {
            h ^= (uint8_t)k[i];
            h *= 0x100000001b3ULL;
        }
        return h ^ (h >> 29);
    }
    int find(uint32_t desc, const char* k, size_t n) const {
        const uint64_t mask = idx.size() - 1;
        for (uint64_t p = hash(desc, k, n) & mask;; p = (p + 1) & mask) {
            const int32_t s = idx[p];
            if (s < 0) return -1;
            const SlotKey& sk = keys[s];
            if (sk.desc == desc && sk.key.size() == n && memcmp(sk.key.data(), k, n) == 0) return s;
        }
    }
    uint32_t intern(uint32_t desc, const std::string& k) {
        const int f = find(desc, k.data(), k.size());
        if (f >= 0) return (uint32_t)f;
        keys.push_back({desc, k});
        if (keys.size() * 2 > idx.size()) {
            idx.assign(idx.size() * 2, -1);
            for (size_t s = 0; s + 1 < keys.size(); s++) place((int32_t)s);
        }
        place((int32_t)keys.size() - 1);
        return (uint32_t)keys.size() - 1;
    }
    void place(int32_t s) {
        const uint64_t mask = idx.size() - 1;
        uint64_t p = hash(keys[s].desc, keys[s].key.data(), keys[s].key.size()) & mask;
        while (idx[p] >= 0) p = (p + 1) & mask;
        idx[p] = s;
    }
};

struct Pred {
    uint32_t slot;
    bool neq;
    std::string lit;
};

struct MLimit {
    std::string ns, name;
    bool has_name = false, deleted = false;
    uint32_t ns_id = 0, varset_id = 0;
    uint64_t max_value = 0, seconds = 0;
    std::vector<std::string> conds, vars;  // sorted, unique (the identity)
    std::vector<Pred> preds;
    std::vector<uint32_t> var_slots;  // same order as vars
};

// ---- the accepted expression grammar (see rl_match.h) -----------------------------------------
inline bool is_ident_start(unsigned char c) { return c == '_' || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }
inline bool is_word(unsigned char c) { return is_ident_start(c) || (c >= '0' && c <= '9'); }  // CEL identifiers are ASCII
inline void skip_ws(const char*& p) {
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v') p++;
}

bool parse_operand(const char*& p, uint32_t& desc, std::string& key) {
    skip_ws(p);
    static const char kDesc[] = "descriptors[";
    if (strncmp(p, kDesc, sizeof kDesc - 1) == 0 && p[sizeof kDesc - 1] >= '0' && p[sizeof kDesc - 1] <= '9') {
        const char* q = p + sizeof kDesc - 1;
        uint64_t n = 0;
        while (*q >= '0' && *q <= '9') {
            n = n * 10 + (uint64_t)(*q - '0');
            if (n >= RL_BIND_ROOT) return false;
            q++;
        }
        if (*q != ']') return false;
        q++;
        if (*q == '.') {
            q++;
            if (!is_ident_start((unsigned char)*q)) return false;
            const char* s = q;
            while (is_word((unsigned char)*q)) q++;
            key.assign(s, q);
        } else if (*q == '[') {
            q++;
            const char kq = *q;
            if (kq != '\'' && kq != '"') return false;
            q++;
            const char* s = q;
            // the closing quote must be the opening one (CEL rejects descriptors[0]['k"] at parse time) and the
            // key must not need CEL's escape processing, which this subset does not do: refuse instead
            while (*q && *q != kq && *q != '\\') q++;
            if (*q != kq || q == s) return false;
            key.assign(s, q);
            q++;
            if (*q != ']') return false;
            q++;
        } else {
            return false;
        }
        desc = (uint32_t)n;
        p = q;
        return true;
    }
    if (!is_ident_start((unsigned char)*p)) return false;
    const char* s = p;
    while (is_word((unsigned char)*p)) p++;
    // `req.method` is member access on the variable `req` in CEL, not a root binding named "req.method"
    // (the reference would never apply such a limit: Predicate::test fails on the unbound `req`,
    // limit/cel.rs:314-322); dotted keys are only reachable as descriptors[0]['req.method'].  Refused.
    if (*p == '.') return false;
    key.assign(s, p);
    desc = RL_BIND_ROOT;
    return true;
}

bool parse_condition(const char* src, uint32_t& desc, std::string& key, bool& neq, std::string& lit) {
    const char* p = src;
    if (!parse_operand(p, desc, key)) return false;
    skip_ws(p);
    if (p[0] == '=' && p[1] == '=') neq = false;
    else if (p[0] == '!' && p[1] == '=') neq = true;
    else return false;
    p += 2;
    skip_ws(p);
    const char quote = *p;
    if (quote != '\'' && quote != '"') return false;
    p++;
    const char* s = p;
    // a literal that needs CEL's escape processing ("a\nb") would compare differently here: refused
    while (*p && *p != quote && *p != '\\') p++;
    if (*p != quote) return false;
    lit.assign(s, p);
    p++;
    skip_ws(p);
    return *p == 0;
}

bool parse_variable(const char* src, uint32_t& desc, std::string& key) {
    const char* p = src;
    if (!parse_operand(p, desc, key)) return false;
    skip_ws(p);
    return *p == 0;
}

// per-thread request scratch: slot -> value of the current request (stamped, never cleared)
struct Scratch {
    std::vector<const char*> val;
    std::vector<uint64_t> stamp;
    uint64_t epoch = 0;
    struct VK {
        uint32_t varset;
        uint64_t lo, hi;
    };
    std::vector<VK> keys;
};
thread_local Scratch tls_scratch;

}  // namespace

struct rl_matcher {
    std::shared_mutex mu;
    uint32_t counter_cap = RL_MAX_COUNTERS_PER_REQUEST;  // rl_matcher_set_counter_cap
    std::mutex err_mu;  // last_error is also written by matching calls, which hold `mu` shared
    std::string last_error;
    SlotTable slots;
    std::vector<MLimit> limits;
    std::unordered_map<std::string, uint32_t> ns_ids;
    std::vector<std::vector<uint32_t>> ns_limits;  // registration order
    std::map<std::string, uint32_t> by_identity;
    std::map<std::string, uint32_t> varsets;  // (namespace, variables) -> id, from 1
};

namespace {

int mfail(rl_matcher* m, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> g(m->err_mu);
    m->last_error = buf;
    return RL_FATAL;
}

std::string joined(const std::vector<std::string>& v) {
    std::string s;
    for (const auto& x : v) {
        s += x;
        s.push_back('\x01');
    }
    return s;
}

// counters_that_apply for one request; returns RL_OK or RL_FATAL (capacity)
int match_one(const rl_matcher* m, uint32_t ns_id, const rl_binding* binds, uint32_t nb, rl_counter* out, uint64_t cap,
              uint64_t& n_out, Scratch& s) {
    n_out = 0;
    if (ns_id >= m->ns_limits.size()) return RL_OK;  // unknown namespace: no limits apply (lib.rs:434-440)
    const size_t nslots = m->slots.keys.size();
    if (s.val.size() < nslots) {
        s.val.resize(nslots, nullptr);
        s.stamp.resize(nslots, 0);
    }
    const uint64_t ep = ++s.epoch;
    for (uint32_t i = 0; i < nb; i++) {
        const rl_binding& b = binds[i];
        if (!b.key || !b.value) continue;
        const int slot = m->slots.find(b.descriptor, b.key, strlen(b.key));
        if (slot < 0) continue;  // nothing refers to this binding
        s.val[slot] = b.value;   // a HashMap holds one value per key: the last one given wins
        s.stamp[slot] = ep;
    }
    s.keys.clear();
    for (const uint32_t lid : m->ns_limits[ns_id]) {
        const MLimit& L = m->limits[lid];
        if (L.deleted) continue;
        bool ok = true;
        for (const Pred& p : L.preds) {
            const char* v = (s.stamp[p.slot] == ep) ? s.val[p.slot] : nullptr;
            if (!v) {  // unbound name / missing key: the predicate is false (cel.rs:315-331)
                ok = false;
                break;
            }
            const bool eq = strcmp(v, p.lit.c_str()) == 0;
            if (eq == p.neq) {
                ok = false;
                break;
            }
        }
        if (!ok) continue;
        for (const uint32_t vs : L.var_slots)
            if (s.stamp[vs] != ep) {  // a variable without a value: no counter (limit.rs:133-148, lib.rs:515-519)
                ok = false;
                break;
            }
        if (!ok) continue;
        // the engine takes at most RL_MAX_COUNTERS_PER_REQUEST counters per request: refuse here, before
        // anything is enqueued, instead of letting the device resolve fail the batch half-applied
        if (n_out >= cap || n_out >= m->counter_cap) return RL_FATAL;
        rl_counter& c = out[n_out++];
        c.limit_id = lid;
        c._pad = 0;
        c.key_lo = c.key_hi = 0;
        if (!L.var_slots.empty()) {
            bool hit = false;
            for (const auto& k : s.keys)
                if (k.varset == L.varset_id) {
                    c.key_lo = k.lo;
                    c.key_hi = k.hi;
                    hit = true;
                    break;
                }
            if (!hit) {
                KeyDigest d;
                for (size_t j = 0; j < L.vars.size(); j++) {
                    d.str(L.vars[j].data(), L.vars[j].size());
                    const char* v = s.val[L.var_slots[j]];
                    d.str(v, strlen(v));
                }
                d.finish(c.key_lo, c.key_hi);
                s.keys.push_back({L.varset_id, c.key_lo, c.key_hi});
            }
        }
    }
    return RL_OK;
}

}  // namespace

extern "C" {

int rl_matcher_create(rl_matcher** out) {
    if (!out) return RL_FATAL;
    *out = new rl_matcher();
    return RL_OK;
}

void rl_matcher_destroy(rl_matcher* m) { delete m; }

const char* rl_matcher_last_error(rl_matcher* m) { return m ? m->last_error.c_str() : "null matcher"; }

int rl_matcher_last_error_copy(rl_matcher* m, char* out, uint32_t cap) {
    if (!m || !out || !cap) return RL_FATAL;
    std::lock_guard<std::mutex> g(m->err_mu);
    snprintf(out, cap, "%s", m->last_error.c_str());
    return RL_OK;
}

int rl_matcher_add_limit(rl_matcher* m, const char* ns, uint64_t max_value, uint64_t seconds,
                         const char* const* conditions, uint32_t n_cond, const char* const* variables, uint32_t n_var,
                         const char* name, rl_limit_desc* out_desc) {
    return rl_matcher_add_limit_ex(m, ns, max_value, seconds, conditions, n_cond, variables, n_var, name, 0, out_desc, nullptr);
}

int rl_matcher_add_limit_ex(rl_matcher* m, const char* ns, uint64_t max_value, uint64_t seconds,
                            const char* const* conditions, uint32_t n_cond, const char* const* variables, uint32_t n_var,
                            const char* name, int keep_existing, rl_limit_desc* out_desc, int* out_existed) {
    if (!m || !ns || !out_desc || (n_cond && !conditions) || (n_var && !variables)) return RL_FATAL;
    if (out_existed) *out_existed = 0;
    std::unique_lock<std::shared_mutex> lock(m->mu);
    MLimit L;
    L.ns = ns;
    L.max_value = max_value;
    L.seconds = seconds;
    if (name) {
        L.name = name;
        L.has_name = true;
    }
    for (uint32_t i = 0; i < n_cond; i++) L.conds.emplace_back(conditions[i] ? conditions[i] : "");
    for (uint32_t i = 0; i < n_var; i++) L.vars.emplace_back(variables[i] ? variables[i] : "");
    // the identity holds SETS of expression sources (limit.rs:31-48: BTreeSet)
    std::sort(L.conds.begin(), L.conds.end());
    L.conds.erase(std::unique(L.conds.begin(), L.conds.end()), L.conds.end());
    std::sort(L.vars.begin(), L.vars.end());
    L.vars.erase(std::unique(L.vars.begin(), L.vars.end()), L.vars.end());
    // parse before touching any table: a refused limit leaves the matcher unchanged
    struct Parsed {
        uint32_t desc;
        std::string key;
        bool neq;
        std::string lit;
    };
    std::vector<Parsed> pc(L.conds.size()), pv(L.vars.size());
    for (size_t i = 0; i < L.conds.size(); i++)
        if (!parse_condition(L.conds[i].c_str(), pc[i].desc, pc[i].key, pc[i].neq, pc[i].lit))
            return mfail(m, "unsupported condition expression: %s", L.conds[i].c_str());
    for (size_t i = 0; i < L.vars.size(); i++)
        if (!parse_variable(L.vars[i].c_str(), pv[i].desc, pv[i].key))
            return mfail(m, "unsupported variable expression: %s", L.vars[i].c_str());

    const std::string ident = L.ns + '\0' + std::to_string(seconds) + '\0' + joined(L.conds) + '\0' + joined(L.vars);
    uint32_t lid;
    auto it = m->by_identity.find(ident);
    if (it != m->by_identity.end()) {
        // update_limit (storage/mod.rs:67-83): same identity, new max_value / name; a deleted one comes back
        lid = it->second;
        MLimit& E = m->limits[lid];
        if (out_existed) *out_existed = E.deleted ? 0 : 1;
        // Storage::add_limit is a HashSet::insert: on an equal (live) element it is a no-op and the OLD
        // max_value / name stay (storage/mod.rs:60-65); only update_limit swaps them (:67-83)
        if (!(keep_existing && !E.deleted)) {
            E.max_value = max_value;
            E.name = L.name;
            E.has_name = L.has_name;
        }
        if (E.deleted) {  // deleted and added again: it is the namespace's newest limit
            auto& order = m->ns_limits[E.ns_id];
            order.erase(std::remove(order.begin(), order.end(), lid), order.end());
            order.push_back(lid);
            E.deleted = false;
        }
    } else {
        lid = (uint32_t)m->limits.size();
        auto nit = m->ns_ids.find(L.ns);
        if (nit == m->ns_ids.end()) {
            nit = m->ns_ids.emplace(L.ns, (uint32_t)m->ns_ids.size()).first;
            m->ns_limits.emplace_back();
        }
        L.ns_id = nit->second;
        if (!L.vars.empty()) {
            const std::string vk = L.ns + '\0' + joined(L.vars);
            auto vit = m->varsets.find(vk);
            if (vit == m->varsets.end()) vit = m->varsets.emplace(vk, (uint32_t)m->varsets.size() + 1).first;
            L.varset_id = vit->second;
        }
        for (const auto& c : pc) L.preds.push_back({m->slots.intern(c.desc, c.key), c.neq, c.lit});
        for (const auto& v : pv) L.var_slots.push_back(m->slots.intern(v.desc, v.key));
        m->by_identity.emplace(ident, lid);
        m->ns_limits[L.ns_id].push_back(lid);
        m->limits.push_back(std::move(L));
    }
    const MLimit& R = m->limits[lid];
    out_desc->limit_id = lid;
    out_desc->ns_id = R.ns_id;
    out_desc->varset_id = R.varset_id;
    out_desc->qualified = R.vars.empty() ? 0 : 1;  // counter.rs:108-110
    out_desc->max_value = R.max_value;
    out_desc->window_us = R.seconds * 1000000ull;  // counter.rs:76-78
    return RL_OK;
}

int rl_matcher_set_counter_cap(rl_matcher* m, uint32_t cap) {
    if (!m || cap == 0) return RL_FATAL;
    std::unique_lock<std::shared_mutex> lock(m->mu);
    m->counter_cap = cap;
    return RL_OK;
}

int rl_matcher_delete_limit(rl_matcher* m, uint32_t limit_id) {
    if (!m) return RL_FATAL;
    std::unique_lock<std::shared_mutex> lock(m->mu);
    if (limit_id >= m->limits.size()) return mfail(m, "unknown limit_id %u", limit_id);
    m->limits[limit_id].deleted = true;
    return RL_OK;
}

int rl_matcher_namespace_id(rl_matcher* m, const char* ns, uint32_t* out_ns_id) {
    if (!m || !ns || !out_ns_id) return RL_FATAL;
    std::shared_lock<std::shared_mutex> lock(m->mu);
    auto it = m->ns_ids.find(ns);
    if (it == m->ns_ids.end()) return RL_FATAL;
    *out_ns_id = it->second;
    return RL_OK;
}

const char* rl_matcher_limit_name(rl_matcher* m, uint32_t limit_id) {
    if (!m) return nullptr;
    std::shared_lock<std::shared_mutex> lock(m->mu);
    if (limit_id >= m->limits.size() || !m->limits[limit_id].has_name) return nullptr;
    return m->limits[limit_id].name.c_str();
}

int rl_matcher_limit_name_copy(rl_matcher* m, uint32_t limit_id, char* out, uint32_t cap, int* out_has_name) {
    if (!m || !out || !cap) return RL_FATAL;
    std::shared_lock<std::shared_mutex> lock(m->mu);
    out[0] = 0;
    if (out_has_name) *out_has_name = 0;
    if (limit_id >= m->limits.size()) return mfail(m, "unknown limit_id %u", limit_id);
    const MLimit& L = m->limits[limit_id];
    if (!L.has_name) return RL_OK;
    if (L.name.size() + 1 > cap) return mfail(m, "limit name needs %zu bytes", L.name.size() + 1);
    memcpy(out, L.name.c_str(), L.name.size() + 1);
    if (out_has_name) *out_has_name = 1;
    return RL_OK;
}

int rl_matcher_counters(rl_matcher* m, uint32_t ns_id, const rl_binding* binds, uint32_t n_binds, rl_counter* out_ctrs,
                        uint32_t cap, uint32_t* out_n) {
    if (!m || !out_n || (n_binds && !binds) || (cap && !out_ctrs)) return RL_FATAL;
    std::shared_lock<std::shared_mutex> lock(m->mu);
    uint64_t n = 0;
    const int r = match_one(m, ns_id, binds, n_binds, out_ctrs, cap, n, tls_scratch);
    *out_n = (uint32_t)n;
    if (r) return mfail(m, "more than %u counters apply to one request (engine limit %d)", cap, RL_MAX_COUNTERS_PER_REQUEST);
    return RL_OK;
}

int rl_matcher_counters_batch(rl_matcher* m, uint64_t n, const uint32_t* ns_id, const uint32_t* bind_off,
                              const rl_binding* binds, uint32_t* out_ctr_off, rl_counter* out_ctrs, uint64_t cap) {
    if (!m || (n && (!ns_id || !bind_off || !out_ctr_off)) || (cap && !out_ctrs)) return RL_FATAL;
    std::shared_lock<std::shared_mutex> lock(m->mu);
    Scratch& s = tls_scratch;
    uint64_t total = 0;
    if (out_ctr_off) out_ctr_off[0] = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t k = 0;
        const int r = match_one(m, ns_id[i], binds + bind_off[i], bind_off[i + 1] - bind_off[i], out_ctrs + total,
                                cap - total, k, s);
        if (r || total + k > 0xFFFFFFFFull) {
            return mfail(m, "counter capacity %llu exhausted, or more than %d counters apply, at request %llu",
                         (unsigned long long)cap, RL_MAX_COUNTERS_PER_REQUEST, (unsigned long long)i);
        }
        total += k;
        out_ctr_off[i + 1] = (uint32_t)total;
    }
    return RL_OK;
}

// CheckResult::response_header (lib.rs:235-275) of one request into `out`: three NUL-terminated values one after the other
// (X-RateLimit-Limit, -Remaining, -Reset).  Returns the bytes written, 0 when they do not fit in cap, or -1 for an unknown
// limit id.  The caller holds the matcher's lock (shared).  No heap traffic: a batching stage calls it per request.
static int64_t format_headers(const rl_matcher* m, const rl_counter* ctrs, const uint64_t* remaining, const uint64_t* ttl_us, uint32_t n,
                              char* out, uint64_t cap) {
    if (n == 0) {
        if (cap < 3) return 0;
        out[0] = out[1] = out[2] = 0;
        return 3;
    }
    uint32_t small[64];
    std::vector<uint32_t> big;
    uint32_t* order = small;
    if (n > 64) {
        big.resize(n);
        order = big.data();
    }
    for (uint32_t i = 0; i < n; i++) {
        if (ctrs[i].limit_id >= m->limits.size()) return -1;
        // insertion sort by remaining, stable: the most restrictive first (lib.rs:238-242; sort_by is stable)
        uint32_t k = i;
        while (k > 0 && remaining[order[k - 1]] > remaining[i]) {
            order[k] = order[k - 1];
            k--;
        }
        order[k] = i;
    }
    uint64_t w = 0;
    auto put = [&](const char* fmt, unsigned long long a, unsigned long long b) {
        if (w >= cap) return false;
        const int k = snprintf(out + w, cap - w, fmt, a, b);
        if (k < 0 || (uint64_t)k >= cap - w) return false;
        w += (uint64_t)k;
        return true;
    };
    const uint32_t f = order[0];
    if (!put("%llu", m->limits[ctrs[f].limit_id].max_value, 0)) return 0;
    for (uint32_t q = 0; q < n; q++) {  // ", <max>;w=<secs>[;name=\"...\"]" for every counter (lib.rs:244-252)
        const MLimit& L = m->limits[ctrs[order[q]].limit_id];
        if (!put(", %llu;w=%llu", L.max_value, L.seconds)) return 0;
        if (L.has_name) {
            if (w + L.name.size() + 9 >= cap) return 0;
            memcpy(out + w, ";name=\"", 7);
            w += 7;
            for (const char c : L.name) out[w++] = c == '"' ? '\'' : c;
            out[w++] = '"';
        }
    }
    if (w >= cap) return 0;
    out[w++] = 0;
    if (!put("%llu", remaining[f], 0)) return 0;
    if (w >= cap) return 0;
    out[w++] = 0;
    if (!put("%llu", ttl_us[f] / 1000000ull, 0)) return 0;  // Duration::as_secs (lib.rs:268-270)
    if (w >= cap) return 0;
    out[w++] = 0;
    return (int64_t)w;
}

int rl_matcher_counters_batch_ns(rl_matcher* m, uint64_t n, const char* const* ns, const uint32_t* bind_off, const rl_binding* binds,
                                 uint32_t* out_ctr_off, rl_counter* out_ctrs, uint64_t cap, uint8_t* out_status) {
    if (!m || !out_ctr_off || (n && (!ns || !bind_off || !out_status)) || (cap && !out_ctrs)) return RL_FATAL;
    std::shared_lock<std::shared_mutex> lock(m->mu);  // one reader section for the whole range
    Scratch& s = tls_scratch;
    uint64_t total = 0;
    out_ctr_off[0] = 0;
    const char* last_ns = nullptr;  // consecutive requests of one namespace look it up once
    bool last_known = false;
    uint32_t last_id = 0;
    for (uint64_t i = 0; i < n; i++) {
        out_status[i] = 0;
        if (!ns[i]) return mfail(m, "request %llu has no namespace", (unsigned long long)i);
        if (!last_ns || strcmp(ns[i], last_ns) != 0) {
            const auto it = m->ns_ids.find(ns[i]);
            last_ns = ns[i];
            last_known = it != m->ns_ids.end();
            last_id = last_known ? it->second : 0;
        }
        if (!last_known) {
            out_status[i] = 1;  // no limit was ever added for the namespace: nothing applies (lib.rs:434-440)
        } else {
            uint64_t k = 0;
            if (cap - total < m->counter_cap) return mfail(m, "counter capacity %llu exhausted at request %llu", (unsigned long long)cap, (unsigned long long)i);
            if (match_one(m, last_id, binds + bind_off[i], bind_off[i + 1] - bind_off[i], out_ctrs + total, cap - total, k, s) != RL_OK) {
                out_status[i] = 2;  // more counters apply than one request may carry: the request gets none
                k = 0;
            }
            total += k;
        }
        if (total > 0xFFFFFFFFull) return mfail(m, "more than 2^32 counters in one batch");
        out_ctr_off[i + 1] = (uint32_t)total;
    }
    return RL_OK;
}

int rl_matcher_response_headers(rl_matcher* m, const rl_counter* ctrs, const uint64_t* remaining, const uint64_t* ttl_us,
                                uint32_t n, char* out_limit, uint32_t cap_limit, char* out_remaining, uint32_t cap_remaining,
                                char* out_reset, uint32_t cap_reset) {
    if (!m || !out_limit || !out_remaining || !out_reset || !cap_limit || !cap_remaining || !cap_reset ||
        (n && (!ctrs || !remaining || !ttl_us)))
        return RL_FATAL;
    out_limit[0] = out_remaining[0] = out_reset[0] = 0;
    if (n == 0) return RL_OK;
    std::shared_lock<std::shared_mutex> lock(m->mu);
    std::vector<char> text((size_t)cap_limit + 64);
    const int64_t w = format_headers(m, ctrs, remaining, ttl_us, n, text.data(), text.size());
    if (w < 0) return mfail(m, "unknown limit_id in the counters of a response");
    const char* lim = text.data();
    const char* rem = w ? lim + strlen(lim) + 1 : lim;
    const char* rst = w ? rem + strlen(rem) + 1 : lim;
    if (w == 0 || strlen(lim) + 1 > cap_limit || strlen(rem) + 1 > cap_remaining || strlen(rst) + 1 > cap_reset)
        return mfail(m, "header buffer too small for X-RateLimit-Limit (%u bytes given)", cap_limit);
    memcpy(out_limit, lim, strlen(lim) + 1);
    memcpy(out_remaining, rem, strlen(rem) + 1);
    memcpy(out_reset, rst, strlen(rst) + 1);
    return RL_OK;
}

int rl_matcher_response_headers_batch(rl_matcher* m, uint64_t n, const uint32_t* ctr_off, const rl_counter* ctrs, const uint64_t* remaining,
                                      const uint64_t* ttl_us, char* out, uint64_t cap, uint64_t* out_off, uint64_t* out_len) {
    if (!m || !out_off || !out_len || (n && (!ctr_off || !ctrs || !remaining || !ttl_us)) || (cap && !out)) return RL_FATAL;
    std::shared_lock<std::shared_mutex> lock(m->mu);  // one reader section for the whole range
    uint64_t w = 0;
    bool fits = true;
    out_off[0] = 0;
    char scratch[4096];
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t o = ctr_off[i], k = ctr_off[i + 1] - o;
        int64_t got = fits ? format_headers(m, ctrs + o, remaining + o, ttl_us + o, k, out + w, cap - w) : 0;
        if (got < 0) return mfail(m, "unknown limit_id in the counters of request %llu", (unsigned long long)i);
        if (got == 0) {  // does not fit (any more): keep measuring so that *out_len tells the caller what to bring
            fits = false;
            std::vector<char> big;
            char* tmp = scratch;
            uint64_t tcap = sizeof scratch;
            while ((got = format_headers(m, ctrs + o, remaining + o, ttl_us + o, k, tmp, tcap)) == 0) {
                big.resize(tcap * 4);
                tmp = big.data();
                tcap = big.size();
            }
            if (got < 0) return mfail(m, "unknown limit_id in the counters of request %llu", (unsigned long long)i);
        }
        w += (uint64_t)got;
        out_off[i + 1] = w;
    }
    *out_len = w;
    return fits ? RL_OK : mfail(m, "header text needs %llu bytes", (unsigned long long)w);
}

void rl_counter_key(const char* const* sources, const char* const* values, uint32_t n, uint64_t* key_lo, uint64_t* key_hi) {
    *key_lo = *key_hi = 0;
    if (n == 0) return;
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return strcmp(sources[a], sources[b]) < 0; });
    KeyDigest d;
    for (const uint32_t i : order) {
        d.str(sources[i], strlen(sources[i]));
        d.str(values[i], strlen(values[i]));
    }
    d.finish(*key_lo, *key_hi);
}

}  // extern "C"
