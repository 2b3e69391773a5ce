// ASan/UBSan driver of the RLS wire surface (limitador_b200/csrc/rl_rls.cpp): the decoder takes bytes from the network, so it
// is fuzzed with mutated and truncated messages; the plan / finish stages run over batches that mix good, malformed and
// unsupported requests (store outputs are made up: no engine here).  The crdt oracle rides along.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "rl_rls.h"

// the engine entry points rl_rls_serve would call: never reached here (the service is created without an engine)
extern "C" {
const char* rl_last_error(rl_engine*) { return "no engine in the sanitizer build"; }
int rl_check_and_update_batch(rl_engine*, uint64_t, const uint32_t*, const rl_counter*, const uint64_t*, const uint64_t*, int, int,
                              uint8_t*, uint32_t*, uint64_t*, uint64_t*) { return RL_FATAL; }
int rl_is_within_limits_batch(rl_engine*, uint64_t, const uint32_t*, const rl_counter*, const uint64_t*, const uint64_t*, int, uint8_t*,
                              uint32_t*) { return RL_FATAL; }
int rl_update_batch(rl_engine*, uint64_t, const uint32_t*, const rl_counter*, const uint64_t*, const uint64_t*, int) { return RL_FATAL; }
int rl_front_check_and_update(rl_front*, const rl_counter*, uint32_t, uint64_t, uint64_t, int, uint8_t*, uint32_t*, uint64_t*, uint64_t*,
                              uint64_t*) { return RL_FATAL; }
}

static void put_varint(std::string& o, uint64_t v) {
    while (v >= 0x80) {
        o.push_back((char)(v | 0x80));
        v >>= 7;
    }
    o.push_back((char)v);
}
static void put_len(std::string& o, uint32_t tag, const std::string& body) {
    put_varint(o, (tag << 3) | 2);
    put_varint(o, body.size());
    o += body;
}
static std::string request(const std::string& domain, const std::vector<std::vector<std::pair<std::string, std::string>>>& descs, uint32_t hits) {
    std::string o;
    if (!domain.empty()) put_len(o, 1, domain);
    for (const auto& d : descs) {
        std::string body;
        for (const auto& kv : d) {
            std::string e;
            if (!kv.first.empty()) put_len(e, 1, kv.first);
            if (!kv.second.empty()) put_len(e, 2, kv.second);
            put_len(body, 1, e);
        }
        put_len(o, 2, body);
    }
    if (hits) {
        put_varint(o, 3 << 3);
        put_varint(o, hits);
    }
    return o;
}

int main() {
    std::mt19937_64 rng(42);
    // 1. decoder fuzz: mutations, truncations, random bytes; small entry capacity so that the overflow path runs too
    const std::string base = request("test_namespace", {{{"req.method", "GET"}, {"app.id", "1"}, {"ü", "日本"}}, {{"y", "2"}}, {}}, 6) +
                             std::string("\x7a\x01\x66\x81\x01\x01\x02\x03\x04\x05\x06\x07\x08\x8d\x01\x01\x02\x03\x04\x7b\x08\x01\x7c", 23);
    uint64_t ok = 0, bad = 0;
    for (int it = 0; it < 200000; it++) {
        std::string b = base;
        const int kind = it % 4;
        if (kind == 0) b.resize(rng() % (b.size() + 1));
        else if (kind == 3) {
            b.resize(rng() % 64);
            for (auto& c : b) c = (char)rng();
        } else
            for (int k = 0; k < 1 + (int)(rng() % 3); k++) b[rng() % b.size()] = (char)rng();
        rl_rls_request q;
        rl_rls_entry ent[3];
        const int r = rl_rls_decode_request((const uint8_t*)b.data(), b.size(), &q, ent, 3);
        if (r == RL_OK) {
            ok++;
            // every range the decoder reports lies inside the message
            if ((uint64_t)q.domain_off + q.domain_len > b.size()) return 2;
            for (uint32_t k = 0; k < q.n_entries && k < 3; k++)
                if ((uint64_t)ent[k].key_off + ent[k].key_len > b.size() || (uint64_t)ent[k].val_off + ent[k].val_len > b.size() ||
                    ent[k].descriptor >= q.n_descriptors)
                    return 3;
        } else {
            bad++;
        }
    }
    // 2. the service's CPU stages over mixed batches
    rl_matcher* m = nullptr;
    if (rl_matcher_create(&m) != RL_OK) return 4;
    rl_limit_desc d;
    const char* c1[] = {"descriptors[0]['req.method'] == 'GET'"};
    const char* v1[] = {"descriptors[0]['app.id']"};
    if (rl_matcher_add_limit(m, "test_namespace", 1, 60, c1, 1, v1, 1, "a \"named\" limit", &d) != RL_OK) return 5;
    if (rl_matcher_add_limit(m, "test_namespace", 100, 3600, nullptr, 0, v1, 1, nullptr, &d) != RL_OK) return 5;
    uint64_t responses = 0;
    for (int threads = 1; threads <= 3; threads += 2) {
        rl_rls* s = nullptr;
        if (rl_rls_create(m, nullptr, RL_RLS_HEADERS_DRAFT_VERSION_03, threads, 1, &s) != RL_OK) return 6;
        for (int method = 0; method < 3; method++) {
            std::string buf;
            std::vector<uint64_t> off{0};
            for (int i = 0; i < 500; i++) {
                std::string msg;
                switch (rng() % 6) {
                    case 0: msg = request("", {{{"a", "b"}}}, 1); break;
                    case 1: msg = request("nobody", {{{"a", "b"}}}, 0); break;
                    case 2: msg = base.substr(0, rng() % base.size()); break;
                    case 3: msg = request("test_namespace", {{{"req.method", "GET"}, {"app.id", std::string("a\0b", 3)}}}, 1); break;
                    default: msg = request("test_namespace", {{{"req.method", rng() % 2 ? "GET" : "POST"}, {"app.id", std::to_string(rng() % 5)}}}, (uint32_t)(rng() % 3)); break;
                }
                buf += msg;
                off.push_back(buf.size());
            }
            if (rl_rls_plan(s, method, 500, (const uint8_t*)buf.data(), off.data(), 1700000000000000ull) != RL_OK) return 7;
            uint64_t n_store = 0;
            const uint32_t* ctr_off = nullptr;
            int lc = 0;
            if (rl_rls_plan_view(s, &n_store, &ctr_off, nullptr, nullptr, nullptr, &lc, nullptr) != RL_OK) return 8;
            std::vector<uint8_t> lim(n_store + 1);
            std::vector<uint32_t> first(n_store + 1, RL_NONE);
            const uint64_t n_ctr = n_store ? ctr_off[n_store] : 0;
            std::vector<uint64_t> rem(n_ctr + 1), ttl(n_ctr + 1);
            for (uint64_t j = 0; j < n_store; j++) {
                lim[j] = (uint8_t)(rng() % 8 == 0 ? RL_VERDICT_ERROR : rng() % 2);
                if (lim[j] == 1) first[j] = (uint32_t)(rng() % 3);  // 2 = an id the matcher does not know
            }
            for (auto& x : rem) x = rng() % 100;
            for (auto& x : ttl) x = rng() % 60000000;
            if (rl_rls_finish(s, method == 2 && threads == 3 ? RL_TRANSIENT : RL_OK, lim.data(), first.data(), rem.data(), ttl.data()) != RL_OK) return 9;
            const uint8_t *out, *grpc, *code;
            const uint64_t* ooff;
            if (rl_rls_responses(s, &out, &ooff, &grpc, &code) != RL_OK) return 10;
            for (int i = 0; i < 500; i++) responses += (ooff[i + 1] - ooff[i]) + grpc[i] + code[i];
        }
        uint64_t need = 0;
        rl_rls_metrics_render(s, nullptr, 0, &need);
        std::vector<char> text(need);
        if (rl_rls_metrics_render(s, text.data(), need, &need) != RL_OK) return 11;
        rl_rls_destroy(s);
    }
    rl_matcher_destroy(m);
    printf("ok decoded=%llu refused=%llu responses=%llu\n", (unsigned long long)ok, (unsigned long long)bad, (unsigned long long)responses);
    return ok > 1000 && bad > 1000 ? 0 : 12;
}
