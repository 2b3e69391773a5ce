// Sanitizer driver (ASan+UBSan build and ThreadSanitizer build) of the batching front with the matcher inside
// (limitador_b200/csrc/rl_front.cu is plain host C++): 8 threads call rl_front_check_and_update_bindings concurrently
// while another thread keeps adding limits.  The store call is a stub that answers from the CSR it is given (no engine
// here), checking on the way that every request's counters arrive intact in the dispatcher's batch.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rl_match.h"

static std::atomic<uint64_t> g_batches{0}, g_requests{0}, g_bad{0};

extern "C" {
const char* rl_last_error(rl_engine*) { return "stub"; }
int rl_check_and_update_batch(rl_engine*, uint64_t n, const uint32_t* off, const rl_counter* ctrs, const uint64_t* delta, const uint64_t* now,
                              int lc, int, uint8_t* limited, uint32_t* first, uint64_t* rem, uint64_t* ttl) {
    g_batches++;
    g_requests += n;
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t m = off[i + 1] - off[i];
        if (m == 0 || m > RL_MAX_COUNTERS_PER_REQUEST || delta[i] == 0 || now[i] == 0) g_bad++;
        limited[i] = (uint8_t)(ctrs[off[i]].key_lo & 1);  // a verdict derived from the request's own first counter
        first[i] = limited[i] ? ctrs[off[i]].limit_id : RL_NONE;
        if (lc)
            for (uint32_t k = 0; k < m; k++) {
                rem[off[i] + k] = ctrs[off[i] + k].key_lo;
                ttl[off[i] + k] = ctrs[off[i] + k].limit_id;
            }
    }
    return RL_OK;
}
}

int main() {
    rl_matcher* m = nullptr;
    rl_front* f = nullptr;
    if (rl_matcher_create(&m) != RL_OK) return 2;
    rl_limit_desc d;
    const char* conds[] = {"descriptors[0].method == 'GET'"};
    const char* vars[] = {"descriptors[0].user"};
    for (int ns = 0; ns < 4; ns++) {
        const std::string name = "ns" + std::to_string(ns);
        if (rl_matcher_add_limit(m, name.c_str(), 10, 60, conds, 1, vars, 1, "per-user", &d) != RL_OK) return 3;
        if (rl_matcher_add_limit(m, name.c_str(), 100, 3600, nullptr, 0, vars, 1, nullptr, &d) != RL_OK) return 3;
    }
    if (rl_front_create((rl_engine*)0x1, 64, 100, &f) != RL_OK) return 4;  // the stub never dereferences the engine
    std::atomic<bool> stop{false};
    std::atomic<uint64_t> wrong{0}, done{0};
    std::thread writer([&] {  // limits come and go while requests are matched (the matcher's reader/writer lock)
        int k = 0;
        while (!stop) {
            const std::string name = "extra" + std::to_string(k++ % 7);
            rl_limit_desc dd;
            rl_matcher_add_limit(m, name.c_str(), 5, 10, nullptr, 0, vars, 1, nullptr, &dd);
            rl_matcher_delete_limit(m, dd.limit_id);
            std::this_thread::yield();
        }
    });
    std::vector<std::thread> callers;
    for (int t = 0; t < 8; t++)
        callers.emplace_back([&, t] {
            for (int i = 0; i < 1500; i++) {
                const std::string ns = i % 11 == 0 ? "nobody" : "ns" + std::to_string((t + i) % 4);
                const std::string user = "u" + std::to_string((t * 31 + i) % 50);
                const rl_binding binds[2] = {{0, 0, "method", i % 3 ? "GET" : "POST"}, {0, 0, "user", user.c_str()}};
                uint8_t lim = 9;
                uint32_t first = 0, n = 99;
                rl_counter ctrs[RL_MAX_COUNTERS_PER_REQUEST];
                uint64_t rem[RL_MAX_COUNTERS_PER_REQUEST], ttl[RL_MAX_COUNTERS_PER_REQUEST], seq = 0;
                const int lc = i % 2;
                if (rl_front_check_and_update_bindings(f, m, ns.c_str(), binds, 2, 1 + i % 3, 1700000000000000ull, lc, &lim, &first,
                                                       i % 5 ? ctrs : nullptr, &n, rem, ttl, &seq) != RL_OK) {
                    wrong++;
                    continue;
                }
                const uint32_t want_n = ns == "nobody" ? 0 : (i % 3 ? 2 : 1);
                if (n != want_n || (n == 0 && (lim != 0 || first != RL_NONE))) wrong++;
                if (n && i % 5) {  // the verdict and the load_counters outputs were derived from OUR counters
                    if (lim != (ctrs[0].key_lo & 1) || (lim && first != ctrs[0].limit_id)) wrong++;
                    if (lc)
                        for (uint32_t k = 0; k < n; k++)
                            if (rem[k] != ctrs[k].key_lo || ttl[k] != ctrs[k].limit_id) wrong++;
                }
                done++;
            }
        });
    for (auto& c : callers) c.join();
    stop = true;
    writer.join();
    uint64_t batches = 0, requests = 0;
    rl_front_stats(f, &batches, &requests);
    rl_front_destroy(f);
    rl_matcher_destroy(m);
    if (wrong || g_bad || batches != g_batches || requests != g_requests) {
        printf("FAILED wrong=%llu bad=%llu\n", (unsigned long long)wrong.load(), (unsigned long long)g_bad.load());
        return 5;
    }
    printf("ok front requests=%llu batches=%llu\n", (unsigned long long)requests, (unsigned long long)batches);
    return 0;
}
