// AddressSanitizer / UBSan driver for the native matcher (include/rl_match.h): limits with accepted and
// refused expressions, updates, deletes, re-adds, random requests from two threads.  Test infrastructure.
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rl_match.h"

static unsigned long long rnd(unsigned long long& s) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return s;
}

int main() {
    rl_matcher* m = nullptr;
    if (rl_matcher_create(&m) != RL_OK) return 1;
    const char* operands[] = {"a", "req.path", "descriptors[0].k", "descriptors[1]['x.y']", "descriptors[0]", "9z", "descriptors[0].u"};
    const char* lits[] = {"'v'", "\"w\"", "''", "'unterminated", "bare"};
    const char* values[] = {"v", "w", "", "zzz"};
    unsigned long long s = 0x9E3779B97F4A7C15ULL;
    unsigned added = 0, refused = 0;
    std::vector<uint32_t> ids;
    for (int i = 0; i < 600; i++) {
        std::vector<std::string> conds, vars;
        for (unsigned k = 0; k < rnd(s) % 4; k++)
            conds.push_back(std::string(operands[rnd(s) % 7]) + (rnd(s) & 1 ? " == " : "!=") + lits[rnd(s) % 5]);
        for (unsigned k = 0; k < rnd(s) % 3; k++) vars.push_back(operands[rnd(s) % 7]);
        std::vector<const char*> pc, pv;
        for (auto& c : conds) pc.push_back(c.c_str());
        for (auto& v : vars) pv.push_back(v.c_str());
        rl_limit_desc d;
        const std::string ns = "ns" + std::to_string(rnd(s) % 5);
        const int r = rl_matcher_add_limit(m, ns.c_str(), rnd(s) % 100, 1 + rnd(s) % 4, pc.data(), (uint32_t)pc.size(), pv.data(),
                                           (uint32_t)pv.size(), (rnd(s) & 1) ? "named" : nullptr, &d);
        if (r == RL_OK) {
            added++;
            ids.push_back(d.limit_id);
            if (rnd(s) % 5 == 0) rl_matcher_delete_limit(m, ids[rnd(s) % ids.size()]);
        } else {
            refused++;
            if (!strstr(rl_matcher_last_error(m), "unsupported")) return 2;
        }
    }
    auto worker = [&](unsigned long long seed) {
        unsigned long long t = seed;
        std::vector<rl_counter> out(256);
        for (int i = 0; i < 20000; i++) {
            rl_binding b[8];
            const unsigned nb = rnd(t) % 8;
            const char* keys[] = {"a", "req.path", "k", "x.y", "u", "other"};
            for (unsigned k = 0; k < nb; k++) {
                b[k].descriptor = (rnd(t) % 3 == 0) ? RL_BIND_ROOT : (uint32_t)(rnd(t) % 3);
                b[k]._pad = 0;
                b[k].key = keys[rnd(t) % 6];
                b[k].value = values[rnd(t) % 4];
            }
            uint32_t n = 0;
            if (rl_matcher_counters(m, (uint32_t)(rnd(t) % 7), b, nb, out.data(), (uint32_t)out.size(), &n) != RL_OK) break;
        }
    };
    std::thread t1(worker, 1234567ULL), t2(worker, 7654321ULL);
    t1.join();
    t2.join();
    uint32_t ns_id = 0;
    rl_matcher_namespace_id(m, "ns0", &ns_id);
    rl_matcher_limit_name(m, 0);
    printf("ok added=%u refused=%u\n", added, refused);
    rl_matcher_destroy(m);
    return 0;
}
