// ASan/UBSan driver of the kernels that run under tests/emu/cuda_shim.h (rl_maint.cuh, rl_crdt.cuh): the same kernel
// source the GPU compiles, over heap buffers sized exactly as rl_maint.cu / rl_crdt.cu size them, so that an out-of-range
// row index, a misaligned 16-byte access or an overflow in the index arithmetic is reported by the sanitizers.
#include "../emu/emu_maint.cpp"

#include <cstdio>
#include <random>

int main() {
    std::mt19937_64 rng(7);
    // metrics: batch sizes around the warp / block / grid-stride boundaries, both record forms
    for (uint32_t n : {0u, 1u, 31u, 32u, 33u, 255u, 256u, 257u, 1792u, 1793u, 5000u}) {
        for (uint32_t words : {4u, 2u}) {
            std::vector<unsigned long long> recs((size_t)n * words);  // exact sizes: one element too far is a report
            std::vector<uint8_t> lim(n);
            std::vector<uint32_t> first(n);
            for (uint32_t i = 0; i < n; i++) {
                const uint64_t ns = rng() % 70, hits = 1 + rng() % 200;
                recs[(size_t)i * words] = words == 4 ? (ns | (hits << 32)) : (ns | (hits << 24) | ((rng() & 0xFFFFFFFFull) << 32));
                lim[i] = (uint8_t)(rng() % 16 == 0 ? 0xFF : rng() % 2);
                first[i] = (uint32_t)(rng() % 20);
            }
            const uint32_t ns_cap = 64, limits_cap = 16;
            std::vector<unsigned long long> out(3 * ns_cap + limits_cap + 1, 0);
            emu_ns_metrics(recs.data(), words, n, lim.data(), first.data(), ns_cap, limits_cap, out.data());
            unsigned long long total = out.back();
            for (uint32_t k = 0; k < ns_cap; k++) total += out[k] + out[2 * ns_cap + k];
            if (total != n) return 2;
        }
    }
    // table rebuild: every geometry, fill / tombstone / rebuild cycles
    for (uint32_t cells : {1u, 3u, 7u}) {
        for (uint32_t log2P : {0u, 2u}) {
            emu_table* t = emu_table_create(cells, log2P, 6);
            std::vector<std::pair<uint64_t, uint64_t>> keys;
            std::vector<unsigned long long> c(2 * cells, 5);
            for (int round = 0; round < 4; round++) {
                for (int i = 0; i < 60 << log2P; i++) {
                    const uint64_t lo = rng() | 1, hi = ((1 + rng() % 7) << 32) | (rng() & 0xFFFFFFFF);
                    if (emu_table_put(t, lo, hi, c.data()) >= 0) keys.emplace_back(lo, hi);
                }
                for (size_t i = 0; i < keys.size(); i += 2) emu_table_tombstone(t, keys[i].first, keys[i].second);
                unsigned long long st[7];
                emu_table_compact(t, round % 2 ? 0 : 10, st, nullptr);
                if (st[6]) return 3;
                std::vector<std::pair<uint64_t, uint64_t>> live;
                for (size_t i = 1; i < keys.size(); i += 2) {
                    if (emu_table_get(t, keys[i].first, keys[i].second, nullptr) < 0) return 4;
                    live.push_back(keys[i]);
                }
                keys.swap(live);
            }
            emu_table_destroy(t);
        }
    }
    // replicated value: sessions on small tables, including a table that fills up
    for (uint32_t actors : {1u, 2u, 5u, 16u}) {
        emu_crdt* c = emu_crdt_create(64, actors, actors - 1);
        uint64_t now = 1700000000000000ull;
        for (int round = 0; round < 30; round++) {
            now += rng() % 2000000;
            const uint32_t n = 1 + rng() % 40;
            std::vector<rl_crdt_key> keys(n);
            std::vector<uint32_t> actor(n);
            std::vector<uint64_t> inc(n), win(n);
            for (uint32_t i = 0; i < n; i++) {
                keys[i] = rl_crdt_key{1 + rng() % 90, (1ull << 32) | (rng() % 3)};
                actor[i] = (uint32_t)(rng() % actors);
                inc[i] = 1 + rng() % 9;
                win[i] = (rng() % 3) * 1000000;
            }
            emu_crdt_inc(c, n, keys.data(), actor.data(), inc.data(), win.data(), now);
            std::vector<rl_crdt_update> ups(n);
            std::vector<uint32_t> ua;
            std::vector<uint64_t> uv;
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t nv = (uint32_t)(rng() % (actors + 1));
                ups[i] = rl_crdt_update{1 + rng() % 90, (1ull << 32) | (rng() % 3), now + rng() % 3000000 - 1000000, (uint32_t)ua.size(), nv};
                for (uint32_t j = 0; j < nv; j++) {
                    ua.push_back((uint32_t)(rng() % actors));
                    uv.push_back(rng() % 1000);
                }
            }
            ua.push_back(0);
            uv.push_back(0);
            emu_crdt_merge(c, n, ups.data(), ua.data(), uv.data(), ua.size() - 1, now);
            std::vector<uint64_t> val(n), exp(n);
            emu_crdt_read(c, n, keys.data(), now, val.data(), exp.data());
            std::vector<rl_crdt_key> ok(64);
            std::vector<uint64_t> oa(64), oe(64), ov(64 * actors);
            if (emu_crdt_scan(c, 0, now, 64, ok.data(), oa.data(), oe.data(), nullptr) > 64) return 5;
            if (emu_crdt_scan(c, 1, 0, 64, ok.data(), nullptr, oe.data(), ov.data()) > 64) return 6;
        }
        emu_crdt_destroy(c);
    }
    printf("ok kernels\n");
    return 0;
}
