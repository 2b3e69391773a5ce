// rl_kernels.cuh — hand-written sm_100a kernels of the batched rate-limit engine.
//
// Pipeline for one batch:
//   k_part<COUNT>  : per-tile histogram of accesses over the P table regions
//   k_colscan      : per-region exclusive prefix over tiles + exclusive scan of region totals
//   k_part<SCATTER>: STABLE scatter of access indices into per-region lists (stream order kept)
//   k_main         : one CTA per region; groups the region's accesses by row key in shared
//                    memory, and one walker thread per key replays that key's requests in
//                    stream order against the row (fixed-window check / increment), so the
//                    result equals one-at-a-time execution on the reference InMemoryStorage
//                    (limitador/src/storage/in_memory.rs:72-156).
// A region (contiguous slab of rows) is touched by exactly one CTA per launch, and a key by
// exactly one thread, so counter values need no atomics at all; the only atomic on the table
// is the 128-bit CAS that claims an empty row for a new key.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rl_engine.h"
#include "rl_core.h"

#define RL_MAIN_THREADS 256  // accesses per chunk of k_main (one per thread)
#define RL_PART_THREADS 256
#define RL_PART_WARPS (RL_PART_THREADS / 32)
#define RL_IDENT_POSORIG 0x0654321006543210ull

struct RlDev {
    uint8_t* rows;
    uint32_t log2P;  // regions
    uint32_t log2R;  // rows per region
    const RlCellDesc* desc;  // [ngroups][8]
    const RlLimitDev* limits;
    uint32_t limits_cap;
    const RlNsDev* ns;
    uint32_t ns_cap;
    const uint32_t* ns_limit_ids;
    uint32_t* err;    // sticky max of RL_DEV_*
    uint32_t* flags;  // bit0: batch has multi-row requests
    unsigned long long tag_mask;  // ~0; tests narrow it to force in-CTA tag collisions
};

struct RlBatch {
    uint32_t n_acc;
    uint32_t n_req;
    // partition workspace
    uint32_t* tile_cnt;      // [num_tiles][P+1]; after k_colscan: exclusive prefix over tiles
    uint32_t* region_total;  // [P+1]
    uint32_t* part_base;     // [P+2]
    uint32_t* part_idx;      // [n_acc]
    uint32_t* scan_ctr;      // last-block-done counter of k_colscan
    uint32_t tile;           // accesses per tile (multiple of 256)
    uint32_t num_tiles;
    // outputs (device)
    uint8_t* out_limited;
    uint32_t* out_first_limited;
    uint64_t* out_remaining;
    uint64_t* out_ttl;
    const uint32_t* out_off;  // CSR: per-request base index of remaining/ttl; null => req*out_stride
    uint32_t out_stride;
    // coupled (multi-row) requests
    uint32_t* fl_prev;
    uint32_t* fl_next;
    int phase;          // RL_PHASE_*
    int load_counters;  // 0/1
    // undo log of the rows a coupled batch touches (RL_PHASE_SNAPSHOT / k_restore)
    uint8_t** log_row;     // [n_acc] row pointer logged at the partition position of a key's first access
    ulonglong2* log_state; // [n_acc][CELLS]
};

// k_main phases.  COMMIT: rows and outputs are written (the normal, single pass).
// Coupled batches (requests spanning several rows) run SNAPSHOT once (log the original
// state of every touched row), then SPEC rounds (rows written, outputs not, fl_next
// accumulated; k_restore puts the rows back after each), then COMMIT.
#define RL_PHASE_COMMIT 0
#define RL_PHASE_SPEC 1
#define RL_PHASE_SNAPSHOT 2

// ---------------------------------------------------------------------------------------
// memory helpers: table traffic bypasses L1 (.cg) — rows have no L1 reuse and .cg keeps the
// CTA's own earlier writes (previous chunk) visible without relying on L1 invalidation.
struct __align__(16) RlU128 {
    unsigned long long x, y;
};
__device__ __forceinline__ ulonglong2 rl_ld_cg(const void* p) {
    return __ldcg(reinterpret_cast<const ulonglong2*>(p));
}
__device__ __forceinline__ void rl_st_cg(void* p, unsigned long long a, unsigned long long b) {
    __stcg(reinterpret_cast<ulonglong2*>(p), make_ulonglong2(a, b));
}
// 128-bit compare-and-swap (PTX atom.cas.b128, sm_90+): claims a row header atomically.
__device__ __forceinline__ ulonglong2 rl_cas128(void* addr, ulonglong2 cmp, ulonglong2 val) {
    ulonglong2 old;
    asm volatile(
        "{\n\t"
        ".reg .b128 c, v, o;\n\t"
        "mov.b128 c, {%3, %4};\n\t"
        "mov.b128 v, {%5, %6};\n\t"
        "atom.global.cas.b128 o, [%2], c, v;\n\t"
        "mov.b128 {%0, %1}, o;\n\t"
        "}\n"
        : "=l"(old.x), "=l"(old.y)
        : "l"(addr), "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y)
        : "memory");
    return old;
}
// streaming (read-once) loads of request data: keep them out of L1
__device__ __forceinline__ ulonglong2 rl_ld_stream(const void* p) {
    return __ldcs(reinterpret_cast<const ulonglong2*>(p));
}
__device__ __forceinline__ void rl_set_err(const RlDev& D, uint32_t code) { atomicMax(D.err, code); }

template <int CELLS>
struct RlGeom {
    static constexpr uint32_t ROW_BYTES = 16u * (1 + CELLS);
};

__device__ __forceinline__ uint64_t rl_region_of(const RlDev& D, uint64_t h) {
    return D.log2P ? (h >> (64 - D.log2P)) : 0ull;
}

// Find the row of (key_lo, hdr_hi); optionally claim an empty/tombstoned row for it.
// Linear probing confined to the key's region.  Returns nullptr when absent (and !create)
// or when the region is full (error flagged).
template <int CELLS>
__device__ uint8_t* rl_probe(const RlDev& D, uint64_t h, uint64_t key_lo, uint64_t hdr_hi, bool create) {
    constexpr uint32_t RB = RlGeom<CELLS>::ROW_BYTES;
    const uint32_t R = 1u << D.log2R;
    const uint64_t base = rl_region_of(D, h) << D.log2R;
    const uint32_t idx = (uint32_t)h & (R - 1);
    int tomb = -1;
    uint32_t restarts = 0;
    for (uint32_t i = 0; i < R;) {
        uint8_t* row = D.rows + (base + ((idx + i) & (R - 1))) * RB;
        const ulonglong2 hdr = rl_ld_cg(row);
        if (hdr.x == key_lo && hdr.y == hdr_hi) return row;
        if (hdr.x == 0 && hdr.y == 0) {
            if (!create) return nullptr;
            uint8_t* target = row;
            ulonglong2 expect = make_ulonglong2(0ull, 0ull);
            if (tomb >= 0) {
                target = D.rows + (base + ((idx + (uint32_t)tomb) & (R - 1))) * RB;
                expect = make_ulonglong2(0ull, RL_TOMB_HI);
            }
            const ulonglong2 old = rl_cas128(target, expect, make_ulonglong2(key_lo, hdr_hi));
            if (old.x == expect.x && old.y == expect.y) return target;
            // another walker of this CTA took the row first: rescan
            if (++restarts > 4 * R) break;
            tomb = -1;
            i = 0;
            continue;
        }
        if (hdr.y == RL_TOMB_HI && tomb < 0) tomb = (int)i;
        i++;
    }
    if (create) {
        if (tomb >= 0) {
            uint8_t* target = D.rows + (base + ((idx + (uint32_t)tomb) & (R - 1))) * RB;
            const ulonglong2 old = rl_cas128(target, make_ulonglong2(0ull, RL_TOMB_HI),
                                             make_ulonglong2(key_lo, hdr_hi));
            if (old.x == 0ull && old.y == RL_TOMB_HI) return target;
        }
        rl_set_err(D, RL_DEV_TABLE_FULL);
    }
    return nullptr;
}

template <int CELLS>
__device__ __forceinline__ void rl_row_load(const uint8_t* row, uint32_t ncells, RlRow<CELLS>& r) {
#pragma unroll
    for (int c = 0; c < CELLS; c++) {
        if ((uint32_t)c < ncells && row) {
            const ulonglong2 v = rl_ld_cg(row + 16 + 16 * c);
            r.value[c] = v.x;
            r.expiry[c] = v.y;
        } else {
            r.value[c] = 0;
            r.expiry[c] = 0;
        }
    }
}
template <int CELLS>
__device__ __forceinline__ void rl_row_store(uint8_t* row, uint32_t dirty, const RlRow<CELLS>& r) {
#pragma unroll
    for (int c = 0; c < CELLS; c++)
        if (dirty & (1u << c)) rl_st_cg(row + 16 + 16 * c, r.value[c], r.expiry[c]);
}

// ---------------------------------------------------------------------------------------
// Access sources.  RecordSrc: access == request, derived on the fly from the 32-B record
// and the namespace table (every namespace is single-row).  AccSrc: materialised accesses
// written by a resolve kernel (general CSR form, or records of multi-row namespaces).
struct RecordSrc {
    static constexpr bool kAccessIsRequest = true;
    const rl_record* recs;
    // identity of access a: false => no row (namespace without limits)
    __device__ __forceinline__ bool ident(const RlDev& D, uint32_t a, uint64_t& key_lo, uint64_t& hdr_hi) const {
        const ulonglong2 w0 = rl_ld_stream(&recs[a]);       // ns_id|hits, key_lo
        const uint32_t ns_id = (uint32_t)w0.x;
        if (ns_id >= D.ns_cap) return false;
        const RlNsDev ns = D.ns[ns_id];
        if (ns.mode != 1) return false;
        if (ns.qualified_row) {
            const unsigned long long key_hi = __ldcs(reinterpret_cast<const unsigned long long*>(&recs[a]) + 2);
            if (key_hi >> 32) {
                rl_set_err(D, RL_DEV_KEY_RANGE);
                return false;
            }
            key_lo = w0.y;
            hdr_hi = ((uint64_t)ns.group << 32) | key_hi;
        } else {
            key_lo = 0;
            hdr_hi = (uint64_t)ns.group << 32;
        }
        return true;
    }
    __device__ __forceinline__ void full(const RlDev& D, uint32_t a, RlAccess& acc, uint64_t& delta,
                                         uint64_t& now) const {
        const ulonglong2 w0 = rl_ld_stream(&recs[a]);
        const ulonglong2 w1 = rl_ld_stream(reinterpret_cast<const ulonglong2*>(&recs[a]) + 1);
        const uint32_t ns_id = (uint32_t)w0.x;
        const RlNsDev ns = D.ns[ns_id];
        acc.key_lo = ns.qualified_row ? w0.y : 0;
        acc.hdr_hi = ((uint64_t)ns.group << 32) | (ns.qualified_row ? w1.x : 0);
        acc.req = a;
        acc.cells = ns.cells;
        acc.posorig = RL_IDENT_POSORIG;
        delta = (uint64_t)(w0.x >> 32);
        now = w1.y;
    }
};

struct AccSrc {
    static constexpr bool kAccessIsRequest = false;
    const RlAccess* acc;
    const uint64_t* delta;  // per request
    const uint64_t* now;    // per request
    __device__ __forceinline__ bool ident(const RlDev&, uint32_t a, uint64_t& key_lo, uint64_t& hdr_hi) const {
        const ulonglong2 w0 = rl_ld_stream(&acc[a]);
        key_lo = w0.x;
        hdr_hi = w0.y;
        return hdr_hi != 0;
    }
    __device__ __forceinline__ void full(const RlDev&, uint32_t a, RlAccess& out, uint64_t& d,
                                         uint64_t& t) const {
        const ulonglong2 w0 = rl_ld_stream(&acc[a]);
        const ulonglong2 w1 = rl_ld_stream(reinterpret_cast<const ulonglong2*>(&acc[a]) + 1);
        out.key_lo = w0.x;
        out.hdr_hi = w0.y;
        out.req = (uint32_t)w1.x;
        out.cells = (uint32_t)(w1.x >> 32);
        out.posorig = w1.y;
        d = delta[out.req];
        t = now[out.req];
    }
};

// ---------------------------------------------------------------------------------------
// Stable partition of the accesses by table region.
// Tile = B.tile consecutive accesses; warp w of the CTA owns the w-th contiguous slice, so
// stream order == (tile, warp, step, lane) and a stable rank is
//   region base + (accesses of earlier tiles) + (accesses of earlier warps) + rank in slice.
template <class Src, bool SCATTER>
__global__ void __launch_bounds__(RL_PART_THREADS) k_part(RlDev D, RlBatch B, Src src) {
    extern __shared__ uint32_t wcnt[];  // [RL_PART_WARPS][P+1]
    const uint32_t P1 = (1u << D.log2P) + 1;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t t0 = tile * B.tile;
    const uint32_t t1 = min(t0 + B.tile, B.n_acc);
    const uint32_t slice = B.tile / RL_PART_WARPS;
    const uint32_t s0 = min(t0 + warp * slice, t1);
    const uint32_t s1 = min(s0 + slice, t1);
    uint32_t* mycnt = wcnt + warp * P1;

    for (uint32_t i = tid; i < RL_PART_WARPS * P1; i += RL_PART_THREADS) wcnt[i] = 0;
    __syncthreads();

    // pass 1: per-warp region counts
    for (uint32_t b = s0; b < s1; b += 32) {
        const uint32_t a = b + lane;
        const bool valid = a < s1;
        uint32_t r = P1 - 1;  // dummy region for accesses without a row
        if (valid) {
            uint64_t klo, hhi;
            if (src.ident(D, a, klo, hhi)) r = (uint32_t)rl_region_of(D, rl_row_hash(klo, hhi));
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned m = __match_any_sync(vmask, r);
            if (lane == (uint32_t)(__ffs(m) - 1)) mycnt[r] += __popc(m);
        }
        __syncwarp();
    }
    __syncthreads();

    if (!SCATTER) {
        for (uint32_t r = tid; r < P1; r += RL_PART_THREADS) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < RL_PART_WARPS; w++) tot += wcnt[w * P1 + r];
            B.tile_cnt[(size_t)tile * P1 + r] = tot;
        }
        return;
    }

    // bases: region base + earlier tiles + earlier warps of this tile
    for (uint32_t r = tid; r < P1; r += RL_PART_THREADS) {
        uint32_t run = B.part_base[r] + B.tile_cnt[(size_t)tile * P1 + r];
#pragma unroll
        for (int w = 0; w < RL_PART_WARPS; w++) {
            const uint32_t c = wcnt[w * P1 + r];
            wcnt[w * P1 + r] = run;
            run += c;
        }
    }
    __syncthreads();

    // pass 2: same sweep, now handing out positions
    for (uint32_t b = s0; b < s1; b += 32) {
        const uint32_t a = b + lane;
        const bool valid = a < s1;
        uint32_t r = P1 - 1;
        if (valid) {
            uint64_t klo, hhi;
            if (src.ident(D, a, klo, hhi)) r = (uint32_t)rl_region_of(D, rl_row_hash(klo, hhi));
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned m = __match_any_sync(vmask, r);
            const int leader = __ffs(m) - 1;
            uint32_t basepos = 0;
            if ((int)lane == leader) {
                basepos = mycnt[r];
                mycnt[r] = basepos + __popc(m);
            }
            basepos = __shfl_sync(m, basepos, leader);
            B.part_idx[basepos + __popc(m & ((1u << lane) - 1))] = a;
            if (Src::kAccessIsRequest && r == P1 - 1 && B.out_limited) {
                // request without any applicable limit: not limited (lib.rs:434-440)
                B.out_limited[a] = 0;
                if (B.out_first_limited) B.out_first_limited[a] = RL_NONE_U32;
            }
        }
        __syncwarp();
    }
}

// Column scan of tile_cnt: CTA c owns regions [32c, 32c+32); warp w a slice of the tiles.
// The last CTA to finish turns region_total into part_base (exclusive scan, P+2 entries).
__global__ void __launch_bounds__(256) k_colscan(RlDev D, RlBatch B) {
    __shared__ uint32_t part[8][32];
    __shared__ uint32_t s_last;
    __shared__ uint32_t s_warp[8];
    const uint32_t P1 = (1u << D.log2P) + 1;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * 32 + lane;
    const uint32_t nt = B.num_tiles;
    const uint32_t per = (nt + 7) / 8;
    const uint32_t a0 = min(warp * per, nt), a1 = min(a0 + per, nt);
    uint32_t sum = 0;
    if (r < P1)
        for (uint32_t t = a0; t < a1; t++) sum += B.tile_cnt[(size_t)t * P1 + r];
    part[warp][lane] = sum;
    __syncthreads();
    uint32_t run = 0;
    for (uint32_t w = 0; w < warp; w++) run += part[w][lane];
    if (r < P1) {
        for (uint32_t t = a0; t < a1; t++) {
            const uint32_t c = B.tile_cnt[(size_t)t * P1 + r];
            B.tile_cnt[(size_t)t * P1 + r] = run;
            run += c;
        }
        if (warp == 7) B.region_total[r] = run;  // warp 7 ends with the full column sum
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(B.scan_ctr, 1) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // exclusive scan over region_total[0..P1) -> part_base[0..P1]
    uint32_t carry = 0;
    for (uint32_t base = 0; base < P1; base += 256) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < P1) ? __ldcg(&B.region_total[i]) : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        uint32_t woff = 0;
        for (uint32_t w = 0; w < warp; w++) woff += s_warp[w];
        uint32_t total = 0;
        for (uint32_t w = 0; w < 8; w++) total += s_warp[w];
        if (i < P1) B.part_base[i] = carry + woff + x - v;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        B.part_base[P1] = carry;
        *B.scan_ctr = 0;  // re-arm for the next batch
    }
}

// ---------------------------------------------------------------------------------------
// The main kernel.  MODE 0 = check_and_update, MODE 2 = update_counters.
//
// One CTA per table region; the region's accesses (already in stream order) are taken in
// chunks of CH (one access per thread).  Per chunk:
//   1. every thread loads its access, and the CTA groups the accesses by row key with a
//      shared-memory hash table (64-bit tag claimed by CAS, full key verified against the
//      claimer's; a mismatch re-inserts under a salted tag);
//   2. every access gets its stable ordinal inside its key group (warps take turns, so the
//      ordinal order is the stream order) and the inclusive prefix sum of the group's
//      deltas (one block scan over the group-sorted deltas);
//   3. the group leader probes / claims the row (one probe per key) and stages the row
//      state in shared memory;
//   4. the group is replayed by ALL its threads in lock-step "run-length" rounds
//      (rl_core.h: hypotheses A and B) — two barriers per round, one round for a saturated
//      or an unconstrained hot key;
//   5. the leader writes the dirty cells back.
// Counter values never need atomics: a region belongs to one CTA, a key to one group.
template <int CELLS, int CH>
struct RlMainSmem {
    static constexpr int GT = 2 * CH;
    unsigned long long g_tag[GT];
    unsigned long long key_lo[CH];
    unsigned long long key_hi[CH];
    unsigned long long s_val[CH * CELLS];  // row state of the group led by thread `gid`
    unsigned long long s_exp[CH * CELLS];
    unsigned long long dsum[CH];           // group-sorted deltas -> inclusive scan
    unsigned long long g_pbase[CH];        // prefix sum of the deltas of the finalised members
    uint32_t g_cnt[GT];
    uint32_t g_start[GT];
    uint32_t g_rep[GT];
    uint32_t g_min[2][2][CH];              // [round parity][A|B][gid]
    uint32_t g_dirty[CH];
    uint32_t g_rowok[CH];
    unsigned long long warp_tot[CH / 32];
};

template <int CELLS, class Src, int MODE, int CH>
__global__ void __launch_bounds__(CH) k_main(RlDev D, RlBatch B, Src src) {
    using Smem = RlMainSmem<CELLS, CH>;
    constexpr int GT = Smem::GT;
    extern __shared__ __align__(16) unsigned char rl_smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(rl_smem_raw);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t P = 1u << D.log2P;
    const bool lc = B.load_counters != 0;
    const bool write_out = (B.phase == RL_PHASE_COMMIT);
    const bool snapshot = (B.phase == RL_PHASE_SNAPSHOT);

    for (uint32_t region = blockIdx.x; region < P; region += gridDim.x) {
        const uint32_t lo = B.part_base[region], hi = B.part_base[region + 1];
        for (uint32_t c0 = lo; c0 < hi; c0 += CH) {
            for (uint32_t i = tid; i < GT; i += CH) {
                sm.g_tag[i] = 0ull;
                sm.g_cnt[i] = 0;
            }
            sm.dsum[tid] = 0ull;
            __syncthreads();

            // ---- 1. load my access, group by key --------------------------------------------
            const uint32_t p = c0 + tid;
            const bool valid = p < hi;
            RlAccess acc;
            acc.key_lo = 0;
            acc.hdr_hi = 0;
            acc.req = 0;
            acc.cells = 0;
            acc.posorig = 0;
            uint64_t delta = 0, now = 0, h = 0;
            if (valid) {
                src.full(D, B.part_idx[p], acc, delta, now);
                h = rl_row_hash(acc.key_lo, acc.hdr_hi);
            }
            sm.key_lo[tid] = acc.key_lo;
            sm.key_hi[tid] = acc.hdr_hi;
            uint32_t slot = 0;
            {
                bool pending = valid;
                uint32_t salt = 0;
                for (;;) {
                    if (pending) {
                        // 56 hash bits + the salt level in the top byte: keys that collided on one
                        // level meet fresh tags on the next, so every level places >= 1 key
                        unsigned long long tag = (salt ? rl_mix64(h + salt) : h) & D.tag_mask;
                        tag = (tag & 0x00FFFFFFFFFFFFFFull) | ((unsigned long long)(salt & 0xFFu) << 56) | 1ull;
                        uint32_t s = (uint32_t)(tag >> 24) & (GT - 1);
                        for (;;) {
                            const unsigned long long old = atomicCAS(&sm.g_tag[s], 0ull, tag);
                            if (old == 0ull) {
                                sm.g_rep[s] = tid;  // I claimed the slot: my key defines the group
                                break;
                            }
                            if (old == tag) break;
                            s = (s + 1) & (GT - 1);
                        }
                        slot = s;
                    }
                    __syncthreads();
                    if (pending) {
                        const uint32_t rep = sm.g_rep[slot];
                        pending = (sm.key_lo[rep] != acc.key_lo) || (sm.key_hi[rep] != acc.hdr_hi);
                        salt++;
                    }
                    if (!__syncthreads_or(pending)) break;  // a tag collision between different keys: re-insert salted
                }
            }

            // ---- 2. stable ordinal inside the key group: warps take turns ---------------------
            uint32_t ord = 0;
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            for (uint32_t w = 0; w < CH / 32; w++) {
                if (warp == w && valid) {
                    const unsigned m = __match_any_sync(vmask, slot);
                    const int leader = __ffs(m) - 1;
                    uint32_t basecnt = 0;
                    if ((int)lane == leader) {
                        basecnt = sm.g_cnt[slot];
                        sm.g_cnt[slot] = basecnt + __popc(m);
                    }
                    basecnt = __shfl_sync(m, basecnt, leader);
                    ord = basecnt + __popc(m & ((1u << lane) - 1));
                }
                __syncthreads();
            }
            // exclusive scan of the group sizes -> where each group starts in group-sorted order
            {
                uint32_t run = 0;
                // GT = 2*CH entries, two per thread
                const uint32_t v0 = sm.g_cnt[2 * tid], v1 = sm.g_cnt[2 * tid + 1];
                uint32_t x = v0 + v1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                    if ((int)lane >= o) x += y;
                }
                if (lane == 31) sm.warp_tot[warp] = x;
                __syncthreads();
                for (uint32_t w = 0; w < warp; w++) run += (uint32_t)sm.warp_tot[w];
                const uint32_t excl = run + x - (v0 + v1);
                sm.g_start[2 * tid] = excl;
                sm.g_start[2 * tid + 1] = excl + v0;
                __syncthreads();
            }
            const uint32_t cnt = valid ? sm.g_cnt[slot] : 0;
            const uint32_t gstart = valid ? sm.g_start[slot] : 0;
            const uint32_t q = gstart + ord;  // my position in group-sorted order
            if (valid) sm.dsum[q] = delta;
            __syncthreads();
            // inclusive scan of the group-sorted deltas; P_i = X[q] - X[gstart-1]
            {
                unsigned long long x = sm.dsum[tid];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
                    if ((int)lane >= o) x += y;
                }
                if (lane == 31) sm.warp_tot[warp] = x;
                __syncthreads();
                unsigned long long run = 0;
                for (uint32_t w = 0; w < warp; w++) run += sm.warp_tot[w];
                sm.dsum[tid] = run + x;
                __syncthreads();
            }
            const unsigned long long p_incl = valid ? sm.dsum[q] - (gstart ? sm.dsum[gstart - 1] : 0ull) : 0ull;
            const unsigned long long p_prev = p_incl - delta;

            // ---- 3. leaders probe the row and stage its state ---------------------------------
            const uint32_t gid = gstart;  // unique per group, < CH
            const uint32_t group = (uint32_t)(acc.hdr_hi >> 32);
            const RlCellDesc* desc = D.desc + (size_t)group * 8;
            uint8_t* row = nullptr;
            const bool is_leader = valid && ord == 0;
            const uint32_t ucells = acc.cells;  // meaningful on the leader; shared below
            if (is_leader) {
                row = rl_probe<CELLS>(D, h, acc.key_lo, acc.hdr_hi, true);
                RlRow<CELLS> st;
                rl_row_load<CELLS>(row, CELLS, st);
#pragma unroll
                for (int c = 0; c < CELLS; c++) {
                    sm.s_val[gid * CELLS + c] = st.value[c];
                    sm.s_exp[gid * CELLS + c] = st.expiry[c];
                }
                sm.g_min[0][0][gid] = sm.g_min[0][1][gid] = 0xFFFFFFFFu;
                sm.g_min[1][0][gid] = sm.g_min[1][1][gid] = 0xFFFFFFFFu;
                sm.g_pbase[gid] = 0ull;
                sm.g_dirty[gid] = 0;
                sm.g_rowok[gid] = (row != nullptr);
                sm.g_rep[slot] = ucells;  // g_rep is free now: publish the leader's cell list
                if (snapshot && row != nullptr) {
                    B.log_row[p] = row;
#pragma unroll
                    for (int c = 0; c < CELLS; c++)
                        B.log_state[(size_t)p * CELLS + c] = make_ulonglong2(st.value[c], st.expiry[c]);
                }
            }
            __syncthreads();

            // ---- 4. lock-step run-length replay -------------------------------------------------
            bool done = !valid || snapshot || !sm.g_rowok[gid];
            const uint32_t lead_cells = valid ? sm.g_rep[slot] : 0;
            const bool multi = rl_cells_multi(acc.cells);
            uint32_t pos = 0;
            for (uint32_t round = 0;; round++) {
                const uint32_t par = round & 1;
                RlRow<CELLS> st;
                if (!done) {
#pragma unroll
                    for (int c = 0; c < CELLS; c++) {
                        st.value[c] = sm.s_val[gid * CELLS + c];
                        st.expiry[c] = sm.s_exp[gid * CELLS + c];
                    }
                    bool aok = false, bok = false;
                    if (!multi) {
                        if (MODE == 0) {
                            aok = rl_eval_deny_noeffect<CELLS>(st, desc, acc.cells, acc.posorig, delta, now, lc);
                            bok = (acc.cells == lead_cells) &&
                                  rl_eval_allow_run<CELLS>(st, desc, acc.cells, p_incl - sm.g_pbase[gid], now);
                        } else {
                            bok = (acc.cells == lead_cells) && rl_eval_update_run<CELLS>(st, acc.cells, now);
                        }
                    }
                    if (!aok) atomicMin(&sm.g_min[par][0][gid], ord);
                    if (!bok) atomicMin(&sm.g_min[par][1][gid], ord);
                }
                __syncthreads();
                if (!done) {
                    const uint32_t mA = min(sm.g_min[par][0][gid], cnt);
                    const uint32_t mB = min(sm.g_min[par][1][gid], cnt);
                    const unsigned long long pbase = sm.g_pbase[gid];
                    uint32_t newpos;
                    bool mine = false;     // am I finalised this round?
                    bool store = false;    // do I publish the new row state?
                    uint32_t dirty = 0;
                    uint32_t fl = RL_NONE_U32;
                    uint64_t* rem = nullptr;
                    uint64_t* ttl = nullptr;
                    if (MODE == 0 && lc && write_out) {
                        const size_t ob = B.out_off ? (size_t)B.out_off[acc.req] : (size_t)acc.req * B.out_stride;
                        if (B.out_remaining) rem = B.out_remaining + ob;
                        if (B.out_ttl) ttl = B.out_ttl + ob;
                    }
                    if (mA > pos) {  // run of denied requests: state untouched
                        newpos = mA;
                        if (ord < mA) {
                            mine = true;
                            fl = rl_walk_check_single<CELLS>(st, dirty, desc, acc.cells, acc.posorig, delta, now, lc, rem, ttl);
                        }
                    } else if (mB > pos) {  // run of allowed requests: values accumulate
                        newpos = mB;
                        if (ord < mB) {
                            mine = true;
                            rl_advance_run<CELLS>(st, acc.cells, p_prev - pbase);
                            if (MODE == 0)
                                fl = rl_walk_check_single<CELLS>(st, dirty, desc, acc.cells, acc.posorig, delta, now, lc, rem, ttl);
                            else
                                rl_walk_update<CELLS>(st, dirty, desc, acc.cells, delta, now);
                            store = (ord == mB - 1);
                            if (store) dirty = 0;
                            if (store) {
                                const uint32_t n = rl_cells_n(acc.cells);
                                for (uint32_t k = 0; k < n; k++) dirty |= 1u << rl_cells_at(acc.cells, k);
                            }
                        }
                    } else {  // the request at `pos` is applied alone, sequential rule
                        newpos = pos + 1;
                        if (ord == pos) {
                            mine = true;
                            store = true;
                            if (MODE == 2) {
                                rl_walk_update<CELLS>(st, dirty, desc, acc.cells, delta, now);
                            } else if (!multi) {
                                fl = rl_walk_check_single<CELLS>(st, dirty, desc, acc.cells, acc.posorig, delta, now, lc, rem, ttl);
                            } else {
                                const uint32_t fl_in = B.fl_prev[acc.req];
                                const uint32_t local = rl_walk_check_multi<CELLS>(st, dirty, desc, acc.cells, acc.posorig,
                                                                                 delta, now, lc, fl_in, rem, ttl);
                                if (!write_out && local != RL_NONE_U32) atomicMin(&B.fl_next[acc.req], local);
                                fl = fl_in;
                            }
                        }
                    }
                    if (mine) {
                        done = true;
                        if (MODE == 0 && write_out) {
                            B.out_limited[acc.req] = (fl != RL_NONE_U32);
                            if (B.out_first_limited) {
                                if (fl == RL_NONE_U32) {
                                    B.out_first_limited[acc.req] = RL_NONE_U32;
                                } else {
                                    // the access holding position fl names the limit
                                    const uint32_t n = rl_cells_n(acc.cells);
                                    for (uint32_t k = 0; k < n; k++)
                                        if (rl_pos_at(acc.posorig, k) == fl)
                                            B.out_first_limited[acc.req] = desc[rl_cells_at(acc.cells, k)].limit_id;
                                }
                            }
                        }
                        if (store && dirty) {
#pragma unroll
                            for (int c = 0; c < CELLS; c++)
                                if (dirty & (1u << c)) {
                                    sm.s_val[gid * CELLS + c] = st.value[c];
                                    sm.s_exp[gid * CELLS + c] = st.expiry[c];
                                }
                            atomicOr(&sm.g_dirty[gid], dirty);
                        }
                        if (ord == newpos - 1) {  // last finalised member re-arms the group
                            sm.g_pbase[gid] = p_incl;
                            sm.g_min[par ^ 1][0][gid] = 0xFFFFFFFFu;
                            sm.g_min[par ^ 1][1][gid] = 0xFFFFFFFFu;
                        }
                    }
                    pos = newpos;
                }
                if (!__syncthreads_or(!done)) break;
            }

            // ---- 5. write the dirty cells back ---------------------------------------------------
            if (is_leader && row != nullptr && !snapshot) {
                const uint32_t dirty = sm.g_dirty[gid];
#pragma unroll
                for (int c = 0; c < CELLS; c++)
                    if (dirty & (1u << c))
                        rl_st_cg(row + 16 + 16 * c, sm.s_val[gid * CELLS + c], sm.s_exp[gid * CELLS + c]);
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------
// Resolve kernels: request -> accesses (rl_resolve_request in rl_core.h).
struct RlResolveOut {
    RlAccess* acc;
    uint64_t* delta;  // records only: materialised per request
    uint64_t* now;
    uint8_t* out_limited;         // defaults for requests without counters
    uint32_t* out_first_limited;  // nullable
};

__global__ void k_resolve_csr(RlDev D, uint32_t n, const uint32_t* __restrict__ off,
                              const rl_counter* __restrict__ ctrs, RlResolveOut O, int write_defaults) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o0 = off[i], m = off[i + 1] - o0;
    if (m == 0) {
        if (write_defaults) {
            O.out_limited[i] = 0;
            if (O.out_first_limited) O.out_first_limited[i] = RL_NONE_U32;
        }
        return;
    }
    auto get = [&](uint32_t j) {
        const rl_counter c = ctrs[o0 + j];
        RlCtrIn r;
        r.limit_id = c.limit_id;
        r.key_lo = c.key_lo;
        r.key_hi = c.key_hi;
        return r;
    };
    RlAccess tmp[RL_MAX_CTRS_PER_REQ];
    const int nacc = rl_resolve_request(i, m, get, D.limits, D.limits_cap, true, tmp);
    if (nacc < 0) {
        rl_set_err(D, (uint32_t)(-nacc));
        for (uint32_t x = 0; x < m && x < RL_MAX_CTRS_PER_REQ; x++) {
            RlAccess z;
            z.key_lo = 0;
            z.hdr_hi = 0;
            z.req = i;
            z.cells = 0;
            z.posorig = 0;
            O.acc[o0 + x] = z;
        }
        // slots beyond RL_MAX_CTRS_PER_REQ (too-many-counters error) are cleared too
        for (uint32_t x = RL_MAX_CTRS_PER_REQ; x < m; x++) {
            RlAccess z;
            z.key_lo = 0;
            z.hdr_hi = 0;
            z.req = i;
            z.cells = 0;
            z.posorig = 0;
            O.acc[o0 + x] = z;
        }
        if (write_defaults) O.out_limited[i] = 0;
        return;
    }
    for (uint32_t x = 0; x < m; x++) O.acc[o0 + x] = tmp[x];
    if (nacc > 1) atomicOr(D.flags, 1u);
}

// Records whose namespaces span several rows: slot base = i * stride.
__global__ void k_resolve_records(RlDev D, uint32_t n, const rl_record* __restrict__ recs, uint32_t stride,
                                  RlResolveOut O, int write_defaults) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rl_record rec = recs[i];
    O.delta[i] = rec.hits_addend;
    O.now[i] = rec.now_us;
    uint32_t m = 0, lim_off = 0;
    if (rec.ns_id < D.ns_cap) {
        const RlNsDev ns = D.ns[rec.ns_id];
        m = ns.lim_cnt;
        lim_off = ns.lim_off;
    }
    const size_t o0 = (size_t)i * stride;
    RlAccess z;
    z.key_lo = 0;
    z.hdr_hi = 0;
    z.req = i;
    z.cells = 0;
    z.posorig = 0;
    if (m == 0) {
        for (uint32_t x = 0; x < stride; x++) O.acc[o0 + x] = z;
        if (write_defaults) {
            O.out_limited[i] = 0;
            if (O.out_first_limited) O.out_first_limited[i] = RL_NONE_U32;
        }
        return;
    }
    auto get = [&](uint32_t j) {
        RlCtrIn r;
        r.limit_id = D.ns_limit_ids[lim_off + j];
        r.key_lo = rec.key_lo;
        r.key_hi = rec.key_hi;
        return r;
    };
    RlAccess tmp[RL_MAX_CTRS_PER_REQ];
    const int nacc = rl_resolve_request(i, m, get, D.limits, D.limits_cap, true, tmp);
    if (nacc < 0) {
        rl_set_err(D, (uint32_t)(-nacc));
        for (uint32_t x = 0; x < stride; x++) O.acc[o0 + x] = z;
        if (write_defaults) O.out_limited[i] = 0;
        return;
    }
    for (uint32_t x = 0; x < stride; x++) O.acc[o0 + x] = (x < m) ? tmp[x] : z;
    if (nacc > 1) atomicOr(D.flags, 1u);
}

// Undo a speculative round: put every logged row back to its state at batch start.
template <int CELLS>
__global__ void k_restore(uint32_t n_acc, uint8_t* const* __restrict__ log_row,
                          const ulonglong2* __restrict__ log_state) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_acc) return;
    uint8_t* row = log_row[p];
    if (!row) return;
#pragma unroll
    for (int c = 0; c < CELLS; c++) {
        const ulonglong2 v = log_state[(size_t)p * CELLS + c];
        rl_st_cg(row + 16 + 16 * c, v.x, v.y);
    }
}

// fixed-point bookkeeping: changed |= (prev != next); prev = next; next = NONE
__global__ void k_fl_step(uint32_t n, uint32_t* fl_prev, uint32_t* fl_next, uint32_t* changed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = fl_prev[i], x = fl_next[i];
    if (p != x) {
        fl_prev[i] = x;
        atomicOr(changed, 1u);
    }
    fl_next[i] = RL_NONE_U32;
}

// ---------------------------------------------------------------------------------------
// is_rate_limited (lib.rs:362-409 over in_memory.rs:20-35): read-only, one thread per
// request, counters in the given order, no insert.
template <int CELLS>
__device__ __forceinline__ bool rl_query_counter(const RlDev& D, uint32_t limit_id, uint64_t key_lo,
                                                 uint64_t key_hi, uint64_t delta, uint64_t now, bool& err) {
    if (limit_id >= D.limits_cap || D.limits[limit_id].group == 0) {
        rl_set_err(D, RL_DEV_UNKNOWN_LIMIT);
        err = true;
        return false;
    }
    const RlLimitDev l = D.limits[limit_id];
    if (l.qualified && (key_hi >> 32)) {
        rl_set_err(D, RL_DEV_KEY_RANGE);
        err = true;
        return false;
    }
    const uint64_t klo = l.qualified ? key_lo : 0;
    const uint64_t hhi = ((uint64_t)l.group << 32) | (l.qualified ? key_hi : 0);
    const uint8_t* row = rl_probe<CELLS>(D, rl_row_hash(klo, hhi), klo, hhi, false);
    uint64_t v = 0;
    if (row) {
        const ulonglong2 c = rl_ld_cg(row + 16 + 16 * l.cell);
        v = rl_value_at(c.x, c.y, now);
    }
    const RlCellDesc d = D.desc[(size_t)l.group * 8 + l.cell];
    return !(d.max_value >= v + delta);  // in_memory.rs:34
}

template <int CELLS>
__global__ void k_query_csr(RlDev D, uint32_t n, const uint32_t* __restrict__ off,
                            const rl_counter* __restrict__ ctrs, const uint64_t* __restrict__ delta,
                            const uint64_t* __restrict__ now, uint8_t* out_limited, uint32_t* out_first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t first = RL_NONE_U32;
    const uint64_t d = delta[i], t = now[i];
    for (uint32_t j = off[i]; j < off[i + 1]; j++) {
        const rl_counter c = ctrs[j];
        bool err = false;
        if (rl_query_counter<CELLS>(D, c.limit_id, c.key_lo, c.key_hi, d, t, err)) {
            first = c.limit_id;
            break;
        }
        if (err) break;
    }
    out_limited[i] = (first != RL_NONE_U32);
    if (out_first) out_first[i] = first;
}

template <int CELLS>
__global__ void k_query_records(RlDev D, uint32_t n, const rl_record* __restrict__ recs, uint8_t* out_limited,
                                uint32_t* out_first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rl_record rec = recs[i];
    uint32_t first = RL_NONE_U32;
    if (rec.ns_id < D.ns_cap) {
        const RlNsDev ns = D.ns[rec.ns_id];
        for (uint32_t j = 0; j < ns.lim_cnt; j++) {
            const uint32_t lid = D.ns_limit_ids[ns.lim_off + j];
            bool err = false;
            if (rl_query_counter<CELLS>(D, lid, rec.key_lo, rec.key_hi, rec.hits_addend, rec.now_us, err)) {
                first = lid;
                break;
            }
            if (err) break;
        }
    }
    out_limited[i] = (first != RL_NONE_U32);
    if (out_first) out_first[i] = first;
}

// ---------------------------------------------------------------------------------------
// Maintenance kernels: streaming passes over the slab, one thread per row.
struct RlScanOut {
    uint32_t* limit_id;
    uint64_t* key_lo;
    uint64_t* key_hi;
    uint64_t* a;  // dump: value      | get_counters: remaining
    uint64_t* b;  // dump: expiry_us  | get_counters: ttl_us
    unsigned long long* count;
    uint64_t cap;
};

// mode 0: dump every present cell; mode 1: get_counters (ns_sel[ns]!=0, ttl>0)
template <int CELLS>
__global__ void k_scan(RlDev D, uint64_t nrows, int mode, uint64_t now, const uint8_t* __restrict__ ns_sel,
                       const uint32_t* __restrict__ group_ns, RlScanOut O) {
    constexpr uint32_t RB = RlGeom<CELLS>::ROW_BYTES;
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const uint8_t* row = D.rows + r * RB;
    const ulonglong2 hdr = rl_ld_cg(row);
    if (hdr.y == 0 || hdr.y == RL_TOMB_HI) return;
    const uint32_t group = (uint32_t)(hdr.y >> 32);
    if (mode == 1 && !ns_sel[group_ns[group]]) return;
    const RlCellDesc* desc = D.desc + (size_t)group * 8;
#pragma unroll
    for (int c = 0; c < CELLS; c++) {
        const RlCellDesc d = desc[c];
        if (d.limit_id == RL_NONE_U32) continue;
        const ulonglong2 cell = rl_ld_cg(row + 16 + 16 * c);
        if (d.qualified && cell.y == 0) continue;  // logically absent
        uint64_t a = cell.x, b = cell.y;
        if (mode == 1) {
            const uint64_t ttl = rl_ttl(cell.y, now);
            if (ttl == 0) continue;                               // in_memory.rs:167-169
            a = d.max_value - rl_value_at(cell.x, cell.y, now);   // wrapping, :164-165
            b = ttl;
        }
        const unsigned long long pos = atomicAdd(O.count, 1ull);
        if (pos < O.cap) {
            O.limit_id[pos] = d.limit_id;
            O.key_lo[pos] = hdr.x;
            O.key_hi[pos] = hdr.y & 0xFFFFFFFFull;
            O.a[pos] = a;
            O.b[pos] = b;
        }
    }
}

// delete_counters (in_memory.rs:241-257): reset the cells of the selected limits.
// sweep (mode 1): invalidate qualified cells with 0 < expiry <= now; rows left with no live
// cell and no unqualified cell become tombstones.
template <int CELLS>
__global__ void k_reset(RlDev D, uint64_t nrows, int mode, uint64_t now, const uint8_t* __restrict__ limit_sel,
                        unsigned long long* count) {
    constexpr uint32_t RB = RlGeom<CELLS>::ROW_BYTES;
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    uint8_t* row = D.rows + r * RB;
    const ulonglong2 hdr = rl_ld_cg(row);
    if (hdr.y == 0 || hdr.y == RL_TOMB_HI) return;
    const uint32_t group = (uint32_t)(hdr.y >> 32);
    const RlCellDesc* desc = D.desc + (size_t)group * 8;
    bool any_live = false, any_unq = false;
    uint32_t dropped = 0;
#pragma unroll
    for (int c = 0; c < CELLS; c++) {
        const RlCellDesc d = desc[c];
        const ulonglong2 cell = rl_ld_cg(row + 16 + 16 * c);
        bool kill = false;
        if (mode == 0) {
            kill = d.limit_id != RL_NONE_U32 && limit_sel[d.limit_id] && (cell.x != 0 || cell.y != 0);
        } else {
            kill = d.limit_id != RL_NONE_U32 && d.qualified && cell.y != 0 && cell.y <= now;
        }
        if (kill) {
            rl_st_cg(row + 16 + 16 * c, 0ull, 0ull);
            dropped++;
        } else if (d.limit_id != RL_NONE_U32) {
            if (!d.qualified) any_unq = true;
            else if (cell.y != 0) any_live = true;
        } else if (cell.x != 0 || cell.y != 0) {
            rl_st_cg(row + 16 + 16 * c, 0ull, 0ull);  // cell of a forgotten limit
        }
    }
    if (mode == 1 && !any_live && !any_unq) rl_st_cg(row, 0ull, RL_TOMB_HI);
    if (dropped && count) atomicAdd(count, (unsigned long long)dropped);
}

// ---------------------------------------------------------------------------------------
// Multi-GPU exchange helpers: stable bucketing of records by owner rank (world <= 32).
// Same tile/warp-slice scheme as k_part, with the owner as the bucket.
__device__ __forceinline__ uint32_t rl_owner_dev(uint32_t ns_id, uint32_t world) {
    return (uint32_t)(rl_mix64((uint64_t)ns_id + 0x51ed270b0a1fULL) % world);
}

template <bool SCATTER>
__global__ void __launch_bounds__(RL_PART_THREADS) k_bucket(const rl_record* __restrict__ recs, uint32_t n,
                                                           uint32_t world, uint32_t tile, uint32_t* tile_cnt,
                                                           const uint32_t* __restrict__ owner_base,
                                                           rl_record* out_recs, uint32_t* out_src) {
    __shared__ uint32_t wcnt[RL_PART_WARPS][32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t t0 = blockIdx.x * tile, t1 = min(t0 + tile, n);
    const uint32_t slice = tile / RL_PART_WARPS;
    const uint32_t s0 = min(t0 + warp * slice, t1), s1 = min(s0 + slice, t1);
    wcnt[warp][lane] = 0;
    __syncthreads();
    for (uint32_t b = s0; b < s1; b += 32) {
        const uint32_t a = b + lane;
        const bool valid = a < s1;
        const uint32_t o = valid ? rl_owner_dev(recs[a].ns_id, world) : 0;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned m = __match_any_sync(vmask, o);
            if (lane == (uint32_t)(__ffs(m) - 1)) wcnt[warp][o] += __popc(m);
        }
        __syncwarp();
    }
    __syncthreads();
    if (!SCATTER) {
        if (tid < world) {
            uint32_t tot = 0;
            for (int w = 0; w < RL_PART_WARPS; w++) tot += wcnt[w][tid];
            tile_cnt[blockIdx.x * 32 + tid] = tot;
        }
        return;
    }
    if (tid < world) {
        uint32_t run = owner_base[tid] + tile_cnt[blockIdx.x * 32 + tid];
        for (int w = 0; w < RL_PART_WARPS; w++) {
            const uint32_t c = wcnt[w][tid];
            wcnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    for (uint32_t b = s0; b < s1; b += 32) {
        const uint32_t a = b + lane;
        const bool valid = a < s1;
        rl_record rec;
        uint32_t o = 0;
        if (valid) {
            rec = recs[a];
            o = rl_owner_dev(rec.ns_id, world);
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned m = __match_any_sync(vmask, o);
            const int leader = __ffs(m) - 1;
            uint32_t basepos = 0;
            if ((int)lane == leader) {
                basepos = wcnt[warp][o];
                wcnt[warp][o] = basepos + __popc(m);
            }
            basepos = __shfl_sync(m, basepos, leader);
            const uint32_t pos = basepos + __popc(m & ((1u << lane) - 1));
            out_recs[pos] = rec;
            out_src[pos] = a;
        }
        __syncwarp();
    }
}

// single CTA: tile_cnt[t][o] -> exclusive prefix over tiles; owner totals -> owner_base
__global__ void k_bucket_scan(uint32_t num_tiles, uint32_t world, uint32_t* tile_cnt, uint32_t* owner_base,
                              unsigned long long* counts_out) {
    __shared__ uint32_t tot[32];
    const uint32_t o = threadIdx.x;
    if (o < 32) {
        uint32_t run = 0;
        if (o < world)
            for (uint32_t t = 0; t < num_tiles; t++) {
                const uint32_t c = tile_cnt[t * 32 + o];
                tile_cnt[t * 32 + o] = run;
                run += c;
            }
        tot[o] = run;
    }
    __syncthreads();
    if (o == 0) {
        uint32_t run = 0;
        for (uint32_t w = 0; w < world; w++) {
            owner_base[w] = run;
            counts_out[w] = tot[w];
            run += tot[w];
        }
    }
}

__global__ void k_unpermute_u8(uint32_t n, const uint8_t* __restrict__ in, const uint32_t* __restrict__ src,
                               uint8_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[src[i]] = in[i];
}
