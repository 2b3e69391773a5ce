// simt_selftest.cpp — the fiber emulator (cuda_simt.h) checked on its own: results of the warp intrinsics and the block
// barrier against closed forms, and — run with an argument — the three mistakes it must refuse (they are hangs or undefined
// behaviour on a GPU): a *_sync mask naming an exited lane, lanes disagreeing on the mask, a barrier not every thread reaches.
#include "cuda_simt.h"

#include <cstdio>
#include <cstring>
#include <vector>

static unsigned g_smem[256];

__global__ void k_ok(unsigned* out_match, unsigned* out_sum, unsigned* out_ballot, unsigned* out_block_sum, unsigned* out_shfl) {
    const unsigned t = threadIdx.x, lane = t & 31u;
    const unsigned key = t % 5;  // five groups per warp
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, key);
    out_match[blockIdx.x * blockDim.x + t] = grp;
    out_sum[blockIdx.x * blockDim.x + t] = __reduce_add_sync(grp, t);  // every group reduces over its own mask
    out_ballot[blockIdx.x * blockDim.x + t] = __ballot_sync(0xFFFFFFFFu, lane % 3 == 0);
    out_shfl[blockIdx.x * blockDim.x + t] = __shfl_sync(0xFFFFFFFFu, t * 10, (lane + 1) & 31);
    // block-wide tree reduction through a shared array: every step separated by the barrier
    g_smem[t] = t + 1;
    __syncthreads();
    for (unsigned s = blockDim.x / 2; s > 0; s >>= 1) {
        if (t < s) g_smem[t] += g_smem[t + s];
        __syncthreads();
    }
    if (t == 0) out_block_sum[blockIdx.x] = g_smem[0];
    __syncthreads();  // nobody overwrites g_smem for the next block's use before thread 0 has read it (blocks run in turn)
}

__global__ void k_exited_lane() {
    if ((threadIdx.x & 31u) == 7) return;  // lane 7 leaves ...
    __ballot_sync(0xFFFFFFFFu, 1);         // ... and the others name it
}
__global__ void k_mask_mismatch() {
    const unsigned lane = threadIdx.x & 31u;
    __reduce_add_sync(lane < 16 ? 0x0000FFFFu : 0xFFFFFFFFu, 1u);  // the upper half names lanes that brought another mask
}
__global__ void k_missed_barrier() {
    if (threadIdx.x != 3) __syncthreads();
    if (threadIdx.x == 3) __syncwarp(1u << 3);
    __syncthreads();
    if (threadIdx.x == 5) return;
    __syncthreads();
    // thread 3 skipped the first barrier: the counts no longer line up and somebody waits for ever
    if (threadIdx.x & 1) __syncthreads();
}

int main(int argc, char** argv) {
    if (argc > 1) {
        if (!strcmp(argv[1], "exited")) simt_launch(1, 64, [] { k_exited_lane(); });
        if (!strcmp(argv[1], "mask")) simt_launch(1, 32, [] { k_mask_mismatch(); });
        if (!strcmp(argv[1], "barrier")) simt_launch(1, 32, [] { k_missed_barrier(); });
        printf("not refused\n");
        return 0;
    }
    const unsigned grid = 3, block = 128, n = grid * block;
    std::vector<unsigned> m(n), s(n), b(n), bs(grid), sh(n);
    simt_launch(grid, block, [&] { k_ok(m.data(), s.data(), b.data(), bs.data(), sh.data()); });
    for (unsigned i = 0; i < n; i++) {
        const unsigned t = i % block, lane = t & 31u, w0 = t - lane;
        unsigned grp = 0, sum = 0, ballot = 0;
        for (unsigned j = 0; j < 32; j++) {
            if ((w0 + j) % 5 == t % 5) {
                grp |= 1u << j;
                sum += w0 + j;
            }
            if (j % 3 == 0) ballot |= 1u << j;
        }
        if (m[i] != grp || s[i] != sum || b[i] != ballot || sh[i] != (w0 + ((lane + 1) & 31)) * 10) return 2;
    }
    for (unsigned g = 0; g < grid; g++)
        if (bs[g] != block * (block + 1) / 2) return 3;
    printf("ok simt\n");
    return 0;
}
