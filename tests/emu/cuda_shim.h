// cuda_shim.h — TEST-ONLY: just enough of the CUDA execution model to run the maintenance / CRDT kernels
// (limitador_b200/csrc/rl_maint.cuh, rl_crdt.cuh) on the host, one CUDA thread after the other, so that their index
// arithmetic, probing, claiming and merge rules are checked against the oracle in the GPU-less dev container.
// Not shipped, not a fallback: the product library never includes this file.
//
// What it models: a 1-D grid of 1-D blocks; every thread runs to completion before the next one starts, the threads
// of a launch in a SHUFFLED order (seeded), so that a kernel that silently relies on thread order is caught.  Atomics
// are plain read-modify-writes.  What it does not model: __syncthreads, warp intrinsics, shared memory — the
// kernels under test keep those inside `#ifndef RL_SHIM` fast paths whose results equal the plain path's.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <random>
#include <utility>
#include <vector>

#define RL_SHIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct shim_dim3 {
    unsigned x = 0, y = 0, z = 0;
};
static thread_local shim_dim3 blockIdx, threadIdx, blockDim, gridDim;

struct ulonglong2 {
    unsigned long long x, y;
};
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

template <class T>
static inline T atomicAdd(T* p, T v) {
    const T o = *p;
    *p = o + v;
    return o;
}
template <class T>
static inline T atomicMax(T* p, T v) {
    const T o = *p;
    if (v > o) *p = v;
    return o;
}
template <class T>
static inline T atomicMin(T* p, T v) {
    const T o = *p;
    if (v < o) *p = v;
    return o;
}
template <class T>
static inline T atomicCAS(T* p, T cmp, T val) {
    const T o = *p;
    if (o == cmp) *p = val;
    return o;
}
template <class T>
static inline T atomicExch(T* p, T v) {
    const T o = *p;
    *p = v;
    return o;
}

static uint64_t shim_seed = 1;

// run body() once per thread of a <<<grid, block>>> launch, threads in a seeded random order
template <class F>
static void shim_launch(unsigned grid, unsigned block, F&& body) {
    std::vector<std::pair<unsigned, unsigned>> order;
    order.reserve((size_t)grid * block);
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < block; t++) order.emplace_back(b, t);
    std::mt19937_64 rng(shim_seed++);
    std::shuffle(order.begin(), order.end(), rng);
    gridDim.x = grid;
    blockDim.x = block;
    for (const auto& bt : order) {
        blockIdx.x = bt.first;
        threadIdx.x = bt.second;
        body();
    }
}
