// cuda_shim.h -- a minimal CUDA-semantics emulator for unit tests (TEST INFRASTRUCTURE ONLY).
//
// Lets g++ compile smcdet_b200/csrc/smcdet_kernels.cu unchanged (-DSMC_HOSTSIM -x c++) and run the
// kernels' own source on the CPU: one OS thread per CUDA thread, blocks executed one after the
// other, __syncthreads / __syncwarp / warp shuffles / ballots implemented with barriers.  This is
// how the index logic, staging, reductions and the Brent loop are checked against the oracle in
// the CPU-only test tier, before any GPU time is spent.  The product package never loads the
// library built from this header (smcdet_b200/_lib.py only accepts CUDA tensors and only loads
// libsmcdet_b200.so); it exists under tests/ and is built on demand by tests/hostsim/sim.py.
#pragma once

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)

using std::max;
using std::min;

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNoDevice = 100 };
enum { cudaDevAttrMultiProcessorCount = 16, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "hostsim"; }
inline cudaError_t cudaGetDevice(int*) { return cudaErrorNoDevice; }
inline cudaError_t cudaDeviceGetAttribute(int*, int, int) { return cudaErrorNoDevice; }
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

namespace hostsim {

struct Dim3 {
    unsigned x = 0, y = 0, z = 0;
};

struct Warp {
    std::unique_ptr<std::barrier<>> bar;
    uint64_t xchg[32];
};

struct Block {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<Warp> warps;
    std::vector<float> dyn_smem;
};

extern thread_local Dim3 t_threadIdx, t_blockIdx;
extern Dim3 g_blockDim, g_gridDim;
extern thread_local Block* t_block;
extern std::mutex g_atomic_mutex;

inline Warp& my_warp() { return t_block->warps[t_threadIdx.x >> 5]; }

template <class F>
void launch(unsigned grid, unsigned block, size_t smem_bytes, F&& body) {
    g_blockDim.x = block; g_blockDim.y = g_blockDim.z = 1;
    g_gridDim.x = grid; g_gridDim.y = g_gridDim.z = 1;
    for (unsigned b = 0; b < grid; ++b) {
        Block blk;
        blk.bar.reset(new std::barrier<>(block));
        const unsigned nw = (block + 31) / 32;
        blk.warps.resize(nw);
        for (unsigned w = 0; w < nw; ++w) {
            const unsigned lanes = std::min(32u, block - w * 32);
            blk.warps[w].bar.reset(new std::barrier<>(lanes));
        }
        blk.dyn_smem.assign((smem_bytes + sizeof(float) - 1) / sizeof(float), 0.0f);  // exact size, so a host address sanitizer sees overruns
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned i = 0; i < block; ++i) {
            threads.emplace_back([&, i]() {
                t_threadIdx.x = i; t_threadIdx.y = t_threadIdx.z = 0;
                t_blockIdx.x = b; t_blockIdx.y = t_blockIdx.z = 0;
                t_block = &blk;
                body();
                // exited threads no longer take part in barriers, as on the GPU
                blk.warps[i >> 5].bar->arrive_and_drop();
                blk.bar->arrive_and_drop();
            });
        }
        for (auto& th : threads) th.join();
    }
}

template <class T>
inline uint64_t to_bits(T v) {
    uint64_t b = 0;
    static_assert(sizeof(T) <= 8, "shuffle payload too large");
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(uint64_t b) {
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace hostsim

#define threadIdx hostsim::t_threadIdx
#define blockIdx hostsim::t_blockIdx
#define blockDim hostsim::g_blockDim
#define gridDim hostsim::g_gridDim

inline void __syncthreads() { hostsim::t_block->bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { hostsim::my_warp().bar->arrive_and_wait(); }

template <class T>
inline T __shfl_xor_sync(unsigned, T v, int o) {
    hostsim::Warp& w = hostsim::my_warp();
    const int lane = threadIdx.x & 31;
    w.xchg[lane] = hostsim::to_bits(v);
    w.bar->arrive_and_wait();
    const T r = hostsim::from_bits<T>(w.xchg[lane ^ o]);
    w.bar->arrive_and_wait();
    return r;
}

template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
    hostsim::Warp& w = hostsim::my_warp();
    const int lane = threadIdx.x & 31;
    w.xchg[lane] = hostsim::to_bits(v);
    w.bar->arrive_and_wait();
    const T r = hostsim::from_bits<T>(w.xchg[src & 31]);
    w.bar->arrive_and_wait();
    return r;
}

template <class T>
inline T __shfl_up_sync(unsigned, T v, int o) {
    hostsim::Warp& w = hostsim::my_warp();
    const int lane = threadIdx.x & 31;
    w.xchg[lane] = hostsim::to_bits(v);
    w.bar->arrive_and_wait();
    const T r = (lane >= o) ? hostsim::from_bits<T>(w.xchg[lane - o]) : v;
    w.bar->arrive_and_wait();
    return r;
}

inline unsigned __ballot_sync(unsigned, bool pred) {
    hostsim::Warp& w = hostsim::my_warp();
    const int lane = threadIdx.x & 31;
    w.xchg[lane] = pred ? 1u : 0u;
    w.bar->arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)(w.xchg[i] & 1u) << i;
    w.bar->arrive_and_wait();
    return m;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }

// rounding intrinsics: single operations that the compiler must not contract
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }

inline float atomicAdd(float* p, float v) {
    std::lock_guard<std::mutex> g(hostsim::g_atomic_mutex);
    const float old = *p;
    *p = old + v;
    return old;
}
inline int atomicAdd(int* p, int v) {
    std::lock_guard<std::mutex> g(hostsim::g_atomic_mutex);
    const int old = *p;
    *p = old + v;
    return old;
}
inline int atomicOr(int* p, int v) {
    std::lock_guard<std::mutex> g(hostsim::g_atomic_mutex);
    const int old = *p;
    *p = old | v;
    return old;
}

// the two shared-memory declarations used by the kernels
#define SMC_SHARED static
#define SMC_DYN_SHARED(type, name) type* name = reinterpret_cast<type*>(hostsim::t_block->dyn_smem.data())
#define SMC_LAUNCH(kern, grid, block, smem, stream, ...) \
    hostsim::launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })

#ifdef SMC_HOSTSIM_IMPL
namespace hostsim {
thread_local Dim3 t_threadIdx, t_blockIdx;
Dim3 g_blockDim, g_gridDim;
thread_local Block* t_block = nullptr;
std::mutex g_atomic_mutex;
}  // namespace hostsim
#endif
