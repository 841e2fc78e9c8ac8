// smcdet_kernels.cu -- sm_100a kernels and C ABI of libsmcdet_b200.so.
//
// Hot path of timwhite0/smcdet (per-tile likelihood-tempered SMC):
//   loglik_kernel         fused render + per-pixel log density + reduction (images.py:85-102, :159-175)
//   mh_kernel             all num_iters single-site MH sweeps in one launch, likelihood updated
//                         incrementally from a resident rate image (kernel.py:26-130, sampler.py:87-91)
//   temper_update_kernel  on-device Brent solve of ESS(delta)=rho*N + softmax weights / ESS / logZ
//                         (sampler.py:93-125, :181-196)
//   resample_kernel       float64 CDF scan + binary search (sampler.py:127-149); gather_kernel (:150-168)
//   prior_*, prune, psf, render: the small pieces around them
//
// Work decomposition of the two heavy kernels: a particle is owned by TPP consecutive lanes of a
// warp (TPP = 1..32, a template parameter); each lane owns RPT = H/TPP rows of the tile and keeps
// its RPT*W expected counts in registers.  TPP = 1 (one thread per particle, no shuffles, no
// redundant scalar work) is used when there are enough particles to fill the 148 SMs; larger TPP
// spreads a small problem over more threads and handles 16x16 / 32x32 tiles.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#ifdef SMC_HOSTSIM
// unit tests compile this very file with g++ against tests/hostsim/cuda_shim.h (CPU emulation of
// the CUDA execution model) to check the kernels' logic without a GPU; never part of the product
#include "cuda_shim.h"
#else
#include <cuda_runtime.h>
#define SMC_SHARED __shared__
#define SMC_DYN_SHARED(type, name) extern __shared__ __align__(16) type name[]
#define SMC_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "smcdet_math.cuh"

using namespace smc;

namespace {

thread_local char g_err[256] = "ok";

int fail(int code, const char* what) {
    if (code > 0) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString((cudaError_t)code));
    else snprintf(g_err, sizeof(g_err), "%s (code %d)", what, code);
    return code;
}

#define SMC_REQUIRE(cond, code, msg) \
    do { if (!(cond)) return fail((code), (msg)); } while (0)

int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, what);
    return 0;
}

#ifndef SMC_KBT
#define SMC_KBT 128
#endif
constexpr int kBT = SMC_KBT;    // threads per block of the particle kernels
constexpr int kMaxStars = 64;   // D limit (shared-memory staging)

// SM count of the calling thread's current device (cached per device ordinal; 148 on a B200)
int num_sms() {
    constexpr int kMaxDev = 64;
    static int cache[kMaxDev] = {0};  // benign race: every writer stores the same value
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 148;
    if (cache[dev] == 0)
        cache[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    return cache[dev];
}

// Makes the device that owns `ptr` current for the duration of a call and restores the previous one: launches go
// to the device of the caller's buffers whatever device the calling thread had selected.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const void* ptr) {
#ifndef SMC_HOSTSIM
        cudaPointerAttributes at;
        int cur = 0;
        if (ptr != nullptr && cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice &&
            cudaGetDevice(&cur) == cudaSuccess && cur != at.device) {
            if (cudaSetDevice(at.device) == cudaSuccess) prev = cur;
        }
        (void)cudaGetLastError();  // a host pointer makes cudaPointerGetAttributes leave an error behind on old drivers
#else
        (void)ptr;
#endif
    }
    ~DeviceGuard() {
#ifndef SMC_HOSTSIM
        if (prev >= 0) cudaSetDevice(prev);
#endif
    }
};

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
// sum over the TPP lanes of a particle; neighbours first, so that with tree_sum() inside each lane the
// additions form one balanced tree over the tile's rows for every TPP (bit-identical results)
template <int TPP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = 1; o < TPP; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// pixels 4g..4g+3 of a lane: resident rate + pending change (pairs acc[2g], acc[2g+1]), as two packed additions
__device__ __forceinline__ float4 rate_plus(const float4 r, const float2 a, const float2 b) {
    const float2 lo = add2(make_float2(r.x, r.y), a), hi = add2(make_float2(r.z, r.w), b);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// stage the tile and the block's particle catalogs (AoS in global memory, coalesced reads) into
// shared memory as s_star[(d*3 + c) * PB + particle], c = 0 row, 1 col, 2 flux
// Shared-memory layout of the tile (and of lgamma(x + 1)).  A lane reads its chunk of PPT consecutive pixels with
// 128-bit loads; chunks that start PPT floats apart begin in the same banks, so the lanes of a particle collided
// (PPT = 64, 16 lanes per particle: a 16-way conflict on every load of a 32 x 32 tile).  Every chunk is therefore
// followed by one float4 of padding: the chunk stride in 16-byte units (PPT / 4 + 1) is odd, which puts eight
// consecutive lanes into eight different bank groups.  One lane per particle (PPT = HW): all lanes read the same
// address (a broadcast), no padding.
template <int PPT, int HW>
struct TileLayout {
    static constexpr int kPad = (PPT < HW) ? 4 : 0;
    static constexpr int kStride = PPT + kPad;          // floats from one lane's chunk to the next
    static constexpr int kSize = (HW / PPT) * kStride;  // floats of one staged image
};

template <int MODEL, int HW, int PPT>
__device__ __forceinline__ void stage_tile(const float* __restrict__ tile, float* s_tile, float* s_lgam) {
    using TL = TileLayout<PPT, HW>;
    for (int i = threadIdx.x; i < HW; i += kBT) {
        const float x = tile[i];
        const int j = i + TL::kPad * (i / PPT);
        s_tile[j] = x;
        s_lgam[j] = (MODEL == SMCDET_MODEL_GAUSS_POISSON) ? lgammaf(x + 1.0f) : 0.0f;
    }
}

template <int PB>
__device__ __forceinline__ void stage_stars(const float* __restrict__ locs, const float* __restrict__ fluxes, int n_here,
                                            int D, float* s_star) {
    const int nl = n_here * 2 * D;
    for (int i = threadIdx.x; i < nl; i += kBT) {
        const int pi = i / (2 * D), r = i - pi * 2 * D;
        s_star[((r >> 1) * 3 + (r & 1)) * PB + pi] = locs[i];
    }
    const int nf = n_here * D;
    for (int i = threadIdx.x; i < nf; i += kBT) {
        const int pi = i / D, d = i - pi * D;
        s_star[(d * 3 + 2) * PB + pi] = fluxes[i];
    }
    // particles past the end of the tile: inert catalogs
    for (int i = threadIdx.x; i < (PB - n_here) * 3 * D; i += kBT) {
        const int pi = n_here + i / (3 * D), r = i % (3 * D);
        s_star[r * PB + pi] = 0.0f;
    }
}

// the same through the resampling indices: particle pi of the block is the source tile's particle idx[pi]
// (the gather fused into the mutation launch; a record is 12 D contiguous bytes, so reads stay sector-sized)
template <int PB>
__device__ __forceinline__ void stage_stars_gather(const float* __restrict__ locs_tile, const float* __restrict__ fluxes_tile,
                                                   const int64_t* __restrict__ idx, int n_here, int D, float* s_star) {
    const int nl = n_here * 2 * D;
    for (int i = threadIdx.x; i < nl; i += kBT) {
        const int pi = i / (2 * D), r = i - pi * 2 * D;
        s_star[((r >> 1) * 3 + (r & 1)) * PB + pi] = locs_tile[(size_t)idx[pi] * 2 * D + r];
    }
    const int nf = n_here * D;
    for (int i = threadIdx.x; i < nf; i += kBT) {
        const int pi = i / D, d = i - pi * D;
        s_star[(d * 3 + 2) * PB + pi] = fluxes_tile[(size_t)idx[pi] * D + d];
    }
    for (int i = threadIdx.x; i < (PB - n_here) * 3 * D; i += kBT) {
        const int pi = n_here + i / (3 * D), r = i % (3 * D);
        s_star[r * PB + pi] = 0.0f;
    }
}

template <int MODEL, int HW, int PB, int PPT>
__device__ __forceinline__ void stage_block(const float* __restrict__ tile, const float* __restrict__ locs,
                                            const float* __restrict__ fluxes, int n_here, int D,
                                            float* s_tile, float* s_lgam, float* s_star) {
    stage_tile<MODEL, HW, PPT>(tile, s_tile, s_lgam);
    stage_stars<PB>(locs, fluxes, n_here, D, s_star);
}

template <int PB>
__device__ __forceinline__ void unstage_block(float* __restrict__ locs, float* __restrict__ fluxes, int n_here,
                                              int D, const float* s_star) {
    const int nl = n_here * 2 * D;
    for (int i = threadIdx.x; i < nl; i += kBT) {
        const int pi = i / (2 * D), r = i - pi * 2 * D;
        locs[i] = s_star[((r >> 1) * 3 + (r & 1)) * PB + pi];
    }
    const int nf = n_here * D;
    for (int i = threadIdx.x; i < nf; i += kBT) {
        const int pi = i / D, d = i - pi * D;
        fluxes[i] = s_star[(d * 3 + 2) * PB + pi];
    }
}

// The block's expected-count images between shared memory ([PPT/4][kBT] float4: lane `thread`, quad g at g * kBT + thread)
// and HBM ([particle][HW] floats, row-major pixels: lane sub of a particle owns floats sub * PPT .. (sub + 1) * PPT), moved
// by the whole block so that a warp touches 512 consecutive bytes of an image.  The HBM layout does not depend on the number
// of lanes per particle, so consecutive launches may use different decompositions.
template <int TPP, int PPT, int HW>
__device__ __forceinline__ void load_rate_images(const float* __restrict__ rates_tile, const int64_t* __restrict__ idx,
                                                 int n_here, float* s_rate) {
    constexpr int Q = HW / 4, QL = PPT / 4;  // float4 per image / per lane
    float4* dst = reinterpret_cast<float4*>(s_rate);
    for (int i = threadIdx.x; i < n_here * Q; i += kBT) {
        const int p = i / Q, q = i - p * Q;
        const float4* src = reinterpret_cast<const float4*>(rates_tile + (size_t)idx[p] * HW);
        dst[(q % QL) * kBT + p * TPP + q / QL] = src[q];
    }
}

template <int TPP, int PPT, int HW>
__device__ __forceinline__ void store_rate_images(float* __restrict__ rates_block, int n_here, const float* s_rate) {
    constexpr int Q = HW / 4, QL = PPT / 4;
    const float4* src = reinterpret_cast<const float4*>(s_rate);
    float4* dst = reinterpret_cast<float4*>(rates_block);
    for (int i = threadIdx.x; i < n_here * Q; i += kBT) {
        const int p = i / Q, q = i - p * Q;
        dst[i] = src[(q % QL) * kBT + p * TPP + q / QL];
    }
}

// full render of the lane's pixels from the staged catalog (without the background)
template <int MODEL, int RPT, int W, int PB>
__device__ __forceinline__ void render_rows(const ModelK& m, const float* s_star, int pi, int D, int row0,
                                            float2 (&acc)[RPT * W / 2]) {
#pragma unroll
    for (int p = 0; p < RPT * W / 2; ++p) acc[p] = make_float2(0.0f, 0.0f);
    for (int d = 0; d < D; ++d) {
        const float f = s_star[(d * 3 + 2) * PB + pi];
        if (f != 0.0f) {  // empty slots contribute exactly zero (prior.py:61-62 zero-fills them)
            const float l0 = s_star[(d * 3 + 0) * PB + pi];
            const float l1 = s_star[(d * 3 + 1) * PB + pi];
            star_accumulate<MODEL, RPT, W>(m, l0, l1, m.c0 * f, row0, acc);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 1: fused render + log-likelihood
// ---------------------------------------------------------------------------------------------
template <int MODEL, int H, int W, int TPP>
__global__ void __launch_bounds__(kBT) loglik_kernel(const ModelK m, const float* __restrict__ tiles,
                                                     const float* __restrict__ locs,
                                                     const float* __restrict__ fluxes, float* __restrict__ out,
                                                     const int32_t* __restrict__ tile_map, int N, int D,
                                                     int blocks_per_tile) {
    constexpr int PB = kBT / TPP, RPT = H / TPP, PPT = RPT * W, HW = H * W;
    SMC_DYN_SHARED(float, smem);
    using TL = TileLayout<PPT, HW>;
    float* s_tile = smem;
    float* s_lgam = s_tile + TL::kSize;
    float* s_star = s_lgam + TL::kSize;

    const int t = blockIdx.x / blocks_per_tile;
    const int n0 = (blockIdx.x - t * blocks_per_tile) * PB;
    const int n_here = min(PB, N - n0);
    const size_t pbase = (size_t)t * N + n0;
    const int ti = tile_map != nullptr ? tile_map[t] : t;  // the segment's image (count strata share their tile's)
    stage_block<MODEL, HW, PB, PPT>(tiles + (size_t)ti * HW, locs + pbase * 2 * D, fluxes + pbase * D, n_here, D, s_tile,
                               s_lgam, s_star);
    __syncthreads();

    const int pi = threadIdx.x / TPP, sub = threadIdx.x % TPP, row0 = sub * RPT;
    float2 acc[PPT / 2];
    render_rows<MODEL, RPT, W, PB>(m, s_star, pi, D, row0, acc);
    float Q, S;
    pixel_loglik_sum<MODEL, RPT, W>(m, s_tile + sub * TL::kStride, s_lgam + sub * TL::kStride, [&](int g) {
        return rate_plus(make_float4(m.bg, m.bg, m.bg, m.bg), acc[2 * g], acc[2 * g + 1]);
    }, Q, S);
    const float ll = finish_loglik<MODEL>(group_sum<TPP>(Q), group_sum<TPP>(S), HW);
    if (sub == 0 && pi < n_here) out[pbase + pi] = ll;
}

// The same for models whose tile is expensive to stage (Poisson: lgamma(x + 1) of every pixel, ~1e2 instructions
// each): the tile is staged once per block and serves groups_per_block groups of PB particles.  With one group per
// block that staging was most of the kernel for 16 x 16 and 32 x 32 tiles, where a group is only 4-32 particles
// (profiles/r02_sweep_loglik_mh.md: 2.5 ms per 10^6 evaluations of a 32 x 32 tile whatever the number of stars).
template <int MODEL, int H, int W, int TPP>
__global__ void __launch_bounds__(kBT) loglik_groups_kernel(const ModelK m, const float* __restrict__ tiles,
                                                            const float* __restrict__ locs,
                                                            const float* __restrict__ fluxes, float* __restrict__ out,
                                                            const int32_t* __restrict__ tile_map, int N, int D,
                                                            int blocks_per_tile, int groups_per_block) {
    constexpr int PB = kBT / TPP, RPT = H / TPP, PPT = RPT * W, HW = H * W;
    SMC_DYN_SHARED(float, smem);
    using TL = TileLayout<PPT, HW>;
    float* s_tile = smem;
    float* s_lgam = s_tile + TL::kSize;
    float* s_star = s_lgam + TL::kSize;

    const int t = blockIdx.x / blocks_per_tile;
    const int g0 = (blockIdx.x - t * blocks_per_tile) * groups_per_block;
    const int ti = tile_map != nullptr ? tile_map[t] : t;
    stage_tile<MODEL, HW, PPT>(tiles + (size_t)ti * HW, s_tile, s_lgam);
    const int pi = threadIdx.x / TPP, sub = threadIdx.x % TPP, row0 = sub * RPT;
#pragma unroll 1
    for (int g = 0; g < groups_per_block; ++g) {
        const int n0 = (g0 + g) * PB;
        if (n0 >= N) break;  // (uniform over the block)
        const int n_here = min(PB, N - n0);
        const size_t pbase = (size_t)t * N + n0;
        if (g > 0) __syncthreads();  // the previous group's catalogs have been read
        stage_stars<PB>(locs + pbase * 2 * D, fluxes + pbase * D, n_here, D, s_star);
        __syncthreads();
        float2 acc[PPT / 2];
        render_rows<MODEL, RPT, W, PB>(m, s_star, pi, D, row0, acc);
        float Q, S;
        pixel_loglik_sum<MODEL, RPT, W>(m, s_tile + sub * TL::kStride, s_lgam + sub * TL::kStride, [&](int q) {
            return rate_plus(make_float4(m.bg, m.bg, m.bg, m.bg), acc[2 * q], acc[2 * q + 1]);
        }, Q, S);
        const float ll = finish_loglik<MODEL>(group_sum<TPP>(Q), group_sum<TPP>(S), HW);
        if (sub == 0 && pi < n_here) out[pbase + pi] = ll;
    }
}

// any tile shape: one warp per particle, lanes stride over pixels, direct PSF evaluation
__global__ void __launch_bounds__(kBT) loglik_generic_kernel(const ModelK m, const float* __restrict__ tiles,
                                                             const float* __restrict__ locs,
                                                             const float* __restrict__ fluxes,
                                                             float* __restrict__ out,
                                                             const int32_t* __restrict__ tile_map, int T, int N, int D,
                                                             int h, int w) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * (kBT / 32) + (threadIdx.x >> 5);
    if (warp >= (size_t)T * N) return;
    const int t = (int)(warp / N);
    const float* tile = tiles + (size_t)(tile_map != nullptr ? tile_map[t] : t) * h * w;
    const float* l = locs + warp * 2 * D;
    const float* f = fluxes + warp * D;
    float acc = 0.0f;
    for (int p = lane; p < h * w; p += 32) {
        const int i = p / w, j = p - i * w;
        float rate = 0.0f;
        for (int d = 0; d < D; ++d) {
            const float fd = f[d];
            if (fd != 0.0f) rate = fmaf(m.c0 * fd, psf_direct(m, l[2 * d], l[2 * d + 1], i, j), rate);
        }
        rate += m.bg;
        const float x = tile[p];
        if (m.kind == SMCDET_MODEL_M71_NORMAL) {
            const float var = fmaf(m.nms * 4096.0f, rate, m.nas * 4096.0f), dd = x - rate;
            acc += fmaf(-0.5f * dd * dd, rcp_fast(var), fmaf(-0.5f * kLn2, lg2_fast(var), -kLogSqrt2Pi));
        } else {
            const float lg = lg2_fast(rate) * kLn2;
            if (rate > m.nswitch) {
                const float dd = x - rate;
                acc += fmaf(-0.5f * dd * dd, rcp_fast(rate), fmaf(-0.5f, lg, -kLogSqrt2Pi));
            } else {
                acc += ((x == 0.0f) ? 0.0f : x * lg) - rate - lgammaf(x + 1.0f);
            }
        }
    }
    acc = group_sum<32>(acc);
    if (lane == 0) out[warp] = acc;
}

// dense PSF stack [T,h,w,N,D] (images.py:28-76) and rate image [T,h,w,N] (images.py:87-89)
__global__ void psf_kernel(const ModelK m, float norm, const float* __restrict__ locs, float* __restrict__ out,
                           int T, int N, int D, int h, int w) {
    const size_t total = (size_t)T * h * w * N * D;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        size_t r = e;
        const int d = (int)(r % D); r /= D;
        const int n = (int)(r % N); r /= N;
        const int j = (int)(r % w); r /= w;
        const int i = (int)(r % h); r /= h;
        const int t = (int)r;
        const float* l = locs + (((size_t)t * N + n) * D + d) * 2;
        out[e] = norm * psf_direct(m, l[0], l[1], i, j);
    }
}

// PSF as a function of the radius (images.py:25-26, :137-145): the elementwise helper behind ImageModel.psf
__global__ void psf_radial_kernel(const ModelK m, float norm, const float* __restrict__ r, float* __restrict__ out,
                                  size_t n) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const float r2 = r[e] * r[e];
        float v = ex2_fast(-m.k1 * r2);
        if (m.kind == SMCDET_MODEL_M71_NORMAL)
            v += m.b * ex2_fast(-m.k2 * r2) + m.p0 * ex2_fast(m.hb * lg2_fast(fmaf(m.cpl, r2, 1.0f)));
        out[e] = norm * v;
    }
}

__global__ void render_kernel(const ModelK m, const float* __restrict__ locs, const float* __restrict__ fluxes,
                              float* __restrict__ out, int T, int N, int D, int h, int w) {
    const size_t total = (size_t)T * h * w * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        size_t r = e;
        const int n = (int)(r % N); r /= N;
        const int j = (int)(r % w); r /= w;
        const int i = (int)(r % h); r /= h;
        const int t = (int)r;
        const size_t pn = (size_t)t * N + n;
        float rate = 0.0f;
        for (int d = 0; d < D; ++d) {
            const float fd = fluxes[pn * D + d];
            if (fd != 0.0f)
                rate = fmaf(m.c0 * fd, psf_direct(m, locs[(pn * D + d) * 2], locs[(pn * D + d) * 2 + 1], i, j), rate);
        }
        out[e] = rate + m.bg;
    }
}

// ---------------------------------------------------------------------------------------------
// prior
// ---------------------------------------------------------------------------------------------
__global__ void prior_logprob_kernel(const smcdet_prior_params p, const float* __restrict__ counts,
                                     const float* __restrict__ locs, const float* __restrict__ fluxes,
                                     float* __restrict__ out, size_t TN, int D) {
    const size_t pn = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pn >= TN) return;
    const float* l = locs + pn * 2 * D;
    const float* f = fluxes + pn * D;
    out[pn] = prior_logprob_catalog(p, counts[pn], D, [&](int d, float& l0, float& l1, float& fv) {
        l0 = l[2 * d]; l1 = l[2 * d + 1]; fv = f[d];
    });
}

__global__ void prior_sample_kernel(const smcdet_prior_params p, const float* __restrict__ u_locs,
                                    const float* __restrict__ u_fluxes, uint64_t seed,
                                    const int64_t* __restrict__ tile_ids, float* __restrict__ counts,
                                    float* __restrict__ locs, float* __restrict__ fluxes, int T, int M,
                                    int num_per_count, int D) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per star slot
    if (e >= (size_t)T * M * D) return;
    const int d = (int)(e % D);
    const size_t pn = e / D;
    const int n = (int)(pn % M), t = (int)(pn / M);
    const float c = (float)(p.min_objects + n / num_per_count);
    if (d == 0) counts[pn] = c;
    float u0, u1, uf;
    if (u_locs != nullptr) {
        u0 = u_locs[2 * e]; u1 = u_locs[2 * e + 1]; uf = u_fluxes[e];
    } else {
        const uint64_t tid = tile_ids ? (uint64_t)tile_ids[t] : (uint64_t)t;
        const Philox4 r = philox4x32_10((uint32_t)n, (uint32_t)tid, (uint32_t)d, kStreamPriorLocs, (uint32_t)seed,
                                        (uint32_t)(seed >> 32));
        u0 = u01_f(r.v[0]); u1 = u01_f(r.v[1]); uf = u01_f(r.v[2]);
    }
    const float mask = ((float)d < c) ? 1.0f : 0.0f;
    // torch Uniform.rsample: low + rand*(high-low) (prior.py:59)
    locs[2 * e] = (p.loc_low[0] + u0 * (p.loc_high[0] - p.loc_low[0])) * mask;
    locs[2 * e + 1] = (p.loc_low[1] + u1 * (p.loc_high[1] - p.loc_low[1])) * mask;
    float fl;
    if (p.flux_kind == SMCDET_FLUX_TRUNCATED_PARETO) {
        // distributions.py:76-85
        const float Ua = powf(p.flux_upper, p.flux_alpha), La = powf(p.flux_lower, p.flux_alpha);
        const float num = Ua - uf * Ua + uf * La;
        fl = powf(num / (La * Ua), -1.0f / p.flux_alpha);
    } else if (p.flux_kind == SMCDET_FLUX_PARETO) {
        // inverse-cdf form of torch Pareto.sample (Exponential(alpha) -> exp -> * scale)
        fl = p.flux_lower * expf(-log1pf(-uf) / p.flux_alpha);
    } else {
        const float q = clamp_f(uf, 1e-7f, 1.0f - 1e-7f);
        fl = p.flux_mean + p.flux_stdev * kSqrt2 * erfinv_f(2.0f * q - 1.0f);
    }
    fluxes[e] = fl * mask;
}

// ---------------------------------------------------------------------------------------------
// Kernel 2: adaptive tempering + weight update, one block per tile
// ---------------------------------------------------------------------------------------------
constexpr int kTB = 256;

__device__ __forceinline__ double block_sum2(double a, double& b_inout, double* s_red) {
    // reduces (a, b) over the block; every thread returns the totals (a as return, b by reference)
    double b = b_inout;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();  // protect s_red from the previous use
    if (lane == 0) { s_red[2 * wid] = a; s_red[2 * wid + 1] = b; }
    __syncthreads();
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int i = 0; i < kTB / 32; ++i) { ta += s_red[2 * i]; tb += s_red[2 * i + 1]; }
    b_inout = tb;
    return ta;
}

__device__ __forceinline__ float block_max(float v, float* s_redf) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_redf[wid] = v;
    __syncthreads();
    float t = s_redf[0];
#pragma unroll
    for (int i = 1; i < kTB / 32; ++i) t = fmaxf(t, s_redf[i]);
    return t;
}

// ESS(delta) - threshold with ESS = (sum e)^2 / sum e^2, e_i = exp(delta*(l_i - max l))
// (sampler.py:93-97; exp(2 delta l) = e^2 so one exponential per particle serves both sums)
__device__ __forceinline__ double ess_objective(const float* __restrict__ ll, int N, float mx, double delta,
                                                double thr, double* s_red) {
    const float dl2e = (float)delta * kLog2e;
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < N; i += kTB) {
        const float l = ll[i];
        float e = 0.0f;
        if (fabsf(l) <= 3.4028234663852886e38f) e = (dl2e == 0.0f) ? 1.0f : ex2_fast(dl2e * (l - mx));
        s1 += (double)e;
        s2 += (double)e * (double)e;
    }
    s1 = block_sum2(s1, s2, s_red);
    return s1 * s1 / s2 - thr;
}

__global__ void __launch_bounds__(kTB) temper_update_kernel(const float* __restrict__ loglik, float* __restrict__ tau,
                                                            float* __restrict__ tau_prev, float ess_threshold,
                                                            int do_temper, float* __restrict__ wlog,
                                                            float* __restrict__ weights, float* __restrict__ ess,
                                                            float* __restrict__ logz, int32_t* __restrict__ funcalls,
                                                            const int32_t* __restrict__ active,
                                                            const smcdet_loop_state loop, int N) {
    SMC_SHARED double s_red[2 * (kTB / 32)];
    SMC_SHARED float s_redf[kTB / 32];
    const int t = blockIdx.x;
    if (active != nullptr && active[t] == 0) {
        if (threadIdx.x == 0 && loop.active_next != nullptr) loop.active_next[t] = 0;
        return;
    }
    const float* ll = loglik + (size_t)t * N;

    float tau_old, tau_new;
    if (do_temper) {
        // A nan or +inf log-likelihood makes the reference's objective nan at every delta, and `nan < 0` is False
        // (sampler.py:111): such a tile jumps to temperature 1.  +inf is the sentinel of that case in the maximum
        // (only finite values enter it otherwise; -inf entries simply carry no weight).
        float mx = -INFINITY;
        for (int i = threadIdx.x; i < N; i += kTB) {
            const float l = ll[i];
            if (fabsf(l) <= 3.4028234663852886e38f) mx = fmaxf(mx, l);
            else if (!(l == -INFINITY)) mx = INFINITY;
        }
        mx = block_max(mx, s_redf);
        const bool poisoned = (mx == INFINITY);
        if (mx == -INFINITY || poisoned) mx = 0.0f;
        tau_old = tau[t];
        const double thr = (double)ess_threshold;
        const double hi = 1.0 - (double)tau_old;
        int calls = 1;
        double delta = hi;
        const double f_hi = poisoned ? 0.0 : ess_objective(ll, N, mx, hi, thr, s_red);  // (uniform over the block)
        if (f_hi < 0.0) {
            // scipy evaluates both ends again before iterating
            const double f_lo = ess_objective(ll, N, mx, 0.0, thr, s_red);
            calls += 2;
            Brent b;
            b.start(0.0, hi, f_lo, f_hi, 1e-6, 1e-6);
            while (!b.done) {  // uniform across the block: every thread holds the same state
                const double f = ess_objective(ll, N, mx, b.x, thr, s_red);
                ++calls;
                b.step(f);
            }
            delta = b.x;
        }
        tau_new = tau_old + (float)delta;
        if (threadIdx.x == 0) {
            tau_prev[t] = tau_old;
            tau[t] = tau_new;
            if (funcalls) funcalls[t] = calls;
        }
    } else {
        tau_old = tau_prev[t];
        tau_new = tau[t];
    }

    // update_weights (sampler.py:181-196)
    const float dt = tau_new - tau_old;
    float* wl = wlog + (size_t)t * N;
    float* wt = weights + (size_t)t * N;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < N; i += kTB) {
        float v = dt * ll[i];
        // torch.nan_to_num(x, -inf): nan -> -inf, +-inf -> +-FLT_MAX
        if (v != v) v = -INFINITY;
        else if (v == INFINITY) v = 3.4028234663852886e38f;
        else if (v == -INFINITY) v = -3.4028234663852886e38f;
        wl[i] = v;
        m = fmaxf(m, v);
    }
    m = block_max(m, s_redf);
    double s = 0.0, unused = 0.0;
    for (int i = threadIdx.x; i < N; i += kTB) s += (double)expf(wl[i] - m);
    s = block_sum2(s, unused, s_red);
    double s2 = 0.0;
    unused = 0.0;
    const float sf = (float)s;
    for (int i = threadIdx.x; i < N; i += kTB) {
        const float wv = expf(wl[i] - m) / sf;
        wt[i] = wv;
        s2 += (double)wv * (double)wv;
    }
    s2 = block_sum2(s2, unused, s_red);
    if (threadIdx.x == 0) {
        ess[t] = (float)(1.0 / s2);
        logz[t] = logz[t] + m + logf(sf / (float)N);
        // loop bookkeeping of a freeze_finished run (the reference's loop test, sampler.py:230, evaluated here)
        const int live = tau_new < 1.0f ? 1 : 0;
        if (loop.active_next != nullptr) loop.active_next[t] = live;
        if (loop.live_count != nullptr && live) atomicAdd(loop.live_count, 1);
        if (loop.acc_count != nullptr && loop.acc_rate != nullptr) {
            loop.acc_rate[t] = loop.acc_count[t] / (float)N;  // accept.float().mean(-1), kernel.py:130
            loop.acc_count[t] = 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 3: resampling -- float64 inclusive CDF (per-thread chunks + warp-shuffle scan of the chunk
// totals) and one binary search per draw; one block per tile
// ---------------------------------------------------------------------------------------------
// SMEM: the CDF lives in shared memory (N doubles; the searches are dependent loads, 14 per draw at N = 10 000, so
// their latency is the kernel -- 57 us per tile from global memory, a third of that from shared) and the draws of a
// tile may be split over `slices` blocks, each of which builds the same CDF (bit-identical: same chunks, same scan) and
// searches its share; block 0 of a tile also writes the CDF to cdf_all.  Without SMEM (N too large for shared memory)
// one block per tile works on cdf_all directly.
template <bool SMEM>
__global__ void __launch_bounds__(kTB) resample_kernel(int method, const float* __restrict__ weights,
                                                       const double* __restrict__ u, uint64_t seed,
                                                       const int64_t* __restrict__ tile_ids,
                                                       const int32_t* __restrict__ active,
                                                       int64_t* __restrict__ index, double* __restrict__ cdf_all,
                                                       int N, int slices) {
    SMC_DYN_SHARED(double, s_dyn);       // [kTB / 32] warp totals, then (SMEM) the CDF
    double* s_warp = s_dyn;
    const int t = blockIdx.x / slices, slice = blockIdx.x - t * slices;
    const int first = slice * kTB + (int)threadIdx.x, stride = slices * kTB;
    if (active != nullptr && active[t] == 0) {
        for (int i = first; i < N; i += stride) index[(size_t)t * N + i] = i;
        return;
    }
    const float* w = weights + (size_t)t * N;
    double* cdf_out = cdf_all + (size_t)t * N;
    double* cdf = SMEM ? (s_dyn + kTB / 32) : cdf_out;
    const int chunk = (N + kTB - 1) / kTB;
    const int lo = min(N, (int)threadIdx.x * chunk), hi = min(N, lo + chunk);
    double local = 0.0;
    for (int i = lo; i < hi; ++i) local += (double)w[i];
    // inclusive scan of the chunk totals across the block
    double incl = local;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    double base = 0.0;
    for (int i = 0; i < wid; ++i) base += s_warp[i];
    double run = base + (incl - local);
    for (int i = lo; i < hi; ++i) {
        run += (double)w[i];
        cdf[i] = run;
        if (SMEM && slice == 0) cdf_out[i] = run;
    }
    __syncthreads();
    const double total = cdf[N - 1];
    const uint64_t tid = tile_ids ? (uint64_t)tile_ids[t] : (uint64_t)t;
    double u_sys = 0.0;
    if (method == SMCDET_RESAMPLE_SYSTEMATIC) {
        if (u != nullptr) u_sys = u[t];
        else {
            const Philox4 r = philox4x32_10(0u, (uint32_t)tid, (uint32_t)(tid >> 32), kStreamResample ^ 1u,
                                            (uint32_t)seed, (uint32_t)(seed >> 32));
            u_sys = u01_d(r.v[0], r.v[1]);
        }
    }
    for (int i = first; i < N; i += stride) {
        double ui;
        if (method == SMCDET_RESAMPLE_SYSTEMATIC) {
            ui = ((double)i + u_sys) / (double)N;
        } else {
            double uu;
            if (u != nullptr) uu = u[(size_t)t * N + i];
            else {
                const Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)tid, (uint32_t)(tid >> 32), kStreamResample,
                                                (uint32_t)seed, (uint32_t)(seed >> 32));
                uu = u01_d(r.v[0], r.v[1]);
            }
            ui = uu * total;
        }
        int a = 0, b = N;  // first k with cdf[k] >= ui (torch.bucketize, right=False)
        while (a < b) {
            const int mid = a + ((b - a) >> 1);
            if (cdf[mid] >= ui) b = mid; else a = mid + 1;
        }
        index[(size_t)t * N + i] = (int64_t)min(max(a, 0), N - 1);
    }
}

// Latency-bound by nature (index -> load -> store per element), so the kernel keeps as many bytes in flight per
// thread as it can: 8-byte elements (VEC = 2: a location pair, two fluxes; needs an even D) and two independent
// elements per loop trip.  IdxT = uint32_t whenever the element count fits (32-bit divisions).
template <typename IdxT, int VEC>
__device__ __forceinline__ void gather_one(IdxT e, IdxT row, const int64_t* __restrict__ index,
                                           const float* __restrict__ counts_in, const float* __restrict__ locs_in,
                                           const float* __restrict__ fluxes_in, float* __restrict__ counts_out,
                                           float* __restrict__ locs_out, float* __restrict__ fluxes_out,
                                           const int32_t* __restrict__ tile_mask, int N, int D) {
    const IdxT pn = e / row;
    const int c = (int)(e - pn * row);
    const IdxT t = pn / (IdxT)N;
    if (tile_mask != nullptr && tile_mask[t] == 0) return;
    const size_t src = (size_t)t * N + (size_t)index[pn];
    if constexpr (VEC == 2) {
        // row = D location pairs, D/2 flux pairs, the count
        if (c < D)
            reinterpret_cast<float2*>(locs_out)[(size_t)pn * D + c] = reinterpret_cast<const float2*>(locs_in)[src * D + c];
        else if (c < D + D / 2)
            reinterpret_cast<float2*>(fluxes_out)[(size_t)pn * (D / 2) + (c - D)] =
                reinterpret_cast<const float2*>(fluxes_in)[src * (D / 2) + (c - D)];
        else counts_out[pn] = counts_in[src];
    } else {
        if (c < 2 * D) locs_out[(size_t)pn * 2 * D + c] = locs_in[src * 2 * D + c];
        else if (c < 3 * D) fluxes_out[(size_t)pn * D + (c - 2 * D)] = fluxes_in[src * D + (c - 2 * D)];
        else counts_out[pn] = counts_in[src];
    }
}

template <typename IdxT, int VEC>
__global__ void gather_kernel(const int64_t* __restrict__ index, const float* __restrict__ counts_in,
                              const float* __restrict__ locs_in, const float* __restrict__ fluxes_in,
                              float* __restrict__ counts_out, float* __restrict__ locs_out,
                              float* __restrict__ fluxes_out, const int32_t* __restrict__ tile_mask, int T, int N,
                              int D) {
    const IdxT row = (IdxT)(VEC == 2 ? D + D / 2 + 1 : 3 * D + 1);
    const IdxT total = (IdxT)T * (IdxT)N * row;
    const IdxT stride = (IdxT)gridDim.x * (IdxT)blockDim.x;
    IdxT e = (IdxT)blockIdx.x * (IdxT)blockDim.x + threadIdx.x;
    for (; e < total && total - e > stride; e += 2 * stride) {
        gather_one<IdxT, VEC>(e, row, index, counts_in, locs_in, fluxes_in, counts_out, locs_out, fluxes_out, tile_mask, N, D);
        gather_one<IdxT, VEC>(e + stride, row, index, counts_in, locs_in, fluxes_in, counts_out, locs_out, fluxes_out,
                              tile_mask, N, D);
    }
    if (e < total)
        gather_one<IdxT, VEC>(e, row, index, counts_in, locs_in, fluxes_in, counts_out, locs_out, fluxes_out, tile_mask, N, D);
}

// ---------------------------------------------------------------------------------------------
// Kernel 4: fused single-component MH
// ---------------------------------------------------------------------------------------------
// count prior log-pmf, out of line (lgammaf is large and this runs once per particle)
__device__ __noinline__ float count_logpmf_scalar(int kind, float rate, int mn, int mx, float c) {
    smcdet_prior_params p;
    p.count_kind = kind; p.count_rate = rate; p.min_objects = mn; p.max_objects = mx;
    return count_logpmf(p, c);
}

// One coordinate of the random-walk proposal (kernel.py:47-61, :71-85, :97-111; distributions.py:25-52):
// draws x ~ TruncNormal(mu, sigma, [lb, ub]) from the uniform u and returns
//   .y = lq = log q(mu | x) - log q(x | mu) = log mass(mu) - log mass(x)
// (the Gaussian parts of the two proposal densities are identical and cancel).  Out of line: it is called
// three times per sweep and inlining its erf / erfinv / log bodies pushed the hot loop past the
// instruction cache.  `wide` selects the one-erf form valid for boxes of at least 12 sigma.
__device__ __noinline__ float2 truncnormal_step(float mu, float sigma, float inv_sigma_sqrt2, float lb, float ub,
                                                 float u, bool wide) {
    TruncNormal q, r;
    if (wide) q = truncnormal_make_wide(mu, inv_sigma_sqrt2, lb, ub);
    else q = truncnormal_make(mu, sigma, lb, ub);
    const float x = truncnormal_draw(q, mu, sigma, lb, ub, u);
    if (wide) r = truncnormal_make_wide(x, inv_sigma_sqrt2, lb, ub);
    else r = truncnormal_make(x, sigma, lb, ub);
    return make_float2(x, q.log_mass - r.log_mass);  // (by value: an out-parameter of a call goes through local memory)
}

// One coordinate of a proposal for boxes of at least 12 sigma (truncnormal_make_wide + truncnormal_draw + the reverse
// box mass): returns x and *lq_term = log mass(mu) - log mass(x), its summand of lq = log q(prev|prop) - log q(prop|prev).
// Products and sums are spelled with the rounding intrinsics, so that the compiler contracts nothing on its own and
// the three-coordinate and the one-coordinate callers below produce the same bits.
__device__ __forceinline__ float truncnormal_coord_wide(float mu, float sigma, float isig, float lb, float ub, float u,
                                                        float& lq_term) {
    const float a = __fadd_rn(mu, -lb), b = __fadd_rn(ub, -mu);
    const float q = __fmaf_rn(-0.5f, erff(__fmul_rn(fminf(a, b), isig)), 0.5f);
    const float cdf_lb = (a < b) ? q : 0.0f;
    const float mass = __fadd_rn(1.0f, -q);
    const float lo = 1e-6f, hi = (float)(1.0 - 1e-6);
    const float p = clamp_f(u, lo, hi);
    const float y = __fmaf_rn(2.0f, clamp_f(__fmaf_rn(p, mass, cdf_lb), lo, hi), -1.0f);
    const float x = clamp_f(__fmaf_rn(__fmul_rn(sigma, erfinv_f(y)), kSqrt2, mu), lb, ub);
    const float qr = __fmaf_rn(-0.5f, erff(__fmul_rn(fminf(__fadd_rn(x, -lb), __fadd_rn(ub, -x)), isig)), 0.5f);
    lq_term = __fmul_rn(__fadd_rn(lg2_fast(mass), -lg2_fast(__fadd_rn(1.0f, -qr))), kLn2);
    return x;
}

// All three coordinates of a proposal in one out-of-line call.  The three independent erf / erfinv / lg2 dependency
// chains sit in one basic block and are interleaved by the scheduler (ILP 3 instead of 1 in the scalar part of a sweep).
__device__ __noinline__ float4 truncnormal_step3_wide(float mu0, float mu1, float mu2, float sl, float isl, float sf,
                                                       float isf, float lb0, float lb1, float lb2, float ub0, float ub1,
                                                       float ub2, float u0, float u1, float u2) {
    float t0, t1, t2;
    const float x0 = truncnormal_coord_wide(mu0, sl, isl, lb0, ub0, u0, t0);
    const float x1 = truncnormal_coord_wide(mu1, sl, isl, lb1, ub1, u1, t1);
    const float x2 = truncnormal_coord_wide(mu2, sf, isf, lb2, ub2, u2, t2);
    // proposal (row, col, flux) and log q(prev|prop) - log q(prop|prev)
    return make_float4(x0, x1, x2, __fadd_rn(__fadd_rn(t0, t1), t2));
}

// One coordinate out of line, for decompositions with several lanes per particle: lanes 0, 1, 2 of a particle take the
// row, the column and the flux and exchange the results, instead of every lane computing all three (the scalar part is
// most of a sweep when a lane owns a single row of the tile).
__device__ __noinline__ float2 truncnormal_step1_wide(float mu, float sigma, float isig, float lb, float ub, float u) {
    float lq_term;
    const float x = truncnormal_coord_wide(mu, sigma, isig, lb, ub, u, lq_term);
    return make_float2(x, lq_term);
}

// MALA (kernel.py:170-195, :214-259): the proposal is a truncated normal around mean = value + step^2/2 * gradient,
// so the Gaussian parts of the forward and reverse densities do not cancel.
//   truncnormal_propose: draw x ~ TruncNormal(mean, sigma, [lb, ub]) from u; returns (x, log q(x | mean))
//   truncnormal_logq   : log q(x | mean)
__device__ __noinline__ float2 truncnormal_propose(float mean, float sigma, float inv_sigma_sqrt2, float lb, float ub,
                                                    float u, bool wide) {
    const TruncNormal q = wide ? truncnormal_make_wide(mean, inv_sigma_sqrt2, lb, ub) : truncnormal_make(mean, sigma, lb, ub);
    const float x = truncnormal_draw(q, mean, sigma, lb, ub, u);
    return make_float2(x, truncnormal_logpdf(q, mean, sigma, x));
}

__device__ __noinline__ float truncnormal_logq(float mean, float sigma, float inv_sigma_sqrt2, float lb, float ub,
                                                bool wide, float x) {
    const TruncNormal q = wide ? truncnormal_make_wide(mean, inv_sigma_sqrt2, lb, ub) : truncnormal_make(mean, sigma, lb, ub);
    return truncnormal_logpdf(q, mean, sigma, x);
}

// Philox4x32-10 block of MH sweep `it` of one particle (counter: particle, global tile id, SMC iteration and sweep)
__device__ __forceinline__ Philox4 mh_draw_block(uint32_t pidx, uint64_t tile_key, uint64_t offset, int it, uint64_t seed) {
    return philox4x32_10(pidx, (uint32_t)tile_key, (uint32_t)(offset << 16) ^ (uint32_t)it, kStreamMHDraws, (uint32_t)seed,
                         (uint32_t)(seed >> 32));
}

struct MHArgs {
    ModelK m;
    PriorK pk;
    int count_kind, min_objects, max_objects;
    float count_rate;
    smcdet_mh_params mh;
    const float* tiles;
    const float* counts;
    float* locs;
    float* fluxes;
    const float* tau;
    float* loglik_out;
    float* acc_count;
    const int32_t* tape_comp;
    const float* tape_u_loc;
    const float* tape_u_flux;
    const float* tape_u_acc;
    float* tr_log_alpha;
    float* tr_target_prop;
    int8_t* tr_accept;
    float* tr_chain_locs;
    float* tr_chain_fluxes;
    uint64_t seed, offset;
    const int64_t* tile_ids;
    const int32_t* active;
    const int32_t* tile_map;
    int32_t* status;
    int T, N, D, blocks_per_tile;
    // gather fused into the launch (mh_kernel<..., GATHER = true>): particle n of tile t is read from the source arrays at
    // gather_index[t, n] (the resampling step's indices) and written, with its count, to locs / fluxes / counts_out
    const int64_t* gather_index;
    const float* counts_src;
    const float* locs_src;
    const float* fluxes_src;
    float* counts_out;
    const int32_t* copy_mask;
    // expected-count images [T,N,HW] carried between the launches of the SMC loop (GATHER only, both nullable): rates_src is
    // read through gather_index in place of the render of the entry state, rates_out receives the render of the final state
    const float* rates_src;
    float* rates_out;
};

#ifndef SMC_MH_MINB
#define SMC_MH_MINB 3
#endif
template <int MODEL, int H, int W, int TPP, bool MALA, bool GATHER = false, bool CARRY = false>
// CARRY (with GATHER): the launches that read / write the carried expected-count images are instantiations of their own, so
// that the sweep loop of the others is compiled as if the feature did not exist (it is that sensitive).
// Resident blocks per SM: 3 for 64 pixels per lane (168 registers); 5 for 8 pixels per lane, the decomposition of
// a single 8x8 tile -- 10 000 particles x 8 lanes are 625 blocks, which 5 x 148 slots take in ONE wave (4 x 148 do not)
__global__ void __launch_bounds__(kBT, MALA ? 2 : (((H / TPP) * W >= 64) ? SMC_MH_MINB : (((H / TPP) * W <= 8) ? 5 : 4)))
    mh_kernel(const MHArgs a) {
    constexpr int PB = kBT / TPP, RPT = H / TPP, PPT = RPT * W, HW = H * W;
    SMC_DYN_SHARED(float, smem);
    using TL = TileLayout<PPT, HW>;
    float* s_tile = smem;
    float* s_lgam = s_tile + TL::kSize;
    float* s_star = s_lgam + TL::kSize;              // [3*D][PB]
    float* s_rate = s_star + 3 * a.D * PB;    // [PPT][kBT]

    const int t = blockIdx.x / a.blocks_per_tile;
    const int N = a.N, D = a.D;
    const int n0 = (blockIdx.x - t * a.blocks_per_tile) * PB;
    const int n_here = min(PB, N - n0);
    const size_t pbase = (size_t)t * N + n0;
    if (a.active != nullptr && a.active[t] == 0) {
        if constexpr (GATHER) {
            // a tile that is not mutated but whose particles have to reach the destination buffers (it finished in
            // the previous iteration: both buffer sets then hold its final particles) is copied through its indices
            if (a.copy_mask != nullptr && a.copy_mask[t] != 0) {
                stage_stars_gather<PB>(a.locs_src + (size_t)t * N * 2 * D, a.fluxes_src + (size_t)t * N * D,
                                       a.gather_index + pbase, n_here, D, s_star);
                __syncthreads();
                unstage_block<PB>(a.locs + pbase * 2 * D, a.fluxes + pbase * D, n_here, D, s_star);
                for (int i = threadIdx.x; i < n_here; i += kBT)
                    a.counts_out[pbase + i] = a.counts_src[(size_t)t * N + a.gather_index[pbase + i]];
            }
        }
        return;
    }
    const int ti = a.tile_map != nullptr ? a.tile_map[t] : t;  // the segment's image (count strata share their tile's)
    if constexpr (GATHER) {
        stage_tile<MODEL, HW, PPT>(a.tiles + (size_t)ti * HW, s_tile, s_lgam);
        stage_stars_gather<PB>(a.locs_src + (size_t)t * N * 2 * D, a.fluxes_src + (size_t)t * N * D, a.gather_index + pbase,
                               n_here, D, s_star);
        if (CARRY && a.rates_src != nullptr)
            load_rate_images<TPP, PPT, HW>(a.rates_src + (size_t)t * N * HW, a.gather_index + pbase, n_here, s_rate);
    } else {
        stage_block<MODEL, HW, PB, PPT>(a.tiles + (size_t)ti * HW, a.locs + pbase * 2 * D, a.fluxes + pbase * D, n_here, D,
                                        s_tile, s_lgam, s_star);
    }
    __syncthreads();

    const ModelK& m = a.m;
    const int pi = threadIdx.x / TPP, sub = threadIdx.x % TPP, row0 = sub * RPT;
    const bool valid = pi < n_here;
    const size_t pn = pbase + pi;
    const float* xs = s_tile + sub * TL::kStride;  // the lane's pixels (row0 * W on, padded layout)
    const float* lg = s_lgam + sub * TL::kStride;
    float4* my_rate = reinterpret_cast<float4*>(s_rate) + threadIdx.x;  // [PPT/4][kBT] float4, conflict-free
    const float* my_star = s_star + pi;

    // ---- per-particle constants
    float count = (float)D;
    if (valid) count = GATHER ? a.counts_src[(size_t)t * N + a.gather_index[pn]] : a.counts[pn];
    const float tau = a.tau[t];
    // log prior (prior.py:67-75, :220-226) kept as: count term + sum of finite star terms, and the number
    // of live stars outside the location prior's support (each contributes -inf); a single-star move
    // then updates it in O(1)
    const float count_lp = count_logpmf_scalar(a.count_kind, a.count_rate, a.min_objects, a.max_objects, count);
    float prior_fin = 0.0f;
    int prior_bad = 0;
    {
        bool oob = false;
        for (int d = 0; d < D; ++d) {
            const float l0 = my_star[(d * 3 + 0) * PB], l1 = my_star[(d * 3 + 1) * PB], f = my_star[(d * 3 + 2) * PB];
            if ((float)d < count) {
                int bad;
                prior_fin += star_prior_term(a.pk, l0, l1, f, bad);
                prior_bad += bad;
            }
            if (!a.mh.live_only || (float)d < count)  // with live_only the empty slots are never proposed from
                oob |= !(l0 >= a.mh.locs_min[0] && l0 <= a.mh.locs_max[0] && l1 >= a.mh.locs_min[1] &&
                         l1 <= a.mh.locs_max[1] && f >= a.mh.fluxes_min && f <= a.mh.fluxes_max);
        }
        if (a.status != nullptr && valid && sub == 0 && oob) atomicOr(a.status, SMCDET_STATUS_OUT_OF_BOX);
    }

    const uint64_t tile_key = a.tile_ids ? (uint64_t)a.tile_ids[t] : (uint64_t)t;
    const uint32_t pidx = (uint32_t)(n0 + pi);
    const float sl = a.mh.locs_stdev, sf = a.mh.fluxes_stdev;
    const float isl = (1.0f / sl) * kInvSqrt2, isf = (1.0f / sf) * kInvSqrt2;
    // every proposal box of the reference is far wider than 12 sigma; narrower ones take the general path
    // (MALA centres its proposals at value + step^2/2 * gradient, which may lie far outside the box, where the
    // box mass underflows; it always takes the reference's two-erf arithmetic so that regime behaves alike)
    const bool wide_l = !MALA && fminf(a.mh.locs_max[0] - a.mh.locs_min[0], a.mh.locs_max[1] - a.mh.locs_min[1]) >= 12.0f * sl;
    const bool wide_f = !MALA && (a.mh.fluxes_max - a.mh.fluxes_min) >= 12.0f * sf;
    // Philox block of the coming sweep: generated inside the star loop of the previous pass (pure integer work that
    // fills issue slots of that MUFU-bound stretch), again at the top of a sweep only if that loop did not run
    Philox4 nxt = {{0u, 0u, 0u, 0u}};
    int nxt_it = -2;
    int last_acc = 0;
    float ll = 0.0f, cached = 0.0f;
    float2 acc[PPT / 2];

    // One loop, one copy of the render / pixel code (the unrolled body is large, so it must not be
    // replicated: instruction-cache misses were 12% of the stalls when it was):
    //   it = -1          full render of the entry state -> rate image, log-likelihood, cached log target
    //                    (the log_denom_target of kernel.py:88-96)
    //   it = 0..iters-1  one MH sweep: rate' = rate - old star + new star on the lane's pixels
    //   it = iters       fresh full render of the final state -> loglik_out (what sampler.py:100-102 recomputes)
    const int it_end = a.mh.num_iters + ((a.loglik_out != nullptr && a.mh.refresh_loglik) ? 1 : 0);
    const bool carried = CARRY && a.rates_src != nullptr;  // the entry state's rate image is in shared memory already
    for (int it = -1; it < it_end; ++it) {
        const bool full = (it < 0) || (it == a.mh.num_iters);
        const bool render = full && !(carried && it < 0);
        int k = 0;
        bool frozen_slot = false;
        float u0 = 0.5f, u1 = 0.5f, uf = 0.5f, ua = 0.5f;
        float l0 = 0.f, l1 = 0.f, f = 0.f, pl0 = 0.f, pl1 = 0.f, pf = 0.f, lq = 0.f;
        if (!full) {
            // ---- draws: component, 2 location uniforms, flux uniform, accept uniform
            if (a.tape_comp != nullptr) {
                if (valid) {
                    const size_t e = ((size_t)it * a.T + t) * N + (n0 + pi);
                    k = a.tape_comp[e];
                    // a component outside the catalog would index past the staged stars: flagged, and clamped
                    if (k < 0 || k >= D) {
                        if (a.status != nullptr) atomicOr(a.status, SMCDET_STATUS_BAD_TAPE);
                        k = min(max(k, 0), D - 1);
                    }
                    u0 = a.tape_u_loc[2 * e]; u1 = a.tape_u_loc[2 * e + 1];
                    uf = a.tape_u_flux[e]; ua = a.tape_u_acc[e];
                }
            } else {
                if (nxt_it != it) nxt = mh_draw_block(pidx, tile_key, a.offset, it, a.seed);
                // one Philox4x32-10 block per sweep: the high 24 bits of its words are the four uniforms, their low
                // bytes together the 32 random bits that pick the component
                u0 = u01_f(nxt.v[0]); u1 = u01_f(nxt.v[1]); uf = u01_f(nxt.v[2]); ua = u01_f(nxt.v[3]);
                const uint32_t cw = (nxt.v[0] & 0xffu) | ((nxt.v[1] & 0xffu) << 8) | ((nxt.v[2] & 0xffu) << 16) | (nxt.v[3] << 24);
                k = (int)(((uint64_t)cw * (uint64_t)(a.mh.live_only ? max((int)count, 1) : D)) >> 32);
            }
            // live_only: only stars j < count move; otherwise the proposal is the current state
            frozen_slot = a.mh.live_only && !((float)k < count);
            if (frozen_slot) k = 0;

            l0 = my_star[(k * 3 + 0) * PB]; l1 = my_star[(k * 3 + 1) * PB]; f = my_star[(k * 3 + 2) * PB];
        }

        // ---- proposal for star k (kernel.py:47-61; distributions.py:40-48)
        if (!full) {
            if constexpr (!MALA && TPP >= 4) {
                if (wide_l && wide_f) {  // (uniform branch; the shuffles below run with the whole warp converged)
                    const int c = sub % 3;
                    const float2 xq = truncnormal_step1_wide(
                        c == 0 ? l0 : (c == 1 ? l1 : f), c == 2 ? sf : sl, c == 2 ? isf : isl,
                        c == 0 ? a.mh.locs_min[0] : (c == 1 ? a.mh.locs_min[1] : a.mh.fluxes_min),
                        c == 0 ? a.mh.locs_max[0] : (c == 1 ? a.mh.locs_max[1] : a.mh.fluxes_max),
                        c == 0 ? u0 : (c == 1 ? u1 : uf));
                    const float xc = xq.x, lqc = xq.y;
                    const int base = (int)(threadIdx.x & 31u) - sub;  // the particle's first lane
                    pl0 = __shfl_sync(0xffffffffu, xc, base); pl1 = __shfl_sync(0xffffffffu, xc, base + 1);
                    pf = __shfl_sync(0xffffffffu, xc, base + 2);
                    lq = __fadd_rn(__fadd_rn(__shfl_sync(0xffffffffu, lqc, base), __shfl_sync(0xffffffffu, lqc, base + 1)),
                                   __shfl_sync(0xffffffffu, lqc, base + 2));
                }
            }
            if (frozen_slot) {
                pl0 = l0; pl1 = l1; pf = f; lq = 0.0f;  // nothing is rendered or re-priced below
            } else if constexpr (!MALA) {
                if (wide_l && wide_f) {
                    if constexpr (TPP < 4) {
                        const float4 pr = truncnormal_step3_wide(l0, l1, f, sl, isl, sf, isf, a.mh.locs_min[0], a.mh.locs_min[1],
                                                                 a.mh.fluxes_min, a.mh.locs_max[0], a.mh.locs_max[1],
                                                                 a.mh.fluxes_max, u0, u1, uf);
                        pl0 = pr.x; pl1 = pr.y; pf = pr.z; lq = pr.w;
                    }
                } else {
                    const float2 r0 = truncnormal_step(l0, sl, isl, a.mh.locs_min[0], a.mh.locs_max[0], u0, wide_l);
                    const float2 r1 = truncnormal_step(l1, sl, isl, a.mh.locs_min[1], a.mh.locs_max[1], u1, wide_l);
                    const float2 rf = truncnormal_step(f, sf, isf, a.mh.fluxes_min, a.mh.fluxes_max, uf, wide_f);
                    pl0 = r0.x; pl1 = r1.x; pf = rf.x;
                    lq = (r0.y + r1.y) + rf.y;
                }
            }
        }

        // ---- expected counts: all D stars (full render) or -old star +new star (MH sweep)
#pragma unroll
        for (int p = 0; p < PPT / 2; ++p) acc[p] = make_float2(0.0f, 0.0f);
        if constexpr (MALA) {
            if (!full) {  // (a frozen slot runs through the same shuffles as its warp's other particles)
                // gradient of the log target wrt star k at the current state (kernel.py:159-167), removing the star
                // from the rate image in the same pass, then the Langevin proposal (kernel.py:170-195)
                float sP, s0, s1;
                const float wgt_old = m.c0 * f;
                star_grad_accumulate<MODEL, RPT, W, true>(m, l0, l1, -wgt_old, row0, [&](int r, float (&wr)[W]) {
#pragma unroll
                    for (int g = 0; g < W / 4; ++g) {
                        const float4 rt = my_rate[(r * (W / 4) + g) * kBT];
                        const float4 xv = reinterpret_cast<const float4*>(xs)[r * (W / 4) + g];
                        wr[4 * g] = pixel_dlogpdf<MODEL>(m, xv.x, rt.x); wr[4 * g + 1] = pixel_dlogpdf<MODEL>(m, xv.y, rt.y);
                        wr[4 * g + 2] = pixel_dlogpdf<MODEL>(m, xv.z, rt.z); wr[4 * g + 3] = pixel_dlogpdf<MODEL>(m, xv.w, rt.w);
                    }
                }, acc, sP, s0, s1);
                sP = group_sum<TPP>(sP); s0 = group_sum<TPP>(s0); s1 = group_sum<TPP>(s1);
                const float dpr = ((float)k < count) ? star_prior_dflux(a.pk, f) : 0.0f;
                const float hl = 0.5f * (sl * sl), hf = 0.5f * (sf * sf);
                const float qm0 = fmaf(hl, tau * wgt_old * s0, l0), qm1 = fmaf(hl, tau * wgt_old * s1, l1);
                const float qmf = fmaf(hf, fmaf(tau * m.c0, sP, dpr), f);
                const float2 p0 = truncnormal_propose(qm0, sl, isl, a.mh.locs_min[0], a.mh.locs_max[0], u0, wide_l);
                const float2 p1 = truncnormal_propose(qm1, sl, isl, a.mh.locs_min[1], a.mh.locs_max[1], u1, wide_l);
                const float2 p2 = truncnormal_propose(qmf, sf, isf, a.mh.fluxes_min, a.mh.fluxes_max, uf, wide_f);
                pl0 = p0.x; pl1 = p1.x; pf = p2.x;
                const float q0 = p0.y, q1 = p1.y, qf = p2.y;
                if (frozen_slot) { pl0 = l0; pl1 = l1; pf = f; }  // the star removed above is put back unchanged
                lq = -((q0 + q1) + qf);  // - log q(proposal | current); the reverse term is added below
                if (pf != 0.0f) star_accumulate<MODEL, RPT, W>(m, pl0, pl1, m.c0 * pf, row0, acc);
            }
        }
        const int ns = full ? (render ? D : 0) : ((MALA || frozen_slot) ? 0 : 2);
#pragma unroll 1
        for (int s = 0; s < ns; ++s) {
            float s0, s1, sw;
            if (full) {
                s0 = my_star[(s * 3 + 0) * PB]; s1 = my_star[(s * 3 + 1) * PB]; sw = m.c0 * my_star[(s * 3 + 2) * PB];
            } else if (s == 0) {
                s0 = l0; s1 = l1; sw = -(m.c0 * f);
            } else {
                s0 = pl0; s1 = pl1; sw = m.c0 * pf;
            }
            if (sw != 0.0f) star_accumulate<MODEL, RPT, W>(m, s0, s1, sw, row0, acc);
            // the next sweep's random block (a pure function of the counters, so generating it here -- by every pass
            // of this loop alike -- changes nothing but where the integer work is issued)
            nxt = mh_draw_block(pidx, tile_key, a.offset, it + 1, a.seed);
            nxt_it = it + 1;
        }
        if (render) {
#pragma unroll
            for (int g = 0; g < PPT / 4; ++g) {
                my_rate[g * kBT] = rate_plus(make_float4(m.bg, m.bg, m.bg, m.bg), acc[2 * g], acc[2 * g + 1]);
                acc[2 * g] = acc[2 * g + 1] = make_float2(0.0f, 0.0f);
            }
        }
        float pq, ps;
        pixel_loglik_sum<MODEL, RPT, W>(m, xs, lg, [&](int g) {
            return rate_plus(my_rate[g * kBT], acc[2 * g], acc[2 * g + 1]);
        }, pq, ps);
        const float llp = finish_loglik<MODEL>(group_sum<TPP>(pq), group_sum<TPP>(ps), HW);

        if (full) {
            ll = llp;
            if (it < 0) cached = (prior_bad ? -INFINITY : count_lp + prior_fin) + tau * ll;
            continue;
        }

        // ---- prior of the proposal and the MH ratio (sampler.py:87-91, kernel.py:114-116)
        float fin_p = prior_fin;
        int bad_p = prior_bad;
        if ((float)k < count && !frozen_slot) {
            int bad_old, bad_new;
            const float t_old = star_prior_term(a.pk, l0, l1, f, bad_old);
            const float t_new = star_prior_term(a.pk, pl0, pl1, pf, bad_new);
            fin_p = (prior_fin - t_old) + t_new;
            bad_p = prior_bad - bad_old + bad_new;
        }
        const float target_p = (bad_p ? -INFINITY : count_lp + fin_p) + tau * llp;
        if constexpr (MALA) {
            // gradient at the proposal (kernel.py:199-213) and the density of the reverse move (kernel.py:214-240)
            float sP, s0, s1;
            const float wgt_new = m.c0 * pf;
            star_grad_accumulate<MODEL, RPT, W, false>(m, pl0, pl1, 0.0f, row0, [&](int r, float (&wr)[W]) {
#pragma unroll
                for (int g = 0; g < W / 4; ++g) {
                    const float4 rt = my_rate[(r * (W / 4) + g) * kBT];
                    const float4 xv = reinterpret_cast<const float4*>(xs)[r * (W / 4) + g];
                    const int p = r * W + 4 * g;
                    wr[4 * g] = pixel_dlogpdf<MODEL>(m, xv.x, rt.x + acc[p / 2].x);
                    wr[4 * g + 1] = pixel_dlogpdf<MODEL>(m, xv.y, rt.y + acc[p / 2].y);
                    wr[4 * g + 2] = pixel_dlogpdf<MODEL>(m, xv.z, rt.z + acc[p / 2 + 1].x);
                    wr[4 * g + 3] = pixel_dlogpdf<MODEL>(m, xv.w, rt.w + acc[p / 2 + 1].y);
                }
            }, acc, sP, s0, s1);
            sP = group_sum<TPP>(sP); s0 = group_sum<TPP>(s0); s1 = group_sum<TPP>(s1);
            const float dpr = ((float)k < count) ? star_prior_dflux(a.pk, pf) : 0.0f;
            const float hl = 0.5f * (sl * sl), hf = 0.5f * (sf * sf);
            const float rm0 = fmaf(hl, tau * wgt_new * s0, pl0), rm1 = fmaf(hl, tau * wgt_new * s1, pl1);
            const float rmf = fmaf(hf, fmaf(tau * m.c0, sP, dpr), pf);
            lq += (truncnormal_logq(rm0, sl, isl, a.mh.locs_min[0], a.mh.locs_max[0], wide_l, l0) +
                   truncnormal_logq(rm1, sl, isl, a.mh.locs_min[1], a.mh.locs_max[1], wide_l, l1)) +
                  truncnormal_logq(rmf, sf, isf, a.mh.fluxes_min, a.mh.fluxes_max, wide_f, f);
            if (frozen_slot) lq = 0.0f;
        }
        const float log_alpha = (target_p - cached) + lq;
        float alpha = ex2_fast(log_alpha * kLog2e);
        if (alpha > 1.0f) alpha = 1.0f;  // clamp(max=1) keeps nan
        const bool accept = (ua <= alpha);

        __syncwarp();  // every lane of the particle has read the old star
        if (accept) {
#pragma unroll
            for (int g = 0; g < PPT / 4; ++g) my_rate[g * kBT] = rate_plus(my_rate[g * kBT], acc[2 * g], acc[2 * g + 1]);
            ll = llp;
            prior_fin = fin_p;
            prior_bad = bad_p;
            if (sub == 0) {
                s_star[(k * 3 + 0) * PB + pi] = pl0;
                s_star[(k * 3 + 1) * PB + pi] = pl1;
                s_star[(k * 3 + 2) * PB + pi] = pf;
            }
        }
        if constexpr (MALA) {
            cached = accept ? target_p : cached;  // torch.where (kernel.py:273)
        } else {
            // arithmetic blend with the bool as in kernel.py:125 (-inf * 0 = nan poisons the cache)
            cached = target_p * (accept ? 1.0f : 0.0f) + cached * (accept ? 0.0f : 1.0f);
        }
        last_acc = accept ? 1 : 0;
        __syncwarp();
        if (valid && sub == 0 && (a.tr_accept != nullptr || a.tr_log_alpha != nullptr || a.tr_target_prop != nullptr ||
                                  a.tr_chain_locs != nullptr)) {
            const size_t e = ((size_t)it * a.T + t) * N + (n0 + pi);
            if (a.tr_log_alpha) a.tr_log_alpha[e] = log_alpha;
            if (a.tr_target_prop) a.tr_target_prop[e] = target_p;
            if (a.tr_accept) a.tr_accept[e] = (int8_t)last_acc;
            if (a.tr_chain_locs != nullptr) {  // the chain of MHsampler (sampler.py:516-523)
                const size_t row = (pn * (size_t)a.mh.num_iters + (size_t)it) * D;
                for (int d = 0; d < D; ++d) {
                    a.tr_chain_locs[(row + d) * 2] = my_star[(d * 3 + 0) * PB];
                    a.tr_chain_locs[(row + d) * 2 + 1] = my_star[(d * 3 + 1) * PB];
                    a.tr_chain_fluxes[row + d] = my_star[(d * 3 + 2) * PB];
                }
            }
        }
    }
    if (a.loglik_out != nullptr && valid && sub == 0) a.loglik_out[pn] = ll;
    if constexpr (GATHER) {
        if (valid && sub == 0) a.counts_out[pn] = count;
    }
    const unsigned votes = __ballot_sync(0xffffffffu, valid && sub == 0 && last_acc);
    if ((threadIdx.x & 31) == 0 && votes != 0) atomicAdd(a.acc_count + t, (float)__popc(votes));
    __syncthreads();
    unstage_block<PB>(a.locs + pbase * 2 * D, a.fluxes + pbase * D, n_here, D, s_star);
    if constexpr (CARRY) {  // (the last pass was the fresh render of the final state: mutate_impl sees to that)
        if (a.rates_out != nullptr) store_rate_images<TPP, PPT, HW>(a.rates_out + pbase * HW, n_here, s_rate);
    }
}

__global__ void divide_kernel(float* v, const int32_t* active, float denom, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T && (active == nullptr || active[t] != 0)) v[t] = v[t] / denom;  // accept.float().mean(-1)
}

__global__ void zero_active_kernel(float* v, const int32_t* active, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T && (active == nullptr || active[t] != 0)) v[t] = 0.0f;
}

// ---------------------------------------------------------------------------------------------
// prune (sampler.py:198-219): order-preserving compaction, one thread per particle
// ---------------------------------------------------------------------------------------------
__global__ void prune_kernel(const float* __restrict__ locs, const float* __restrict__ fluxes, float tile_h,
                             float tile_w, float thr, int64_t* __restrict__ counts, float* __restrict__ locs_out,
                             float* __restrict__ fluxes_out, size_t TN, int D) {
    const size_t pn = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pn >= TN) return;
    const float* l = locs + pn * 2 * D;
    const float* f = fluxes + pn * D;
    float* lo = locs_out + pn * 2 * D;
    float* fo = fluxes_out + pn * D;
    int k = 0;
    for (int d = 0; d < D; ++d) {
        const float l0 = l[2 * d], l1 = l[2 * d + 1], fv = f[d];
        const bool keep = (l0 > 0.0f && l0 < tile_h) && (l1 > 0.0f && l1 < tile_w) && (fv > thr);
        if (keep) { lo[2 * k] = l0; lo[2 * k + 1] = l1; fo[k] = fv; ++k; }
    }
    counts[pn] = k;
    for (; k < D; ++k) { lo[2 * k] = 0.0f; lo[2 * k + 1] = 0.0f; fo[k] = 0.0f; }
}

// ---------------------------------------------------------------------------------------------
// match_catalogs (metrics.py:8-84): one thread per (tile, drawn catalog) solves the rectangular
// assignment problem of true vs estimated stars with scipy's algorithm (Crouse 2016 shortest
// augmenting path, float64, same operation order -- see oracle/smcdet_oracle.c: oracle_lsap) and
// counts total / matched stars per magnitude bin.  Evaluation utility, not a hot kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxMatch = 96;  // stars per side of one matching problem (local-memory arrays)

__device__ __forceinline__ int bucket_of(float x, const float* __restrict__ bins, int B) {
    int k = 0;  // torch.bucketize(right=False): boundaries strictly below x
    while (k < B && bins[k] < x) ++k;
    return k;
}

__global__ void match_kernel(const float* __restrict__ true_counts, const float* __restrict__ true_locs,
                             const float* __restrict__ true_fluxes, const float* __restrict__ est_counts,
                             const float* __restrict__ est_locs, const float* __restrict__ est_fluxes,
                             const int64_t* __restrict__ index, const float* __restrict__ mag_bins, float locs_tol,
                             float mags_tol, float* __restrict__ true_total, float* __restrict__ true_match,
                             float* __restrict__ est_total, float* __restrict__ est_match, int32_t* status, int T,
                             int n, int M, int Dt, int De, int B) {
    const int prob = blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= T * n) return;
    const int t = prob / n;
    const size_t cat = (size_t)t * M + (size_t)index[prob];
    float* o_tt = true_total + (size_t)prob * B;
    float* o_tm = true_match + (size_t)prob * B;
    float* o_et = est_total + (size_t)prob * B;
    float* o_em = est_match + (size_t)prob * B;
    for (int b = 0; b < B; ++b) { o_tt[b] = 0.f; o_tm[b] = 0.f; o_et[b] = 0.f; o_em[b] = 0.f; }
    const int nt = (int)true_counts[t], ne = (int)est_counts[cat];
    if (nt < 0 || ne < 0 || nt > kMaxMatch || ne > kMaxMatch || nt > Dt || ne > De) {
        if (status) atomicOr(status, 4);
        return;
    }
    const float* tl = true_locs + (size_t)t * Dt * 2;
    const float* el = est_locs + cat * De * 2;
    float tm[kMaxMatch], em[kMaxMatch];
    for (int i = 0; i < nt; ++i) tm[i] = 22.5f - mul_unfused(2.5f, log10f(true_fluxes[(size_t)t * Dt + i]));
    for (int j = 0; j < ne; ++j) em[j] = 22.5f - mul_unfused(2.5f, log10f(est_fluxes[cat * De + j]));
    for (int i = 0; i < nt; ++i) { const int b = bucket_of(tm[i], mag_bins, B); if (b < B) o_tt[b] += 1.f; }
    for (int j = 0; j < ne; ++j) { const int b = bucket_of(em[j], mag_bins, B); if (b < B) o_et[b] += 1.f; }
    if (nt == 0 || ne == 0) return;

    // rows = the smaller side (scipy transposes when there are more rows than columns)
    const bool tr = ne < nt;
    const int R = tr ? ne : nt, Cn = tr ? nt : ne;
    auto pair_cost = [&](int r, int c, bool& oob) -> double {
        const int i = tr ? c : r, j = tr ? r : c;   // i: true star, j: estimated star
        const float dx = tl[2 * i] - el[2 * j], dy = tl[2 * i + 1] - el[2 * j + 1];
        const float dist = sqrtf(mul_unfused(dx, dx) + mul_unfused(dy, dy));
        oob = dist > locs_tol || fabsf(tm[i] - em[j]) > mags_tol;
        return oob ? (double)(dist + 1e20f) : (double)dist;   // metrics.py:60 adds the penalty in float32
    };
    double u[kMaxMatch], v[kMaxMatch], spc[kMaxMatch];
    int path[kMaxMatch], col4row[kMaxMatch], row4col[kMaxMatch], remaining[kMaxMatch];
    bool SR[kMaxMatch], SC[kMaxMatch];
    for (int i = 0; i < R; ++i) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = 0; j < Cn; ++j) { v[j] = 0.0; path[j] = -1; row4col[j] = -1; }
    const double inf = (double)INFINITY;
    for (int cur = 0; cur < R; ++cur) {
        double min_val = 0.0;
        int num_remaining = Cn, sink = -1, i = cur;
        for (int it = 0; it < Cn; ++it) { remaining[it] = Cn - it - 1; SC[it] = false; spc[it] = inf; }
        for (int r = 0; r < R; ++r) SR[r] = false;
        while (sink == -1) {
            int idx = -1;
            double lowest = inf;
            SR[i] = true;
            for (int it = 0; it < num_remaining; ++it) {
                const int j = remaining[it];
                bool o;
                const double r = min_val + pair_cost(i, j, o) - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; idx = it; }
            }
            min_val = lowest;
            const int j = remaining[idx];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = true;
            remaining[idx] = remaining[--num_remaining];
        }
        u[cur] += min_val;
        for (int r = 0; r < R; ++r) if (SR[r] && r != cur) u[r] += min_val - spc[col4row[r]];
        for (int j = 0; j < Cn; ++j) if (SC[j]) v[j] -= min_val - spc[j];
        int j = sink;
        while (true) {
            const int r = path[j];
            row4col[j] = r;
            const int old = col4row[r];
            col4row[r] = j;
            j = old;
            if (r == cur) break;
        }
    }
    for (int r = 0; r < R; ++r) {
        bool o;
        pair_cost(r, col4row[r], o);
        if (o) continue;
        const int i = tr ? col4row[r] : r, j = tr ? r : col4row[r];
        const int bt = bucket_of(tm[i], mag_bins, B), be = bucket_of(em[j], mag_bins, B);
        if (bt < B) o_tm[bt] += 1.f;
        if (be < B) o_em[be] += 1.f;
    }
}

// ---------------------------------------------------------------------------------------------
// Aggregate tree merge (aggregate.py:189-324, :105-128): merging two neighbouring tiles' catalogs
// and mutating the merged catalogs under the bridge target
//     log prior(parent) + (1 - tau) * [loglik(child 1) + loglik(child 2)] + tau * loglik(parent)
// ---------------------------------------------------------------------------------------------
// drop_sources_from_overlap + join (aggregate.py:189-265) for one parent particle per thread:
// the first child keeps stars with 0 != loc_axis < dim, the second those with loc_axis > 0 (shifted by
// dim into the parent's frame); kept values are compacted to the front, field by field as the
// reference's three sorts do.  Output catalogs have 2*M slots; counts_out = number of kept stars.
__global__ void agg_join_kernel(const float* __restrict__ locs, const float* __restrict__ fluxes, int axis, float dim,
                                float* __restrict__ counts_out, float* __restrict__ locs_out,
                                float* __restrict__ fluxes_out, int nH, int nW, int N, int M) {
    const int pH = axis == 0 ? nH / 2 : nH, pW = axis == 1 ? nW / 2 : nW;
    const size_t total = (size_t)pH * pW * N;
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    const int n = (int)(id % N);
    const int pw = (int)((id / N) % pW), ph = (int)(id / ((size_t)N * pW));
    float* lo = locs_out + id * (size_t)(4 * M);
    float* fo = fluxes_out + id * (size_t)(2 * M);
    int k0 = 0, k1 = 0, kf = 0, kept = 0;
    for (int c = 0; c < 2; ++c) {
        const int h = axis == 0 ? 2 * ph + c : ph, w = axis == 1 ? 2 * pw + c : pw;
        const size_t src = ((size_t)h * nW + w) * N + n;
        const float* l = locs + src * (size_t)(2 * M);
        const float* f = fluxes + src * (size_t)M;
        for (int d = 0; d < M; ++d) {
            float l0 = l[2 * d], l1 = l[2 * d + 1], fv = f[d];
            const float la = axis == 0 ? l0 : l1;
            const bool keep = c == 0 ? (la < dim && la != 0.0f) : (la > 0.0f);
            if (!keep) continue;
            ++kept;
            if (c == 1) {  // shift into the parent's frame; exact zeros stay "empty" (aggregate.py:243-248)
                if (axis == 0) { if (l0 != 0.0f) l0 += dim; } else { if (l1 != 0.0f) l1 += dim; }
            }
            if (l0 != 0.0f) lo[2 * k0++] = l0;
            if (l1 != 0.0f) lo[2 * k1++ + 1] = l1;
            if (fv != 0.0f) fo[kf++] = fv;
        }
    }
    for (; k0 < 2 * M; ++k0) lo[2 * k0] = 0.0f;
    for (; k1 < 2 * M; ++k1) lo[2 * k1 + 1] = 0.0f;
    for (; kf < 2 * M; ++kf) fo[kf] = 0.0f;
    counts_out[id] = (float)kept;
}

// unjoin (aggregate.py:267-324): split parent catalogs at loc_axis <= half into the two children's frames;
// children are written parent-major ([parent tile][child 0/1][particle]), field-wise compaction as above
__global__ void agg_unjoin_kernel(const float* __restrict__ locs, const float* __restrict__ fluxes, int axis, float half,
                                  float* __restrict__ counts_out, float* __restrict__ locs_out,
                                  float* __restrict__ fluxes_out, int T, int N, int D) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (size_t)T * N) return;
    const size_t t = id / N, n = id % N;
    const float* l = locs + id * (size_t)(2 * D);
    const float* f = fluxes + id * (size_t)D;
    for (int c = 0; c < 2; ++c) {
        const size_t dst = (t * 2 + c) * N + n;
        float* lo = locs_out + dst * (size_t)(2 * D);
        float* fo = fluxes_out + dst * (size_t)D;
        int k0 = 0, k1 = 0, kf = 0, cnt = 0;
        for (int d = 0; d < D; ++d) {
            float l0 = l[2 * d], l1 = l[2 * d + 1];
            const float fv = f[d];
            const bool first = (axis == 0 ? l0 : l1) <= half;
            if (first != (c == 0)) continue;
            if (l0 != 0.0f && l1 != 0.0f) ++cnt;
            if (c == 1) { if (axis == 0) { if (l0 != 0.0f) l0 -= half; } else { if (l1 != 0.0f) l1 -= half; } }
            if (l0 != 0.0f) lo[2 * k0++] = l0;
            if (l1 != 0.0f) lo[2 * k1++ + 1] = l1;
            if (fv != 0.0f) fo[kf++] = fv;
        }
        for (; k0 < D; ++k0) lo[2 * k0] = 0.0f;
        for (; k1 < D; ++k1) lo[2 * k1 + 1] = 0.0f;
        for (; kf < D; ++kf) fo[kf] = 0.0f;
        counts_out[dst] = (float)cnt;
    }
}

struct AggOut {
    float* loglik_diff;    // [T,N] parent - (child 1 + child 2) at the final state (aggregate.py:539-541)
    float* parent_loglik;  // [T,N] nullable
    float* child_loglik;   // [T,N] nullable: sum over the two children
    float* log_target;     // [T,N] nullable: bridge target of the final state
    float half;
};

// The bridge-target sweep.  A particle is owned by TPP lanes as in loglik_kernel and mh_kernel.  Two rate
// images live in shared memory for the whole launch: the parent image and the "children" image, in which a star
// only reaches the pixels of its own half -- which is what evaluating the two child tiles on their own catalogs
// amounts to: the half boundary is an integer, so pixel-to-star offsets and the floor()-anchored PSF patch are
// the same in the child's frame.  A sweep renders the old and the new star once each (star_accumulate) and
// updates both images' lane pixels incrementally, like mh_kernel; it = -1 is the full render of the entry state.
template <int RPT, int W, int H, int AXIS>
__device__ __forceinline__ void agg_split_add(const float2 (&tmp)[RPT * W / 2], bool second, int row0,
                                              float2 (&accP)[RPT * W / 2], float2 (&accC)[RPT * W / 2]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int c = 0; c < W / 2; ++c) {  // a pair of columns never straddles the split (W / 2 is even)
            const bool pix_second = AXIS == 0 ? (row0 + r >= H / 2) : (2 * c >= W / 2);
            const float2 v = tmp[r * (W / 2) + c];
            accP[r * (W / 2) + c] = add2(accP[r * (W / 2) + c], v);
            accC[r * (W / 2) + c] = add2(accC[r * (W / 2) + c], (pix_second == second) ? v : make_float2(0.0f, 0.0f));
        }
}

template <int MODEL, int H, int W, int TPP, int AXIS>
__global__ void __launch_bounds__(kBT) agg_mh_kernel(const MHArgs a, const AggOut o) {
    constexpr int PB = kBT / TPP, RPT = H / TPP, PPT = RPT * W, HW = H * W;
    SMC_DYN_SHARED(float, smem);
    using TL = TileLayout<PPT, HW>;
    float* s_tile = smem;
    float* s_lgam = s_tile + TL::kSize;
    float* s_star = s_lgam + TL::kSize;               // [3*D][PB]
    float* s_rateP = s_star + 3 * a.D * PB;    // [PPT][kBT]
    float* s_rateC = s_rateP + PPT * kBT;      // [PPT][kBT]

    const int t = blockIdx.x / a.blocks_per_tile;
    if (a.active != nullptr && a.active[t] == 0) return;
    const int N = a.N, D = a.D;
    const int n0 = (blockIdx.x - t * a.blocks_per_tile) * PB;
    const int n_here = min(PB, N - n0);
    const size_t pbase = (size_t)t * N + n0;
    stage_block<MODEL, HW, PB, PPT>(a.tiles + (size_t)t * HW, a.locs + pbase * 2 * D, a.fluxes + pbase * D, n_here, D,
                               s_tile, s_lgam, s_star);
    __syncthreads();

    const ModelK& m = a.m;
    const int pi = threadIdx.x / TPP, sub = threadIdx.x % TPP, row0 = sub * RPT;
    const bool valid = pi < n_here;
    const size_t pn = pbase + pi;
    const float* my_star = s_star + pi;
    float4* rateP = reinterpret_cast<float4*>(s_rateP) + threadIdx.x;  // [PPT/4][kBT] float4
    float4* rateC = reinterpret_cast<float4*>(s_rateC) + threadIdx.x;
    const float* xs = s_tile + sub * TL::kStride;  // the lane's pixels (row0 * W on, padded layout)
    const float* lg = s_lgam + sub * TL::kStride;
    const float count = valid ? a.counts[pn] : 0.0f;
    const int icount = (int)count;
    const float tau = a.tau[t];
    const float count_lp = count_logpmf_scalar(a.count_kind, a.count_rate, a.min_objects, a.max_objects, count);
    const uint64_t tile_key = a.tile_ids ? (uint64_t)a.tile_ids[t] : (uint64_t)t;
    const uint32_t pidx = (uint32_t)(n0 + pi);
    const float sl = a.mh.locs_stdev, sf = a.mh.fluxes_stdev;
    const float isl = (1.0f / sl) * kInvSqrt2, isf = (1.0f / sf) * kInvSqrt2;
    const bool wide_l = fminf(a.mh.locs_max[0] - a.mh.locs_min[0], a.mh.locs_max[1] - a.mh.locs_min[1]) >= 12.0f * sl;
    const bool wide_f = (a.mh.fluxes_max - a.mh.fluxes_min) >= 12.0f * sf;
    // log prior kept as in mh_kernel: finite part + number of live stars outside the support
    float prior_fin = 0.0f;
    int prior_bad = 0;
    for (int d = 0; d < D; ++d)
        if ((float)d < count) {
            int bad;
            prior_fin += star_prior_term(a.pk, my_star[(d * 3 + 0) * PB], my_star[(d * 3 + 1) * PB],
                                         my_star[(d * 3 + 2) * PB], bad);
            prior_bad += bad;
        }
    int last_acc = 0;
    float cached = 0.0f, ll_par = 0.0f, ll_chi = 0.0f;

    for (int it = -1; it < a.mh.num_iters; ++it) {
        const bool full = it < 0;
        int k = 0;
        bool live = false;
        float u0 = 0.5f, u1 = 0.5f, uf = 0.5f, ua = 0.5f;
        float l0 = 0.f, l1 = 0.f, f = 0.f, pl0 = 0.f, pl1 = 0.f, pf = 0.f, lq = 0.f;
        const size_t e = ((size_t)max(it, 0) * a.T + t) * N + (n0 + pi);
        if (!full) {
            if (a.tape_comp != nullptr) {
                if (valid) {
                    k = a.tape_comp[e];
                    u0 = a.tape_u_loc[2 * e]; u1 = a.tape_u_loc[2 * e + 1];
                    uf = a.tape_u_flux[e]; ua = a.tape_u_acc[e];
                }
            } else {
                const uint32_t c2 = (uint32_t)(a.offset << 16) ^ (uint32_t)it;
                const Philox4 r = philox4x32_10(pidx, (uint32_t)tile_key, c2, kStreamMHDraws, (uint32_t)a.seed,
                                                (uint32_t)(a.seed >> 32));
                const Philox4 rc = philox4x32_10(pidx, (uint32_t)tile_key, c2, kStreamMHComp, (uint32_t)a.seed,
                                                 (uint32_t)(a.seed >> 32));
                u0 = u01_f(r.v[0]); u1 = u01_f(r.v[1]); uf = u01_f(r.v[2]); ua = u01_f(r.v[3]);
                k = (int)(((uint64_t)rc.v[0] * (uint64_t)max(icount, 1)) >> 32);  // uniform over the live stars
            }
            live = k >= 0 && k < icount && k < D;  // anything else (also a bad tape entry) leaves the catalog as it is
            if (!live) k = 0;
            l0 = my_star[(k * 3 + 0) * PB]; l1 = my_star[(k * 3 + 1) * PB]; f = my_star[(k * 3 + 2) * PB];
            pl0 = l0; pl1 = l1; pf = f;
            if (live) {
                if (wide_l && wide_f) {  // the three coordinates interleaved in one call, as in mh_kernel
                    const float4 pr = truncnormal_step3_wide(l0, l1, f, sl, isl, sf, isf, a.mh.locs_min[0], a.mh.locs_min[1],
                                                             a.mh.fluxes_min, a.mh.locs_max[0], a.mh.locs_max[1],
                                                             a.mh.fluxes_max, u0, u1, uf);
                    pl0 = pr.x; pl1 = pr.y; pf = pr.z; lq = pr.w;
                } else {
                    const float2 r0 = truncnormal_step(l0, sl, isl, a.mh.locs_min[0], a.mh.locs_max[0], u0, wide_l);
                    const float2 r1 = truncnormal_step(l1, sl, isl, a.mh.locs_min[1], a.mh.locs_max[1], u1, wide_l);
                    const float2 rf = truncnormal_step(f, sf, isf, a.mh.fluxes_min, a.mh.fluxes_max, uf, wide_f);
                    pl0 = r0.x; pl1 = r1.x; pf = rf.x;
                    lq = (r0.y + r1.y) + rf.y;
                }
            }
        }
        // ---- changes of both images on the lane's pixels: all D stars (entry) or -old star +new star (sweep)
        float2 accP[PPT / 2], accC[PPT / 2];
#pragma unroll
        for (int p = 0; p < PPT / 2; ++p) { accP[p] = make_float2(0.0f, 0.0f); accC[p] = make_float2(0.0f, 0.0f); }
        const int ns = full ? D : (live ? 2 : 0);
#pragma unroll 1
        for (int s = 0; s < ns; ++s) {
            float s0, s1, sw;
            if (full) {
                s0 = my_star[(s * 3 + 0) * PB]; s1 = my_star[(s * 3 + 1) * PB]; sw = m.c0 * my_star[(s * 3 + 2) * PB];
            } else if (s == 0) {
                s0 = l0; s1 = l1; sw = -(m.c0 * f);
            } else {
                s0 = pl0; s1 = pl1; sw = m.c0 * pf;
            }
            if (sw == 0.0f) continue;
            float2 tmp[PPT / 2];
#pragma unroll
            for (int p = 0; p < PPT / 2; ++p) tmp[p] = make_float2(0.0f, 0.0f);
            star_accumulate<MODEL, RPT, W>(m, s0, s1, sw, row0, tmp);
            // aggregate.py:279-281: a star belongs to the first child iff loc_axis <= half
            agg_split_add<RPT, W, H, AXIS>(tmp, (AXIS == 0 ? s0 : s1) > o.half, row0, accP, accC);
        }
        if (full) {
#pragma unroll
            for (int g = 0; g < PPT / 4; ++g) {
                const float4 bg4 = make_float4(m.bg, m.bg, m.bg, m.bg);
                rateP[g * kBT] = rate_plus(bg4, accP[2 * g], accP[2 * g + 1]);
                rateC[g * kBT] = rate_plus(bg4, accC[2 * g], accC[2 * g + 1]);
                accP[2 * g] = accP[2 * g + 1] = accC[2 * g] = accC[2 * g + 1] = make_float2(0.0f, 0.0f);
            }
        }
        float q, sg;
        pixel_loglik_sum<MODEL, RPT, W>(m, xs, lg, [&](int g) {
            return rate_plus(rateP[g * kBT], accP[2 * g], accP[2 * g + 1]);
        }, q, sg);
        const float llp = finish_loglik<MODEL>(group_sum<TPP>(q), group_sum<TPP>(sg), HW);
        pixel_loglik_sum<MODEL, RPT, W>(m, xs, lg, [&](int g) {
            return rate_plus(rateC[g * kBT], accC[2 * g], accC[2 * g + 1]);
        }, q, sg);
        const float llc = finish_loglik<MODEL>(group_sum<TPP>(q), group_sum<TPP>(sg), HW);
        float fin_p = prior_fin;
        int bad_p = prior_bad;
        if (!full && live) {
            int bad_old, bad_new;
            const float t_old = star_prior_term(a.pk, l0, l1, f, bad_old);
            const float t_new = star_prior_term(a.pk, pl0, pl1, pf, bad_new);
            fin_p = (prior_fin - t_old) + t_new;
            bad_p = prior_bad - bad_old + bad_new;
        }
        const float target = ((bad_p ? -INFINITY : count_lp + fin_p) + (1.0f - tau) * llc) + tau * llp;
        if (full) {
            cached = target; ll_par = llp; ll_chi = llc;
            continue;
        }
        const float log_alpha = (target - cached) + lq;
        float alpha = ex2_fast(log_alpha * kLog2e);
        if (alpha > 1.0f) alpha = 1.0f;
        const bool accept = (ua <= alpha);
        __syncwarp();  // every lane of the particle has read the old star
        if (accept) {
#pragma unroll
            for (int g = 0; g < PPT / 4; ++g) {
                rateP[g * kBT] = rate_plus(rateP[g * kBT], accP[2 * g], accP[2 * g + 1]);
                rateC[g * kBT] = rate_plus(rateC[g * kBT], accC[2 * g], accC[2 * g + 1]);
            }
            ll_par = llp; ll_chi = llc;
            prior_fin = fin_p; prior_bad = bad_p;
            if (sub == 0 && live) {
                s_star[(k * 3 + 0) * PB + pi] = pl0;
                s_star[(k * 3 + 1) * PB + pi] = pl1;
                s_star[(k * 3 + 2) * PB + pi] = pf;
            }
        }
        cached = target * (accept ? 1.0f : 0.0f) + cached * (accept ? 0.0f : 1.0f);  // kernel.py:125
        last_acc = accept ? 1 : 0;
        __syncwarp();
        if (valid && sub == 0) {
            if (a.tr_log_alpha) a.tr_log_alpha[e] = log_alpha;
            if (a.tr_target_prop) a.tr_target_prop[e] = target;
            if (a.tr_accept) a.tr_accept[e] = (int8_t)last_acc;
        }
    }
    if (valid && sub == 0) {
        if (o.loglik_diff) o.loglik_diff[pn] = ll_par - ll_chi;
        if (o.parent_loglik) o.parent_loglik[pn] = ll_par;
        if (o.child_loglik) o.child_loglik[pn] = ll_chi;
        if (o.log_target) o.log_target[pn] = cached;
    }
    const unsigned votes = __ballot_sync(0xffffffffu, valid && sub == 0 && last_acc);
    if ((threadIdx.x & 31) == 0 && votes != 0) atomicAdd(a.acc_count + t, (float)__popc(votes));
    __syncthreads();
    unstage_block<PB>(a.locs + pbase * 2 * D, a.fluxes + pbase * D, n_here, D, s_star);
}

template <int MODEL, int H, int W, int TPP, int AXIS>
int launch_agg_t(MHArgs a, const AggOut& o, cudaStream_t st) {
    constexpr int PB = kBT / TPP;
    a.blocks_per_tile = (a.N + PB - 1) / PB;
    constexpr int PPT = (H / TPP) * W;
    const size_t smem = sizeof(float) * ((size_t)2 * TileLayout<PPT, H * W>::kSize + (size_t)3 * a.D * PB + (size_t)2 * PPT * kBT);
    const long long grid = (long long)a.T * a.blocks_per_tile;
    if (grid >= (1LL << 31)) return fail(SMCDET_E_TOO_LARGE, "smcdet_agg_mutate: grid too large");
#ifndef SMC_HOSTSIM
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(agg_mh_kernel<MODEL, H, W, TPP, AXIS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(agg_mh_kernel)");
    }
#endif
    SMC_LAUNCH((agg_mh_kernel<MODEL, H, W, TPP, AXIS>), (unsigned)grid, kBT, smem, st, a, o);
    return launch_status("agg_mh_kernel");
}

template <int MODEL>
int dispatch_agg(int h, int w, int axis, const MHArgs& a, const AggOut& o, cudaStream_t st) {
    // a parent tile is two child tiles side by side: 2s x s after a merge along rows, 2s x 2s after the next one
    if (h == 16 && w == 8 && axis == 0) return launch_agg_t<MODEL, 16, 8, 8, 0>(a, o, st);
    if (h == 16 && w == 16 && axis == 1) return launch_agg_t<MODEL, 16, 16, 16, 1>(a, o, st);
    if (h == 32 && w == 16 && axis == 0) return launch_agg_t<MODEL, 32, 16, 32, 0>(a, o, st);
    if (h == 32 && w == 32 && axis == 1) return launch_agg_t<MODEL, 32, 32, 32, 1>(a, o, st);
    return fail(SMCDET_E_UNSUPPORTED, "smcdet_agg_mutate: parent tile must be 16x8, 16x16, 32x16 or 32x32 with the matching axis");
}

// ---------------------------------------------------------------------------------------------
// dispatch over (model, tile, threads-per-particle)
// ---------------------------------------------------------------------------------------------
// smallest TPP whose grid fills the machine, within what is instantiated for the tile size
int choose_tpp(int side, long long particles) {
    const int min_tpp = side == 8 ? 1 : (side == 16 ? 4 : 16);
    const int max_tpp = side == 8 ? 8 : (side == 16 ? 16 : 32);
    // measured on B200 (scripts/gpu_probe3.py, N = 10 000, 8x8): splitting a particle over more lanes only pays
    // while the grid has fewer than ~256 threads per SM
    const long long want = (long long)num_sms() * 256;
    int tpp = min_tpp;
    while (tpp < max_tpp && particles * tpp < want) tpp *= 2;
    return tpp;
}

thread_local int g_force_tpp = 0;  // diagnostic override of the calling thread (smcdet_debug_force_tpp)

// smallest tile side whose Poisson-model likelihood takes loglik_groups_kernel (measured on B200: the loop costs
// registers, which 8 x 8 tiles -- 0.5 lgamma per thread to stage -- do not get back)
#ifndef SMC_GROUPS_MIN_SIDE
#define SMC_GROUPS_MIN_SIDE 32
#endif
// ... and the largest catalog: with many stars per catalog the staging is amortised anyway and the plain kernel's
// register allocation wins (32 x 32, B200: D = 1 0.88 against 1.08 ms per 10^6 evaluations, D = 16 4.11 against 3.76)
#ifndef SMC_GROUPS_MAX_STARS
#define SMC_GROUPS_MAX_STARS 4
#endif
template <int MODEL, int H, int TPP>
int launch_loglik_t(const ModelK& m, const float* tiles, const float* locs, const float* fluxes, float* out,
                    const int32_t* tile_map, int T, int N, int D, cudaStream_t st) {
    constexpr int PB = kBT / TPP;
    const int groups = (N + PB - 1) / PB;  // groups of PB particles per tile
    const size_t smem = sizeof(float) * (2 * TileLayout<(H / TPP) * H, H * H>::kSize + 3 * (size_t)D * PB);
    if constexpr (MODEL == SMCDET_MODEL_GAUSS_POISSON && H >= SMC_GROUPS_MIN_SIDE) if (D <= SMC_GROUPS_MAX_STARS) {
        // several groups per block as long as the grid keeps >= 8 blocks per resident slot; blocks of a tile get
        // equal shares
        const long long slots = (long long)num_sms() * 4 * 8;
        const long long want = ((long long)T * groups) / slots;
        const int gmax = (int)std::min<long long>(std::max<long long>(want, 1), (H >= 16) ? 64 : 16);
        const int nb = (groups + gmax - 1) / gmax;
        const int gpb = (groups + nb - 1) / nb;
        const int bpt = (groups + gpb - 1) / gpb;
        if ((long long)T * bpt >= (1LL << 31)) return fail(SMCDET_E_TOO_LARGE, "smcdet_loglik: grid too large");
        auto kern = loglik_groups_kernel<MODEL, H, H, TPP>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(loglik)");
        }
        SMC_LAUNCH(kern, (unsigned)((size_t)T * bpt), kBT, smem, st, m, tiles, locs, fluxes, out, tile_map, N, D, bpt, gpb);
        return launch_status("loglik_groups_kernel");
    }
    {
        const int bpt = groups;
        if ((long long)T * bpt >= (1LL << 31)) return fail(SMCDET_E_TOO_LARGE, "smcdet_loglik: grid too large");
        auto kern = loglik_kernel<MODEL, H, H, TPP>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(loglik)");
        }
        SMC_LAUNCH(kern, (unsigned)((size_t)T * bpt), kBT, smem, st, m, tiles, locs, fluxes, out, tile_map, N, D, bpt);
        return launch_status("loglik_kernel");
    }
}

template <int MODEL, int H, int TPP, bool MALA, bool GATHER = false, bool CARRY = false>
int launch_mh_t(MHArgs& a, cudaStream_t st) {
    constexpr int PB = kBT / TPP, PPT = (H / TPP) * H;
    a.blocks_per_tile = (a.N + PB - 1) / PB;
    const size_t smem = sizeof(float) * (2 * TileLayout<PPT, H * H>::kSize + 3 * (size_t)a.D * PB + (size_t)PPT * kBT);
    if ((long long)a.T * a.blocks_per_tile >= (1LL << 31)) return fail(SMCDET_E_TOO_LARGE, "smcdet_mh_mutate: grid too large");
    auto kern = mh_kernel<MODEL, H, H, TPP, MALA, GATHER, CARRY>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(mh)");
    }
    SMC_LAUNCH(kern, (unsigned)((size_t)a.T * a.blocks_per_tile), kBT, smem, st, a);
    return launch_status("mh_kernel");
}

#define SMC_DISPATCH_TPP(FN, MODEL, H, tpp, ...)                         \
    switch (tpp) {                                                       \
        case 1: if constexpr (H == 8) return FN<MODEL, H, 1>(__VA_ARGS__); break;   \
        case 2: if constexpr (H == 8) return FN<MODEL, H, 2>(__VA_ARGS__); break;   \
        case 4: if constexpr (H <= 16) return FN<MODEL, H, 4>(__VA_ARGS__); break;  \
        case 8: if constexpr (H <= 16) return FN<MODEL, H, 8>(__VA_ARGS__); break;  \
        case 16: if constexpr (H >= 16) return FN<MODEL, H, 16>(__VA_ARGS__); break; \
        case 32: if constexpr (H == 32) return FN<MODEL, H, 32>(__VA_ARGS__); break; \
        default: break;                                                  \
    }                                                                    \
    return fail(SMCDET_E_UNSUPPORTED, "threads-per-particle not instantiated for this tile size");

template <int MODEL, int H>
int dispatch_loglik_tpp(int tpp, const ModelK& m, const float* tiles, const float* locs, const float* fluxes,
                        float* out, const int32_t* tile_map, int T, int N, int D, cudaStream_t st) {
    SMC_DISPATCH_TPP(launch_loglik_t, MODEL, H, tpp, m, tiles, locs, fluxes, out, tile_map, T, N, D, st)
}

template <int MODEL, int H, int TPP>
int launch_mh_plain(MHArgs& a, cudaStream_t st) { return launch_mh_t<MODEL, H, TPP, false>(a, st); }
template <int MODEL, int H, int TPP>
int launch_mh_mala(MHArgs& a, cudaStream_t st) { return launch_mh_t<MODEL, H, TPP, true>(a, st); }
template <int MODEL, int H, int TPP>
int launch_mh_gather(MHArgs& a, cudaStream_t st) { return launch_mh_t<MODEL, H, TPP, false, true>(a, st); }
template <int MODEL, int H, int TPP>
int launch_mh_carry(MHArgs& a, cudaStream_t st) { return launch_mh_t<MODEL, H, TPP, false, true, true>(a, st); }

template <int MODEL, int H>
int dispatch_mh_tpp(int tpp, bool mala, MHArgs& a, cudaStream_t st) {
    if (mala) { SMC_DISPATCH_TPP(launch_mh_mala, MODEL, H, tpp, a, st) }
    if (a.gather_index != nullptr && (a.rates_src != nullptr || a.rates_out != nullptr)) {
        SMC_DISPATCH_TPP(launch_mh_carry, MODEL, H, tpp, a, st)
    }
    if (a.gather_index != nullptr) { SMC_DISPATCH_TPP(launch_mh_gather, MODEL, H, tpp, a, st) }
    SMC_DISPATCH_TPP(launch_mh_plain, MODEL, H, tpp, a, st)
}

template <int MODEL>
int dispatch_loglik_side(int side, int tpp, const ModelK& m, const float* tiles, const float* locs,
                         const float* fluxes, float* out, const int32_t* tile_map, int T, int N, int D, cudaStream_t st) {
    if (side == 8) return dispatch_loglik_tpp<MODEL, 8>(tpp, m, tiles, locs, fluxes, out, tile_map, T, N, D, st);
    if (side == 16) return dispatch_loglik_tpp<MODEL, 16>(tpp, m, tiles, locs, fluxes, out, tile_map, T, N, D, st);
    return dispatch_loglik_tpp<MODEL, 32>(tpp, m, tiles, locs, fluxes, out, tile_map, T, N, D, st);
}

template <int MODEL>
int dispatch_mh_side(int side, int tpp, bool mala, MHArgs& a, cudaStream_t st) {
    if (side == 8) return dispatch_mh_tpp<MODEL, 8>(tpp, mala, a, st);
    if (side == 16) return dispatch_mh_tpp<MODEL, 16>(tpp, mala, a, st);
    return dispatch_mh_tpp<MODEL, 32>(tpp, mala, a, st);
}

bool model_ok(const smcdet_model_params* p) {
    return p != nullptr &&
           (p->model_kind == SMCDET_MODEL_GAUSS_POISSON || p->model_kind == SMCDET_MODEL_M71_NORMAL) &&
           p->psf_radius >= 0;
}

unsigned grid_for(size_t total, int block) {
    const size_t want = (total + block - 1) / block;
    const size_t cap = (size_t)num_sms() * 32;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int smcdet_version(void) { return SMCDET_ABI_VERSION; }

const char* smcdet_last_error_string(void) { return g_err; }

// diagnostic: force the threads-per-particle choice of the calling thread's next loglik / mh launches (0 = automatic)
int smcdet_debug_force_tpp(int tpp) {
    g_force_tpp = tpp;
    return 0;
}

static int loglik_impl(const smcdet_model_params* model, const float* tiles, const int32_t* tile_of_segment,
                       const float* locs, const float* fluxes, float* loglik, int T, int N, int D, int h, int w,
                       void* stream) {
    DeviceGuard guard(locs);
    SMC_REQUIRE(model_ok(model), SMCDET_E_INVALID, "smcdet_loglik: bad model parameters");
    SMC_REQUIRE(tiles && locs && fluxes && loglik, SMCDET_E_INVALID, "smcdet_loglik: null pointer");
    SMC_REQUIRE(T > 0 && N > 0 && D > 0 && h > 0 && w > 0, SMCDET_E_INVALID, "smcdet_loglik: non-positive size");
    cudaStream_t st = (cudaStream_t)stream;
    const ModelK m = make_model_k(*model);
    const bool fast = (h == w) && (h == 8 || h == 16 || h == 32) && D <= kMaxStars;
    if (fast) {
        const int tpp = g_force_tpp ? g_force_tpp : choose_tpp(h, (long long)T * N);
        if (model->model_kind == SMCDET_MODEL_M71_NORMAL)
            return dispatch_loglik_side<SMCDET_MODEL_M71_NORMAL>(h, tpp, m, tiles, locs, fluxes, loglik, tile_of_segment, T, N, D, st);
        return dispatch_loglik_side<SMCDET_MODEL_GAUSS_POISSON>(h, tpp, m, tiles, locs, fluxes, loglik, tile_of_segment, T, N, D, st);
    }
    const size_t warps = (size_t)T * N;
    const size_t blocks = (warps + (kBT / 32) - 1) / (kBT / 32);
    SMC_REQUIRE(blocks < 0x7fffffffull, SMCDET_E_TOO_LARGE, "smcdet_loglik: too many particles for one launch");
    SMC_LAUNCH(loglik_generic_kernel, (unsigned)blocks, kBT, 0, st, m, tiles, locs, fluxes, loglik, tile_of_segment, T, N, D, h, w);
    return launch_status("loglik_generic_kernel");
}

int smcdet_loglik(const smcdet_model_params* model, const float* tiles, const float* locs, const float* fluxes,
                  float* loglik, int T, int N, int D, int h, int w, void* stream) {
    return loglik_impl(model, tiles, nullptr, locs, fluxes, loglik, T, N, D, h, w, stream);
}

int smcdet_loglik_segments(const smcdet_model_params* model, const float* tiles, const int32_t* tile_of_segment,
                           const float* locs, const float* fluxes, float* loglik, int S, int N, int D, int h, int w,
                           void* stream) {
    SMC_REQUIRE(tile_of_segment != nullptr, SMCDET_E_INVALID, "smcdet_loglik_segments: null segment map");
    return loglik_impl(model, tiles, tile_of_segment, locs, fluxes, loglik, S, N, D, h, w, stream);
}

int smcdet_psf(const smcdet_model_params* model, const float* locs, float* psf, int T, int N, int D, int h, int w,
               void* stream) {
    DeviceGuard guard(locs);
    SMC_REQUIRE(model_ok(model), SMCDET_E_INVALID, "smcdet_psf: bad model parameters");
    SMC_REQUIRE(locs && psf, SMCDET_E_INVALID, "smcdet_psf: null pointer");
    SMC_REQUIRE(T > 0 && N > 0 && D > 0 && h > 0 && w > 0, SMCDET_E_INVALID, "smcdet_psf: non-positive size");
    const ModelK m = make_model_k(*model);
    const float norm = m.cn;
    const size_t total = (size_t)T * h * w * N * D;
    SMC_LAUNCH(psf_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, m, norm, locs, psf, T, N, D, h, w);
    return launch_status("psf_kernel");
}

int smcdet_psf_radial(const smcdet_model_params* model, int normalized, const float* r, float* out, long long n,
                      void* stream) {
    DeviceGuard guard(r);
    SMC_REQUIRE(model_ok(model), SMCDET_E_INVALID, "smcdet_psf_radial: bad model parameters");
    SMC_REQUIRE(r && out && n > 0, SMCDET_E_INVALID, "smcdet_psf_radial: null pointer or non-positive size");
    const ModelK m = make_model_k(*model);
    // m.cn carries 1 / ((1 + b + p0) Z) for the M71 PSF: the un-normalised form drops the Z
    const float norm = (normalized || model->model_kind != SMCDET_MODEL_M71_NORMAL) ? m.cn : m.cn * model->psf_norm;
    SMC_LAUNCH(psf_radial_kernel, grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream, m, norm, r, out, (size_t)n);
    return launch_status("psf_radial_kernel");
}

int smcdet_render(const smcdet_model_params* model, const float* locs, const float* fluxes, float* rate, int T,
                  int N, int D, int h, int w, void* stream) {
    DeviceGuard guard(locs);
    SMC_REQUIRE(model_ok(model), SMCDET_E_INVALID, "smcdet_render: bad model parameters");
    SMC_REQUIRE(locs && fluxes && rate, SMCDET_E_INVALID, "smcdet_render: null pointer");
    SMC_REQUIRE(T > 0 && N > 0 && D > 0 && h > 0 && w > 0, SMCDET_E_INVALID, "smcdet_render: non-positive size");
    const ModelK m = make_model_k(*model);
    const size_t total = (size_t)T * h * w * N;
    SMC_LAUNCH(render_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, m, locs, fluxes, rate, T, N, D, h, w);
    return launch_status("render_kernel");
}

int smcdet_prior_logprob(const smcdet_prior_params* prior, const float* counts, const float* locs,
                         const float* fluxes, float* out, int T, int N, int D, void* stream) {
    DeviceGuard guard(counts);
    SMC_REQUIRE(prior && counts && locs && fluxes && out, SMCDET_E_INVALID, "smcdet_prior_logprob: null pointer");
    SMC_REQUIRE(T > 0 && N > 0 && D > 0, SMCDET_E_INVALID, "smcdet_prior_logprob: non-positive size");
    const size_t TN = (size_t)T * N;
    SMC_LAUNCH(prior_logprob_kernel, (unsigned)((TN + 255) / 256), 256, 0, (cudaStream_t)stream, *prior, counts, locs, fluxes,
                                                                                          out, TN, D);
    return launch_status("prior_logprob_kernel");
}

int smcdet_prior_sample(const smcdet_prior_params* prior, const float* u_locs, const float* u_fluxes, uint64_t seed,
                        const int64_t* tile_ids, float* counts, float* locs, float* fluxes, int T,
                        int num_per_count, int D, void* stream) {
    DeviceGuard guard(counts);
    SMC_REQUIRE(prior && counts && locs && fluxes, SMCDET_E_INVALID, "smcdet_prior_sample: null pointer");
    SMC_REQUIRE((u_locs == nullptr) == (u_fluxes == nullptr), SMCDET_E_INVALID,
                "smcdet_prior_sample: give both uniform tapes or neither");
    SMC_REQUIRE(T > 0 && num_per_count > 0 && D > 0 && prior->max_objects >= prior->min_objects, SMCDET_E_INVALID,
                "smcdet_prior_sample: bad sizes");
    const int M = (prior->max_objects - prior->min_objects + 1) * num_per_count;
    const size_t total = (size_t)T * M * D;
    SMC_LAUNCH(prior_sample_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream,
        *prior, u_locs, u_fluxes, seed, tile_ids, counts, locs, fluxes, T, M, num_per_count, D);
    return launch_status("prior_sample_kernel");
}

int smcdet_temper_update(const float* loglik, float* tau, float* tau_prev, float ess_threshold, int do_temper,
                         float* wlog, float* weights, float* ess, float* logz, int32_t* funcalls,
                         const int32_t* active, const smcdet_loop_state* loop, int T, int N, void* stream) {
    DeviceGuard guard(loglik);
    SMC_REQUIRE(loglik && tau && tau_prev && wlog && weights && ess && logz, SMCDET_E_INVALID,
                "smcdet_temper_update: null pointer");
    SMC_REQUIRE(T > 0 && N > 0, SMCDET_E_INVALID, "smcdet_temper_update: non-positive size");
    smcdet_loop_state ls;
    memset(&ls, 0, sizeof(ls));
    if (loop != nullptr) ls = *loop;
    SMC_LAUNCH(temper_update_kernel, T, kTB, 0, (cudaStream_t)stream, loglik, tau, tau_prev, ess_threshold, do_temper, wlog,
                                                              weights, ess, logz, funcalls, active, ls, N);
    return launch_status("temper_update_kernel");
}

int smcdet_resample(int method, const float* weights, const double* u, uint64_t seed, const int64_t* tile_ids,
                    const int32_t* active, int64_t* index, double* cdf_scratch, int T, int N, void* stream) {
    DeviceGuard guard(weights);
    SMC_REQUIRE(method == SMCDET_RESAMPLE_MULTINOMIAL || method == SMCDET_RESAMPLE_SYSTEMATIC, SMCDET_E_INVALID,
                "smcdet_resample: unknown method");
    SMC_REQUIRE(weights && index && cdf_scratch, SMCDET_E_INVALID, "smcdet_resample: null pointer");
    SMC_REQUIRE(T > 0 && N > 0, SMCDET_E_INVALID, "smcdet_resample: non-positive size");
    const size_t warp_bytes = sizeof(double) * (kTB / 32);
    const size_t smem = warp_bytes + sizeof(double) * (size_t)N;
    if (smem <= 200 * 1024) {
        // the CDF in shared memory; few tiles: the draws of a tile split over several blocks (about two blocks per SM)
        const int slices = (int)std::min<long long>(16, std::max<long long>(1, (2LL * num_sms()) / T));
        SMC_REQUIRE((long long)T * slices < (1LL << 31), SMCDET_E_TOO_LARGE, "smcdet_resample: grid too large");
#ifndef SMC_HOSTSIM
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(resample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(resample_kernel)");
        }
#endif
        SMC_LAUNCH(resample_kernel<true>, (unsigned)(T * slices), kTB, smem, (cudaStream_t)stream, method, weights, u, seed,
                   tile_ids, active, index, cdf_scratch, N, slices);
    } else {
        SMC_LAUNCH(resample_kernel<false>, T, kTB, warp_bytes, (cudaStream_t)stream, method, weights, u, seed, tile_ids, active,
                   index, cdf_scratch, N, 1);
    }
    return launch_status("resample_kernel");
}

int smcdet_gather(const int64_t* index, const float* counts_in, const float* locs_in, const float* fluxes_in,
                  float* counts_out, float* locs_out, float* fluxes_out, const int32_t* tile_mask, int T, int N, int D,
                  void* stream) {
    DeviceGuard guard(index);
    SMC_REQUIRE(index && counts_in && locs_in && fluxes_in && counts_out && locs_out && fluxes_out, SMCDET_E_INVALID,
                "smcdet_gather: null pointer");
    SMC_REQUIRE(T > 0 && N > 0 && D > 0, SMCDET_E_INVALID, "smcdet_gather: non-positive size");
    SMC_REQUIRE(counts_in != counts_out && locs_in != locs_out && fluxes_in != fluxes_out, SMCDET_E_INVALID,
                "smcdet_gather: in-place gather is not supported");
    // 8-byte elements need an even D and 8-byte aligned arrays (every record then is)
    const bool vec = (D % 2 == 0) && (((uintptr_t)locs_in | (uintptr_t)locs_out | (uintptr_t)fluxes_in | (uintptr_t)fluxes_out) % 8 == 0);
    const size_t total = (size_t)T * N * (vec ? D + D / 2 + 1 : 3 * D + 1);
    cudaStream_t st = (cudaStream_t)stream;
#define SMC_GATHER(IDX, VEC) \
    SMC_LAUNCH((gather_kernel<IDX, VEC>), grid_for(total, 256), 256, 0, st, index, counts_in, locs_in, fluxes_in, counts_out, \
               locs_out, fluxes_out, tile_mask, T, N, D)
    // 32-bit element arithmetic with headroom for two grid-stride increments
    if (total < (size_t)3000000000u) { if (vec) SMC_GATHER(uint32_t, 2); else SMC_GATHER(uint32_t, 1); }
    else { if (vec) SMC_GATHER(size_t, 2); else SMC_GATHER(size_t, 1); }
#undef SMC_GATHER
    return launch_status("gather_kernel");
}

static int mutate_impl(bool mala, const smcdet_model_params* model, const smcdet_prior_params* prior, const smcdet_mh_params* mh,
                     const float* tiles, const smcdet_resampled_source* source, const float* counts, float* locs, float* fluxes,
                     const float* tau,
                     float* loglik_out, float* acc_rate, const smcdet_draw_tape* tape, const smcdet_mh_trace* trace,
                     uint64_t seed, uint64_t offset, const int64_t* tile_ids, const int32_t* active, int32_t* status,
                     int T, int N, int D, int h, int w, void* stream) {
    DeviceGuard guard(locs);
    SMC_REQUIRE(model_ok(model) && prior && mh, SMCDET_E_INVALID, "smcdet_mh_mutate: bad parameters");
    SMC_REQUIRE(tiles && (counts || source) && locs && fluxes && tau && acc_rate, SMCDET_E_INVALID,
                "smcdet_mh_mutate: null pointer");
    if (source != nullptr) {
        SMC_REQUIRE(!mala, SMCDET_E_UNSUPPORTED, "smcdet_mh_mutate_resampled: not available for the MALA kernel");
        SMC_REQUIRE(source->index && source->counts && source->locs && source->fluxes && source->counts_out, SMCDET_E_INVALID,
                    "smcdet_mh_mutate_resampled: null pointer in the source");
        SMC_REQUIRE(source->locs != locs && source->fluxes != fluxes && source->counts != source->counts_out, SMCDET_E_INVALID,
                    "smcdet_mh_mutate_resampled: source and destination arrays must differ");
    }
    SMC_REQUIRE(T > 0 && N > 0 && D > 0 && mh->num_iters >= 0, SMCDET_E_INVALID, "smcdet_mh_mutate: bad sizes");
    SMC_REQUIRE(h == w && (h == 8 || h == 16 || h == 32), SMCDET_E_UNSUPPORTED,
                "smcdet_mh_mutate: tile must be 8x8, 16x16 or 32x32");
    SMC_REQUIRE(D <= kMaxStars, SMCDET_E_TOO_LARGE, "smcdet_mh_mutate: too many stars per catalog");
    if (tape != nullptr)
        SMC_REQUIRE(tape->comp && tape->u_loc && tape->u_flux && tape->u_acc, SMCDET_E_INVALID,
                    "smcdet_mh_mutate: incomplete draw tape");
    cudaStream_t st = (cudaStream_t)stream;
    MHArgs a;
    memset(&a, 0, sizeof(a));
    a.m = make_model_k(*model);
    a.pk = make_prior_k(*prior);
    a.count_kind = prior->count_kind; a.min_objects = prior->min_objects; a.max_objects = prior->max_objects;
    a.count_rate = prior->count_rate;
    a.mh = *mh;
    a.tiles = tiles; a.counts = counts; a.locs = locs; a.fluxes = fluxes; a.tau = tau;
    a.loglik_out = loglik_out; a.acc_count = acc_rate;
    if (tape) { a.tape_comp = tape->comp; a.tape_u_loc = tape->u_loc; a.tape_u_flux = tape->u_flux; a.tape_u_acc = tape->u_acc; }
    if (trace) { a.tr_log_alpha = trace->log_alpha; a.tr_target_prop = trace->target_prop; a.tr_accept = trace->accept;
                 a.tr_chain_locs = trace->chain_locs; a.tr_chain_fluxes = trace->chain_fluxes; }
    a.seed = seed; a.offset = offset; a.tile_ids = tile_ids; a.active = active; a.status = status;
    a.tile_map = mh->tile_of_segment;
    a.T = T; a.N = N; a.D = D;
    if (source != nullptr) {
        a.gather_index = source->index; a.counts_src = source->counts; a.locs_src = source->locs;
        a.fluxes_src = source->fluxes; a.counts_out = source->counts_out; a.copy_mask = source->copy_mask;
        // the carried images are those of a fresh render of the final state, which only the refresh pass makes
        if (loglik_out != nullptr && mh->refresh_loglik) {
            SMC_REQUIRE(source->rates == nullptr || source->rates != source->rates_out, SMCDET_E_INVALID,
                        "smcdet_mh_mutate_resampled: rates and rates_out must differ");
            a.rates_src = source->rates; a.rates_out = source->rates_out;
        }
    }
    // acc_as_count: the caller keeps acc_rate zero-filled between launches and divides by N itself
    // (smcdet_temper_update does both through smcdet_loop_state), which saves two small launches per call
    if (!mh->acc_as_count) SMC_LAUNCH(zero_active_kernel, (T + 255) / 256, 256, 0, st, acc_rate, active, T);
    // the decomposition is chosen for the tiles that will really run, when the caller knows how many that is
    const long long live_tiles = (mh->live_tiles_hint > 0 && mh->live_tiles_hint < T) ? mh->live_tiles_hint : T;
    const int tpp = g_force_tpp ? g_force_tpp : choose_tpp(h, live_tiles * N);
    int rc;
    if (model->model_kind == SMCDET_MODEL_M71_NORMAL) rc = dispatch_mh_side<SMCDET_MODEL_M71_NORMAL>(h, tpp, mala, a, st);
    else rc = dispatch_mh_side<SMCDET_MODEL_GAUSS_POISSON>(h, tpp, mala, a, st);
    if (rc != 0 || mh->acc_as_count) return rc;
    SMC_LAUNCH(divide_kernel, (T + 255) / 256, 256, 0, st, acc_rate, active, (float)N, T);
    return launch_status("divide_kernel");
}

int smcdet_mh_mutate(const smcdet_model_params* model, const smcdet_prior_params* prior, const smcdet_mh_params* mh,
                     const float* tiles, const float* counts, float* locs, float* fluxes, const float* tau,
                     float* loglik_out, float* acc_rate, const smcdet_draw_tape* tape, const smcdet_mh_trace* trace,
                     uint64_t seed, uint64_t offset, const int64_t* tile_ids, const int32_t* active, int32_t* status,
                     int T, int N, int D, int h, int w, void* stream) {
    return mutate_impl(false, model, prior, mh, tiles, nullptr, counts, locs, fluxes, tau, loglik_out, acc_rate, tape, trace,
                       seed, offset, tile_ids, active, status, T, N, D, h, w, stream);
}

int smcdet_mh_mutate_resampled(const smcdet_model_params* model, const smcdet_prior_params* prior, const smcdet_mh_params* mh,
                               const float* tiles, const smcdet_resampled_source* source, float* locs, float* fluxes,
                               const float* tau, float* loglik_out, float* acc_rate, const smcdet_draw_tape* tape,
                               const smcdet_mh_trace* trace, uint64_t seed, uint64_t offset, const int64_t* tile_ids,
                               const int32_t* active, int32_t* status, int T, int N, int D, int h, int w, void* stream) {
    if (source == nullptr) return fail(SMCDET_E_INVALID, "smcdet_mh_mutate_resampled: null source");
    return mutate_impl(false, model, prior, mh, tiles, source, nullptr, locs, fluxes, tau, loglik_out, acc_rate, tape, trace,
                       seed, offset, tile_ids, active, status, T, N, D, h, w, stream);
}

int smcdet_mala_mutate(const smcdet_model_params* model, const smcdet_prior_params* prior, const smcdet_mh_params* mh,
                       const float* tiles, const float* counts, float* locs, float* fluxes, const float* tau,
                       float* loglik_out, float* acc_rate, const smcdet_draw_tape* tape, const smcdet_mh_trace* trace,
                       uint64_t seed, uint64_t offset, const int64_t* tile_ids, const int32_t* active, int32_t* status,
                       int T, int N, int D, int h, int w, void* stream) {
    return mutate_impl(true, model, prior, mh, tiles, nullptr, counts, locs, fluxes, tau, loglik_out, acc_rate, tape, trace,
                       seed, offset, tile_ids, active, status, T, N, D, h, w, stream);
}

int smcdet_prune(const float* locs, const float* fluxes, float tile_h, float tile_w, float flux_threshold,
                 int64_t* counts_out, float* locs_out, float* fluxes_out, int T, int N, int D, void* stream) {
    DeviceGuard guard(locs);
    SMC_REQUIRE(locs && fluxes && counts_out && locs_out && fluxes_out, SMCDET_E_INVALID, "smcdet_prune: null pointer");
    SMC_REQUIRE(T > 0 && N > 0 && D > 0, SMCDET_E_INVALID, "smcdet_prune: non-positive size");
    const size_t TN = (size_t)T * N;
    SMC_LAUNCH(prune_kernel, (unsigned)((TN + 255) / 256), 256, 0, (cudaStream_t)stream, locs, fluxes, tile_h, tile_w,
                                                                                  flux_threshold, counts_out, locs_out,
                                                                                  fluxes_out, TN, D);
    return launch_status("prune_kernel");
}

int smcdet_match_catalogs(const float* true_counts, const float* true_locs, const float* true_fluxes,
                          const float* est_counts, const float* est_locs, const float* est_fluxes, const int64_t* index,
                          const float* mag_bins, float locs_tol, float mags_tol, float* true_total, float* true_match,
                          float* est_total, float* est_match, int32_t* status, int T, int n, int M, int Dt, int De,
                          int B, void* stream) {
    DeviceGuard guard(true_counts);
    SMC_REQUIRE(true_counts && true_locs && true_fluxes && est_counts && est_locs && est_fluxes && index && mag_bins,
                SMCDET_E_INVALID, "smcdet_match_catalogs: null input pointer");
    SMC_REQUIRE(true_total && true_match && est_total && est_match, SMCDET_E_INVALID,
                "smcdet_match_catalogs: null output pointer");
    SMC_REQUIRE(T > 0 && n > 0 && M > 0 && Dt > 0 && De > 0 && B > 0, SMCDET_E_INVALID,
                "smcdet_match_catalogs: non-positive size");
    const long long probs = (long long)T * n;
    SMC_REQUIRE(probs < (1LL << 31), SMCDET_E_INVALID, "smcdet_match_catalogs: too many matching problems");
    SMC_LAUNCH(match_kernel, (unsigned)((probs + 63) / 64), 64, 0, (cudaStream_t)stream, true_counts, true_locs,
               true_fluxes, est_counts, est_locs, est_fluxes, index, mag_bins, locs_tol, mags_tol, true_total,
               true_match, est_total, est_match, status, T, n, M, Dt, De, B);
    return launch_status("match_kernel");
}

int smcdet_agg_join(const float* locs, const float* fluxes, int axis, float dim, float* counts_out, float* locs_out,
                    float* fluxes_out, int nH, int nW, int N, int M, void* stream) {
    DeviceGuard guard(locs);
    SMC_REQUIRE(locs && fluxes && counts_out && locs_out && fluxes_out, SMCDET_E_INVALID, "smcdet_agg_join: null pointer");
    SMC_REQUIRE(nH > 0 && nW > 0 && N > 0 && M > 0 && (axis == 0 || axis == 1), SMCDET_E_INVALID,
                "smcdet_agg_join: bad sizes");
    SMC_REQUIRE((axis == 0 ? nH : nW) % 2 == 0, SMCDET_E_INVALID, "smcdet_agg_join: odd number of tiles along the merge axis");
    const size_t total = (size_t)nH * nW * N / 2;
    SMC_LAUNCH(agg_join_kernel, (unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream, locs, fluxes, axis, dim,
               counts_out, locs_out, fluxes_out, nH, nW, N, M);
    return launch_status("agg_join_kernel");
}

int smcdet_agg_unjoin(const float* locs, const float* fluxes, int axis, float half, float* counts_out, float* locs_out,
                      float* fluxes_out, int T, int N, int D, void* stream) {
    DeviceGuard guard(locs);
    SMC_REQUIRE(locs && fluxes && counts_out && locs_out && fluxes_out, SMCDET_E_INVALID, "smcdet_agg_unjoin: null pointer");
    SMC_REQUIRE(T > 0 && N > 0 && D > 0 && (axis == 0 || axis == 1), SMCDET_E_INVALID, "smcdet_agg_unjoin: bad sizes");
    const size_t total = (size_t)T * N;
    SMC_LAUNCH(agg_unjoin_kernel, (unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream, locs, fluxes, axis, half,
               counts_out, locs_out, fluxes_out, T, N, D);
    return launch_status("agg_unjoin_kernel");
}

int smcdet_agg_mutate(const smcdet_model_params* model, const smcdet_prior_params* prior, const smcdet_mh_params* mh,
                      int axis, const float* tiles, const float* counts, float* locs, float* fluxes, const float* tau,
                      float* loglik_diff_out, float* parent_loglik_out, float* child_loglik_out, float* log_target_out,
                      float* acc_rate, const smcdet_draw_tape* tape, const smcdet_mh_trace* trace, uint64_t seed,
                      uint64_t offset, const int64_t* tile_ids, const int32_t* active, int T, int N, int D, int h, int w,
                      void* stream) {
    DeviceGuard guard(locs);
    SMC_REQUIRE(model_ok(model) && prior && mh, SMCDET_E_INVALID, "smcdet_agg_mutate: bad parameters");
    SMC_REQUIRE(tiles && counts && locs && fluxes && tau && acc_rate, SMCDET_E_INVALID, "smcdet_agg_mutate: null pointer");
    SMC_REQUIRE(T > 0 && N > 0 && D > 0 && mh->num_iters >= 0 && (axis == 0 || axis == 1), SMCDET_E_INVALID,
                "smcdet_agg_mutate: bad sizes");
    SMC_REQUIRE(D <= kMaxStars, SMCDET_E_TOO_LARGE, "smcdet_agg_mutate: too many stars per catalog");
    if (tape != nullptr)
        SMC_REQUIRE(tape->comp && tape->u_loc && tape->u_flux && tape->u_acc, SMCDET_E_INVALID,
                    "smcdet_agg_mutate: incomplete draw tape");
    cudaStream_t st = (cudaStream_t)stream;
    MHArgs a;
    memset(&a, 0, sizeof(a));
    a.m = make_model_k(*model);
    a.pk = make_prior_k(*prior);
    a.count_kind = prior->count_kind; a.min_objects = prior->min_objects; a.max_objects = prior->max_objects;
    a.count_rate = prior->count_rate;
    a.mh = *mh;
    a.tiles = tiles; a.counts = counts; a.locs = locs; a.fluxes = fluxes; a.tau = tau;
    a.acc_count = acc_rate;
    if (tape) { a.tape_comp = tape->comp; a.tape_u_loc = tape->u_loc; a.tape_u_flux = tape->u_flux; a.tape_u_acc = tape->u_acc; }
    if (trace) { a.tr_log_alpha = trace->log_alpha; a.tr_target_prop = trace->target_prop; a.tr_accept = trace->accept; }
    a.seed = seed; a.offset = offset; a.tile_ids = tile_ids; a.active = active;
    a.T = T; a.N = N; a.D = D;
    AggOut o;
    o.loglik_diff = loglik_diff_out; o.parent_loglik = parent_loglik_out; o.child_loglik = child_loglik_out;
    o.log_target = log_target_out;
    o.half = axis == 0 ? 0.5f * (float)h : 0.5f * (float)w;
    SMC_LAUNCH(zero_active_kernel, (T + 255) / 256, 256, 0, st, acc_rate, active, T);
    int rc;
    if (model->model_kind == SMCDET_MODEL_M71_NORMAL) rc = dispatch_agg<SMCDET_MODEL_M71_NORMAL>(h, w, axis, a, o, st);
    else rc = dispatch_agg<SMCDET_MODEL_GAUSS_POISSON>(h, w, axis, a, o, st);
    if (rc != 0) return rc;
    SMC_LAUNCH(divide_kernel, (T + 255) / 256, 256, 0, st, acc_rate, active, (float)N, T);
    return launch_status("divide_kernel");
}

}  // extern "C"
