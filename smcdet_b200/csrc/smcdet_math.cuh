// smcdet_math.cuh -- per-thread arithmetic of the smcdet hot path for sm_100a.
//
// Everything here is a __host__ __device__ inline function over plain values and pointers, so
// the kernels in smcdet_kernels.cu are thin loops around it and tests/hostsim can compile the
// very same code with g++ to check the index logic against the CPU oracle without a GPU
// (that host build is test infrastructure; the product library has no CPU path).
//
// Reference semantics are cited as file:line in the reference tree (timwhite0/smcdet).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/smcdet_b200.h"

#if defined(__CUDACC__)
#define SMC_HD __host__ __device__ __forceinline__
#define SMC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SMC_HD inline
#define SMC_HD_NOINLINE inline
#endif

#if !defined(__CUDACC__)
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif

namespace smc {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLogSqrt2Pi = 0.9189385332046727f;
constexpr float kQuadScale = 1.0f / 4096.0f;  // 2^-12: scale of the noise variance carried in ModelK::nas / nms
constexpr float kSqrt2 = 1.4142135623730951f;
constexpr float kInvSqrt2 = 0.7071067811865476f;

// ---------------------------------------------------------------------------------------------
// MUFU wrappers: one SFU instruction each on the device
// ---------------------------------------------------------------------------------------------
SMC_HD float ex2_fast(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}

SMC_HD float lg2_fast(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return log2f(x);
#endif
}

// product that the compiler may not fuse into a following add (keeps the summation tree identical for every
// lanes-per-particle decomposition: a fused multiply-add in one instantiation would round differently)
SMC_HD float mul_unfused(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float p = a * b;
    return p;
#endif
}

SMC_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}

// ---------------------------------------------------------------------------------------------
// Packed pairs: sm_100a executes two independent float32 operations per FFMA2 / FADD2 / FMUL2 instruction
// (a scalar operand is broadcast for free), which halves the issue slots of the per-pixel arithmetic.  Each
// half is the correctly rounded scalar operation, so results are bit-identical to the scalar host build.
// ---------------------------------------------------------------------------------------------
SMC_HD float2 bcast2(float s) { return make_float2(s, s); }

SMC_HD float2 fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

SMC_HD float2 add2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    volatile float x = a.x + b.x, y = a.y + b.y;
    return make_float2(x, y);
#endif
}

SMC_HD float2 sub2(float2 a, float2 b) { return add2(a, make_float2(-b.x, -b.y)); }

SMC_HD float2 mul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    volatile float x = a.x * b.x, y = a.y * b.y;
    return make_float2(x, y);
#endif
}

SMC_HD float2 ex2_fast2(float2 x) { return make_float2(ex2_fast(x.x), ex2_fast(x.y)); }
SMC_HD float2 lg2_fast2(float2 x) { return make_float2(lg2_fast(x.x), lg2_fast(x.y)); }
SMC_HD float2 rcp_fast2(float2 x) { return make_float2(rcp_fast(x.x), rcp_fast(x.y)); }

SMC_HD uint32_t f32_bits(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
#endif
}

SMC_HD float bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float x;
    memcpy(&x, &u, 4);
    return x;
#endif
}

// 2^x on the FMA / ALU pipes instead of the MUFU pipe (Cody-Waite: x = n + f, |f| <= 1/2, 2^f by a degree-5
// minimax polynomial, 2^n added to the exponent field), for a pair of arguments.  The mutation kernel is bound by
// the 16-lane MUFU pipe while the FMA pipe is two-thirds idle, so a fixed share of the power-law wing's ex2 goes
// through here (kSoftPairs).  Relative error <= 2.2e-7 in float32 (ex2.approx: 2 ulp = 2.4e-7); arguments <= -127 (masked pixels
// arrive as -inf) give exactly 0, arguments must stay below 128.
SMC_HD float2 ex2_soft2(float2 x) {
    const float magic = 12582912.0f;  // 1.5 * 2^23: adding it rounds to the nearest integer
    x.x = fmaxf(x.x, -127.0f);
    x.y = fmaxf(x.y, -127.0f);
    const float2 j = add2(x, bcast2(magic));
    const float2 n = add2(j, bcast2(-magic));
    const float2 f = sub2(x, n);
    float2 p = fma2(bcast2(1.327647245e-3f), f, bcast2(9.675540961e-3f));  // minimax on [-1/2, 1/2], 7.5e-8
    p = fma2(p, f, bcast2(5.550713092e-2f));
    p = fma2(p, f, bcast2(2.402212024e-1f));
    p = fma2(p, f, bcast2(6.931469440e-1f));
    p = fma2(p, f, bcast2(1.000000119f));
    // the low bits of j hold n in two's complement: shifted into the exponent field
    return make_float2(bits_f32(f32_bits(p.x) + (f32_bits(j.x) << 23)), bits_f32(f32_bits(p.y) + (f32_bits(j.y) << 23)));
}

SMC_HD float erfinv_f(float y) {
#if defined(__CUDA_ARCH__)
    return erfinvf(y);
#else
    // host build (tests only): Giles' approximation + Newton steps in double
    if (!(y > -1.0f && y < 1.0f)) return (y == 1.0f) ? INFINITY : ((y == -1.0f) ? -INFINITY : NAN);
    double yd = y, w = -log((1.0 - yd) * (1.0 + yd)), x;
    if (w < 5.0) {
        w -= 2.5;
        x = 2.81022636e-08; x = 3.43273939e-07 + x * w; x = -3.5233877e-06 + x * w;
        x = -4.39150654e-06 + x * w; x = 0.00021858087 + x * w; x = -0.00125372503 + x * w;
        x = -0.00417768164 + x * w; x = 0.246640727 + x * w; x = 1.50140941 + x * w;
    } else {
        w = sqrt(w) - 3.0;
        x = -0.000200214257; x = 0.000100950558 + x * w; x = 0.00134934322 + x * w;
        x = -0.00367342844 + x * w; x = 0.00573950773 + x * w; x = -0.0076224613 + x * w;
        x = 0.00943887047 + x * w; x = 1.00167406 + x * w; x = 2.83297682 + x * w;
    }
    x *= yd;
    for (int i = 0; i < 3; ++i) x -= (erf(x) - yd) / (1.1283791670955126 * exp(-x * x));
    return (float)x;
#endif
}

// torch.nan_to_num with default arguments (smcdet/distributions.py:35)
SMC_HD float nan_to_num_f(float x) {
    if (x != x) return 0.0f;
    if (x == INFINITY) return 3.4028234663852886e38f;
    if (x == -INFINITY) return -3.4028234663852886e38f;
    return x;
}

// torch.clamp propagates NaN
SMC_HD float clamp_f(float x, float lo, float hi) {
    if (x != x) return x;
    return x < lo ? lo : (x > hi ? hi : x);
}

// ---------------------------------------------------------------------------------------------
// Image model constants folded once per launch (smcdet/images.py:17, :25-26, :137-145, :162-167)
// ---------------------------------------------------------------------------------------------
struct ModelK {
    int kind;
    float radius;   // psf_radius as float
    float k1, k2;   // exp(-r^2/(2 s)) = ex2(-k r^2)
    float cpl;      // power-law wing: t = 1 + cpl r^2
    float hb;       // -beta/2
    float b, p0;
    float cn;       // PSF normalisation: M71 1/((1+b+p0) Z); Gaussian 1/(stdev sqrt(2pi))
    float c0;       // flux -> weight: cn * adu_per_nmgy (M71), cn (Gaussian)
    float bg, nas, nms, nswitch;  // nas / nms: noise_additive / noise_multiplicative times kQuadScale (M71 model)
    float is1, is2, isp;  // 1/sigma1, 1/sigma2, 1/sigmap (M71); is1 = 1/stdev^2 (Gaussian): PSF gradients (MALA)
    float lp0;            // lg2(p0)
};

inline ModelK make_model_k(const smcdet_model_params& p) {
    ModelK m;
    m.kind = p.model_kind;
    m.radius = (float)p.psf_radius;
    m.bg = p.background;
    m.nswitch = p.normal_switch_rate;
    if (p.model_kind == SMCDET_MODEL_M71_NORMAL) {
        m.k1 = (float)(1.4426950408889634 / (2.0 * (double)p.sigma1));
        m.k2 = (float)(1.4426950408889634 / (2.0 * (double)p.sigma2));
        m.cpl = (float)(1.0 / ((double)p.beta * (double)p.sigmap));
        m.hb = (float)(-0.5 * (double)p.beta);
        m.b = p.b;
        m.p0 = p.p0;
        m.cn = (float)(1.0 / ((1.0 + (double)p.b + (double)p.p0) * (double)p.psf_norm));
        m.c0 = (float)((double)p.adu_per_nmgy / ((1.0 + (double)p.b + (double)p.p0) * (double)p.psf_norm));
        m.nas = p.noise_additive * kQuadScale;   // exact: a power of two
        m.nms = p.noise_multiplicative * kQuadScale;
        m.is1 = (float)(1.0 / (double)p.sigma1); m.is2 = (float)(1.0 / (double)p.sigma2);
        m.isp = (float)(1.0 / (double)p.sigmap);
        m.lp0 = (float)log2((double)p.p0);
    } else {
        double s = (double)p.psf_stdev;
        m.k1 = (float)(1.4426950408889634 / (2.0 * s * s));
        m.k2 = 0.f; m.cpl = 0.f; m.hb = 0.f; m.b = 0.f; m.p0 = 0.f;
        m.cn = (float)(1.0 / (s * 2.5066282746310002));
        m.c0 = m.cn;
        m.nas = 0.f; m.nms = kQuadScale;
        m.is1 = (float)(1.0 / (s * s)); m.is2 = 0.f; m.isp = 0.f; m.lp0 = 0.f;
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// Separable PSF evaluation.
//
// A thread owns RPT consecutive rows [row0, row0+RPT) and all W columns of its particle's tile.
// For one star (l0 = row coordinate, l1 = column coordinate, signed weight wgt = c0*flux):
//   Gaussian terms:  exp(-k (dy^2+dx^2)) = exp(-k dy^2) * exp(-k dx^2)   -> RPT+W ex2 per term
//   power-law wing:  p0 (1 + cpl r^2)^(-beta/2) = ex2(hb*lg2(ay+bx) + lg2(|wgt| p0))
// The (2R+1)^2 patch anchored at floor(loc) (smcdet/images.py:33-43) is separable as well:
// rows/columns outside it get zero Gaussian factors and an infinite wing argument.
// ---------------------------------------------------------------------------------------------
// Column pairs (2 jp, 2 jp + 1) whose power-law wing takes the FMA-pipe exponential (ex2_soft2) instead of two
// MUFU ex2: a 4-bit mask over the four pairs of every group of 8 columns.  It depends on the column only, so a
// pixel's value does not depend on how rows are split over the lanes of a particle.  Tuned on B200.
#ifndef SMC_SOFT_EX2_MASK
#define SMC_SOFT_EX2_MASK 0x1
#endif
constexpr int kSoftEx2Mask = SMC_SOFT_EX2_MASK;
// Separable Gaussian factors of a star (W column + RPT row factors per Gaussian term) on the FMA pipe as well:
// bit 0 = column factors, bit 1 = row factors.
#ifndef SMC_SOFT_GAUSS
#define SMC_SOFT_GAUSS 0
#endif
constexpr int kSoftGauss = SMC_SOFT_GAUSS;

template <int MODEL, int W>
struct ColFactors {
    float2 e1[W / 2];
    float2 e2[MODEL == SMCDET_MODEL_M71_NORMAL ? W / 2 : 1];
    float2 bx[MODEL == SMCDET_MODEL_M71_NORMAL ? W / 2 : 1];
};

// element p of an array of pairs (p is a compile-time constant wherever this is used)
SMC_HD float& pair_elem(float2* v, int p) { return (p & 1) ? v[p >> 1].y : v[p >> 1].x; }
SMC_HD float pair_elem(const float2* v, int p) { return (p & 1) ? v[p >> 1].y : v[p >> 1].x; }

template <int MODEL, int W>
SMC_HD void col_factors(const ModelK& m, float l1, ColFactors<MODEL, W>& c) {
    // columns outside the (2R+1) patch anchored at floor(l1) get squared distance +inf: their Gaussian
    // factors become ex2(-inf) = 0 and their wing argument +inf
    const float lo = floorf(l1) - m.radius, hi = floorf(l1) + m.radius;
#pragma unroll
    for (int jp = 0; jp < W / 2; ++jp) {
        const float j0 = (float)(2 * jp), j1 = (float)(2 * jp + 1);
        const float dx0 = (j0 + 0.5f) - l1, dx1 = (j1 + 0.5f) - l1;
        const float2 d2 = make_float2((j0 >= lo && j0 <= hi) ? dx0 * dx0 : INFINITY,
                                      (j1 >= lo && j1 <= hi) ? dx1 * dx1 : INFINITY);
        c.e1[jp] = (kSoftGauss & 1) ? ex2_soft2(mul2(bcast2(-m.k1), d2)) : ex2_fast2(mul2(bcast2(-m.k1), d2));
        if (MODEL == SMCDET_MODEL_M71_NORMAL) {
            c.e2[jp] = (kSoftGauss & 1) ? ex2_soft2(mul2(bcast2(-m.k2), d2)) : ex2_fast2(mul2(bcast2(-m.k2), d2));
            c.bx[jp] = mul2(bcast2(m.cpl), d2);
        }
    }
}

// acc[(r*W + j)/2] += wgt * psf(star, pixel (row0+r, j)) for the thread's RPT x W pixels (pairs of columns).
template <int MODEL, int RPT, int W>
SMC_HD void star_accumulate(const ModelK& m, float l0, float l1, float wgt, int row0, float2 (&acc)[RPT * W / 2]) {
    ColFactors<MODEL, W> c;
    col_factors<MODEL, W>(m, l1, c);
    const float lo = floorf(l0) - m.radius, hi = floorf(l0) + m.radius;
    const float wb = wgt * m.b, wp = wgt * m.p0;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const float fi = (float)(row0 + r);
        const float dy = (fi + 0.5f) - l0;
        const float d2 = (fi >= lo && fi <= hi) ? dy * dy : INFINITY;
        float g1, g2 = 0.0f;
        if (MODEL == SMCDET_MODEL_M71_NORMAL && (kSoftGauss & 2)) {
            const float2 g = ex2_soft2(make_float2(-m.k1 * d2, -m.k2 * d2));
            g1 = wgt * g.x; g2 = wb * g.y;
        } else {
            g1 = wgt * ex2_fast(-m.k1 * d2);
            if (MODEL == SMCDET_MODEL_M71_NORMAL) g2 = wb * ex2_fast(-m.k2 * d2);
        }
        if (MODEL == SMCDET_MODEL_M71_NORMAL) {
            const float ay = fmaf(m.cpl, d2, 1.0f);
#pragma unroll
            for (int jp = 0; jp < W / 2; ++jp) {
                // wing p0 t^(-beta/2) = ex2(hb lg2 t); masked rows / columns have t = +inf -> exactly 0
                const float2 t = add2(bcast2(ay), c.bx[jp]);
                const float2 arg = mul2(bcast2(m.hb), lg2_fast2(t));
                const float2 pw = ((kSoftEx2Mask >> (jp & 3)) & 1) ? ex2_soft2(arg) : ex2_fast2(arg);
                float2 v = fma2(bcast2(g1), c.e1[jp], acc[r * (W / 2) + jp]);
                v = fma2(bcast2(g2), c.e2[jp], v);
                acc[r * (W / 2) + jp] = fma2(bcast2(wp), pw, v);
            }
        } else {
#pragma unroll
            for (int jp = 0; jp < W / 2; ++jp) acc[r * (W / 2) + jp] = fma2(bcast2(g1), c.e1[jp], acc[r * (W / 2) + jp]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Per-pixel log density, summed over the thread's NPIX pixels.
//   M71 (images.py:169-175): Normal(rate, sqrt(na + nm rate)).log_prob(x); pixels are paired so
//     that two of them share one rcp and one lg2.
//   Gaussian-PSF model (images.py:91-102): Poisson(rate).log_prob(x), Normal(rate, sqrt(rate))
//     where rate > 50000.  lgam[p] = lgamma(x[p]+1).
// rate_at(p) returns the expected count of the thread's p-th pixel.
// ---------------------------------------------------------------------------------------------
// balanced pairwise sum of N = 2^k values: ((v0+v1)+(v2+v3))+...  Together with the xor-butterfly over the lanes
// of a particle this gives ONE summation tree over the tile's rows whatever the number of lanes per particle,
// so results do not depend on how a launch was decomposed (and hence not on how tiles are sharded over GPUs).
template <int N>
SMC_HD float tree_sum(float (&v)[N]) {
#pragma unroll
    for (int s = 1; s < N; s <<= 1) {
#pragma unroll
        for (int i = 0; i + s < N; i += 2 * s) v[i] += v[i + s];
    }
    return v[0];
}

// Per-pixel log density summed over the lane's RPT rows of W pixels; returns the two partial sums (Q, S) that
// finish_loglik combines after the reduction over the lanes of the particle:
//   M71 (images.py:169-175): Q = sum (x-r)^2 / v, S = sum lg2 v, v = (na + nm r) * 2^-12.  Four pixels share one rcp
//     and one lg2; an octet of pixels is handled as two such quads in the two halves of packed (FFMA2) operations:
//     the even pixels of the octet in one half, the odd ones in the other.
//   Gaussian-PSF model (images.py:91-102): Q = sum of Poisson / Normal terms, S = 0; the Normal branch (rate >
//     normal_switch_rate) is only evaluated for rows that hold such a pixel.
// x / lgam: the lane's observed pixels and lgamma(x+1) (16-byte aligned); rate4(g) = expected counts of pixels 4g..4g+3.
template <int MODEL, int RPT, int W, class Rate4>
SMC_HD void pixel_loglik_sum(const ModelK& m, const float* x, const float* lgam, Rate4 rate4, float& Q, float& S) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float qrow[RPT], srow[RPT];
    if (MODEL == SMCDET_MODEL_M71_NORMAL) {
        const float2 nm = bcast2(m.nms), na = bcast2(m.nas), neg1 = bcast2(-1.0f);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            float2 q = make_float2(0.f, 0.f), s = make_float2(0.f, 0.f);
#pragma unroll
            for (int g = r * (W / 4); g < (r + 1) * (W / 4); g += 2) {
                const float4 ra = rate4(g), rb = rate4(g + 1);
                const float4 xa = x4[g], xb = x4[g + 1];
                const float2 r0 = make_float2(ra.x, ra.y), r1 = make_float2(ra.z, ra.w);
                const float2 r2 = make_float2(rb.x, rb.y), r3 = make_float2(rb.z, rb.w);
                // per quad {a, b, c, d}:  sum d_i^2 / v_i = (n_ab v_cd + n_cd v_ab) / (v_ab v_cd),  sum lg2 v_i = lg2(v_ab v_cd)
                // on variances scaled by kQuadScale = 2^-12 (exact), which keeps the products of four far from the
                // float range for any pixel value below ~3e9; finish_loglik undoes the scale
                const float2 va = fma2(nm, r0, na), vb = fma2(nm, r1, na), vc = fma2(nm, r2, na), vd = fma2(nm, r3, na);
                const float2 da = fma2(neg1, r0, make_float2(xa.x, xa.y)), db = fma2(neg1, r1, make_float2(xa.z, xa.w));
                const float2 dc = fma2(neg1, r2, make_float2(xb.x, xb.y)), dd = fma2(neg1, r3, make_float2(xb.z, xb.w));
                const float2 vab = mul2(va, vb), vcd = mul2(vc, vd);
                const float2 nab = fma2(mul2(da, da), vb, mul2(mul2(db, db), va));
                const float2 ncd = fma2(mul2(dc, dc), vd, mul2(mul2(dd, dd), vc));
                const float2 den = mul2(vab, vcd);
                q = fma2(fma2(nab, vcd, mul2(ncd, vab)), rcp_fast2(den), q);
                s = add2(s, lg2_fast2(den));
            }
            qrow[r] = q.x + q.y; srow[r] = s.x + s.y;
        }
        Q = tree_sum<RPT>(qrow);
        S = tree_sum<RPT>(srow);
    } else {
        // Poisson terms (x log r - r) - lgamma(x + 1) of pairs of pixels.  Fast path: no per-pixel selects; a lane
        // whose pixels need the reference's special cases -- a rate above normal_switch_rate (Normal branch,
        // images.py:96-101), or x log r with x = 0 at a vanishing rate (torch.xlogy gives 0) -- sees it in the
        // maximum rate / a non-finite sum and redoes its pixels with the per-pixel selects.  Ordinary pixels get
        // bit-identical terms on both paths, so results do not depend on which lanes took the slow one.
        const float4* l4 = reinterpret_cast<const float4*>(lgam);
        const float2 neg1 = bcast2(-1.0f), ln2 = bcast2(kLn2);
        float rmax = 0.0f;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int g = r * (W / 4); g < (r + 1) * (W / 4); ++g) {
                const float4 rt = rate4(g), xv4 = x4[g], lg4 = l4[g];
                rmax = fmaxf(fmaxf(rmax, fmaxf(rt.x, rt.y)), fmaxf(rt.z, rt.w));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 rv = h ? make_float2(rt.z, rt.w) : make_float2(rt.x, rt.y);
                    const float2 xv = h ? make_float2(xv4.z, xv4.w) : make_float2(xv4.x, xv4.y);
                    const float2 gg = h ? make_float2(lg4.z, lg4.w) : make_float2(lg4.x, lg4.y);
                    const float2 lg = mul2(lg2_fast2(rv), ln2);
                    acc = add2(acc, fma2(neg1, gg, fma2(xv, lg, make_float2(-rv.x, -rv.y))));
                }
            }
            qrow[r] = acc.x + acc.y;
        }
        float total = 0.f;
#pragma unroll
        for (int r = 0; r < RPT; ++r) total += qrow[r];
        if (rmax > m.nswitch || !(fabsf(total) <= 3.4028234663852886e38f)) {
            // (the barrier makes the rare path reload and recompute: nothing of the fast path is kept alive for it)
            asm volatile("" ::: "memory");
#pragma unroll
            for (int r = 0; r < RPT; ++r) {  // (unrolled: the lane's pixels live in registers, which cannot be indexed)
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int g = 0; g < W / 4; ++g) {
                    const float4 rt = rate4(r * (W / 4) + g), xv4 = x4[r * (W / 4) + g], lg4 = l4[r * (W / 4) + g];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float2 rv = h ? make_float2(rt.z, rt.w) : make_float2(rt.x, rt.y);
                        const float2 xv = h ? make_float2(xv4.z, xv4.w) : make_float2(xv4.x, xv4.y);
                        const float2 gg = h ? make_float2(lg4.z, lg4.w) : make_float2(lg4.x, lg4.y);
                        const float2 lg = mul2(lg2_fast2(rv), ln2);
                        float2 xl = fma2(xv, lg, make_float2(-rv.x, -rv.y));   // x log r - r; x = 0: -r (xlogy)
                        xl.x = (xv.x == 0.0f) ? -rv.x : xl.x;
                        xl.y = (xv.y == 0.0f) ? -rv.y : xl.y;
                        float2 term = fma2(neg1, gg, xl);
                        const float2 d = fma2(neg1, rv, xv);
                        const float2 nrm = fma2(mul2(bcast2(-0.5f), mul2(d, d)), rcp_fast2(rv), fma2(bcast2(-0.5f), lg, bcast2(-kLogSqrt2Pi)));
                        term.x = (rv.x > m.nswitch) ? nrm.x : term.x;
                        term.y = (rv.y > m.nswitch) ? nrm.y : term.y;
                        acc = add2(acc, term);
                    }
                }
                qrow[r] = acc.x + acc.y;
            }
        }
        Q = tree_sum<RPT>(qrow);
        S = 0.f;
    }
}

// log-likelihood of the tile from the totals of pixel_loglik_sum over all NPIX_TOTAL pixels
template <int MODEL>
SMC_HD float finish_loglik(float Q, float S, int npix_total) {
    if (MODEL == SMCDET_MODEL_M71_NORMAL)  // Q and S arrive on variances scaled by kQuadScale = 2^-12
        return fmaf(-0.5f * kQuadScale, Q, fmaf(-0.5f * kLn2, S, -(float)npix_total * (kLogSqrt2Pi + 6.0f * kLn2)));
    return Q;
}

// ---------------------------------------------------------------------------------------------
// Gradient pieces for MALA (reference kernel.py:133-275 differentiates log_target with autograd).
//   pixel_dlogpdf : d log p(x | rate) / d rate of one pixel (images.py:91-102, :169-175)
//   star_grad_accumulate : for one star, with per-pixel weights w_p = d loglik / d rate_p (w_row(r, w[W])
//        fills those of the lane's r-th row),
//        sP = sum_p w_p P_p,  s0 = sum_p w_p Q_p dy_p,  s1 = sum_p w_p Q_p dx_p
//     where the star's contribution to the rate is wgt * P (P = e1 + b e2 + p0 t^(-beta/2), un-normalised
//     as in star_accumulate) and dP/dl0 = Q dy, dP/dl1 = Q dx with
//     Q = e1/s1 + b e2/s2 + (p0/sp) t^(-beta/2 - 1)   (Gaussian model: Q = P / stdev^2).
//     Optionally accumulates acc_wgt * P into acc (to remove the star from the rate image in the same pass).
// ---------------------------------------------------------------------------------------------
template <int MODEL>
SMC_HD float pixel_dlogpdf(const ModelK& m, float x, float r) {
    if (MODEL == SMCDET_MODEL_M71_NORMAL) {
        const float nm = m.nms * 4096.0f;
        const float v = fmaf(nm, r, m.nas * 4096.0f), d = x - r, iv = rcp_fast(v);
        return fmaf(0.5f * nm * iv, fmaf(d, d, -v) * iv, d * iv);
    }
    const float ir = rcp_fast(r);
    if (r > m.nswitch) {
        const float d = x - r;
        return fmaf(0.5f * ir, fmaf(d, d, -r) * ir, d * ir);
    }
    return fmaf(x, ir, -1.0f);
}

template <int MODEL, int RPT, int W, bool ACC, class WRow>
SMC_HD void star_grad_accumulate(const ModelK& m, float l0, float l1, float acc_wgt, int row0, WRow w_row,
                                 float2 (&acc)[RPT * W / 2], float& sP, float& s0, float& s1) {
    ColFactors<MODEL, W> c;
    col_factors<MODEL, W>(m, l1, c);
    float dxs[W];
#pragma unroll
    for (int j = 0; j < W; ++j) dxs[j] = ((float)j + 0.5f) - l1;
    const float lo = floorf(l0) - m.radius, hi = floorf(l0) + m.radius;
    float pP[RPT], p0[RPT], p1[RPT];  // per-row partial sums, combined by the same tree as the log-likelihood
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const float fi = (float)(row0 + r);
        const float dy = (fi + 0.5f) - l0;
        const float d2 = (fi >= lo && fi <= hi) ? dy * dy : INFINITY;
        const float g1 = ex2_fast(-m.k1 * d2);
        float rowP = 0.f, rowQ0 = 0.f, row1 = 0.f;  // sums over the row of w P, w Q (times dy afterwards), w Q dx
        float wr[W];        // w_p = d loglik / d rate_p of the row's pixels
        w_row(r, wr);
        if (MODEL == SMCDET_MODEL_M71_NORMAL) {
            const float g2 = m.b * ex2_fast(-m.k2 * d2);
            const float ay = fmaf(m.cpl, d2, 1.0f);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const float t = ay + pair_elem(c.bx, j);
                const float pw = ex2_fast(fmaf(m.hb, lg2_fast(t), m.lp0));
                const float e1 = g1 * pair_elem(c.e1, j), e2 = g2 * pair_elem(c.e2, j);
                const float P = (e1 + e2) + pw;
                const float Q = fmaf(pw * rcp_fast(t), m.isp, fmaf(e2, m.is2, e1 * m.is1));
                const float wp = wr[j];
                if (ACC) pair_elem(acc, r * W + j) = fmaf(acc_wgt, P, pair_elem(acc, r * W + j));
                rowP = fmaf(wp, P, rowP);
                const float wq = wp * Q;
                rowQ0 += wq;
                row1 = fmaf(wq, dxs[j], row1);
            }
        } else {
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const float P = g1 * pair_elem(c.e1, j);
                const float wp = wr[j];
                if (ACC) pair_elem(acc, r * W + j) = fmaf(acc_wgt, P, pair_elem(acc, r * W + j));
                rowP = fmaf(wp, P, rowP);
                const float wq = wp * (P * m.is1);
                rowQ0 += wq;
                row1 = fmaf(wq, dxs[j], row1);
            }
        }
        // a masked row has dy finite but every Q exactly 0
        pP[r] = rowP; p0[r] = mul_unfused(rowQ0, dy); p1[r] = row1;
    }
    sP = tree_sum<RPT>(pP); s0 = tree_sum<RPT>(p0); s1 = tree_sum<RPT>(p1);
}

// ---------------------------------------------------------------------------------------------
// Direct (non-separable) PSF value for the generic kernels (psf / render / arbitrary tile size)
// ---------------------------------------------------------------------------------------------
SMC_HD float psf_direct(const ModelK& m, float l0, float l1, int i, int j) {
    const float di = (float)i - floorf(l0), dj = (float)j - floorf(l1);
    if (!(di >= -m.radius && di <= m.radius && dj >= -m.radius && dj <= m.radius)) return 0.0f;
    const float dy = ((float)i + 0.5f) - l0, dx = ((float)j + 0.5f) - l1;
    const float r2 = fmaf(dy, dy, dx * dx);
    if (m.kind == SMCDET_MODEL_M71_NORMAL) {
        const float t = fmaf(m.cpl, r2, 1.0f);
        const float v = ex2_fast(-m.k1 * r2) + m.b * ex2_fast(-m.k2 * r2) + m.p0 * ex2_fast(m.hb * lg2_fast(t));
        return v;  // un-normalised by c0/A: caller scales
    }
    return ex2_fast(-m.k1 * r2);
}

// ---------------------------------------------------------------------------------------------
// Truncated normal proposal (smcdet/distributions.py:22-52 over torch.distributions.Normal)
// ---------------------------------------------------------------------------------------------
SMC_HD float normal_cdf_f(float x, float mu, float sigma) {
    return 0.5f * (1.0f + erff((x - mu) * (1.0f / sigma) * kInvSqrt2));
}

struct TruncNormal {
    float cdf_lb;
    float mass;      // cdf(ub) - cdf(lb)
    float log_mass;  // nan_to_num(log(mass))
};

// the reference's arithmetic, any box (distributions.py:33-35)
SMC_HD TruncNormal truncnormal_make(float mu, float sigma, float lb, float ub) {
    TruncNormal d;
    d.cdf_lb = normal_cdf_f(lb, mu, sigma);
    d.log_mass = nan_to_num_f(logf(normal_cdf_f(ub, mu, sigma) - d.cdf_lb));
    d.mass = expf(d.log_mass);
    return d;
}

// Box at least 12 sigma wide (every configuration of the reference): at most one bound is within
// 6 sigma of mu, and beyond 5.6 sigma the float32 normal cdf is exactly 0 / 1, so
//   cdf(ub) - cdf(lb) = 1 - q,  q = Phi(-min(mu - lb, ub - mu)/sigma) = 0.5 - 0.5 erf(z / sqrt 2)
// with ONE erf instead of two, to float32 rounding of the reference's expression.
SMC_HD TruncNormal truncnormal_make_wide(float mu, float inv_sigma_sqrt2, float lb, float ub) {
    TruncNormal d;
    const float a = mu - lb, b = ub - mu;
    const float q = 0.5f - 0.5f * erff(fminf(a, b) * inv_sigma_sqrt2);
    d.cdf_lb = (a < b) ? q : 0.0f;
    d.mass = 1.0f - q;
    d.log_mass = lg2_fast(d.mass) * kLn2;  // mu inside the box => mass in [0.5, 1]: abs error <= 2^-22, never 0
    return d;
}

SMC_HD float truncnormal_draw(const TruncNormal& d, float mu, float sigma, float lb, float ub, float u) {
    const float lo = 1e-6f, hi = (float)(1.0 - 1e-6);
    const float p = clamp_f(u, lo, hi);
    const float pt = d.cdf_lb + p * d.mass;  // reference: p * exp(log_prob_in_box)
    const float q = clamp_f(pt, lo, hi);
    const float x = mu + sigma * erfinv_f(2.0f * q - 1.0f) * kSqrt2;
    return clamp_f(x, lb, ub);
}

// log density of x under the truncated normal d with mean mu (distributions.py:50-52)
SMC_HD float truncnormal_logpdf(const TruncNormal& d, float mu, float sigma, float x) {
    const float z = x - mu;
    return -(z * z) / (2.0f * (sigma * sigma)) - logf(sigma) - kLogSqrt2Pi - d.log_mass;
}

// ---------------------------------------------------------------------------------------------
// Prior log density of one catalog (smcdet/prior.py:67-75, :183-189, :220-226).
// star(d, l0, l1, f) yields the d-th star.
// ---------------------------------------------------------------------------------------------
SMC_HD float count_logpmf(const smcdet_prior_params& p, float c) {
    if (p.count_kind == SMCDET_COUNT_POISSON) {
        const float xl = (c == 0.0f) ? 0.0f : c * logf(p.count_rate);
        return xl - p.count_rate - lgammaf(c + 1.0f);
    }
    if (p.count_kind == SMCDET_COUNT_NONE) return 0.0f;
    if (c >= (float)p.min_objects && c <= (float)p.max_objects)
        return logf(1.0f / (float)(p.max_objects - p.min_objects + 1));
    return -INFINITY;
}

SMC_HD float flux_logpdf(const smcdet_prior_params& p, float f) {
    if (p.flux_kind == SMCDET_FLUX_TRUNCATED_PARETO) {
        const float v = (f == 0.0f) ? p.flux_lower : f;
        return p.flux_logpdf_const - (p.flux_alpha + 1.0f) * logf(v);
    }
    if (p.flux_kind == SMCDET_FLUX_PARETO) {
        const float v = (f == 0.0f) ? p.flux_lower : f;
        const float x = logf(v / p.flux_lower);
        return (logf(p.flux_alpha) - p.flux_alpha * x) - x - logf(p.flux_lower);
    }
    const float d = f - p.flux_mean;
    return -(d * d) / (2.0f * p.flux_stdev * p.flux_stdev) - logf(p.flux_stdev) - kLogSqrt2Pi;
}

SMC_HD float loc_logpdf(const smcdet_prior_params& p, float l0, float l1) {
    // torch Uniform.log_prob: log(1[low <= x] * 1[x < high]) - log(high - low), summed over 2 coords
    const float in0 = (p.loc_low[0] <= l0 && p.loc_high[0] > l0) ? 0.0f : -INFINITY;
    const float in1 = (p.loc_low[1] <= l1 && p.loc_high[1] > l1) ? 0.0f : -INFINITY;
    return (in0 - logf(p.loc_high[0] - p.loc_low[0])) + (in1 - logf(p.loc_high[1] - p.loc_low[1]));
}

// Launch constants of the prior folded for the MH kernel.  Both Pareto families have a flux log density of
// the form A - B log f (distributions.py:87-89; torch Pareto = Exponential -> exp -> affine), so one
// branch-free expression serves them; empty slots (f == 0) are evaluated at `repl` (prior.py:187, :224).
struct PriorK {
    float loc_low[2], loc_high[2];
    float loc_norm;          // log(h0-l0) + log(h1-l1)
    float flux_a, flux_b;    // log density = flux_a - flux_b * log f
    float repl;
    int flux_is_normal;
    float flux_mean, flux_inv2var, flux_lognorm;
};

inline PriorK make_prior_k(const smcdet_prior_params& p) {
    PriorK k;
    for (int c = 0; c < 2; ++c) { k.loc_low[c] = p.loc_low[c]; k.loc_high[c] = p.loc_high[c]; }
    k.loc_norm = logf(p.loc_high[0] - p.loc_low[0]) + logf(p.loc_high[1] - p.loc_low[1]);
    k.flux_is_normal = (p.flux_kind == SMCDET_FLUX_NORMAL) ? 1 : 0;
    k.repl = p.flux_lower;
    k.flux_a = k.flux_b = 0.f;
    k.flux_mean = p.flux_mean;
    k.flux_inv2var = 1.0f / (2.0f * p.flux_stdev * p.flux_stdev);
    k.flux_lognorm = logf(p.flux_stdev) + kLogSqrt2Pi;
    if (p.flux_kind == SMCDET_FLUX_TRUNCATED_PARETO) {
        k.flux_a = p.flux_logpdf_const;
        k.flux_b = p.flux_alpha + 1.0f;
    } else if (p.flux_kind == SMCDET_FLUX_PARETO) {
        k.flux_a = (float)(log((double)p.flux_alpha) + (double)p.flux_alpha * log((double)p.flux_lower));
        k.flux_b = p.flux_alpha + 1.0f;
    }
    return k;
}

// One star's contribution to the prior, split into its finite part and an "outside the support of
// the location prior" flag (whose log density is -inf), so that the MH kernel can update the prior
// of a catalog incrementally when a single star moves.
SMC_HD float star_prior_term(const PriorK& p, float l0, float l1, float f, int& bad) {
    const bool inside = (p.loc_low[0] <= l0 && p.loc_high[0] > l0) && (p.loc_low[1] <= l1 && p.loc_high[1] > l1);
    bad = inside ? 0 : 1;
    float t;
    if (p.flux_is_normal) {
        const float d = f - p.flux_mean;
        t = -(d * d) * p.flux_inv2var - p.flux_lognorm;
    } else {
        t = p.flux_a - p.flux_b * (lg2_fast((f == 0.0f) ? p.repl : f) * kLn2);
    }
    return t - p.loc_norm;
}

// d log prior / d flux of one live star (autograd of prior.py:183-189, :220-226, :152-154)
SMC_HD float star_prior_dflux(const PriorK& p, float f) {
    if (p.flux_is_normal) return -(f - p.flux_mean) * (2.0f * p.flux_inv2var);
    return -p.flux_b / ((f == 0.0f) ? p.repl : f);
}

template <class StarAt>
SMC_HD float prior_logprob_catalog(const smcdet_prior_params& p, float count, int D, StarAt star) {
    float loc_acc = 0.f, flux_acc = 0.f;
    for (int d = 0; d < D; ++d) {
        float l0, l1, f;
        star(d, l0, l1, f);
        const float mask = ((float)d < count) ? 1.0f : 0.0f;
        loc_acc += loc_logpdf(p, l0, l1) * mask;  // -inf * 0 = nan, as in the reference
        flux_acc += flux_logpdf(p, f) * mask;
    }
    return (count_logpmf(p, count) + loc_acc) + flux_acc;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011)
// ---------------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t v[4];
};

SMC_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

SMC_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// 24-bit uniform in [0,1), the resolution of torch.rand for float32
SMC_HD float u01_f(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }
// 53-bit uniform in [0,1)
SMC_HD double u01_d(uint32_t hi, uint32_t lo) {
    const uint64_t v = (((uint64_t)hi << 32) | (uint64_t)lo) >> 11;
    return (double)v * 1.1102230246251565e-16;
}

// domain-separation constants for the Philox counter's 4th word
enum : uint32_t {
    kStreamPriorLocs = 0x50524c4fu, kStreamPriorFlux = 0x5052464cu, kStreamResample = 0x52455341u,
    kStreamMHDraws = 0x4d484452u, kStreamMHComp = 0x4d48434fu
};

// ---------------------------------------------------------------------------------------------
// Brent's root finder as an explicit state machine, so that a whole thread block can evaluate
// the objective cooperatively between steps.  Same algorithm, tolerances and tie-breaking as
// scipy.optimize.brentq (scipy/optimize/Zeros/brentq.c), the solver the reference calls at
// smcdet/sampler.py:114-120.
//   Brent b; b.start(xa, xb, fa, fb, xtol, rtol);  while (!b.done) { f = F(b.x); b.step(f); }
// ---------------------------------------------------------------------------------------------
struct Brent {
    double x_prev, x_cur, x_blk, f_prev, f_cur, f_blk, s_prev, s_cur, xtol, rtol;
    double x;  // next abscissa to evaluate, or the root once done
    int done, iters;

    SMC_HD void start(double xa, double xb, double fa, double fb, double xtol_, double rtol_) {
        x_prev = xa; x_cur = xb; f_prev = fa; f_cur = fb;
        x_blk = 0.0; f_blk = 0.0; s_prev = 0.0; s_cur = 0.0;
        xtol = xtol_; rtol = rtol_; done = 0; iters = 0; x = xb;
        if (f_prev == 0.0) { x = x_prev; done = 1; return; }
        if (f_cur == 0.0) { x = x_cur; done = 1; return; }
        if (signbit(f_prev) == signbit(f_cur)) { x = 0.0; done = 2; return; }
        advance();
    }

    SMC_HD void step(double f_new) {
        f_cur = f_new;
        advance();
    }

    SMC_HD void advance() {
        if (iters >= 100) { x = x_cur; done = 3; return; }
        ++iters;
        if (f_prev != 0.0 && f_cur != 0.0 && (signbit(f_prev) != signbit(f_cur))) {
            x_blk = x_prev; f_blk = f_prev;
            s_prev = s_cur = x_cur - x_prev;
        }
        if (fabs(f_blk) < fabs(f_cur)) {
            x_prev = x_cur; x_cur = x_blk; x_blk = x_prev;
            f_prev = f_cur; f_cur = f_blk; f_blk = f_prev;
        }
        const double tol = (xtol + rtol * fabs(x_cur)) / 2;
        const double s_bis = (x_blk - x_cur) / 2;
        if (f_cur == 0.0 || fabs(s_bis) < tol) { x = x_cur; done = 1; return; }
        if (fabs(s_prev) > tol && fabs(f_cur) < fabs(f_prev)) {
            double s_try;
            if (x_prev == x_blk) {
                s_try = -f_cur * (x_cur - x_prev) / (f_cur - f_prev);
            } else {
                const double d_prev = (f_prev - f_cur) / (x_prev - x_cur);
                const double d_blk = (f_blk - f_cur) / (x_blk - x_cur);
                s_try = -f_cur * (f_blk * d_blk - f_prev * d_prev) / (d_blk * d_prev * (f_blk - f_prev));
            }
            const double lim = fmin(fabs(s_prev), 3 * fabs(s_bis) - tol);
            if (2 * fabs(s_try) < lim) { s_prev = s_cur; s_cur = s_try; }
            else { s_prev = s_bis; s_cur = s_bis; }
        } else {
            s_prev = s_bis; s_cur = s_bis;
        }
        x_prev = x_cur; f_prev = f_cur;
        if (fabs(s_cur) > tol) x_cur += s_cur;
        else x_cur += (s_bis > 0 ? tol : -tol);
        x = x_cur;
    }
};

}  // namespace smc
