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
template <int MODEL, int W>
struct ColFactors {
    float e1[W];
    float e2[MODEL == SMCDET_MODEL_M71_NORMAL ? W : 1];
    float bx[MODEL == SMCDET_MODEL_M71_NORMAL ? W : 1];
};

template <int MODEL, int W>
SMC_HD void col_factors(const ModelK& m, float l1, ColFactors<MODEL, W>& c) {
    // columns outside the (2R+1) patch anchored at floor(l1) get squared distance +inf: their Gaussian
    // factors become ex2(-inf) = 0 and their wing argument +inf
    const float lo = floorf(l1) - m.radius, hi = floorf(l1) + m.radius;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const float dx = ((float)j + 0.5f) - l1;
        const float d2 = ((float)j >= lo && (float)j <= hi) ? dx * dx : INFINITY;
        c.e1[j] = ex2_fast(-m.k1 * d2);
        if (MODEL == SMCDET_MODEL_M71_NORMAL) {
            c.e2[j] = ex2_fast(-m.k2 * d2);
            c.bx[j] = m.cpl * d2;
        }
    }
}

// acc[r*W + j] += wgt * psf(star, pixel (row0+r, j)) for the thread's RPT x W pixels.
template <int MODEL, int RPT, int W>
SMC_HD void star_accumulate(const ModelK& m, float l0, float l1, float wgt, int row0, float (&acc)[RPT * W]) {
    ColFactors<MODEL, W> c;
    col_factors<MODEL, W>(m, l1, c);
    const float lo = floorf(l0) - m.radius, hi = floorf(l0) + m.radius;
    float lw = 0.f, sgn = 1.f;
    if (MODEL == SMCDET_MODEL_M71_NORMAL) {
        lw = lg2_fast(fabsf(wgt) * m.p0);
        sgn = (wgt < 0.f) ? -1.f : 1.f;
    }
    const float wb = wgt * m.b;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const float fi = (float)(row0 + r);
        const float dy = (fi + 0.5f) - l0;
        const float d2 = (fi >= lo && fi <= hi) ? dy * dy : INFINITY;
        const float g1 = wgt * ex2_fast(-m.k1 * d2);
        if (MODEL == SMCDET_MODEL_M71_NORMAL) {
            const float g2 = wb * ex2_fast(-m.k2 * d2);
            const float ay = fmaf(m.cpl, d2, 1.0f);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const float t = ay + c.bx[j];
                const float pw = ex2_fast(fmaf(m.hb, lg2_fast(t), lw));
                float v = fmaf(g1, c.e1[j], acc[r * W + j]);
                v = fmaf(g2, c.e2[j], v);
                acc[r * W + j] = fmaf(sgn, pw, v);
            }
        } else {
#pragma unroll
            for (int j = 0; j < W; ++j) acc[r * W + j] = fmaf(g1, c.e1[j], acc[r * W + j]);
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
//   M71 (images.py:169-175): Q = sum (x-r)^2 / v, S = sum lg2 v, v = (na + nm r) * 2^-12 (four pixels share a rcp and a lg2)
//   Gaussian-PSF model (images.py:91-102): Q = sum of Poisson / Normal terms, S = 0
// x / lgam: the lane's observed pixels and lgamma(x+1) (16-byte aligned); rate4(g) = expected counts of pixels 4g..4g+3.
template <int MODEL, int RPT, int W, class Rate4>
SMC_HD void pixel_loglik_sum(const ModelK& m, const float* x, const float* lgam, Rate4 rate4, float& Q, float& S) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float qrow[RPT], srow[RPT];
    if (MODEL == SMCDET_MODEL_M71_NORMAL) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            float q = 0.f, s = 0.f;
#pragma unroll
            for (int g = r * (W / 4); g < (r + 1) * (W / 4); ++g) {
                const float4 rt = rate4(g);
                const float4 xv = x4[g];
                // four pixels share one rcp and one lg2:
                //   sum d_i^2 / v_i = (n_ab v_cd + n_cd v_ab) / (v_ab v_cd),  sum lg2 v_i = lg2(v_ab v_cd)
                // on variances scaled by kQuadScale = 2^-12 (exact), which keeps the products of four far from the
                // float range for any pixel value below ~3e9; finish_loglik undoes the scale
                const float va = fmaf(m.nms, rt.x, m.nas), vb = fmaf(m.nms, rt.y, m.nas);
                const float vc = fmaf(m.nms, rt.z, m.nas), vd = fmaf(m.nms, rt.w, m.nas);
                const float da = xv.x - rt.x, db = xv.y - rt.y, dc = xv.z - rt.z, dd = xv.w - rt.w;
                const float vab = va * vb, vcd = vc * vd;
                const float nab = fmaf(da * da, vb, (db * db) * va), ncd = fmaf(dc * dc, vd, (dd * dd) * vc);
                const float den = vab * vcd;
                q = fmaf(fmaf(nab, vcd, ncd * vab), rcp_fast(den), q);
                s += lg2_fast(den);
            }
            qrow[r] = q; srow[r] = s;
        }
        Q = tree_sum<RPT>(qrow);
        S = tree_sum<RPT>(srow);
    } else {
        const float4* l4 = reinterpret_cast<const float4*>(lgam);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int g = r * (W / 4); g < (r + 1) * (W / 4); ++g) {
                const float4 r4 = rate4(g);
                const float4 xv4 = x4[g];
                const float4 lg4 = l4[g];
                const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
                const float xx[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
                const float gg[4] = {lg4.x, lg4.y, lg4.z, lg4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float rv = rr[e], xv = xx[e];
                    const float lg = lg2_fast(rv) * kLn2;
                    // both branches are evaluated and selected: a per-pixel branch cost more than the extra rcp
                    const float d = xv - rv;
                    const float normal = fmaf(-0.5f * (d * d), rcp_fast(rv), fmaf(-0.5f, lg, -kLogSqrt2Pi));
                    const float xl = (xv == 0.0f) ? 0.0f : xv * lg;
                    acc += (rv > m.nswitch) ? normal : ((xl - rv) - gg[e]);
                }
            }
            qrow[r] = acc;
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
                                 float (&acc)[RPT * W], float& sP, float& s0, float& s1) {
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
                const float t = ay + c.bx[j];
                const float pw = ex2_fast(fmaf(m.hb, lg2_fast(t), m.lp0));
                const float e1 = g1 * c.e1[j], e2 = g2 * c.e2[j];
                const float P = (e1 + e2) + pw;
                const float Q = fmaf(pw * rcp_fast(t), m.isp, fmaf(e2, m.is2, e1 * m.is1));
                const float wp = wr[j];
                if (ACC) acc[r * W + j] = fmaf(acc_wgt, P, acc[r * W + j]);
                rowP = fmaf(wp, P, rowP);
                const float wq = wp * Q;
                rowQ0 += wq;
                row1 = fmaf(wq, dxs[j], row1);
            }
        } else {
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const float P = g1 * c.e1[j];
                const float wp = wr[j];
                if (ACC) acc[r * W + j] = fmaf(acc_wgt, P, acc[r * W + j]);
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
