// fp64 device math for the radio model, sized for the accuracy the path needs (<= 1e-13 relative) instead of the
// 1-ulp general-purpose libm routines (CUDA's pow() alone is ~300 instructions; the whole SNR chain below is ~45).
//
// Table-driven log2 / exp2 with 16-entry tables: a 16 x 8-byte table is exactly one row of the 32 shared-memory
// banks, so ANY per-lane index pattern is conflict-free (equal indices broadcast).  The tables are built once per
// handle with the host libm (dcb_create) and live in shared memory (dcb_math_init copies them in).
//
//   log2(x)   x = 2^e * m, j = top 4 mantissa bits, c_j = 1 + (j + 1/2)/16, r = m * INV[j] - 1, |r| <= 1/32
//             log2(x) = e + L2C[j] + log2(1 + r),   L2C[j] = -log2(INV[j]) with INV[j] = fl(1/c_j)   (consistent)
//             log2(1 + r) = r * (a1 + a2 r + ... + a8 r^7),  truncation (1/32)^9/9/ln2 = 4.5e-15
//   exp2(y)   k = rint(16 y), f = y - k/16, |f| <= 1/32, 2^y = 2^(k >> 4) * EXPT[k & 15] * 2^f
//             2^f = sum_{n<=6} (f ln2)^n / n!,  truncation (ln2/32)^7/7! = 4.4e-16
//   log1p2(s) log2(1 + s) for 0 <= s < 1/32 as the reference rounds it: t = fl(1 + s), s' = t - 1 (exact), series in s'
//   rcp(x)    MUFU.RCP64H seed + 2 Newton steps (<= 1 ulp) for the divisions that need not be IEEE-exact
//
// The polynomials are evaluated as two independent even / odd Horner chains (dependency depth ~5 instead of 8, one
// literal per FMA): the path is issue- and latency-bound at the occupancy a 1024-env batch gives a B200.
#pragma once

#include <cuda_runtime.h>

struct MathTables {
    double inv[16];   // fl(1 / c_j)
    double l2c[16];   // -log2(inv[j])
    double ex2[16];   // 2^(j/16)
    double pwm[16];   // inv[j]^h            = (mantissa segment centre)^(-h)   (dcb_snr_inrange)
    double pwe[16];   // 2^(c0 - h e)        for binary exponents e = 0..15     (dcb_snr_inrange)
};

// The tables are built once per handle on the host (dcb_create: host libm, dcb_host_math_tables) and copied into shared
// memory by the first 80 threads of the CTA, then __syncthreads(): no libm call in any kernel prologue.
#define DCB_MATH_TABLE_DOUBLES 80
#define DCB_VTHR_DOUBLES 16      // snap thresholds of the drawn velocities 0..15 follow the math tables in the host block
__device__ __forceinline__ void dcb_math_init(MathTables *t, double *vthr, int tid, int nthreads, const double *src) {
    for (int j = tid; j < DCB_MATH_TABLE_DOUBLES + DCB_VTHR_DOUBLES; j += nthreads) {
        const double v = src[j];
        if (j < DCB_MATH_TABLE_DOUBLES) reinterpret_cast<double *>(t)[j] = v;
        else if (vthr) vthr[j - DCB_MATH_TABLE_DOUBLES] = v;
    }
}

#define DCB_INV_LN2 1.4426950408889634074
#define DCB_LN2 0.69314718055994530942

// 1/x for a positive normal x, <= 1 ulp
__device__ __forceinline__ double dcb_rcp(double x) {
    double y;
#ifdef __CUDA_ARCH__
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    // host compilation of this header (tests/native/dcb_math_host.cpp: the table math checked on the CPU): a seed of the
    // same ~20-bit quality, so that the two Newton steps below are exercised as on the device
    int ex_;
    const double m_ = frexp(x, &ex_);
    y = ldexp((double)(1.0f / (float)m_), -ex_);
#endif
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// log2(1 + r) for |r| <= 1/32: Taylor series r (c1 + c2 r + ... + c8 r^7), c_k = (-1)^(k+1) / (k ln 2), split into
// even and odd halves, each a Horner chain in r^2.  One literal per FMA (a second one would cost two MOVs), two
// independent chains of depth 3.
__device__ __forceinline__ double dcb_log2_1p_small(double r) {
    const double r2 = r * r;
    double pa = DCB_INV_LN2 / 7.0;                  // c1 + c3 r^2 + c5 r^4 + c7 r^6
    pa = fma(pa, r2, DCB_INV_LN2 / 5.0);
    pa = fma(pa, r2, DCB_INV_LN2 / 3.0);
    pa = fma(pa, r2, DCB_INV_LN2);
    double pb = -DCB_INV_LN2 / 8.0;                 // c2 + c4 r^2 + c6 r^4 + c8 r^6
    pb = fma(pb, r2, -DCB_INV_LN2 / 6.0);
    pb = fma(pb, r2, -DCB_INV_LN2 / 4.0);
    pb = fma(pb, r2, -DCB_INV_LN2 / 2.0);
    return fma(pb, r, pa) * r;
}

// log2 of a positive, normal double
__device__ __forceinline__ double dcb_log2(const MathTables *t, double x) {
    const int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const int j = (hi >> 16) & 15;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double r = fma(m, t->inv[j], -1.0);
    // (double)e without a conversion instruction: 2^52 + 2^31 + e, minus the same constant
    const double ef = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;
    return (ef + t->l2c[j]) + dcb_log2_1p_small(r);
}

// 2^y for |y| < 1000 (result stays a normal double)
__device__ __forceinline__ double dcb_exp2(const MathTables *t, double y) {
    const double magic = 6755399441055744.0;             // 1.5 * 2^52: rint() in the low word
    const double kd = fma(y, 16.0, magic);
    const int k = __double2loint(kd);
    const double f = fma(kd - magic, -0.0625, y);         // exact: |f| <= 1/32
    const double z = f * DCB_LN2;
    const double z2 = z * z;
    double pe = 1.0 / 720.0;                        // 1 + z^2/2 + z^4/24 + z^6/720
    pe = fma(pe, z2, 1.0 / 24.0);
    pe = fma(pe, z2, 0.5);
    pe = fma(pe, z2, 1.0);
    double po = 1.0 / 120.0;                        // 1 + z^2/6 + z^4/120
    po = fma(po, z2, 1.0 / 6.0);
    po = fma(po, z2, 1.0);
    const double p = fma(po, z, pe);
    const double v = t->ex2[k & 15] * p;
    return __hiloint2double(__double2hiint(v) + ((k >> 4) << 20), __double2loint(v));
}

// snr = 2^(c0 - h log2(d2)) for 1 <= d2 < 65536 without the log2 -> exp2 round trip: d2 = 2^e c_j (1 + r) with the
// 16 mantissa segments of dcb_log2, so snr = 2^(c0 - h e) * c_j^(-h) * (1 + r)^(-h): two table entries and the
// binomial series in r (|r| <= 1/32, degree 9: truncation 5e-15; even / odd Horner chains), pw = its coefficients.
// Every link the step kernel evaluates is in range (d2 <= 4750.5), so this is the common path.
__device__ __forceinline__ double dcb_snr_inrange(const MathTables *t, const double *pw, double d2) {
    const int hi = __double2hiint(d2);
    const int lo = __double2loint(d2);
    const int e = ((hi >> 20) - 1023) & 15;
    const int j = (hi >> 16) & 15;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double r = fma(m, t->inv[j], -1.0);
    const double r2 = r * r;
    double pa = pw[8];                              // even powers
    pa = fma(pa, r2, pw[6]);
    pa = fma(pa, r2, pw[4]);
    pa = fma(pa, r2, pw[2]);
    pa = fma(pa, r2, pw[0]);
    double pb = pw[9];                              // odd powers
    pb = fma(pb, r2, pw[7]);
    pb = fma(pb, r2, pw[5]);
    pb = fma(pb, r2, pw[3]);
    pb = fma(pb, r2, pw[1]);
    return (t->pwe[e] * t->pwm[j]) * fma(pb, r, pa);
}

// log2(1 + s) for s >= 0 with the reference's rounding of 1 + s (station.py:137 np.log2(1 + snr))
__device__ __forceinline__ double dcb_log2_1p(const MathTables *t, double s) {
    const double u = 1.0 + s;
    if (s < 0.03125) return dcb_log2_1p_small(u - 1.0);
    return dcb_log2(t, u);
}
