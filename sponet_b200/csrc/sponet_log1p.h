// log1p with the exact rounding sequence of glibc 2.39's x86-64 FMA build of s_log1p.c
// (the fdlibm algorithm, polynomial split four ways, compiled with FMA contraction).
//
// Why this exists: the reference's `rng.exponential` is numba's 256-layer ziggurat
// (numba/np/random/distributions.py:105-119); its tail branch returns
// `ziggurat_exp_r - log1p(-u)` and that value is added to the event time.  Replay mode promises
// bit-exact event times, so the device needs the same log1p bits the host libm produces.  CUDA's
// own log1p is < 1 ulp but not bit-identical.  Every operation below is spelled out with explicit
// round-to-nearest primitives so neither nvcc nor gcc can re-contract it.
//
// Verified against libm log1p on 5e7 arguments x = -u, u in [0,1) (tests/test_log1p.py re-checks).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define SPL_HD __host__ __device__ __forceinline__
#define SPL_MUL(a, b) __dmul_rn((a), (b))
#define SPL_ADD(a, b) __dadd_rn((a), (b))
#define SPL_SUB(a, b) __dsub_rn((a), (b))
#define SPL_DIV(a, b) __ddiv_rn((a), (b))
#define SPL_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#include <math.h>
#if defined(__CUDACC__)
#define SPL_HD __host__ __device__ inline
#else
#define SPL_HD static inline
#endif
// Host build: compile the including TU with -ffp-contract=off.
#define SPL_MUL(a, b) ((a) * (b))
#define SPL_ADD(a, b) ((a) + (b))
#define SPL_SUB(a, b) ((a) - (b))
#define SPL_DIV(a, b) ((a) / (b))
#define SPL_FMA(a, b, c) fma((a), (b), (c))
#endif

SPL_HD int32_t spl_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return (int32_t)(b >> 32);
#endif
}

SPL_HD double spl_sethi(double x, int32_t h) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(h, __double2loint(x));
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    b = (b & 0xffffffffULL) | ((uint64_t)(uint32_t)h << 32);
    memcpy(&x, &b, 8);
    return x;
#endif
}

// Valid for x in (-1, 0.41]; the ziggurat only ever calls it with x = -u, u in [0, 1).
SPL_HD double sponet_log1p(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01,
                 Lp3 = 2.857142874366239149e-01, Lp4 = 2.222219843214978396e-01,
                 Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
                 Lp7 = 1.479819860511658591e-01;
    double f = 0.0, c = 0.0, u;
    int32_t hx = spl_hi(x), ax = hx & 0x7fffffff, hu = 0, k = 1;
    if (hx < 0x3FDA827A) {
        if (ax >= 0x3ff00000) {  // x <= -1: -inf / nan (never reached from the ziggurat)
            return spl_sethi(0.0, (x == -1.0) ? (int32_t)0xfff00000 : (int32_t)0x7ff80000);
        }
        if (ax < 0x3e200000) {  // |x| < 2^-29
            if (ax < 0x3c900000) return x;
            return SPL_FMA(-0.5, SPL_MUL(x, x), x);
        }
        if (hx > 0 || hx <= (int32_t)0xbfd2bec3) {  // -0.2929 < x < 0.41422
            k = 0;
            f = x;
            hu = 1;
        }
    }
    if (k != 0) {
        u = SPL_ADD(1.0, x);
        hu = spl_hi(u);
        k = (hu >> 20) - 1023;
        c = (k > 0) ? SPL_SUB(1.0, SPL_SUB(u, x)) : SPL_SUB(x, SPL_SUB(u, 1.0));
        c = SPL_DIV(c, u);
        hu &= 0x000fffff;
        if (hu < 0x6a09e) {
            u = spl_sethi(u, hu | 0x3ff00000);
        } else {
            k += 1;
            u = spl_sethi(u, hu | 0x3fe00000);
            hu = (0x00100000 - hu) >> 2;
        }
        f = SPL_SUB(u, 1.0);
    }
    const double dk = (double)k;
    const double hfsq = SPL_MUL(SPL_MUL(0.5, f), f);
    if (hu == 0) {  // |f| < 2^-20
        if (f == 0.0) {
            if (k == 0) return 0.0;
            c = SPL_FMA(dk, ln2_lo, c);
            return SPL_FMA(dk, ln2_hi, c);
        }
        double R = SPL_MUL(hfsq, SPL_FMA(-0.66666666666666666, f, 1.0));
        if (k == 0) return SPL_SUB(f, R);
        double t = SPL_SUB(SPL_SUB(R, SPL_FMA(dk, ln2_lo, c)), f);
        return SPL_FMA(dk, ln2_hi, -t);
    }
    const double s = SPL_DIV(f, SPL_ADD(2.0, f));
    const double z = SPL_MUL(s, s);
    const double R4 = SPL_FMA(z, Lp7, Lp6);
    const double z2 = SPL_MUL(z, z);
    const double R2 = SPL_FMA(z, Lp3, Lp2);
    const double R3 = SPL_FMA(z, Lp5, Lp4);
    const double z4 = SPL_MUL(z2, z2);
    const double z6 = SPL_MUL(z4, z2);
    double R = SPL_FMA(Lp1, z, SPL_MUL(R2, z2));
    R = SPL_FMA(z4, R3, R);
    R = SPL_FMA(R4, z6, R);
    const double sR = SPL_MUL(SPL_ADD(R, hfsq), s);
    if (k == 0) return SPL_SUB(f, SPL_SUB(hfsq, sR));
    double t = SPL_SUB(SPL_SUB(hfsq, SPL_ADD(SPL_FMA(dk, ln2_lo, c), sR)), f);
    return SPL_FMA(dk, ln2_hi, -t);
}
