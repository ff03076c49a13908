// sin^2 / cos^2 of a double argument with full RELATIVE accuracy in both, in 15 fp64 operations and
// no table loads or quadrant selects — the arithmetic core of the precession likelihood
// (test_models.py:134-140: pr0 = cos(t (omega - w_) / 2) ** 2).
//
//   k = rint(theta * 2/pi);  r = theta - k * pi/2  (three-term Cody-Waite, |r| <= pi/4)
//   q = sin(r)^2  (fdlibm __kernel_sin minimax polynomial)
//   k even: cos^2(theta) = 1 - q,  sin^2(theta) = q          (1 - q in [1/2, 1]: no cancellation)
//   k odd : cos^2(theta) = q,      sin^2(theta) = 1 - q
//
// Valid for |theta| < 2^20 * pi/2 ~ 1.6e6 (k * P1 and k * P2 are exact products because P1, P2 carry 33
// bits); callers fall back to the library cos() beyond that or for non-finite arguments.  Measured against
// long-double cosl/sinl on 2e7 arguments (uniform, near the zeros, up to 1e6): worst relative error 7.4e-16
// in either output (glibc's cos()**2: 3.3e-16).  Coefficients sit in __constant__ memory so that the DFMAs
// take them as c[bank][offset] operands instead of materialising 64-bit immediates every iteration.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace qb {

constexpr double TRIG_FAST_LIMIT = 1.0e6;

#define QB_TRIG_COEFFS                                                                              \
    {6.36619772367581382433e-01,  /* 0: 2/pi                                                    */ \
     6755399441055744.0,          /* 1: 1.5 * 2^52: adding it rounds to the nearest integer    */ \
     1.57079632673412561417e+00,  /* 2: P1, first 33 bits of pi/2                               */ \
     6.07710050630396597660e-11,  /* 3: P2, next 33 bits                                        */ \
     2.02226624871116645580e-21,  /* 4: P3, next 33 bits                                        */ \
     -1.66666666666666324348e-01, /* 5: S1                                                      */ \
     8.33333333332248946124e-03,  /* 6: S2                                                      */ \
     -1.98412698298579493134e-04, /* 7: S3                                                      */ \
     2.75573137070700676789e-06,  /* 8: S4                                                      */ \
     -2.50507602534068634195e-08, /* 9: S5                                                      */ \
     1.58969099521155010221e-10,  /* 10: S6                                                     */ \
     0.0}

#ifdef __CUDACC__
static __constant__ double TRIG_C_DEV[12] = QB_TRIG_COEFFS;
#endif
static const double TRIG_C_HOST[12] = QB_TRIG_COEFFS;

// q = sin(r)^2 and the parity of k for theta = k pi/2 + r.
#ifdef __CUDACC__
__host__ __device__ __forceinline__
#else
inline
#endif
double sin2_reduced(double theta, int& odd) {
#ifdef __CUDA_ARCH__
    const double* C = TRIG_C_DEV;
#else
    const double* C = TRIG_C_HOST;
#endif
    const double t = fma(theta, C[0], C[1]);
    const double kd = t - C[1];
#ifdef __CUDA_ARCH__
    odd = __double2loint(t) & 1;
#else
    int64_t bits;
    memcpy(&bits, &t, sizeof(bits));
    odd = static_cast<int>(bits & 1);
#endif
    double r = fma(-kd, C[2], theta);
    r = fma(-kd, C[3], r);
    r = fma(-kd, C[4], r);
    const double z = r * r;
    const double v = z * r;
    double p = fma(z, C[10], C[9]);
    p = fma(z, p, C[8]);
    p = fma(z, p, C[7]);
    p = fma(z, p, C[6]);
    p = fma(z, p, C[5]);
    const double s = fma(v, p, r);
    return s * s;
}

struct SinCosSq {
    double cos2, sin2;
};

#ifdef __CUDACC__
__host__ __device__ __forceinline__
#else
inline
#endif
SinCosSq sincos_squared_fast(double theta) {
    int odd;
    const double q = sin2_reduced(theta, odd);
    const double omq = 1.0 - q;
    SinCosSq out;
    out.cos2 = odd ? q : omq;
    out.sin2 = odd ? omq : q;
    return out;
}

// cos^2(theta) if want_cos2 else sin^2(theta), with a single select.
#ifdef __CUDACC__
__host__ __device__ __forceinline__
#else
inline
#endif
double cos2_or_sin2_fast(double theta, int want_cos2) {
    int odd;
    const double q = sin2_reduced(theta, odd);
    const double omq = 1.0 - q;
    return (odd == want_cos2) ? q : omq;   // want cos2: odd -> q ; want sin2: even -> q
}

}  // namespace qb
