// Counter-based device RNG shared by the stand-alone generators (qb_rng.cu) and the fused
// Liu-West draw+move kernels (qb_resample.cu): Philox4x32-10 (Salmon et al., SC'11).
//
// Element i of stream (seed, offset) uses counter (offset + i/2), key = seed, and is
//   uniform: ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from two of the four 32-bit words of that counter
//            (words 0,1 for even i, words 2,3 for odd i) — NumPy's random_sample construction;
//   normal : Box-Muller on the two uniforms of one counter value (cos branch for even i, sin for odd i).
// Both users go through philox_uniform_pair / philox_normal_pair, so a fused kernel that regenerates
// element i on the fly sees exactly the value qb_rng_uniform / qb_rng_normal would have stored.
#pragma once
#include "qb_common.cuh"

namespace qb {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return (static_cast<double>(a >> 5) * 67108864.0 + static_cast<double>(b >> 6)) * (1.0 / 9007199254740992.0);
}

// the two uniforms of counter value `ctr`
__device__ __forceinline__ void philox_uniform_pair(uint64_t seed, uint64_t ctr, double& a, double& b) {
    uint32_t r[4];
    philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u, 0u, static_cast<uint32_t>(seed),
                  static_cast<uint32_t>(seed >> 32), r);
    a = u53(r[0], r[1]);
    b = u53(r[2], r[3]);
}

// the two standard normals of counter value `ctr`
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t ctr, double& a, double& b) {
    double ua, ub;
    philox_uniform_pair(seed, ctr, ua, ub);
    const double rad = sqrt(-2.0 * log(1.0 - ua));  // 1 - ua in (0, 1]
    double sn, cs;
    sincospi(2.0 * ub, &sn, &cs);
    a = rad * cs;
    b = rad * sn;
}

// element `i` of a stream (for the rare single-element uses: retries, odd alignments)
__device__ __forceinline__ double philox_uniform_elem(uint64_t seed, uint64_t offset, int64_t i) {
    double a, b;
    philox_uniform_pair(seed, offset + static_cast<uint64_t>(i >> 1), a, b);
    return (i & 1) ? b : a;
}
__device__ __forceinline__ double philox_normal_elem(uint64_t seed, uint64_t offset, int64_t i) {
    double a, b;
    philox_normal_pair(seed, offset + static_cast<uint64_t>(i >> 1), a, b);
    return (i & 1) ? b : a;
}

}  // namespace qb
