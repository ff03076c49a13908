// TomographyModel.canonicalize on the device (SURVEY §8 a10;
// tomography/models.py:149-209): the reference spends 94 % of configuration C4
// here (np.linalg.eig per particle, 100 us each).  One thread owns one particle:
//   rho  = sum_a x_a conj(B_a)                       (models.py:183)
//   rho  = V diag(lambda) V^H  by cyclic complex Jacobi rotations (Hermitian, DIM <= 4)
//   if any lambda < 0:  rho' = V diag(max(lambda,0)) V^H ;  x_a = Re sum_ij B_a[ij] rho'[ij]
//   x /= x_0 sqrt(DIM)   unless allow_subnormalized  (models.py:194-209)
// A particle whose spectrum is already non-negative is passed through bit for bit
// (models.py:185-186) before the renormalising division, like the reference.
#include "qb_common.cuh"

namespace qb {

struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cconj(cplx a) { return {a.re, -a.im}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return {a.re * s, a.im * s}; }

constexpr int TOMO_THREADS = 128;

// rho[r][c] = sum_a x_a conj(B_a[r][c]), then made exactly Hermitian (models.py:183)
template <int DIM>
__device__ __forceinline__ void build_rho(const double* xv, const double* bs, cplx (&A)[DIM][DIM]) {
    constexpr int D2 = DIM * DIM;
#pragma unroll
    for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int a = 0; a < D2; ++a) {
                re = fma(xv[a], bs[((a * DIM + r) * DIM + c) * 2 + 0], re);
                im = fma(xv[a], -bs[((a * DIM + r) * DIM + c) * 2 + 1], im);
            }
            A[r][c] = {re, im};
        }
    // symmetrise the rounding noise so the iteration sees an exactly Hermitian matrix
#pragma unroll
    for (int r = 0; r < DIM; ++r) {
        A[r][r].im = 0.0;
#pragma unroll
        for (int c = r + 1; c < DIM; ++c) {
            const cplx m = {0.5 * (A[r][c].re + A[c][r].re), 0.5 * (A[r][c].im - A[c][r].im)};
            A[r][c] = m;
            A[c][r] = cconj(m);
        }
    }
}

// Screening pass: most particles of a resampled cloud are comfortably positive definite, and for those
// canonicalize is the identity followed by the renormalising division (models.py:185-186, 194-209).  An LDL^H
// factorisation certifies that without an eigendecomposition: if every pivot exceeds 1e-8 * max diagonal the
// smallest eigenvalue is positive beyond any rounding doubt, the particle is renormalised in place and done;
// otherwise it is flagged for the Jacobi kernel below (which decides exactly as before).  One coalesced pass
// over the slab, no local-memory traffic.
template <int DIM>
__global__ void __launch_bounds__(TOMO_THREADS) tomo_screen_kernel(double* __restrict__ x, int64_t n, int ld,
                                                                   const double* __restrict__ basis, int allow_subnorm,
                                                                   uint8_t* __restrict__ flags) {
    constexpr int D2 = DIM * DIM;
    __shared__ double bs[D2 * D2 * 2];
    for (int j = threadIdx.x; j < D2 * D2 * 2; j += TOMO_THREADS) bs[j] = basis[j];
    __syncthreads();
    const double sqrt_dim = sqrt(static_cast<double>(DIM));
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        // the basis is loop-invariant: without this barrier the compiler hoists all 2 D2^2 shared loads out of the
        // particle loop and spills them (3 KB of local memory per thread for DIM = 4)
        asm volatile("" ::: "memory");
        double xv[D2];
#pragma unroll
        for (int a = 0; a < D2; ++a) xv[a] = x[i * ld + a];
        cplx A[DIM][DIM];
        build_rho<DIM>(xv, bs, A);
        double dmax = 0.0;
#pragma unroll
        for (int r = 0; r < DIM; ++r) dmax = fmax(dmax, A[r][r].re);
        const double floor_ = 1e-8 * dmax;
        bool pd = dmax > 0.0 && isfinite(dmax);
        // right-looking LDL^H on the lower triangle: after step k, A[r][c] (r, c > k) holds the Schur complement
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            const double piv = A[k][k].re;
            pd = pd && (piv > floor_);
            const double ipiv = 1.0 / piv;
#pragma unroll
            for (int r = k + 1; r < DIM; ++r) {
                const cplx lrk = cscale(A[r][k], ipiv);
#pragma unroll
                for (int c = k + 1; c <= r; ++c) {
                    const cplx t = cmul(lrk, cconj(A[c][k]));
                    A[r][c].re -= t.re;
                    A[r][c].im -= t.im;
                }
            }
        }
        flags[i] = pd ? 0 : 1;
        if (pd && !allow_subnorm) {
            const double norm = xv[0] * sqrt_dim;
#pragma unroll
            for (int a = 0; a < D2; ++a) x[i * ld + a] = xv[a] / norm;
        }
    }
}

// idxs == NULL: every particle; otherwise the particles idxs[0 .. *count) left over by the screening pass.
template <int DIM>
__global__ void __launch_bounds__(TOMO_THREADS) tomo_canonicalize_kernel(double* __restrict__ x, int64_t n, int ld,
                                                                         const double* __restrict__ basis,
                                                                         int allow_subnorm,
                                                                         const int64_t* __restrict__ idxs,
                                                                         const int64_t* __restrict__ count) {
    constexpr int D2 = DIM * DIM;
    __shared__ double bs[D2 * D2 * 2];  // basis[a][i][j] as (re, im)
    for (int j = threadIdx.x; j < D2 * D2 * 2; j += TOMO_THREADS) bs[j] = basis[j];
    __syncthreads();
    const double sqrt_dim = sqrt(static_cast<double>(DIM));
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t total = (idxs != nullptr) ? *count : n;
    for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r0 < total; r0 += stride) {
        const int64_t i = (idxs != nullptr) ? idxs[r0] : r0;
        asm volatile("" ::: "memory");  // keep the basis in shared memory (see tomo_screen_kernel)
        double xv[D2];
#pragma unroll
        for (int a = 0; a < D2; ++a) xv[a] = x[i * ld + a];

        cplx A[DIM][DIM], V[DIM][DIM];
        build_rho<DIM>(xv, bs, A);
#pragma unroll
        for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int c = 0; c < DIM; ++c) V[r][c] = {(r == c) ? 1.0 : 0.0, 0.0};

        double scale = 0.0;
#pragma unroll
        for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int c = 0; c < DIM; ++c) scale += A[r][c].re * A[r][c].re + A[r][c].im * A[r][c].im;
        const double tol = scale * 1e-34;  // off-diagonal Frobenius^2 threshold ~ (1e-17 ||A||)^2

        for (int sweep = 0; sweep < 16; ++sweep) {
            double off = 0.0;
#pragma unroll
            for (int r = 0; r < DIM; ++r)
#pragma unroll
                for (int c = r + 1; c < DIM; ++c) off += A[r][c].re * A[r][c].re + A[r][c].im * A[r][c].im;
            if (off <= tol) break;
#pragma unroll
            for (int p = 0; p < DIM - 1; ++p)
#pragma unroll
                for (int q = p + 1; q < DIM; ++q) {
                    const cplx apq = A[p][q];
                    const double r2 = apq.re * apq.re + apq.im * apq.im;
                    if (r2 == 0.0) continue;
                    const double r = sqrt(r2);
                    const cplx ph = {apq.re / r, apq.im / r};  // e^{i phi}
                    const double tau = (A[q][q].re - A[p][p].re) / (2.0 * r);
                    const double t = ((tau >= 0.0) ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    const double cth = 1.0 / sqrt(1.0 + t * t);
                    const double sth = t * cth;
                    // G = D R: G_pp = c, G_pq = s, G_qp = -s e^{-i phi}, G_qq = c e^{-i phi}
                    const cplx gqp = cscale(cconj(ph), -sth);
                    const cplx gqq = cscale(cconj(ph), cth);
                    // A <- A G (columns p, q)
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        const cplx akp = A[k][p], akq = A[k][q];
                        A[k][p] = cadd(cscale(akp, cth), cmul(akq, gqp));
                        A[k][q] = cadd(cscale(akp, sth), cmul(akq, gqq));
                        const cplx vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = cadd(cscale(vkp, cth), cmul(vkq, gqp));
                        V[k][q] = cadd(cscale(vkp, sth), cmul(vkq, gqq));
                    }
                    // A <- G^H A (rows p, q)
                    const cplx hpq = cconj(gqp);  // (G^H)_pq = conj(G_qp)
                    const cplx hqq = cconj(gqq);
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        const cplx apk = A[p][k], aqk = A[q][k];
                        A[p][k] = cadd(cscale(apk, cth), cmul(hpq, aqk));
                        A[q][k] = cadd(cscale(apk, sth), cmul(hqq, aqk));
                    }
                    A[p][q] = {0.0, 0.0};
                    A[q][p] = {0.0, 0.0};
                    A[p][p].im = 0.0;
                    A[q][q].im = 0.0;
                }
        }

        bool all_nonneg = true;
#pragma unroll
        for (int k = 0; k < DIM; ++k) all_nonneg = all_nonneg && (A[k][k].re >= 0.0);

        asm volatile("" ::: "memory");
        if (!all_nonneg) {
            cplx N[DIM][DIM];
#pragma unroll
            for (int r = 0; r < DIM; ++r)
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    cplx acc = {0.0, 0.0};
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        const double lam = fmax(A[k][k].re, 0.0);
                        acc = cadd(acc, cscale(cmul(V[r][k], cconj(V[c][k])), lam));
                    }
                    N[r][c] = acc;
                }
#pragma unroll
            for (int a = 0; a < D2; ++a) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < DIM; ++r)
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        const double bre = bs[((a * DIM + r) * DIM + c) * 2 + 0];
                        const double bim = bs[((a * DIM + r) * DIM + c) * 2 + 1];
                        acc += bre * N[r][c].re - bim * N[r][c].im;
                    }
                xv[a] = acc;
            }
        }
        if (!allow_subnorm) {
            const double norm = xv[0] * sqrt_dim;
#pragma unroll
            for (int a = 0; a < D2; ++a) xv[a] = xv[a] / norm;
        }
        if (!all_nonneg || !allow_subnorm) {
#pragma unroll
            for (int a = 0; a < D2; ++a) x[i * ld + a] = xv[a];
        }
    }
}

}  // namespace qb

using namespace qb;

static int tomo_grid(int64_t n) {
    int64_t want = (n + TOMO_THREADS - 1) / TOMO_THREADS;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (want > cap) want = cap;
    return static_cast<int>(want < 1 ? 1 : want);
}

extern "C" int qb_tomo_canonicalize(double* d_x, int64_t n, int32_t dim, const double* d_basis,
                                    int32_t allow_subnormalized, void* stream) {
    return qb_tomo_canonicalize_ld(d_x, n, dim, dim * dim, d_basis, allow_subnormalized, stream);
}

extern "C" int qb_tomo_canonicalize_ld(double* d_x, int64_t n, int32_t dim, int32_t ld, const double* d_basis,
                                       int32_t allow_subnormalized, void* stream) {
    QB_REQUIRE(d_x && d_basis && n >= 1 && ld >= dim * dim, QB_ERR_INVALID_ARGUMENT, "qb_tomo_canonicalize: bad arguments");
    QB_REQUIRE(dim >= 2 && dim <= 4, QB_ERR_UNSUPPORTED_MODEL,
               "qb_tomo_canonicalize: Hilbert-space dimension %d not in {2,3,4}", dim);
    const int grid = tomo_grid(n);
    cudaStream_t st = as_stream(stream);
    if (dim == 2)
        tomo_canonicalize_kernel<2><<<grid, TOMO_THREADS, 0, st>>>(d_x, n, ld, d_basis, allow_subnormalized, nullptr, nullptr);
    else if (dim == 3)
        tomo_canonicalize_kernel<3><<<grid, TOMO_THREADS, 0, st>>>(d_x, n, ld, d_basis, allow_subnormalized, nullptr, nullptr);
    else
        tomo_canonicalize_kernel<4><<<grid, TOMO_THREADS, 0, st>>>(d_x, n, ld, d_basis, allow_subnormalized, nullptr, nullptr);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}

extern "C" int qb_tomo_canonicalize_screened(double* d_x, int64_t n, int32_t dim, const double* d_basis,
                                             int32_t allow_subnormalized, uint8_t* d_flags, int64_t* d_idxs,
                                             int64_t* d_count, void* d_ws, size_t ws_bytes, void* stream) {
    return qb_tomo_canonicalize_screened_ld(d_x, n, dim, dim * dim, d_basis, allow_subnormalized, d_flags, d_idxs,
                                            d_count, d_ws, ws_bytes, stream);
}

extern "C" int qb_tomo_canonicalize_screened_ld(double* d_x, int64_t n, int32_t dim, int32_t ld, const double* d_basis,
                                                int32_t allow_subnormalized, uint8_t* d_flags, int64_t* d_idxs,
                                                int64_t* d_count, void* d_ws, size_t ws_bytes, void* stream) {
    QB_REQUIRE(d_x && d_basis && d_flags && d_idxs && d_count && d_ws && n >= 1 && ld >= dim * dim, QB_ERR_INVALID_ARGUMENT,
               "qb_tomo_canonicalize_screened: bad arguments");
    QB_REQUIRE(dim >= 2 && dim <= 4, QB_ERR_UNSUPPORTED_MODEL,
               "qb_tomo_canonicalize: Hilbert-space dimension %d not in {2,3,4}", dim);
    const int grid = tomo_grid(n);
    cudaStream_t st = as_stream(stream);
    if (dim == 2)
        tomo_screen_kernel<2><<<grid, TOMO_THREADS, 0, st>>>(d_x, n, ld, d_basis, allow_subnormalized, d_flags);
    else if (dim == 3)
        tomo_screen_kernel<3><<<grid, TOMO_THREADS, 0, st>>>(d_x, n, ld, d_basis, allow_subnormalized, d_flags);
    else
        tomo_screen_kernel<4><<<grid, TOMO_THREADS, 0, st>>>(d_x, n, ld, d_basis, allow_subnormalized, d_flags);
    QB_CUDA_CHECK(cudaGetLastError());
    int rc = qb_compact_invalid(d_flags, n, d_idxs, d_count, d_ws, ws_bytes, stream);
    if (rc != QB_OK) return rc;
    // the list length stays on the device: the Jacobi kernel reads it (no host round trip); a quarter of the full
    // grid is plenty for the few per cent of particles that usually remain
    const int g2 = (grid + 3) / 4;
    if (dim == 2)
        tomo_canonicalize_kernel<2><<<g2, TOMO_THREADS, 0, st>>>(d_x, n, ld, d_basis, allow_subnormalized, d_idxs, d_count);
    else if (dim == 3)
        tomo_canonicalize_kernel<3><<<g2, TOMO_THREADS, 0, st>>>(d_x, n, ld, d_basis, allow_subnormalized, d_idxs, d_count);
    else
        tomo_canonicalize_kernel<4><<<g2, TOMO_THREADS, 0, st>>>(d_x, n, ld, d_basis, allow_subnormalized, d_idxs, d_count);
    QB_CUDA_CHECK(cudaGetLastError());
    return QB_OK;
}
