// Scratch probe (not product): accuracy of CUDA cos() near its zeros vs glibc cosl, for the
// argument pattern of the precession likelihood, with and without --fmad.
#include <cstdio>
#include <cmath>
#include <vector>
#include <cstdlib>
__global__ void k(const double* x, double* y, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = cos(x[i]);
}
int main() {
    const int n = 1 << 22;
    std::vector<double> hx(n), hy(n);
    srand48(1);
    for (int i = 0; i < n; ++i) {
        // half the points uniform on [0, 6e4], half within 1e-6 of a zero of cos
        if (i & 1) hx[i] = drand48() * 6e4;
        else { long kk = lrand48() % 38000; hx[i] = (double)((kk + 0.5L) * 3.14159265358979323846264338327950288L) + (drand48() - 0.5) * 2e-6; }
    }
    double *dx, *dy;
    cudaMalloc(&dx, n * 8); cudaMalloc(&dy, n * 8);
    cudaMemcpy(dx, hx.data(), n * 8, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(dx, dy, n);
    cudaMemcpy(hy.data(), dy, n * 8, cudaMemcpyDeviceToHost);
    double worst_rel = 0, worst_x = 0; int nbad = 0;
    for (int i = 0; i < n; ++i) {
        long double ref = cosl((long double)hx[i]);
        double rel = fabs((double)(((long double)hy[i] - ref) / ref));
        if (rel > worst_rel) { worst_rel = rel; worst_x = hx[i]; }
        if (rel > 1e-15) ++nbad;
    }
    printf("worst rel err %.3e at x=%.17g ; %d of %d above 1e-15\n", worst_rel, worst_x, nbad, n);
    return 0;
}
