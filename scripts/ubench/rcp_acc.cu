// accuracy of MUFU.RCP64H / MUFU.RSQ64H seeds and of the refinements built on them
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double seed_rcp(double b) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b)); return r; }
__device__ double seed_rsq(double b) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b)); return r; }
__global__ void k(double *out, int n) {
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double b = 1.0 + (double)i / n * 3.0 + 1e-9 * i;   // [1, 4)
        const double ex = 1.0 / b;
        double r = seed_rcp(b);
        m0 = fmax(m0, fabs(r * b - 1.0));
        double e = fma(-b, r, 1.0);
        double e2 = fma(e, e, e);
        double rc = fma(r, e2, r);                     // cubic
        m1 = fmax(m1, fabs(rc / ex - 1.0));
        double rn = fma(r, e, r);                      // two Newton
        e = fma(-b, rn, 1.0);
        rn = fma(rn, e, rn);
        m2 = fmax(m2, fabs(rn / ex - 1.0));
        double y = seed_rsq(b);
        m3 = fmax(m3, fabs(y * y * b - 1.0));
    }
    atomicMax((unsigned long long *)&out[0], (unsigned long long)__double_as_longlong(m0));
    atomicMax((unsigned long long *)&out[1], (unsigned long long)__double_as_longlong(m1));
    atomicMax((unsigned long long *)&out[2], (unsigned long long)__double_as_longlong(m2));
    atomicMax((unsigned long long *)&out[3], (unsigned long long)__double_as_longlong(m3));
}
int main() {
    double *d, h[4];
    cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
    k<<<592, 256>>>(d, 1 << 26);
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("MUFU.RCP64H seed max rel err %.3e (2^%.1f)\ncubic step        max rel err %.3e\ntwo Newton steps  max rel err %.3e\nMUFU.RSQ64H seed: max |y^2 b - 1| %.3e (2^%.1f)\n",
           h[0], log2(h[0]), h[1], h[2], h[3], log2(h[3]));
    return 0;
}
