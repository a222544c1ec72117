// Microbenchmark for the Euler-Maruyama table row (open item 5): every lane of a warp needs the SAME
// NK doubles per step.  How much does it cost to fetch them (a) with per-lane global loads (__ldg, what em_run does
// today), (b) from shared memory, (c) as constant-bank operands indexed by a warp-uniform step counter (ULDC / LDCU:
// one value per warp instead of 32)?  Each variant runs the same arithmetic: SPT samples x NK FMAs per step.
// Build: nvcc -arch=sm_100a -O3 -o row_broadcast row_broadcast.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NK = 8, STEPS = 896;         // 896 x 8 x 8 B = 56 KB: fits the 64 KB constant bank
__constant__ double c_tab[STEPS * NK];

template <int MODE, int SPT>
__global__ void __launch_bounds__(128) k(const double *__restrict__ g_tab, double *out, int reps) {
    extern __shared__ double s_tab[];
    if (MODE == 1) {
        for (int i = threadIdx.x; i < STEPS * NK; i += blockDim.x) s_tab[i] = g_tab[i];
        __syncthreads();
    }
    double z[SPT][NK], w[SPT];
#pragma unroll
    for (int q = 0; q < SPT; q++) {
        w[q] = 0.0;
#pragma unroll
        for (int j = 0; j < NK; j++) z[q][j] = 1e-3 * (threadIdx.x + 7 * j + 13 * q);
    }
    for (int r = 0; r < reps; r++) {
        for (int s = 0; s < STEPS; s++) {
            double row[NK];
            if (MODE == 0) {
                const double2 *p = reinterpret_cast<const double2 *>(g_tab + (long long)s * NK);
#pragma unroll
                for (int j = 0; j < NK / 2; j++) { const double2 v = __ldg(p + j); row[2 * j] = v.x; row[2 * j + 1] = v.y; }
            } else if (MODE == 1) {
                const double2 *p = reinterpret_cast<const double2 *>(s_tab + s * NK);
#pragma unroll
                for (int j = 0; j < NK / 2; j++) { const double2 v = p[j]; row[2 * j] = v.x; row[2 * j + 1] = v.y; }
            } else {
#pragma unroll
                for (int j = 0; j < NK; j++) row[j] = c_tab[s * NK + j];
            }
#pragma unroll
            for (int q = 0; q < SPT; q++) {
                double acc = w[q] * 0.999;
#pragma unroll
                for (int j = 0; j < NK; j++) acc = fma(z[q][j], row[j], acc);
                w[q] = acc;
            }
        }
    }
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < SPT; q++) t += w[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE, int SPT>
void run(const char *name, const double *g_tab, double *out, int sms) {
    const int reps = 40, blocks = sms * 3;
    const size_t smem = MODE == 1 ? sizeof(double) * STEPS * NK : 0;
    if (smem) cudaFuncSetAttribute(k<MODE, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, SPT><<<blocks, 128, smem>>>(g_tab, out, 2);
    cudaEventRecord(e0);
    k<MODE, SPT><<<blocks, 128, smem>>>(g_tab, out, reps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)blocks * 128 * reps * STEPS * SPT * (NK + 1);
    printf("%-34s SPT %d  %8.3f ms  %6.2f TFLOP/s (FMA = 2)  err %s\n", name, SPT, ms, 2.0 * fma / ms / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *h = new double[STEPS * NK];
    for (int i = 0; i < STEPS * NK; i++) h[i] = 1e-3 * ((i * 2654435761u) % 1000) - 0.5;
    double *g_tab, *out;
    cudaMalloc(&g_tab, sizeof(double) * STEPS * NK);
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 3 * 128);
    cudaMemcpy(g_tab, h, sizeof(double) * STEPS * NK, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_tab, h, sizeof(double) * STEPS * NK);
    printf("%s, %d SMs, 3 blocks of 128 per SM, %d steps x %d doubles per row\n", p.name, p.multiProcessorCount, STEPS, NK);
    run<0, 1>("global __ldg, per lane", g_tab, out, p.multiProcessorCount);
    run<0, 3>("global __ldg, per lane", g_tab, out, p.multiProcessorCount);
    run<1, 1>("shared memory broadcast", g_tab, out, p.multiProcessorCount);
    run<1, 3>("shared memory broadcast", g_tab, out, p.multiProcessorCount);
    run<2, 1>("constant bank, uniform index", g_tab, out, p.multiProcessorCount);
    run<2, 3>("constant bank, uniform index", g_tab, out, p.multiProcessorCount);
    return 0;
}
