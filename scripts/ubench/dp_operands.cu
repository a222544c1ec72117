// FP64-pipe microbenchmark 2: does the DFMA issue rate of a B200 SM depend on how many of its
// operands are (distinct) vector registers, and are ALU / IMAD instructions free next to it?
// Build: nvcc -arch=sm_100a -O3 -o dp_operands dp_operands.cu ; the instruction forms are pinned
// with inline PTX and were checked in the SASS (cuobjdump -sass).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b, int ia) {
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x * 1e-3 + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-9 * i + 1e-12 * threadIdx.x; }
    int n[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) n[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(z[i]));      // R,R,R (3 distinct)
                if (MODE == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(z[i]));         // R,UR,R (2 regs)
                if (MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));            // R,UR,UR?? (1 reg)
                if (MODE == 3) asm volatile("fma.rn.f64 %0, %0, %0, %1;" : "+d"(x[i]) : "d"(z[i]));                 // R,R(same),R
                if (MODE == 4) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(y[i]));                     // DMUL R,R
                if (MODE == 5) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(z[i]));                     // DADD R,R
                if (MODE == 6) asm volatile("fma.rn.f64 %0, %0, %1, 0d3FF8000000000000;" : "+d"(x[i]) : "d"(y[i])); // R,R,imm
                if (MODE == 7) {  // R,R,R + 1 LOP3
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(z[i]));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[i]) : "r"(ia));
                }
                if (MODE == 8) {  // R,R,R + 2 LOP3
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(z[i]));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[i]) : "r"(ia));
                    asm volatile("and.b32 %0, %0, %1;" : "+r"(n[i]) : "r"(~ia));
                }
                if (MODE == 9) {  // R,UR,R + 1 LOP3
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(z[i]));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[i]) : "r"(ia));
                }
                if (MODE == 10) {  // R,R,R + 1 IMAD
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(z[i]));
                    asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(n[i]) : "r"(ia));
                }
                if (MODE == 11) {  // R,UR,R + 1 FFMA-pipe-free op: setp+selp
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(z[i]));
                    asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.s32 %0, %0, %1, p;}" : "+r"(n[i]) : "r"(ia));
                }
                if (MODE == 12) {  // R,UR,R + 2 LOP3
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(z[i]));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[i]) : "r"(ia));
                    asm volatile("and.b32 %0, %0, %1;" : "+r"(n[i]) : "r"(~ia));
                }
                if (MODE == 13) {  // R,UR,R + FFMA
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(z[i]));
                    float f = __int_as_float(n[i]);
                    asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f));
                    n[i] = __float_as_int(f);
                }
            }
        }
    }
    double s = 0; int m = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) { s += x[i]; m += n[i]; }
    if (s == 123456.789 || m == 123457) out[0] = s + m;
}

template <int MODE, int ILP>
void run(const char *name, int blocks_per_sm, int threads, int sms, double *d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<MODE, ILP><<<sms * blocks_per_sm, threads>>>(d, iters, 1.0000001, 1e-9, 3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    double dp = (double)iters * 8 * ILP * sms * blocks_per_sm * threads;
    double per_clk = dp / 32 / (best * 1e-3) / 1.965e9 / (sms * 4);
    printf("%-28s warps/SM %2d ILP %d : %.3f DP warp-inst/clk/SMSP  = %.2f clk per DP inst (+ its companions)\n", name,
           blocks_per_sm * threads / 32, ILP, per_clk, 1.0 / per_clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *d; cudaMalloc(&d, 64);
    run<0, 8>("DFMA R,R,R", 8, 256, sms, d);
    run<1, 8>("DFMA R,UR,R", 8, 256, sms, d);
    run<2, 8>("DFMA R,UR,UR", 8, 256, sms, d);
    run<3, 8>("DFMA R,R(same),R", 8, 256, sms, d);
    run<4, 8>("DMUL R,R", 8, 256, sms, d);
    run<5, 8>("DADD R,R", 8, 256, sms, d);
    run<6, 8>("DFMA R,R,imm", 8, 256, sms, d);
    run<7, 8>("DFMA R,R,R + LOP3", 8, 256, sms, d);
    run<8, 8>("DFMA R,R,R + 2 LOP3", 8, 256, sms, d);
    run<9, 8>("DFMA R,UR,R + LOP3", 8, 256, sms, d);
    run<12, 8>("DFMA R,UR,R + 2 LOP3", 8, 256, sms, d);
    run<10, 8>("DFMA R,R,R + IMAD", 8, 256, sms, d);
    run<11, 8>("DFMA R,UR,R + ISETP+SEL", 8, 256, sms, d);
    run<13, 8>("DFMA R,UR,R + FFMA", 8, 256, sms, d);
    run<0, 2>("DFMA R,R,R", 4, 256, sms, d);
    run<1, 2>("DFMA R,UR,R", 4, 256, sms, d);
    run<0, 1>("DFMA R,R,R", 4, 256, sms, d);
    run<1, 1>("DFMA R,UR,R", 4, 256, sms, d);
    run<0, 1>("DFMA R,R,R", 2, 256, sms, d);
    run<0, 1>("DFMA R,R,R", 1, 128, sms, d);
    return 0;
}
