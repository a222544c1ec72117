// expect_prelude.cuh -- first part of the NVRTC translation unit, placed BEFORE the user's model
// source.  The host library injects, ahead of this text:
//   #define B200E_NSTATE / B200E_NPARAM / B200E_NX / B200E_NOUT / B200E_NKKL   (dims)
//   #define B200E_SOLVER   0 Tsit5 | 1 Euler-Maruyama
//   #define B200E_FP32     0 | 1
//   #define B200E_BLOCK    threads per block,  B200E_MINBLOCKS  resident blocks per SM
//   #define B200E_CLIP_DTMAX  1 when the model sets dtmax < t1 - t0
//   #define B200E_NSAVE    number of save times of a time-sampled observable (0: end-state observable)
//   #define B200E_TABLE_PDF   1 when a factor of the density is tabulated
//   #define B200E_DYNAMIC / B200E_REPRO   work distribution (B200E_SCHED_*): dynamic instantiations, reproducible sums
// and the text of b200e_kparams.h.
//
// Dialect of the model source (valid as CUDA device code here and as C in the test oracle):
//   `real` scalars, B200E_FN linkage, B200E_REAL(1.5) literals, math via sin/cos/exp/log/sqrt/
//   fabs/pow/fmin/fmax (CUDA resolves the float overloads in FP32 mode).

#ifndef B200E_DYNAMIC
#define B200E_DYNAMIC 0
#endif
#ifndef B200E_REPRO
#define B200E_REPRO 0
#endif

#if B200E_FP32
typedef float real;
#define B200E_REAL(x) x##f
#else
typedef double real;
#define B200E_REAL(x) x
#endif

#define B200E_FN static __device__ __forceinline__

// arrays of length 0 are ill-formed; dims of 0 (e.g. no parameters) get one dummy slot
#define B200E_DIM(n) ((n) > 0 ? (n) : 1)
