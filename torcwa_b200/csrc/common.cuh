// Shared definitions for the sm_100a RCWA kernels.
//
// Two build modes:
//   * nvcc (product): everything below is device code for sm_100a.
//   * g++ -DRCWA_EMU (tests only): the *single-CTA, phase-structured* kernels (LU panel, QR window
//     chase, small QR, triangular eigenvector solves ...) are compiled as plain C++ where ONE host
//     thread plays the whole CTA (tid = 0, nthreads = 1, barriers are no-ops).  This lets the CPU
//     test-suite exercise the exact control flow / index arithmetic of those kernels without a
//     GPU.  The emulation library is never loaded by the product package.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#ifdef RCWA_EMU
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#define HD inline
#define DEV inline
#define CTA_SYNC() ((void)0)
#define WARP_SYNC() ((void)0)
#else
#include <cuda_runtime.h>
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#define CTA_SYNC() __syncthreads()
#define WARP_SYNC() __syncwarp()
#endif

typedef double2 cplx;

// ------------------------------------------------------------------ complex arithmetic (fp64)
HD cplx C(double re, double im) { return make_double2(re, im); }
HD cplx cadd(cplx a, cplx b) { return C(a.x + b.x, a.y + b.y); }
HD cplx csub(cplx a, cplx b) { return C(a.x - b.x, a.y - b.y); }
HD cplx cneg(cplx a) { return C(-a.x, -a.y); }
HD cplx cconj(cplx a) { return C(a.x, -a.y); }
HD cplx cmul(cplx a, cplx b) { return C(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
HD cplx cmulc(cplx a, cplx b) { /* conj(a)*b */ return C(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
HD cplx cscale(cplx a, double s) { return C(a.x * s, a.y * s); }
HD cplx cfma(cplx a, cplx b, cplx c) { /* a*b + c */
    return C(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
HD double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }
HD double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
HD double cabs_(cplx a) { return hypot(a.x, a.y); }
HD bool cis_zero(cplx a) { return a.x == 0.0 && a.y == 0.0; }
// robust complex division (Smith)
HD cplx cdiv(cplx a, cplx b) {
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, d = b.x + b.y * r;
        return C((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y, d = b.x * r + b.y;
        return C((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}
HD cplx cinv(cplx b) { return cdiv(C(1.0, 0.0), b); }
// principal square root
HD cplx csqrt_(cplx z) {
    double m = hypot(z.x, z.y);
    if (m == 0.0) return C(0.0, 0.0);
    double re, im;
    if (z.x >= 0.0) { re = sqrt(0.5 * (m + z.x)); im = z.y / (2.0 * re); }
    else { im = sqrt(0.5 * (m - z.x)); if (z.y < 0.0) im = -im; re = z.y / (2.0 * im); }
    return C(re, im);
}

#define RCWA_EPS 2.220446049250313e-16        /* fp64 unit roundoff * 2 (LAPACK "precision") */
#define RCWA_SAFMIN 2.2250738585072014e-308

// ------------------------------------------------------------------ CTA context (device or emulated)
struct Cta {
    int tid, nthreads;   // thread index / number of cooperating threads
    int bid;             // which batch entry (matrix) this CTA owns
    char* smem;          // dynamic shared memory base
    int warp_only;       // 1: the cooperating group is ONE warp (tid = lane, nthreads = 32): GROUP_SYNC is __syncwarp
};
// barrier of the cooperating group of a Cta context (whole CTA, or a single warp for the latency-bound
// small-matrix solvers where a 512-thread barrier per Givens rotation would dominate)
#ifdef RCWA_EMU
#define GROUP_SYNC(c) ((void)0)
#else
#define GROUP_SYNC(c) do { if ((c).warp_only) __syncwarp(); else __syncthreads(); } while (0)
#endif

#ifndef RCWA_EMU
DEV Cta make_cta(int bid, char* smem) { Cta c; c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = bid; c.smem = smem; c.warp_only = 0; return c; }

DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DEV double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif

// CTA-wide sum of one double per thread; result returned to every thread.  `scratch` = 33 doubles
// of shared memory.  Contains barriers: must be called by all threads.
DEV double cta_sum(const Cta& c, double v, double* scratch) {
#ifdef RCWA_EMU
    (void)c; (void)scratch; return v;
#else
    v = warp_sum(v);
    int w = c.tid >> 5, l = c.tid & 31, nw = (c.nthreads + 31) >> 5;
    CTA_SYNC();
    if (l == 0) scratch[w] = v;
    CTA_SYNC();
    if (w == 0) { double t = (l < nw) ? scratch[l] : 0.0; t = warp_sum(t); if (l == 0) scratch[32] = t; }
    CTA_SYNC();
    return scratch[32];
#endif
}
DEV double cta_max(const Cta& c, double v, double* scratch) {
#ifdef RCWA_EMU
    (void)c; (void)scratch; return v;
#else
    v = warp_max(v);
    int w = c.tid >> 5, l = c.tid & 31, nw = (c.nthreads + 31) >> 5;
    CTA_SYNC();
    if (l == 0) scratch[w] = v;
    CTA_SYNC();
    if (w == 0) { double t = (l < nw) ? scratch[l] : 0.0; t = warp_max(t); if (l == 0) scratch[32] = t; }
    CTA_SYNC();
    return scratch[32];
#endif
}

// ------------------------------------------------------------------ grouped GEMM problem descriptor
// C(M x N) = alpha * op(A) * op(B) + beta * C, row-major, interleaved complex128.
// M == 0 marks an inactive entry.
#define ZGEMM_B_UPPER 1      // op(B) is upper triangular (K == N frame): output column tile [n0, n0+BN) only needs k < n0 + BN
#define ZGEMM_C_UPPER 2      // only the upper triangle of C is wanted: tiles entirely below the diagonal are skipped (left untouched)
struct ZGemmProblem {
    const cplx* A; const cplx* B; cplx* C;
    int M, N, K;
    int lda, ldb, ldc;
    int flags;               // ZGEMM_* structure hints (0 = dense)
    // ZGEMM_A_BAND / ZGEMM_B_BAND: the small operand U (<= 64 x 64, a QR window unitary: banded, see eig.cu) has
    // nonzeros of its 8-column tile t only in rows [4*klo[t], 4*khi[t]); the kernel skips the other k groups.
    unsigned char klo[8], khi[8];
};
#define ZGEMM_A_BAND 4       // op(A) = U^H (row update): output ROW tile t <-> U's column tile t
#define ZGEMM_B_BAND 8       // B = U (column / Z update): output COLUMN tile t <-> U's column tile t

// status / error codes of the C ABI (LAPACK style: <0 = bad argument #k)
#define RCWA_OK 0
#define RCWA_ERR_CUDA (-1000)
