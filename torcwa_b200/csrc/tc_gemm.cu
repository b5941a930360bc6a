// Complex fp64-grade GEMM on the 5th-generation tensor cores (tcgen05, kind::i8, accumulators in TMEM, operands
// staged by TMA) -- "Ozaki scheme": tcgen05 has no f64 kind, but int8 x int8 -> int32 products are EXACT, so a
// product of fixed-point numbers can be assembled from products of their 8-bit digits.
//
//   C (M x N) = alpha * op(A) * op(B) + beta * C,   row-major interleaved complex128,  alpha real.
//
// 1. split (split_rows_kernel / split_cols_kernel):  every row i of op(A) (column j of op(B)) gets a power-of-two
//    scale 2^eA[i] > max_k(|re| + |im|); re, im and re+im are rounded to integers X = rint(x * 2^(8s-2-e)) and
//    written as s balanced radix-256 digits d_0 .. d_{s-1} in [-128, 127] (d_0 most significant), one int8 plane per
//    (component, digit), K contiguous ("K-major"), zero-padded to a multiple of 128.  re+im is split from the exact
//    integer sum X_re + X_im, so the 3-multiplication complex product below is exact in integer arithmetic.
// 2. tc_gemm_kernel: per 128 x 128 output tile and per real product t in {re*re, im*im, (re+im)*(re+im)}:
//        S_d = sum_{p+q=d} A_p B_q^T   (d = 0 .. s-1: the s most significant "levels"; int32, exact)
//    Digits stream through a ring of 16 KB shared-memory slots filled by TMA (cp.async.bulk.tensor, 128-byte swizzle)
//    and are consumed by tcgen05.mma.kind::i8 (M = 128, N <= 128, K = 32) issued by one thread; up to four levels
//    accumulate concurrently in the four 128-column quarters of TMEM, so a digit plane loaded once serves up to four
//    products.  The order of loads / MMAs / slot releases is a host-built table (tc_build_schedule) that producer
//    and issuer walk in lock step.  Eight epilogue warps read the levels with tcgen05.ld, merge them exactly in
//    int64, and keep the running fp64 sum of the tile in registers;  re = P1 - P2, im = P3 - P1 - P2 are combined
//    in C itself (the tile belongs to one CTA).
//    Error: digit products below level s-1 are dropped: norm-wise per row and column scale, |err_ij| / 2^(eA_i + eB_j)
//    = 4e-8 (s = 4), 2e-10 (5), 8e-13 (6), 3e-15 (7), 3e-16 (8, fp64-grade) on random matrices (tests/test_tc_gemm.py).
//
// Replaces the dense torch.matmul products of /root/reference/torcwa/rcwa.py:1236,1264,1276-1281,1291-1294 (and the
// Schur-vector / back-transformation products that stand in for torch.linalg.eig, torch_eig.py:14).
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include "common.cuh"
#include "kernels.h"

namespace rcwa {

// ------------------------------------------------------------------------------------------------ schedule (host)
void tc_build_schedule(int s, int nl, TcSchedule* sch) {
    sch->s = s; sch->ngroups = 0; sch->nops = 0;
    for (int d0 = 0; d0 < s; d0 += nl) {
        TcGroup& g = sch->g[sch->ngroups++];
        const int d1 = (d0 + nl - 1 < s - 1) ? d0 + nl - 1 : s - 1;
        g.d0 = d0; g.nl = d1 - d0 + 1; g.op0 = sch->nops; g.nloads = 0;
        int loaded_b[TC_MAXS]; bool first[4] = {true, true, true, true};
        for (int q = 0; q < TC_MAXS; ++q) loaded_b[q] = -1;
        const int pmax = d1;                                   // d1 <= s - 1
        for (int p = 0; p <= pmax; ++p) {
            const int qlo = (d0 - p > 0) ? d0 - p : 0, qhi = (d1 - p < s - 1) ? d1 - p : s - 1;
            if (qlo > qhi) continue;
            for (int q = qhi; q >= qlo; --q)
                if (loaded_b[q] < 0) { sch->ops[sch->nops++] = TC_OP_LOAD_B | (q << 2); loaded_b[q] = g.nloads++; }
            sch->ops[sch->nops++] = TC_OP_LOAD_A | (p << 2);
            const int ia = g.nloads++;
            for (int q = qhi; q >= qlo; --q) {
                const int l = p + q - d0;
                const int last_p_of_q = (d1 - q < pmax) ? d1 - q : pmax;
                unsigned op = TC_OP_MMA | (ia << 2) | (loaded_b[q] << 7) | (l << 12);
                if (first[l]) op |= 1u << 14;
                if (q == qlo) op |= 1u << 15;                  // last use of A_p
                if (p == last_p_of_q) op |= 1u << 16;          // last use of B_q
                first[l] = false;
                sch->ops[sch->nops++] = op;
            }
        }
        g.nops = sch->nops - g.op0;
    }
}

}  // namespace rcwa

#ifndef RCWA_TC_HOST_ONLY
namespace {
using namespace rcwa;

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 128;          // tile rows / columns, K bytes per ring slot
constexpr int TC_SLOT = TC_BM * TC_BK;                          // 16 KB
constexpr int TC_RING = 12;
constexpr int TC_THREADS = 384;                                  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr int TC_SMEM = TC_RING * TC_SLOT + 1024 /*alignment slack*/ + 512 /*barriers*/;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol error traps (launch failure reported to the host) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, M = 128, N from idesc, K = 32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor: K-major operand tile [rows][128 B], 128-byte swizzle (what TMA SWIZZLE_128B writes):
// start address >> 4 | LBO (unused for swizzled K-major) | SBO = 8 rows * 128 B = 1024 B | version 1 (sm_100) | layout SWIZZLE_128B
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t tc_idesc(int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ double pow2d(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }   // -1022 <= e <= 1023

struct TcParams {
    int M, N, K;            // product shape
    int nkc;                // K chunks of 128 bytes
    int nk_last;            // MMAs (K = 32 each) of the last chunk
    int mt, nt;             // tiles per matrix
    int nmat;               // matrices in this launch
    double alpha; cplx beta;
    cplx* C; int ldc; long long stride_c;
    const int* eA; const int* eB;     // [nmat, M], [nmat, N] exponents
};

// ------------------------------------------------------------------------------------------------ the GEMM kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ TcSchedule sch, const __grid_constant__ TcParams prm) {
    extern __shared__ char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;           // 1024-byte aligned ring (swizzle atom)
    const uint32_t bars = sbase + TC_RING * TC_SLOT;                          // full[RING], empty[RING], tfull, tempty (8 B each), tmem ptr
    const uint32_t bar_full = bars, bar_empty = bars + 8 * TC_RING, bar_tfull = bars + 16 * TC_RING, bar_tempty = bar_tfull + 8;
    const uint32_t tmem_slot = bar_tempty + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = sch.s;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_RING; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    const int items = prm.nmat * prm.mt * prm.nt;

    // register budget: 168 per thread at launch (64 K / 384); the data-movement warpgroup hands its surplus to the epilogue
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory");
    if (warp == 0) {
        // ===================================================== TMA producer (one thread)
        if (lane == 0) {
            unsigned cnt = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int b = item / (prm.mt * prm.nt), rem = item % (prm.mt * prm.nt);
                const int m0 = (rem / prm.nt) * TC_BM, n0 = (rem % prm.nt) * TC_BN;
                for (int t = 0; t < 3; ++t) {
                    const int plane0 = (b * 3 + t) * s;
                    for (int g = 0; g < sch.ngroups; ++g) {
                        const int op0 = sch.g[g].op0, op1 = op0 + sch.g[g].nops;
                        for (int kc = 0; kc < prm.nkc; ++kc) {
                            for (int o = op0; o < op1; ++o) {
                                const unsigned op = sch.ops[o];
                                const unsigned type = op & 3u;
                                if (type == TC_OP_MMA) continue;
                                const unsigned slot = cnt % TC_RING, par = (cnt / TC_RING) & 1u;
                                mbar_wait(bar_empty + 8 * slot, par ^ 1u);
                                mbar_expect_tx(bar_full + 8 * slot, TC_SLOT);
                                const int slice = (op >> 2) & 15;
                                if (type == TC_OP_LOAD_A) tma_load_3d(sbase + slot * TC_SLOT, &mapA, bar_full + 8 * slot, kc * TC_BK, m0, plane0 + slice);
                                else tma_load_3d(sbase + slot * TC_SLOT, &mapB, bar_full + 8 * slot, kc * TC_BK, n0, plane0 + slice);
                                ++cnt;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            unsigned base = 0, acq = 0, grp = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int rem = item % (prm.mt * prm.nt);
                const int n0 = (rem % prm.nt) * TC_BN;
                int ncols = prm.N - n0; if (ncols > TC_BN) ncols = TC_BN;
                const uint32_t idesc = tc_idesc((ncols + 15) & ~15);
                for (int t = 0; t < 3; ++t) {
                    for (int g = 0; g < sch.ngroups; ++g, ++grp) {
                        // the epilogue must have drained the accumulators of the previous group
                        mbar_wait(bar_tempty, (grp & 1u) ^ 1u);
                        tc_fence_after();
                        const int op0 = sch.g[g].op0, op1 = op0 + sch.g[g].nops, nloads = sch.g[g].nloads;
                        for (int kc = 0; kc < prm.nkc; ++kc) {
                            const int nk = (kc == prm.nkc - 1) ? prm.nk_last : TC_BK / 32;
                            for (int o = op0; o < op1; ++o) {
                                const unsigned op = sch.ops[o];
                                if ((op & 3u) != TC_OP_MMA) continue;
                                const unsigned ia = base + ((op >> 2) & 31u), ib = base + ((op >> 7) & 31u);
                                const unsigned need = (ia > ib ? ia : ib);
                                while (acq <= need) { mbar_wait(bar_full + 8 * (acq % TC_RING), (acq / TC_RING) & 1u); ++acq; }
                                tc_fence_after();
                                const uint32_t sa = ia % TC_RING, sb = ib % TC_RING;
                                const uint64_t da = tc_smem_desc(sbase + sa * TC_SLOT), db = tc_smem_desc(sbase + sb * TC_SLOT);
                                const uint32_t dcol = tmem_base + ((op >> 12) & 3u) * TC_BN;
                                const bool fresh = (kc == 0) && ((op >> 14) & 1u);
                                for (int k = 0; k < nk; ++k)
                                    tc_mma_i8(dcol, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (fresh && k == 0) ? 0u : 1u);
                                if ((op >> 15) & 1u) tc_commit(bar_empty + 8 * sa);
                                if ((op >> 16) & 1u) tc_commit(bar_empty + 8 * sb);
                            }
                            base += nloads;
                        }
                        tc_commit(bar_tfull);
                    }
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
        // ===================================================== epilogue: 8 warps; warp%4 = TMEM lane quarter, (warp-4)/4 = column half
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        const int lrow = quarter * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * 64;
        unsigned grp = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int b = item / (prm.mt * prm.nt), rem = item % (prm.mt * prm.nt);
            const int m0 = (rem / prm.nt) * TC_BM, n0 = (rem % prm.nt) * TC_BN;
            const int row = m0 + lrow, col0 = n0 + half * 64;
            const bool row_ok = row < prm.M;
            cplx* crow = prm.C + (size_t)b * prm.stride_c + (size_t)(row_ok ? row : 0) * prm.ldc + col0;
            const int* eB = prm.eB + (size_t)b * prm.N + col0;
            int ea = row_ok ? prm.eA[(size_t)b * prm.M + row] : 0;
            const double fa = pow2d(ea - 6) * prm.alpha;
            for (int t = 0; t < 3; ++t) {
                double acc[64];
#pragma unroll
                for (int c = 0; c < 64; ++c) acc[c] = 0.0;
                for (int g = 0; g < sch.ngroups; ++g, ++grp) {
                    const int nl = sch.g[g].nl;
                    const double w = pow2d(-8 * (sch.g[g].d0 + nl - 1));
                    mbar_wait(bar_tfull, grp & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        int r0[8], r1[8], r2[8], r3[8];
                        tmem_ld8(tlane + cc * 8, r0);
                        if (nl > 1) tmem_ld8(tlane + TC_BN + cc * 8, r1);
                        if (nl > 2) tmem_ld8(tlane + 2 * TC_BN + cc * 8, r2);
                        if (nl > 3) tmem_ld8(tlane + 3 * TC_BN + cc * 8, r3);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            long long v = r0[c];
                            if (nl > 1) v = v * 256 + r1[c];
                            if (nl > 2) v = v * 256 + r2[c];
                            if (nl > 3) v = v * 256 + r3[c];
                            acc[cc * 8 + c] = fma((double)v, w, acc[cc * 8 + c]);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty);
                }
                // fold product t into C:  t = 0: C = beta C + alpha (P1, -P1);  t = 1: C -= alpha (P2, P2);  t = 2: Im C += alpha P3
                if (row_ok) {
#pragma unroll
                    for (int c = 0; c < 64; ++c) {
                        if (col0 + c < prm.N) {
                            const double v = acc[c] * fa * pow2d(eB[c] - 6);
                            cplx o;
                            if (t == 0) {
                                o = C(v, -v);
                                if (prm.beta.x != 0.0 || prm.beta.y != 0.0) o = cadd(o, cmul(prm.beta, crow[c]));
                            } else if (t == 1) {
                                o = crow[c]; o.x -= v; o.y -= v;
                            } else {
                                o = crow[c]; o.y += v;
                            }
                            crow[c] = o;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ split kernels
// digits of one integer: d[s-1] least significant ... d[0] most significant, each in [-128, 127]
template <typename F>
__device__ __forceinline__ void tc_digits(long long x, int s, F&& put) {
    for (int j = s - 1; j >= 0; --j) {
        const int d = (int)(signed char)(x & 0xFF);
        put(j, d);
        x = (x - d) >> 8;
    }
}
__device__ __forceinline__ int tc_exponent(double mx) {      // 2^e > mx ; e clamped so that pow2d stays normal
    if (!(mx > 0.0)) return TC_E_ZERO;
    int e = ilogb(mx) + 1;
    if (e < -900) return TC_E_ZERO;
    return e;
}

// operand whose rows are contiguous in memory: src [R rows][Kc] (ld), one CTA per row.
// out: planes [(b*3 + comp)*s + digit][R][Kp] int8; ex[b*R + r]
__global__ void __launch_bounds__(256)
tc_split_rows_kernel(const cplx* __restrict__ src, int ld, long long stride, int R, int Kc, int Kp, int s, int conj,
                     signed char* __restrict__ out, int* __restrict__ ex) {
    const int r = blockIdx.x, b = blockIdx.y;
    const cplx* row = src + (size_t)b * stride + (size_t)r * ld;
    __shared__ double red[9];
    double mx = 0.0;
    for (int k = threadIdx.x; k < Kc; k += blockDim.x) { const cplx v = row[k]; mx = fmax(mx, fabs(v.x) + fabs(v.y)); }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = (threadIdx.x < 8) ? red[threadIdx.x] : 0.0;
        t = warp_max(t);
        if (threadIdx.x == 0) red[8] = t;
    }
    __syncthreads();
    const int e = tc_exponent(red[8]);
    if (threadIdx.x == 0) ex[(size_t)b * R + r] = (e == TC_E_ZERO) ? 0 : e;
    const double scale = (e == TC_E_ZERO) ? 0.0 : pow2d(8 * s - 2 - e > 1023 ? 1023 : 8 * s - 2 - e);
    const size_t plane = (size_t)R * Kp;
    signed char* o = out + ((size_t)b * 3 * s) * plane + (size_t)r * Kp;
    for (int k4 = threadIdx.x * 4; k4 < Kp; k4 += blockDim.x * 4) {
        unsigned pk[3][TC_MAXS];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < TC_MAXS; ++j) pk[c][j] = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k4 + i;
            long long xr = 0, xi = 0;
            if (k < Kc) { const cplx v = row[k]; xr = __double2ll_rn(v.x * scale); xi = __double2ll_rn(v.y * scale); if (conj) xi = -xi; }
            const long long xs = xr + xi;
            long long x3[3] = {xr, xi, xs};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                long long x = x3[c];
#pragma unroll
                for (int j = TC_MAXS - 1; j >= 0; --j) {
                    if (j < s) {
                        const int d = (int)(signed char)(x & 0xFF);
                        pk[c][j] |= (unsigned)(d & 0xFF) << (8 * i);
                        x = (x - d) >> 8;
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < TC_MAXS; ++j)
                if (j < s) *reinterpret_cast<unsigned*>(o + ((size_t)c * s + j) * plane + k4) = pk[c][j];
    }
}

// column maxima of src [Kc rows][R cols]: thread per column (coalesced across the warp)
__global__ void __launch_bounds__(128)
tc_colmax_kernel(const cplx* __restrict__ src, int ld, long long stride, int R, int Kc, int* __restrict__ ex) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (j >= R) return;
    const cplx* p = src + (size_t)b * stride + j;
    double mx = 0.0;
#pragma unroll 4
    for (int k = 0; k < Kc; ++k) { const cplx v = p[(size_t)k * ld]; mx = fmax(mx, fabs(v.x) + fabs(v.y)); }
    const int e = tc_exponent(mx);
    ex[(size_t)b * R + j] = e;          // TC_E_ZERO marks an all-zero column until the split kernel has consumed it
}

// operand whose "rows" are the COLUMNS of src [Kc rows][R cols] (ld): tile of 32 columns x 64 k, transposed through
// shared memory.  ex holds the exponents from tc_colmax_kernel (TC_E_ZERO -> stored back as 0 by tc_fix_exponent_kernel).
constexpr int TCS_J = 32, TCS_K = 64, TCS_LDW = 17;   // smem row: 64 bytes + 4 pad = 17 words
__global__ void __launch_bounds__(256)
tc_split_cols_kernel(const cplx* __restrict__ src, int ld, long long stride, int R, int Kc, int Kp, int s, int conj,
                     signed char* __restrict__ out, const int* __restrict__ ex) {
    extern __shared__ unsigned tcs_smem[];          // [3*s][TCS_J][TCS_LDW] words
    signed char* sb = reinterpret_cast<signed char*>(tcs_smem);
    const int j0 = blockIdx.x * TCS_J, k0 = blockIdx.y * TCS_K, b = blockIdx.z;
    const int tj = threadIdx.x & 31, tk = threadIdx.x >> 5;
    const int j = j0 + tj;
    int e = TC_E_ZERO;
    if (j < R) e = ex[(size_t)b * R + j];
    const double scale = (e == TC_E_ZERO) ? 0.0 : pow2d(8 * s - 2 - e > 1023 ? 1023 : 8 * s - 2 - e);
    const cplx* p = src + (size_t)b * stride + j;
    for (int kk = tk; kk < TCS_K; kk += 8) {
        const int k = k0 + kk;
        long long xr = 0, xi = 0;
        if (j < R && k < Kc) { const cplx v = p[(size_t)k * ld]; xr = __double2ll_rn(v.x * scale); xi = __double2ll_rn(v.y * scale); if (conj) xi = -xi; }
        long long x3[3] = {xr, xi, xr + xi};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            long long x = x3[c];
            for (int d = s - 1; d >= 0; --d) {
                const int dg = (int)(signed char)(x & 0xFF);
                sb[(((size_t)c * s + d) * TCS_J + tj) * (TCS_LDW * 4) + kk] = (signed char)dg;
                x = (x - dg) >> 8;
            }
        }
    }
    __syncthreads();
    const size_t plane = (size_t)R * Kp;
    signed char* o = out + ((size_t)b * 3 * s) * plane;
    const int nrows = 3 * s * TCS_J;
    for (int idx = threadIdx.x; idx < nrows * 16; idx += blockDim.x) {
        const int rowi = idx >> 4, w = idx & 15;
        const int pl = rowi / TCS_J, jj = rowi % TCS_J;
        if (j0 + jj < R && k0 + 4 * w < Kp)
            *reinterpret_cast<unsigned*>(o + (size_t)pl * plane + (size_t)(j0 + jj) * Kp + k0 + 4 * w) = tcs_smem[(size_t)rowi * TCS_LDW + w];
    }
}
__global__ void tc_fix_exponent_kernel(int* ex, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && ex[i] == TC_E_ZERO) ex[i] = 0;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_encoder() {
    static EncodeTiledFn fn = nullptr;            // idempotent lookup (same value from every thread)
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// planes [nplanes][rows][Kp] int8 -> 3-D map (Kp, rows, nplanes), box 128 B x 128 rows x 1, 128-byte swizzle
bool tc_make_map(CUtensorMap* m, const void* base, int Kp, int rows, long long nplanes) {
    EncodeTiledFn enc = tc_encoder();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)nplanes};
    cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * (cuuint64_t)rows};
    cuuint32_t box[3] = {TC_BK, TC_BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }
inline int kpad(int K) { return (K + TC_BK - 1) / TC_BK * TC_BK; }
// bytes of split storage for one matrix of the batch
size_t tc_bytes_per_matrix(int M, int N, int K, int s) {
    const size_t Kp = kpad(K);
    return al256((size_t)3 * s * M * Kp) + al256((size_t)3 * s * N * Kp) + al256(sizeof(int) * (size_t)M) + al256(sizeof(int) * (size_t)N);
}

// split op(X) so that its rows (want_rows = the index that stays a row of the int8 planes) are K-contiguous
cudaError_t tc_split(const cplx* X, int ld, long long stride, bool rows_contiguous, int R, int Kc, int Kp, int s, int conj,
                     signed char* out, int* ex, int nmat, cudaStream_t st) {
    if (rows_contiguous) {
        tc_split_rows_kernel<<<dim3(R, nmat), 256, 0, st>>>(X, ld, stride, R, Kc, Kp, s, conj, out, ex);
    } else {
        tc_colmax_kernel<<<dim3((R + 127) / 128, nmat), 128, 0, st>>>(X, ld, stride, R, Kc, ex);
        const size_t smem = (size_t)3 * s * TCS_J * TCS_LDW * 4;
        if (smem > 48 * 1024) cudaFuncSetAttribute(tc_split_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_split_cols_kernel<<<dim3((R + TCS_J - 1) / TCS_J, Kp / TCS_K, nmat), 256, smem, st>>>(X, ld, stride, R, Kc, Kp, s, conj, out, ex);
        const size_t n = (size_t)R * nmat;
        tc_fix_exponent_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ex, n);
    }
    return cudaGetLastError();
}

}  // namespace

namespace rcwa {

size_t tc_workspace_bytes(int M, int N, int K, int nb, int s) {
    // enough for the whole batch; the routine also works (in chunks) with less, down to one matrix
    return tc_bytes_per_matrix(M, N, K, s) * (size_t)(nb > 0 ? nb : 1);
}
size_t tc_workspace_min_bytes(int M, int N, int K, int s) { return tc_bytes_per_matrix(M, N, K, s); }

bool tc_supported(int s, int M, int N, int K) { return s >= 2 && s <= TC_MAXS && M > 0 && N > 0 && K > 0 && tc_encoder() != nullptr; }

cudaError_t tc_zgemm_strided(int s, int opa, int opb, int M, int N, int K, double alpha, const cplx* A, int lda, long long sa,
                             const cplx* B, int ldb, long long sb, cplx beta, cplx* Cm, int ldc, long long sc, int batch,
                             char* ws, size_t ws_bytes, cudaStream_t st) {
    if (s < 2 || s > TC_MAXS) return cudaErrorInvalidValue;
    const size_t per = tc_bytes_per_matrix(M, N, K, s);
    int chunk = (int)(ws_bytes / per);
    if (chunk < 1) return cudaErrorInvalidValue;
    if (chunk > batch) chunk = batch;
    const int Kp = kpad(K);
    int dev = 0, nsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);   // per device; cheap, so unconditional
    TcSchedule sch;
    tc_build_schedule(s, 4, &sch);
    char* p = ws;
    signed char* Asl = (signed char*)p; p += al256((size_t)3 * s * M * Kp) * chunk;
    signed char* Bsl = (signed char*)p; p += al256((size_t)3 * s * N * Kp) * chunk;
    int* eA = (int*)p; p += al256(sizeof(int) * (size_t)M) * chunk;
    int* eB = (int*)p;
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        const int nm = (batch - b0 < chunk) ? batch - b0 : chunk;
        cudaError_t e;
        // op(A): rows i, K along.  opa = N: stored [M][K] -> rows contiguous; T/H: stored [K][M] -> columns
        e = tc_split(A + (size_t)b0 * sa, lda, sa, opa == OP_N, M, K, Kp, s, opa == OP_H, Asl, eA, nm, st);
        if (e != cudaSuccess) return e;
        // op(B): "rows" j (columns of op(B)), K along.  opb = N: stored [K][N] -> columns; T/H: stored [N][K] -> rows contiguous
        e = tc_split(B + (size_t)b0 * sb, ldb, sb, opb != OP_N, N, K, Kp, s, opb == OP_H, Bsl, eB, nm, st);
        if (e != cudaSuccess) return e;
        CUtensorMap mA, mB;
        if (!tc_make_map(&mA, Asl, Kp, M, (long long)nm * 3 * s) || !tc_make_map(&mB, Bsl, Kp, N, (long long)nm * 3 * s)) return cudaErrorNotSupported;
        TcParams prm;
        prm.M = M; prm.N = N; prm.K = K;
        prm.nkc = Kp / TC_BK;
        const int klast = K - (prm.nkc - 1) * TC_BK;
        prm.nk_last = (klast + 31) / 32;
        prm.mt = (M + TC_BM - 1) / TC_BM; prm.nt = (N + TC_BN - 1) / TC_BN;
        prm.nmat = nm;
        prm.alpha = alpha; prm.beta = beta;
        prm.C = Cm + (size_t)b0 * sc; prm.ldc = ldc; prm.stride_c = sc;
        prm.eA = eA; prm.eB = eB;
        const int items = nm * prm.mt * prm.nt;
        const int grid = items < nsm ? items : nsm;
        tc_gemm_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(mA, mB, sch, prm);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// debugging / tests: the split on its own
cudaError_t tc_split_debug(const cplx* X, int ld, long long stride, int rows_contiguous, int R, int Kc, int s, int conj,
                           signed char* out, int* ex, int nmat, cudaStream_t st) {
    return tc_split(X, ld, stride, rows_contiguous != 0, R, Kc, kpad(Kc), s, conj, out, ex, nmat, st);
}

}  // namespace rcwa
#endif  // RCWA_TC_HOST_ONLY
