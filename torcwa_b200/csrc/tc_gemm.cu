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
        g.d0 = d0; g.nl = d1 - d0 + 1; g.op0 = sch->nops; g.nloads = 0; g.nmma = 0;
        int loaded_b[TC_MAXS]; bool first[4] = {true, true, true, true};
        for (int q = 0; q < TC_MAXS; ++q) loaded_b[q] = -1;
        const int pmax = d1;                                   // d1 <= s - 1
        // every B plane of the group first, most significant digit last (q = d1 .. 0), into CONSECUTIVE ring slots: planes
        // q and q-1 then sit in adjacent slots, and one N = 256 MMA can multiply an A plane with both (see tc_issue_entry)
        for (int q = d1; q >= 0; --q) { sch->ops[sch->nops++] = TC_OP_LOAD_B | (q << 2); loaded_b[q] = g.nloads++; }
        for (int p = 0; p <= pmax; ++p) {
            const int qlo = (d0 - p > 0) ? d0 - p : 0, qhi = (d1 - p < s - 1) ? d1 - p : s - 1;
            if (qlo > qhi) continue;
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
                ++g.nmma;
            }
        }
        g.nops = sch->nops - g.op0;
    }
}

void tc_compact_schedule(const TcSchedule* sch, TcTables* tab) {
    tab->s = sch->s; tab->ngroups = sch->ngroups;
    for (int gi = 0; gi < sch->ngroups; ++gi) {
        const TcGroup& g = sch->g[gi];
        int nl = 0, nm = 0, seen = 0, acquired = 0;
        for (int o = g.op0; o < g.op0 + g.nops; ++o) {
            const unsigned op = sch->ops[o], type = op & 3u;
            if (type != TC_OP_MMA) {
                tab->loads[gi * TC_MAXLOADS + nl++] = (type == TC_OP_LOAD_B ? 1u : 0u) | (((op >> 2) & 15u) << 1);
                ++seen;
            } else {
                const unsigned ia = (op >> 2) & 31u, ib = (op >> 7) & 31u;
                const int need = (int)(ia > ib ? ia : ib) + 1;
                const int nacq = need > acquired ? need - acquired : 0;
                acquired += nacq;
                tab->mmas[gi * TC_MAXMMAS + nm++] = ia | (ib << 5) | (((op >> 12) & 3u) << 10) | (((op >> 14) & 1u) << 12) |
                                                    (((op >> 15) & 1u) << 13) | (((op >> 16) & 1u) << 14) | ((unsigned)nacq << 15);
            }
        }
        (void)seen;
        tab->d0[gi] = g.d0; tab->nl[gi] = g.nl; tab->nloads[gi] = nl; tab->nmma[gi] = nm;
        // steps: consecutive MMA ops that share their A load
        int ns = 0;
        for (int i = 0; i < nm;) {
            const unsigned w = tab->mmas[gi * TC_MAXMMAS + i], ia = w & 31u;
            unsigned w0 = ia, w1 = 0, nacq = 0, relb = 0;
            int np = 0;
            while (i < nm && (tab->mmas[gi * TC_MAXMMAS + i] & 31u) == ia && np < 4) {
                const unsigned m = tab->mmas[gi * TC_MAXMMAS + i];
                w1 |= ((((m >> 5) & 31u)) | (((m >> 10) & 3u) << 5) | (((m >> 12) & 1u) << 7)) << (8 * np);
                nacq += (m >> 15) & 15u;
                if ((m >> 14) & 1u) relb = 1u + ((m >> 5) & 31u);     // at most one B plane has its last use in a step
                ++np; ++i;
            }
            w0 |= (unsigned)np << 5 | nacq << 8 | relb << 12;
            tab->steps[(gi * TC_MAXS + ns) * 2] = w0;
            tab->steps[(gi * TC_MAXS + ns) * 2 + 1] = w1;
            ++ns;
        }
        tab->nsteps[gi] = ns;
    }
}

}  // namespace rcwa

#ifndef RCWA_TC_HOST_ONLY
namespace {
using namespace rcwa;

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 128;          // tile rows / columns, K bytes per ring slot
constexpr int TC_SLOT = TC_BM * TC_BK;                          // 16 KB
constexpr int TC_THREADS = 384;                                  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr int TC_TAB_GROUPS = 2;                                 // level groups the kernel supports (s <= 8 at 4 levels per group)
constexpr int TC_TAB_BYTES = TC_TAB_GROUPS * TC_RING * TC_MAXS * 64;
constexpr int TC_SMEM = TC_RING * TC_SLOT + 1024 /*alignment slack*/ + 512 /*barriers*/ + TC_TAB_BYTES;

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
// one lane of a converged warp (the predicate ptxas understands as "single thread": no ELECT loops around uniform-datapath ops)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n .reg .b32 rx;\n .reg .pred px;\n elect.sync rx|px, 0xFFFFFFFF;\n @px mov.s32 %0, 1;\n}" : "+r"(pred));
    return pred != 0;
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

// the MMAs of one (A plane, B plane) pair over one K chunk (nk x 32 bytes)
__device__ __forceinline__ void tc_issue_pair(uint32_t dcol, uint64_t da, uint64_t db, uint32_t idesc, uint32_t keep, int nk) {
    tc_mma_i8(dcol, da, db, idesc, keep);
    if (nk > 1) tc_mma_i8(dcol, da + 2, db + 2, idesc, 1u);
    if (nk > 2) tc_mma_i8(dcol, da + 4, db + 4, idesc, 1u);
    if (nk > 3) tc_mma_i8(dcol, da + 6, db + 6, idesc, 1u);
}
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
               const __grid_constant__ TcTables sch, const __grid_constant__ TcParams prm) {
    extern __shared__ char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;           // 1024-byte aligned ring (swizzle atom)
    const uint32_t bars = sbase + TC_RING * TC_SLOT;                          // full[RING], empty[RING], tfull, tempty (8 B each), tmem ptr
    const uint32_t bar_full = bars, bar_empty = bars + 8 * TC_RING, bar_tfull = bars + 16 * TC_RING, bar_tempty = bar_tfull + 8;
    const uint32_t tmem_slot = bar_tempty + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = sch.s;
    // Issue table: for every (level group, ring position of the iteration's first load, step) the ready-made operands of
    // the issue (tc_issue_entry, kernels.h, with ring slots turned into descriptor low words / barrier offsets here):
    // the issuer's per-step work is three 16-byte shared loads, the waits and the tcgen05 instructions themselves.
    const uint32_t tab = bars + 512;
    {
        const uint32_t lo0 = (sbase & 0x3FFFFu) >> 4;
        uint32_t* T = reinterpret_cast<uint32_t*>(smem_raw + (tab - smem_u32(smem_raw)));
        for (int idx = threadIdx.x; idx < sch.ngroups * TC_RING * TC_MAXS; idx += blockDim.x) {
            const int g = idx / (TC_RING * TC_MAXS), bs = (idx / TC_MAXS) % TC_RING, st = idx % TC_MAXS;
            unsigned e[16];
            tc_issue_entry(sch, g, bs, st, e);
            uint32_t* o = T + (size_t)idx * 16;
            o[0] = lo0 + e[0] * (TC_SLOT >> 4);
            o[1] = (e[1] & 0xFFu) | (((e[1] >> 8) & 255u) * 8u) << 8 | (e[1] & (1u << 16)) | (((e[1] >> 17) & 255u) * 8u) << 17;
            for (int j = 0; j < 4; ++j) { o[2 + 2 * j] = lo0 + e[2 + 2 * j] * (TC_SLOT >> 4); o[3 + 2 * j] = e[3 + 2 * j]; }
        }
    }
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

    // Both data-movement roles run as WARP-UNIFORM loops (all 32 lanes walk the table; the asynchronous instruction itself
    // is issued by one elected lane): the compiler keeps the bookkeeping in uniform registers, and the per-op work is a
    // few dozen instructions.  (A loop inside `if (lane == 0)` made ptxas wrap every UTCIMMA in an ELECT loop -- 130
    // instructions per four MMAs, measured 880 cycles per op against 256 cycles of tensor work.)
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory");
    if (warp == 0) {
        // ===================================================== TMA producer
        unsigned slot = 0, par = 1;                      // parity to wait for on empty[slot]: starts at 1 (fresh barrier passes)
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int b = item / (prm.mt * prm.nt), rem = item % (prm.mt * prm.nt);
            const int m0 = (rem / prm.nt) * TC_BM, n0 = (rem % prm.nt) * TC_BN;
            for (int t = 0; t < 3; ++t) {
                const int plane0 = (b * 3 + t) * s;
                for (int g = 0; g < sch.ngroups; ++g) {
                    const int l0 = g * TC_MAXLOADS, nloads = sch.nloads[g];
                    for (int kc = 0; kc < prm.nkc; ++kc) {
                        for (int i = 0; i < nloads; ++i) {
                            const uint32_t w = sch.loads[l0 + i];                 // bit 0: B operand, bits 1-4: digit
                            mbar_wait(bar_empty + 8 * slot, par);
                            if (elect_one()) {
                                mbar_expect_tx(bar_full + 8 * slot, TC_SLOT);
                                tma_load_3d(sbase + slot * TC_SLOT, (w & 1u) ? &mapB : &mapA, bar_full + 8 * slot, kc * TC_BK,
                                            (w & 1u) ? n0 : m0, plane0 + (int)(w >> 1));
                            }
                            __syncwarp();
                            if (++slot == TC_RING) { slot = 0; par ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        const uint32_t desc_hi = (uint32_t)(tc_smem_desc(0) >> 32);
        unsigned aslot = 0, apar = 0;                    // acquire cursor: next load to wait for on full[aslot]
        unsigned bslot = 0;                              // ring slot of load 0 of the current (group, K chunk) iteration
        unsigned grp = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int rem = item % (prm.mt * prm.nt);
            const int n0 = (rem % prm.nt) * TC_BN;
            int ncols = prm.N - n0; if (ncols > TC_BN) ncols = TC_BN;
            const uint32_t idesc = tc_idesc((ncols + 15) & ~15);
            const uint32_t idesc2 = tc_idesc(TC_BN + ((ncols + 15) & ~15));   // two planes at once: 128 rows of the first + the live rows of the second
            for (int t = 0; t < 3; ++t) {
                for (int g = 0; g < sch.ngroups; ++g, ++grp) {
                    mbar_wait(bar_tempty, (grp & 1u) ^ 1u);          // the epilogue has drained the previous group's accumulators
                    tc_fence_after();
                    const int ns = sch.nsteps[g], nloads = sch.nloads[g];
                    for (int kc = 0; kc < prm.nkc; ++kc) {
                        const int nk = (kc == prm.nkc - 1) ? prm.nk_last : TC_BK / 32;
                        const uint32_t tp = tab + (uint32_t)((g * TC_RING + bslot) * TC_MAXS) * 64u;
                        const uint32_t fresh_mask = (kc == 0) ? (1u << 9) : 0u;
                        for (int i = 0; i < ns; ++i) {
                            uint4 e0, e1, e2;
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e0.x), "=r"(e0.y), "=r"(e0.z), "=r"(e0.w) : "r"(tp + 64u * i));
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e1.x), "=r"(e1.y), "=r"(e1.z), "=r"(e1.w) : "r"(tp + 64u * i + 16u));
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e2.x), "=r"(e2.y), "=r"(e2.z), "=r"(e2.w) : "r"(tp + 64u * i + 32u));
                            for (unsigned n = (e0.y >> 4) & 15u; n > 0; --n) {
                                mbar_wait(bar_full + 8 * aslot, apar);
                                if (++aslot == TC_RING) { aslot = 0; apar ^= 1u; }
                            }
                            tc_fence_after();
                            if (elect_one()) {
                                const uint32_t ng = e0.y & 15u;
                                const uint64_t da = ((uint64_t)desc_hi << 32) | e0.x;
                                tc_issue_pair(tmem_base + (e0.w & 511u), da, ((uint64_t)desc_hi << 32) | e0.z, (e0.w & 1024u) ? idesc2 : idesc, (e0.w & fresh_mask) ? 0u : 1u, nk);
                                if (ng > 1) tc_issue_pair(tmem_base + (e1.y & 511u), da, ((uint64_t)desc_hi << 32) | e1.x, (e1.y & 1024u) ? idesc2 : idesc, (e1.y & fresh_mask) ? 0u : 1u, nk);
                                if (ng > 2) tc_issue_pair(tmem_base + (e1.w & 511u), da, ((uint64_t)desc_hi << 32) | e1.z, (e1.w & 1024u) ? idesc2 : idesc, (e1.w & fresh_mask) ? 0u : 1u, nk);
                                if (ng > 3) tc_issue_pair(tmem_base + (e2.y & 511u), da, ((uint64_t)desc_hi << 32) | e2.x, (e2.y & 1024u) ? idesc2 : idesc, (e2.y & fresh_mask) ? 0u : 1u, nk);
                                tc_commit(bar_empty + ((e0.y >> 8) & 255u));                           // A plane: its last use is this step
                                if (e0.y & (1u << 16)) tc_commit(bar_empty + ((e0.y >> 17) & 255u));    // the B plane that is done
                            }
                            __syncwarp();
                        }
                        bslot += nloads;
                        if (bslot >= TC_RING) bslot -= TC_RING;
                        if (bslot >= TC_RING) bslot -= TC_RING;
                    }
                    if (elect_one()) tc_commit(bar_tfull);
                    __syncwarp();
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
                    const int nl = sch.nl[g];
                    const double w = pow2d(-8 * (sch.d0[g] + nl - 1));
                    mbar_wait(bar_tfull, grp & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        int r0[8], r1[8], r2[8], r3[8];
                        // level l of the group lives at column (nl-1-l)*128 (adjacent levels side by side for the N = 256 MMAs)
                        tmem_ld8(tlane + (nl - 1) * TC_BN + cc * 8, r0);
                        if (nl > 1) tmem_ld8(tlane + (nl - 2) * TC_BN + cc * 8, r1);
                        if (nl > 2) tmem_ld8(tlane + (nl - 3) * TC_BN + cc * 8, r2);
                        if (nl > 3) tmem_ld8(tlane + (nl - 4) * TC_BN + cc * 8, r3);
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

// column maxima of src [Kc rows][R cols]: thread per column (coalesced across the warp), the K range split over
// blockIdx.y; the per-range exponents meet in an atomicMax on ex (pre-filled with TC_E_ZERO)
constexpr int TCM_K = 64;
__global__ void tc_fill_int_kernel(int* p, size_t n, int v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(128)
tc_colmax_kernel(const cplx* __restrict__ src, int ld, long long stride, int R, int Kc, int* __restrict__ ex) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.z;
    if (j >= R) return;
    const int k0 = blockIdx.y * TCM_K, k1 = (k0 + TCM_K < Kc) ? k0 + TCM_K : Kc;
    const cplx* p = src + (size_t)b * stride + j;
    double mx = 0.0;
#pragma unroll 8
    for (int k = k0; k < k1; ++k) { const cplx v = p[(size_t)k * ld]; mx = fmax(mx, fabs(v.x) + fabs(v.y)); }
    const int e = tc_exponent(mx);
    if (e != TC_E_ZERO) atomicMax(&ex[(size_t)b * R + j], e);      // TC_E_ZERO marks an all-zero column until the split kernel has consumed it
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
        const size_t nex = (size_t)R * nmat;
        tc_fill_int_kernel<<<(unsigned)((nex + 255) / 256), 256, 0, st>>>(ex, nex, TC_E_ZERO);
        tc_colmax_kernel<<<dim3((R + 127) / 128, (Kc + TCM_K - 1) / TCM_K, nmat), 128, 0, st>>>(X, ld, stride, R, Kc, ex);
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

bool tc_supported(int s, int M, int N, int K) { return s >= 2 && s <= TC_MAXS && (s + 3) / 4 <= TC_TAB_GROUPS && M > 0 && N > 0 && K > 0 && tc_encoder() != nullptr; }

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
    TcSchedule full;
    tc_build_schedule(s, 4, &full);
    TcTables sch;
    tc_compact_schedule(&full, &sch);
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

size_t tc_ctx_bytes(int n, int nb, int slices) {
    if (slices < 2) return 0;
    const int chunk = nb < 16 ? (nb > 0 ? nb : 1) : 16;          // digits of 16 matrices at a time keep the kernel's grid full
    return tc_bytes_per_matrix(n, n, n, slices) * (size_t)chunk;
}

cudaError_t gemm_auto(const TcCtx* tc, int opa, int opb, int M, int N, int K, double alpha, const cplx* A, int lda, long long sa,
                      const cplx* B, int ldb, long long sb, cplx beta, cplx* Cm, int ldc, long long sc, int batch,
                      ZGemmProblem* gscratch, cudaStream_t st) {
    // the digit path pays a split of both operands (~O(MK + KN) bytes each way) and an epilogue of three read-modify-write passes
    // over C: only for products with enough work per element.  Measured at 7 digits, split included: 71 TF/s-equivalent at
    // M = N = K = 1922 (DMMA 38), but 24 at 481 (DMMA ~30: the symmetry blocks of order 15), and K = 512 block updates were
    // slower than DMMA too (profiles/r2_summary.md sections 3 and 6) -- hence the K threshold.
    const bool big = M >= 256 && N >= 128 && K >= 768;
    if (tc && tc->slices >= 2 && big && tc_supported(tc->slices, M, N, K) && tc->ws_bytes >= tc_workspace_min_bytes(M, N, K, tc->slices))
        return tc_zgemm_strided(tc->slices, opa, opb, M, N, K, alpha, A, lda, sa, B, ldb, sb, beta, Cm, ldc, sc, batch, tc->ws, tc->ws_bytes, st);
    return zgemm_strided(opa, opb, M, N, K, C(alpha, 0.0), A, lda, sa, B, ldb, sb, beta, Cm, ldc, sc, batch, gscratch, st);
}

// debugging / tests: the split on its own
cudaError_t tc_split_debug(const cplx* X, int ld, long long stride, int rows_contiguous, int R, int Kc, int s, int conj,
                           signed char* out, int* ex, int nmat, cudaStream_t st) {
    return tc_split(X, ld, stride, rows_contiguous != 0, R, Kc, kpad(Kc), s, conj, out, ex, nmat, st);
}

}  // namespace rcwa
#endif  // RCWA_TC_HOST_ONLY
