// Batched complex128 LU for ROW-MAJOR RIGHT-SOLVES:   X * A = B   (X, B are nrows x n).
//
// Every linear solve of the RCWA path can be put in this form without a single transpose
// (inv(eps_conv); the two solves of the minimal layer S-matrix; the one solve of
// the Redheffer product, SURVEY.md A.6 via the push-through identity).  In row-major storage the
// natural elimination is then by ROWS with COLUMN pivoting,
//
//      A * Pi = L * U,    L lower (non-unit),   U unit upper,
//
// i.e. LAPACK getrf applied to A^T, which makes pivot search, scaling and the rank-1 update all
// contiguous along rows.  Blocked right-looking, panel = NB rows:
//   panel kernel (one CTA per matrix; phase-structured, also compiled by the CPU emulation build)
//   -> column swaps outside the panel -> rows below: A21 := A21 * U11^-1 (inverted 32 x 32 block, K = 32 GEMM)
//   -> trailing update A22 -= A21 * A12 on the DMMA grouped GEMM.
// Solve: X = B*Pi (gather) ; X := X * U^-1 (forward over column blocks) ; X := X * L^-1 (backward),
// each block step = a per-row small triangular solve + one GEMM.
//
// Replaces torch.linalg.inv at /root/reference/torcwa/rcwa.py:1226,1230,1248,1266-1267,
// 1271,1273,1287-1288.
#include "common.cuh"
#ifndef RCWA_EMU
#include "kernels.h"
#else
#include <vector>
#endif

#define LU_NB 32
#define SOLVE_NB 128        // block size of the triangular solves on the fp64 (DMMA) path (diagonal blocks are inverted once per factorisation)
#define SOLVE_NB_TC 512     // ... on the tcgen05 path: right-looking updates with K = 512 (tc_gemm.cu)

// ------------------------------------------------------------------------------------------------
// panel factorisation: rows [k0, k0+nbe) of A (n x n, leading dim lda), columns [k0, n)
// smem: 40 doubles + 40 ints scratch
DEV void lu_panel_body(const Cta& c, cplx* A, int n, int lda, int k0, int nbe, int* ipiv, int* info) {
    double* red_v = reinterpret_cast<double*>(c.smem);          // [33]
    int* red_i = reinterpret_cast<int*>(red_v + 40);             // [33]
    for (int j = k0; j < k0 + nbe; ++j) {
        // ---- pivot search along row j
        cplx* rowj = A + (size_t)j * lda;
        double best = -1.0; int bidx = j;
        for (int col = j + c.tid; col < n; col += c.nthreads) {
            double v = cabs1(rowj[col]);
            if (v > best) { best = v; bidx = col; }
        }
#ifndef RCWA_EMU
        // warp argmax (ties -> smaller index), then across warps
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
        {
            int w = c.tid >> 5, l = c.tid & 31, nw = (c.nthreads + 31) >> 5;
            CTA_SYNC();
            if (l == 0) { red_v[w] = best; red_i[w] = bidx; }
            CTA_SYNC();
            if (w == 0) {
                double v = (l < nw) ? red_v[l] : -2.0; int i2 = (l < nw) ? red_i[l] : 0x7fffffff;
                for (int o = 16; o > 0; o >>= 1) {
                    double ov = __shfl_xor_sync(0xffffffffu, v, o);
                    int oi = __shfl_xor_sync(0xffffffffu, i2, o);
                    if (ov > v || (ov == v && oi < i2)) { v = ov; i2 = oi; }
                }
                if (l == 0) { red_v[32] = v; red_i[32] = i2; }
            }
            CTA_SYNC();
            best = red_v[32]; bidx = red_i[32];
        }
#else
        (void)red_v; (void)red_i;
#endif
        const int p = bidx;
        if (c.tid == 0) {
            ipiv[j] = p;
            if (best == 0.0 && *info == 0) *info = j + 1;
        }
        // ---- swap columns j <-> p inside the panel rows
        if (p != j) {
            for (int r = k0 + c.tid; r < k0 + nbe; r += c.nthreads) {
                cplx* row = A + (size_t)r * lda;
                cplx t = row[j]; row[j] = row[p]; row[p] = t;
            }
        }
        CTA_SYNC();
        // ---- scale row j right of the diagonal by 1/pivot
        const cplx piv = rowj[j];
        if (best != 0.0) {
            const cplx ip = cinv(piv);
            for (int col = j + 1 + c.tid; col < n; col += c.nthreads) rowj[col] = cmul(rowj[col], ip);
        }
        CTA_SYNC();
        // ---- rank-1 update of the remaining panel rows
        const int nr = k0 + nbe - (j + 1), ncol = n - (j + 1);
        if (nr > 0 && ncol > 0) {
            for (long long idx = c.tid; idx < (long long)nr * ncol; idx += c.nthreads) {
                int r = j + 1 + (int)(idx / ncol), col = j + 1 + (int)(idx % ncol);
                cplx* row = A + (size_t)r * lda;
                row[col] = csub(row[col], cmul(row[j], rowj[col]));
            }
        }
        CTA_SYNC();
    }
}

// perm[c] = original column sitting at position c after all swaps (so (B*Pi)[r][c] = B[r][perm[c]])
DEV void lu_perm_body(const Cta& c, const int* ipiv, int n, int* perm) {
    int* idx = reinterpret_cast<int*>(c.smem);   // [n]
    for (int i = c.tid; i < n; i += c.nthreads) idx[i] = i;
    CTA_SYNC();
    if (c.tid == 0) {
        for (int j = 0; j < n; ++j) { int p = ipiv[j]; if (p != j) { int t = idx[j]; idx[j] = idx[p]; idx[p] = t; } }
    }
    CTA_SYNC();
    for (int i = c.tid; i < n; i += c.nthreads) perm[i] = idx[i];
}

// x * T = a for one row: mode 0: T unit upper (forward over columns); mode 1: T lower non-unit (backward).
// T is nbe x nbe with leading dimension ldt; x overwrites a (nbe contiguous entries).
DEV void trsm_row(cplx* x, const cplx* T, int ldt, int nbe, int mode) {
    if (mode == 0) {
        for (int col = 1; col < nbe; ++col) {
            cplx acc = x[col];
            for (int s = 0; s < col; ++s) acc = csub(acc, cmul(x[s], T[s * ldt + col]));
            x[col] = acc;
        }
    } else {
        for (int col = nbe - 1; col >= 0; --col) {
            cplx acc = x[col];
            for (int s = col + 1; s < nbe; ++s) acc = csub(acc, cmul(x[s], T[s * ldt + col]));
            x[col] = cdiv(acc, T[col * ldt + col]);
        }
    }
}

#ifdef RCWA_EMU
// ------------------------------------------------------------------ CPU emulation entry points (tests only)
extern "C" int emu_lu_factor(cplx* A, int n, int lda, int* ipiv, int* perm, int* info) {
    char smem[65536];
    Cta c; c.tid = 0; c.nthreads = 1; c.bid = 0; c.smem = smem; c.warp_only = 0;
    *info = 0;
    for (int k0 = 0; k0 < n; k0 += LU_NB) {
        int nbe = (n - k0 < LU_NB) ? n - k0 : LU_NB;
        lu_panel_body(c, A, n, lda, k0, nbe, ipiv, info);
        // column swaps outside the panel
        for (int r = 0; r < n; ++r) {
            if (r >= k0 && r < k0 + nbe) continue;
            for (int j = k0; j < k0 + nbe; ++j) { int p = ipiv[j]; if (p != j) { cplx t = A[(size_t)r * lda + j]; A[(size_t)r * lda + j] = A[(size_t)r * lda + p]; A[(size_t)r * lda + p] = t; } }
        }
        // rows below: x * U11 = a ; then A22 -= A21 * A12
        for (int r = k0 + nbe; r < n; ++r) trsm_row(A + (size_t)r * lda + k0, A + (size_t)k0 * lda + k0, lda, nbe, 0);
        for (int r = k0 + nbe; r < n; ++r)
            for (int col = k0 + nbe; col < n; ++col) {
                cplx acc = A[(size_t)r * lda + col];
                for (int s = 0; s < nbe; ++s) acc = csub(acc, cmul(A[(size_t)r * lda + k0 + s], A[(size_t)(k0 + s) * lda + col]));
                A[(size_t)r * lda + col] = acc;
            }
    }
    std::vector<char> big((size_t)n * sizeof(int) + 64);
    c.smem = big.data();
    lu_perm_body(c, ipiv, n, perm);
    return 0;
}
extern "C" int emu_lu_solve(const cplx* LU, int n, int lda, const int* perm, const cplx* B, int nrows, int ldb, cplx* X, int ldx) {
    for (int r = 0; r < nrows; ++r) for (int col = 0; col < n; ++col) X[(size_t)r * ldx + col] = B[(size_t)r * ldb + perm[col]];
    for (int k0 = 0; k0 < n; k0 += LU_NB) {
        int nbe = (n - k0 < LU_NB) ? n - k0 : LU_NB;
        for (int r = 0; r < nrows; ++r) trsm_row(X + (size_t)r * ldx + k0, LU + (size_t)k0 * lda + k0, lda, nbe, 0);
        for (int r = 0; r < nrows; ++r)
            for (int col = k0 + nbe; col < n; ++col) {
                cplx acc = X[(size_t)r * ldx + col];
                for (int s = 0; s < nbe; ++s) acc = csub(acc, cmul(X[(size_t)r * ldx + k0 + s], LU[(size_t)(k0 + s) * lda + col]));
                X[(size_t)r * ldx + col] = acc;
            }
    }
    int nblk = (n + LU_NB - 1) / LU_NB;
    for (int kb = nblk - 1; kb >= 0; --kb) {
        int k0 = kb * LU_NB, nbe = (n - k0 < LU_NB) ? n - k0 : LU_NB;
        for (int r = 0; r < nrows; ++r) trsm_row(X + (size_t)r * ldx + k0, LU + (size_t)k0 * lda + k0, lda, nbe, 1);
        for (int r = 0; r < nrows; ++r)
            for (int col = 0; col < k0; ++col) {
                cplx acc = X[(size_t)r * ldx + col];
                for (int s = 0; s < nbe; ++s) acc = csub(acc, cmul(X[(size_t)r * ldx + k0 + s], LU[(size_t)(k0 + s) * lda + col]));
                X[(size_t)r * ldx + col] = acc;
            }
    }
    return 0;
}
#else
// ------------------------------------------------------------------ device kernels
namespace {

__global__ void __launch_bounds__(512, 1)
lu_panel_kernel(cplx* A, long long stride, int n, int lda, int k0, int nbe, int* ipiv, int* info) {
    extern __shared__ __align__(16) char smem_raw[];
    Cta c = make_cta(blockIdx.x, smem_raw);
    lu_panel_body(c, A + (size_t)blockIdx.x * stride, n, lda, k0, nbe, ipiv + (size_t)blockIdx.x * n, info + blockIdx.x);
}

__global__ void lu_perm_kernel(const int* ipiv, int n, int* perm) {
    extern __shared__ __align__(16) char smem_raw[];
    Cta c = make_cta(blockIdx.x, smem_raw);
    lu_perm_body(c, ipiv + (size_t)blockIdx.x * n, n, perm + (size_t)blockIdx.x * n);
}

// grid (ceil(n/128), B): thread per row outside the panel applies the panel's column swaps in order
__global__ void lu_colswap_kernel(cplx* A, long long stride, int n, int lda, int k0, int nbe, const int* ipiv) {
    __shared__ int piv[LU_NB];
    const int b = blockIdx.y;
    if (threadIdx.x < nbe) piv[threadIdx.x] = ipiv[(size_t)b * n + k0 + threadIdx.x];
    __syncthreads();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || (r >= k0 && r < k0 + nbe)) return;
    cplx* row = A + (size_t)b * stride + (size_t)r * lda;
    for (int jj = 0; jj < nbe; ++jj) {
        int p = piv[jj], j = k0 + jj;
        if (p != j) { cplx t = row[j]; row[j] = row[p]; row[p] = t; }
    }
}

// Inverses of the SOLVE_NB x SOLVE_NB diagonal blocks of the factors, so that the triangular solves
// become large-K GEMMs:  grid (nblk, B, 2): z = 0 -> unit upper U_kk^-1, z = 1 -> lower L_kk^-1.
// One thread per column of the inverse (back-substitution down its own column; every thread reads the
// same factor entry -> broadcast loads; the inverse column lives in the output itself).
// Out-of-range rows/columns of the last block behave like an identity extension.
__global__ void __launch_bounds__(SOLVE_NB_TC)
tri_inv_kernel(const cplx* __restrict__ LU, long long lustride, int n, int lda, cplx* __restrict__ tinv, int nblk, int SNB) {
    const int kb = blockIdx.x, b = blockIdx.y, which = blockIdx.z, j = threadIdx.x;
    const int k0 = kb * SNB;
    const int w = (n - k0 < SNB) ? n - k0 : SNB;
    const cplx* F = LU + (size_t)b * lustride + (size_t)k0 * lda + k0;
    cplx* X = tinv + (((size_t)b * 2 + which) * nblk + kb) * SNB * SNB;
    for (int i = 0; i < SNB; ++i) X[i * SNB + j] = C(i == j ? 1.0 : 0.0, 0.0);
    if (j >= w) return;
    if (which == 0) {
        // U unit upper:  x_j = 1 ; x_i = -sum_{k=i+1..j} U[i][k] x_k   (i = j-1 .. 0)
        for (int i = j - 1; i >= 0; --i) {
            cplx acc = C(0, 0);
            for (int k = i + 1; k <= j; ++k) acc = cfma(F[(size_t)i * lda + k], X[k * SNB + j], acc);
            X[i * SNB + j] = cneg(acc);
        }
    } else {
        // L lower (non-unit):  x_j = 1/L_jj ; x_i = -(sum_{k=j..i-1} L[i][k] x_k) / L_ii   (i = j+1 .. w-1)
        X[j * SNB + j] = cinv(F[(size_t)j * lda + j]);
        for (int i = j + 1; i < w; ++i) {
            cplx acc = C(0, 0);
            for (int k = j; k < i; ++k) acc = cfma(F[(size_t)i * lda + k], X[k * SNB + j], acc);
            X[i * SNB + j] = cneg(cdiv(acc, F[(size_t)i * lda + i]));
        }
    }
}

// Inverse of the unit-upper nbe x nbe diagonal block U11 of the current panel (identity-extended to LU_NB), so that
// the rows below get  A21 := A21 * U11^-1  as a K = 32 GEMM instead of a per-row substitution by 32-thread CTAs.
// grid (B), block (LU_NB): thread j builds column j by back-substitution in shared memory.
__global__ void __launch_bounds__(LU_NB)
tri_inv_panel_kernel(const cplx* __restrict__ LU, long long lustride, int lda, int k0, int nbe, cplx* __restrict__ uinv, long long ustride) {
    __shared__ cplx F[LU_NB][LU_NB + 1];
    __shared__ cplx X[LU_NB][LU_NB + 1];
    const int b = blockIdx.x, j = threadIdx.x;
    const cplx* src = LU + (size_t)b * lustride + (size_t)k0 * lda + k0;
    for (int i = 0; i < nbe; ++i) F[i][j] = (j < nbe) ? src[(size_t)i * lda + j] : C(0, 0);
    for (int i = 0; i < LU_NB; ++i) X[i][j] = C(i == j ? 1.0 : 0.0, 0.0);
    __syncthreads();
    if (j < nbe) {
        for (int i = j - 1; i >= 0; --i) {
            cplx acc = C(0, 0);
            for (int k = i + 1; k <= j; ++k) acc = cfma(F[i][k], X[k][j], acc);
            X[i][j] = cneg(acc);
        }
    }
    __syncthreads();
    cplx* dst = uinv + (size_t)b * ustride;
    for (int i = 0; i < LU_NB; ++i) dst[i * LU_NB + j] = X[i][j];
}

// grid (nrows, B): X[r][c] = Bm[r][perm[c]]
__global__ void gather_cols_kernel(const cplx* __restrict__ Bm, long long bstride, int ldb, const int* __restrict__ perm,
                                   int n, cplx* __restrict__ X, long long xstride, int ldx) {
    const int r = blockIdx.x, b = blockIdx.y;
    const cplx* src = Bm + (size_t)b * bstride + (size_t)r * ldb;
    cplx* dst = X + (size_t)b * xstride + (size_t)r * ldx;
    const int* pm = perm + (size_t)b * n;
    for (int col = threadIdx.x; col < n; col += blockDim.x) dst[col] = src[pm[col]];
}

}  // namespace

namespace rcwa {

// A: [B] matrices n x n (lda, stride) overwritten by L\U; ipiv, perm: [B,n] ints; info: [B] ints
int lu_solve_block(int tc_slices) { return tc_slices >= 2 ? SOLVE_NB_TC : SOLVE_NB; }
size_t lu_tinv_elems(int n, int nb, int tc_slices) {
    const int S = lu_solve_block(tc_slices);
    return (size_t)nb * 2 * ((n + S - 1) / S) * S * S;
}

cudaError_t lu_factor(cplx* A, long long stride, int n, int lda, int nb, int* ipiv, int* perm, int* info, cplx* tinv,
                      ZGemmProblem* gscratch, cudaStream_t st, bool clear_info, int tc_slices) {
    const int SNB = lu_solve_block(tc_slices);
    if (clear_info) cudaMemsetAsync(info, 0, sizeof(int) * nb, st);
    const cplx one = C(1, 0), mone = C(-1, 0);
    for (int k0 = 0; k0 < n; k0 += LU_NB) {
        const int nbe = (n - k0 < LU_NB) ? n - k0 : LU_NB;
        lu_panel_kernel<<<nb, 512, 1024, st>>>(A, stride, n, lda, k0, nbe, ipiv, info);
        lu_colswap_kernel<<<dim3((n + 127) / 128, nb), 128, 0, st>>>(A, stride, n, lda, k0, nbe, ipiv);
        const int rem = n - k0 - nbe;
        if (rem > 0) {
            // A21 := A21 * U11^-1 (in place: one tile spans the nbe <= 32 columns).  `tinv` is free until the end of the
            // factorisation and serves as the scratch of the inverted diagonal block.
            const long long ustride = (long long)2 * ((n + SNB - 1) / SNB) * SNB * SNB;
            tri_inv_panel_kernel<<<nb, LU_NB, 0, st>>>(A, stride, lda, k0, nbe, tinv, ustride);
            cudaError_t e = zgemm_strided(OP_N, OP_N, rem, nbe, nbe, one, A + (size_t)(k0 + nbe) * lda + k0, lda, stride,
                                          tinv, LU_NB, ustride, C(0, 0), A + (size_t)(k0 + nbe) * lda + k0, lda, stride, nb, gscratch, st);
            if (e != cudaSuccess) return e;
            e = zgemm_strided(OP_N, OP_N, rem, rem, nbe, mone,
                                          A + (size_t)(k0 + nbe) * lda + k0, lda, stride,
                                          A + (size_t)k0 * lda + (k0 + nbe), lda, stride, one,
                                          A + (size_t)(k0 + nbe) * lda + (k0 + nbe), lda, stride, nb, gscratch, st);
            if (e != cudaSuccess) return e;
        }
    }
    lu_perm_kernel<<<nb, 256, n * sizeof(int), st>>>(ipiv, n, perm);
    const int nblk = (n + SNB - 1) / SNB;
    tri_inv_kernel<<<dim3(nblk, nb, 2), SNB, 0, st>>>(A, stride, n, lda, tinv, nblk, SNB);
    return cudaGetLastError();
}

// X (nrows x n, ldx, xstride) = Bm * A^-1 ; X must not alias Bm.  `Yw`: work buffer like X (same ld / stride).
//   X A = B,  A Pi = L U   =>   Y U = B Pi (forward over column blocks),  X L = Y (backward).
// Left-looking over SOLVE_NB-wide column blocks with pre-inverted diagonal blocks: every step is a
// GEMM with K = (columns already solved) followed by a K = SOLVE_NB GEMM -- large-K DMMA work instead
// of rank-32 updates.
cudaError_t lu_solve_right(const cplx* LU, long long lustride, int n, int lda, const int* perm, const cplx* tinv,
                           const cplx* Bm, long long bstride, int ldb, int nrows, cplx* X, long long xstride, int ldx,
                           cplx* Yw, int nb, ZGemmProblem* gscratch, cudaStream_t st, const TcCtx* tc) {
    const cplx one = C(1, 0), mone = C(-1, 0), zero = C(0, 0);
    if (tc && tc->slices >= 2) {
        // tcgen05 path: RIGHT-looking over SOLVE_NB_TC-wide column blocks, so that every product has K = 512 and each
        // operand block is split into digits once:   Y_k = X_k Uinv_k ;  X_{>k} -= Y_k U_{k,>k}   (forward), then
        // Z_k = Y_k Linv_k ;  Y_{<k} -= Z_k L_{k,<k}   (backward; Z overwrites the X buffer).
        const int S = SOLVE_NB_TC, nblk = (n + S - 1) / S;
        const long long tstride = (long long)2 * nblk * S * S;
        const cplx* Uinv = tinv;
        const cplx* Linv = tinv + (size_t)nblk * S * S;
        gather_cols_kernel<<<dim3(nrows, nb), 256, 0, st>>>(Bm, bstride, ldb, perm, n, X, xstride, ldx);
#define SK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)
        for (int kb = 0; kb < nblk; ++kb) {
            const int c0 = kb * S, w = (n - c0 < S) ? n - c0 : S, c1 = c0 + w;
            SK(gemm_auto(tc, OP_N, OP_N, nrows, w, w, 1.0, X + c0, ldx, xstride, Uinv + (size_t)kb * S * S, S, tstride, zero,
                         Yw + c0, ldx, xstride, nb, gscratch, st));
            if (c1 < n)
                SK(gemm_auto(tc, OP_N, OP_N, nrows, n - c1, w, -1.0, Yw + c0, ldx, xstride, LU + (size_t)c0 * lda + c1, lda, lustride, one,
                             X + c1, ldx, xstride, nb, gscratch, st));
        }
        for (int kb = nblk - 1; kb >= 0; --kb) {
            const int c0 = kb * S, w = (n - c0 < S) ? n - c0 : S;
            SK(gemm_auto(tc, OP_N, OP_N, nrows, w, w, 1.0, Yw + c0, ldx, xstride, Linv + (size_t)kb * S * S, S, tstride, zero,
                         X + c0, ldx, xstride, nb, gscratch, st));
            if (c0 > 0)
                SK(gemm_auto(tc, OP_N, OP_N, nrows, c0, w, -1.0, X + c0, ldx, xstride, LU + (size_t)c0 * lda, lda, lustride, one,
                             Yw, ldx, xstride, nb, gscratch, st));
        }
#undef SK
        return cudaGetLastError();
    }
    const int nblk = (n + SOLVE_NB - 1) / SOLVE_NB;
    const long long tstride = (long long)2 * nblk * SOLVE_NB * SOLVE_NB;      // per matrix
    const cplx* Uinv = tinv;
    const cplx* Linv = tinv + (size_t)nblk * SOLVE_NB * SOLVE_NB;
    gather_cols_kernel<<<dim3(nrows, nb), 256, 0, st>>>(Bm, bstride, ldb, perm, n, X, xstride, ldx);
#define SK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)
    // forward:  Y_k = (Xw_k - Y_{<k} U_{<k,k}) Uinv_k          (Xw = X buffer, Y = Yw buffer)
    for (int kb = 0; kb < nblk; ++kb) {
        const int c0 = kb * SOLVE_NB, w = (n - c0 < SOLVE_NB) ? n - c0 : SOLVE_NB;
        if (kb > 0)
            SK(zgemm_strided(OP_N, OP_N, nrows, w, c0, mone, Yw, ldx, xstride, LU + c0, lda, lustride, one, X + c0, ldx, xstride, nb, gscratch, st));
        SK(zgemm_strided(OP_N, OP_N, nrows, w, w, one, X + c0, ldx, xstride, Uinv + (size_t)kb * SOLVE_NB * SOLVE_NB, SOLVE_NB, tstride,
                         zero, Yw + c0, ldx, xstride, nb, gscratch, st));
    }
    // backward: Z_k = (Y_k - Z_{>k} L_{>k,k}) Linv_k           (Z overwrites the X buffer)
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int c0 = kb * SOLVE_NB, w = (n - c0 < SOLVE_NB) ? n - c0 : SOLVE_NB, c1 = c0 + w;
        if (c1 < n)
            SK(zgemm_strided(OP_N, OP_N, nrows, w, n - c1, mone, X + c1, ldx, xstride, LU + (size_t)c1 * lda + c0, lda, lustride, one,
                             Yw + c0, ldx, xstride, nb, gscratch, st));
        SK(zgemm_strided(OP_N, OP_N, nrows, w, w, one, Yw + c0, ldx, xstride, Linv + (size_t)kb * SOLVE_NB * SOLVE_NB, SOLVE_NB, tstride,
                         zero, X + c0, ldx, xstride, nb, gscratch, st));
    }
#undef SK
    return cudaGetLastError();
}

}  // namespace rcwa
#endif
