// Stage 2: batched complex128 non-Hermitian eigendecomposition, entirely on the device.
// Replaces torch.linalg.eig in Eig.forward (/root/reference/torcwa/torch_eig.py:11-17; LAPACK
// zgeev on CPU, cuSOLVER/MAGMA hybrid on CUDA).  Three phases per batch of matrices:
//
//  (1) Blocked Householder Hessenberg reduction  A = Z H Z^H  (hess.cu): per column one read-only
//      streaming mat-vec over the trailing matrix (the HBM-bound kernel of the stage, exactly the
//      algorithmic 16 n^3/3 bytes), per 64-column panel compact-WY block updates of A and Z on the
//      DMMA GEMM.
//  (2) Windowed multishift QR with aggressive early deflation, Z <- Z U:  chains of up to QR_NS
//      single-shift Givens bulges (spacing 2) are chased through a QR_W x QR_W diagonal window held
//      in shared memory by ONE CTA per matrix, all bulges advancing one position per step; the
//      window's accumulated unitary U is then applied to the off-diagonal row panel (main stream),
//      the column panel and Z (side stream) by grouped DMMA GEMMs (zgemm.cu).  H is kept current
//      only inside the active block.  AED: the trailing 48 x 48 window is Schur-factored on a copy
//      (two-phase explicit-shift QR, barrier-free warp version), converged eigenvalues are deflated
//      by the spike test, the rest become the next shifts.  Every serial piece is a resumable time
//      slice (SM-clock budget): a launch lasts as long as its slowest matrix.  The batch runs as
//      independently pipelined groups; the host only enqueues and polls a device counter through
//      pinned memory.
//  (3) Schur form T = Z^H A0 Z (upper tiles), eigenvectors of T by blocked back-substitution (one
//      triangular-hinted GEMM + one per-column small triangular solve per 32-row block), then
//      V = Z X (GEMM) and unit 2-norm columns (LAPACK geev convention).
//
// The single-CTA bodies (qr pass, shift solver, triangular solves) are phase-structured and are
// also compiled for the CPU by the emulation build (-DRCWA_EMU, tests only).
#include "common.cuh"
#ifndef RCWA_EMU
#include "kernels.h"
#else
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#endif

#define QR_W 64            // window size
#define QR_LD 65           // shared-memory leading dimension (odd: conflict-free column access)
#define QR_NS 16           // max simultaneous shifts / bulges
#define QR_SMALL 48         // active blocks up to this size are Schur-factored directly in shared memory
#ifndef QR_AED_W
#define QR_AED_W 48         // aggressive-early-deflation window
#endif
// Time-slice budgets of the serial pieces of a pass.  A launch lasts as long as its slowest matrix, and with
// ~100 matrices in a batch some matrix is in its most expensive segment in EVERY launch (measured: mean
// own work 117 us per pass, launch duration 400 us), so every segment type is cut to about the duration of a
// bulge-chain window (60-80 us).
struct QrBudget {
    int schur;      // rotations per launch while Schur-factoring an AED window or a small block
    int swaps;      // eigenvalue swaps per launch during the AED deflation scan
    int restore;    // Householder steps per launch while folding the AED spike back to Hessenberg form
    long long cycles;   // SM-clock budget of a serial slice (0: count budgets only); the counts above are upper bounds
    int aed_w;          // aggressive-early-deflation window of this call (<= QR_AED_W, the size the buffers are laid out for)
};
#define QR_BUDGET_SCHUR 4000
#define QR_BUDGET_SWAPS 300
#define QR_BUDGET_RESTORE 32
#define QR_BUDGET_US 90

// Deadline test of a time slice, uniform over the cooperating warp (lane 0 reads the clock).
DEV bool slice_expired(const Cta& c, long long deadline) {
#ifdef RCWA_EMU
    (void)c; (void)deadline; return false;
#else
    if (deadline == 0 || !c.warp_only) return false;
    int e = (clock64() > deadline) ? 1 : 0;
    e = __shfl_sync(0xffffffffu, e, 0);
    return e != 0;
#endif
}
#define QR_AED_W_SMALL 24   // AED window for matrices up to QR_AED_SMALL_N ...
#define QR_AED_W_LARGE 32   // ... and above (the buffers are laid out for QR_AED_W = 48, the largest allowed)
#define QR_AED_SMALL_N 1024
#define QR_MODE_ALL 0
#define QR_MODE_WIN 1
#define QR_MODE_SMALL 2
#define QR_ROWS_SMALL ((QR_SMALL > QR_AED_W) ? QR_SMALL : QR_AED_W)     // rows of the window buffers a small dense solve needs
#define QR_GRAPH_ITERS 8     // iterations of one group recorded into one CUDA graph (even: the window-unitary buffers alternate)
#define QR_MAXG 8           // most independently pipelined groups of the QR phase (host_flag holds 2 ints per group)
#define QR_SMS 148         // SMs of a B200 (the pass kernel runs one CTA per SM)
#define QR_MAXSTALL 40     // sweeps without deflation before giving up on a matrix
#define TV_NB 32           // eigenvector back-substitution block

struct QrState {
    int lo, hi;          // active block [lo, hi] (inclusive); done when hi < 1
    int phase;           // 0: start a new sweep (deflation scan + shifts); 1: chain in flight; 2: small-block solve; 3: AED window Schur
    int p;               // window start of the next pass
    int nbulge;          // bulges currently in flight
    int nintro;          // bulges introduced so far in this sweep
    int ns;              // shifts of this sweep
    int stall;           // consecutive sweeps without any deflation
    int sweeps, passes;  // statistics
    int done, info;
    int hi_prev;
    int ss_i;            // small-block solve (phase 2): current bottom row of the unconverged part (local)
    int ss_its;          //   QR iterations spent on the current eigenvalue
    int ss_fresh;        //   1: U must be initialised to identity at the next slice
    int small_solves;    // statistics
    int aed_kw, aed_nw;  // AED window [aed_kw, aed_kw + aed_nw) (phase 3)
    int p_last;          // window start of the previous chase pass (its column update may still be in flight)
    int aed_off;         // AED disabled for this matrix (its window failed to converge)
    int aed_stage;       // 0: Schur slices, 1: deflation scan slices, 2: Hessenberg-restore slices (then finish)
    int aed_prog[4];     // scan progress: ns, ilst, knt, kcur
    int shifts_ready;    // phase 1 may start with st.shifts as they are (supplied by AED)
    int aeds, aed_deflated;   // statistics
    int pad0, pad1;
    long long band_active, band_total;   // profiling: k groups the banded update GEMMs keep / would run dense
    long long cyc_max[6]; // profiling: longest single segment
    long long cyc[6];    // profiling: SM cycles spent per pass segment (0 sweep start: scan + shifts, 1 chase, 2 small-block
    int cnt[6];          //   slice, 3 AED Schur slice, 4 AED scan slice, 5 AED finish: scan + restore + write-back) and counts
    int kpos[QR_NS];     // column of each bulge (leading first): bulge element is H[k+2][k]
    cplx shifts[QR_NS];
};

// ------------------------------------------------------------------------------------------------
// Givens rotation G = [[c, s], [-conj(s), c]] with G [a; b] = [r; 0], c real >= 0 (LAPACK zlartg).
HD void givens(cplx a, cplx b, double& c, cplx& s, cplx& r) {
    if (cis_zero(b)) { c = 1.0; s = C(0, 0); r = a; return; }
    if (cis_zero(a)) { c = 0.0; double nb = cabs_(b); s = cscale(cconj(b), 1.0 / nb); r = C(nb, 0); return; }
    const double na2 = cabs2(a), nb2 = cabs2(b), n2 = na2 + nb2;
    if (na2 > 1e-280 && nb2 > 1e-280 && n2 < 1e280) {
        // common, well-scaled case.  This sits on the serial critical path of every chase step and of every
        // small-Schur rotation: two INDEPENDENT reciprocal square roots (they pipeline) and multiplications
        // instead of two square roots and three divisions in sequence (several hundred cycles in fp64).
#ifdef RCWA_EMU
        const double ina = 1.0 / sqrt(na2), inr = 1.0 / sqrt(n2);
#else
        const double ina = rsqrt(na2), inr = rsqrt(n2);
#endif
        const double w = ina * inr;                 // 1 / (|a| * norm)
        c = na2 * w;                                // |a| / norm
        s = cscale(cmul(a, cconj(b)), w);
        r = cscale(a, n2 * w);                      // a * norm / |a|
        return;
    }
    const double na = cabs_(a), nb = cabs_(b);
    const double sc = fmax(na, nb);
    const double nrm = sc * sqrt((na / sc) * (na / sc) + (nb / sc) * (nb / sc));
    c = na / nrm;
    const cplx ph = cscale(a, 1.0 / na);                 // a/|a|
    s = cscale(cmul(ph, cconj(b)), 1.0 / nrm);
    r = cscale(ph, nrm);
}

// LAPACK zlahqr deflation test for subdiagonal h10 = H[k][k-1] given its neighbours.
HD bool negligible_subdiag(cplx h10, cplx h00, cplx h11, cplx h01, double extra) {
    const double smlnum = RCWA_SAFMIN * (1.0 / RCWA_EPS);
    const double a10 = cabs1(h10);
    if (a10 <= smlnum) return true;
    double tst = cabs1(h00) + cabs1(h11);
    if (tst == 0.0) tst = extra;
    if (a10 > RCWA_EPS * tst) return false;     // subdiagonals are complex here (zlahqr makes them real first)
    const double a01 = cabs1(h01);
    const double ab = fmax(a10, a01), ba = fmin(a10, a01);
    const double d = cabs1(csub(h00, h11));
    const double aa = fmax(cabs1(h11), d), bb = fmin(cabs1(h11), d);
    const double s = aa + ab;
    return ba * (ab / s) <= fmax(smlnum, RCWA_EPS * (bb * (aa / s)));
}

// ------------------------------------------------------------------------------------------------
// Eigenvalues of a small upper-Hessenberg matrix (m <= QR_NS) held in shared memory, by ONE WARP
// (lanes loop over columns / rows; WARP_SYNC between dependent phases).  Single-shift implicit QR
// with Wilkinson shifts (zlahqr without Schur vectors; only the active block is updated).
// Returns the number of eigenvalues that failed to converge (their diagonal entry is returned).
DEV int tiny_hqr_eigs(int lane, int nlanes, cplx* T, int ldt, int m, cplx* wout) {
    int fails = 0;
    int i = m - 1;
    int guard = 0;
    while (i >= 0 && guard < 64 * QR_NS) {
        int l = 0, its = 0;
        bool conv = false;
        for (its = 0; its <= 40; ++its, ++guard) {
            // locate a negligible subdiagonal (uniform: every lane evaluates the same scalars)
            for (l = i; l > 0; --l) {
                cplx h10 = T[l * ldt + l - 1];
                double extra = 0.0;
                if (l - 2 >= 0) extra += cabs1(T[(l - 1) * ldt + l - 2]);
                if (l + 1 <= i) extra += cabs1(T[(l + 1) * ldt + l]);
                if (negligible_subdiag(h10, T[(l - 1) * ldt + l - 1], T[l * ldt + l], T[(l - 1) * ldt + l], extra)) break;
            }
            WARP_SYNC();
            if (l > 0 && lane == 0) T[l * ldt + l - 1] = C(0, 0);
            WARP_SYNC();
            if (l >= i) { conv = true; break; }
            // shift
            cplx sig;
            if (its == 10 || its == 20 || its == 30) {
                sig = cadd(T[l * ldt + l], C(0.75 * cabs1(T[(l + 1) * ldt + l]), 0.0));
            } else {
                // Wilkinson: eigenvalue of [[a,b],[c,d]] (trailing 2x2) closer to d
                cplx a = T[(i - 1) * ldt + i - 1], b = T[(i - 1) * ldt + i], cc = T[i * ldt + i - 1], d = T[i * ldt + i];
                cplx tr2 = cscale(csub(a, d), 0.5);
                cplx disc = csqrt_(cadd(cmul(tr2, tr2), cmul(b, cc)));
                // d + (bc)/(tr2 +- disc) with the larger denominator
                cplx den1 = cadd(tr2, disc), den2 = csub(tr2, disc);
                cplx den = (cabs2(den1) >= cabs2(den2)) ? den1 : den2;
                sig = cis_zero(den) ? d : csub(d, cdiv(cmul(b, cc), den));
            }
            // one QR sweep l..i
            for (int k = l; k < i; ++k) {
                cplx a, b;
                if (k == l) { a = csub(T[k * ldt + k], sig); b = T[(k + 1) * ldt + k]; }
                else { a = T[k * ldt + k - 1]; b = T[(k + 1) * ldt + k - 1]; }
                double cs; cplx sn, r;
                givens(a, b, cs, sn, r);
                WARP_SYNC();
                if (k > l && lane == 0) { T[k * ldt + k - 1] = r; T[(k + 1) * ldt + k - 1] = C(0, 0); }
                // left: rows k, k+1, columns k..i
                for (int j = k + lane; j <= i; j += nlanes) {
                    cplx x = T[k * ldt + j], y = T[(k + 1) * ldt + j];
                    T[k * ldt + j] = cadd(cscale(x, cs), cmul(sn, y));
                    T[(k + 1) * ldt + j] = csub(cscale(y, cs), cmul(cconj(sn), x));
                }
                WARP_SYNC();
                // right: columns k, k+1, rows l..min(k+2, i)
                const int rmax = (k + 2 < i) ? k + 2 : i;
                for (int r2 = l + lane; r2 <= rmax; r2 += nlanes) {
                    cplx x = T[r2 * ldt + k], y = T[r2 * ldt + k + 1];
                    T[r2 * ldt + k] = cadd(cscale(x, cs), cmul(y, cconj(sn)));
                    T[r2 * ldt + k + 1] = csub(cscale(y, cs), cmul(x, sn));
                }
                WARP_SYNC();
            }
        }
        if (!conv) { ++fails; l = i; }
        // block [l..i] with l == i (converged) or forced
        i = l - 1;
        if (conv) { /* eigenvalue at index l == old i */ }
    }
    WARP_SYNC();
    for (int j = lane; j < m; j += nlanes) wout[j] = T[j * ldt + j];
    WARP_SYNC();
    return fails;
}

// ------------------------------------------------------------------------------------------------
// Largest l in (0, i] whose subdiagonal H[l][l-1] is negligible (LAPACK zlahqr test); 0 if none.
// Warp version: every lane tests one candidate, one ballot per 32 candidates.
DEV int schur_find_split(const Cta& c, const cplx* Hs, int i) {
#ifndef RCWA_EMU
    if (c.warp_only && c.nthreads == 32) {
        for (int base = i; base > 0; base -= 32) {
            const int k = base - c.tid;
            bool z = false;
            if (k > 0) {
                const cplx h10 = Hs[k * QR_LD + k - 1];
                z = cis_zero(h10);
                if (!z) {
                    double extra = 0.0;
                    if (k - 2 >= 0) extra += cabs1(Hs[(k - 1) * QR_LD + k - 2]);
                    if (k + 1 <= i) extra += cabs1(Hs[(k + 1) * QR_LD + k]);
                    z = negligible_subdiag(h10, Hs[(k - 1) * QR_LD + k - 1], Hs[k * QR_LD + k], Hs[(k - 1) * QR_LD + k], extra);
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, z);
            if (mask) return base - (__ffs(mask) - 1);
        }
        return 0;
    }
#endif
    int l;
    for (l = i; l > 0; --l) {
        cplx h10 = Hs[l * QR_LD + l - 1];
        if (cis_zero(h10)) break;
        double extra = 0.0;
        if (l - 2 >= 0) extra += cabs1(Hs[(l - 1) * QR_LD + l - 2]);
        if (l + 1 <= i) extra += cabs1(Hs[(l + 1) * QR_LD + l]);
        if (negligible_subdiag(h10, Hs[(l - 1) * QR_LD + l - 1], Hs[l * QR_LD + l], Hs[(l - 1) * QR_LD + l], extra)) break;
    }
    return l;
}

// ------------------------------------------------------------------------------------------------
// Resumable single-shift QR (Schur form with Schur vectors) on an m x m upper-Hessenberg block held in
// shared memory (Hs, Us with leading dimension QR_LD; Us accumulates the unitary).  Runs at most
// `budget` rotations, then returns; state (i, its) lives in QrState so the next launch continues.
// Returns 1 when the block is upper triangular, 0 if more slices are needed, -1 on failure.
//
// Each QR iteration on the active block [l, i] is an EXPLICIT shifted step, H - sig I = Q R,
// H' = R Q + sig I, in two phases that need one group barrier per rotation instead of three:
//   phase 1 (thread j owns COLUMN j): for k = l..i-1 every thread forms the same rotation G_k from
//           (R[k][k], H[k+1][k]) and applies it to rows (k, k+1) of its own columns; the rotations are
//           parked in rot_c / rot_s;
//   phase 2 (thread r owns ROW r of H or of U): applies the whole parked sequence to its row, carrying
//           the running element in a register -- no barriers at all.
// sig is subtracted from / added back to the diagonal of the active block only (the rotations are the
// identity outside it, so this is the same similarity as the implicit step).
DEV int small_schur_slice(const Cta& c, cplx* Hs, cplx* Us, int m, int* pi, int* pits, int budget, long long deadline, double* rot_c, cplx* rot_s) {
    int i = *pi, its = *pits, used = 0;
    while (i >= 1) {
        const int l = schur_find_split(c, Hs, i);
        GROUP_SYNC(c);
        if (l > 0 && c.tid == 0) Hs[l * QR_LD + l - 1] = C(0, 0);
        GROUP_SYNC(c);
        if (l >= i) { --i; its = 0; continue; }
        if (its > 60) { *pi = i; *pits = its; return -1; }
        if (used > 0 && (used + (i - l) > budget || slice_expired(c, deadline))) break;       // out of time: resume at the next launch
        // ---- shift
        cplx sig;
        if (its == 10 || its == 30) sig = cadd(Hs[l * QR_LD + l], C(0.75 * cabs1(Hs[(l + 1) * QR_LD + l]), 0.0));
        else if (its == 20 || its == 40) sig = cadd(Hs[i * QR_LD + i], C(0.75 * cabs1(Hs[i * QR_LD + i - 1]), 0.0));
        else {
            cplx a = Hs[(i - 1) * QR_LD + i - 1], b = Hs[(i - 1) * QR_LD + i], cc = Hs[i * QR_LD + i - 1], d = Hs[i * QR_LD + i];
            cplx tr2 = cscale(csub(a, d), 0.5);
            cplx disc = csqrt_(cadd(cmul(tr2, tr2), cmul(b, cc)));
            cplx den1 = cadd(tr2, disc), den2 = csub(tr2, disc);
            cplx den = (cabs2(den1) >= cabs2(den2)) ? den1 : den2;
            sig = cis_zero(den) ? d : csub(d, cdiv(cmul(b, cc), den));
        }
        GROUP_SYNC(c);                                         // every thread has read the shift operands
        for (int d = l + c.tid; d <= i; d += c.nthreads) Hs[d * QR_LD + d] = csub(Hs[d * QR_LD + d], sig);
        GROUP_SYNC(c);
        // ---- phase 1: R = Q^H (H - sig I)
#ifndef RCWA_EMU
        if (c.warp_only && c.nthreads == 32 && m <= 64) {
            // Warp version without any barrier inside the rotation loop: lane t owns columns t and t + 32 and keeps
            // the running row (row k after G_{k-1}) of its columns in registers; the pivot element travels by
            // shuffle; row k + 1 is still the untouched original in shared memory when step k reads it; final rows
            // are stored as they complete, and the zeroed subdiagonal is written after the loop (nobody may clear
            // H[k+1][k] while another lane can still be reading it as the rotation's second operand).
            const int lane = c.tid;
            cplx top[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) { const int j = lane + 32 * q; top[q] = (j >= l && j < m) ? Hs[l * QR_LD + j] : C(0, 0); }
            for (int k = l; k < i; ++k) {
                const int ko = k & 31, so = k >> 5;
                cplx bot[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) { const int j = lane + 32 * q; bot[q] = (j > k && j < m) ? Hs[(k + 1) * QR_LD + j] : C(0, 0); }
                const cplx y = Hs[(k + 1) * QR_LD + k];
                cplx x;
                x.x = __shfl_sync(0xffffffffu, so ? top[1].x : top[0].x, ko);
                x.y = __shfl_sync(0xffffffffu, so ? top[1].y : top[0].y, ko);
                double cs; cplx sn, r;
                givens(x, y, cs, sn, r);
                if (lane == ko) { rot_c[k] = cs; rot_s[k] = sn; Hs[k * QR_LD + k] = r; }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int j = lane + 32 * q;
                    if (j > k && j < m) {
                        Hs[k * QR_LD + j] = cadd(cscale(top[q], cs), cmul(sn, bot[q]));
                        top[q] = csub(cscale(bot[q], cs), cmul(cconj(sn), top[q]));
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) { const int j = lane + 32 * q; if (j >= i && j < m) Hs[i * QR_LD + j] = top[q]; }
            GROUP_SYNC(c);
            for (int k = l + lane; k < i; k += 32) Hs[(k + 1) * QR_LD + k] = C(0, 0);
            GROUP_SYNC(c);
        } else
#endif
        {
        cplx r_prev = C(0, 0);
        for (int k = l; k < i; ++k) {
            double cs; cplx sn, r;
            givens(Hs[k * QR_LD + k], Hs[(k + 1) * QR_LD + k], cs, sn, r);
            if (c.tid == 0) {
                rot_c[k] = cs; rot_s[k] = sn;
                // column k-1 gets its final (r, 0) one step late: every thread has read it by now
                if (k > l) { Hs[(k - 1) * QR_LD + k - 1] = r_prev; Hs[k * QR_LD + k - 1] = C(0, 0); }
            }
            r_prev = r;
            for (int j = k + 1 + c.tid; j < m; j += c.nthreads) {
                const cplx x = Hs[k * QR_LD + j], y = Hs[(k + 1) * QR_LD + j];
                Hs[k * QR_LD + j] = cadd(cscale(x, cs), cmul(sn, y));
                Hs[(k + 1) * QR_LD + j] = csub(cscale(y, cs), cmul(cconj(sn), x));
            }
            GROUP_SYNC(c);
        }
        if (c.tid == 0) { Hs[(i - 1) * QR_LD + i - 1] = r_prev; Hs[i * QR_LD + i - 1] = C(0, 0); }
        GROUP_SYNC(c);
        }
        // ---- phase 2: H' = R Q (rows 0..i; row r only holds columns >= r) and U <- U Q (all m rows)
        for (int idx = c.tid; idx < (i + 1) + m; idx += c.nthreads) {
            cplx* row; int ks;
            if (idx <= i) { row = Hs + idx * QR_LD; ks = (idx - 1 > l) ? idx - 1 : l; }
            else { row = Us + (idx - i - 1) * QR_LD; ks = l; }
            cplx x = row[ks];
            for (int k = ks; k < i; ++k) {
                const cplx y = row[k + 1];
                const double cs = rot_c[k]; const cplx sn = rot_s[k];
                row[k] = cadd(cscale(x, cs), cmul(y, cconj(sn)));
                x = csub(cscale(y, cs), cmul(x, sn));
            }
            row[i] = x;
        }
        GROUP_SYNC(c);
        for (int d = l + c.tid; d <= i; d += c.nthreads) Hs[d * QR_LD + d] = cadd(Hs[d * QR_LD + d], sig);
        GROUP_SYNC(c);
        used += i - l;
        ++its;
    }
    *pi = i; *pits = its;
    return (i < 1) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// One window pass of the multishift QR for one matrix.  H: n x n (ldh), state in global memory.
// Outputs: U (QR_W x QR_W, ld QR_W, global), three GEMM problems (rows, cols, Z) -- M = 0 if idle.
// Shared memory: Hs[QR_W*QR_LD] + Us[QR_W*QR_LD] cplx + small scratch (see qr_pass_smem_bytes()).
struct QrScratch {
    QrState st;                 // working copy of the per-matrix state (thread 0 mutates, barriers publish)
    double cs[QR_NS];
    cplx sn[QR_NS];
    int act[QR_NS];             // rotation of bulge slot b exists in the current step
    int rr1[QR_NS];             // its upper local row (0 for an introduction)
    int start[QR_NS];           // closed-form schedule of the pass: first step of slot b,
    int q0[QR_NS];              //   local column at that step (-1 = virtual column of an introduction),
    int nrot[QR_NS];            //   number of rotations it performs in this pass,
    int shift_id[QR_NS];        //   shift used by an introduction
    int nslots, tmax;
    int aed_ns;                 // result of the AED deflation analysis
    int ss_rc, ss_i, ss_its;    // results of a small-Schur slice run by warp 0
    cplx hv[QR_W];              // Householder vector (AED restore)
    double red[40];
    unsigned char col_lo[QR_W], col_hi[QR_W];   // first / last nonzero row of each column of the window unitary
    unsigned char band_lo[8], band_hi[8];       // per 8-column tile: range of 4-row groups holding nonzeros
    double rot_c[QR_W];         // rotations of one small-Schur iteration (phase 1 -> phase 2)
    cplx rot_s[QR_W];
};

HD size_t qr_pass_smem_bytes(int n, int mode = QR_MODE_ALL) {
    const size_t rows = (mode == QR_MODE_SMALL) ? QR_ROWS_SMALL : QR_W;
    return 2 * rows * QR_LD * sizeof(cplx) + sizeof(QrScratch) + (size_t)(QR_NS * (QR_NS + 1)) * sizeof(cplx)
           + ((mode == QR_MODE_SMALL) ? 0 : (size_t)(n + 16)) + 64;
}

// ------------------------------------------------------------------------------------------------
// AED helpers.  T (upper triangular Schur form of the window, nw x nw) in Hs, its Schur vectors in Us.
//
// Deflation analysis (LAPACK zlaqr2): the window's coupling to the rest of H is the "spike"
// s * conj(V[0,:]); eigenvalues whose spike entry is negligible are converged; the others are moved
// to the top with adjacent swaps (ztrexc).  Returns ns = number of undeflatable eigenvalues, which
// then occupy T[0:ns, 0:ns].
// Resumable: at most `budget` swaps per call; progress (ns, ilst, knt, kcur) lives in `prog[4]`.
// Returns 1 when the scan is complete (prog[0] = ns), 0 if it must be continued.
DEV int aed_deflation_scan(const Cta& c, cplx* T, cplx* V, int nw, cplx s, int* prog, int budget, long long deadline) {
    const double smlnum = RCWA_SAFMIN * (1.0 / RCWA_EPS);
    int ns = prog[0], ilst = prog[1], knt = prog[2], kcur = prog[3], used = 0;
    while (knt < nw) {
        if (kcur < 0) {
            double foo = cabs1(T[(ns - 1) * QR_LD + ns - 1]);
            if (foo == 0.0) foo = cabs1(s);
            if (cabs1(s) * cabs1(V[ns - 1]) <= fmax(smlnum, RCWA_EPS * foo)) { --ns; ++knt; continue; }
            kcur = ns - 2;            // undeflatable: move position ns-1 up to ilst by adjacent swaps
        }
        while (kcur >= ilst) {
            if (used >= budget || (used > 0 && (used & 7) == 0 && slice_expired(c, deadline))) { GROUP_SYNC(c); prog[0] = ns; prog[1] = ilst; prog[2] = knt; prog[3] = kcur; return 0; }
            const int k = kcur;
            const cplx t11 = T[k * QR_LD + k], t22 = T[(k + 1) * QR_LD + k + 1];
            double cs; cplx sn, r;
            givens(T[k * QR_LD + k + 1], csub(t22, t11), cs, sn, r);
            // rows (k,k+1) x cols k+2..nw-1 | cols (k,k+1) x rows 0..k-1 of T | cols (k,k+1) x all rows of V
            const int n_left = nw - (k + 2), n_right = k;
            for (int idx = c.tid; idx < n_left + n_right + nw; idx += c.nthreads) {
                if (idx < n_left) {
                    const int j = k + 2 + idx;
                    cplx x = T[k * QR_LD + j], y = T[(k + 1) * QR_LD + j];
                    T[k * QR_LD + j] = cadd(cscale(x, cs), cmul(sn, y));
                    T[(k + 1) * QR_LD + j] = csub(cscale(y, cs), cmul(cconj(sn), x));
                } else {
                    cplx* base = (idx < n_left + n_right) ? (T + (idx - n_left) * QR_LD) : (V + (idx - n_left - n_right) * QR_LD);
                    cplx x = base[k], y = base[k + 1];
                    base[k] = cadd(cscale(x, cs), cmul(y, cconj(sn)));
                    base[k + 1] = csub(cscale(y, cs), cmul(x, sn));
                }
            }
            GROUP_SYNC(c);
            if (c.tid == 0) { T[k * QR_LD + k] = t22; T[(k + 1) * QR_LD + k + 1] = t11; }
            GROUP_SYNC(c);
            --kcur; ++used;
        }
        ++ilst; ++knt; kcur = -1;
    }
    prog[0] = ns; prog[1] = ilst; prog[2] = knt; prog[3] = kcur;
    return 1;
}

// Hermitian unitary reflector H = I - u u^H (|u|^2 = 2) with H x = beta e_1 for x = M[r0.., col] (rows
// r0..r1-1 of column `col`, stride QR_LD), or for an explicit vector when M == nullptr (then `u`
// holds x on entry).  Thread 0 builds u (indexed by absolute row) into `u`; returns beta via *beta.
DEV void small_reflector(const Cta& c, cplx* u, int r0, int r1, cplx* beta_out) {
    // every thread forms the same scalars (partial sums combined by warp shuffles on the device)
    double ss = 0.0;
#ifndef RCWA_EMU
    if (c.warp_only && c.nthreads == 32) {
        for (int r = r0 + c.tid; r < r1; r += 32) ss += cabs2(u[r]);
        ss = warp_sum(ss);
    } else
#endif
    { for (int r = r0; r < r1; ++r) ss += cabs2(u[r]); }
    const double sigma = sqrt(ss);
    const cplx x1 = u[r0];
    const double ax = cabs_(x1);
    GROUP_SYNC(c);                                   // all reads of u done before anyone rewrites it
    if (sigma == 0.0) {
        for (int r = r0 + c.tid; r < r1; r += c.nthreads) u[r] = C(0, 0);
        *beta_out = x1;
    } else {
        const cplx ph = (ax == 0.0) ? C(1, 0) : cscale(x1, 1.0 / ax);
        const cplx beta = cscale(ph, -sigma);
        const double scl = 1.0 / sqrt(sigma * (sigma + ax));
        for (int r = r0 + c.tid; r < r1; r += c.nthreads) {
            cplx v = u[r];
            if (r == r0) v = csub(v, beta);
            u[r] = cscale(v, scl);
        }
        *beta_out = beta;
    }
    GROUP_SYNC(c);
}

// M <- (I - u u^H) M on rows r0..r1-1, columns c0..c1-1 (one thread per column: dot then update)
DEV void small_hh_left(const Cta& c, cplx* M, const cplx* u, int r0, int r1, int c0, int c1) {
    for (int j = c0 + c.tid; j < c1; j += c.nthreads) {
        cplx w = C(0, 0), w2 = C(0, 0);              // two chains: the dot product is latency-bound
        int r = r0;
        for (; r + 1 < r1; r += 2) { w = cadd(w, cmulc(u[r], M[r * QR_LD + j])); w2 = cadd(w2, cmulc(u[r + 1], M[(r + 1) * QR_LD + j])); }
        if (r < r1) w = cadd(w, cmulc(u[r], M[r * QR_LD + j]));
        w = cadd(w, w2);
        for (r = r0; r < r1; ++r) M[r * QR_LD + j] = csub(M[r * QR_LD + j], cmul(u[r], w));
    }
}
// M <- M (I - u u^H) on rows q0..q1-1, columns r0..r1-1 (one thread per row)
DEV void small_hh_right(const Cta& c, cplx* M, const cplx* u, int q0, int q1, int r0, int r1) {
    for (int i = q0 + c.tid; i < q1; i += c.nthreads) {
        cplx y = C(0, 0), y2 = C(0, 0);
        int r = r0;
        for (; r + 1 < r1; r += 2) { y = cfma(M[i * QR_LD + r], u[r], y); y2 = cfma(M[i * QR_LD + r + 1], u[r + 1], y2); }
        if (r < r1) y = cfma(M[i * QR_LD + r], u[r], y);
        y = cadd(y, y2);
        for (r = r0; r < r1; ++r) M[i * QR_LD + r] = csub(M[i * QR_LD + r], cmul(y, cconj(u[r])));
    }
}

// After the scan: fold the remaining spike s*conj(V[0,0:ns]) back into Hessenberg form.
// T[0:ns,0:ns] <- Hessenberg, T[0:ns, ns:nw] and V[:,0:ns] updated accordingly (LAPACK zlaqr2:
// zlarfg on the spike, zlarf x3, zgehrd, zunmhr -- done here with Hermitian reflectors).
// Resumable: *pj = -1 on entry of the first slice (the spike reflector), then the column index of the
// reduction; at most `budget` Householder steps per call.  Returns 1 when finished.
DEV int aed_restore_hessenberg(const Cta& c, cplx* T, cplx* V, int nw, int ns, cplx* u, int* pj, int budget, long long deadline) {
    if (ns <= 1) return 1;
    cplx beta;
    int j = *pj, used = 0;
    if (j < 0) {
        // reflector that maps the spike direction conj(V[0,0:ns]) onto e_0
        for (int r = c.tid; r < ns; r += c.nthreads) u[r] = cconj(V[r]);
        GROUP_SYNC(c);
        small_reflector(c, u, 0, ns, &beta);
        small_hh_left(c, T, u, 0, ns, 0, nw);
        GROUP_SYNC(c);
        small_hh_right(c, T, u, 0, ns, 0, ns);
        small_hh_right(c, V, u, 0, nw, 0, ns);
        GROUP_SYNC(c);
        j = 0; used = 2;
    }
    // Hessenberg reduction of T[0:ns,0:ns]; reflectors act on indices >= 1, so the spike stays on e_0
    for (; j + 2 < ns; ++j) {
        if (used >= budget || (used > 0 && slice_expired(c, deadline))) { *pj = j; return 0; }
        for (int r = j + 1 + c.tid; r < ns; r += c.nthreads) u[r] = T[r * QR_LD + j];
        GROUP_SYNC(c);
        small_reflector(c, u, j + 1, ns, &beta);
        if (c.tid == 0) {
            T[(j + 1) * QR_LD + j] = beta;
            for (int r = j + 2; r < ns; ++r) T[r * QR_LD + j] = C(0, 0);
        }
        small_hh_left(c, T, u, j + 1, ns, j + 1, nw);
        GROUP_SYNC(c);
        small_hh_right(c, T, u, 0, ns, j + 1, ns);
        small_hh_right(c, V, u, 0, nw, j + 1, ns);
        GROUP_SYNC(c);
        ++used;
    }
    *pj = j;
    return 1;
}

// emit the three GEMM problems that apply the window unitary U (wl x wl at Ug) for the window [p, p+wl)
// of the active block [lo, hi].  During the QR phase H is only kept current INSIDE the active block
// (LAPACK's wantt = false): the row panel is updated up to column hi, the column panel from row lo.
// The off-diagonal blocks of the final Schur form are recovered afterwards in two large GEMMs,
// T = Z^H A0 Z (A0 = the input matrix, Z = all accumulated transformations) -- 2 n^3 complex MACs at full tensor rate instead of ~1/3 of all K = 64 panel updates.
DEV void emit_window_gemms(cplx* H, int ldh, int n, cplx* Zm, int ldz, cplx* Ug, int p, int wl, int lo, int hi,
                           ZGemmProblem* prob_rows, ZGemmProblem* prob_cols, ZGemmProblem* prob_z,
                           const unsigned char* band_lo = nullptr, const unsigned char* band_hi = nullptr) {
    const int wend = p + wl;
    const int ncol = hi + 1 - wend, nrow = p - lo;
    ZGemmProblem g;
    g.flags = 0;
    for (int t = 0; t < 8; ++t) { g.klo[t] = band_lo ? band_lo[t] : 0; g.khi[t] = band_hi ? band_hi[t] : 0; }
    g.A = Ug; g.lda = QR_W; g.B = H + (size_t)p * ldh + wend; g.ldb = ldh; g.C = H + (size_t)p * ldh + wend; g.ldc = ldh;
    g.M = (ncol > 0) ? wl : 0; g.N = ncol; g.K = wl; g.flags = band_lo ? ZGEMM_A_BAND : 0; *prob_rows = g;   // H[p:wend, wend:hi+1] <- U^H * (.)
    g.flags = band_lo ? ZGEMM_B_BAND : 0;
    g.A = H + (size_t)lo * ldh + p; g.lda = ldh; g.B = Ug; g.ldb = QR_W; g.C = H + (size_t)lo * ldh + p; g.ldc = ldh;
    g.M = (nrow > 0) ? nrow : 0; g.N = wl; g.K = wl; *prob_cols = g;                      // H[lo:p, p:wend] <- (.) * U
    g.A = Zm + p; g.lda = ldz; g.B = Ug; g.ldb = QR_W; g.C = Zm + p; g.ldc = ldz;
    g.M = n; g.N = wl; g.K = wl; *prob_z = g;                                            // Z[:, p:wend] <- (.) * U
}

#ifdef RCWA_EMU
#define QR_CLOCK() 0LL
#else
#define QR_CLOCK() clock64()
#endif
#define QR_ACCOUNT(seg) do { if (c.tid == 0) { const long long _t = QR_CLOCK(); st.cyc[seg] += _t - tseg; st.cnt[seg]++; if (_t - tseg > st.cyc_max[seg]) st.cyc_max[seg] = _t - tseg; tseg = _t; } } while (0)

DEV void qr_pass_body(const Cta& c, cplx* H, int ldh, int n, cplx* Zm, int ldz, QrState* stg,
                      cplx* Ug, cplx* Vg, cplx* Tg, ZGemmProblem* prob_rows, ZGemmProblem* prob_cols_main,
                      ZGemmProblem* prob_cols, ZGemmProblem* prob_z, QrBudget bud, int mode = QR_MODE_ALL) {
    // mode: QR_MODE_ALL -- one launch does whatever segment the matrix is in.  QR_MODE_WIN / QR_MODE_SMALL -- the same
    // iteration as TWO launches: the first handles the sweep start and the bulge-chase windows (64-row window buffers,
    // 512 threads, one CTA per SM), the second the time slices of the small dense solves (AED window, small active
    // block: at most QR_ROWS_SMALL rows, mostly one warp at work), launched with less shared memory and fewer threads so
    // that two of its CTAs share an SM.  A launch returns at once for matrices in a segment of the other kind.
    // Ug: this pass's window unitary (double-buffered by the host: its GEMMs may still run while the next
    //     pass works); Vg/Tg: persistent copies for the time-sliced AED / small-block solves.
    // prob_rows, prob_cols_main run on the main stream before the next pass; prob_cols, prob_z on the side
    // stream (they touch only rows above / columns of Z that later passes of a downward-moving chase never
    // read).  AED and small-block solves may be followed by a window that reaches upwards, so their column
    // update goes to the main stream.
    const int rows = (mode == QR_MODE_SMALL) ? QR_ROWS_SMALL : QR_W;     // see qr_pass_smem_bytes()
    cplx* Hs = reinterpret_cast<cplx*>(c.smem);
    cplx* Us = Hs + rows * QR_LD;
    QrScratch* sc = reinterpret_cast<QrScratch*>(Us + rows * QR_LD);
    cplx* Ts = reinterpret_cast<cplx*>(sc + 1);                         // [QR_NS][QR_NS+1] shift block
    unsigned char* negl = reinterpret_cast<unsigned char*>(Ts + QR_NS * (QR_NS + 1));   // [n] deflation flags (not in QR_MODE_SMALL)
    QrState& st = sc->st;

    // the second launch of an iteration must not clear what the first one emitted
    if (c.tid == 0) { if (mode != QR_MODE_SMALL) { prob_rows->M = 0; prob_cols_main->M = 0; prob_cols->M = 0; prob_z->M = 0; } st = *stg; }
    CTA_SYNC();
    if (st.done) return;
    if (mode != QR_MODE_ALL && (mode == QR_MODE_SMALL) != (st.phase == 2 || st.phase == 3)) return;
    long long tseg = QR_CLOCK();
    const long long deadline = (bud.cycles > 0) ? tseg + bud.cycles : 0;
    const bool ran0 = (st.phase == 0);

    if (st.phase == 0) {
        // ---------------- new sweep: deflation scan over the whole remaining matrix [1..hi]
        const int hi0 = st.hi;
        for (int k = 1 + c.tid; k <= hi0; k += c.nthreads) {
            cplx h10 = H[(size_t)k * ldh + k - 1];
            bool z = cis_zero(h10);
            if (!z) {
                double extra = 0.0;
                if (k - 2 >= 0) extra += cabs1(H[(size_t)(k - 1) * ldh + k - 2]);
                if (k + 1 <= hi0) extra += cabs1(H[(size_t)(k + 1) * ldh + k]);
                // H[k-1][k] belongs to the side-stream column update of the window that started at k; if that
                // window was the previous pass, it may still be in flight: use the plain criterion there
                const cplx h01 = (k == st.p_last) ? h10 : H[(size_t)(k - 1) * ldh + k];
                z = negligible_subdiag(h10, H[(size_t)(k - 1) * ldh + k - 1], H[(size_t)k * ldh + k], h01, extra);
                if (z) H[(size_t)k * ldh + k - 1] = C(0, 0);
            }
            negl[k] = z ? 1 : 0;
        }
        CTA_SYNC();
        if (c.tid == 0) {
            int hi = hi0;
            while (hi >= 1 && negl[hi]) --hi;
            int lo = hi;
            while (lo >= 1 && !negl[lo]) --lo;
            if (hi < 1) { st.done = 1; st.hi = hi; }
            else {
                if (hi < st.hi_prev || lo > st.lo) st.stall = 0; else st.stall++;
                st.hi_prev = hi; st.lo = lo; st.hi = hi;
                if (st.stall > QR_MAXSTALL) { st.done = 1; st.info = hi + 1; }
                else if (hi - lo + 1 <= QR_SMALL) {
                    // small active block: Schur-factor it directly in shared memory (time-sliced)
                    st.phase = 2; st.p = lo; st.ss_i = hi - lo; st.ss_its = 0; st.ss_fresh = 1; st.small_solves++;
                } else if (!st.aed_off) {
                    // aggressive early deflation on the trailing window before (or instead of) a sweep
                    const int aw = (bud.aed_w > 0 && bud.aed_w < QR_AED_W) ? bud.aed_w : QR_AED_W;
                    const int nw = (aw < hi - lo) ? aw : hi - lo;     // spike entry H[kw][kw-1] stays inside the block
                    st.phase = 3; st.aed_stage = 0; st.aed_nw = nw; st.aed_kw = hi - nw + 1; st.ss_i = nw - 1; st.ss_its = 0; st.ss_fresh = 1; st.aeds++;
                } else {
                    st.ns = QR_NS; st.nintro = 0; st.nbulge = 0; st.p = lo; st.phase = 1; st.sweeps++; st.shifts_ready = 0;
                }
            }
            if (st.done) *stg = st;
        }
        CTA_SYNC();
        if (st.done) return;
        if (st.phase == 1 && !st.shifts_ready) {
            // ---------------- shifts: eigenvalues of the trailing ns x ns block (warp 0)
            const int ns = st.ns, hi = st.hi;
            for (int idx = c.tid; idx < ns * ns; idx += c.nthreads) {
                int r = idx / ns, q = idx % ns;
                Ts[r * (QR_NS + 1) + q] = (r <= q + 1) ? H[(size_t)(hi - ns + 1 + r) * ldh + (hi - ns + 1 + q)] : C(0, 0);
            }
            CTA_SYNC();
#ifndef RCWA_EMU
            if (c.tid < 32) tiny_hqr_eigs(c.tid, 32, Ts, QR_NS + 1, ns, st.shifts);
#else
            tiny_hqr_eigs(0, 1, Ts, QR_NS + 1, ns, st.shifts);
#endif
            CTA_SYNC();
            if (c.tid == 0 && (st.stall % 6 == 5)) {
                // exceptional shifts (LAPACK-style) when the block has not deflated for a while
                const double mag = 0.75 * cabs1(H[(size_t)hi * ldh + hi - 1]);
                for (int j = 0; j < ns; ++j) st.shifts[j] = cadd(H[(size_t)hi * ldh + hi], C(mag * ((j & 1) ? -1.0 : 1.0), mag * 0.5 * (j % 3 - 1)));
            }
            CTA_SYNC();
        }
    }

    if (ran0) QR_ACCOUNT(0);
    if (mode == QR_MODE_WIN && (st.phase == 2 || st.phase == 3)) {        // the sweep start chose a small dense solve: next launch
        if (c.tid == 0) *stg = st;
        return;
    }

    if (st.phase == 3) {
        // ---------------- AED: time slices of the window's Schur factorisation on a COPY (Tg, Ug), then
        // the deflation analysis; H itself is only touched if something deflates
        const int kw = st.aed_kw, nw = st.aed_nw;
        const int fresh = st.ss_fresh;
        for (int idx = c.tid; idx < nw * nw; idx += c.nthreads) {
            int r = idx / nw, q = idx % nw;
            Hs[r * QR_LD + q] = fresh ? ((r <= q + 1) ? H[(size_t)(kw + r) * ldh + (kw + q)] : C(0, 0)) : Tg[r * QR_W + q];
            Us[r * QR_LD + q] = fresh ? C(r == q ? 1.0 : 0.0, 0.0) : Vg[r * QR_W + q];
        }
        CTA_SYNC();
        // latency-bound serial work: one warp with __syncwarp (a 512-thread barrier per rotation would dominate)
        Cta w1 = c; w1.warp_only = 1; w1.nthreads = (c.nthreads < 32) ? c.nthreads : 32;
        if (st.aed_stage == 0) {
            if (c.tid < w1.nthreads) {
                int si = st.ss_i, sits = st.ss_its;
                const int rc1 = small_schur_slice(w1, Hs, Us, nw, &si, &sits, bud.schur, deadline, sc->rot_c, sc->rot_s);
                if (c.tid == 0) { sc->ss_rc = rc1; sc->ss_i = si; sc->ss_its = sits; }
            }
            CTA_SYNC();
            const int rc = sc->ss_rc;
            if (rc < 0) {               // the window did not converge: fall back to plain sweeps for this matrix
                QR_ACCOUNT(3);
                if (c.tid == 0) { st.aed_off = 1; st.phase = 0; st.passes++; *stg = st; }
                return;
            }
            // park the copy; the deflation scan starts at the next launch
            for (int idx = c.tid; idx < nw * nw; idx += c.nthreads) {
                int r = idx / nw, q = idx % nw;
                Tg[r * QR_W + q] = Hs[r * QR_LD + q];
                Vg[r * QR_W + q] = Us[r * QR_LD + q];
            }
            QR_ACCOUNT(3);
            if (c.tid == 0) {
                st.ss_i = sc->ss_i; st.ss_its = sc->ss_its; st.ss_fresh = 0; st.passes++;
                if (rc == 1) { st.aed_stage = 1; st.aed_prog[0] = nw; st.aed_prog[1] = 0; st.aed_prog[2] = 0; st.aed_prog[3] = -1; }
                *stg = st;
            }
            return;
        }
        // ---- stage 1: deflation scan in time slices; stage 2: Hessenberg restore in time slices; then finish
        const cplx spike = H[(size_t)kw * ldh + kw - 1];
        if (st.aed_stage == 1) {
            if (c.tid < w1.nthreads) {
                int prog[4] = {st.aed_prog[0], st.aed_prog[1], st.aed_prog[2], st.aed_prog[3]};
                const int done1 = aed_deflation_scan(w1, Hs, Us, nw, spike, prog, bud.swaps, deadline);
                if (c.tid == 0) { sc->ss_rc = done1; sc->aed_ns = prog[0]; sc->ss_i = prog[1]; sc->ss_its = prog[2]; sc->nslots = prog[3]; }
            }
            CTA_SYNC();
            const int scan_done = sc->ss_rc, ns1 = sc->aed_ns;
            const bool need_restore = scan_done && (nw - ns1 > 0) && (ns1 > 1);
            if (scan_done && c.tid == 0) {                       // shifts offered to the next sweep: trailing
                const int nsh1 = (ns1 < QR_NS) ? ns1 : QR_NS;    // undeflated eigenvalues of the window
                for (int j = 0; j < nsh1; ++j) st.shifts[j] = Hs[(ns1 - nsh1 + j) * QR_LD + (ns1 - nsh1 + j)];
            }
            if (!scan_done || need_restore) {
                for (int idx = c.tid; idx < nw * nw; idx += c.nthreads) {
                    int r = idx / nw, q = idx % nw;
                    Tg[r * QR_W + q] = Hs[r * QR_LD + q];
                    Vg[r * QR_W + q] = Us[r * QR_LD + q];
                }
                QR_ACCOUNT(4);
                if (c.tid == 0) {
                    st.aed_prog[0] = ns1; st.aed_prog[1] = sc->ss_i; st.aed_prog[2] = sc->ss_its; st.aed_prog[3] = sc->nslots;
                    if (need_restore) { st.aed_stage = 2; st.aed_prog[1] = -1; }
                    st.passes++; *stg = st;
                }
                return;
            }
            CTA_SYNC();
        } else {
            if (c.tid < w1.nthreads) {
                int rj = st.aed_prog[1];
                const int done2 = aed_restore_hessenberg(w1, Hs, Us, nw, st.aed_prog[0], sc->hv, &rj, bud.restore, deadline);
                if (c.tid == 0) { sc->ss_rc = done2; sc->ss_i = rj; sc->aed_ns = st.aed_prog[0]; }
            }
            CTA_SYNC();
            if (!sc->ss_rc) {
                for (int idx = c.tid; idx < nw * nw; idx += c.nthreads) {
                    int r = idx / nw, q = idx % nw;
                    Tg[r * QR_W + q] = Hs[r * QR_LD + q];
                    Vg[r * QR_W + q] = Us[r * QR_LD + q];
                }
                QR_ACCOUNT(5);
                if (c.tid == 0) { st.aed_prog[1] = sc->ss_i; st.passes++; *stg = st; }
                return;
            }
        }
        const int ns = sc->aed_ns;
        const int nd = nw - ns;
        const int nsh = (ns < QR_NS) ? ns : QR_NS;
        if (nd > 0) {
            for (int idx = c.tid; idx < nw * nw; idx += c.nthreads) {
                int r = idx / nw, q = idx % nw;
                const bool below = (q < ns) ? (r > q + 1) : (r > q);        // exact zeros below the (quasi-)triangle
                H[(size_t)(kw + r) * ldh + (kw + q)] = below ? C(0, 0) : Hs[r * QR_LD + q];
                Ug[r * QR_W + q] = Us[r * QR_LD + q];
            }
        }
        QR_ACCOUNT(5);
        if (c.tid == 0) {
            st.passes++;
            if (nd > 0) {
                H[(size_t)kw * ldh + kw - 1] = (ns > 0) ? cmul(spike, cconj(Us[0])) : C(0, 0);
                emit_window_gemms(H, ldh, n, Zm, ldz, Ug, kw, nw, st.lo, st.hi, prob_rows, prob_cols_main, prob_z);
                st.aed_deflated += nd;
                st.hi = kw + ns - 1;          // the nd trailing eigenvalues are converged
                st.stall = 0;
            }
            const int m_new = st.hi - st.lo + 1;
            if (m_new <= QR_SMALL || ns == 0 || nd * 100 > 14 * nw) {
                st.phase = 0;                 // good harvest (or tiny rest): look again before sweeping
            } else {
                st.ns = nsh; st.nintro = 0; st.nbulge = 0; st.p = st.lo; st.phase = 1; st.sweeps++; st.shifts_ready = 1;
            }
            *stg = st;
        }
        return;
    }

    if (st.phase == 2) {
        // ---------------- one time slice of the small-block solve on [lo, hi]
        const int p2 = st.lo, m2 = st.hi - st.lo + 1;
        const int fresh = st.ss_fresh;
        for (int idx = c.tid; idx < m2 * m2; idx += c.nthreads) {
            int r = idx / m2, q = idx % m2;
            Hs[r * QR_LD + q] = H[(size_t)(p2 + r) * ldh + (p2 + q)];
            Us[r * QR_LD + q] = fresh ? C(r == q ? 1.0 : 0.0, 0.0) : Vg[r * QR_W + q];
        }
        CTA_SYNC();
        Cta w1 = c; w1.warp_only = 1; w1.nthreads = (c.nthreads < 32) ? c.nthreads : 32;
        if (c.tid < w1.nthreads) {
            int si1 = st.ss_i, sits1 = st.ss_its;
            const int rc1 = small_schur_slice(w1, Hs, Us, m2, &si1, &sits1, bud.schur, deadline, sc->rot_c, sc->rot_s);
            if (c.tid == 0) { sc->ss_rc = rc1; sc->ss_i = si1; sc->ss_its = sits1; }
        }
        CTA_SYNC();
        const int rc = sc->ss_rc, si = sc->ss_i, sits = sc->ss_its;
        for (int idx = c.tid; idx < m2 * m2; idx += c.nthreads) {
            int r = idx / m2, q = idx % m2;
            H[(size_t)(p2 + r) * ldh + (p2 + q)] = Hs[r * QR_LD + q];
            Vg[r * QR_W + q] = Us[r * QR_LD + q];
            if (rc != 0) Ug[r * QR_W + q] = Us[r * QR_LD + q];
        }
        QR_ACCOUNT(2);
        if (c.tid == 0) {
            st.ss_i = si; st.ss_its = sits; st.ss_fresh = 0;
            st.passes++;
            if (rc != 0) {
                // finished (or failed): apply the accumulated unitary to the off-diagonal panels and Z
                emit_window_gemms(H, ldh, n, Zm, ldz, Ug, p2, m2, st.lo, st.hi, prob_rows, prob_cols_main, prob_z);
                st.phase = 0;
                if (rc < 0) { st.done = 1; st.info = st.lo + si + 1; }
            }
            *stg = st;
        }
        return;
    }

    // ---------------- window [p, wend): chase the chain of bulges through it
    const int p = st.p;
    const int wend = (p + QR_W < st.hi + 1) ? p + QR_W : st.hi + 1;
    const int wl = wend - p;
    const bool at_bottom = (wend - 1 == st.hi);
    for (int idx = c.tid; idx < wl * wl; idx += c.nthreads) {
        int r = idx / wl, q = idx % wl;
        Hs[r * QR_LD + q] = H[(size_t)(p + r) * ldh + (p + q)];
        Us[r * QR_LD + q] = C(r == q ? 1.0 : 0.0, 0.0);
    }
    // closed-form schedule: slot b performs rotations at steps start[b] .. start[b]+nrot[b]-1, the
    // i-th of them on rows (q0[b]+i+1, q0[b]+i+2) (an introduction starts from the virtual column -1).
    if (c.tid == 0) {
        int slots = 0, tmax = 0;
        for (int b = 0; b < st.nbulge; ++b) {               // bulges already in flight (leading first)
            const int q = st.kpos[b] - p;
            const int fin = at_bottom ? (wl - 2) : (wl - 3 - 2 * b);
            int m = fin - q; if (m < 0) m = 0;
            sc->start[slots] = 0; sc->q0[slots] = q; sc->nrot[slots] = m; sc->shift_id[slots] = 0;
            if (m > tmax) tmax = m;
            ++slots;
        }
        if (p == st.lo && st.nbulge == 0) {                 // first window of the sweep: introduce the chain
            int j = 0;
            for (; st.nintro + j < st.ns && slots < QR_NS; ++j) {
                const int fin = at_bottom ? (wl - 2) : (wl - 3 - 2 * j);
                if (fin < 1) break;                           // no room for another bulge in this window
                sc->start[slots] = 2 * j; sc->q0[slots] = -1; sc->nrot[slots] = 1 + fin; sc->shift_id[slots] = st.nintro + j;
                if (2 * j + 1 + fin > tmax) tmax = 2 * j + 1 + fin;
                ++slots;
            }
        }
        sc->nslots = slots; sc->tmax = tmax;
    }
    CTA_SYNC();
    const int nslots = sc->nslots, tmax = sc->tmax;
    for (int t = 0; t < tmax; ++t) {
        // ---- rotation parameters (one thread per bulge slot)
        for (int b = c.tid; b < nslots; b += c.nthreads) {
            const int i = t - sc->start[b];
            const bool on = (i >= 0 && i < sc->nrot[b]);
            sc->act[b] = on ? 1 : 0;
            if (!on) continue;
            const int pos = sc->q0[b] + i, r1 = pos + 1;
            cplx a, bb;
            if (pos < 0) { a = csub(Hs[0], st.shifts[sc->shift_id[b]]); bb = Hs[1 * QR_LD + 0]; }
            else { a = Hs[r1 * QR_LD + pos]; bb = Hs[(r1 + 1) * QR_LD + pos]; }
            // (a, b) both at round-off level (the chain runs over an already converged spot): a rotation
            // built from noise would scramble converged rows -> identity.  A tiny b next to a
            // non-negligible a is kept: small bulges still carry the shifts.
            if (cabs1(a) + cabs1(bb) <= RCWA_EPS * (cabs1(Hs[r1 * QR_LD + r1]) + cabs1(Hs[(r1 + 1) * QR_LD + r1 + 1]))) { a = C(0, 0); bb = C(0, 0); }
            double cs; cplx sn, r;
            givens(a, bb, cs, sn, r);
            sc->cs[b] = cs; sc->sn[b] = sn; sc->rr1[b] = r1;
            if (pos >= 0) { Hs[r1 * QR_LD + pos] = r; Hs[(r1 + 1) * QR_LD + pos] = C(0, 0); }
        }
        CTA_SYNC();
        // ---- left: rows (r1, r1+1), columns r1 .. wl-1 (column r1-1 was set explicitly above)
        for (int idx = c.tid; idx < nslots * QR_W; idx += c.nthreads) {
            const int b = idx / QR_W, j = idx % QR_W;
            if (!sc->act[b]) continue;
            const int r1 = sc->rr1[b];
            if (j < r1 || j >= wl) continue;
            const double cs = sc->cs[b]; const cplx sn = sc->sn[b];
            cplx x = Hs[r1 * QR_LD + j], y = Hs[(r1 + 1) * QR_LD + j];
            Hs[r1 * QR_LD + j] = cadd(cscale(x, cs), cmul(sn, y));
            Hs[(r1 + 1) * QR_LD + j] = csub(cscale(y, cs), cmul(cconj(sn), x));
        }
        CTA_SYNC();
        // ---- right: columns (r1, r1+1), rows 0..min(r1+2, wl-1) of H and all rows of U
        for (int idx = c.tid; idx < nslots * 2 * QR_W; idx += c.nthreads) {
            const int b = idx / (2 * QR_W), rem = idx % (2 * QR_W);
            if (!sc->act[b]) continue;
            const int r1 = sc->rr1[b];
            const double cs = sc->cs[b]; const cplx sn = sc->sn[b];
            cplx* base;
            if (rem < QR_W) {
                const int imax = (r1 + 2 < wl - 1) ? r1 + 2 : wl - 1;
                if (rem > imax) continue;
                base = Hs + rem * QR_LD;
            } else {
                if (rem - QR_W >= wl) continue;
                base = Us + (rem - QR_W) * QR_LD;
            }
            cplx x = base[r1], y = base[r1 + 1];
            base[r1] = cadd(cscale(x, cs), cmul(y, cconj(sn)));
            base[r1 + 1] = csub(cscale(y, cs), cmul(x, sn));
        }
        CTA_SYNC();
    }

    // ---------------- band of the window unitary.  A bulge carries content only along its own travel range and the
    // bulges never overtake each other, so U is banded (lower bandwidth = number of bulges, upper ~ travel + 1):
    // about 30 % of a 64 x 64 U are EXACT zeros (never touched since the identity).  The update GEMMs skip the
    // 4-row k groups that are zero for a whole 8-column tile of U; the table is read off the actual zeros.
    for (int j = c.tid; j < QR_W; j += c.nthreads) {
        int lo_r = 255, hi_r = 0;
        if (j < wl)
            for (int r = 0; r < wl; ++r)
                if (!cis_zero(Us[r * QR_LD + j])) { if (r < lo_r) lo_r = r; hi_r = r; }
        sc->col_lo[j] = (unsigned char)lo_r; sc->col_hi[j] = (unsigned char)hi_r;
    }
    CTA_SYNC();
    for (int t = c.tid; t < 8; t += c.nthreads) {
        int lo_r = 255, hi_r = -1;
        for (int q = 0; q < 8; ++q) {
            const int j = 8 * t + q;
            if (j < wl && sc->col_lo[j] != 255) { if (sc->col_lo[j] < lo_r) lo_r = sc->col_lo[j]; if (sc->col_hi[j] > hi_r) hi_r = sc->col_hi[j]; }
        }
        sc->band_lo[t] = (unsigned char)((hi_r < 0) ? 0 : lo_r / 4);
        sc->band_hi[t] = (unsigned char)((hi_r < 0) ? 0 : hi_r / 4 + 1);
    }
    // ---------------- write back window and U, emit GEMM problems, advance state
    for (int idx = c.tid; idx < wl * wl; idx += c.nthreads) {
        int r = idx / wl, q = idx % wl;
        H[(size_t)(p + r) * ldh + (p + q)] = Hs[r * QR_LD + q];
        Ug[r * QR_W + q] = Us[r * QR_LD + q];
    }
    CTA_SYNC();
    if (c.tid == 0) {
        for (int t = 0; t < 8; ++t) if (8 * t < wl) { st.band_active += sc->band_hi[t] - sc->band_lo[t]; st.band_total += (wl + 3) / 4; }
        int nb_after = 0, introduced = 0;
        for (int b = 0; b < nslots; ++b) {
            if (sc->q0[b] < 0) ++introduced;
            if (at_bottom) continue;                          // every bulge ran off the bottom
            const int fin = sc->q0[b] + sc->nrot[b];          // column after its last rotation
            st.kpos[nb_after++] = p + fin;
        }
        // the pass that ends a sweep is followed by a deflation scan / AED window that may reach above this
        // window: its column update must be complete by then -> main stream; otherwise side stream
        emit_window_gemms(H, ldh, n, Zm, ldz, Ug, p, wl, st.lo, st.hi, prob_rows, (nb_after == 0) ? prob_cols_main : prob_cols, prob_z,
                          sc->band_lo, sc->band_hi);
        st.p_last = (nb_after == 0) ? -1 : p;
        st.nintro += introduced;
        if (p == st.lo && st.nintro < st.ns) st.ns = st.nintro;      // window could not take more: cap this sweep
        st.nbulge = nb_after;
        { const long long _t = QR_CLOCK(); st.cyc[1] += _t - tseg; st.cnt[1]++; if (_t - tseg > st.cyc_max[1]) st.cyc_max[1] = _t - tseg; }
        st.passes++;
        if (nb_after == 0) { st.phase = 0; st.shifts_ready = 0; }   // chain gone: next pass starts a new sweep
        else st.p = st.kpos[nb_after - 1];                    // next window starts at the trailing bulge
        *stg = st;
    }
}

// ------------------------------------------------------------------------------------------------
// Eigenvectors of the triangular T: diagonal-block solves for block rows [r0, r1).
//   for columns j >= r1 :  (T[I,I] - t_jj) x = R[:, j]   (R already in X[I, j])
//   for columns j in I  :  x_j = 1, (T[r0:j, r0:j] - t_jj) x = -T[r0:j, j]
// One (virtual) thread per column; T[I,I] is read through `Tb` (nbk x nbk, ld TV_NB).
DEV void trevc_col(const cplx* Tb, int nbk, int r0, int j, cplx tjj, double smin, cplx* X, int ldx, const cplx* T, int ldt) {
    int top;                     // number of rows of the block to solve
    if (j >= r0 + nbk) top = nbk;
    else {
        top = j - r0;
        for (int r = 0; r < top; ++r) X[(size_t)(r0 + r) * ldx + j] = cneg(T[(size_t)(r0 + r) * ldt + j]);
        X[(size_t)j * ldx + j] = C(1, 0);
        for (int r = top + 1; r < nbk; ++r) X[(size_t)(r0 + r) * ldx + j] = C(0, 0);
    }
    for (int r = top - 1; r >= 0; --r) {
        cplx acc = X[(size_t)(r0 + r) * ldx + j];
        for (int s = r + 1; s < top; ++s) acc = csub(acc, cmul(Tb[r * TV_NB + s], X[(size_t)(r0 + s) * ldx + j]));
        cplx d = csub(Tb[r * TV_NB + r], tjj);
        if (cabs1(d) < smin) d = C(smin, 0);
        X[(size_t)(r0 + r) * ldx + j] = cdiv(acc, d);
    }
}

#ifdef RCWA_EMU
// =============================================================================== CPU emulation
static void emu_gemm(const ZGemmProblem& g, int opa) {
    if (g.M <= 0 || g.N <= 0) return;
    std::vector<cplx> out((size_t)g.M * g.N);
    for (int i = 0; i < g.M; ++i)
        for (int j = 0; j < g.N; ++j) {
            cplx acc = C(0, 0);
            for (int k = 0; k < g.K; ++k) {
                cplx a = (opa == 0) ? g.A[(size_t)i * g.lda + k] : cconj(g.A[(size_t)k * g.lda + i]);
                acc = cfma(a, g.B[(size_t)k * g.ldb + j], acc);
            }
            out[(size_t)i * g.N + j] = acc;
        }
    for (int i = 0; i < g.M; ++i) for (int j = 0; j < g.N; ++j) g.C[(size_t)i * g.ldc + j] = out[(size_t)i * g.N + j];
}

// H (upper Hessenberg, n x n) -> T in place, Z <- Z U.  Returns info; stats[0..2] = sweeps, passes, done.
// split != 0: every iteration as the two launches of the large-batch device path (QR_MODE_WIN, then QR_MODE_SMALL);
// aed_w: AED window (0 = the compile-time maximum)
extern "C" int emu_qr_opts(cplx* H, cplx* Z, int n, int max_passes, int* stats, int split, int aed_w);
extern "C" int emu_qr(cplx* H, cplx* Z, int n, int max_passes, int* stats) { return emu_qr_opts(H, Z, n, max_passes, stats, 0, 0); }
extern "C" int emu_qr_opts(cplx* H, cplx* Z, int n, int max_passes, int* stats, int split, int aed_w) {
    std::vector<char> smem(qr_pass_smem_bytes(n));
    std::vector<cplx> U((size_t)QR_W * QR_W), Tgbuf((size_t)QR_W * QR_W), Vgbuf((size_t)QR_W * QR_W);
    // the QR phase keeps H current only inside active blocks; the caller's Z holds the Hessenberg
    // transformation Z0 on entry, so the input matrix is A0 = Z0 H0 Z0^H
    std::vector<cplx> H0(H, H + (size_t)n * n), Z0(Z, Z + (size_t)n * n), A0((size_t)n * n);
    {
        std::vector<cplx> t0((size_t)n * n), Z0h((size_t)n * n);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Z0h[(size_t)i * n + j] = cconj(Z0[(size_t)j * n + i]);
        ZGemmProblem g; g.flags = 0; g.M = n; g.N = n; g.K = n; g.lda = n; g.ldb = n; g.ldc = n;
        g.A = Z0.data(); g.B = H0.data(); g.C = t0.data(); emu_gemm(g, 0);
        g.A = t0.data(); g.B = Z0h.data(); g.C = A0.data(); emu_gemm(g, 0);
    }
    QrState st; memset(&st, 0, sizeof(st));
    st.lo = 0; st.hi = n - 1; st.hi_prev = n - 1; st.p_last = -1;
    Cta c; c.tid = 0; c.nthreads = 1; c.bid = 0; c.smem = smem.data(); c.warp_only = 0;
    ZGemmProblem pr, pcm, pc, pz;
    int it = 0;
    for (; it < max_passes && !st.done; ++it) {
        QrBudget bud; bud.schur = 160; bud.swaps = 100; bud.restore = 12; bud.cycles = 0; bud.aed_w = aed_w;     // count budgets exercise the slicing on the CPU
        if (!split) qr_pass_body(c, H, n, n, Z, n, &st, U.data(), Vgbuf.data(), Tgbuf.data(), &pr, &pcm, &pc, &pz, bud);
        else {
            qr_pass_body(c, H, n, n, Z, n, &st, U.data(), Vgbuf.data(), Tgbuf.data(), &pr, &pcm, &pc, &pz, bud, QR_MODE_WIN);
            qr_pass_body(c, H, n, n, Z, n, &st, U.data(), Vgbuf.data(), Tgbuf.data(), &pr, &pcm, &pc, &pz, bud, QR_MODE_SMALL);
        }
        emu_gemm(pr, 2); emu_gemm(pcm, 0); emu_gemm(pc, 0); emu_gemm(pz, 0);
    }
    {   // T = Z^H A0 Z (as the device path does); strictly lower part set to exact zero
        std::vector<cplx> tmp((size_t)n * n);
        ZGemmProblem g; g.flags = 0; g.M = n; g.N = n; g.K = n; g.lda = n; g.ldb = n; g.ldc = n;
        g.A = A0.data(); g.B = Z; g.C = tmp.data(); emu_gemm(g, 0);
        g.A = Z; g.B = tmp.data(); g.C = H; emu_gemm(g, 2);
        for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) H[(size_t)i * n + j] = C(0, 0);
    }
    stats[0] = st.sweeps; stats[1] = st.passes; stats[2] = st.done; stats[3] = st.small_solves;
    stats[4] = st.aeds; stats[5] = st.aed_deflated;
    stats[6] = st.cnt[1];            // bulge-chain windows (each one costs three update GEMMs on the device)
    stats[7] = (int)(100.0 * (double)st.band_active / (double)(st.band_total > 0 ? st.band_total : 1));
    return st.done ? st.info : -1;
}

extern "C" int emu_tiny_eigs(cplx* T, int m, cplx* w) {
    std::vector<cplx> buf((size_t)QR_NS * (QR_NS + 1));
    for (int r = 0; r < m; ++r) for (int q = 0; q < m; ++q) buf[r * (QR_NS + 1) + q] = T[r * m + q];
    return tiny_hqr_eigs(0, 1, buf.data(), QR_NS + 1, m, w);
}

// eigenvectors of upper-triangular T (n x n): X (n x n) unit upper triangular, column j = eigenvector j
extern "C" int emu_trevc(const cplx* T, int n, cplx* X) {
    memset(X, 0, sizeof(cplx) * (size_t)n * n);
    double tnorm = 0.0;
    for (int j = 0; j < n; ++j) tnorm = fmax(tnorm, cabs1(T[(size_t)j * n + j]));
    const int nblk = (n + TV_NB - 1) / TV_NB;
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int r0 = kb * TV_NB, nbk = (n - r0 < TV_NB) ? n - r0 : TV_NB, r1 = r0 + nbk;
        if (r1 < n) {
            ZGemmProblem g; g.flags = 0; g.A = T + (size_t)r0 * n + r1; g.lda = n; g.B = X + (size_t)r1 * n + r1; g.ldb = n;
            g.C = X + (size_t)r0 * n + r1; g.ldc = n; g.M = nbk; g.N = n - r1; g.K = n - r1;
            emu_gemm(g, 0);
            for (int i = 0; i < nbk; ++i) for (int j = r1; j < n; ++j) X[(size_t)(r0 + i) * n + j] = cneg(X[(size_t)(r0 + i) * n + j]);
        }
        cplx Tb[TV_NB * TV_NB];
        for (int r = 0; r < nbk; ++r) for (int s = 0; s < nbk; ++s) Tb[r * TV_NB + s] = T[(size_t)(r0 + r) * n + r0 + s];
        for (int j = r0; j < n; ++j) {
            cplx tjj = T[(size_t)j * n + j];
            double smin = fmax(RCWA_EPS * cabs1(tjj), RCWA_SAFMIN / RCWA_EPS);
            smin = fmax(smin, RCWA_EPS * tnorm * 1e-3);
            trevc_col(Tb, nbk, r0, j, tjj, smin, X, n, T, n);
        }
    }
    return 0;
}
#else
// =============================================================================== device code
namespace {

// ---------------------------------------------------------------- phase 2: QR passes
__global__ void __launch_bounds__(512, 1)
qr_pass_kernel(cplx* H, long long hstride, int ldh, int n, cplx* Z, long long zstride, int ldz, QrState* states,
               cplx* U, cplx* Vg, cplx* Tg, ZGemmProblem* prows, ZGemmProblem* pcols_main, ZGemmProblem* pcolsz, QrBudget bud, int mode) {
    extern __shared__ __align__(16) char smem_raw[];
    const int b = blockIdx.x;
    Cta c = make_cta(b, smem_raw);
    qr_pass_body(c, H + (size_t)b * hstride, ldh, n, Z + (size_t)b * zstride, ldz, states + b,
                 U + (size_t)b * QR_W * QR_W, Vg + (size_t)b * QR_W * QR_W, Tg + (size_t)b * QR_W * QR_W,
                 prows + b, pcols_main + b, pcolsz + 2 * b, pcolsz + 2 * b + 1, bud, mode);
}

__global__ void qr_init_kernel(QrState* states, int n, int nb) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    QrState st;
    memset(&st, 0, sizeof(st));
    st.lo = 0; st.hi = n - 1; st.hi_prev = n - 1; st.p_last = -1;
    if (n < 2) st.done = 1;
    states[b] = st;
}

// counts unfinished matrices into *flag_dev
__global__ void qr_count_kernel(const QrState* states, int nb, int* flag_dev) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) local += states[b].done ? 0 : 1;
    atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) *flag_dev = cnt;
}

__global__ void qr_stats_kernel(const QrState* states, int nb, int* out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    out[4 * b + 0] = states[b].sweeps; out[4 * b + 1] = states[b].passes;
    out[4 * b + 2] = states[b].aeds; out[4 * b + 3] = states[b].done ? states[b].info : -1;
}

__global__ void qr_profile_kernel(const QrState* states, int nb, long long* out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    for (int k = 0; k < 6; ++k) { out[18 * b + 3 * k] = states[b].cnt[k]; out[18 * b + 3 * k + 1] = states[b].cyc[k]; out[18 * b + 3 * k + 2] = states[b].cyc_max[k]; }
}

__global__ void qr_finish_kernel(const QrState* states, int nb, int* info) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    info[b] = states[b].done ? states[b].info : (states[b].hi + 1);
}

// ---------------------------------------------------------------- phase 3: eigenvectors
__global__ void diag_extract_kernel(const cplx* T, long long tstride, int ldt, int n, cplx* w) {
    int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (j < n) w[(size_t)b * n + j] = T[(size_t)b * tstride + (size_t)j * ldt + j];
}

// max_j |t_jj|_1 per matrix -> tnorm[b]
__global__ void tnorm_kernel(const cplx* w, int n, double* tnorm) {
    __shared__ double red[40];
    Cta c = make_cta(blockIdx.x, nullptr);
    double m = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) m = fmax(m, cabs1(w[(size_t)blockIdx.x * n + j]));
    m = cta_max(c, m, red);
    if (threadIdx.x == 0) tnorm[blockIdx.x] = m;
}

// grid (ceil((n - r0)/128), B): columns j >= r0
__global__ void __launch_bounds__(128)
trevc_block_kernel(const cplx* T, long long tstride, int ldt, int n, int r0, int nbk, const cplx* w, const double* tnorm,
                   cplx* X, long long xstride, int ldx) {
    __shared__ cplx Tb[TV_NB * TV_NB];
    const int b = blockIdx.y;
    const cplx* Tm = T + (size_t)b * tstride;
    for (int i = threadIdx.x; i < nbk * nbk; i += blockDim.x) Tb[(i / nbk) * TV_NB + (i % nbk)] = Tm[(size_t)(r0 + i / nbk) * ldt + r0 + (i % nbk)];
    __syncthreads();
    const int j = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const cplx tjj = w[(size_t)b * n + j];
    double smin = fmax(RCWA_EPS * cabs1(tjj), RCWA_SAFMIN / RCWA_EPS);
    smin = fmax(smin, RCWA_EPS * tnorm[b] * 1e-3);
    trevc_col(Tb, nbk, r0, j, tjj, smin, X + (size_t)b * xstride, ldx, Tm, ldt);
}

// column 2-norms: grid (ceil(n/32), B), block (32, 8): coalesced row segments
__global__ void colnorm_kernel(const cplx* V, long long vstride, int ldv, int n, double* nrm) {
    __shared__ double part[8][33];
    const int j = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
    double s = 0.0;
    if (j < n) for (int i = threadIdx.y; i < n; i += 8) s += cabs2(V[(size_t)b * vstride + (size_t)i * ldv + j]);
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && j < n) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += part[q][threadIdx.x];
        nrm[(size_t)b * n + j] = sqrt(t);
    }
}
__global__ void colscale_kernel(cplx* V, long long vstride, int ldv, int n, const double* nrm) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= n) return;
    const double s = nrm[(size_t)b * n + j];
    cplx* p = V + (size_t)b * vstride + (size_t)i * ldv + j;
    if (s > 0.0) *p = cscale(*p, 1.0 / s);
}

inline size_t al(size_t x) { return (x + 255) & ~size_t(255); }

struct EigWs {
    cplx *Z, *X, *U, *Vg, *Tg; char* hess; QrState* states; ZGemmProblem *prows, *pcols_main, *pcolsz, *gs; int* flag; double *tnorm, *nrm;
    size_t total;
};
EigWs carve(char* base, int n, int nb) {
    EigWs w; size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += al(bytes); return p; };
    w.Z = (cplx*)take(sizeof(cplx) * (size_t)n * n * nb);
    w.X = (cplx*)take(sizeof(cplx) * (size_t)n * n * nb);
    w.U = (cplx*)take(sizeof(cplx) * (size_t)QR_W * QR_W * nb * 2);      // double-buffered window unitaries
    w.Vg = (cplx*)take(sizeof(cplx) * (size_t)QR_W * QR_W * nb);
    w.Tg = (cplx*)take(sizeof(cplx) * (size_t)QR_W * QR_W * nb);
    w.hess = take(rcwa::hessenberg_workspace_bytes(n, nb / 2) + rcwa::hessenberg_workspace_bytes(n, nb - nb / 2) + 4096);   // two half batches
    w.states = (QrState*)take(sizeof(QrState) * (size_t)nb);
    w.prows = (ZGemmProblem*)take(sizeof(ZGemmProblem) * (size_t)nb);
    w.pcols_main = (ZGemmProblem*)take(sizeof(ZGemmProblem) * (size_t)nb);
    w.pcolsz = (ZGemmProblem*)take(sizeof(ZGemmProblem) * (size_t)nb * 2 * 2);   // double-buffered
    w.gs = (ZGemmProblem*)take(sizeof(ZGemmProblem) * (size_t)nb * 4);
    w.flag = (int*)take(256);
    w.tnorm = (double*)take(sizeof(double) * (size_t)nb);
    w.nrm = (double*)take(sizeof(double) * (size_t)nb * n);
    w.total = off;
    return w;
}

}  // namespace

// Streams / events of the QR phase: one set per host thread and device, created on first use, reused, never destroyed
// (a few driver objects per thread; they die with the process).  Thread-local => re-entrant across host threads.
struct EigStreams {
    bool ready;
    cudaStream_t sa[QR_MAXG], sb[QR_MAXG];
    cudaEvent_t ev_fork, ev_join[QR_MAXG], ev_pass[QR_MAXG][2], ev_side[QR_MAXG][2], ev[QR_MAXG][2];
    cudaEvent_t cap_pass[QR_MAXG][2], cap_side[QR_MAXG][2];      // the same roles inside a stream capture (CUDA graph of the QR loop)
};
EigStreams* eig_streams() {
    static thread_local EigStreams cache[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    EigStreams& r = cache[dev];
    if (r.ready) return &r;
    int prio_lo = 0, prio_hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) return nullptr;
    bool ok = cudaEventCreateWithFlags(&r.ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int g = 0; g < QR_MAXG && ok; ++g) {
        ok = ok && cudaStreamCreateWithPriority(&r.sa[g], cudaStreamNonBlocking, prio_hi) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&r.sb[g], cudaStreamNonBlocking, prio_lo) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&r.ev_join[g], cudaEventDisableTiming) == cudaSuccess;
        for (int q = 0; q < 2 && ok; ++q) {
            ok = ok && cudaEventCreateWithFlags(&r.ev_pass[g][q], cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&r.ev_side[g][q], cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&r.ev[g][q], cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&r.cap_pass[g][q], cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&r.cap_side[g][q], cudaEventDisableTiming) == cudaSuccess;
        }
    }
    if (!ok) return nullptr;
    r.ready = true;
    return &r;
}

namespace rcwa {

size_t eig_workspace_bytes(int n, int nb) { return carve(nullptr, n, nb).total; }

#define EK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)

// The Hessenberg phase alternates an HBM-bound column phase (streaming mat-vec + a latency-bound per-column kernel:
// tensor pipes idle) with a tensor-bound block-update phase (HBM idle).  Two half batches on two streams, the second
// started when the first has finished its first column phase, run in anti-phase and use both resources at once.
static cudaError_t hessenberg_phase(cplx* A, int n, int nb, const EigWs& ws, cudaStream_t st) {
    if (!gemm_get_tuning(11) || nb < 16) return rcwa::hessenberg_blocked(A, n, nb, ws.Z, ws.hess, st);      // hess.cu
    const int nb0 = nb / 2, nb1 = nb - nb0;
    const long long ms = (long long)n * n;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t fork = nullptr, first = nullptr, j0 = nullptr, j1 = nullptr;
    cudaError_t e;
#define HP(expr) do { e = (expr); if (e != cudaSuccess) return e; } while (0)
    HP(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
    HP(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    HP(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    HP(cudaEventCreateWithFlags(&first, cudaEventDisableTiming));
    HP(cudaEventCreateWithFlags(&j0, cudaEventDisableTiming));
    HP(cudaEventCreateWithFlags(&j1, cudaEventDisableTiming));
    HP(cudaEventRecord(fork, st));
    HP(cudaStreamWaitEvent(s0, fork, 0));
    HP(cudaStreamWaitEvent(s1, fork, 0));
    char* w1 = ws.hess + al(rcwa::hessenberg_workspace_bytes(n, nb0));
    HP(rcwa::hessenberg_blocked(A, n, nb0, ws.Z, ws.hess, s0, first));
    HP(cudaStreamWaitEvent(s1, first, 0));
    HP(rcwa::hessenberg_blocked(A + (size_t)nb0 * ms, n, nb1, ws.Z + (size_t)nb0 * ms, w1, s1));
    HP(cudaEventRecord(j0, s0));
    HP(cudaEventRecord(j1, s1));
    HP(cudaStreamWaitEvent(st, j0, 0));
    HP(cudaStreamWaitEvent(st, j1, 0));
#undef HP
    cudaEventDestroy(fork); cudaEventDestroy(first); cudaEventDestroy(j0); cudaEventDestroy(j1);
    cudaStreamDestroy(s0); cudaStreamDestroy(s1);
    return cudaSuccess;
}

// diagnostics of the last eig() run that used this workspace: per matrix {sweeps, passes, AED windows, info}
cudaError_t eig_stats(const char* wsb, int n, int nb, int* out, cudaStream_t st) {
    EigWs ws = carve(const_cast<char*>(wsb), n, nb);
    qr_stats_kernel<<<(nb + 127) / 128, 128, 0, st>>>(ws.states, nb, out);
    return cudaGetLastError();
}

// per matrix 6 x {count, SM cycles} of the QR pass segments (see QrState::cyc)
cudaError_t eig_profile(const char* wsb, int n, int nb, long long* out, cudaStream_t st) {
    EigWs ws = carve(const_cast<char*>(wsb), n, nb);
    qr_profile_kernel<<<(nb + 127) / 128, 128, 0, st>>>(ws.states, nb, out);
    return cudaGetLastError();
}

cudaError_t eig_matvec_probe(const cplx* A, int n, int nb, int j, char* wsb, size_t ws_bytes, cudaStream_t st) {
    EigWs ws = carve(wsb, n, nb);
    if (ws_bytes < ws.total) return cudaErrorInvalidValue;
    return rcwa::hessenberg_matvec_probe(A, n, nb, j, ws.hess, st);
}

// A -> H (upper Hessenberg, in place), Zout = accumulated reflectors (A_in = Z H Z^H)
cudaError_t hessenberg(cplx* A, int n, int nb, cplx* Zout, char* wsb, size_t ws_bytes, cudaStream_t st) {
    EigWs ws = carve(wsb, n, nb);
    if (ws_bytes < ws.total) return cudaErrorInvalidValue;
    EK(hessenberg_phase(A, n, nb, ws, st));
    return cudaMemcpyAsync(Zout, ws.Z, sizeof(cplx) * (size_t)n * n * nb, cudaMemcpyDeviceToDevice, st);
}

cudaError_t eig(cplx* A, int n, int nb, cplx* wout, cplx* V, char* wsb, size_t ws_bytes, int* info, volatile int* host_flag, cudaStream_t st, int phases) {
    EigWs ws = carve(wsb, n, nb);
    if (ws_bytes < ws.total) return cudaErrorInvalidValue;
    const long long ms = (long long)n * n;
    const cplx one = C(1, 0), zero = C(0, 0);

    // ---------------- phase 1: keep a copy A0 of the input for the final T = Z^H A0 Z; Hessenberg, Z accumulated
    if (phases & 1) {
        EK(cudaMemcpyAsync(ws.X, A, sizeof(cplx) * (size_t)n * n * nb, cudaMemcpyDeviceToDevice, st));
        EK(hessenberg_phase(A, n, nb, ws, st));
    }
    if (!(phases & 2)) return cudaGetLastError();

    // ---------------- phase 2: QR passes (host enqueues, polls the pinned flag every `poll` passes)
    qr_init_kernel<<<(nb + 127) / 128, 128, 0, st>>>(ws.states, n, nb);
    // Optional (tuning key 4): let every pass CTA claim the whole shared memory of its SM so that no GEMM CTA
    // becomes co-resident.  Measured on B200 (profiles/): no difference -- the pass is bound by its own
    // dependent fp64 chain, not by sharing the fp64 pipe with DMMA warps -- so it is off by default.
    size_t smem = qr_pass_smem_bytes(n);
    if (gemm_get_tuning(4) && smem < 227 * 1024) smem = 227 * 1024;
    EK(cudaFuncSetAttribute(qr_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QrBudget bud;
    bud.schur = gemm_get_tuning(5) > 0 ? gemm_get_tuning(5) : QR_BUDGET_SCHUR;
    bud.swaps = gemm_get_tuning(6) > 0 ? gemm_get_tuning(6) : QR_BUDGET_SWAPS;
    bud.restore = gemm_get_tuning(7) > 0 ? gemm_get_tuning(7) : QR_BUDGET_RESTORE;
    // AED window: the Schur factorisation of the window costs O(w^3) serial work per AED, and on small matrices that is what
    // the pass kernel spends its time on.  Measured (profiles/r2_summary.md): 512 matrices of n = 481 -- eig 872 ms at w = 48,
    // 790 at 40, 700 at 32, 662 at 28, 651 at 24 (sweeps per matrix 26.6 -> 33.7); 64 matrices of n = 1922 -- step 4.14 s at 48,
    // 4.00 at 40, 3.91 at 32.
    bud.aed_w = gemm_get_tuning(15) > 0 ? gemm_get_tuning(15) : ((n <= QR_AED_SMALL_N) ? QR_AED_W_SMALL : QR_AED_W_LARGE);
    if (bud.aed_w < 8) bud.aed_w = 8;
    {
        int dev = 0, khz = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
        const int us = gemm_get_tuning(8) != 0 ? gemm_get_tuning(8) : QR_BUDGET_US;
        bud.cycles = (us > 0) ? (long long)us * (khz > 0 ? khz : 1900000) / 1000 : 0;         // us < 0: count budgets only
    }
    const long long max_passes = 80LL * n + 4000;       // generous: ~(n/NS) sweeps x (n/(W-2NS)) windows x iterations
    const int poll = 64;
    // in-place window updates need one tile across the K-side dimension (zgemm.cu): rows 64x128 or 64x64,
    // columns/Z 128x64 or 64x64; the 128-thread 64x64 tile (two CTAs per SM) is the measured best at K = 64
    const int m3 = gemm_get_tuning(0) ? GEMM_M3 : 0;
    const int cfg_rows = ((gemm_get_tuning(1) == GEMM_TILE_64x128) ? GEMM_TILE_64x128 : GEMM_TILE_64x64) | m3;
    const int cfg_cz = ((gemm_get_tuning(2) == GEMM_TILE_128x64) ? GEMM_TILE_128x64 : GEMM_TILE_64x64) | m3;
    const int band = (m3 && gemm_get_tuning(10)) ? GEMM_BAND : 0;
    const int cfg_rows_b = ((cfg_rows & 7) == GEMM_TILE_64x64) ? (cfg_rows | band) : cfg_rows;
    const int cfg_cz_b = ((cfg_cz & 7) == GEMM_TILE_64x64) ? (cfg_cz | band) : cfg_cz;
    const int max_tiles_rows = gemm_tiles(cfg_rows, QR_W, n);
    const int max_tiles_cz = gemm_tiles(cfg_cz, n, QR_W);
    // Two streams: the main stream carries the serial chain  pass -> row-panel GEMM -> next pass ; the
    // column-panel and Z updates of a chase pass (2/3 of the flops) run on a side stream, overlapping the
    // next pass.  Window unitaries and side-stream descriptors are double-buffered; pass k+2 waits for the
    // side GEMMs of pass k.  The host polls convergence with a lag of one group: the count of group g is
    // copied to pinned memory asynchronously and examined after group g+1 has been enqueued.
    // Priorities: the latency-bound chain (pass kernel + the GEMM the next pass needs) runs on an internal
    // HIGH-priority stream forked from the caller's stream; the bulk column/Z GEMMs run on a low-priority side
    // stream, so that a pass kernel's CTAs are dispatched as soon as SMs free up instead of queueing behind
    // thousands of GEMM tiles.  Both are joined back into the caller's stream before the eigenvector phase.
    // The batch is split into G groups with their own stream pairs: the chain of one group (pass -> row GEMM ->
    // next pass) leaves the tensor pipes idle while its pass kernel runs and most SMs idle while its row GEMM
    // runs; the other groups' chains fill those gaps.
    // Default: two groups; more when a group's pass launch (one CTA per matrix, one CTA per SM) would not fit in one wave.
    // Many small matrices (the blocks of a symmetry-reduced layer: more matrices than two waves of SMs) make the QR phase
    // bound by the SM time of the pass kernel, 2/3 of it in the one-warp small dense solves: run those as a second launch
    // with two CTAs per SM (tuning key 13: 0 = automatic, 1 = one launch, 2 = two launches).
    const size_t smem_small = qr_pass_smem_bytes(n, QR_MODE_SMALL);
    const bool split = (gemm_get_tuning(13) == 2) || (gemm_get_tuning(13) == 0 && nb > 2 * QR_SMS);
    int G = gemm_get_tuning(9) > 0 ? gemm_get_tuning(9) : ((nb > 2 * QR_SMS) ? (nb + QR_SMS - 1) / QR_SMS : 2);
    if (gemm_get_tuning(9) <= 0 && G > 4) G = 4;          // more directly enqueued groups are host-bound (measured, see the graph note below)
    if (G > QR_MAXG) G = QR_MAXG;
    while (G > 1 && nb < 8 * G) --G;
    // Internal streams and events are created ONCE per host thread and device and reused by later calls (thread-local
    // cache below): no per-call cudaStreamCreate / Destroy, nothing shared between host threads.  Whatever happens below,
    // the Join guard makes the caller's stream wait for everything that was forked before this function returns.
    EigStreams* res = eig_streams();
    if (!res) return cudaErrorMemoryAllocation;
    cudaStream_t user_st = st;
    cudaStream_t* sa = res->sa; cudaStream_t* sb = res->sb;
    cudaEvent_t ev_fork = res->ev_fork;
    cudaEvent_t (*ev_pass)[2] = res->ev_pass; cudaEvent_t (*ev_side)[2] = res->ev_side; cudaEvent_t (*ev)[2] = res->ev;
    int gb0[QR_MAXG + 1];
    for (int g = 0; g <= G; ++g) gb0[g] = (int)((long long)nb * g / G);
    EK(cudaEventRecord(ev_fork, user_st));
    int* hf = const_cast<int*>(host_flag);
    struct Join {
        EigStreams* r; cudaStream_t user; int G; bool side_used[QR_MAXG][2];
        void run() {
            for (int g = 0; g < G; ++g) {
                for (int q = 0; q < 2; ++q) if (side_used[g][q]) cudaStreamWaitEvent(r->sa[g], r->ev_side[g][q], 0);
                cudaEventRecord(r->ev_join[g], r->sa[g]);
                cudaStreamWaitEvent(user, r->ev_join[g], 0);
            }
            G = 0;
        }
        ~Join() { run(); }      // also on every early error return: never leave forked work un-joined behind the caller's stream
    } join_guard = {res, user_st, G, {}};
    for (int g = 0; g < G; ++g) {
        EK(cudaStreamWaitEvent(sa[g], ev_fork, 0));
        for (int q = 0; q < 2; ++q)
            if (hf) hf[g * 2 + q] = gb0[g + 1] - gb0[g];
    }
    long long group = 0;
    bool fin[QR_MAXG] = {};
    int nfin = 0;
    const size_t ustride = (size_t)QR_W * QR_W * nb;
    const size_t w2 = (size_t)QR_W * QR_W;
    // One iteration of group g: pass launch(es), the row-panel GEMM and the main-stream column GEMM on sm, the bulk column / Z
    // GEMMs on ss.  wait2 / wait1: whether the side GEMMs of iterations it-2 / it-1 exist and must be waited for (evs = the
    // side events, evp = the pass events: the cached pair of the group, or the capture pair when this is being recorded
    // into a CUDA graph).
    auto enqueue = [&](int g, int buf, bool wait2, bool wait1, cudaEvent_t* evp, cudaEvent_t* evs) -> cudaError_t {
        const int b0 = gb0[g], nbg = gb0[g + 1] - gb0[g];
        cudaStream_t sm = sa[g], ss = sb[g];
        ZGemmProblem* pcz = ws.pcolsz + (size_t)buf * 2 * nb + 2 * (size_t)b0;
        if (wait2) EK(cudaStreamWaitEvent(sm, evs[buf], 0));                   // U[buf] / descriptors[buf] are free again
        if (!split) {
            qr_pass_kernel<<<nbg, 512, smem, sm>>>(A + (size_t)b0 * ms, ms, n, n, ws.Z + (size_t)b0 * ms, ms, n, ws.states + b0,
                                                   ws.U + buf * ustride + b0 * w2, ws.Vg + b0 * w2, ws.Tg + b0 * w2,
                                                   ws.prows + b0, ws.pcols_main + b0, pcz, bud, QR_MODE_ALL);
        } else {
            qr_pass_kernel<<<nbg, 512, smem, sm>>>(A + (size_t)b0 * ms, ms, n, n, ws.Z + (size_t)b0 * ms, ms, n, ws.states + b0,
                                                   ws.U + buf * ustride + b0 * w2, ws.Vg + b0 * w2, ws.Tg + b0 * w2,
                                                   ws.prows + b0, ws.pcols_main + b0, pcz, bud, QR_MODE_WIN);
            qr_pass_kernel<<<nbg, 256, smem_small, sm>>>(A + (size_t)b0 * ms, ms, n, n, ws.Z + (size_t)b0 * ms, ms, n, ws.states + b0,
                                                         ws.U + buf * ustride + b0 * w2, ws.Vg + b0 * w2, ws.Tg + b0 * w2,
                                                         ws.prows + b0, ws.pcols_main + b0, pcz, bud, QR_MODE_SMALL);
        }
        EK(cudaEventRecord(evp[buf], sm));
        EK(zgemm_grouped(cfg_rows_b, OP_H, OP_N, ws.prows + b0, nbg, max_tiles_rows, one, zero, sm));
        // A main-stream column update (last window of a sweep, AED, small block) overlaps the columns of the
        // previous window's side-stream column update and must be applied AFTER it: wait for the previous
        // iteration's side GEMMs.  (Without this the order was only a matter of timing -- the low-priority side
        // GEMM normally finishes long before -- and a second group's kernels delaying it corrupted results.)
        if (wait1) EK(cudaStreamWaitEvent(sm, evs[buf ^ 1], 0));
        EK(zgemm_grouped(cfg_cz_b, OP_N, OP_N, ws.pcols_main + b0, nbg, max_tiles_cz, one, zero, sm));
        EK(cudaStreamWaitEvent(ss, evp[buf], 0));
        EK(zgemm_grouped(cfg_cz_b, OP_N, OP_N, pcz, 2 * nbg, max_tiles_cz, one, zero, ss));
        EK(cudaEventRecord(evs[buf], ss));
        return cudaSuccess;
    };
    auto poll_groups = [&]() -> cudaError_t {
        const int slot = (int)(group & 1);
        for (int g = 0; g < G; ++g) {
            if (fin[g]) continue;
            if (group >= 1) {       // examine the previous poll's count (its copy was enqueued one poll period ago)
                EK(cudaEventSynchronize(ev[g][slot ^ 1]));
                if (hf[g * 2 + (slot ^ 1)] == 0) { fin[g] = true; ++nfin; continue; }
            }
            qr_count_kernel<<<1, 128, 0, sa[g]>>>(ws.states + gb0[g], gb0[g + 1] - gb0[g], ws.flag + g * 2 + slot);
            EK(cudaMemcpyAsync(hf + g * 2 + slot, ws.flag + g * 2 + slot, sizeof(int), cudaMemcpyDeviceToHost, sa[g]));
            EK(cudaEventRecord(ev[g][slot], sa[g]));
        }
        ++group;
        return cudaSuccess;
    };
    // Optional (tuning key 14 = 2): record QR_GRAPH_ITERS iterations of a group ONCE as a CUDA graph (after two directly
    // enqueued iterations, which also take care of one-time kernel attributes) and replay it.  Inside the graph the side
    // GEMMs overlap the following passes exactly as in the direct loop; replays of one group serialise on its main stream.
    // Measured on 512 matrices of n = 481 (profiles/r2_summary.md): the direct loop costs the host ~60 us of API time per
    // group and iteration, so 6 or 8 groups are host-bound (988 / 1148 ms against 872 ms with 4); graphs remove that (885
    // / 877 ms) -- and show that 4 directly enqueued groups already sit on the device-side bound (pass SM time + GEMM time,
    // which cannot share an SM's shared memory).  Off by default.  Any failure while recording falls back to the direct loop.
    long long it = 0;
    bool graphs = hf && gemm_get_tuning(14) == 2;
    cudaGraphExec_t gexec[QR_MAXG] = {};
    struct GraphGuard { cudaGraphExec_t* e; ~GraphGuard() { for (int g = 0; g < QR_MAXG; ++g) if (e[g]) cudaGraphExecDestroy(e[g]); } } graph_guard = {gexec};
    if (graphs) {
        for (; it < 2; ++it)
            for (int g = 0; g < G; ++g) {
                EK(enqueue(g, (int)(it & 1), false, it >= 1, ev_pass[g], ev_side[g]));
                join_guard.side_used[g][it & 1] = true;
            }
        for (int g = 0; g < G && graphs; ++g) {
            EK(cudaStreamWaitEvent(sa[g], ev_side[g][0], 0));
            EK(cudaStreamWaitEvent(sa[g], ev_side[g][1], 0));
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(sa[g], cudaStreamCaptureModeThreadLocal) != cudaSuccess) { graphs = false; break; }
            cudaError_t ce = cudaSuccess;
            for (int j = 0; j < QR_GRAPH_ITERS && ce == cudaSuccess; ++j)
                ce = enqueue(g, j & 1, j >= 2, j >= 1, res->cap_pass[g], res->cap_side[g]);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(sa[g], res->cap_side[g][0], 0);      // join the side stream
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(sa[g], res->cap_side[g][1], 0);
            const cudaError_t ee = cudaStreamEndCapture(sa[g], &graph);
            if (ce == cudaSuccess && ee == cudaSuccess && graph) ce = cudaGraphInstantiate(&gexec[g], graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (ce != cudaSuccess || ee != cudaSuccess) { cudaGetLastError(); gexec[g] = nullptr; graphs = false; }
        }
    }
    if (graphs) {
        for (; it < max_passes && nfin < G; it += QR_GRAPH_ITERS) {
            for (int g = 0; g < G; ++g)
                if (!fin[g]) EK(cudaGraphLaunch(gexec[g], sa[g]));
            if (((it - 2) / QR_GRAPH_ITERS) % (poll / QR_GRAPH_ITERS) == (poll / QR_GRAPH_ITERS) - 1) EK(poll_groups());
        }
    } else {
        for (; it < max_passes && nfin < G; ++it) {
            const int buf = (int)(it & 1);
            for (int g = 0; g < G; ++g) {
                if (fin[g]) continue;
                EK(enqueue(g, buf, it >= 2, it >= 1, ev_pass[g], ev_side[g]));
                join_guard.side_used[g][buf] = true;
            }
            if (hf && (it % poll) == poll - 1) EK(poll_groups());
        }
    }
    // join all internal streams back into the caller's stream before anything reads H or Z
    st = user_st;
    join_guard.run();
    qr_finish_kernel<<<(nb + 127) / 128, 128, 0, st>>>(ws.states, nb, info);

    // ---------------- phase 3: Schur form T = Z^H A0 Z (upper triangle; the strictly lower part is round-off
    // and is never read), eigenvalues, eigenvectors of T, back-transformation, normalisation
    cplx* Tm = ws.X;            // A0 -> T
    cplx* Xv = A;               // scratch for A0*Z, then the eigenvector matrix of T
    EK(zgemm_strided(OP_N, OP_N, n, n, n, one, ws.X, n, ms, ws.Z, n, ms, zero, A, n, ms, nb, ws.gs, st));        // A  = A0 Z
    EK(zgemm_strided(OP_H, OP_N, n, n, n, one, ws.Z, n, ms, A, n, ms, zero, Tm, n, ms, nb, ws.gs, st, ZGEMM_C_UPPER));   // T  = Z^H (A0 Z), upper tiles only
    diag_extract_kernel<<<dim3((n + 255) / 256, nb), 256, 0, st>>>(Tm, ms, n, n, wout);
    tnorm_kernel<<<nb, 256, 0, st>>>(wout, n, ws.tnorm);
    EK(cudaMemsetAsync(Xv, 0, sizeof(cplx) * (size_t)ms * nb, st));
    const int nblk = (n + TV_NB - 1) / TV_NB;
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int r0 = kb * TV_NB, nbk = (n - r0 < TV_NB) ? n - r0 : TV_NB, r1 = r0 + nbk;
        if (r1 < n) {
            // X[I, r1:n] = -T[I, r1:n] * X[r1:n, r1:n]
            EK(zgemm_strided(OP_N, OP_N, nbk, n - r1, n - r1, C(-1, 0), Tm + (size_t)r0 * n + r1, n, ms,
                             Xv + (size_t)r1 * n + r1, n, ms, zero, Xv + (size_t)r0 * n + r1, n, ms, nb, ws.gs, st, ZGEMM_B_UPPER));
        }
        trevc_block_kernel<<<dim3((n - r0 + 127) / 128, nb), 128, 0, st>>>(Tm, ms, n, n, r0, nbk, wout, ws.tnorm, Xv, ms, n);
    }
    EK(zgemm_strided(OP_N, OP_N, n, n, n, one, ws.Z, n, ms, Xv, n, ms, zero, V, n, ms, nb, ws.gs, st, ZGEMM_B_UPPER));   // X is upper triangular
    colnorm_kernel<<<dim3((n + 31) / 32, nb), dim3(32, 8), 0, st>>>(V, ms, n, n, ws.nrm);
    colscale_kernel<<<dim3((n + 255) / 256, n, nb), 256, 0, st>>>(V, ms, n, n, ws.nrm);
    return cudaGetLastError();
}

}  // namespace rcwa
#endif
