// Stage 2: batched complex128 non-Hermitian eigendecomposition, entirely on the device.
// Replaces torch.linalg.eig in Eig.forward (/root/reference/torcwa/torch_eig.py:11-17; LAPACK
// zgeev on CPU, cuSOLVER/MAGMA hybrid on CUDA).  Four phases per batch of matrices:
//
//  (1) Hessenberg reduction  A = Z H Z^H   by Householder reflectors H_k = I - u u^H (|u|^2 = 2).
//      One fused streaming pass per column over the stacked matrix [A; Z]:
//          a_ij <- a_ij - u_i w~_j - y~_i conj(u_j)          (two-sided rank-2 update of step k)
//      while the same pass accumulates  y' = A_new u'  and  w' = u'^H A_new  for step k+1
//      (u' is built first from the updated column k+1 by a small per-matrix kernel).  Rows of Z
//      ride along with u_i = 0 (right-multiplication only).  HBM-bound: one read + one write of
//      the trailing region per column; row dot-products by warp shuffles, column sums as
//      per-row-band partials (deterministic, no atomics).
//  (2) Windowed multishift QR  H -> T (upper triangular), Z <- Z U:  chains of up to QR_NS
//      single-shift Givens bulges (spacing 2) are chased through a QR_W x QR_W diagonal window held
//      in shared memory by ONE CTA per matrix, all bulges advancing one position per step; the
//      window's accumulated unitary U is then applied to the off-diagonal row panel, column panel
//      and to Z by three grouped DMMA GEMMs (zgemm.cu) over the whole batch.  Shifts =
//      eigenvalues of the trailing QR_NS x QR_NS block (single-warp shifted QR in shared memory);
//      deflation by the conservative LAPACK zlahqr criterion.  The host only enqueues
//      (pass kernel, 2 GEMM launches) repeatedly and polls a device counter through pinned memory.
//  (3) Eigenvectors of T by blocked back-substitution (one GEMM + one per-column small triangular
//      solve per 32-row block), then V = Z X (GEMM) and unit 2-norm columns (LAPACK geev convention).
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
#define QR_SLICE_ROT 160    // rotations per launch of a small-block solve (time slice: keeps the batch in lockstep)
#define QR_MAXSTALL 40     // sweeps without deflation before giving up on a matrix
#define TV_NB 32           // eigenvector back-substitution block

struct QrState {
    int lo, hi;          // active block [lo, hi] (inclusive); done when hi < 1
    int phase;           // 0: start a new sweep (deflation scan + shifts); 1: chain in flight; 2: small-block solve
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
    int pad0;
    int kpos[QR_NS];     // column of each bulge (leading first): bulge element is H[k+2][k]
    cplx shifts[QR_NS];
};

// ------------------------------------------------------------------------------------------------
// Givens rotation G = [[c, s], [-conj(s), c]] with G [a; b] = [r; 0], c real >= 0 (LAPACK zlartg).
HD void givens(cplx a, cplx b, double& c, cplx& s, cplx& r) {
    if (cis_zero(b)) { c = 1.0; s = C(0, 0); r = a; return; }
    if (cis_zero(a)) { c = 0.0; double nb = cabs_(b); s = cscale(cconj(b), 1.0 / nb); r = C(nb, 0); return; }
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
// Resumable single-shift QR (zlahqr with Schur vectors) on an m x m upper-Hessenberg block held in
// shared memory (Hs, Us with leading dimension QR_LD; Us accumulates the unitary).  Runs at most
// `budget` rotations, then returns; state (i, its) lives in QrState so the next launch continues.
// Returns 1 when the block is upper triangular, 0 if more slices are needed, -1 on failure.
DEV int small_schur_slice(const Cta& c, cplx* Hs, cplx* Us, int m, int* pi, int* pits, int budget) {
    int i = *pi, its = *pits, used = 0;
    while (i >= 1) {
        // ---- deflation scan (uniform scalar code; every thread reads the same shared values)
        int l;
        for (l = i; l > 0; --l) {
            cplx h10 = Hs[l * QR_LD + l - 1];
            if (cis_zero(h10)) break;
            double extra = 0.0;
            if (l - 2 >= 0) extra += cabs1(Hs[(l - 1) * QR_LD + l - 2]);
            if (l + 1 <= i) extra += cabs1(Hs[(l + 1) * QR_LD + l]);
            if (negligible_subdiag(h10, Hs[(l - 1) * QR_LD + l - 1], Hs[l * QR_LD + l], Hs[(l - 1) * QR_LD + l], extra)) break;
        }
        CTA_SYNC();
        if (l > 0 && c.tid == 0) Hs[l * QR_LD + l - 1] = C(0, 0);
        CTA_SYNC();
        if (l >= i) { --i; its = 0; continue; }
        if (its > 60) { *pi = i; *pits = its; return -1; }
        if (used + (i - l) > budget && used > 0) break;       // out of time: resume at the next launch
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
        // ---- one QR sweep l..i with Schur-vector accumulation (full rows/columns of the block)
        for (int k = l; k < i; ++k) {
            cplx a, b;
            if (k == l) { a = csub(Hs[k * QR_LD + k], sig); b = Hs[(k + 1) * QR_LD + k]; }
            else { a = Hs[k * QR_LD + k - 1]; b = Hs[(k + 1) * QR_LD + k - 1]; }
            double cs; cplx sn, r;
            givens(a, b, cs, sn, r);
            CTA_SYNC();
            if (k > l && c.tid == 0) { Hs[k * QR_LD + k - 1] = r; Hs[(k + 1) * QR_LD + k - 1] = C(0, 0); }
            for (int j = k + c.tid; j < m; j += c.nthreads) {
                cplx x = Hs[k * QR_LD + j], y = Hs[(k + 1) * QR_LD + j];
                Hs[k * QR_LD + j] = cadd(cscale(x, cs), cmul(sn, y));
                Hs[(k + 1) * QR_LD + j] = csub(cscale(y, cs), cmul(cconj(sn), x));
            }
            CTA_SYNC();
            const int rmax = (k + 2 < i) ? k + 2 : i;
            for (int idx = c.tid; idx < (rmax + 1) + m; idx += c.nthreads) {
                cplx* base = (idx <= rmax) ? (Hs + idx * QR_LD) : (Us + (idx - rmax - 1) * QR_LD);
                cplx x = base[k], y = base[k + 1];
                base[k] = cadd(cscale(x, cs), cmul(y, cconj(sn)));
                base[k + 1] = csub(cscale(y, cs), cmul(x, sn));
            }
            CTA_SYNC();      // the next rotation reads the bulge this right-update just created
        }
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
    double cs[QR_NS + 1];
    cplx sn[QR_NS + 1];
    int rot_row[QR_NS + 1];     // local upper row index of each rotation this step
    int rot_col0[QR_NS + 1];    // first local column of the left update
    int nrot;
    int moved_any;
    int flag;
    double red[40];
};

HD size_t qr_pass_smem_bytes(int n) {
    return 2 * (size_t)QR_W * QR_LD * sizeof(cplx) + sizeof(QrScratch) + (size_t)(QR_NS * (QR_NS + 1)) * sizeof(cplx) + (size_t)(n + 16) + 64;
}

DEV void qr_pass_body(const Cta& c, cplx* H, int ldh, int n, cplx* Zm, int ldz, QrState* stg,
                      cplx* Ug, ZGemmProblem* prob_rows, ZGemmProblem* prob_cols, ZGemmProblem* prob_z) {
    cplx* Hs = reinterpret_cast<cplx*>(c.smem);
    cplx* Us = Hs + QR_W * QR_LD;
    QrScratch* sc = reinterpret_cast<QrScratch*>(Us + QR_W * QR_LD);
    cplx* Ts = reinterpret_cast<cplx*>(sc + 1);                         // [QR_NS][QR_NS+1] shift block
    unsigned char* negl = reinterpret_cast<unsigned char*>(Ts + QR_NS * (QR_NS + 1));   // [n] deflation flags

    // default: idle problems
    if (c.tid == 0) { prob_rows->M = 0; prob_cols->M = 0; prob_z->M = 0; }
    QrState st = *stg;          // every thread holds a private copy; thread 0 writes it back
    CTA_SYNC();
    if (st.done) return;

    if (st.phase == 0) {
        // ---------------- new sweep: deflation scan over the whole remaining matrix [1..hi]
        for (int k = 1 + c.tid; k <= st.hi; k += c.nthreads) {
            cplx h10 = H[(size_t)k * ldh + k - 1];
            bool z = cis_zero(h10);
            if (!z) {
                double extra = 0.0;
                if (k - 2 >= 0) extra += cabs1(H[(size_t)(k - 1) * ldh + k - 2]);
                if (k + 1 <= st.hi) extra += cabs1(H[(size_t)(k + 1) * ldh + k]);
                z = negligible_subdiag(h10, H[(size_t)(k - 1) * ldh + k - 1], H[(size_t)k * ldh + k], H[(size_t)(k - 1) * ldh + k], extra);
                if (z) H[(size_t)k * ldh + k - 1] = C(0, 0);
            }
            negl[k] = z ? 1 : 0;
        }
        CTA_SYNC();
        // walk down from hi (uniform scalar code on shared flags)
        int hi = st.hi;
        while (hi >= 1 && negl[hi]) --hi;
        int lo = hi;
        while (lo >= 1 && !negl[lo]) --lo;
        if (hi < 1) {
            if (c.tid == 0) { st.done = 1; st.hi = hi; *stg = st; }
            return;
        }
        if (hi < st.hi_prev || lo > st.lo) st.stall = 0; else st.stall++;
        st.hi_prev = hi;
        st.lo = lo; st.hi = hi;
        if (st.stall > QR_MAXSTALL) {
            if (c.tid == 0) { st.done = 1; st.info = hi + 1; *stg = st; }
            return;
        }
        const int m = hi - lo + 1;
        if (m <= QR_SMALL) {
            // small active block: Schur-factor it directly in shared memory (time-sliced)
            st.phase = 2; st.p = lo; st.ss_i = m - 1; st.ss_its = 0; st.ss_fresh = 1; st.small_solves++;
        }
        const int ns = (m < QR_NS) ? m : QR_NS;
        if (st.phase != 2) {
        // ---------------- shifts: eigenvalues of the trailing ns x ns block (warp 0)
        for (int idx = c.tid; idx < ns * ns; idx += c.nthreads) {
            int r = idx / ns, q = idx % ns;
            Ts[r * (QR_NS + 1) + q] = (r <= q + 1) ? H[(size_t)(hi - ns + 1 + r) * ldh + (hi - ns + 1 + q)] : C(0, 0);
        }
        CTA_SYNC();
        const bool exceptional = (st.stall % 6 == 5);
#ifndef RCWA_EMU
        if (c.tid < 32) {
            tiny_hqr_eigs(c.tid, 32, Ts, QR_NS + 1, ns, st.shifts);
        }
        // broadcast the shifts computed by warp 0 through shared memory
        CTA_SYNC();
        if (c.tid < 32) { for (int j = c.tid; j < ns; j += 32) Ts[j] = st.shifts[j]; }
        CTA_SYNC();
        for (int j = 0; j < ns; ++j) st.shifts[j] = Ts[j];
        CTA_SYNC();
#else
        tiny_hqr_eigs(0, 1, Ts, QR_NS + 1, ns, st.shifts);
#endif
        if (exceptional) {
            const double mag = 0.75 * cabs1(H[(size_t)hi * ldh + hi - 1]);
            for (int j = 0; j < ns; ++j) st.shifts[j] = cadd(H[(size_t)hi * ldh + hi], C(mag * ((j & 1) ? -1.0 : 1.0), mag * 0.5 * (j % 3 - 1)));
        }
        st.ns = ns; st.nintro = 0; st.nbulge = 0; st.p = lo; st.phase = 1; st.sweeps++;
        }
#ifdef RCWA_EMU
        if (getenv("RCWA_EMU_DEBUG")) fprintf(stderr, "sweep %d: lo=%d hi=%d ns=%d stall=%d sub=%.3e small=%d\n", st.sweeps, lo, hi, ns, st.stall, cabs1(H[(size_t)hi * ldh + hi - 1]), st.small_solves);
#endif
    }

    if (st.phase == 2) {
        // ---------------- one time slice of the small-block solve on [lo, hi]
        const int p2 = st.lo, m2 = st.hi - st.lo + 1;
        for (int idx = c.tid; idx < m2 * m2; idx += c.nthreads) {
            int r = idx / m2, q = idx % m2;
            Hs[r * QR_LD + q] = H[(size_t)(p2 + r) * ldh + (p2 + q)];
            Us[r * QR_LD + q] = st.ss_fresh ? C(r == q ? 1.0 : 0.0, 0.0) : Ug[r * QR_W + q];
        }
        CTA_SYNC();
        int si = st.ss_i, sits = st.ss_its;
        const int rc = small_schur_slice(c, Hs, Us, m2, &si, &sits, QR_SLICE_ROT);
        st.ss_i = si; st.ss_its = sits; st.ss_fresh = 0;
        for (int idx = c.tid; idx < m2 * m2; idx += c.nthreads) {
            int r = idx / m2, q = idx % m2;
            H[(size_t)(p2 + r) * ldh + (p2 + q)] = Hs[r * QR_LD + q];
            Ug[r * QR_W + q] = Us[r * QR_LD + q];
        }
        if (c.tid == 0) {
            st.passes++;
            if (rc != 0) {
                // finished (or failed): apply the accumulated unitary to the off-diagonal panels and Z
                const int wend2 = st.hi + 1;
                ZGemmProblem g;
                g.A = Ug; g.lda = QR_W; g.B = H + (size_t)p2 * ldh + wend2; g.ldb = ldh; g.C = H + (size_t)p2 * ldh + wend2; g.ldc = ldh;
                g.M = (n - wend2 > 0) ? m2 : 0; g.N = n - wend2; g.K = m2; *prob_rows = g;
                g.A = H + p2; g.lda = ldh; g.B = Ug; g.ldb = QR_W; g.C = H + p2; g.ldc = ldh;
                g.M = p2; g.N = m2; g.K = m2; *prob_cols = g;
                g.A = Zm + p2; g.lda = ldz; g.B = Ug; g.ldb = QR_W; g.C = Zm + p2; g.ldc = ldz;
                g.M = n; g.N = m2; g.K = m2; *prob_z = g;
                st.phase = 0;
                if (rc < 0) { st.done = 1; st.info = st.lo + si + 1; }
            }
            *stg = st;
        }
        return;
    }

    // ---------------- window [p, wend)
    const int p = st.p;
    const int wend = (p + QR_W < st.hi + 1) ? p + QR_W : st.hi + 1;
    const int wl = wend - p;
    const bool at_bottom = (wend - 1 == st.hi);
    for (int idx = c.tid; idx < wl * wl; idx += c.nthreads) {
        int r = idx / wl, q = idx % wl;
        Hs[r * QR_LD + q] = H[(size_t)(p + r) * ldh + (p + q)];
        Us[r * QR_LD + q] = C(r == q ? 1.0 : 0.0, 0.0);
    }
    CTA_SYNC();

    // local bulge columns (relative to p); -1 marks the virtual column of an introduction
    for (;;) {
        // ---- decide which rotations happen this step (uniform scalar code, every thread)
        int nrot = 0;
        int new_k[QR_NS];
        int prev_new = 1 << 30;      // new position of the bulge ahead
        int nb_after = 0;
        int rrow[QR_NS + 1], rcol0[QR_NS + 1], rb[QR_NS + 1];
        for (int b = 0; b < st.nbulge; ++b) {
            const int k = st.kpos[b] - p;          // local column of the bulge (>= 0)
            bool can = false, exits = false;
            if (at_bottom) { if (k + 2 <= wl - 1) { can = true; exits = (k + 2 == wl - 1); } }
            else if (k + 3 <= wl - 1) can = true;
            if (can && !exits && !(k + 1 + 2 <= prev_new)) can = false;
            if (can) { rrow[nrot] = k + 1; rcol0[nrot] = k; rb[nrot] = b; ++nrot; new_k[b] = exits ? -999 : k + 1; }
            else new_k[b] = k;
            if (new_k[b] != -999) prev_new = new_k[b];
        }
        // introduction of the next bulge at the top of the active block
        bool intro = false;
        if (p == st.lo && st.nintro < st.ns && wl >= 2) {
            // rows (0,1) must be free this step and the new bulge (column 0) needs spacing 2 behind the
            // trailing one: new position of the trailing bulge >= 2 (prev_new is huge when none is left)
            if (prev_new >= 2) { intro = true; rrow[nrot] = 0; rcol0[nrot] = -1; rb[nrot] = -1; ++nrot; }
        }
        if (nrot == 0) break;
        // ---- rotation parameters
        if (c.tid < nrot) {
            const int t = c.tid, r1 = rrow[t];
            cplx a, b;
            if (rcol0[t] < 0) { a = csub(Hs[0], st.shifts[st.nintro]); b = Hs[1 * QR_LD + 0]; }
            else { a = Hs[r1 * QR_LD + rcol0[t]]; b = Hs[(r1 + 1) * QR_LD + rcol0[t]]; }
            // (a, b) both at round-off level (the chain runs over an already converged spot): a rotation
            // built from noise would scramble converged rows -> identity.  A tiny b next to a
            // non-negligible a is kept: small bulges still carry the shifts.
            if (cabs1(a) + cabs1(b) <= RCWA_EPS * (cabs1(Hs[r1 * QR_LD + r1]) + cabs1(Hs[(r1 + 1) * QR_LD + r1 + 1]))) { a = C(0, 0); b = C(0, 0); }
            double cs; cplx sn, r;
            givens(a, b, cs, sn, r);
            sc->cs[t] = cs; sc->sn[t] = sn;
            if (rcol0[t] >= 0) { Hs[r1 * QR_LD + rcol0[t]] = r; Hs[(r1 + 1) * QR_LD + rcol0[t]] = C(0, 0); }
#ifdef RCWA_EMU
            if (getenv("RCWA_EMU_DEBUG2")) fprintf(stderr, "  rot t=%d r1=%d col0=%d nintro=%d shift=(%g,%g) a=(%g,%g) b=(%g,%g) c=%g s=(%g,%g)\n", t, r1, rcol0[t], st.nintro, st.shifts[st.nintro].x, st.shifts[st.nintro].y, a.x, a.y, b.x, b.y, cs, sn.x, sn.y);
#endif
        }
#ifdef RCWA_EMU
        for (int t = 1; t < nrot; ++t) {     // the single emulated thread plays threads 1..nrot-1
            const int r1 = rrow[t];
            cplx a, b;
            if (rcol0[t] < 0) { a = csub(Hs[0], st.shifts[st.nintro]); b = Hs[1 * QR_LD + 0]; }
            else { a = Hs[r1 * QR_LD + rcol0[t]]; b = Hs[(r1 + 1) * QR_LD + rcol0[t]]; }
            // (a, b) both at round-off level (the chain runs over an already converged spot): a rotation
            // built from noise would scramble converged rows -> identity.  A tiny b next to a
            // non-negligible a is kept: small bulges still carry the shifts.
            if (cabs1(a) + cabs1(b) <= RCWA_EPS * (cabs1(Hs[r1 * QR_LD + r1]) + cabs1(Hs[(r1 + 1) * QR_LD + r1 + 1]))) { a = C(0, 0); b = C(0, 0); }
            double cs; cplx sn, r;
            givens(a, b, cs, sn, r);
            sc->cs[t] = cs; sc->sn[t] = sn;
            if (rcol0[t] >= 0) { Hs[r1 * QR_LD + rcol0[t]] = r; Hs[(r1 + 1) * QR_LD + rcol0[t]] = C(0, 0); }
        }
#endif
        CTA_SYNC();
        // ---- left: rows (r1, r1+1), columns from col0+1 (or 0 for an introduction) to wl-1
        for (int idx = c.tid; idx < nrot * QR_W; idx += c.nthreads) {
            const int t = idx / QR_W, j = idx % QR_W, r1 = rrow[t];
            if (j <= rcol0[t] || j >= wl) continue;
            const double cs = sc->cs[t]; const cplx sn = sc->sn[t];
            cplx x = Hs[r1 * QR_LD + j], y = Hs[(r1 + 1) * QR_LD + j];
            Hs[r1 * QR_LD + j] = cadd(cscale(x, cs), cmul(sn, y));
            Hs[(r1 + 1) * QR_LD + j] = csub(cscale(y, cs), cmul(cconj(sn), x));
        }
        CTA_SYNC();
        // ---- right: columns (r1, r1+1), rows 0..min(r1+2, wl-1) of H and all rows of U
        for (int idx = c.tid; idx < nrot * 2 * QR_W; idx += c.nthreads) {
            const int t = idx / (2 * QR_W), rem = idx % (2 * QR_W), r1 = rrow[t];
            const double cs = sc->cs[t]; const cplx sn = sc->sn[t];
            if (rem < QR_W) {
                const int i = rem;
                const int imax = (r1 + 2 < wl - 1) ? r1 + 2 : wl - 1;
                if (i > imax) continue;
                cplx x = Hs[i * QR_LD + r1], y = Hs[i * QR_LD + r1 + 1];
                Hs[i * QR_LD + r1] = cadd(cscale(x, cs), cmul(y, cconj(sn)));
                Hs[i * QR_LD + r1 + 1] = csub(cscale(y, cs), cmul(x, sn));
            } else {
                const int i = rem - QR_W;
                if (i >= wl) continue;
                cplx x = Us[i * QR_LD + r1], y = Us[i * QR_LD + r1 + 1];
                Us[i * QR_LD + r1] = cadd(cscale(x, cs), cmul(y, cconj(sn)));
                Us[i * QR_LD + r1 + 1] = csub(cscale(y, cs), cmul(x, sn));
            }
        }
        CTA_SYNC();
        // ---- bookkeeping (uniform)
        nb_after = 0;
        for (int b = 0; b < st.nbulge; ++b) if (new_k[b] != -999) st.kpos[nb_after++] = new_k[b] + p;
        if (intro) {
            ++st.nintro;
            if (wl >= 3) st.kpos[nb_after++] = p + 0;        // bulge element now at H[p+2][p]
        }
        st.nbulge = nb_after;
        (void)rb;
    }

    // ---------------- write back window and U, emit GEMM problems, advance state
    for (int idx = c.tid; idx < wl * wl; idx += c.nthreads) {
        int r = idx / wl, q = idx % wl;
        H[(size_t)(p + r) * ldh + (p + q)] = Hs[r * QR_LD + q];
        Ug[r * QR_W + q] = Us[r * QR_LD + q];
    }
    if (c.tid == 0) {
        ZGemmProblem g;
        // rows: H[p:wend, wend:n] <- U^H * H[p:wend, wend:n]
        g.A = Ug; g.lda = QR_W; g.B = H + (size_t)p * ldh + wend; g.ldb = ldh; g.C = H + (size_t)p * ldh + wend; g.ldc = ldh;
        g.M = (n - wend > 0) ? wl : 0; g.N = n - wend; g.K = wl; *prob_rows = g;
        // cols: H[0:p, p:wend] <- H[0:p, p:wend] * U
        g.A = H + p; g.lda = ldh; g.B = Ug; g.ldb = QR_W; g.C = H + p; g.ldc = ldh;
        g.M = p; g.N = wl; g.K = wl; *prob_cols = g;
        // Z[:, p:wend] <- Z[:, p:wend] * U
        g.A = Zm + p; g.lda = ldz; g.B = Ug; g.ldb = QR_W; g.C = Zm + p; g.ldc = ldz;
        g.M = n; g.N = wl; g.K = wl; *prob_z = g;
        st.passes++;
        if (st.nbulge == 0 && (st.nintro >= st.ns)) st.phase = 0;       // chain gone: next pass starts a new sweep
        else if (st.nbulge == 0) st.phase = 0;                            // nothing could be introduced (tiny block)
        else {
            // next window starts at the trailing bulge (or stays at lo while bulges remain to be introduced)
            int trail = st.kpos[st.nbulge - 1];
            st.p = (st.nintro < st.ns && p == st.lo) ? trail : trail;
            if (st.nintro < st.ns && p == st.lo) st.ns = st.nintro;      // window full: cap this sweep's shifts
        }
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
extern "C" int emu_qr(cplx* H, cplx* Z, int n, int max_passes, int* stats) {
    std::vector<char> smem(qr_pass_smem_bytes(n));
    std::vector<cplx> U((size_t)QR_W * QR_W);
    QrState st; memset(&st, 0, sizeof(st));
    st.lo = 0; st.hi = n - 1; st.hi_prev = n - 1;
    Cta c; c.tid = 0; c.nthreads = 1; c.bid = 0; c.smem = smem.data();
    ZGemmProblem pr, pc, pz;
    int it = 0;
    for (; it < max_passes && !st.done; ++it) {
        qr_pass_body(c, H, n, n, Z, n, &st, U.data(), &pr, &pc, &pz);
        emu_gemm(pr, 2); emu_gemm(pc, 0); emu_gemm(pz, 0);
    }
    stats[0] = st.sweeps; stats[1] = st.passes; stats[2] = st.done; stats[3] = st.small_solves;
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
            ZGemmProblem g; g.A = T + (size_t)r0 * n + r1; g.lda = n; g.B = X + (size_t)r1 * n + r1; g.ldb = n;
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

// ---------------------------------------------------------------- phase 1: Hessenberg reduction
// Extended row space: rows [0,n) = A, rows [n,2n) = Z.
#define HS_ROWS_PER_CTA 64     // row band per CTA in the fused pass (8 warps x 8 rows)
#define HS_CHUNK 4             // columns per lane per chunk (=> 128 columns per warp iteration)

// Per-matrix vectors, all length 2n unless noted (workspace layout, complex):
//   u[n], unext[n], wt[n] (w~), y[2n] (raw y of the *next* step, written by the fused pass),
//   yt[2n] (y~ of the current step), wpart[nbands][n] (raw partial w of the next step)
struct HessVecs { cplx *u, *unext, *wt, *y, *yt, *wpart; cplx* beta; };

__device__ __forceinline__ HessVecs hess_vecs(cplx* base, int n, int nbands) {
    HessVecs v;
    v.u = base; v.unext = v.u + n; v.wt = v.unext + n; v.y = v.wt + n; v.yt = v.y + 2 * n; v.wpart = v.yt + 2 * n;
    v.beta = v.wpart + (size_t)nbands * n;
    return v;
}
__host__ __device__ inline size_t hess_vec_elems(int n, int nbands) { return (size_t)n * 3 + 4 * (size_t)n + (size_t)nbands * n + 8; }

// Step kernel, one CTA per matrix.  On entry (k >= 0): u = u_k, y = raw A u_k (2n), wpart = raw
// partial sums of u_k^H A.  Produces y~, w~ of step k, then the reflector u_{k+1} from the updated
// column k+1.  For k == -1 (bootstrap) there is no pending update: it only builds u_0 from column 0
// and the caller then runs a "pure accumulate" fused pass.
__global__ void __launch_bounds__(512, 1)
hess_step_kernel(cplx* A, long long astride, int lda, int n, int k, cplx* vecs, long long vstride, int nbands) {
    __shared__ double red[40];
    __shared__ cplx sh_c[2];
    const int b = blockIdx.x;
    cplx* Ab = A + (size_t)b * astride;
    HessVecs v = hess_vecs(vecs + (size_t)b * vstride, n, nbands);
    Cta c = make_cta(b, nullptr);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int kn = k + 1;                 // column that becomes the next reflector source
    if (k >= 0) {
        // raw w_j = sum over bands ; gamma = sum_j w_j u_j
        double gr = 0.0, gi = 0.0;
        for (int j = k + 1 + tid; j < n; j += nt) {
            cplx w = C(0, 0);
            const int nba = (n + HS_ROWS_PER_CTA - 1) / HS_ROWS_PER_CTA;     // bands that contain rows of A
            for (int q = 0; q < nba; ++q) w = cadd(w, v.wpart[(size_t)q * n + j]);
            v.wt[j] = w;
            cplx t = cmul(w, v.u[j]);
            gr += t.x; gi += t.y;
        }
        gr = cta_sum(c, gr, red); gi = cta_sum(c, gi, red);
        const cplx hg = C(0.5 * gr, 0.5 * gi);
        for (int j = k + 1 + tid; j < n; j += nt) v.wt[j] = csub(v.wt[j], cmul(hg, cconj(v.u[j])));
        for (int i = tid; i < 2 * n; i += nt) {
            cplx ui = (i < n && i > k) ? v.u[i] : C(0, 0);
            v.yt[i] = csub(v.y[i], cmul(hg, ui));
        }
        __syncthreads();
    }
    if (kn > n - 3) {       // no further reflector: mark u_next = 0
        for (int i = tid; i < n; i += nt) v.unext[i] = C(0, 0);
        return;
    }
    // updated column kn, rows i >= kn+1 (held in unext for now)
    double ss = 0.0;
    for (int i = kn + 1 + tid; i < n; i += nt) {
        cplx a = Ab[(size_t)i * lda + kn];
        if (k >= 0) {
            a = csub(a, cmul(v.u[i], v.wt[kn]));
            a = csub(a, cmul(v.yt[i], cconj(v.u[kn])));
        }
        v.unext[i] = a;
        ss += cabs2(a);
    }
    for (int i = tid; i <= kn && i < n; i += nt) v.unext[i] = C(0, 0);
    // scaled norm is unnecessary here: entries are O(|A|) and fp64 range is ample
    ss = cta_sum(c, ss, red);
    const double sigma = sqrt(ss);
    if (tid == 0) {
        cplx x1 = v.unext[kn + 1];
        cplx beta, ph;
        double ax = cabs_(x1);
        if (sigma == 0.0) { sh_c[0] = C(0, 0); sh_c[1] = C(0, 0); v.beta[0] = x1; }
        else {
            ph = (ax == 0.0) ? C(1, 0) : cscale(x1, 1.0 / ax);
            beta = cscale(ph, -sigma);
            sh_c[0] = beta;
            sh_c[1] = C(1.0 / sqrt(sigma * (sigma + ax)), 0.0);
            v.beta[0] = beta;
        }
    }
    __syncthreads();
    const cplx beta = sh_c[0];
    const double sc = sh_c[1].x;
    if (sc == 0.0) {
        for (int i = kn + 1 + tid; i < n; i += nt) v.unext[i] = C(0, 0);
    } else {
        for (int i = kn + 1 + tid; i < n; i += nt) {
            cplx a = v.unext[i];
            if (i == kn + 1) a = csub(a, beta);
            v.unext[i] = cscale(a, sc);
        }
    }
}

// Fused streaming pass over rows of [A; Z], columns [k+1, n):
//   a_ij <- a_ij - u_i w~_j - y~_i conj(u_j)      (skipped when k < 0)
//   column k+1: rows <= k+1 keep the updated value, row k+2 <- beta', rows > k+2 <- 0  (A rows only)
//   y'_i = sum_{j >= k+2} a_ij u'_j ;  wpart[band][j] = sum_{i in band} conj(u'_i) a_ij   (j >= k+2)
// grid (nbands_total = ceil(2n/64), B), 256 threads.
__global__ void __launch_bounds__(256)
hess_fused_kernel(cplx* A, long long astride, int lda, cplx* Z, long long zstride, int ldz, int n, int k,
                  cplx* vecs, long long vstride, int nbands) {
    __shared__ cplx wsh[8][128];       // per-warp partial column sums of the current chunk
    const int b = blockIdx.y, band = blockIdx.x;
    HessVecs v = hess_vecs(vecs + (size_t)b * vstride, n, nbands);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = band * HS_ROWS_PER_CTA + warp * 8;
    const bool a_band = (band * HS_ROWS_PER_CTA < n);     // bands never straddle A/Z when n % 64 != 0? handled per row
    const int kn = k + 1;
    const bool has_next = (kn <= n - 3);
    const cplx beta_n = v.beta[0];
    cplx* rowp[8]; cplx ui[8], uin[8], yti[8]; bool is_a[8]; int ri[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = row0 + r;
        ri[r] = i;
        if (i < n) { rowp[r] = A + (size_t)b * astride + (size_t)i * lda; is_a[r] = true; }
        else if (i < 2 * n) { rowp[r] = Z + (size_t)b * zstride + (size_t)(i - n) * ldz; is_a[r] = false; }
        else { rowp[r] = nullptr; is_a[r] = false; }
        ui[r] = (k >= 0 && i < n && i > k) ? v.u[i] : C(0, 0);
        uin[r] = (i < n && i > kn) ? cconj(v.unext[i]) : C(0, 0);
        yti[r] = (k >= 0 && i < 2 * n) ? v.yt[i] : C(0, 0);
    }
    cplx yacc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) yacc[r] = C(0, 0);
    const int jstart = kn - (kn % 128);        // aligned chunk start keeps 512-byte coalescing
    (void)a_band;
    for (int j0 = jstart; j0 < n; j0 += 128) {
        cplx wt[HS_CHUNK], ucj[HS_CHUNK], unj[HS_CHUNK], wacc[HS_CHUNK];
        int jj[HS_CHUNK];
#pragma unroll
        for (int q = 0; q < HS_CHUNK; ++q) {
            const int j = j0 + q * 32 + lane;
            jj[q] = j;
            const bool in = (j >= kn && j < n);
            wt[q] = (in && k >= 0) ? v.wt[j] : C(0, 0);
            ucj[q] = (in && k >= 0) ? cconj(v.u[j]) : C(0, 0);
            unj[q] = (in && j > kn) ? v.unext[j] : C(0, 0);
            wacc[q] = C(0, 0);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (!rowp[r]) continue;
#pragma unroll
            for (int q = 0; q < HS_CHUNK; ++q) {
                const int j = jj[q];
                if (j < kn || j >= n) continue;
                cplx a = rowp[r][j];
                if (k >= 0) {
                    a = csub(a, cmul(ui[r], wt[q]));
                    a = csub(a, cmul(yti[r], ucj[q]));
                }
                if (j == kn && is_a[r] && has_next) {
                    // this column is the source of reflector k+1: H_{k+1} x = beta e_1
                    if (ri[r] == kn + 1) a = beta_n;
                    else if (ri[r] > kn + 1) a = C(0, 0);
                }
                if (k >= 0 || (j == kn && is_a[r] && has_next)) rowp[r][j] = a;
                if (j > kn) {
                    yacc[r] = cfma(a, unj[q], yacc[r]);
                    wacc[q] = cfma(uin[r], a, wacc[q]);
                }
            }
        }
        // column partial sums of this chunk: warp-private slots, then reduce over the 8 warps
#pragma unroll
        for (int q = 0; q < HS_CHUNK; ++q) wsh[warp][q * 32 + lane] = wacc[q];
        __syncthreads();
        if (threadIdx.x < 128) {
            const int j = j0 + threadIdx.x;
            if (j > kn && j < n && band * HS_ROWS_PER_CTA < n) {
                cplx s = C(0, 0);
#pragma unroll
                for (int w = 0; w < 8; ++w) s = cadd(s, wsh[w][threadIdx.x]);
                v.wpart[(size_t)band * n + j] = s;
            }
        }
        __syncthreads();
    }
    // row dot products
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double yr = warp_sum(yacc[r].x), yi = warp_sum(yacc[r].y);
        if (lane == 0 && ri[r] < 2 * n) v.y[ri[r]] = C(yr, yi);
    }
}

__global__ void hess_advance_kernel(cplx* vecs, long long vstride, int n, int nbands) {
    // u <- unext
    HessVecs v = hess_vecs(vecs + (size_t)blockIdx.y * vstride, n, nbands);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v.u[i] = v.unext[i];
}

// ---------------------------------------------------------------- phase 2: QR passes
__global__ void __launch_bounds__(512, 1)
qr_pass_kernel(cplx* H, long long hstride, int ldh, int n, cplx* Z, long long zstride, int ldz, QrState* states,
               cplx* U, ZGemmProblem* prows, ZGemmProblem* pcolsz) {
    extern __shared__ __align__(16) char smem_raw[];
    const int b = blockIdx.x;
    Cta c = make_cta(b, smem_raw);
    qr_pass_body(c, H + (size_t)b * hstride, ldh, n, Z + (size_t)b * zstride, ldz, states + b,
                 U + (size_t)b * QR_W * QR_W, prows + b, pcolsz + 2 * b, pcolsz + 2 * b + 1);
}

__global__ void qr_init_kernel(QrState* states, int n, int nb) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    QrState st;
    memset(&st, 0, sizeof(st));
    st.lo = 0; st.hi = n - 1; st.hi_prev = n - 1;
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
    cplx *Z, *X, *U, *vecs; QrState* states; ZGemmProblem *prows, *pcolsz, *gs; int* flag; double *tnorm, *nrm;
    size_t total; int nbands; long long vstride;
};
EigWs carve(char* base, int n, int nb) {
    EigWs w; size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += al(bytes); return p; };
    w.nbands = (2 * n + HS_ROWS_PER_CTA - 1) / HS_ROWS_PER_CTA;
    w.vstride = (long long)hess_vec_elems(n, w.nbands);
    w.Z = (cplx*)take(sizeof(cplx) * (size_t)n * n * nb);
    w.X = (cplx*)take(sizeof(cplx) * (size_t)n * n * nb);
    w.U = (cplx*)take(sizeof(cplx) * (size_t)QR_W * QR_W * nb);
    w.vecs = (cplx*)take(sizeof(cplx) * (size_t)w.vstride * nb);
    w.states = (QrState*)take(sizeof(QrState) * (size_t)nb);
    w.prows = (ZGemmProblem*)take(sizeof(ZGemmProblem) * (size_t)nb);
    w.pcolsz = (ZGemmProblem*)take(sizeof(ZGemmProblem) * (size_t)nb * 2);
    w.gs = (ZGemmProblem*)take(sizeof(ZGemmProblem) * (size_t)nb * 4);
    w.flag = (int*)take(256);
    w.tnorm = (double*)take(sizeof(double) * (size_t)nb);
    w.nrm = (double*)take(sizeof(double) * (size_t)nb * n);
    w.total = off;
    return w;
}

}  // namespace

namespace rcwa {

size_t eig_workspace_bytes(int n, int nb) { return carve(nullptr, n, nb).total; }

#define EK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)

static cudaError_t hessenberg_phase(cplx* A, int n, int nb, const EigWs& ws, cudaStream_t st) {
    const long long ms = (long long)n * n;
    EK(set_identity(ws.Z, n, n, ms, nb, st));
    if (n >= 3) {
        const dim3 fgrid(ws.nbands, nb);
        // bootstrap: u_0 from column 0, then a pure accumulate pass (k = -1)
        hess_step_kernel<<<nb, 512, 0, st>>>(A, ms, n, n, -1, ws.vecs, ws.vstride, ws.nbands);
        hess_fused_kernel<<<fgrid, 256, 0, st>>>(A, ms, n, ws.Z, ms, n, n, -1, ws.vecs, ws.vstride, ws.nbands);
        hess_advance_kernel<<<dim3((n + 255) / 256, nb), 256, 0, st>>>(ws.vecs, ws.vstride, n, ws.nbands);
        for (int k = 0; k <= n - 3; ++k) {
            hess_step_kernel<<<nb, 512, 0, st>>>(A, ms, n, n, k, ws.vecs, ws.vstride, ws.nbands);
            hess_fused_kernel<<<fgrid, 256, 0, st>>>(A, ms, n, ws.Z, ms, n, n, k, ws.vecs, ws.vstride, ws.nbands);
            hess_advance_kernel<<<dim3((n + 255) / 256, nb), 256, 0, st>>>(ws.vecs, ws.vstride, n, ws.nbands);
        }
        EK(cudaGetLastError());
    }
    return cudaSuccess;
}

// A -> H (upper Hessenberg, in place), Zout = accumulated reflectors (A_in = Z H Z^H)
cudaError_t hessenberg(cplx* A, int n, int nb, cplx* Zout, char* wsb, size_t ws_bytes, cudaStream_t st) {
    EigWs ws = carve(wsb, n, nb);
    if (ws_bytes < ws.total) return cudaErrorInvalidValue;
    EK(hessenberg_phase(A, n, nb, ws, st));
    return cudaMemcpyAsync(Zout, ws.Z, sizeof(cplx) * (size_t)n * n * nb, cudaMemcpyDeviceToDevice, st);
}

cudaError_t eig(cplx* A, int n, int nb, cplx* wout, cplx* V, char* wsb, size_t ws_bytes, int* info, volatile int* host_flag, cudaStream_t st) {
    EigWs ws = carve(wsb, n, nb);
    if (ws_bytes < ws.total) return cudaErrorInvalidValue;
    const long long ms = (long long)n * n;
    const cplx one = C(1, 0), zero = C(0, 0);

    // ---------------- phase 1: Hessenberg, Z accumulated alongside
    EK(hessenberg_phase(A, n, nb, ws, st));

    // ---------------- phase 2: QR passes (host enqueues, polls the pinned flag every `poll` passes)
    qr_init_kernel<<<(nb + 127) / 128, 128, 0, st>>>(ws.states, n, nb);
    const size_t smem = qr_pass_smem_bytes(n);
    EK(cudaFuncSetAttribute(qr_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long max_passes = 40LL * n + 2000;       // generous: ~(n/NS) sweeps x (n/(W-2NS)) windows x iterations
    const int poll = 64;
    const int max_tiles_rows = gemm_tiles(GEMM_TILE_64x128, QR_W, n);
    const int max_tiles_cz = gemm_tiles(GEMM_TILE_128x64, n, QR_W);
    // The host polls convergence with a lag of one group: the count of group g is copied to pinned
    // memory asynchronously and examined after group g+1 has been enqueued, so the device never idles.
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int* hf = const_cast<int*>(host_flag);
    if (hf) { EK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming)); EK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming)); hf[0] = hf[1] = nb; }
    long long group = 0;
    bool finished = false;
    for (long long it = 0; it < max_passes && !finished; ++it) {
        qr_pass_kernel<<<nb, 512, smem, st>>>(A, ms, n, n, ws.Z, ms, n, ws.states, ws.U, ws.prows, ws.pcolsz);
        EK(zgemm_grouped(GEMM_TILE_64x128, OP_H, OP_N, ws.prows, nb, max_tiles_rows, one, zero, st));
        EK(zgemm_grouped(GEMM_TILE_128x64, OP_N, OP_N, ws.pcolsz, 2 * nb, max_tiles_cz, one, zero, st));
        if (hf && (it % poll) == poll - 1) {
            const int slot = (int)(group & 1);
            if (group >= 1) {       // examine the previous group's count (its copy was enqueued one group ago)
                EK(cudaEventSynchronize(ev[slot ^ 1]));
                if (hf[slot ^ 1] == 0) finished = true;
            }
            qr_count_kernel<<<1, 128, 0, st>>>(ws.states, nb, ws.flag + slot);
            EK(cudaMemcpyAsync(hf + slot, ws.flag + slot, sizeof(int), cudaMemcpyDeviceToHost, st));
            EK(cudaEventRecord(ev[slot], st));
            ++group;
        }
    }
    if (ev[0]) cudaEventDestroy(ev[0]);
    if (ev[1]) cudaEventDestroy(ev[1]);
    qr_finish_kernel<<<(nb + 127) / 128, 128, 0, st>>>(ws.states, nb, info);

    // ---------------- phase 3: eigenvalues, eigenvectors of T, back-transformation, normalisation
    diag_extract_kernel<<<dim3((n + 255) / 256, nb), 256, 0, st>>>(A, ms, n, n, wout);
    tnorm_kernel<<<nb, 256, 0, st>>>(wout, n, ws.tnorm);
    EK(cudaMemsetAsync(ws.X, 0, sizeof(cplx) * (size_t)ms * nb, st));
    const int nblk = (n + TV_NB - 1) / TV_NB;
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int r0 = kb * TV_NB, nbk = (n - r0 < TV_NB) ? n - r0 : TV_NB, r1 = r0 + nbk;
        if (r1 < n) {
            // X[I, r1:n] = -T[I, r1:n] * X[r1:n, r1:n]
            EK(zgemm_strided(OP_N, OP_N, nbk, n - r1, n - r1, C(-1, 0), A + (size_t)r0 * n + r1, n, ms,
                             ws.X + (size_t)r1 * n + r1, n, ms, zero, ws.X + (size_t)r0 * n + r1, n, ms, nb, ws.gs, st));
        }
        trevc_block_kernel<<<dim3((n - r0 + 127) / 128, nb), 128, 0, st>>>(A, ms, n, n, r0, nbk, wout, ws.tnorm, ws.X, ms, n);
    }
    EK(zgemm_strided(OP_N, OP_N, n, n, n, one, ws.Z, n, ms, ws.X, n, ms, zero, V, n, ms, nb, ws.gs, st));
    colnorm_kernel<<<dim3((n + 31) / 32, nb), dim3(32, 8), 0, st>>>(V, ms, n, n, ws.nrm);
    colscale_kernel<<<dim3((n + 255) / 256, n, nb), 256, 0, st>>>(V, ms, n, n, ws.nrm);
    return cudaGetLastError();
}

}  // namespace rcwa
#endif
