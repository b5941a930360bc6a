// extern "C" boundary of librcwa_b200.so -- see include/rcwa_b200.h for the contract.
#include "common.cuh"
#include "kernels.h"
#include "../../include/rcwa_b200.h"

using namespace rcwa;

namespace {
inline int cu(cudaError_t e) { return e == cudaSuccess ? 0 : RCWA_ERR_CUDA - (int)e; }
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
#define CK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return cu(_e); } while (0)
}  // namespace

extern "C" {

int rcwa_b200_abi_version(void) { return RCWA_B200_ABI_VERSION; }

size_t rcwa_gemm_scratch_bytes(int nb) { return align256(sizeof(ZGemmProblem) * (size_t)(nb > 0 ? nb : 1) * 4); }

size_t rcwa_convmat_workspace_bytes(int nx, int ny, int nb, int ox, int oy) {
    return align256(convmat_workspace_elems(nx, ny, nb, ox, oy) * sizeof(cplx));
}

int rcwa_convmat(const void* grid, int grid_type, long long grid_stride, int nx, int ny, int nb, int ox, int oy,
                 void* E, void* ws, void* stream) {
    if (!grid) return -1;
    if (grid_type < 0 || grid_type > 3) return -2;
    if (nx < 4 * ox + 1) return -4;
    if (ny < 4 * oy + 1) return -5;
    if (nb <= 0) return -6;
    if (ox < 0) return -7;
    if (oy < 0) return -8;
    if (!E) return -9;
    if (!ws) return -10;
    return cu(convmat(grid, grid_type, grid_stride, nx, ny, nb, ox, oy, (cplx*)E, (cplx*)ws, S(stream)));
}

int rcwa_zgemm_batched(int opa, int opb, int M, int N, int K, double alpha_re, double alpha_im,
                       const void* A, int lda, long long sa, const void* B, int ldb, long long sb,
                       double beta_re, double beta_im, void* Cm, int ldc, long long sc, int nb, void* gs, void* stream) {
    if (opa < 0 || opa > 2) return -1;
    if (opb < 0 || opb > 2) return -2;
    if (M < 0) return -3;
    if (N < 0) return -4;
    if (K < 0) return -5;
    if (!A) return -8;
    if (!B) return -11;
    if (!Cm) return -16;
    if (nb <= 0) return -19;
    if (!gs) return -20;
    return cu(zgemm_strided(opa, opb, M, N, K, C(alpha_re, alpha_im), (const cplx*)A, lda, sa, (const cplx*)B, ldb, sb,
                            C(beta_re, beta_im), (cplx*)Cm, ldc, sc, nb, (ZGemmProblem*)gs, S(stream)));
}

int rcwa_zgemm_batched_cfg(int cfg, int opa, int opb, int M, int N, int K, double alpha_re, double alpha_im,
                           const void* A, int lda, long long sa, const void* B, int ldb, long long sb,
                           double beta_re, double beta_im, void* Cm, int ldc, long long sc, int nb, void* gs, void* stream) {
    if (cfg < -1 || cfg > 15 || (cfg >= 0 && (cfg & 7) > 4)) return -1;
    if (opa < 0 || opa > 2) return -2;
    if (opb < 0 || opb > 2) return -3;
    if (M < 0 || N < 0 || K < 0) return -4;
    if (!A || !B || !Cm || !gs) return -9;
    if (nb <= 0) return -20;
    return cu(zgemm_strided_cfg(cfg, opa, opb, M, N, K, C(alpha_re, alpha_im), (const cplx*)A, lda, sa, (const cplx*)B, ldb, sb,
                                C(beta_re, beta_im), (cplx*)Cm, ldc, sc, nb, (ZGemmProblem*)gs, S(stream)));
}

size_t rcwa_zgemm_tc_workspace_bytes(int M, int N, int K, int nb, int slices) {
    if (M <= 0 || N <= 0 || K <= 0 || slices < 2 || slices > TC_MAXS) return 0;
    return tc_workspace_bytes(M, N, K, nb, slices);
}

int rcwa_zgemm_tc_batched(int slices, int opa, int opb, int M, int N, int K, double alpha,
                          const void* A, int lda, long long sa, const void* B, int ldb, long long sb,
                          double beta_re, double beta_im, void* Cm, int ldc, long long sc, int nb, void* ws, size_t ws_bytes, void* stream) {
    if (slices < 2 || slices > TC_MAXS) return -1;
    if (opa < 0 || opa > 2) return -2;
    if (opb < 0 || opb > 2) return -3;
    if (M <= 0) return -4;
    if (N <= 0) return -5;
    if (K <= 0 || K > 16000) return -6;
    if (!A) return -8;
    if (!B) return -11;
    if (!Cm) return -16;
    if (nb <= 0) return -19;
    if (!ws || ws_bytes < tc_workspace_min_bytes(M, N, K, slices)) return -20;
    if (!tc_supported(slices, M, N, K)) return RCWA_ERR_CUDA - (int)cudaErrorNotSupported;
    return cu(tc_zgemm_strided(slices, opa, opb, M, N, K, alpha, (const cplx*)A, lda, sa, (const cplx*)B, ldb, sb,
                               C(beta_re, beta_im), (cplx*)Cm, ldc, sc, nb, (char*)ws, ws_bytes, S(stream)));
}

int rcwa_tc_split(const void* X, int ld, long long stride, int rows_contiguous, int R, int Kc, int slices, int conj,
                  void* planes, int* ex, int nb, void* stream) {
    if (!X) return -1;
    if (R <= 0) return -5;
    if (Kc <= 0) return -6;
    if (slices < 2 || slices > TC_MAXS) return -7;
    if (!planes) return -9;
    if (!ex) return -10;
    if (nb <= 0) return -11;
    return cu(tc_split_debug((const cplx*)X, ld, stride, rows_contiguous, R, Kc, slices, conj, (signed char*)planes, ex, nb, S(stream)));
}

int rcwa_tc_schedule(int slices, int levels, unsigned* ops, int* meta, unsigned* loads, unsigned* mmas) {
    if (slices < 2 || slices > TC_MAXS) return -1;
    if (levels < 1 || levels > 4) return -2;
    if (!ops) return -3;
    if (!meta) return -4;
    TcSchedule sch;
    tc_build_schedule(slices, levels, &sch);
    TcTables tab;
    tc_compact_schedule(&sch, &tab);
    for (int i = 0; i < TC_MAXOPS; ++i) ops[i] = i < sch.nops ? sch.ops[i] : 0u;
    meta[0] = sch.ngroups; meta[1] = sch.nops;
    for (int g = 0; g < sch.ngroups; ++g) {
        meta[2 + 6 * g] = sch.g[g].d0; meta[3 + 6 * g] = sch.g[g].nl; meta[4 + 6 * g] = sch.g[g].op0;
        meta[5 + 6 * g] = sch.g[g].nops; meta[6 + 6 * g] = tab.nloads[g]; meta[7 + 6 * g] = tab.nmma[g];
    }
    if (loads) for (int i = 0; i < TC_MAXS * TC_MAXLOADS; ++i) loads[i] = tab.loads[i];
    if (mmas) for (int i = 0; i < TC_MAXS * TC_MAXMMAS; ++i) mmas[i] = tab.mmas[i];
    return 0;
}

int rcwa_tc_issue_entry(int slices, int levels, int group, int ring_pos, int step, unsigned* entry, int* nsteps) {
    if (slices < 2 || slices > TC_MAXS) return -1;
    if (levels < 1 || levels > 4) return -2;
    if (!entry) return -6;
    TcSchedule sch;
    tc_build_schedule(slices, levels, &sch);
    TcTables tab;
    tc_compact_schedule(&sch, &tab);
    if (group < 0 || group >= tab.ngroups) return -3;
    if (ring_pos < 0 || ring_pos >= TC_RING) return -4;
    if (nsteps) *nsteps = tab.nsteps[group];
    if (step < 0 || step >= tab.nsteps[group]) return -5;
    tc_issue_entry(tab, group, ring_pos, step, entry);
    return 0;
}

int rcwa_set_tuning(int key, int value) {
    if (key < 0 || key > 15) return -1;
    gemm_set_tuning(key, value);
    return 0;
}
int rcwa_get_tuning(int key) { return gemm_get_tuning(key); }

size_t rcwa_lu_tinv_bytes(int n, int nb) { return align256(lu_tinv_elems(n, nb > 0 ? nb : 1) * sizeof(cplx)); }
static size_t tinv_bytes(int n, int nb, int slices) { return align256(lu_tinv_elems(n, nb > 0 ? nb : 1, slices) * sizeof(cplx)); }
static int norm_slices(int s) { return (s >= 2 && s <= TC_MAXS) ? s : 0; }
// Triangular solves on the tcgen05 engine (512-wide right-looking blocks): off by default -- measured on B200 (profiles/
// r2g_stage_kernels_*.log) the K = 512 products are bound by the kernel's epilogue and the 512-wide block inverses cost
// more than they save; tuning key 12 = 1 turns them on for experiments.
static int solve_slices(int sl) { return gemm_get_tuning(12) ? sl : 0; }

int rcwa_lu_factor(void* A, long long stride, int n, int lda, int nb, int* ipiv, int* perm, int* info, void* tinv, void* gs, void* stream) {
    if (!A) return -1;
    if (n <= 0) return -3;
    if (lda < n) return -4;
    if (nb <= 0) return -5;
    if (!ipiv) return -6;
    if (!perm) return -7;
    if (!info) return -8;
    if (!tinv) return -9;
    if (!gs) return -10;
    return cu(lu_factor((cplx*)A, stride, n, lda, nb, ipiv, perm, info, (cplx*)tinv, (ZGemmProblem*)gs, S(stream)));
}

int rcwa_lu_solve_right(const void* LU, long long lus, int n, int lda, const int* perm, const void* tinv, const void* B, long long bs, int ldb,
                        int nrows, void* X, long long xs, int ldx, void* work, int nb, void* gs, void* stream) {
    if (!LU) return -1;
    if (n <= 0) return -3;
    if (!perm) return -5;
    if (!tinv) return -6;
    if (!B) return -7;
    if (nrows <= 0) return -10;
    if (!X || X == B) return -11;
    if (!work || work == X || work == B) return -14;
    if (nb <= 0) return -15;
    if (!gs) return -16;
    return cu(lu_solve_right((const cplx*)LU, lus, n, lda, perm, (const cplx*)tinv, (const cplx*)B, bs, ldb, nrows, (cplx*)X, xs, ldx,
                             (cplx*)work, nb, (ZGemmProblem*)gs, S(stream)));
}

int rcwa_pq_assemble(const void* eta, const void* E, const void* Mc, const void* nu, const void* mu_scalar,
                     const void* kx, const void* ky, int nb, int N, void* P, void* Q, void* stream) {
    if (!eta) return -1;
    if (!E) return -2;
    if ((Mc == nullptr) != (nu == nullptr)) return -3;
    if (!Mc && !mu_scalar) return -5;
    if (!kx) return -6;
    if (!ky) return -7;
    if (nb <= 0) return -8;
    if (N <= 0) return -9;
    if (!P) return -10;
    if (!Q) return -11;
    return cu(pq_assemble((const cplx*)eta, (const cplx*)E, (const cplx*)Mc, (const cplx*)nu, (const cplx*)mu_scalar,
                          (const cplx*)kx, (const cplx*)ky, nb, N, (cplx*)P, (cplx*)Q, S(stream)));
}

size_t rcwa_eig_workspace_bytes(int n, int nb) { return eig_workspace_bytes(n, nb); }

int rcwa_eig(void* A, int n, int nb, void* w, void* V, void* ws, size_t ws_bytes, int* info, void* host_flag, void* stream) {
    return rcwa_eig_phases(A, n, nb, w, V, ws, ws_bytes, info, host_flag, 3, stream);
}

int rcwa_eig_phases(void* A, int n, int nb, void* w, void* V, void* ws, size_t ws_bytes, int* info, void* host_flag, int phases, void* stream) {
    if (!A) return -1;
    if (n <= 0) return -2;
    if (nb <= 0) return -3;
    if (!w) return -4;
    if (!V) return -5;
    if (!ws || ws_bytes < eig_workspace_bytes(n, nb)) return -6;
    if (!info) return -8;
    if (phases < 1 || phases > 3) return -10;
    return cu(eig((cplx*)A, n, nb, (cplx*)w, (cplx*)V, (char*)ws, ws_bytes, info, (volatile int*)host_flag, S(stream), phases));
}

int rcwa_eig_stats(const void* ws, int n, int nb, int* out, void* stream) {
    if (!ws) return -1;
    if (n <= 0) return -2;
    if (nb <= 0) return -3;
    if (!out) return -4;
    return cu(eig_stats((const char*)ws, n, nb, out, S(stream)));
}

int rcwa_eig_profile(const void* ws, int n, int nb, long long* out, void* stream) {
    if (!ws) return -1;
    if (n <= 0) return -2;
    if (nb <= 0) return -3;
    if (!out) return -4;
    return cu(eig_profile((const char*)ws, n, nb, out, S(stream)));
}

int rcwa_hessenberg(void* A, int n, int nb, void* Z, void* ws, size_t ws_bytes, void* stream) {
    if (!A) return -1;
    if (n <= 0) return -2;
    if (nb <= 0) return -3;
    if (!Z) return -4;
    if (!ws || ws_bytes < eig_workspace_bytes(n, nb)) return -5;
    return cu(hessenberg((cplx*)A, n, nb, (cplx*)Z, (char*)ws, ws_bytes, S(stream)));
}

int rcwa_hessenberg_matvec_probe(const void* A, int n, int nb, int j, void* ws, size_t ws_bytes, void* stream) {
    if (!A) return -1;
    if (n < 3) return -2;
    if (nb <= 0) return -3;
    if (j < 0 || j > n - 3) return -4;
    if (!ws || ws_bytes < eig_workspace_bytes(n, nb)) return -5;
    return cu(eig_matvec_probe((const cplx*)A, n, nb, j, (char*)ws, ws_bytes, S(stream)));
}
int rcwa_hessenberg_panel_width(void) { return hessenberg_panel_width(); }

int rcwa_kz_branch(const void* lam, void* kz, long long total, void* stream) {
    if (!lam) return -1;
    if (!kz) return -2;
    if (total <= 0) return -3;
    return cu(kz_branch((const cplx*)lam, (cplx*)kz, (size_t)total, S(stream)));
}

// ---- Eig.backward (torcwa/torch_eig.py:19-44):  grad = X^-H (diag(g_lambda) + conj(F) o (X^H g_X)) X^H
size_t rcwa_eig_backward_workspace_bytes(int n, int nb) {
    const size_t mat = align256((size_t)n * n * nb * sizeof(cplx));
    return 4 * mat + 2 * align256((size_t)n * nb * sizeof(int)) + rcwa_lu_tinv_bytes(n, nb) + rcwa_gemm_scratch_bytes(nb);
}

int rcwa_eig_backward(const void* lam, const void* X, const void* glam, const void* gX, double delta, int nb, int n,
                      void* grad, void* ws, int* info, void* stream) {
    if (!lam) return -1;
    if (!X) return -2;
    if (delta < 0.0) return -5;
    if (nb <= 0) return -6;
    if (n <= 0) return -7;
    if (!grad) return -8;
    if (!ws) return -9;
    if (!info) return -10;
    cudaStream_t st = S(stream);
    const long long ms = (long long)n * n;
    const size_t mat = align256((size_t)ms * nb * sizeof(cplx));
    char* p = (char*)ws;
    cplx* b0 = (cplx*)p; p += mat;
    cplx* b1 = (cplx*)p; p += mat;
    cplx* b2 = (cplx*)p; p += mat;
    cplx* b3 = (cplx*)p; p += mat;
    int* ipiv = (int*)p; p += align256((size_t)n * nb * sizeof(int));
    int* perm = (int*)p; p += align256((size_t)n * nb * sizeof(int));
    cplx* tinv = (cplx*)p; p += rcwa_lu_tinv_bytes(n, nb);
    ZGemmProblem* gs = (ZGemmProblem*)p;
    const cplx one = C(1, 0), zero = C(0, 0);
    const cplx* Xc = (const cplx*)X;
    // T = X^H gX -> b0 ;  M = diag(glam) + conj(F) o T -> b1
    if (gX) CK(zgemm_strided(OP_H, OP_N, n, n, n, one, Xc, n, ms, (const cplx*)gX, n, ms, zero, b0, n, ms, nb, gs, st));
    CK(eig_backward_combine((const cplx*)lam, (const cplx*)glam, gX ? b0 : nullptr, delta, nb, n, b1, st));
    // grad = X^-H M X^H  =  (X M^H X^-1)^H :   Y = X M^H -> b0 ;  R = Y X^-1 (right-solve on the LU of X) -> b2 ;  grad = R^H
    CK(zgemm_strided(OP_N, OP_H, n, n, n, one, Xc, n, ms, b1, n, ms, zero, b0, n, ms, nb, gs, st));
    CK(cudaMemcpyAsync(b1, Xc, (size_t)ms * nb * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    CK(lu_factor(b1, ms, n, n, nb, ipiv, perm, info, tinv, gs, st));
    CK(lu_solve_right(b1, ms, n, n, perm, tinv, b0, ms, n, n, b2, ms, n, b3, nb, gs, st));
    CK(conj_transpose(b2, nb, n, (cplx*)grad, st));
    return 0;
}

// workspace layout helpers ---------------------------------------------------------------------
size_t rcwa_layer_smatrix_workspace_bytes(int N, int nb, int gemm_slices) {
    const size_t n = 2 * (size_t)N, mat = align256(n * n * nb * sizeof(cplx));
    const int sl = norm_slices(gemm_slices);
    return 6 * mat + 2 * align256(n * nb * sizeof(int)) + tinv_bytes((int)n, nb, solve_slices(sl)) + rcwa_gemm_scratch_bytes(nb) + align256(tc_ctx_bytes((int)n, nb, sl));
}

int rcwa_layer_smatrix(const void* W, const void* kz, const void* Q, const void* vfinv, const double* omega,
                       const double* thickness, int nb, int N, void* S11, void* S21, void* ws, int* info, int gemm_slices, void* stream) {
    if (!W) return -1;
    if (!kz) return -2;
    if (!Q) return -3;
    if (!vfinv) return -4;
    if (!omega) return -5;
    if (!thickness) return -6;
    if (nb <= 0) return -7;
    if (N <= 0) return -8;
    if (!S11) return -9;
    if (!S21) return -10;
    if (!ws) return -11;
    if (!info) return -12;
    cudaStream_t st = S(stream);
    const int n = 2 * N;
    const long long ms = (long long)n * n;
    const size_t mat = align256((size_t)ms * nb * sizeof(cplx));
    char* p = (char*)ws;
    cplx* b0 = (cplx*)p; p += mat;
    cplx* b1 = (cplx*)p; p += mat;
    cplx* b2 = (cplx*)p; p += mat;
    cplx* b3 = (cplx*)p; p += mat;
    cplx* b4 = (cplx*)p; p += mat;
    cplx* b5 = (cplx*)p; p += mat;
    int* ipiv = (int*)p; p += align256((size_t)n * nb * sizeof(int));
    int* perm = (int*)p; p += align256((size_t)n * nb * sizeof(int));
    const int sl = norm_slices(gemm_slices);
    cplx* tinv = (cplx*)p; p += tinv_bytes(n, nb, solve_slices(sl));
    ZGemmProblem* gs = (ZGemmProblem*)p; p += rcwa_gemm_scratch_bytes(nb);
    TcCtx tc = {sl, p, tc_ctx_bytes(n, nb, sl)};
    const int ssl = solve_slices(sl);
    // QW = Q * W
    CK(gemm_auto(&tc, OP_N, OP_N, n, n, n, 1.0, (const cplx*)Q, n, ms, (const cplx*)W, n, ms, C(0, 0), b0, n, ms, nb, gs, st));
    // M+ (b1), M- (b2), R+ (b3), R- (b4)
    CK(layer_form((const cplx*)W, b0, (const cplx*)kz, (const cplx*)vfinv, omega, thickness, nb, N, b1, b2, b3, b4, st));
    // T+ = R+ M+^-1  -> b0
    CK(lu_factor(b1, ms, n, n, nb, ipiv, perm, info, tinv, gs, st, true, ssl));
    CK(lu_solve_right(b1, ms, n, n, perm, tinv, b3, ms, n, n, b0, ms, n, b5, nb, gs, st, ssl ? &tc : nullptr));
    // T- = R- M-^-1  -> b3   (info keeps the first failure: the second factorisation does not clear it)
    CK(lu_factor(b2, ms, n, n, nb, ipiv, perm, info, tinv, gs, st, false, ssl));
    CK(lu_solve_right(b2, ms, n, n, perm, tinv, b4, ms, n, n, b3, ms, n, b5, nb, gs, st, ssl ? &tc : nullptr));
    CK(layer_finish(b0, b3, nb, n, (cplx*)S11, (cplx*)S21, st));
    return 0;
}

size_t rcwa_redheffer_workspace_bytes(int n, int nb, int gemm_slices) {
    const size_t mat = align256((size_t)n * n * nb * sizeof(cplx));
    const int sl = norm_slices(gemm_slices);
    return 5 * mat + 2 * align256((size_t)n * nb * sizeof(int)) + tinv_bytes(n, nb, solve_slices(sl)) + rcwa_gemm_scratch_bytes(nb) + align256(tc_ctx_bytes(n, nb, sl));
}

int rcwa_redheffer(const void* const Sm[4], const void* const Sn[4], void* const out[4], int nb, int n, void* ws, int* info, int gemm_slices, void* stream) {
    if (!Sm) return -1;
    if (!Sn) return -2;
    if (!out) return -3;
    for (int k = 0; k < 4; ++k) {
        if (!Sm[k]) return -1;
        if (!Sn[k]) return -2;
        if (!out[k]) return -3;
    }
    if (nb <= 0) return -4;
    if (n <= 0) return -5;
    if (!ws) return -6;
    if (!info) return -7;
    cudaStream_t st = S(stream);
    const long long ms = (long long)n * n;
    const size_t mat = align256((size_t)ms * nb * sizeof(cplx));
    char* p = (char*)ws;
    cplx* D = (cplx*)p; p += mat;
    cplx* Y1 = (cplx*)p; p += mat;
    cplx* Y2 = (cplx*)p; p += mat;
    cplx* G = (cplx*)p; p += mat;
    cplx* T = (cplx*)p; p += mat;
    int* ipiv = (int*)p; p += align256((size_t)n * nb * sizeof(int));
    int* perm = (int*)p; p += align256((size_t)n * nb * sizeof(int));
    const int sl = norm_slices(gemm_slices);
    cplx* tinv = (cplx*)p; p += tinv_bytes(n, nb, solve_slices(sl));
    ZGemmProblem* gs = (ZGemmProblem*)p; p += rcwa_gemm_scratch_bytes(nb);
    TcCtx tc = {sl, p, tc_ctx_bytes(n, nb, sl)};
    const int ssl = solve_slices(sl);
    const cplx *Sm11 = (const cplx*)Sm[0], *Sm21 = (const cplx*)Sm[1], *Sm12 = (const cplx*)Sm[2], *Sm22 = (const cplx*)Sm[3];
    const cplx *Sn11 = (const cplx*)Sn[0], *Sn21 = (const cplx*)Sn[1], *Sn12 = (const cplx*)Sn[2], *Sn22 = (const cplx*)Sn[3];
    cplx *O11 = (cplx*)out[0], *O21 = (cplx*)out[1], *O12 = (cplx*)out[2], *O22 = (cplx*)out[3];
    const cplx one = C(1, 0), zero = C(0, 0);
    const size_t bytes = (size_t)ms * nb * sizeof(cplx);
#define GEMM(a, b, beta, c) CK(gemm_auto(&tc, OP_N, OP_N, n, n, n, 1.0, a, n, ms, b, n, ms, beta, c, n, ms, nb, gs, st))
    // D = I - Sm12 Sn21
    CK(set_identity(D, n, n, ms, nb, st));
    CK(gemm_auto(&tc, OP_N, OP_N, n, n, n, -1.0, Sm12, n, ms, Sn21, n, ms, one, D, n, ms, nb, gs, st));
    CK(lu_factor(D, ms, n, n, nb, ipiv, perm, info, tinv, gs, st, true, ssl));
    // Y1 = Sn11 D^-1 ; Y2 = Sn21 D^-1     (T is free until later: it is the solves' work buffer)
    CK(lu_solve_right(D, ms, n, n, perm, tinv, Sn11, ms, n, n, Y1, ms, n, T, nb, gs, st, ssl ? &tc : nullptr));
    CK(lu_solve_right(D, ms, n, n, perm, tinv, Sn21, ms, n, n, Y2, ms, n, T, nb, gs, st, ssl ? &tc : nullptr));
    // G = Sm12 Sn22
    GEMM(Sm12, Sn22, zero, G);
    // S11 = Y1 Sm11
    GEMM(Y1, Sm11, zero, O11);
    // S12 = Sn12 + Y1 G
    CK(cudaMemcpyAsync(O12, Sn12, bytes, cudaMemcpyDeviceToDevice, st));
    GEMM(Y1, G, one, O12);
    // S21 = Sm21 + Sm22 (Y2 Sm11)
    GEMM(Y2, Sm11, zero, T);
    CK(cudaMemcpyAsync(O21, Sm21, bytes, cudaMemcpyDeviceToDevice, st));
    GEMM(Sm22, T, one, O21);
    // S22 = Sm22 (Sn22 + Y2 G)
    CK(cudaMemcpyAsync(D, Sn22, bytes, cudaMemcpyDeviceToDevice, st));
    GEMM(Y2, G, one, D);
    GEMM(Sm22, D, zero, O22);
#undef GEMM
    return 0;
}

int rcwa_redheffer_bdleft(const void* const Sm_bd[4], const void* const Sn[4], void* const out[4], int nb, int N, void* ws, int* info, int gemm_slices, void* stream) {
    if (!Sm_bd) return -1;
    if (!Sn) return -2;
    if (!out) return -3;
    for (int k = 0; k < 4; ++k) {
        if (!Sm_bd[k]) return -1;
        if (!Sn[k]) return -2;
        if (!out[k]) return -3;
    }
    if (nb <= 0) return -4;
    if (N <= 0) return -5;
    if (!ws) return -6;
    if (!info) return -7;
    cudaStream_t st = S(stream);
    const int n = 2 * N;
    const long long ms = (long long)n * n;
    const size_t mat = align256((size_t)ms * nb * sizeof(cplx));
    char* p = (char*)ws;
    cplx* D = (cplx*)p; p += mat;
    cplx* Y1 = (cplx*)p; p += mat;
    cplx* Y2 = (cplx*)p; p += mat;
    cplx* G = (cplx*)p; p += mat;
    cplx* T = (cplx*)p; p += mat;
    int* ipiv = (int*)p; p += align256((size_t)n * nb * sizeof(int));
    int* perm = (int*)p; p += align256((size_t)n * nb * sizeof(int));
    const int sl = norm_slices(gemm_slices);
    cplx* tinv = (cplx*)p; p += tinv_bytes(n, nb, solve_slices(sl));
    ZGemmProblem* gs = (ZGemmProblem*)p; p += rcwa_gemm_scratch_bytes(nb);
    TcCtx tc = {sl, p, tc_ctx_bytes(n, nb, sl)};
    const int ssl = solve_slices(sl);
    const cplx *m11 = (const cplx*)Sm_bd[0], *m21 = (const cplx*)Sm_bd[1], *m12 = (const cplx*)Sm_bd[2], *m22 = (const cplx*)Sm_bd[3];
    const cplx *Sn11 = (const cplx*)Sn[0], *Sn21 = (const cplx*)Sn[1], *Sn12 = (const cplx*)Sn[2], *Sn22 = (const cplx*)Sn[3];
    cplx *O11 = (cplx*)out[0], *O21 = (cplx*)out[1], *O12 = (cplx*)out[2], *O22 = (cplx*)out[3];
    const cplx one = C(1, 0), zero = C(0, 0), mone = C(-1, 0);
    const size_t bytes = (size_t)ms * nb * sizeof(cplx);
    // D = I - Sm12 Sn21          (Sm12 is four diagonals: an O(n^2) row combination, not a GEMM)
    CK(set_identity(D, n, n, ms, nb, st));
    CK(bd_left_mul(m12, Sn21, nb, N, n, mone, one, D, st));
    CK(lu_factor(D, ms, n, n, nb, ipiv, perm, info, tinv, gs, st, true, ssl));
    CK(lu_solve_right(D, ms, n, n, perm, tinv, Sn11, ms, n, n, Y1, ms, n, T, nb, gs, st, ssl ? &tc : nullptr));      // Y1 = Sn11 D^-1 (T = work)
    CK(lu_solve_right(D, ms, n, n, perm, tinv, Sn21, ms, n, n, Y2, ms, n, T, nb, gs, st, ssl ? &tc : nullptr));      // Y2 = Sn21 D^-1
    CK(bd_left_mul(m12, Sn22, nb, N, n, one, zero, G, st));                             // G = Sm12 Sn22
    CK(bd_right_mul(m11, Y1, nb, N, n, one, zero, O11, st));                            // S11 = Y1 Sm11
    CK(cudaMemcpyAsync(O12, Sn12, bytes, cudaMemcpyDeviceToDevice, st));                // S12 = Sn12 + Y1 G
    CK(gemm_auto(&tc, OP_N, OP_N, n, n, n, 1.0, Y1, n, ms, G, n, ms, one, O12, n, ms, nb, gs, st));
    CK(bd_right_mul(m11, Y2, nb, N, n, one, zero, T, st));                              // S21 = Sm21 + Sm22 (Y2 Sm11)
    CK(bd_left_mul(m22, T, nb, N, n, one, zero, O21, st));
    CK(bd_add(m21, nb, N, one, O21, st));
    CK(cudaMemcpyAsync(T, Sn22, bytes, cudaMemcpyDeviceToDevice, st));                  // S22 = Sm22 (Sn22 + Y2 G)
    CK(gemm_auto(&tc, OP_N, OP_N, n, n, n, 1.0, Y2, n, ms, G, n, ms, one, T, n, ms, nb, gs, st));
    CK(bd_left_mul(m22, T, nb, N, n, one, zero, O22, st));
    return 0;
}

int rcwa_blockdiag_dense(const void* d4, int nb, int N, void* D, void* stream) {
    if (!d4) return -1;
    if (nb <= 0) return -2;
    if (N <= 0) return -3;
    if (!D) return -4;
    return cu(blockdiag_dense((const cplx*)d4, nb, N, (cplx*)D, S(stream)));
}

int rcwa_sym_project(const void* X, int nb, int n, const int* il, const void* cl, const int* ir, const void* cr,
                     int G, int nkl, int nkr, void* out, void* stream) {
    if (!X) return -1;
    if (nb <= 0) return -2;
    if (n <= 0) return -3;
    if (!il || !cl) return -4;
    if (!ir || !cr) return -6;
    if (G < 1 || G > 4) return -8;
    if (nkl <= 0 || nkl > 65535) return -9;
    if (nkr <= 0) return -10;
    if (!out) return -11;
    return cu(sym_project((const cplx*)X, nb, n, il, (const cplx*)cl, ir, (const cplx*)cr, G, nkl, nkr, (cplx*)out, S(stream)));
}

}  // extern "C"
