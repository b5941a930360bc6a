// Internal C++ interface between the .cu translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

#define GEMM_TILE_64x128 0     // 8 warps, 1 CTA/SM  (long K)
#define GEMM_TILE_128x64 1
#define GEMM_TILE_64x64 2      // 4 warps, 2 CTAs/SM (short K)
#define GEMM_TILE_128x32 3     // 4 warps, 2 CTAs/SM (N <= 32)
#define GEMM_TILE_32x128 4     // 4 warps, 2 CTAs/SM (M <= 32)
#define GEMM_M3 8              // flag: 3-multiplication complex product
#define GEMM_BAND 16           // flag (64x64 + M3 only): honour the ZGEMM_A_BAND / ZGEMM_B_BAND tables of the descriptors
#define GEMM_SHORT_K 128
#define OP_N 0
#define OP_T 1
#define OP_H 2

namespace rcwa {

// ---- zgemm.cu
int gemm_tiles(int tile_cfg, int M, int N);
cudaError_t zgemm_grouped(int tile_cfg, int opa, int opb, const ZGemmProblem* probs, int nprob, int max_tiles,
                          cplx alpha, cplx beta, cudaStream_t st);
bool gemm_cfg_supports(int tile_cfg, int opa, int opb);
int gemm_pick_cfg(int opa, int opb, int M, int N, int K);
void gemm_set_tuning(int key, int value);
int gemm_get_tuning(int key);
cudaError_t zgemm_strided_cfg(int cfg, int opa, int opb, int M, int N, int K, cplx alpha, const cplx* A, int lda, long long sa,
                              const cplx* B, int ldb, long long sb, cplx beta, cplx* C, int ldc, long long sc,
                              int batch, ZGemmProblem* scratch, cudaStream_t st, int flags = 0);
// `scratch`: device array of >= batch ZGemmProblem
cudaError_t zgemm_strided(int opa, int opb, int M, int N, int K, cplx alpha, const cplx* A, int lda, long long sa,
                          const cplx* B, int ldb, long long sb, cplx beta, cplx* C, int ldc, long long sc,
                          int batch, ZGemmProblem* scratch, cudaStream_t st, int flags = 0);   // flags: ZGEMM_* structure hints


// ---- tc_gemm.cu: complex fp64-grade GEMM from int8 digit products on tcgen05 (see the file header)
#define TC_MAXS 8            // most digits (slices) per number
#define TC_MAXOPS 128
#define TC_MAXLOADS 16       // loads / MMA ops of one level group per K chunk
#define TC_MAXMMAS 32
#define TC_OP_LOAD_A 0u      // op word: bits 0-1 type; loads: bits 2-5 digit index;
#define TC_OP_LOAD_B 1u      //   MMA: bits 2-6 / 7-11 = index of the A / B load inside this (group, K chunk) iteration, bits 12-13 level - d0,
#define TC_OP_MMA 2u         //   bit 14 first MMA of that level in the iteration, bit 15 / 16 = last use of the A / B slot (release it)
#define TC_E_ZERO (-100000)  // exponent marker of an all-zero row / column
struct TcGroup { int d0, nl, op0, nops, nloads, nmma; };
// The schedule in two forms: `ops` (loads and MMAs interleaved in issue order; what the tests reason about) and the
// compact per-role tables the kernel walks:  loads[g][i] = bit 0: B operand, bits 1-4: digit;
// mmas[g][i] = bits 0-4 / 5-9: A / B load index, 10-11 level, 12 first, 13 / 14 release A / B, 15-18 loads to acquire first.
struct TcSchedule { int s, ngroups, nops, pad; TcGroup g[TC_MAXS]; unsigned ops[TC_MAXOPS]; };
// steps[g][p] (two words): one A digit plane with ALL the B planes it multiplies in this group (<= 4 pairs, 16 MMAs):
//   word 0 = bits 0-4 A load index, 5-7 pairs, 8-11 loads to acquire first, 12-16 (1 + index of the B load released after the step, 0 = none)
//   word 1 = per pair j, byte j: bits 0-4 B load index, 5-6 level, 7 first MMA of that level in the iteration
struct TcTables { int s, ngroups; int d0[TC_MAXS], nl[TC_MAXS], nloads[TC_MAXS], nmma[TC_MAXS], nsteps[TC_MAXS];
                  unsigned loads[TC_MAXS * TC_MAXLOADS]; unsigned mmas[TC_MAXS * TC_MAXMMAS]; unsigned steps[TC_MAXS * TC_MAXS * 2]; };
void tc_compact_schedule(const TcSchedule* sch, TcTables* tab);
#define TC_RING 12           // ring slots of 16 KB (one digit plane tile: 128 rows x 128 bytes of K)
// Issue-table entry of one step for a given ring position `bs` of the iteration's first load (16 words, what the issuer
// thread reads): [0] ring slot of the A plane, [1] MMA groups | loads to acquire << 4 | A slot << 8 | (B released ? 1 : 0) << 16
// | its slot << 17; per MMA group j (<= 4): [2+2j] ring slot of its (first) B plane, [3+2j] accumulator column | first << 9 |
// double << 10.  A "double" multiplies the A plane with TWO B planes that sit in adjacent ring slots (digits q and q-1)
// in one N = 256 MMA: the accumulator columns of level l and l-1 are adjacent in TMEM (level l of a group of nl levels
// lives at column (nl-1-l)*128).  Host and device share this function (tests/test_tc_gemm.py checks it on the CPU).
inline
#ifdef __CUDACC__
__host__ __device__
#endif
void tc_issue_entry(const TcTables& t, int g, int bs, int st, unsigned* e) {
    const unsigned w0 = t.steps[(g * TC_MAXS + st) * 2], w1 = t.steps[(g * TC_MAXS + st) * 2 + 1];
    const unsigned sa = (bs + (w0 & 31u)) % TC_RING, np = (w0 >> 5) & 7u, relb = (w0 >> 12) & 31u;
    const unsigned nl = (unsigned)t.nl[g];
    unsigned ng = 0;
    for (unsigned j = 0; j < 8; ++j) e[2 + j] = 0;
    for (unsigned j = 0; j < np;) {
        const unsigned pj = (w1 >> (8 * j)) & 255u;
        const unsigned sb = (bs + (pj & 31u)) % TC_RING, lvl = (pj >> 5) & 3u, first = (pj >> 7) & 1u;
        unsigned dbl = 0;
        if (j + 1 < np) {
            const unsigned pk = (w1 >> (8 * (j + 1))) & 255u;
            const unsigned sb2 = (bs + (pk & 31u)) % TC_RING, lvl2 = (pk >> 5) & 3u, first2 = (pk >> 7) & 1u;
            if (sb + 1 < TC_RING && sb2 == sb + 1 && lvl >= 1 && lvl2 == lvl - 1 && first2 == first) dbl = 1;
        }
        e[2 + 2 * ng] = sb;
        e[3 + 2 * ng] = ((nl - 1 - lvl) * 128u) | (first << 9) | (dbl << 10);
        ++ng;
        j += 1 + dbl;
    }
    e[0] = sa;
    e[1] = ng | (((w0 >> 8) & 15u) << 4) | (sa << 8) | (relb ? (1u << 16) | (((bs + relb - 1u) % TC_RING) << 17) : 0u);
}
void tc_build_schedule(int s, int nl, TcSchedule* sch);
size_t tc_workspace_bytes(int M, int N, int K, int nb, int s);      // split storage for the whole batch
size_t tc_workspace_min_bytes(int M, int N, int K, int s);          // ... for one matrix (the routine then runs in chunks)
bool tc_supported(int s, int M, int N, int K);
// C = alpha op(A) op(B) + beta C with s digits per number (alpha real); ws >= tc_workspace_min_bytes
cudaError_t tc_zgemm_strided(int s, int opa, int opb, int M, int N, int K, double alpha, const cplx* A, int lda, long long sa,
                             const cplx* B, int ldb, long long sb, cplx beta, cplx* C, int ldc, long long sc, int batch,
                             char* ws, size_t ws_bytes, cudaStream_t st);
cudaError_t tc_split_debug(const cplx* X, int ld, long long stride, int rows_contiguous, int R, int Kc, int s, int conj,
                           signed char* out, int* ex, int nmat, cudaStream_t st);

// Which GEMM engine a composite routine uses for its dense products: slices = 0 -> fp64 DMMA (zgemm.cu) everywhere;
// 2..8 -> products large enough to pay for the digit split go through the tcgen05 kernel with that many digits.
struct TcCtx { int slices; char* ws; size_t ws_bytes; };
// alpha real; falls back to the DMMA kernel for small shapes or when tc is null / off
cudaError_t gemm_auto(const TcCtx* tc, int opa, int opb, int M, int N, int K, double alpha, const cplx* A, int lda, long long sa,
                      const cplx* B, int ldb, long long sb, cplx beta, cplx* C, int ldc, long long sc, int batch,
                      ZGemmProblem* gscratch, cudaStream_t st);
size_t tc_ctx_bytes(int n, int nb, int slices);     // digit workspace a composite routine reserves for n x n x n products of a batch of nb

// ---- convmat.cu
size_t convmat_workspace_elems(int nx, int ny, int nb, int ox, int oy);
cudaError_t convmat(const void* grid, int grid_type, long long grid_stride, int nx, int ny, int nb, int ox, int oy,
                    cplx* E, cplx* ws, cudaStream_t st);

// ---- assemble.cu
cudaError_t pq_assemble(const cplx* eta, const cplx* E, const cplx* Mc, const cplx* nu, const cplx* mu_s,
                        const cplx* kx, const cplx* ky, int nb, int N, cplx* P, cplx* Q, cudaStream_t st);
cudaError_t kz_branch(const cplx* lam, cplx* kz, size_t total, cudaStream_t st);
cudaError_t eig_backward_combine(const cplx* lam, const cplx* glam, const cplx* T, double delta, int nb, int n, cplx* M, cudaStream_t st);
cudaError_t conj_transpose(const cplx* A, int nb, int n, cplx* Bm, cudaStream_t st);
cudaError_t layer_form(const cplx* W, const cplx* QW, const cplx* kz, const cplx* vfinv, const double* omega,
                       const double* thick, int nb, int N, cplx* Mp, cplx* Mm, cplx* Rp, cplx* Rm, cudaStream_t st);
cudaError_t layer_finish(const cplx* Tp, const cplx* Tm, int nb, int n, cplx* S11, cplx* S21, cudaStream_t st);
cudaError_t blockdiag_dense(const cplx* d4, int nb, int N, cplx* D, cudaStream_t st);
cudaError_t sym_project(const cplx* X, int nb, int n, const int* il, const cplx* cl, const int* ir, const cplx* cr,
                        int G, int nkl, int nkr, cplx* out, cudaStream_t st);
cudaError_t bd_left_mul(const cplx* d4, const cplx* X, int nb, int N, int ncols, cplx alpha, cplx beta, cplx* Y, cudaStream_t st);
cudaError_t bd_right_mul(const cplx* d4, const cplx* X, int nb, int N, int nrows, cplx alpha, cplx beta, cplx* Y, cudaStream_t st);
cudaError_t bd_add(const cplx* d4, int nb, int N, cplx alpha, cplx* D, cudaStream_t st);
cudaError_t set_identity(cplx* A, int n, int lda, long long stride, int nb, cudaStream_t st);
cudaError_t axpby(cplx alpha, const cplx* X, cplx beta, cplx* Y, size_t total, cudaStream_t st);

// ---- lu.cu
// tc_slices >= 2: the factorisation prepares 512-wide inverted diagonal blocks for the tcgen05 solve (else 128-wide, DMMA)
size_t lu_tinv_elems(int n, int nb, int tc_slices = 0);     // complex elements of the inverted-diagonal-block buffer of lu_factor
cudaError_t lu_factor(cplx* A, long long stride, int n, int lda, int nb, int* ipiv, int* perm, int* info, cplx* tinv,
                      ZGemmProblem* gscratch, cudaStream_t st, bool clear_info = true, int tc_slices = 0);
// Yw: work buffer shaped like X; tc (optional) must carry the same slices value the factorisation was given
cudaError_t lu_solve_right(const cplx* LU, long long lustride, int n, int lda, const int* perm, const cplx* tinv,
                           const cplx* Bm, long long bstride, int ldb, int nrows, cplx* X, long long xstride, int ldx,
                           cplx* Yw, int nb, ZGemmProblem* gscratch, cudaStream_t st, const TcCtx* tc = nullptr);

// ---- hess.cu
size_t hessenberg_workspace_bytes(int n, int nb);
// after_first_columns (optional): recorded on `st` once the column phase of the first panel has been enqueued
cudaError_t hessenberg_blocked(cplx* A, int n, int nb, cplx* Z, char* ws, cudaStream_t st, cudaEvent_t after_first_columns = nullptr);
cudaError_t hessenberg_matvec_probe(const cplx* A, int n, int nb, int j, char* ws, cudaStream_t st);
int hessenberg_panel_width();

// ---- eig.cu
size_t eig_workspace_bytes(int n, int nb);
cudaError_t eig_stats(const char* ws, int n, int nb, int* out, cudaStream_t st);
cudaError_t eig_profile(const char* ws, int n, int nb, long long* out, cudaStream_t st);
cudaError_t hessenberg(cplx* A, int n, int nb, cplx* Zout, char* ws, size_t ws_bytes, cudaStream_t st);
cudaError_t eig_matvec_probe(const cplx* A, int n, int nb, int j, char* ws, size_t ws_bytes, cudaStream_t st);
// phases: 1 = Hessenberg reduction only (state stays in A / ws), 2 = QR iteration + eigenvectors of a reduced problem, 3 = both
cudaError_t eig(cplx* A, int n, int nb, cplx* w, cplx* V, char* ws, size_t ws_bytes, int* info, volatile int* host_flag, cudaStream_t st, int phases = 3);

}  // namespace rcwa
