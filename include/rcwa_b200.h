/* rcwa_b200.h -- C ABI of the B200-native RCWA hot path (librcwa_b200.so).
 *
 * This is the drop-in boundary.  The reference (kch3782/torcwa) has NO native
 * boundary: every step below is a chain of ATen calls issued from torcwa/rcwa.py.  Each entry point
 * names the reference code it replaces (file:line under /root/reference).  INTEGRATION.md shows the
 * ctypes stub a torcwa maintainer would add.
 *
 * Conventions
 *   - every matrix pointer is a DEVICE pointer to row-major, interleaved (re,im) fp64 ("c128")
 *     data, 16-byte aligned; batched tensors are [nb, rows, cols] contiguous unless a stride is given;
 *   - the caller allocates everything (outputs, workspaces, info); the library never allocates or frees device
 *     memory.  Every entry point EXCEPT rcwa_eig / rcwa_eig_phases only enqueues work on `stream` (a cudaStream_t
 *     passed as void*), never synchronises, and is CUDA-graph capturable;
 *   - rcwa_eig is the exception, and says so: its QR phase forks onto internal streams (created once per host thread
 *     and device on first use, reused by later calls, joined back into `stream` before it returns -- also on error
 *     paths) and, to stop as soon as every matrix has converged, the HOST polls a counter through the caller's pinned
 *     `host_flag` with cudaEventSynchronize.  It therefore blocks the calling host thread for the duration of the QR
 *     phase and cannot be captured into a CUDA graph.  It is re-entrant: concurrent calls from different host threads
 *     (on their own streams and workspaces) share nothing (tests/test_gpu_eig.py);
 *   - process-wide state: the tuning table behind rcwa_set_tuning (defaults = the measured best; results do not
 *     depend on it) and per-device "attribute already set" flags of the kernels.  Nothing else;
 *   - return value: 0 ok; -k = argument k invalid (LAPACK style); <= -1000 = CUDA runtime error
 *     (-1000 - cudaError_t);
 *   - numerical status is reported per batch entry in device int32 info[nb]
 *     (0 ok; >0 = zero pivot / unconverged eigenvalue index), never by the return value.
 */
#ifndef RCWA_B200_H
#define RCWA_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RCWA_B200_ABI_VERSION 2

/* grid element types for rcwa_convmat */
#define RCWA_GRID_F32 0
#define RCWA_GRID_F64 1
#define RCWA_GRID_C64 2
#define RCWA_GRID_C128 3
/* GEMM operand ops */
#define RCWA_OP_N 0
#define RCWA_OP_T 1
#define RCWA_OP_H 2

int rcwa_b200_abi_version(void);

/* bytes of the grouped-GEMM descriptor scratch every dense routine needs for a batch of nb */
size_t rcwa_gemm_scratch_bytes(int nb);

/* ---- stage 1: Fourier factorisation --------------------------------------------------------
 * E[b] (N x N, N=(2ox+1)(2oy+1)) = Toeplitz matrix of the 2-D Fourier coefficients of grid[b]
 * (nx x ny samples, element type grid_type; grid_stride = elements between consecutive grids,
 * 0 = one grid shared by the whole batch).  Requires nx >= 4ox+1, ny >= 4oy+1.
 * Replaces rcwa._material_conv, torcwa/rcwa.py:1183-1204 (fft2 + index gather). */
size_t rcwa_convmat_workspace_bytes(int nx, int ny, int nb, int ox, int oy);
int rcwa_convmat(const void* grid, int grid_type, long long grid_stride, int nx, int ny, int nb,
                 int ox, int oy, void* E, void* ws, void* stream);

/* ---- dense building blocks -----------------------------------------------------------------
 * C[b] = alpha op(A[b]) op(B[b]) + beta C[b]; strides in elements.
 * Replaces torch.matmul call sites of the path (rcwa.py:1228,1232,1236,1264,1276-1281,1291-1294). */
int rcwa_zgemm_batched(int opa, int opb, int M, int N, int K, double alpha_re, double alpha_im,
                       const void* A, int lda, long long stride_a, const void* B, int ldb, long long stride_b,
                       double beta_re, double beta_im, void* C, int ldc, long long stride_c,
                       int nb, void* gemm_scratch, void* stream);

/* The same product on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM, operands staged by
 * TMA): every number is split into `slices` (2..8) signed 8-bit digits after a per-row (A) / per-column (B) power-of-two
 * scaling, the digit products are exact int8 x int8 -> int32 GEMMs, and the result is re-assembled in fp64 -- error
 * relative to (row scale of A) x (column scale of B), measured on random matrices: slices 4: 4e-8, 5: 2e-10, 6: 8e-13,
 * 7: 3e-15, 8: 3e-16 (fp64-grade); 4-5 are enough for the complex64 API.  alpha is real.  `ws`: rcwa_zgemm_tc_workspace_bytes(M,N,K,nb,
 * slices) for one pass over the batch; any size down to rcwa_zgemm_tc_workspace_bytes(M,N,K,1,slices) works (chunked).
 * Replaces the same torch.matmul call sites (rcwa.py:1236,1264,1276-1281,1291-1294). */
size_t rcwa_zgemm_tc_workspace_bytes(int M, int N, int K, int nb, int slices);
int rcwa_zgemm_tc_batched(int slices, int opa, int opb, int M, int N, int K, double alpha,
                          const void* A, int lda, long long stride_a, const void* B, int ldb, long long stride_b,
                          double beta_re, double beta_im, void* C, int ldc, long long stride_c,
                          int nb, void* ws, size_t ws_bytes, void* stream);
/* Test / diagnostic entry points of that path.  rcwa_tc_split: the digit split on its own -- X [nb] matrices (ld, stride);
 * rows_contiguous = 1: the R scaled vectors are the rows of X (Kc entries each), 0: its columns (X is Kc x R);
 * planes: int8 [nb, 3 (re, im, re+im), slices, R, Kp], Kp = Kc rounded up to 128, digit 0 most significant;
 * ex: int32 [nb, R] with value = 2^(ex + 2 - 8 slices) * sum_d digit_d 256^(slices-1-d).
 * rcwa_tc_schedule (host only): the load / MMA / release schedule of one K chunk; ops [128] words (issue order), meta [64] =
 * {groups, ops, then per group: first level, levels, first op, ops, loads, MMAs}; loads [8*16] and mmas [8*32] (optional):
 * the compact per-role tables the kernel walks (bit layout in csrc/kernels.h). */
int rcwa_tc_split(const void* X, int ld, long long stride, int rows_contiguous, int R, int Kc, int slices, int conj,
                  void* planes, int* ex, int nb, void* stream);
int rcwa_tc_schedule(int slices, int levels, unsigned* ops, int* meta, unsigned* loads, unsigned* mmas);
/* (host only) the issue-table entry (16 words, layout in csrc/kernels.h: tc_issue_entry) the kernel's MMA issuer reads for
 * step `step` of level group `group` when the iteration's first load sits in ring slot `ring_pos`. */
int rcwa_tc_issue_entry(int slices, int levels, int group, int ring_pos, int step, unsigned* entry, int* nsteps);

/* Tuning / profiling entry points (not needed by a binding; used by bench.py and tools/).
 * rcwa_zgemm_batched_cfg: the same product on an explicit kernel configuration:
 *   cfg = tile | 8*m3;  tile 0: 64x128, 1: 128x64 (256 threads, one CTA/SM, long K), 2: 64x64, 3: 128x32,
 *   4: 32x128 (128 threads, two CTAs/SM: short K or one skinny dimension); m3 = 3-multiplication complex
 *   product (25 % fewer fp64 tensor instructions, norm-wise accuracy); cfg = -1: automatic choice.
 *   Tiles 2-4 and m3 exist for the op pairs (N,N), (N,H), (H,N) only.
 * rcwa_set_tuning(key, value): process-wide knobs of the automatic choice -- key 0: use m3 (default 1);
 *   1 / 2: tile of the QR row / column updates; 3: use the 128-thread tiles (default 1); 4: the QR pass
 *   kernel claims a whole SM per matrix (default 0); 5-7: count limits of the serial QR slices (Schur
 *   rotations, AED swaps, AED restore steps); 8: time budget of a serial QR slice in us (default 90);
 *   9: number of independently pipelined matrix groups of the QR phase (default 2; one per 148 matrices, at most 4,
 *   for batches above 296); 10: skip the zero
 *   k groups of the banded window unitaries in the QR update GEMMs (default 0, measured no gain);
 *   11: Hessenberg phase as two staggered half batches (default 0, measured slower); 12: triangular solves of the
 *   S-matrix stage on the tcgen05 engine too when gemm_slices >= 2 (default 0: measured slower at K = 512); 13: QR pass
 *   as two launches per iteration -- bulge-chase windows, then the small dense solves at two CTAs per SM (0 = automatic:
 *   batches above 296 matrices, 1 = never, 2 = always); 14 = 2: replay the QR loop from CUDA graphs of 8 iterations per
 *   matrix group (default off: measured no gain at the default group count); 15: aggressive-early-deflation window of the QR
 *   phase (0 = automatic: 24 for n <= 1024, else 32).  Call before
 *   asking for workspace sizes and enqueuing work; the numerical contract does not depend on them. */
int rcwa_zgemm_batched_cfg(int cfg, int opa, int opb, int M, int N, int K, double alpha_re, double alpha_im,
                           const void* A, int lda, long long stride_a, const void* B, int ldb, long long stride_b,
                           double beta_re, double beta_im, void* C, int ldc, long long stride_c,
                           int nb, void* gemm_scratch, void* stream);
int rcwa_set_tuning(int key, int value);
int rcwa_get_tuning(int key);

/* In-place LU with column pivoting for right-solves X*A = B (A*Pi = L*U, L lower, U unit upper).
 * ipiv, perm: int32 [nb,n]; info int32 [nb]; tinv: rcwa_lu_tinv_bytes(n, nb) bytes receiving the
 * inverses of the 128 x 128 diagonal blocks of L and U (they turn the triangular solves into GEMMs).
 * Replaces torch.linalg.inv call sites (rcwa.py:1226,1230,1248,1266-1267,1271,1273,1287-1288). */
size_t rcwa_lu_tinv_bytes(int n, int nb);
int rcwa_lu_factor(void* A, long long stride, int n, int lda, int nb, int* ipiv, int* perm, int* info,
                   void* tinv, void* gemm_scratch, void* stream);
/* X[b] (nrows x n) = B[b] * A[b]^-1 given rcwa_lu_factor output; X must not alias B; `work` is a
 * buffer shaped and strided like X. */
int rcwa_lu_solve_right(const void* LU, long long lu_stride, int n, int lda, const int* perm, const void* tinv,
                        const void* B, long long b_stride, int ldb, int nrows,
                        void* X, long long x_stride, int ldx, void* work, int nb, void* gemm_scratch, void* stream);

/* ---- stage 1 -> 2: P, Q of the layer eigenproblem -----------------------------------------
 * eta = E^-1, optional Mc (mu conv. matrix) and nu = Mc^-1 (both NULL => homogeneous mu given per
 * batch entry in mu_scalar[nb]); kx, ky: [nb,N] normalised wavevectors.  P, Q: [nb,2N,2N].
 * Replaces the assembly in rcwa._eigen_decomposition, rcwa.py:1224-1232. */
int rcwa_pq_assemble(const void* eta, const void* E, const void* Mc, const void* nu, const void* mu_scalar,
                     const void* kx, const void* ky, int nb, int N, void* P, void* Q, void* stream);

/* ---- stage 2: non-Hermitian eigendecomposition --------------------------------------------
 * A[b] (n x n, destroyed) -> eigenvalues w[b] (n) and right eigenvectors V[b] (n x n, columns,
 * unit 2-norm, arbitrary order) with A V = V diag(w).  Householder Hessenberg reduction, windowed
 * multishift QR with Schur-vector accumulation, blocked triangular eigenvector solve,
 * back-transformation; everything on the device.  host_flag: 64 bytes of PINNED host memory the
 * routine uses to poll convergence without a device-wide synchronisation (may be NULL: then the
 * routine runs its full sweep budget).
 * Replaces torch.linalg.eig in Eig.forward, torcwa/torch_eig.py:11-17 (called at rcwa.py:1236/1238). */
size_t rcwa_eig_workspace_bytes(int n, int nb);
int rcwa_eig(void* A, int n, int nb, void* w, void* V, void* ws, size_t ws_bytes, int* info,
             void* host_flag, void* stream);
/* The same routine in two calls on the same arguments: phases = 1: Hessenberg reduction only (only enqueues; the reduced
 * problem stays in A and ws), phases = 2: QR iteration and eigenvectors of a problem reduced by a phases = 1 call,
 * phases = 3: both (= rcwa_eig).  Lets a host pipeline sub-batches: the HBM-bound reduction of one sub-batch runs under
 * the latency-bound QR iteration of another (torcwa_b200/rcwa.py). */
int rcwa_eig_phases(void* A, int n, int nb, void* w, void* V, void* ws, size_t ws_bytes, int* info,
                    void* host_flag, int phases, void* stream);
/* Diagnostics of the last rcwa_eig that used workspace `ws`: out[4*b..] = {QR sweeps, window passes,
 * AED windows, info} of matrix b (device int32 [nb,4]). */
int rcwa_eig_stats(const void* ws, int n, int nb, int* out, void* stream);
/* Profile of the QR phase of the last rcwa_eig on `ws`: out = device int64 [nb,6,3] = {launch count, SM cycles,
 * longest single segment in cycles} per pass segment (0 sweep start: deflation scan + shifts, 1 bulge-chain window, 2 small-block slice,
 * 3 AED Schur slice, 4 AED deflation-scan slice, 5 AED finish). */
int rcwa_eig_profile(const void* ws, int n, int nb, long long* out, void* stream);
/* First phase of rcwa_eig on its own (profiling / building block): A[b] -> H[b] upper Hessenberg in
 * place, Z[b] unitary with A_in = Z H Z^H.  Workspace as for rcwa_eig.  This is the HBM-bound
 * streaming kernel of the eigen stage (one fused pass over [A; Z] per column). */
int rcwa_hessenberg(void* A, int n, int nb, void* Z, void* ws, size_t ws_bytes, void* stream);
/* Profiling entry: ONE launch of the streaming mat-vec kernel of rcwa_hessenberg for column j (0 <= j <= n-3)
 * on A[nb,n,n] (read only; the vector it multiplies with is whatever the workspace holds).  It reads the
 * (n - k0 - 1) x (n - j - 1) trailing block of every matrix once, k0 = j rounded down to the panel width
 * (rcwa_hessenberg_panel_width()): the algorithmic bytes of that launch are 16 (n-k0-1)(n-j-1) nb. */
int rcwa_hessenberg_matvec_probe(const void* A, int n, int nb, int j, void* ws, size_t ws_bytes, void* stream);
int rcwa_hessenberg_panel_width(void);
/* kz = sqrt(lambda), negated where Im < 0 (rcwa.py:1240-1241). total = nb*n elements. */
int rcwa_kz_branch(const void* lam, void* kz, long long total, void* stream);

/* ---- stage 2, reverse mode: gradient of the eigendecomposition ------------------------------
 * grad[b] (n x n) = X^-H (diag(g_lambda) + conj(F) o (X^H g_X)) X^H with F_ij = conj(s_ij) / (|s_ij|^2 + delta),
 * s_ij = lambda_j - lambda_i, F_ii = 0 (Lorentzian broadening, delta = torcwa.Eig.broadening_parameter).
 * lam [nb,n], X [nb,n,n] = the outputs of rcwa_eig; glam [nb,n] and gX [nb,n,n] the incoming gradients (either
 * may be NULL = zero).  Two GEMMs, one LU of X, one right-solve; info[nb] reports a singular X.
 * Replaces Eig.backward, torcwa/torch_eig.py:19-44 (3 matmul + 1 inverse). */
size_t rcwa_eig_backward_workspace_bytes(int n, int nb);
int rcwa_eig_backward(const void* lam, const void* X, const void* glam, const void* gX, double delta, int nb, int n,
                      void* grad, void* ws, int* info, void* stream);

/* ---- stage 3a: layer S-matrix -------------------------------------------------------------
 * Inputs: eigenvectors W [nb,n,n], kz [nb,n], Q [nb,n,n], vfinv [nb,4,N] = the four diagonals of
 * Vf^-1 (free-space E->H matrix, rcwa.py:1143-1147), omega[nb], thickness[nb] (fp64).
 * Outputs S11 (= S22) and S21 (= S12) of the single layer, [nb,n,n].
 * Minimal algebra: V = Q W Kz^-1, two LU right-solves; replaces
 * rcwa._solve_layer_smatrix, rcwa.py:1244-1281 (dense inv of the 4N x 4N coupling matrix).
 * gemm_slices (here and in the star products): 0 = every dense product on the fp64 tensor pipe (DMMA; the complex128
 * contract); 2..8 = the n x n x n products (and, with tuning key 12, the K = 512 block updates of the triangular solves) run
 * on the tcgen05 int8-digit GEMM with that many digits when n >= 768 -- below that the DMMA kernel is faster and is used (see rcwa_zgemm_tc_batched; the Python host uses 7 for complex64 simulations:
 * the stage amplifies a product's error by up to ~1e6 into the far-evanescent S-block entries).  The
 * workspace size depends on it. */
size_t rcwa_layer_smatrix_workspace_bytes(int N, int nb, int gemm_slices);
int rcwa_layer_smatrix(const void* W, const void* kz, const void* Q, const void* vfinv,
                       const double* omega, const double* thickness, int nb, int N,
                       void* S11, void* S21, void* ws, int* info, int gemm_slices, void* stream);

/* ---- stage 3b: Redheffer star product -----------------------------------------------------
 * out = Sm (*) Sn, each S = {S11,S21,S12,S22} of [nb,n,n]; outputs must not alias inputs.
 * One LU + two right-solves + 8 GEMMs; replaces rcwa._RS_prod, rcwa.py:1283-1294. */
size_t rcwa_redheffer_workspace_bytes(int n, int nb, int gemm_slices);
int rcwa_redheffer(const void* const Sm[4], const void* const Sn[4], void* const out[4],
                   int nb, int n, void* ws, int* info, int gemm_slices, void* stream);

/* Same product when the LEFT factor is a half-space / homogeneous-layer S-matrix, i.e. each of its four
 * blocks is itself four diagonals: Sm_bd[k] = [nb,4,N] (order 11,12,21,22 inside each block; the
 * reference builds these densely, rcwa.py:1157-1164).  Six of the eight GEMMs become O(n^2) row/column
 * combinations.  n = 2N; workspace as rcwa_redheffer. */
int rcwa_redheffer_bdleft(const void* const Sm_bd[4], const void* const Sn[4], void* const out[4],
                          int nb, int N, void* ws, int* info, int gemm_slices, void* stream);

/* dense [nb,2N,2N] from four diagonals d4 [nb,4,N] (order 11,12,21,22): half-space and
 * homogeneous-layer blocks (rcwa.py:1157-1181, :1206-1222). */
int rcwa_blockdiag_dense(const void* d4, int nb, int N, void* D, void* stream);

/* Symmetry-adapted block of a dense operator: out [nb,nkl,nkr] = T_L^H X T_R for X [nb,n,n], where column k of T_L is
 * sum_{t<G} cl[t][k] e_{il[t][k]} (il, cl: [G,nkl] row-major, int32 / complex128; likewise ir, cr [G,nkr]; G <= 4; unused
 * slots carry a zero coefficient and any valid index).  The bases are the joint eigenvectors of the mirror / C2 operators
 * that commute with P and Q of a symmetric cell (torcwa_b200/symmetry.py); the reference has no counterpart -- it always
 * solves the full n x n problem (rcwa.py:1224-1247). */
int rcwa_sym_project(const void* X, int nb, int n, const int* il, const void* cl, const int* ir, const void* cr,
                     int G, int nkl, int nkr, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
