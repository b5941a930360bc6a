// Elementwise assembly kernels around the dense stages (all HBM-bound, O(n^2) per design point).
//
//  * pq_assemble      : P, Q of the layer eigenproblem from the convolution matrices
//                       (reference: rcwa._eigen_decomposition, torcwa/rcwa.py:1224-1232).  The
//                       reference multiplies by dense diag(Kx), diag(Ky); those are row/column
//                       scalings, done here in one pass.
//  * kz_branch        : kz = sqrt(lambda), negated where Im < 0 (rcwa.py:1240-1241).
//  * layer_form       : operands of the two right-solves of the minimal layer S-matrix
//                       (reference: rcwa._solve_layer_smatrix, rcwa.py:1244-1281).
//  * layer_finish     : S11 = T+ + T-,  S21 = T+ - T- - I.
//  * blockdiag_dense  : scatter four diagonals into a dense 2N x 2N matrix (half-space and
//                       homogeneous-layer S blocks, rcwa.py:1157-1181 / :1206-1222).
//  * small utilities  : identity, axpby.
#include "common.cuh"
#include "kernels.h"

namespace {

// grid (ceil(N/256), N, B)
__global__ void pq_assemble_kernel(const cplx* __restrict__ eta, const cplx* __restrict__ E,
                                   const cplx* __restrict__ Mc, const cplx* __restrict__ nu,
                                   const cplx* __restrict__ mu_s,   // [B] scalar mu when Mc == nullptr
                                   const cplx* __restrict__ kx, const cplx* __restrict__ ky,   // [B,N]
                                   int N, cplx* __restrict__ P, cplx* __restrict__ Q) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= N) return;
    const int n = 2 * N;
    const size_t e = ((size_t)b * N + i) * N + j;
    const cplx kxi = kx[(size_t)b * N + i], kyi = ky[(size_t)b * N + i];
    const cplx kxj = kx[(size_t)b * N + j], kyj = ky[(size_t)b * N + j];
    const cplx et = eta[e], ep = E[e];
    cplx m, v;
    if (Mc) { m = Mc[e]; v = nu[e]; }
    else { cplx mu = mu_s[b]; m = (i == j) ? mu : C(0, 0); v = (i == j) ? cinv(mu) : C(0, 0); }
    const cplx xe = cmul(kxi, et), ye = cmul(kyi, et), xv = cmul(kxi, v), yv = cmul(kyi, v);
    cplx* p = P + (size_t)b * n * n;
    cplx* q = Q + (size_t)b * n * n;
    const size_t r0 = (size_t)i * n + j, r1 = (size_t)(i + N) * n + j;
    p[r0] = cmul(xe, kyj);                    // Kx eta Ky
    p[r0 + N] = csub(m, cmul(xe, kxj));       // M - Kx eta Kx
    p[r1] = csub(cmul(ye, kyj), m);           // Ky eta Ky - M
    p[r1 + N] = cneg(cmul(ye, kxj));          // -Ky eta Kx
    q[r0] = cneg(cmul(xv, kyj));              // -Kx nu Ky
    q[r0 + N] = csub(cmul(xv, kxj), ep);      // Kx nu Kx - E
    q[r1] = csub(ep, cmul(yv, kyj));          // E - Ky nu Ky
    q[r1 + N] = cmul(yv, kxj);                // Ky nu Kx
}

// Eig.backward, elementwise part (torcwa/torch_eig.py:24-38):  M = diag(g_lambda) + conj(F) o T,
//   s_ij = lambda_j - lambda_i,  F_ij = conj(s_ij) / (|s_ij|^2 + delta),  F_ii = 0   =>   conj(F_ij) = s_ij / (|s_ij|^2 + delta).
// grid (ceil(n/256), n, B).  glam may be nullptr (zero eigenvalue gradient); T may be nullptr (zero eigenvector gradient).
__global__ void eig_backward_combine_kernel(const cplx* __restrict__ lam, const cplx* __restrict__ glam, const cplx* __restrict__ T,
                                            double delta, int n, cplx* __restrict__ M) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= n) return;
    const size_t e = ((size_t)b * n + i) * n + j;
    cplx out;
    if (i == j) out = glam ? glam[(size_t)b * n + i] : C(0, 0);
    else if (!T) out = C(0, 0);
    else {
        const cplx sij = csub(lam[(size_t)b * n + j], lam[(size_t)b * n + i]);
        const double den = cabs_(sij) * cabs_(sij) + delta;        // torch.abs(s)**2 + delta, as the reference rounds it
        out = cmul(C(sij.x / den, sij.y / den), T[e]);
    }
    M[e] = out;
}

// B[b][j][i] = conj(A[b][i][j])   (tile transpose through shared memory), grid (ceil(n/32), ceil(n/32), B), block (32, 8)
__global__ void conj_transpose_kernel(const cplx* __restrict__ A, int n, cplx* __restrict__ Bm) {
    __shared__ cplx tile[32][33];
    const int b = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const cplx* a = A + (size_t)b * n * n;
    cplx* o = Bm + (size_t)b * n * n;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = y0 + r, j = x0 + threadIdx.x;
        if (i < n && j < n) tile[r][threadIdx.x] = a[(size_t)i * n + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = x0 + r, j = y0 + threadIdx.x;          // output row = input column
        if (i < n && j < n) o[(size_t)i * n + j] = cconj(tile[threadIdx.x][r]);
    }
}

__global__ void kz_branch_kernel(const cplx* __restrict__ lam, cplx* __restrict__ kz, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    cplx r = csqrt_(lam[i]);
    if (r.y < 0.0) r = cneg(r);
    kz[i] = r;
}

// grid (ceil(n/256), N, B); thread handles column j of rows i and i+N
__global__ void layer_form_kernel(const cplx* __restrict__ W, const cplx* __restrict__ QW, const cplx* __restrict__ kz,
                                  const cplx* __restrict__ vfinv,          // [B,4,N]: Vf^-1 diagonals (11,12,21,22)
                                  const double* __restrict__ omega, const double* __restrict__ thick,   // [B]
                                  int N, cplx* __restrict__ Mp, cplx* __restrict__ Mm, cplx* __restrict__ Rp, cplx* __restrict__ Rm) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    const int n = 2 * N;
    if (j >= n) return;
    const cplx kzj = kz[(size_t)b * n + j];
    // X_j = exp(i * omega * kz_j * d)
    const double od = omega[b] * thick[b];
    const double mag = exp(-od * kzj.y);
    double s, c;
    sincos(od * kzj.x, &s, &c);
    const cplx X = C(mag * c, mag * s);
    const cplx onep = C(1.0 + X.x, X.y), onem = C(1.0 - X.x, -X.y);
    const cplx ikz = cinv(kzj);
    const size_t base = (size_t)b * n * n;
    const size_t e0 = base + (size_t)i * n + j, e1 = base + (size_t)(i + N) * n + j;
    const cplx w0 = W[e0], w1 = W[e1];
    const cplx v0 = cmul(QW[e0], ikz), v1 = cmul(QW[e1], ikz);      // V = Q W Kz^-1
    const cplx* vi = vfinv + (size_t)b * 4 * N;
    const cplx b0 = cadd(cmul(vi[i], v0), cmul(vi[N + i], v1));      // B = Vf^-1 V (per-order 2x2)
    const cplx b1 = cadd(cmul(vi[2 * N + i], v0), cmul(vi[3 * N + i], v1));
    const cplx rp0 = cmul(w0, onep), rp1 = cmul(w1, onep);
    const cplx wm0 = cmul(w0, onem), wm1 = cmul(w1, onem);
    Rp[e0] = rp0; Rp[e1] = rp1;
    Rm[e0] = cneg(wm0); Rm[e1] = cneg(wm1);
    Mp[e0] = cadd(rp0, cmul(b0, onem)); Mp[e1] = cadd(rp1, cmul(b1, onem));
    Mm[e0] = cadd(wm0, cmul(b0, onep)); Mm[e1] = cadd(wm1, cmul(b1, onep));
}

// grid (ceil(n/256), n, B)
__global__ void layer_finish_kernel(const cplx* __restrict__ Tp, const cplx* __restrict__ Tm, int n,
                                    cplx* __restrict__ S11, cplx* __restrict__ S21) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= n) return;
    const size_t e = ((size_t)b * n + i) * n + j;
    const cplx a = Tp[e], c = Tm[e];
    S11[e] = cadd(a, c);
    cplx d = csub(a, c);
    if (i == j) d.x -= 1.0;
    S21[e] = d;
}

// grid (ceil(n/256), n, B): D[b] = [[diag d0, diag d1],[diag d2, diag d3]]
__global__ void blockdiag_dense_kernel(const cplx* __restrict__ d4, int N, cplx* __restrict__ D) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    const int n = 2 * N;
    if (j >= n) return;
    cplx v = C(0, 0);
    const int ii = i % N, jj = j % N;
    if (ii == jj) v = d4[((size_t)b * 4 + (i / N) * 2 + (j / N)) * N + ii];
    D[((size_t)b * n + i) * n + j] = v;
}

// Y = alpha * BD * X + beta * Y   (BD = four diagonals [B,4,N]; X, Y dense [B,2N,ncols]); grid (ceil(ncols/256), N, B)
__global__ void bd_left_mul_kernel(const cplx* __restrict__ d4, const cplx* __restrict__ X, int N, int ncols, cplx alpha, cplx beta,
                                   cplx* __restrict__ Y) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= ncols) return;
    const cplx* d = d4 + (size_t)b * 4 * N;
    const size_t base = (size_t)b * 2 * N * ncols;
    const size_t e0 = base + (size_t)i * ncols + j, e1 = base + (size_t)(i + N) * ncols + j;
    const cplx x0 = X[e0], x1 = X[e1];
    cplx y0 = cmul(alpha, cadd(cmul(d[i], x0), cmul(d[N + i], x1)));
    cplx y1 = cmul(alpha, cadd(cmul(d[2 * N + i], x0), cmul(d[3 * N + i], x1)));
    if (!(beta.x == 0.0 && beta.y == 0.0)) { y0 = cadd(y0, cmul(beta, Y[e0])); y1 = cadd(y1, cmul(beta, Y[e1])); }
    Y[e0] = y0; Y[e1] = y1;
}

// Y = alpha * X * BD + beta * Y   (X, Y dense [B,nrows,2N]); grid (ceil(N/256), nrows, B)
__global__ void bd_right_mul_kernel(const cplx* __restrict__ d4, const cplx* __restrict__ X, int N, int nrows, cplx alpha, cplx beta,
                                    cplx* __restrict__ Y) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= N) return;
    const cplx* d = d4 + (size_t)b * 4 * N;
    const size_t e0 = ((size_t)b * nrows + i) * 2 * N + j, e1 = e0 + N;
    const cplx x0 = X[e0], x1 = X[e1];
    cplx y0 = cmul(alpha, cadd(cmul(x0, d[j]), cmul(x1, d[2 * N + j])));
    cplx y1 = cmul(alpha, cadd(cmul(x0, d[N + j]), cmul(x1, d[3 * N + j])));
    if (!(beta.x == 0.0 && beta.y == 0.0)) { y0 = cadd(y0, cmul(beta, Y[e0])); y1 = cadd(y1, cmul(beta, Y[e1])); }
    Y[e0] = y0; Y[e1] = y1;
}

// D += alpha * dense(BD)   (adds the four diagonals); grid (ceil(N/256), 1, B)
__global__ void bd_add_kernel(const cplx* __restrict__ d4, int N, cplx alpha, cplx* __restrict__ D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.z;
    if (i >= N) return;
    const int n = 2 * N;
    const cplx* d = d4 + (size_t)b * 4 * N;
    cplx* m = D + (size_t)b * n * n;
    m[(size_t)i * n + i] = cadd(m[(size_t)i * n + i], cmul(alpha, d[i]));
    m[(size_t)i * n + i + N] = cadd(m[(size_t)i * n + i + N], cmul(alpha, d[N + i]));
    m[(size_t)(i + N) * n + i] = cadd(m[(size_t)(i + N) * n + i], cmul(alpha, d[2 * N + i]));
    m[(size_t)(i + N) * n + i + N] = cadd(m[(size_t)(i + N) * n + i + N], cmul(alpha, d[3 * N + i]));
}

__global__ void identity_kernel(cplx* __restrict__ A, int n, int lda, long long stride) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= n) return;
    A[(size_t)b * stride + (size_t)i * lda + j] = C(i == j ? 1.0 : 0.0, 0.0);
}

// Y = alpha * X + beta * Y (flat)
__global__ void axpby_kernel(cplx alpha, const cplx* __restrict__ X, cplx beta, cplx* __restrict__ Y, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    Y[i] = cadd(cmul(alpha, X[i]), cmul(beta, Y[i]));
}

}  // namespace

namespace rcwa {

cudaError_t pq_assemble(const cplx* eta, const cplx* E, const cplx* Mc, const cplx* nu, const cplx* mu_s,
                        const cplx* kx, const cplx* ky, int nb, int N, cplx* P, cplx* Q, cudaStream_t st) {
    pq_assemble_kernel<<<dim3((N + 255) / 256, N, nb), 256, 0, st>>>(eta, E, Mc, nu, mu_s, kx, ky, N, P, Q);
    return cudaGetLastError();
}
cudaError_t eig_backward_combine(const cplx* lam, const cplx* glam, const cplx* T, double delta, int nb, int n, cplx* M, cudaStream_t st) {
    eig_backward_combine_kernel<<<dim3((n + 255) / 256, n, nb), 256, 0, st>>>(lam, glam, T, delta, n, M);
    return cudaGetLastError();
}
cudaError_t conj_transpose(const cplx* A, int nb, int n, cplx* Bm, cudaStream_t st) {
    conj_transpose_kernel<<<dim3((n + 31) / 32, (n + 31) / 32, nb), dim3(32, 8), 0, st>>>(A, n, Bm);
    return cudaGetLastError();
}
cudaError_t kz_branch(const cplx* lam, cplx* kz, size_t total, cudaStream_t st) {
    kz_branch_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(lam, kz, total);
    return cudaGetLastError();
}
// ---- symmetry-adapted blocks (torcwa_b200/symmetry.py): out = T_L^H X T_R for bases whose columns are combinations of at most
// G <= 4 unit vectors: column k of T_L is sum_t cl[t][k] e_{il[t][k]} (padding entries carry a zero coefficient).  One
// pass over the G*G gathered entries per output element; rows il[t][k] are read along l, so the loads coalesce wherever the
// orbit representatives ir[u][l] run contiguously (they do within one kx order).
__global__ void sym_project_kernel(const cplx* __restrict__ X, int n, const int* __restrict__ il, const cplx* __restrict__ cl,
                                   const int* __restrict__ ir, const cplx* __restrict__ cr, int G, int nkl, int nkr,
                                   cplx* __restrict__ out) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    const long long b = blockIdx.z;
    if (l >= nkr) return;
    const cplx* Xb = X + b * (long long)n * n;
    int jr[4];
    cplx wr[4];
    for (int u = 0; u < G; ++u) { jr[u] = ir[u * nkr + l]; wr[u] = cr[u * nkr + l]; }
    double sr = 0.0, si = 0.0;
    for (int t = 0; t < G; ++t) {
        const cplx w = cl[t * nkl + k];
        if (w.x == 0.0 && w.y == 0.0) continue;
        const cplx* row = Xb + (long long)il[t * nkl + k] * n;
        double ar = 0.0, ai = 0.0;
        for (int u = 0; u < G; ++u) {
            if (wr[u].x == 0.0 && wr[u].y == 0.0) continue;
            const cplx x = row[jr[u]];
            ar += x.x * wr[u].x - x.y * wr[u].y;
            ai += x.x * wr[u].y + x.y * wr[u].x;
        }
        sr += w.x * ar + w.y * ai;          // conj(w) * a
        si += w.x * ai - w.y * ar;
    }
    out[(b * nkl + k) * nkr + l] = make_double2(sr, si);
}

cudaError_t layer_form(const cplx* W, const cplx* QW, const cplx* kz, const cplx* vfinv, const double* omega,
                       const double* thick, int nb, int N, cplx* Mp, cplx* Mm, cplx* Rp, cplx* Rm, cudaStream_t st) {
    layer_form_kernel<<<dim3((2 * N + 255) / 256, N, nb), 256, 0, st>>>(W, QW, kz, vfinv, omega, thick, N, Mp, Mm, Rp, Rm);
    return cudaGetLastError();
}
cudaError_t layer_finish(const cplx* Tp, const cplx* Tm, int nb, int n, cplx* S11, cplx* S21, cudaStream_t st) {
    layer_finish_kernel<<<dim3((n + 255) / 256, n, nb), 256, 0, st>>>(Tp, Tm, n, S11, S21);
    return cudaGetLastError();
}
cudaError_t blockdiag_dense(const cplx* d4, int nb, int N, cplx* D, cudaStream_t st) {
    blockdiag_dense_kernel<<<dim3((2 * N + 255) / 256, 2 * N, nb), 256, 0, st>>>(d4, N, D);
    return cudaGetLastError();
}
cudaError_t bd_left_mul(const cplx* d4, const cplx* X, int nb, int N, int ncols, cplx alpha, cplx beta, cplx* Y, cudaStream_t st) {
    bd_left_mul_kernel<<<dim3((ncols + 255) / 256, N, nb), 256, 0, st>>>(d4, X, N, ncols, alpha, beta, Y);
    return cudaGetLastError();
}
cudaError_t bd_right_mul(const cplx* d4, const cplx* X, int nb, int N, int nrows, cplx alpha, cplx beta, cplx* Y, cudaStream_t st) {
    bd_right_mul_kernel<<<dim3((N + 255) / 256, nrows, nb), 256, 0, st>>>(d4, X, N, nrows, alpha, beta, Y);
    return cudaGetLastError();
}
cudaError_t bd_add(const cplx* d4, int nb, int N, cplx alpha, cplx* D, cudaStream_t st) {
    bd_add_kernel<<<dim3((N + 255) / 256, 1, nb), 256, 0, st>>>(d4, N, alpha, D);
    return cudaGetLastError();
}
cudaError_t sym_project(const cplx* X, int nb, int n, const int* il, const cplx* cl, const int* ir, const cplx* cr,
                        int G, int nkl, int nkr, cplx* out, cudaStream_t st) {
    sym_project_kernel<<<dim3((nkr + 127) / 128, nkl, nb), 128, 0, st>>>(X, n, il, cl, ir, cr, G, nkl, nkr, out);
    return cudaGetLastError();
}
cudaError_t set_identity(cplx* A, int n, int lda, long long stride, int nb, cudaStream_t st) {
    identity_kernel<<<dim3((n + 255) / 256, n, nb), 256, 0, st>>>(A, n, lda, stride);
    return cudaGetLastError();
}
cudaError_t axpby(cplx alpha, const cplx* X, cplx beta, cplx* Y, size_t total, cudaStream_t st) {
    axpby_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(alpha, X, beta, Y, total);
    return cudaGetLastError();
}

}  // namespace rcwa
