// Phase 1 of rcwa_eig: batched BLOCKED Householder Hessenberg reduction  A = Z H Z^H  (complex128).
//
// Panel of HB_NB = 64 columns (compact WY, LAPACK zgehrd/zlahr2 structure, Hermitian reflectors
// H_j = I - u_j u_j^H with |u_j|^2 = 2, i.e. tau = 1):
//
//   per column j of the panel
//     hb_col_kernel    (one CTA per matrix)  finishes Y/T of the previous column, applies the
//                      panel's earlier reflectors to column j, builds u_j           -- O(n * NB) work
//     hb_matvec_kernel (row bands x batch)   y = A[k0+1:n, j+1:n] * u_j             -- THE streaming kernel:
//                      read-only, one pass over the trailing matrix per column, 16 independent 16-byte
//                      loads in flight per lane, warp-shuffle row reductions.  Its bytes are exactly
//                      the algorithmic bytes of the reduction: sum_j (n-j)^2 * 16 B = 16 n^3/3.
//   per panel (DMMA grouped GEMM, zgemm.cu)
//     Y_top = A_top V T ;  A[:, right] -= Y V^H ;  A[below, right] -= V T^H (V^H A) ;  Z -= (Z V T) V^H
//
// Replaces the Hessenberg stage inside LAPACK zgeev reached through torch.linalg.eig
// (/root/reference/torcwa/torch_eig.py:14).
#include "common.cuh"
#include "kernels.h"

#ifndef HB_NB
#define HB_NB 64          // panel width (compile-time).  64: the panel GEMMs are K = 64 / N = 64 products, which run at
                          // 21-31 TFLOP/s instead of 14-27 at 32 (profiles/r1b_gemm_probe.json); measured -7 % on the phase
#endif
#define HB_ROWS 32          // rows per CTA in the matvec (8 warps x 4 rows)

namespace {

__device__ __forceinline__ cplx ld_nc(const cplx* p) {
    cplx r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

struct HbPtrs { cplx *V, *Y, *T, *u, *tv; };    // V,Y: [n][NB]; T: [NB][NB]; u: [n]; tv: [NB] (+beta at tv[NB])
__host__ __device__ inline size_t hb_elems(int n) { return (size_t)2 * n * HB_NB + HB_NB * HB_NB + n + HB_NB + 16; }
__device__ __forceinline__ HbPtrs hb_ptrs(cplx* base, int n) {
    HbPtrs p; p.V = base; p.Y = p.V + (size_t)n * HB_NB; p.T = p.Y + (size_t)n * HB_NB; p.u = p.T + HB_NB * HB_NB; p.tv = p.u + n;
    return p;
}

// One CTA per matrix.  Column j = k0 + i of the current panel.  `finish_only`: only complete Y/T of
// column i-1 (called once after the last column of a panel).
// dynamic smem: b[n] cplx + red[NB][17] cplx
__global__ void __launch_bounds__(512, 1)
hb_col_kernel(cplx* A, long long astride, int lda, int n, int k0, int i, int finish_only, cplx* wsb, long long wstride) {
    extern __shared__ __align__(16) char smem_raw[];
    cplx* bv = reinterpret_cast<cplx*>(smem_raw);             // [n]
    cplx* red = bv + n;                                        // [16][NB] partials
    __shared__ cplx wv[HB_NB], tvs[HB_NB];
    __shared__ double sred[40];
    __shared__ cplx sh_beta; __shared__ double sh_scl;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    cplx* Ab = A + (size_t)b * astride;
    HbPtrs p = hb_ptrs(wsb + (size_t)b * wstride, n);
    Cta c = make_cta(b, nullptr);
    const int r0 = k0 + 1;                                     // first row touched by the panel's reflectors
    const int tcomp = tid & (HB_NB - 1), tchunk = tid / HB_NB, nchunk = nt / HB_NB;     // (component, row chunk)

    // ---------------- phase 0: finish column c = i-1:  tv = V[:,0:c]^H u_c ; Y[:,c] -= Y[:,0:c] tv ; T[:,c]
    if (i > 0) {
        const int cc = i - 1;
        if (cc > 0) {
            cplx acc = C(0, 0);
            if (tcomp < cc) for (int r = r0 + tchunk; r < n; r += nchunk) acc = cadd(acc, cmulc(p.V[(size_t)r * HB_NB + tcomp], p.V[(size_t)r * HB_NB + cc]));
            red[tchunk * HB_NB + tcomp] = acc;
            __syncthreads();
            if (tid < cc) { cplx s = C(0, 0); for (int q = 0; q < nchunk; ++q) s = cadd(s, red[q * HB_NB + tid]); tvs[tid] = s; }
            __syncthreads();
            for (int r = r0 + tid; r < n; r += nt) {
                cplx y = p.Y[(size_t)r * HB_NB + cc];
                for (int t = 0; t < cc; ++t) y = csub(y, cmul(p.Y[(size_t)r * HB_NB + t], tvs[t]));
                p.Y[(size_t)r * HB_NB + cc] = y;
            }
            if (tid < cc) {                                   // T[0:cc, cc] = -T[0:cc,0:cc] * tv   (upper triangular T)
                cplx s = C(0, 0);
                for (int t = tid; t < cc; ++t) s = cadd(s, cmul(p.T[tid * HB_NB + t], tvs[t]));
                p.T[tid * HB_NB + cc] = cneg(s);
            }
        }
        if (tid == 0) p.T[cc * HB_NB + cc] = C(1, 0);
        if (tid < HB_NB && tid > cc) p.T[tid * HB_NB + cc] = C(0, 0);
        __syncthreads();
    }
    if (finish_only) return;

    const int j = k0 + i;
    // ---------------- phase 1: b = A[r0:n, j] updated by the panel's earlier reflectors
    for (int r = r0 + tid; r < n; r += nt) {
        cplx v = Ab[(size_t)r * lda + j];
        for (int t = 0; t < i; ++t) v = csub(v, cmul(p.Y[(size_t)r * HB_NB + t], cconj(p.V[(size_t)j * HB_NB + t])));   // right: A (I - V T V^H)
        bv[r] = v;
    }
    __syncthreads();
    if (i > 0) {
        cplx acc = C(0, 0);                                    // w = V^H b
        if (tcomp < i) for (int r = r0 + tchunk; r < n; r += nchunk) acc = cadd(acc, cmulc(p.V[(size_t)r * HB_NB + tcomp], bv[r]));
        red[tchunk * HB_NB + tcomp] = acc;
        __syncthreads();
        if (tid < i) { cplx s = C(0, 0); for (int q = 0; q < nchunk; ++q) s = cadd(s, red[q * HB_NB + tid]); wv[tid] = s; }
        __syncthreads();
        if (tid < i) {                                         // w <- T^H w
            cplx s = C(0, 0);
            for (int t = 0; t <= tid; ++t) s = cadd(s, cmulc(p.T[t * HB_NB + tid], wv[t]));
            tvs[tid] = s;
        }
        __syncthreads();
        for (int r = r0 + tid; r < n; r += nt) {               // left: b -= V (T^H V^H b)
            cplx v = bv[r];
            for (int t = 0; t < i; ++t) v = csub(v, cmul(p.V[(size_t)r * HB_NB + t], tvs[t]));
            bv[r] = v;
        }
        __syncthreads();
    }
    // ---------------- phase 2: reflector from b[j+1:n]
    double ss = 0.0;
    for (int r = j + 1 + tid; r < n; r += nt) ss += cabs2(bv[r]);
    ss = cta_sum(c, ss, sred);
    if (tid == 0) {
        const double sigma = sqrt(ss);
        const cplx x1 = bv[j + 1];
        const double ax = cabs_(x1);
        if (sigma == 0.0) { sh_beta = x1; sh_scl = 0.0; }
        else {
            const cplx ph = (ax == 0.0) ? C(1, 0) : cscale(x1, 1.0 / ax);
            sh_beta = cscale(ph, -sigma);
            sh_scl = 1.0 / sqrt(sigma * (sigma + ax));
        }
    }
    __syncthreads();
    const cplx beta = sh_beta; const double scl = sh_scl;
    for (int r = tid; r < n; r += nt) {
        cplx u = C(0, 0);
        if (r > j && scl != 0.0) { u = bv[r]; if (r == j + 1) u = csub(u, beta); u = cscale(u, scl); }
        p.V[(size_t)r * HB_NB + i] = u;
        p.u[r] = u;
        if (r >= r0) {                                         // write the finished column back
            cplx v = (r <= j) ? bv[r] : ((r == j + 1) ? beta : C(0, 0));
            Ab[(size_t)r * lda + j] = v;
        }
    }
}

// y[r] = sum_{c > j} A[r][c] u[c]  for rows r in [k0+1, n); stored to Y[r][i].
// grid (ceil((n-k0-1)/HB_ROWS), B), 256 threads = 8 warps x 4 rows: small row bands keep the grid large
// (>= 2 waves of CTAs even at batch 8) and each lane has 16 independent 16-byte loads per 128-column chunk.
__global__ void __launch_bounds__(256, 3)
hb_matvec_kernel(const cplx* __restrict__ A, long long astride, int lda, int n, int k0, int i, cplx* wsb, long long wstride) {
    const int b = blockIdx.y;
    HbPtrs p = hb_ptrs(wsb + (size_t)b * wstride, n);
    const cplx* Ab = A + (size_t)b * astride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = k0 + i;
    const int row0 = k0 + 1 + blockIdx.x * HB_ROWS + warp * 4;
    cplx acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = C(0, 0);
    const int cstart = (j + 1) - ((j + 1) % 128);
    for (int c0 = cstart; c0 < n; c0 += 128) {
        cplx uc[4], av[4][4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = row0 + rr;
            const cplx* rp = Ab + (size_t)r * lda;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = c0 + q * 32 + lane;
                av[rr][q] = (r < n && col > j && col < n) ? ld_nc(rp + col) : C(0, 0);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = c0 + q * 32 + lane;
            uc[q] = (col > j && col < n) ? p.u[col] : C(0, 0);
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[rr] = cfma(av[rr][q], uc[q], acc[rr]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double yr = warp_sum(acc[r].x), yi = warp_sum(acc[r].y);
        const int row = row0 + r;
        if (lane == 0 && row < n) p.Y[(size_t)row * HB_NB + i] = C(yr, yi);
    }
}

__global__ void hb_zero_kernel(cplx* wsb, long long wstride, int n) {
    HbPtrs p = hb_ptrs(wsb + (size_t)blockIdx.y * wstride, n);
    const size_t tot = (size_t)2 * n * HB_NB + HB_NB * HB_NB;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (size_t)gridDim.x * blockDim.x) p.V[idx] = C(0, 0);
}

inline size_t al(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

namespace rcwa {

// workspace: per-matrix panel storage + W buffers for the block updates + GEMM descriptors
size_t hessenberg_workspace_bytes(int n, int nb) {
    return al(sizeof(cplx) * hb_elems(n) * nb) + 2 * al(sizeof(cplx) * (size_t)HB_NB * n * nb) + 2 * al(sizeof(cplx) * (size_t)n * HB_NB * nb)
         + al(sizeof(ZGemmProblem) * (size_t)nb * 4);
}

#define HK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)

// One launch of the streaming mat-vec of column j = k0 + i exactly as hessenberg_blocked issues it (profiling
// entry: bench.py times these launches alone with CUDA events for the roofline figure).  Reads the
// (n - k0 - 1) x (n - j - 1) trailing block of every matrix once: 16 (n-k0-1)(n-j-1) nb bytes.
cudaError_t hessenberg_matvec_probe(const cplx* A, int n, int nb, int j, char* wsb, cudaStream_t st) {
    if (j < 0 || j > n - 3) return cudaErrorInvalidValue;
    const int k0 = (j / HB_NB) * HB_NB, i = j - k0;
    cplx* pan = (cplx*)wsb;
    hb_matvec_kernel<<<dim3((n - k0 - 1 + HB_ROWS - 1) / HB_ROWS, nb), 256, 0, st>>>(A, (long long)n * n, n, n, k0, i, pan, (long long)hb_elems(n));
    return cudaGetLastError();
}
int hessenberg_panel_width() { return HB_NB; }

// A [nb] (n x n) -> upper Hessenberg in place; Z [nb] (n x n) <- accumulated reflectors (A_in = Z H Z^H)
cudaError_t hessenberg_blocked(cplx* A, int n, int nb, cplx* Z, char* wsb, cudaStream_t st, cudaEvent_t after_first_columns) {
    const long long ms = (long long)n * n;
    HK(set_identity(Z, n, n, ms, nb, st));
    if (n < 3) return cudaSuccess;
    char* q = wsb;
    cplx* pan = (cplx*)q; q += al(sizeof(cplx) * hb_elems(n) * nb);
    cplx* W = (cplx*)q; q += al(sizeof(cplx) * (size_t)HB_NB * n * nb);          // [NB][n]
    cplx* W1 = (cplx*)q; q += al(sizeof(cplx) * (size_t)HB_NB * n * nb);         // [NB][n]
    cplx* X = (cplx*)q; q += al(sizeof(cplx) * (size_t)n * HB_NB * nb);          // [n][NB]
    cplx* X1 = (cplx*)q; q += al(sizeof(cplx) * (size_t)n * HB_NB * nb);         // [n][NB]
    ZGemmProblem* gs = (ZGemmProblem*)q;
    const long long wstride = (long long)hb_elems(n);
    const long long sW = (long long)HB_NB * n, sX = (long long)n * HB_NB;
    cplx* V = pan;                                           // per-matrix offsets inside `pan` (stride wstride)
    cplx* Y = pan + (size_t)n * HB_NB;
    cplx* T = Y + (size_t)n * HB_NB;
    const size_t smem_col = sizeof(cplx) * ((size_t)n + 16 * HB_NB);
    {   // per-device attribute (one process may drive several GPUs)
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            HK(cudaFuncSetAttribute(hb_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    if (smem_col > 200 * 1024) return cudaErrorInvalidValue;
    const cplx one = C(1, 0), zero = C(0, 0), mone = C(-1, 0);
    const int last = n - 3;                                  // last column that gets a reflector
    for (int k0 = 0; k0 <= last; k0 += HB_NB) {
        const int nbe = (last - k0 + 1 < HB_NB) ? last - k0 + 1 : HB_NB;
        hb_zero_kernel<<<dim3(32, nb), 256, 0, st>>>(pan, wstride, n);
        for (int i = 0; i < nbe; ++i) {
            hb_col_kernel<<<nb, 512, smem_col, st>>>(A, ms, n, n, k0, i, 0, pan, wstride);
            hb_matvec_kernel<<<dim3((n - k0 - 1 + HB_ROWS - 1) / HB_ROWS, nb), 256, 0, st>>>(A, ms, n, n, k0, i, pan, wstride);
        }
        hb_col_kernel<<<nb, 512, smem_col, st>>>(A, ms, n, n, k0, nbe, 1, pan, wstride);
        HK(cudaGetLastError());
        if (k0 == 0 && after_first_columns) HK(cudaEventRecord(after_first_columns, st));
        const int r0 = k0 + 1, nr = n - r0;                  // rows touched by this panel's reflectors
        const int c1 = k0 + nbe, nc = n - c1;                // trailing columns (right of the panel)
        // ---- Y_top = A[0:r0, r0:n] V[r0:n,:] T
        HK(zgemm_strided(OP_N, OP_N, r0, nbe, nr, one, A + r0, n, ms, V + (size_t)r0 * HB_NB, HB_NB, wstride, zero, X, HB_NB, sX, nb, gs, st));
        HK(zgemm_strided(OP_N, OP_N, r0, nbe, nbe, one, X, HB_NB, sX, T, HB_NB, wstride, zero, Y, HB_NB, wstride, nb, gs, st));
        // ---- right update: A[:, c1:n] -= Y V[c1:n,:]^H ; top rows of the panel's own columns too
        if (nc > 0)
            HK(zgemm_strided(OP_N, OP_H, n, nc, nbe, mone, Y, HB_NB, wstride, V + (size_t)c1 * HB_NB, HB_NB, wstride, one, A + c1, n, ms, nb, gs, st));
        if (nbe > 1)
            HK(zgemm_strided(OP_N, OP_H, r0, nbe - 1, nbe, mone, Y, HB_NB, wstride, V + (size_t)r0 * HB_NB, HB_NB, wstride, one, A + r0, n, ms, nb, gs, st));
        // ---- left update: A[r0:n, c1:n] -= V T^H (V^H A[r0:n, c1:n])
        if (nc > 0) {
            HK(zgemm_strided(OP_H, OP_N, nbe, nc, nr, one, V + (size_t)r0 * HB_NB, HB_NB, wstride, A + (size_t)r0 * n + c1, n, ms, zero, W, n, sW, nb, gs, st));
            HK(zgemm_strided(OP_H, OP_N, nbe, nc, nbe, one, T, HB_NB, wstride, W, n, sW, zero, W1, n, sW, nb, gs, st));
            HK(zgemm_strided(OP_N, OP_N, nr, nc, nbe, mone, V + (size_t)r0 * HB_NB, HB_NB, wstride, W1, n, sW, one, A + (size_t)r0 * n + c1, n, ms, nb, gs, st));
        }
        // ---- Z[:, r0:n] -= (Z[:, r0:n] V T) V^H
        HK(zgemm_strided(OP_N, OP_N, n, nbe, nr, one, Z + r0, n, ms, V + (size_t)r0 * HB_NB, HB_NB, wstride, zero, X, HB_NB, sX, nb, gs, st));
        HK(zgemm_strided(OP_N, OP_N, n, nbe, nbe, one, X, HB_NB, sX, T, HB_NB, wstride, zero, X1, HB_NB, sX, nb, gs, st));
        HK(zgemm_strided(OP_N, OP_H, n, nr, nbe, mone, X1, HB_NB, sX, V + (size_t)r0 * HB_NB, HB_NB, wstride, one, Z + r0, n, ms, nb, gs, st));
    }
    return cudaGetLastError();
}

}  // namespace rcwa
