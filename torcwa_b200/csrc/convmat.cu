// Stage 1: Fourier factorisation of a sampled unit cell  (replaces rcwa._material_conv,
// /root/reference/torcwa/rcwa.py:1183-1204).
//
// The reference takes a full nx x ny FFT and then gathers the (4ox+1) x (4oy+1) low-order
// coefficients into the N x N Toeplitz ("convolution") matrix, N = (2ox+1)(2oy+1).  Only those
// coefficients are ever used, so we compute exactly them with a separable pruned DFT in fp64
// (exact integer phase reduction (p*x mod nx) + sincospi twiddles), then write the Toeplitz
// matrix with coalesced rows:
//
//   G[x][q]  = sum_y grid[x][y] * exp(-2 pi i q y / ny)                       (dft_rows)
//   F[p][q]  = 1/(nx ny) * sum_x G[x][q] * exp(-2 pi i p x / nx)              (dft_cols)
//   E[i][j]  = F[m_i - m_j][n_i - n_j],  i = (2oy+1)(m+ox) + (n+oy)           (toeplitz)
//
// HBM-bound and tiny next to the eigen stage: 16 N^2 bytes written per design point.
#include "common.cuh"
#include "kernels.h"

namespace {

template <typename T> __device__ __forceinline__ cplx load_as_cplx(const T* p, size_t i);
template <> __device__ __forceinline__ cplx load_as_cplx<float>(const float* p, size_t i) { return C((double)p[i], 0.0); }
template <> __device__ __forceinline__ cplx load_as_cplx<double>(const double* p, size_t i) { return C(p[i], 0.0); }
template <> __device__ __forceinline__ cplx load_as_cplx<float2>(const float2* p, size_t i) { float2 v = p[i]; return C((double)v.x, (double)v.y); }
template <> __device__ __forceinline__ cplx load_as_cplx<double2>(const double2* p, size_t i) { return p[i]; }

// grid (nx, B): one CTA per real-space row
template <typename T>
__global__ void dft_rows_kernel(const T* __restrict__ grid, long long grid_stride, int nx, int ny, int oy, cplx* __restrict__ G) {
    extern __shared__ __align__(16) char smem_raw[];
    cplx* row = reinterpret_cast<cplx*>(smem_raw);   // [ny]
    cplx* tw = row + ny;                             // [ny]  exp(-2 pi i k / ny)
    const int x = blockIdx.x, b = blockIdx.y, QY = 4 * oy + 1;
    const T* src = grid + (size_t)b * grid_stride + (size_t)x * ny;
    for (int y = threadIdx.x; y < ny; y += blockDim.x) {
        row[y] = load_as_cplx<T>(src, y);
        double s, c;
        sincospi(-2.0 * (double)y / (double)ny, &s, &c);
        tw[y] = C(c, s);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < QY; q += blockDim.x) {
        long long qm = ((q - 2 * oy) % ny + ny) % ny;
        cplx acc = C(0.0, 0.0);
        long long ph = 0;
        for (int y = 0; y < ny; ++y) {
            acc = cfma(row[y], tw[ph], acc);
            ph += qm; if (ph >= ny) ph -= ny;
        }
        G[((size_t)b * nx + x) * QY + q] = acc;
    }
}

// grid (QX, B): one CTA per output p; threads over q
__global__ void dft_cols_kernel(const cplx* __restrict__ G, int nx, int ny, int ox, int oy, cplx* __restrict__ F) {
    extern __shared__ __align__(16) char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);    // [nx] exp(-2 pi i k / nx)
    const int p = blockIdx.x, b = blockIdx.y, QY = 4 * oy + 1, QX = 4 * ox + 1;
    for (int x = threadIdx.x; x < nx; x += blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)x / (double)nx, &s, &c);
        tw[x] = C(c, s);
    }
    __syncthreads();
    const long long pm = ((p - 2 * ox) % nx + nx) % nx;
    const double scale = 1.0 / ((double)nx * (double)ny);
    for (int q = threadIdx.x; q < QY; q += blockDim.x) {
        cplx acc = C(0.0, 0.0);
        long long ph = 0;
        const cplx* g = G + (size_t)b * nx * QY + q;
        for (int x = 0; x < nx; ++x) {
            acc = cfma(g[(size_t)x * QY], tw[ph], acc);
            ph += pm; if (ph >= nx) ph -= nx;
        }
        F[((size_t)b * QX + p) * QY + q] = cscale(acc, scale);
    }
}

// grid (N, B): one CTA per output row i
__global__ void toeplitz_kernel(const cplx* __restrict__ F, int ox, int oy, cplx* __restrict__ E) {
    const int NY = 2 * oy + 1, N = (2 * ox + 1) * NY, QY = 4 * oy + 1, QX = 4 * ox + 1;
    const int i = blockIdx.x, b = blockIdx.y;
    const int mi = i / NY, ni = i % NY;
    const cplx* f = F + (size_t)b * QX * QY;
    cplx* e = E + ((size_t)b * N + i) * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        int mj = j / NY, nj = j % NY;
        e[j] = __ldg(&f[(size_t)(mi - mj + 2 * ox) * QY + (ni - nj + 2 * oy)]);
    }
}

template <typename T>
cudaError_t run(const void* grid, long long grid_stride, int nx, int ny, int nb, int ox, int oy, cplx* E, cplx* ws, cudaStream_t st) {
    const int QX = 4 * ox + 1, QY = 4 * oy + 1, N = (2 * ox + 1) * (2 * oy + 1);
    cplx* G = ws;                                    // [B, nx, QY]
    cplx* F = ws + (size_t)nb * nx * QY;             // [B, QX, QY]
    dft_rows_kernel<T><<<dim3(nx, nb), 128, 2 * ny * sizeof(cplx), st>>>((const T*)grid, grid_stride, nx, ny, oy, G);
    dft_cols_kernel<<<dim3(QX, nb), 128, nx * sizeof(cplx), st>>>(G, nx, ny, ox, oy, F);
    toeplitz_kernel<<<dim3(N, nb), 256, 0, st>>>(F, ox, oy, E);
    return cudaGetLastError();
}

}  // namespace

namespace rcwa {

size_t convmat_workspace_elems(int nx, int ny, int nb, int ox, int oy) {
    return (size_t)nb * ((size_t)nx * (4 * oy + 1) + (size_t)(4 * ox + 1) * (4 * oy + 1));
}

cudaError_t convmat(const void* grid, int grid_type, long long grid_stride, int nx, int ny, int nb, int ox, int oy,
                    cplx* E, cplx* ws, cudaStream_t st) {
    switch (grid_type) {
        case 0: return run<float>(grid, grid_stride, nx, ny, nb, ox, oy, E, ws, st);
        case 1: return run<double>(grid, grid_stride, nx, ny, nb, ox, oy, E, ws, st);
        case 2: return run<float2>(grid, grid_stride, nx, ny, nb, ox, oy, E, ws, st);
        default: return run<double2>(grid, grid_stride, nx, ny, nb, ox, oy, E, ws, st);
    }
}

}  // namespace rcwa
