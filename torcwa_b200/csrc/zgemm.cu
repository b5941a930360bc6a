// Grouped / batched complex128 GEMM on the fp64 tensor path (DMMA, mma.sync m8n8k4.f64).
//
//   C_p (M x N) = alpha * op(A_p) * op(B_p) + beta * C_p        for every problem p
//
// Row-major, interleaved (re,im) fp64.  op in {N, T, H}.  tcgen05 has no f64 kind, so the fp64
// tensor path on sm_100a is mma.sync; a complex product is four real MMAs
// on (re,im) fragments that share one 16-byte shared-memory load per operand element.
//
// Tiling: CTA tile BM x BN (64x128 or 128x64), 8 warps, warp tile 32x32 complex (4x4 m8n8 tiles,
// 64 accumulator pairs), BK = 8, 3-stage cp.async pipeline (16 B = one complex element per copy,
// which lets the copy itself transpose A into the k-major layout the fragments want; zero-fill for
// ragged edges).  Shared-memory leading dimensions are == 2 (mod 8) elements, which makes the
// per-quarter-warp LDS.128 fragment loads conflict-free.
//
// In-place use (needed by the QR sweeps): C may alias B when M <= BM (one tile covers all rows)
// and may alias A when N <= BN, because a CTA reads its whole K extent before the epilogue writes
// and no other CTA touches those rows/columns.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BK = 8;
constexpr int STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int BM, int BN, int OPA, int OPB>
__global__ void __launch_bounds__(256, 1)
zgemm_grouped_kernel(const ZGemmProblem* __restrict__ probs, cplx alpha, cplx beta) {
    constexpr int WN = BN / 32;
    constexpr int LDA_S = BM + 2, LDB_S = BN + 2;
    constexpr int A_STAGE = BK * LDA_S, B_STAGE = BK * LDB_S;
    static_assert((BM / 32) * WN == 8, "8 warps");
    extern __shared__ __align__(16) char smem_raw[];
    cplx* As = reinterpret_cast<cplx*>(smem_raw);
    cplx* Bs = As + STAGES * A_STAGE;

    const ZGemmProblem p = probs[blockIdx.y];
    if (p.M <= 0 || p.N <= 0) return;
    const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
    if ((int)blockIdx.x >= tiles_m * tiles_n) return;
    const int m0 = ((int)blockIdx.x / tiles_n) * BM, n0 = ((int)blockIdx.x % tiles_n) * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int nk = (p.K + BK - 1) / BK;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * BK;
        cplx* as = As + stage * A_STAGE;
        cplx* bs = Bs + stage * B_STAGE;
#pragma unroll
        for (int i = 0; i < BM * BK / 256; ++i) {
            int idx = tid + i * 256, m, k;
            if (OPA == 0) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            bool ok = (m0 + m < p.M) && (k0 + k < p.K);
            const cplx* src = p.A;
            if (ok) src = (OPA == 0) ? p.A + (size_t)(m0 + m) * p.lda + (k0 + k) : p.A + (size_t)(k0 + k) * p.lda + (m0 + m);
            cp_async16(as + k * LDA_S + m, src, ok ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < BN * BK / 256; ++i) {
            int idx = tid + i * 256, n, k;
            if (OPB == 0) { n = idx % BN; k = idx / BN; } else { k = idx % BK; n = idx / BK; }
            bool ok = (n0 + n < p.N) && (k0 + k < p.K);
            const cplx* src = p.B;
            if (ok) src = (OPB == 0) ? p.B + (size_t)(k0 + k) * p.ldb + (n0 + n) : p.B + (size_t)(n0 + n) * p.ldb + (k0 + k);
            cp_async16(bs + k * LDB_S + n, src, ok ? 16 : 0);
        }
    };

    double acc_re[4][4][2], acc_im[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc_re[i][j][0] = acc_re[i][j][1] = 0.0; acc_im[i][j][0] = acc_im[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) { if (s < nk) load_stage(s, s); cp_async_commit(); }

    const int fr = lane >> 2, fk = lane & 3;
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        { int nx = kt + STAGES - 1; if (nx < nk) load_stage(nx % STAGES, nx); cp_async_commit(); }
        const cplx* as = As + (kt % STAGES) * A_STAGE + wm * 32 + fr;
        const cplx* bs = Bs + (kt % STAGES) * B_STAGE + wn * 32 + fr;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            cplx a[4], b[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                a[t] = as[(kk + fk) * LDA_S + t * 8];
                b[t] = bs[(kk + fk) * LDB_S + t * 8];
                if (OPA == 2) a[t].y = -a[t].y;
                if (OPB == 2) b[t].y = -b[t].y;
            }
            // four real MMAs per complex tile, issued pass by pass so that the two MMAs that accumulate into
            // the same registers are 16 instructions apart (back-to-back they serialise on the DMMA latency)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc_re[i][j][0], acc_re[i][j][1], a[i].x, b[j].x);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc_im[i][j][0], acc_im[i][j][1], a[i].x, b[j].y);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double nai = -a[i].y;
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc_re[i][j][0], acc_re[i][j][1], nai, b[j].y);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc_im[i][j][0], acc_im[i][j][1], a[i].y, b[j].x);
        }
    }
    cp_async_wait<0>();

    const bool use_beta = !(beta.x == 0.0 && beta.y == 0.0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + wm * 32 + i * 8 + fr;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = n0 + wn * 32 + j * 8 + fk * 2 + e;
                if (col >= p.N) continue;
                cplx* dst = p.C + (size_t)row * p.ldc + col;
                cplx v = cmul(alpha, C(acc_re[i][j][e], acc_im[i][j][e]));
                if (use_beta) v = cadd(v, cmul(beta, *dst));
                *dst = v;
            }
        }
    }
}

__global__ void fill_strided_kernel(ZGemmProblem* probs, int batch, const cplx* A, const cplx* B, cplx* C,
                                    long long sa, long long sb, long long sc, int M, int N, int K, int lda, int ldb, int ldc) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    ZGemmProblem p;
    p.A = A + (size_t)b * sa; p.B = B + (size_t)b * sb; p.C = C + (size_t)b * sc;
    p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    probs[b] = p;
}

template <int BM, int BN, int OPA, int OPB>
cudaError_t launch_cfg(const ZGemmProblem* probs, int nprob, int max_tiles, cplx alpha, cplx beta, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * BK * ((BM + 2) + (BN + 2)) * sizeof(cplx);
    static bool attr_set = false;   // idempotent attribute; benign if raced
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(zgemm_grouped_kernel<BM, BN, OPA, OPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid(max_tiles, nprob);
    zgemm_grouped_kernel<BM, BN, OPA, OPB><<<grid, 256, smem, st>>>(probs, alpha, beta);
    return cudaGetLastError();
}

template <int BM, int BN>
cudaError_t launch_ops(int opa, int opb, const ZGemmProblem* probs, int nprob, int max_tiles, cplx alpha, cplx beta, cudaStream_t st) {
    switch (opa * 3 + opb) {
        case 0: return launch_cfg<BM, BN, 0, 0>(probs, nprob, max_tiles, alpha, beta, st);
        case 1: return launch_cfg<BM, BN, 0, 1>(probs, nprob, max_tiles, alpha, beta, st);
        case 2: return launch_cfg<BM, BN, 0, 2>(probs, nprob, max_tiles, alpha, beta, st);
        case 3: return launch_cfg<BM, BN, 1, 0>(probs, nprob, max_tiles, alpha, beta, st);
        case 4: return launch_cfg<BM, BN, 1, 1>(probs, nprob, max_tiles, alpha, beta, st);
        case 5: return launch_cfg<BM, BN, 1, 2>(probs, nprob, max_tiles, alpha, beta, st);
        case 6: return launch_cfg<BM, BN, 2, 0>(probs, nprob, max_tiles, alpha, beta, st);
        case 7: return launch_cfg<BM, BN, 2, 1>(probs, nprob, max_tiles, alpha, beta, st);
        default: return launch_cfg<BM, BN, 2, 2>(probs, nprob, max_tiles, alpha, beta, st);
    }
}

}  // namespace

namespace rcwa {

int gemm_tiles(int tile_cfg, int M, int N) {
    int bm = tile_cfg == GEMM_TILE_64x128 ? 64 : 128, bn = tile_cfg == GEMM_TILE_64x128 ? 128 : 64;
    return ((M + bm - 1) / bm) * ((N + bn - 1) / bn);
}

cudaError_t zgemm_grouped(int tile_cfg, int opa, int opb, const ZGemmProblem* probs, int nprob, int max_tiles,
                          cplx alpha, cplx beta, cudaStream_t st) {
    if (nprob <= 0 || max_tiles <= 0) return cudaSuccess;
    if (tile_cfg == GEMM_TILE_64x128) return launch_ops<64, 128>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
    return launch_ops<128, 64>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
}

cudaError_t zgemm_strided(int opa, int opb, int M, int N, int K, cplx alpha, const cplx* A, int lda, long long sa,
                          const cplx* B, int ldb, long long sb, cplx beta, cplx* Cm, int ldc, long long sc,
                          int batch, ZGemmProblem* scratch, cudaStream_t st) {
    if (batch <= 0 || M <= 0 || N <= 0) return cudaSuccess;
    fill_strided_kernel<<<(batch + 127) / 128, 128, 0, st>>>(scratch, batch, A, B, Cm, sa, sb, sc, M, N, K, lda, ldb, ldc);
    // pick the tile shape that wastes less of the output
    int cfg = (gemm_tiles(GEMM_TILE_64x128, M, N) * 64 * 128 <= gemm_tiles(GEMM_TILE_128x64, M, N) * 128 * 64) ? GEMM_TILE_64x128 : GEMM_TILE_128x64;
    return zgemm_grouped(cfg, opa, opb, scratch, batch, gemm_tiles(cfg, M, N), alpha, beta, st);
}

}  // namespace rcwa
