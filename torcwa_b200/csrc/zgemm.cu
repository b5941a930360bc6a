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

// Register-lean operand loader.  One k-step moves an [R x BK] operand tile into shared memory as
// s[k][r] (leading dimension R + 2).  Stages are always loaded in k order, so the global pointer is a
// running one and the per-copy addresses are (running pointer + i * constant): no per-copy index math.
//   KFAST:  operand stored [r][k] (ld between r)  -- A with op N, B with op T/H
//   !KFAST: operand stored [k][r] (ld between k)  -- A with op T/H, B with op N
template <int R, int NTHR, bool KFAST>
struct TileLoader {
    static constexpr int NI = R * BK / NTHR;
    static constexpr int ISTEP = KFAST ? NTHR / BK : NTHR / R;      // rows (KFAST) or k (!KFAST) between copies of one thread
    static_assert((R * BK) % NTHR == 0 && (KFAST ? NTHR % BK == 0 : NTHR % R == 0), "tile / thread-count mismatch");
    const cplx* base; const cplx* ptr;
    long long istride; int kstride, rlim, klim, sdst;
    __device__ __forceinline__ void init(const cplx* b, int ld, int r0, int rtot, int ktot, int tid) {
        base = b;
        if (KFAST) { const int k = tid % BK, r = tid / BK; ptr = b + (size_t)(r0 + r) * ld + k; istride = (long long)ISTEP * ld; kstride = BK; rlim = rtot - r0 - r; klim = ktot - k; sdst = k * (R + 2) + r; }
        else { const int r = tid % R, k = tid / R; ptr = b + (size_t)k * ld + (r0 + r); istride = (long long)ISTEP * ld; kstride = BK * ld; rlim = rtot - r0 - r; klim = ktot - k; sdst = k * (R + 2) + r; }
    }
    __device__ __forceinline__ void load(cplx* stage, int k0) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const bool ok = KFAST ? (i * ISTEP < rlim && k0 < klim) : (rlim > 0 && k0 + i * ISTEP < klim);
            cplx* dst = stage + sdst + (KFAST ? i * ISTEP : i * ISTEP * (R + 2));
            cp_async16(dst, ok ? ptr + i * istride : base, ok ? 16 : 0);
        }
        ptr += kstride;
    }
};

// Generic kernel: CTA tile BM x BN, warp tile WM x WN (multiples of 8), (BM/WM)*(BN/WN) warps,
// MINB = CTAs per SM the register budget is planned for.  M3 = 3-multiplication complex product
// (three real accumulators  P1 = sum ar*br, P2 = sum ai*bi, P3 = sum (ar+ai)(br+bi);  re = P1 - P2,
// im = P3 - P1 - P2): 25 % fewer DMMAs at norm-wise (not component-wise) fp64 accuracy.
// BAND (3M kernels only): 1 = op(A) is a banded window unitary (ZGEMM_A_BAND), 2 = B is (ZGEMM_B_BAND): the MMAs of
// the 4-row k groups that are zero for a whole 8-wide tile of U are skipped (exact; table in the descriptor).
template <int BM, int BN, int WM, int WN, int OPA, int OPB, bool M3, int MINB, int BAND = 0>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB)
zgemm_grouped_kernel(const ZGemmProblem* __restrict__ probs, cplx alpha, cplx beta) {
    constexpr int NTHR = (BM / WM) * (BN / WN) * 32;
    constexpr int WNC = BN / WN;                 // warps along n
    constexpr int MT = WM / 8, NT = WN / 8;      // m8n8 tiles per warp
    constexpr int LDA_S = BM + 2, LDB_S = BN + 2;
    constexpr int A_STAGE = BK * LDA_S, B_STAGE = BK * LDB_S;
    extern __shared__ __align__(16) char smem_raw[];
    cplx* As = reinterpret_cast<cplx*>(smem_raw);
    cplx* Bs = As + STAGES * A_STAGE;

    const ZGemmProblem p = probs[blockIdx.y];
    if (p.M <= 0 || p.N <= 0) return;
    const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
    if ((int)blockIdx.x >= tiles_m * tiles_n) return;
    const int m0 = ((int)blockIdx.x / tiles_n) * BM, n0 = ((int)blockIdx.x % tiles_n) * BN;
    if ((p.flags & ZGEMM_C_UPPER) && m0 >= n0 + BN) return;             // tile entirely below the diagonal
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WNC, wn = warp % WNC;
    // triangular op(B): rows k >= n0 + BN of this column tile are zero -- shorten the k loop (exact: skips zeros)
    const int kmax = ((p.flags & ZGEMM_B_UPPER) && n0 + BN < p.K) ? n0 + BN : p.K;
    const int nk = (kmax + BK - 1) / BK;

    TileLoader<BM, NTHR, OPA == 0> la;
    TileLoader<BN, NTHR, OPB != 0> lb;
    la.init(p.A, p.lda, m0, p.M, kmax, tid);
    lb.init(p.B, p.ldb, n0, p.N, kmax, tid);
    auto load_stage = [&](int stage, int kt) {
        la.load(As + stage * A_STAGE, kt * BK);
        lb.load(Bs + stage * B_STAGE, kt * BK);
    };

    // 4M: acc1 = re, acc2 = im.   3M: acc1 = P1, acc2 = P2, acc3 = P3.
    double acc1[MT][NT][2], acc2[MT][NT][2], acc3[M3 ? MT : 1][M3 ? NT : 1][2];
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) { if (s < nk) load_stage(s, s); cp_async_commit(); }

    // band table of this warp's four 8-wide tiles of U, one byte each (all-active when the problem carries no table)
    unsigned band_lo4 = 0u, band_hi4 = 0xffffffffu;
    if (BAND != 0 && (p.flags & (BAND == 1 ? ZGEMM_A_BAND : ZGEMM_B_BAND))) {
        const int t0 = (BAND == 1) ? (m0 + wm * WM) / 8 : (n0 + wn * WN) / 8;
        band_lo4 = 0u; band_hi4 = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = t0 + q;
            band_lo4 |= (unsigned)(t < 8 ? p.klo[t] : 0) << (8 * q);
            band_hi4 |= (unsigned)(t < 8 ? p.khi[t] : 0) << (8 * q);
        }
    }

    // beta != 0 (rank-k updates C -= A B): the old C tile is folded into the INITIAL accumulators as
    // (beta/alpha) C, loaded here -- all loads independent and in flight together with the first operand
    // stages -- instead of a load -> fma -> store chain per element in the epilogue (the compiler cannot
    // hoist those loads above the stores, which made short-K updates latency-bound).
    const bool use_beta = !(beta.x == 0.0 && beta.y == 0.0);
    const bool pre_beta = use_beta && !(alpha.x == 0.0 && alpha.y == 0.0);
    const cplx boa = pre_beta ? cdiv(beta, alpha) : C(0, 0);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int row = m0 + wm * WM + i * 8 + fr;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = n0 + wn * WN + j * 8 + fk * 2 + e;
                cplx c0 = C(0, 0);
                if (pre_beta && row < p.M && col < p.N) c0 = cmul(boa, p.C[(size_t)row * p.ldc + col]);
                acc1[i][j][e] = c0.x;
                if (M3) { acc2[i][j][e] = 0.0; acc3[M3 ? i : 0][M3 ? j : 0][e] = c0.x + c0.y; }
                else acc2[i][j][e] = c0.y;
            }
        }
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        { int nx = kt + STAGES - 1; if (nx < nk) load_stage(nx % STAGES, nx); cp_async_commit(); }
        const cplx* as = As + (kt % STAGES) * A_STAGE + wm * WM + fr;
        const cplx* bs = Bs + (kt % STAGES) * B_STAGE + wn * WN + fr;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            cplx a[MT], b[NT];
#pragma unroll
            for (int t = 0; t < MT; ++t) { a[t] = as[(kk + fk) * LDA_S + t * 8]; if (OPA == 2) a[t].y = -a[t].y; }
#pragma unroll
            for (int t = 0; t < NT; ++t) { b[t] = bs[(kk + fk) * LDB_S + t * 8]; if (OPB == 2) b[t].y = -b[t].y; }
            if (M3) {
                // three real MMAs per complex tile; each pass touches MT*NT independent accumulators
                unsigned act = 0xfu;
                if (BAND != 0) {
                    const unsigned k4 = (unsigned)(kt * BK + kk) >> 2;
                    act = 0u;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (k4 >= ((band_lo4 >> (8 * q)) & 0xffu) && k4 < ((band_hi4 >> (8 * q)) & 0xffu)) act |= 1u << q;
                }
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    if (BAND == 1 && !((act >> i) & 1u)) continue;
#pragma unroll
                    for (int j = 0; j < NT; ++j) { if (BAND == 2 && !((act >> j) & 1u)) continue; dmma(acc1[i][j][0], acc1[i][j][1], a[i].x, b[j].x); }
                }
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    if (BAND == 1 && !((act >> i) & 1u)) continue;
#pragma unroll
                    for (int j = 0; j < NT; ++j) { if (BAND == 2 && !((act >> j) & 1u)) continue; dmma(acc2[i][j][0], acc2[i][j][1], a[i].y, b[j].y); }
                }
                double bsum[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) bsum[j] = b[j].x + b[j].y;
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    if (BAND == 1 && !((act >> i) & 1u)) continue;
                    const double asum = a[i].x + a[i].y;
#pragma unroll
                    for (int j = 0; j < NT; ++j) { if (BAND == 2 && !((act >> j) & 1u)) continue; dmma(acc3[M3 ? i : 0][M3 ? j : 0][0], acc3[M3 ? i : 0][M3 ? j : 0][1], asum, bsum[j]); }
                }
            } else {
                // four real MMAs per complex tile, issued pass by pass so that the two MMAs that accumulate into
                // the same registers are MT*NT instructions apart (back-to-back they serialise on the DMMA latency)
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma(acc1[i][j][0], acc1[i][j][1], a[i].x, b[j].x);
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma(acc2[i][j][0], acc2[i][j][1], a[i].x, b[j].y);
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    const double nai = -a[i].y;
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma(acc1[i][j][0], acc1[i][j][1], nai, b[j].y);
                }
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma(acc2[i][j][0], acc2[i][j][1], a[i].y, b[j].x);
            }
        }
    }
    cp_async_wait<0>();

    const bool post_beta = use_beta && !pre_beta;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int row = m0 + wm * WM + i * 8 + fr;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = n0 + wn * WN + j * 8 + fk * 2 + e;
                if (col >= p.N) continue;
                cplx* dst = p.C + (size_t)row * p.ldc + col;
                cplx r;
                if (M3) r = C(acc1[i][j][e] - acc2[i][j][e], acc3[M3 ? i : 0][M3 ? j : 0][e] - acc1[i][j][e] - acc2[i][j][e]);
                else r = C(acc1[i][j][e], acc2[i][j][e]);
                cplx v = cmul(alpha, r);
                if (post_beta) v = cadd(v, cmul(beta, *dst));
                *dst = v;
            }
        }
    }
}

__global__ void fill_strided_kernel(ZGemmProblem* probs, int batch, const cplx* A, const cplx* B, cplx* C,
                                    long long sa, long long sb, long long sc, int M, int N, int K, int lda, int ldb, int ldc, int flags) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    ZGemmProblem p;
    p.A = A + (size_t)b * sa; p.B = B + (size_t)b * sb; p.C = C + (size_t)b * sc;
    p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.flags = flags;
    probs[b] = p;
}

template <int BM, int BN, int WM, int WN, int OPA, int OPB, bool M3, int MINB, int BAND = 0>
cudaError_t launch_cfg(const ZGemmProblem* probs, int nprob, int max_tiles, cplx alpha, cplx beta, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * BK * ((BM + 2) + (BN + 2)) * sizeof(cplx);
    constexpr int nthr = (BM / WM) * (BN / WN) * 32;
    // the attribute is per DEVICE: remember it per device (idempotent; benign if raced between host threads)
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(zgemm_grouped_kernel<BM, BN, WM, WN, OPA, OPB, M3, MINB, BAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    dim3 grid(max_tiles, nprob);
    zgemm_grouped_kernel<BM, BN, WM, WN, OPA, OPB, M3, MINB, BAND><<<grid, nthr, smem, st>>>(probs, alpha, beta);
    return cudaGetLastError();
}

// all nine op combinations (the two original large tiles)
template <int BM, int BN, int WM, int WN, bool M3, int MINB>
cudaError_t launch_ops9(int opa, int opb, const ZGemmProblem* probs, int nprob, int max_tiles, cplx alpha, cplx beta, cudaStream_t st) {
    switch (opa * 3 + opb) {
        case 0: return launch_cfg<BM, BN, WM, WN, 0, 0, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
        case 1: return launch_cfg<BM, BN, WM, WN, 0, 1, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
        case 2: return launch_cfg<BM, BN, WM, WN, 0, 2, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
        case 3: return launch_cfg<BM, BN, WM, WN, 1, 0, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
        case 4: return launch_cfg<BM, BN, WM, WN, 1, 1, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
        case 5: return launch_cfg<BM, BN, WM, WN, 1, 2, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
        case 6: return launch_cfg<BM, BN, WM, WN, 2, 0, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
        case 7: return launch_cfg<BM, BN, WM, WN, 2, 1, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
        default: return launch_cfg<BM, BN, WM, WN, 2, 2, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
    }
}
// the three op combinations the RCWA path uses: (N,N), (N,H), (H,N)
template <int BM, int BN, int WM, int WN, bool M3, int MINB>
cudaError_t launch_ops3(int opa, int opb, const ZGemmProblem* probs, int nprob, int max_tiles, cplx alpha, cplx beta, cudaStream_t st) {
    if (opa == 0 && opb == 0) return launch_cfg<BM, BN, WM, WN, 0, 0, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
    if (opa == 0 && opb == 2) return launch_cfg<BM, BN, WM, WN, 0, 2, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
    if (opa == 2 && opb == 0) return launch_cfg<BM, BN, WM, WN, 2, 0, M3, MINB>(probs, nprob, max_tiles, alpha, beta, st);
    return cudaErrorNotSupported;
}

inline void tile_dims(int cfg, int& bm, int& bn) {
    switch (cfg & 7) {
        case GEMM_TILE_64x128: bm = 64; bn = 128; break;
        case GEMM_TILE_128x64: bm = 128; bn = 64; break;
        case GEMM_TILE_64x64: bm = 64; bn = 64; break;
        case GEMM_TILE_128x32: bm = 128; bn = 32; break;
        default: bm = 32; bn = 128; break;
    }
}

}  // namespace

namespace rcwa {

int gemm_tiles(int tile_cfg, int M, int N) {
    int bm, bn;
    tile_dims(tile_cfg, bm, bn);
    return ((M + bm - 1) / bm) * ((N + bn - 1) / bn);
}

bool gemm_cfg_supports(int tile_cfg, int opa, int opb) {
    if (tile_cfg == GEMM_TILE_64x128 || tile_cfg == GEMM_TILE_128x64) return true;
    return (opa == 0 && opb == 0) || (opa == 0 && opb == 2) || (opa == 2 && opb == 0);
}

int gemm_pick_cfg_base(int opa, int opb, int M, int N, int K);

cudaError_t zgemm_grouped(int tile_cfg, int opa, int opb, const ZGemmProblem* probs, int nprob, int max_tiles,
                          cplx alpha, cplx beta, cudaStream_t st) {
    if (nprob <= 0 || max_tiles <= 0) return cudaSuccess;
    switch (tile_cfg) {
        case GEMM_TILE_64x128: return launch_ops9<64, 128, 32, 32, false, 1>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        case GEMM_TILE_128x64: return launch_ops9<128, 64, 32, 32, false, 1>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        case GEMM_TILE_64x64: return launch_ops3<64, 64, 32, 32, false, 2>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        case GEMM_TILE_128x32: return launch_ops3<128, 32, 32, 32, false, 2>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        case GEMM_TILE_32x128: return launch_ops3<32, 128, 32, 32, false, 2>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        case GEMM_TILE_64x128 | GEMM_M3: return launch_ops3<64, 128, 32, 32, true, 1>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        case GEMM_TILE_128x64 | GEMM_M3: return launch_ops3<128, 64, 32, 32, true, 1>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        case GEMM_TILE_64x64 | GEMM_M3: return launch_ops3<64, 64, 32, 32, true, 2>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        case GEMM_TILE_64x64 | GEMM_M3 | GEMM_BAND:      // QR window updates: U^H * rows (band on A) or panel * U (band on B)
            if (opa == 2 && opb == 0) return launch_cfg<64, 64, 32, 32, 2, 0, true, 2, 1>(probs, nprob, max_tiles, alpha, beta, st);
            if (opa == 0 && opb == 0) return launch_cfg<64, 64, 32, 32, 0, 0, true, 2, 2>(probs, nprob, max_tiles, alpha, beta, st);
            return cudaErrorNotSupported;
        case GEMM_TILE_128x32 | GEMM_M3: return launch_ops3<128, 32, 32, 32, true, 2>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        case GEMM_TILE_32x128 | GEMM_M3: return launch_ops3<32, 128, 32, 32, true, 2>(opa, opb, probs, nprob, max_tiles, alpha, beta, st);
        default: return cudaErrorInvalidValue;
    }
}

// Shape heuristic (measured on B200, profiles/): short-K products are prologue/epilogue bound with one
// 256-thread CTA per SM, so they use the 128-thread tiles (two CTAs per SM overlap each other's load and
// store phases); products with one dimension <= 32 use the matching skinny tile instead of computing a
// half-empty 64-wide one.
static int g_tune[16] = {1, GEMM_TILE_64x64, GEMM_TILE_64x64, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
// tuning knobs (process-wide, set before use; not part of the numerical contract):
//   0: use the 3-multiplication complex product where a kernel exists      (default 1)
//   1: tile of the QR row-panel updates   2: tile of the QR column/Z updates
//   3: use the short-K / skinny tiles in the automatic choice               (default 1)
//   4: the QR pass kernel claims a whole SM per matrix (no GEMM CTA co-resident)  (default 0)
//   5 / 6 / 7: per-launch count limits of the QR pass: small-Schur rotations, AED swaps, AED restore steps (0 = default)
//   8: time budget of a serial QR slice in microseconds (0 = default 90, < 0 = count limits only)
//   9: number of independently pipelined matrix groups in the QR phase (0 = default: 2, or one per wave of 148 matrices up to 4)
//  10: skip the zero k groups of the banded window unitaries in the QR update GEMMs (default 0: measured no gain --
//      34 % fewer MMAs, same time: those launches are bound by their load/store phases and by sharing SMs with the pass kernels)
//  11: run the Hessenberg phase as two staggered half batches on two streams (default 0: measured 9 % SLOWER --
//      the latency-bound per-column kernel does not shrink with the batch and the halves do not stay in anti-phase)
//  12: triangular solves of the S-matrix stage on tcgen05 too (default 0: measured slower, profiles/r2_summary.md section 3)
//  13: QR pass as two launches per iteration, bulge-chase windows / small dense solves (0 = automatic: when the batch
//      exceeds two waves of SMs, 1 = never, 2 = always)
//  14: 2 = replay the QR loop of a group from a CUDA graph of 8 iterations instead of enqueuing every launch (default off:
//      measured no gain with the default group count, eig.cu)
//  15: AED window of the QR phase (0 = automatic: 24 for n <= 1024, else 32; at most 48)
void gemm_set_tuning(int key, int value) { if (key >= 0 && key < 16) g_tune[key] = value; }
int gemm_get_tuning(int key) { return (key >= 0 && key < 16) ? g_tune[key] : 0; }

int gemm_pick_cfg(int opa, int opb, int M, int N, int K) {
    const int big = (gemm_tiles(GEMM_TILE_64x128, M, N) * 64 * 128 <= gemm_tiles(GEMM_TILE_128x64, M, N) * 128 * 64) ? GEMM_TILE_64x128 : GEMM_TILE_128x64;
    int cfg = gemm_pick_cfg_base(opa, opb, M, N, K);
    if (!g_tune[3]) cfg = big;
    if (g_tune[0] && gemm_cfg_supports(GEMM_TILE_64x64, opa, opb)) cfg |= GEMM_M3;
    return cfg;
}

int gemm_pick_cfg_base(int opa, int opb, int M, int N, int K) {
    // measured table: profiles/r1b_gemm_probe.json (B200, n = 1922).  The 128-thread tiles win on every
    // shape of the path, the 32x128 one by the widest margin (its B-operand copies are 2 KB contiguous rows).
    (void)K;
    int cfg;
    if (N <= 32 && M > 32) cfg = GEMM_TILE_128x32;
    else if (M <= 32) cfg = GEMM_TILE_32x128;
    else if (N <= 64) cfg = GEMM_TILE_64x64;
    else cfg = GEMM_TILE_32x128;
    if (!gemm_cfg_supports(cfg, opa, opb))
        cfg = (gemm_tiles(GEMM_TILE_64x128, M, N) * 64 * 128 <= gemm_tiles(GEMM_TILE_128x64, M, N) * 128 * 64) ? GEMM_TILE_64x128 : GEMM_TILE_128x64;
    return cfg;
}

cudaError_t zgemm_strided_cfg(int cfg, int opa, int opb, int M, int N, int K, cplx alpha, const cplx* A, int lda, long long sa,
                              const cplx* B, int ldb, long long sb, cplx beta, cplx* Cm, int ldc, long long sc,
                              int batch, ZGemmProblem* scratch, cudaStream_t st, int flags) {
    if (batch <= 0 || M <= 0 || N <= 0) return cudaSuccess;
    if (cfg < 0) cfg = gemm_pick_cfg(opa, opb, M, N, K);
    if (!gemm_cfg_supports(cfg & 7, opa, opb) || ((cfg & GEMM_M3) && !gemm_cfg_supports(GEMM_TILE_64x64, opa, opb))) return cudaErrorNotSupported;
    fill_strided_kernel<<<(batch + 127) / 128, 128, 0, st>>>(scratch, batch, A, B, Cm, sa, sb, sc, M, N, K, lda, ldb, ldc, flags);
    return zgemm_grouped(cfg, opa, opb, scratch, batch, gemm_tiles(cfg, M, N), alpha, beta, st);
}

cudaError_t zgemm_strided(int opa, int opb, int M, int N, int K, cplx alpha, const cplx* A, int lda, long long sa,
                          const cplx* B, int ldb, long long sb, cplx beta, cplx* Cm, int ldc, long long sc,
                          int batch, ZGemmProblem* scratch, cudaStream_t st, int flags) {
    return zgemm_strided_cfg(-1, opa, opb, M, N, K, alpha, A, lda, sa, B, ldb, sb, beta, Cm, ldc, sc, batch, scratch, st, flags);
}

}  // namespace rcwa
