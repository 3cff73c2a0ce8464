// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) GEMM used by the grid quadrature (generic N) and the AO->MO transforms.
//
//   C[m,n] (+)= sum_k  A(m,k) * scale[k] * B(k,n)
//
// tcgen05 has no FP64 kind and wgmma does not exist on sm_100, so FP64 tensor work on B200 is the warp-level
// DMMA family.  CTA tile BM x 64 x 16 (BM = 64 or 32), 4 warps as 2 x 2, each warp a (BM/2) x 32 sub-tile
// (m8n8 accumulators in registers), operands staged through shared memory in k-major layout with a row stride
// == 4 (mod 16): the four k-rows of a fragment land on disjoint banks, so fragment loads are conflict free.
// Global -> shared staging is register double-buffered (next tile's loads are in flight during the MMAs).
// B operand addressing modes (the MP2 transforms read symmetric-packed intermediates without expanding them):
//   B_ROWS   : B(k,n) = B[k*sb_k + n]
//   B_KOFF   : B(k,n) = B[koff[k] + n]                    (row offsets from a table, e.g. pair(p,q)*npair)
//   B_SYMPK  : B(k,n) = B[pair(max(k,n), min(k,n))]       (symmetric matrix stored as a packed lower triangle)
#pragma once
#include "common.cuh"

namespace sqcc {

enum GemmEpilogue {
    EPI_STORE = 0,       // C = acc
    EPI_ATOMIC = 1,      // C += acc (split-K partial sums, red.global.add.f64)
    EPI_ROWDOT = 2       // out[m] += sum_n acc[m,n] * X[m,n]   (X has the shape/strides of C)
};
enum GemmBMode { B_ROWS = 0, B_KOFF = 1, B_SYMPK = 2 };

struct GemmArgs {
    const double *A;
    const double *B;
    double *C;            // EPI_ROWDOT: out vector [M] (per batch: + z * batch_c)
    const double *X;      // EPI_ROWDOT only
    const double *scale;  // optional per-k scale (nullptr = 1)
    const int64_t *koff;  // B_KOFF only: [batch][K] row offsets
    int M, N, K;
    int64_t sa_m, sa_k;   // A(m,k) = A[m*sa_m + k*sa_k]
    int64_t sb_k;         // B_ROWS: B(k,n) = B[k*sb_k + n]
    int64_t sc_m;         // C(m,n) = C[m*sc_m + n]   (and X)
    int batch;            // blockIdx.z = batch * splitk + split
    int64_t batch_a, batch_b, batch_c, batch_scale, batch_koff;
    int splitk;           // K is cut into `splitk` contiguous ranges (EPI_ATOMIC / EPI_ROWDOT only)
};

constexpr int GEMM_BN = 64, GEMM_BK = 16, GEMM_THREADS = 128;

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <bool A_MFAST, int EPI, int BMODE, int BM>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_f64_kernel(const GemmArgs g)
{
    constexpr int LDA_S = BM + 4, LDB_S = GEMM_BN + 4;   // 68 / 36: == 4 (mod 16)
    constexpr int MT = BM / 16;                          // m8 tiles per warp
    constexpr int EA = BM * GEMM_BK / GEMM_THREADS;      // A elements staged per thread
    __shared__ double As[2][GEMM_BK][LDA_S];
    __shared__ double Bs[2][GEMM_BK][LDB_S];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * (BM / 2), wn = (warp & 1) * 32;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * GEMM_BN;
    const int z = blockIdx.z, bz = z / g.splitk, sz = z % g.splitk;
    const double *A = g.A + (int64_t)bz * g.batch_a;
    const double *B = g.B + (int64_t)bz * g.batch_b;
    const double *scale = g.scale ? g.scale + (int64_t)bz * g.batch_scale : nullptr;
    const int64_t *koff = BMODE == B_KOFF ? g.koff + (int64_t)bz * g.batch_koff : nullptr;
    // split-K range (multiples of BK)
    const int ktiles = (g.K + GEMM_BK - 1) / GEMM_BK;
    const int per = (ktiles + g.splitk - 1) / g.splitk;
    const int kt_lo = sz * per, kt_hi = (kt_lo + per) < ktiles ? (kt_lo + per) : ktiles;

    double acc[MT][4][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    double ra[EA], rb[8];
    auto load_tile = [&](int kt) {
        const int kbase = kt * GEMM_BK;
#pragma unroll
        for (int e = 0; e < EA; ++e) {
            const int idx = tid + e * GEMM_THREADS;
            int m, k;
            if (A_MFAST) { m = idx % BM; k = idx / BM; } else { k = idx & 15; m = idx >> 4; }
            const int gm = m0 + m, gk = kbase + k;
            double v = 0.0;
            if (gm < g.M && gk < g.K) {
                v = __ldg(A + (int64_t)gm * g.sa_m + (int64_t)gk * g.sa_k);
                if (scale) v *= __ldg(scale + gk);
            }
            ra[e] = v;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + e * GEMM_THREADS;
            const int n = idx & 63, kb = idx >> 6;
            const int gn = n0 + n, gkb = kbase + kb;
            double v = 0.0;
            if (gn < g.N && gkb < g.K) {
                if (BMODE == B_ROWS) v = __ldg(B + (int64_t)gkb * g.sb_k + gn);
                else if (BMODE == B_KOFF) v = __ldg(B + __ldg(koff + gkb) + gn);
                else {
                    const int64_t hi = gkb > gn ? gkb : gn, lo = gkb > gn ? gn : gkb;
                    v = __ldg(B + hi * (hi + 1) / 2 + lo);
                }
            }
            rb[e] = v;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int e = 0; e < EA; ++e) {
            const int idx = tid + e * GEMM_THREADS;
            int m, k;
            if (A_MFAST) { m = idx % BM; k = idx / BM; } else { k = idx & 15; m = idx >> 4; }
            As[buf][k][m] = ra[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + e * GEMM_THREADS;
            Bs[buf][idx >> 6][idx & 63] = rb[e];
        }
    };

    if (kt_lo < kt_hi) {
        load_tile(kt_lo);
        store_tile(0);
        __syncthreads();
        for (int kt = kt_lo; kt < kt_hi; ++kt) {
            const int buf = (kt - kt_lo) & 1;
            if (kt + 1 < kt_hi) load_tile(kt + 1);
#pragma unroll
            for (int ks = 0; ks < GEMM_BK; ks += 4) {
                double af[MT], bf[4];
#pragma unroll
                for (int i = 0; i < MT; ++i) af[i] = As[buf][ks + (lane & 3)][wm + i * 8 + (lane >> 2)];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = Bs[buf][ks + (lane & 3)][wn + j * 8 + (lane >> 2)];
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            if (kt + 1 < kt_hi) {
                store_tile(buf ^ 1);
                __syncthreads();
            }
        }
    }

    // epilogue: lane holds C[row = lane/4][col = 2*(lane%4) + {0,1}] of every 8x8 tile
    if (EPI == EPI_ROWDOT) {
        double *out = g.C + (int64_t)bz * g.batch_c;
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int gm = m0 + wm + i * 8 + (lane >> 2);
            double part = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = n0 + wn + j * 8 + 2 * (lane & 3);
                if (gm < g.M) {
                    if (gn < g.N) part = fma(acc[i][j][0], __ldg(g.X + (int64_t)gm * g.sc_m + gn), part);
                    if (gn + 1 < g.N) part = fma(acc[i][j][1], __ldg(g.X + (int64_t)gm * g.sc_m + gn + 1), part);
                }
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if ((lane & 3) == 0 && gm < g.M) atomicAdd(out + gm, part);
        }
    } else {
        double *C = g.C + (int64_t)bz * g.batch_c;
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int gm = m0 + wm + i * 8 + (lane >> 2);
            if (gm >= g.M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = n0 + wn + j * 8 + 2 * (lane & 3);
                double *dst = C + (int64_t)gm * g.sc_m + gn;
                if (EPI == EPI_ATOMIC) {
                    if (gn < g.N) atomicAdd(dst, acc[i][j][0]);
                    if (gn + 1 < g.N) atomicAdd(dst + 1, acc[i][j][1]);
                } else {
                    if (gn < g.N) dst[0] = acc[i][j][0];
                    if (gn + 1 < g.N) dst[1] = acc[i][j][1];
                }
            }
        }
    }
}

// bmode / small-M tiles are only instantiated for the A-transposed STORE form the AO->MO transforms use.
inline int gemm_f64(sqcc_ctx *ctx, const GemmArgs &g, int epi, int bmode = B_ROWS)
{
    SQCC_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.batch > 0 && g.splitk > 0, "bad GEMM shape");
    SQCC_REQUIRE(g.sa_m == 1 || g.sa_k == 1, "GEMM A operand needs a unit stride");
    SQCC_REQUIRE(epi != EPI_STORE || g.splitk == 1, "split-K needs an accumulating epilogue");
    const bool mfast = g.sa_m == 1 && g.sa_k != 1;
    const bool small_m = g.M <= 32 && mfast && epi == EPI_STORE;
    SQCC_REQUIRE(bmode == B_ROWS || (mfast && epi == EPI_STORE), "packed B operands need the A^T / STORE form");
    const int bm = small_m ? 32 : 64;
    dim3 grid((g.N + GEMM_BN - 1) / GEMM_BN, (g.M + bm - 1) / bm, g.batch * g.splitk);
    SQCC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "GEMM grid too large");
#define SQCC_GEMM_LAUNCH(MF, E, BMD, BMV) gemm_f64_kernel<MF, E, BMD, BMV><<<grid, GEMM_THREADS, 0, ctx->stream>>>(g)
    if (mfast && epi == EPI_STORE) {
        if (small_m) {
            if (bmode == B_ROWS) SQCC_GEMM_LAUNCH(true, EPI_STORE, B_ROWS, 32);
            else if (bmode == B_KOFF) SQCC_GEMM_LAUNCH(true, EPI_STORE, B_KOFF, 32);
            else SQCC_GEMM_LAUNCH(true, EPI_STORE, B_SYMPK, 32);
        } else {
            if (bmode == B_ROWS) SQCC_GEMM_LAUNCH(true, EPI_STORE, B_ROWS, 64);
            else if (bmode == B_KOFF) SQCC_GEMM_LAUNCH(true, EPI_STORE, B_KOFF, 64);
            else SQCC_GEMM_LAUNCH(true, EPI_STORE, B_SYMPK, 64);
        }
    } else if (mfast) {
        if (epi == EPI_ATOMIC) SQCC_GEMM_LAUNCH(true, EPI_ATOMIC, B_ROWS, 64);
        else SQCC_GEMM_LAUNCH(true, EPI_ROWDOT, B_ROWS, 64);
    } else {
        if (epi == EPI_STORE) SQCC_GEMM_LAUNCH(false, EPI_STORE, B_ROWS, 64);
        else if (epi == EPI_ATOMIC) SQCC_GEMM_LAUNCH(false, EPI_ATOMIC, B_ROWS, 64);
        else SQCC_GEMM_LAUNCH(false, EPI_ROWDOT, B_ROWS, 64);
    }
#undef SQCC_GEMM_LAUNCH
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

}  // namespace sqcc
