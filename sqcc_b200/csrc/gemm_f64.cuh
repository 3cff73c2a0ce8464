// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) GEMM used by the grid quadrature and the AO->MO transforms.
//
//   C[m,n] (+)= sum_k  A(m,k) * scale[k] * B(k,n)
//
// tcgen05 has no FP64 kind and wgmma does not exist on sm_100, so FP64 tensor work on B200 is the warp-level
// DMMA family.  CTA tile 64x64x16, 4 warps, each warp a 32x32 sub-tile (4x4 m8n8 accumulators in registers),
// operands staged through shared memory in k-major layout with a row stride of 68 doubles (== 4 mod 16):
// the four k-rows of a fragment then land on disjoint banks, so fragment loads are conflict free.
// Global -> shared staging is register double-buffered (next tile's loads are in flight during the MMAs).
#pragma once
#include "common.cuh"

namespace sqcc {

enum GemmEpilogue {
    EPI_STORE = 0,       // C = acc
    EPI_ATOMIC = 1,      // C += acc (split-K partial sums, red.global.add.f64)
    EPI_ROWDOT = 2       // out[m] += sum_n acc[m,n] * X[m,n]   (X has the shape/strides of C)
};

struct GemmArgs {
    const double *A;
    const double *B;
    double *C;            // EPI_ROWDOT: out vector [M] (per batch: + z * batch_c)
    const double *X;      // EPI_ROWDOT only
    const double *scale;  // optional per-k scale (nullptr = 1)
    int M, N, K;
    int64_t sa_m, sa_k;   // A(m,k) = A[m*sa_m + k*sa_k]
    int64_t sb_k;         // B(k,n) = B[k*sb_k + n]
    int64_t sc_m;         // C(m,n) = C[m*sc_m + n]   (and X)
    int batch;            // blockIdx.z = batch * splitk + split
    int64_t batch_a, batch_b, batch_c, batch_scale;
    int splitk;           // K is cut into `splitk` contiguous ranges (EPI_ATOMIC / EPI_ROWDOT only)
};

constexpr int GEMM_BM = 64, GEMM_BN = 64, GEMM_BK = 16, GEMM_LDS = 68, GEMM_THREADS = 128;

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <bool A_MFAST, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_f64_kernel(const GemmArgs g)
{
    __shared__ double As[2][GEMM_BK][GEMM_LDS];
    __shared__ double Bs[2][GEMM_BK][GEMM_LDS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * GEMM_BN;
    const int z = blockIdx.z, bz = z / g.splitk, sz = z % g.splitk;
    const double *A = g.A + (int64_t)bz * g.batch_a;
    const double *B = g.B + (int64_t)bz * g.batch_b;
    const double *scale = g.scale ? g.scale + (int64_t)bz * g.batch_scale : nullptr;
    // split-K range (multiples of BK)
    const int ktiles = (g.K + GEMM_BK - 1) / GEMM_BK;
    const int per = (ktiles + g.splitk - 1) / g.splitk;
    const int kt_lo = sz * per, kt_hi = (kt_lo + per) < ktiles ? (kt_lo + per) : ktiles;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // staging registers: 1024 elements per operand tile / 128 threads = 8 each
    double ra[8], rb[8];
    auto load_tile = [&](int kt) {
        const int kbase = kt * GEMM_BK;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + e * GEMM_THREADS;
            int m, k;
            if (A_MFAST) { m = idx & 63; k = idx >> 6; } else { k = idx & 15; m = idx >> 4; }
            const int gm = m0 + m, gk = kbase + k;
            double v = 0.0;
            if (gm < g.M && gk < g.K) {
                v = __ldg(A + (int64_t)gm * g.sa_m + (int64_t)gk * g.sa_k);
                if (scale) v *= __ldg(scale + gk);
            }
            ra[e] = v;
            const int n = idx & 63, kb = idx >> 6;
            const int gn = n0 + n, gkb = kbase + kb;
            rb[e] = (gn < g.N && gkb < g.K) ? __ldg(B + (int64_t)gkb * g.sb_k + gn) : 0.0;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + e * GEMM_THREADS;
            int m, k;
            if (A_MFAST) { m = idx & 63; k = idx >> 6; } else { k = idx & 15; m = idx >> 4; }
            As[buf][k][m] = ra[e];
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
                double af[4], bf[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = As[buf][ks + (lane & 3)][wm + i * 8 + (lane >> 2)];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = Bs[buf][ks + (lane & 3)][wn + j * 8 + (lane >> 2)];
#pragma unroll
                for (int i = 0; i < 4; ++i)
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
        for (int i = 0; i < 4; ++i) {
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
        for (int i = 0; i < 4; ++i) {
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

inline int gemm_f64(sqcc_ctx *ctx, const GemmArgs &g, int epi)
{
    SQCC_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.batch > 0 && g.splitk > 0, "bad GEMM shape");
    SQCC_REQUIRE(g.sa_m == 1 || g.sa_k == 1, "GEMM A operand needs a unit stride");
    SQCC_REQUIRE(epi != EPI_STORE || g.splitk == 1, "split-K needs an accumulating epilogue");
    dim3 grid((g.N + GEMM_BN - 1) / GEMM_BN, (g.M + GEMM_BM - 1) / GEMM_BM, g.batch * g.splitk);
    SQCC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "GEMM grid too large");
    const bool mfast = g.sa_m == 1 && g.sa_k != 1;
#define SQCC_GEMM_LAUNCH(MF, E) gemm_f64_kernel<MF, E><<<grid, GEMM_THREADS, 0, ctx->stream>>>(g)
    if (mfast) {
        if (epi == EPI_STORE) SQCC_GEMM_LAUNCH(true, EPI_STORE);
        else if (epi == EPI_ATOMIC) SQCC_GEMM_LAUNCH(true, EPI_ATOMIC);
        else SQCC_GEMM_LAUNCH(true, EPI_ROWDOT);
    } else {
        if (epi == EPI_STORE) SQCC_GEMM_LAUNCH(false, EPI_STORE);
        else if (epi == EPI_ATOMIC) SQCC_GEMM_LAUNCH(false, EPI_ATOMIC);
        else SQCC_GEMM_LAUNCH(false, EPI_ROWDOT);
    }
#undef SQCC_GEMM_LAUNCH
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

}  // namespace sqcc
