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
//   B_PAIRQ  : B(k,n) = B[pair(max(k,q), min(k,q)) * sb_k + n], q = batch index  (row pair(k,q) of the pair matrix:
//              the first AO->MO quarter; same rows as a B_KOFF table, but computed instead of loaded)
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace sqcc {

enum GemmEpilogue {
    EPI_STORE = 0,       // C = acc
    EPI_ATOMIC = 1,      // C += acc (split-K partial sums, red.global.add.f64)
    EPI_ROWDOT = 2       // out[m] += sum_n acc[m,n] * X[m,n]   (X has the shape/strides of C)
};
enum GemmBMode { B_ROWS = 0, B_KOFF = 1, B_SYMPK = 2, B_PAIRQ = 3 };

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
    int symmetric;        // EPI_ATOMIC, M == N, A == B^T: only the tiles on and above the diagonal are computed and their
                          // sums added to both triangles (the grid quadrature phi^T diag(w v) phi)
    int ktri;             // B is upper triangular (B(k,n) = 0 for k > n): a column block only runs over k < n0 + BN
};

constexpr int GEMM_BN = 64, GEMM_BK = 16, GEMM_THREADS = 128;

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Implemented in gemm_f64.cu (one translation unit instantiates the kernels).
int gemm_f64(sqcc_ctx *ctx, const GemmArgs &g, int epi, int bmode = B_ROWS);

}  // namespace sqcc
