// AO -> MO four-index quarter transforms and the closed-shell MP2 energy on FP64 tensor cores.
//
//   (pq|rs) -> (iq|rs) -> (iq|js) -> (ia|js) -> (ia|jb)        mp2.py:123-170
//   E_corr = sum_ijab (2 x^2 - x x') / (e_i + e_j - e_a - e_b),  x = (ia|jb), x' = (ib|ja)
//                                                               mp2.py:89-99, :240-254
// The reference repeats the identical transform a second time for (ib|ja) (mp2.py:185-228); the result is the same
// array, so it is read transposed instead.  The full N^4 tensor is never materialised:
//   S   [npair x npair]      the packed lower triangle of the pair matrix, symmetrised by a tiled transpose
//   Q1  per q: [n1 x npair]  = C1^T [n1 x N] . S[pair(p,q), :]      (B rows addressed by computed pair index;
//                              only canonical pairs rs are transformed: half the flops of the reference's step 1)
//   Q2  per (a,q): [n3 x N]  = C3^T [n3 x N] . T1_aq (symmetric in (r,s), read from its packed lower triangle)
//   Q3  per a:  [n2 x n3 N]  = C2^T [n2 x N] . T2_a [N x n3 N]
//   Q4  [n1 n2 n3 x n4]      = T3 [n1 n2 n3 x N] . C4 [N x n4]
// Each quarter is a (batched) DMMA GEMM (gemm_f64.cuh).
#include <math.h>
#include <stdlib.h>

#include "gemm_f64.cuh"

namespace sqcc {

static int ensure_mp2work(sqcc_ctx *ctx, size_t bytes)
{
    if (ctx->mp2work_bytes >= bytes) return SQCC_OK;
    if (ctx->d_mp2work) cudaFree(ctx->d_mp2work);
    ctx->d_mp2work = nullptr;
    ctx->mp2work_bytes = 0;
    // the symmetrised pair matrix is npair^2 doubles (4.9 GB at N = 222, ~120 GB at N = 494): say so instead of failing
    // inside cudaMalloc
    size_t free_b = 0, total_b = 0;
    SQCC_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (bytes > free_b) {
        static char msg[192];
        snprintf(msg, sizeof(msg), "AO->MO transform at N = %d needs %.1f GB of work space (pair matrix + intermediates), "
                                   "%.1f GB free on the device", ctx->n, (double)bytes / 1e9, (double)free_b / 1e9);
        SQCC_REQUIRE(false, msg);
    }
    SQCC_CUDA(cudaMalloc(&ctx->d_mp2work, bytes));
    ctx->mp2work_bytes = bytes;
    return SQCC_OK;
}

__device__ __forceinline__ void pair_unrank_dev(int64_t ij, int &i, int &j)
{
    int64_t t = (int64_t)((sqrt(8.0 * (double)ij + 1.0) - 1.0) * 0.5);
    while (t * (t + 1) / 2 > ij) --t;
    while ((t + 1) * (t + 2) / 2 <= ij) ++t;
    i = (int)t;
    j = (int)(ij - t * (t + 1) / 2);
}

// S[pq][rs] = (pq|rs) for all pair indices: 128 x 32 tiles of the packed lower triangle are read along rows (mostly
// contiguous in the native layout), written as they are and written transposed through shared memory.  Tall tiles
// amortise the per-thread index work (pair unranking of the column, sub-row offset) over 16 rows per thread.
constexpr int SYM_TR = 128, SYM_TC = 32;

__global__ void __launch_bounds__(256) symmetrize_pairs_kernel(const double *__restrict__ P,
                                                               const int64_t *__restrict__ ibase, int n, int64_t npair,
                                                               double *__restrict__ S)
{
    __shared__ double tile[SYM_TR][SYM_TC + 1];
    const int64_t R0 = (int64_t)blockIdx.y * SYM_TR, C0 = (int64_t)blockIdx.x * SYM_TC;
    if (C0 > R0 + SYM_TR - 1) return;   // tile entirely above the diagonal
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t rs = C0 + lane;
    int k = 0, l = 0;
    if (rs < npair) pair_unrank_dev(rs, k, l);
    // native v2 address (layout.cuh) = tile(i, j) + imax * colA + colB + slot(i, j) * colw + colx: the column part is
    // computed once per thread, the row part once per row
    const int cch = l >> 6, cm = k - LAY_CHUNK * cch;
    const int colw = lay_w(cm), colx = l - LAY_CHUNK * cch;
    const int64_t colA = (int64_t)LAY_ROWS * 64 * cch;
    const int64_t colB = LAY_ROWS * (64 * (-(int64_t)23 * cch - 32 * (int64_t)cch * (cch - 1)) + lay_S(cm));
    const double fkl = k == l ? 2.0 : 1.0;          // 1 / degeneracy factors are exact powers of two
    // (i, j) of this warp's first row, then stepped by 8 rows at a time (no sqrt per row)
    int i = 0, j = 0;
    pair_unrank_dev(R0 + warp < npair ? R0 + warp : npair - 1, i, j);
#pragma unroll 4
    for (int r = warp; r < SYM_TR; r += 8) {
        const int64_t pq = R0 + r;
        double v = 0.0;
        if (pq < npair && rs <= pq) {
            const double f = fkl * (i == j ? 2.0 : 1.0) * (rs == pq ? 2.0 : 1.0);
            const int ib = lay_iblock(n, i), ihi = lay_ihi(n, ib), imax = ihi - 1;
            const int64_t off = ibase[ib] + (int64_t)(j >> 2) * lay_tile_len(imax) + (int64_t)imax * colA + colB +
                                (int64_t)lay_slot(ihi, i, j) * colw + colx;
            v = P[off] * f;
            S[pq * npair + rs] = v;
        }
        tile[r][lane] = v;
        j += 8;
        while (j > i) { j -= i + 1; ++i; }
    }
    __syncthreads();
    for (int c = warp; c < SYM_TC; c += 8) {
        // elements (pq = R0 + lane + 32 m, rs = C0 + c) go to S[rs][pq]
        const int64_t rs2 = C0 + c;
        if (rs2 >= npair) break;
#pragma unroll
        for (int m = 0; m < SYM_TR / 32; ++m) {
            const int64_t pq = R0 + lane + 32 * m;
            if (pq < npair && rs2 < pq) S[rs2 * npair + pq] = tile[lane + 32 * m][c];
        }
    }
}

int ao2mo(sqcc_ctx *ctx, const double *d_C1, int n1, const double *d_C2, int n2, const double *d_C3, int n3,
          const double *d_C4, int n4, double *d_out)
{
    // C blocks are column blocks of an [N,N] AO x MO matrix: element (p, a) at C[p*N + a].
    SQCC_REQUIRE(ctx->d_eri && ctx->whole,
                 "the AO->MO transform needs an unsharded ERI attached");
    const int64_t n = ctx->n, npair = n * (n + 1) / 2;
    const size_t s_sz = (size_t)npair * npair;
    const size_t t1 = (size_t)n1 * n * npair, t2 = (size_t)n1 * n * n3 * n, t3 = (size_t)n1 * n2 * n3 * n;
    const size_t t13 = t1 > t3 ? t1 : t3;   // T3 reuses T1's space (T1 is dead after Q2)
    SQCC_TRY(ensure_mp2work(ctx, sizeof(double) * (s_sz + t13 + t2)));
    double *S = ctx->d_mp2work, *T1 = S + s_sz, *T2 = T1 + t13, *T3 = T1;

    {   // symmetrised pair matrix
        const dim3 sgrid((unsigned)((npair + SYM_TC - 1) / SYM_TC), (unsigned)((npair + SYM_TR - 1) / SYM_TR));
        SQCC_REQUIRE(sgrid.y <= 65535, "pair matrix too large for one launch");
        symmetrize_pairs_kernel<<<sgrid, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->d_ibase, (int)n, npair, S);
        ctx->launches += 1;
        SQCC_CUDA(cudaGetLastError());
    }

    GemmArgs g{};
    g.scale = nullptr;
    g.X = nullptr;
    g.koff = nullptr;
    g.splitk = 1;
    g.batch_scale = 0;
    g.batch_koff = 0;
    // Q1: T1[a, q, rs] = sum_p C1[p, a] S[pair(p,q), rs]        batch over q, canonical pairs rs only
    {
        GemmArgs h = g;
        h.A = d_C1; h.sa_m = 1; h.sa_k = n;
        h.B = S; h.sb_k = npair;
        h.C = T1; h.sc_m = n * npair;
        h.M = n1; h.N = (int)npair; h.K = (int)n;
        h.batch = (int)n; h.batch_a = 0; h.batch_b = 0; h.batch_c = npair;
        SQCC_REQUIRE(npair < ((int64_t)1 << 31), "npair too large");
        SQCC_TRY(gemm_f64(ctx, h, EPI_STORE, B_PAIRQ));
    }
    // Q2: T2[a, q, c, s] = sum_r C3[r, c] T1[a, q, pair(r,s)]   batch over (a, q); T1_aq read as packed symmetric
    {
        const int64_t nb = (int64_t)n1 * n;
        const int64_t bmax = 32768;
        for (int64_t b0 = 0; b0 < nb; b0 += bmax) {
            GemmArgs h = g;
            h.A = d_C3; h.sa_m = 1; h.sa_k = n;
            h.B = T1 + b0 * npair; h.sb_k = 0;
            h.C = T2 + b0 * n3 * n; h.sc_m = n;
            h.M = n3; h.N = (int)n; h.K = (int)n;
            h.batch = (int)((nb - b0) < bmax ? (nb - b0) : bmax);
            h.batch_a = 0; h.batch_b = npair; h.batch_c = (int64_t)n3 * n;
            SQCC_TRY(gemm_f64(ctx, h, EPI_STORE, B_SYMPK));
        }
    }
    // Q3: T3[a, b, c, s] = sum_q C2[q, b] T2[a, q, c, s]         batch over a
    {
        GemmArgs h = g;
        h.A = d_C2; h.sa_m = 1; h.sa_k = n;
        h.B = T2; h.sb_k = (int64_t)n3 * n;
        h.C = T3; h.sc_m = (int64_t)n3 * n;
        h.M = n2; h.N = (int)(n3 * n); h.K = (int)n;
        h.batch = n1; h.batch_a = 0; h.batch_b = n * n3 * n; h.batch_c = (int64_t)n2 * n3 * n;
        SQCC_TRY(gemm_f64(ctx, h, EPI_STORE));
    }
    // Q4: out[(a b c), d] = sum_s T3[(a b c), s] C4[s, d]
    {
        const int64_t rows = (int64_t)n1 * n2 * n3;
        const int64_t rmax = (int64_t)65535 * 64;
        for (int64_t r0 = 0; r0 < rows; r0 += rmax) {
            GemmArgs h = g;
            h.A = T3 + r0 * n; h.sa_m = n; h.sa_k = 1;
            h.B = d_C4; h.sb_k = n;
            h.C = d_out + r0 * n4; h.sc_m = n4;
            h.M = (int)((rows - r0) < rmax ? (rows - r0) : rmax); h.N = n4; h.K = (int)n;
            h.batch = 1; h.batch_a = h.batch_b = h.batch_c = 0;
            SQCC_TRY(gemm_f64(ctx, h, EPI_STORE));
        }
    }
    return SQCC_OK;
}

// E_corr = sum_{i,a,j,b} (2 x^2 - x x') / (e_i + e_j - e_a - e_b),  x = iajb[i,a,j,b], x' = iajb[i,b,j,a]
__global__ void __launch_bounds__(256) mp2_energy_kernel(const double *__restrict__ iajb, const double *__restrict__ eps,
                                                         int no, int nv, double *__restrict__ ecorr)
{
    const int64_t total = (int64_t)no * nv * no * nv;
    double part = 0.0;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx % nv);
        const int j = (int)((idx / nv) % no);
        const int a = (int)((idx / ((int64_t)nv * no)) % nv);
        const int i = (int)(idx / ((int64_t)nv * no * nv));
        const double x = iajb[idx];
        const double xp = iajb[(((int64_t)i * nv + b) * no + j) * nv + a];
        const double den = eps[i] + eps[j] - eps[no + a] - eps[no + b];
        part += (2.0 * x * x - x * xp) / den;
    }
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(ecorr, s);
    }
}

int mp2(sqcc_ctx *ctx, const double *d_C, const double *d_eps, int n_occ, double *d_ecorr, double *d_iajb)
{
    const int n = ctx->n, nv = n - n_occ;
    const size_t nout = (size_t)n_occ * nv * n_occ * nv;
    double *out = d_iajb;
    if (!out) {
        if (ctx->mp2out_n < nout) {
            if (ctx->d_mp2out) cudaFree(ctx->d_mp2out);
            ctx->d_mp2out = nullptr;
            ctx->mp2out_n = 0;
            SQCC_CUDA(cudaMalloc(&ctx->d_mp2out, sizeof(double) * nout));
            ctx->mp2out_n = nout;
        }
        out = ctx->d_mp2out;
    }
    SQCC_TRY(timer_start(ctx, 4));
    SQCC_TRY(ao2mo(ctx, d_C, n_occ, d_C + n_occ, nv, d_C, n_occ, d_C + n_occ, nv, out));
    SQCC_TRY(timer_stop(ctx, 4));
    SQCC_TRY(timer_start(ctx, 5));
    SQCC_CUDA(cudaMemsetAsync(d_ecorr, 0, sizeof(double), ctx->stream));
    mp2_energy_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(out, d_eps, n_occ, nv, d_ecorr);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    SQCC_TRY(timer_stop(ctx, 5));
    return SQCC_OK;
}

}  // namespace sqcc
