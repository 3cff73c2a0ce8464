// AO -> MO four-index quarter transforms and the closed-shell MP2 energy on FP64 tensor cores.
//
//   (pq|rs) -> (iq|rs) -> (iq|js) -> (ia|js) -> (ia|jb)        mp2.py:123-170
//   E_corr = sum_ijab (2 x^2 - x x') / (e_i + e_j - e_a - e_b),  x = (ia|jb), x' = (ib|ja)
//                                                               mp2.py:89-99, :240-254
// The reference repeats the identical transform a second time for (ib|ja) (mp2.py:185-228); the result is
// the same array, so it is read transposed instead.  Each quarter is a (batched) DMMA GEMM (gemm_f64.cuh):
//   Q1  [n1 x N^3]        = C1^T [n1 x N] . E [N x N^3]
//   Q2  per (a,q): [n3 x N] = C3^T [n3 x N] . T1_aq [N x N]
//   Q3  per a:  [n2 x n3 N] = C2^T [n2 x N] . T2_a [N x n3 N]
//   Q4  [n1 n2 n3 x n4]   = T3 [n1 n2 n3 x N] . C4 [N x n4]
// E is the full tensor expanded once from the packed resident copy (eri_unpack_full).
#include "gemm_f64.cuh"

namespace sqcc {

static int ensure_mp2work(sqcc_ctx *ctx, size_t bytes)
{
    if (ctx->mp2work_bytes >= bytes) return SQCC_OK;
    if (ctx->d_mp2work) cudaFree(ctx->d_mp2work);
    ctx->d_mp2work = nullptr;
    ctx->mp2work_bytes = 0;
    SQCC_CUDA(cudaMalloc(&ctx->d_mp2work, bytes));
    ctx->mp2work_bytes = bytes;
    return SQCC_OK;
}

int ao2mo(sqcc_ctx *ctx, const double *d_C1, int n1, const double *d_C2, int n2, const double *d_C3, int n3,
          const double *d_C4, int n4, double *d_out)
{
    // C blocks are [N, n_x] row-major with leading dimension ld_x passed through n (columns of the AO x MO matrix):
    // callers pass pointers into the [N,N] coefficient matrix, so the leading dimension is N.
    const int64_t n = ctx->n;
    const int64_t n3_ = n * n * n;
    const size_t full = (size_t)n * n3_;
    const size_t t1 = (size_t)n1 * n3_, t2 = (size_t)n1 * n * n3 * n, t3 = (size_t)n1 * n2 * n3 * n;
    // workspace: [E full][T1][T2] ; T3 reuses T1's space (T1 is dead after Q2)
    size_t t13 = t1 > t3 ? t1 : t3;
    SQCC_TRY(ensure_mp2work(ctx, sizeof(double) * (full + t13 + t2)));
    double *E = ctx->d_mp2work, *T1 = E + full, *T2 = T1 + t13, *T3 = T1;
    SQCC_TRY(eri_unpack_full(ctx, E));

    GemmArgs g{};
    g.scale = nullptr;
    g.X = nullptr;
    g.splitk = 1;
    g.batch_scale = 0;
    // Q1: T1[a, (q r s)] = sum_p C1[p, a] E[p, (q r s)]
    g.A = d_C1; g.sa_m = 1; g.sa_k = n;
    g.B = E; g.sb_k = n3_;
    g.C = T1; g.sc_m = n3_;
    g.M = n1; g.N = (int)0; g.K = (int)n;
    g.batch = 1; g.batch_a = g.batch_b = g.batch_c = 0;
    // N^3 can exceed what one launch's grid.x handles comfortably; chunk the column range
    {
        const int64_t chunk = (int64_t)1 << 24;
        for (int64_t c0 = 0; c0 < n3_; c0 += chunk) {
            GemmArgs h = g;
            h.B = E + c0;
            h.C = T1 + c0;
            h.N = (int)((n3_ - c0) < chunk ? (n3_ - c0) : chunk);
            SQCC_TRY(gemm_f64(ctx, h, EPI_STORE));
        }
    }
    // Q2: T2[a, q, c, s] = sum_r C3[r, c] T1[a, q, r, s]     batch over (a, q)
    {
        const int64_t nb = (int64_t)n1 * n;
        const int64_t bmax = 32768;
        for (int64_t b0 = 0; b0 < nb; b0 += bmax) {
            GemmArgs h = g;
            h.A = d_C3; h.sa_m = 1; h.sa_k = n;
            h.B = T1 + b0 * n * n; h.sb_k = n;
            h.C = T2 + b0 * n3 * n; h.sc_m = n;
            h.M = n3; h.N = (int)n; h.K = (int)n;
            h.batch = (int)((nb - b0) < bmax ? (nb - b0) : bmax);
            h.batch_a = 0; h.batch_b = n * n; h.batch_c = (int64_t)n3 * n;
            SQCC_TRY(gemm_f64(ctx, h, EPI_STORE));
        }
    }
    // Q3: T3[a, b, c, s] = sum_q C2[q, b] T2[a, q, c, s]     batch over a
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
        const int64_t rmax = (int64_t)65535 * GEMM_BM;
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
