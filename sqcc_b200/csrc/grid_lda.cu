// LDA-exchange grid quadrature on FP64 tensor cores (DMMA GEMMs with fused epilogues).
//
//   rho_n  = phi_n D phi_n^T                     hf_ksdft.py:426-428, :634-636 (RKS), :434-440, :640-646 (UKS)
//   v_n    = -(3/pi)^(1/3) rho^(1/3)             ksdft_functional.py:41-65   (x 2^(1/3) per spin: :98-134)
//   E_x    = -(3/4)(3/pi)^(1/3) sum w rho^(4/3)  ksdft_functional.py:68-95   (x 2^(1/3), both spins: :137-177)
//   V_pq   = sum_n phi_np (w_n v_n) phi_nq       hf_ksdft.py:456-461 (RKS), :469-476 (UKS)
//
// rho is a GEMM (G x N)(N x N) whose epilogue dots each output row with the phi row (the G x N product is
// never written); V_xc is a split-K GEMM (N x G)(G x N) with the w*v scaling fused into the A-operand load.
#include <math.h>

#include "grid_fused.cuh"

namespace sqcc {

static int ensure_gridwork(sqcc_ctx *ctx, size_t bytes)
{
    if (ctx->gridwork_bytes >= bytes) return SQCC_OK;
    if (ctx->d_gridwork) cudaFree(ctx->d_gridwork);
    ctx->d_gridwork = nullptr;
    ctx->gridwork_bytes = 0;
    SQCC_CUDA(cudaMalloc(&ctx->d_gridwork, bytes));
    ctx->gridwork_bytes = bytes;
    return SQCC_OK;
}

// wv[s][n] = w_n * c * rho_s[n]^(1/3);  exc += c_e * sum_n w_n rho_s[n]^(4/3)
// pow(x, 1/3) (not cbrt) mirrors numpy's `**(1.0/3.0)`: NaN for negative densities, like the reference.
__global__ void __launch_bounds__(256) lda_pointwise_kernel(const double *__restrict__ rho, const double *__restrict__ w,
                                                            int64_t G, int nspin, double c_pot, double c_en,
                                                            double *__restrict__ wv, double *__restrict__ exc)
{
    double part = 0.0;
    const int64_t total = G * nspin;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = idx % G;
        const double r = rho[idx], wn = w[n];
        wv[idx] = wn * (c_pot * pow(r, 1.0 / 3.0));
        part += wn * pow(r, 4.0 / 3.0);
    }
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
        atomicAdd(exc, c_en * s);
    }
}

__global__ void mul_kernel(const double *a, const double *b, int64_t n, double *o)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = a[i] * b[i];
}

// U_mn = D_mn + D_nm (m < n), U_mm = D_mm, 0 below the diagonal: rho_p = sum_{m <= n} phi_pm U_mn phi_pn for any D
__global__ void upper_fold_kernel(const double *__restrict__ D, int nspin, int n, double *__restrict__ U)
{
    const int64_t tot = (int64_t)nspin * n * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = e / ((int64_t)n * n);
        const int r = (int)((e / n) % n), c = (int)(e % n);
        U[e] = r < c ? D[e] + D[s * n * n + (int64_t)c * n + r] : (r == c ? D[e] : 0.0);
    }
}

static int rho_gemm(sqcc_ctx *ctx, const double *d_D, int nspin, double *d_rho)
{
    const int n = ctx->grid_n;
    SQCC_CUDA(cudaMemsetAsync(d_rho, 0, sizeof(double) * nspin * ctx->G, ctx->stream));
    // triangular form: a 64-column block of T = phi U only needs the k-tiles with k < its last column
    double *U = ctx->d_gridwork + 4 * ctx->G;
    upper_fold_kernel<<<(unsigned)(((int64_t)nspin * n * n + 255) / 256), 256, 0, ctx->stream>>>(d_D, nspin, n, U);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    GemmArgs g{};
    g.ktri = 1;
    g.A = ctx->d_phi;
    g.B = U;
    g.C = d_rho;
    g.X = ctx->d_phi;
    g.scale = nullptr;
    g.M = (int)ctx->G;
    g.N = n;
    g.K = n;
    g.sa_m = n;
    g.sa_k = 1;
    g.sb_k = n;
    g.sc_m = n;
    g.batch = nspin;
    g.batch_a = 0;
    g.batch_b = (int64_t)n * n;
    g.batch_c = ctx->G;
    g.batch_scale = 0;
    g.splitk = 1;
    return gemm_f64(ctx, g, EPI_ROWDOT);
}

static int vxc_gemm(sqcc_ctx *ctx, const double *d_wv, int nspin, double *d_V)
{
    const int n = ctx->grid_n;
    SQCC_CUDA(cudaMemsetAsync(d_V, 0, sizeof(double) * nspin * n * n, ctx->stream));
    GemmArgs g{};
    g.A = ctx->d_phi;   // A(m = p, k = grid point) = phi[k*n + m]
    g.B = ctx->d_phi;   // B(k, q) = phi[k*n + q]
    g.C = d_V;
    g.X = nullptr;
    g.scale = d_wv;
    g.M = n;
    g.N = n;
    g.K = (int)ctx->G;
    g.sa_m = 1;
    g.sa_k = n;
    g.sb_k = n;
    g.sc_m = n;
    g.batch = nspin;
    g.batch_a = 0;
    g.batch_b = 0;
    g.batch_c = (int64_t)n * n;
    g.batch_scale = ctx->G;
    g.symmetric = 1;   // phi^T diag(w v) phi: tiles on and above the diagonal, mirrored in the epilogue
    // enough k-splits to fill the machine ~2x with the computed 64x64 output blocks
    const int nt = (n + 63) / 64;
    int blocks = nt * (nt + 1) / 2 * nspin;
    int splitk = (2 * ctx->sm_count * 4 + blocks - 1) / blocks;
    int ktiles = (int)((ctx->G + GEMM_BK - 1) / GEMM_BK);
    if (splitk > ktiles) splitk = ktiles;
    if (splitk < 1) splitk = 1;
    if ((int64_t)splitk * nspin > 65535) splitk = 65535 / nspin;
    g.splitk = splitk;
    return gemm_f64(ctx, g, EPI_ATOMIC);
}

int grid_rho(sqcc_ctx *ctx, const double *d_D, double *d_rho)
{
    SQCC_REQUIRE(ctx->G < (int64_t)1 << 31, "grid too large");
    SQCC_TRY(ensure_gridwork(ctx, sizeof(double) * (4 * ctx->G + 2 * (size_t)ctx->grid_n * ctx->grid_n)));
    return rho_gemm(ctx, d_D, 1, d_rho);
}

int grid_quadrature(sqcc_ctx *ctx, const double *d_v, double *d_V)
{
    // caller supplies v; the weights are applied here: wv = w * v
    SQCC_TRY(ensure_gridwork(ctx, sizeof(double) * 4 * ctx->G));
    double *wv = ctx->d_gridwork;
    mul_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->d_w, d_v, ctx->G, wv);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    if (grid_fused_ok(ctx)) {
        SQCC_CUDA(cudaMemsetAsync(d_V, 0, sizeof(double) * ctx->grid_n * ctx->grid_n, ctx->stream));
        return launch_vxc<1>(ctx, wv, d_V);
    }
    return vxc_gemm(ctx, wv, 1, d_V);
}

int lda_vxc(sqcc_ctx *ctx, const double *d_D, int nspin, double *d_Vxc, double *d_Exc, double *d_rho)
{
    SQCC_REQUIRE(ctx->G < (int64_t)1 << 31, "grid too large");
    const int64_t G = ctx->G;
    SQCC_TRY(ensure_gridwork(ctx, sizeof(double) * (4 * G + 2 * (size_t)ctx->grid_n * ctx->grid_n)));
    double *rho = d_rho ? d_rho : ctx->d_gridwork;  // [nspin][G]
    double *wv = ctx->d_gridwork + 2 * G;           // [nspin][G]
    // constants exactly as ksdft_functional.py:59, :89, :125, :173 evaluate them
    const double cx = pow(3.0 / M_PI, 1.0 / 3.0);
    double c_pot = -cx, c_en = (-3.0 / 4.0) * cx;
    if (nspin == 2) {
        c_pot = -pow(2.0, 1.0 / 3.0) * cx;
        c_en = (-3.0 / 4.0) * cx * pow(2.0, 1.0 / 3.0);
    }
    SQCC_CUDA(cudaMemsetAsync(d_Exc, 0, sizeof(double), ctx->stream));
    if (grid_fused_ok(ctx)) {   // N <= 64: persistent fused DMMA kernels (grid_fused.cuh)
        const int n = ctx->grid_n;
        SQCC_TRY(timer_start(ctx, 2));
        if (nspin == 1) SQCC_TRY(launch_rho_lda<1>(ctx, d_D, c_pot, c_en, rho, wv, d_Exc));
        else SQCC_TRY(launch_rho_lda<2>(ctx, d_D, c_pot, c_en, rho, wv, d_Exc));
        SQCC_TRY(timer_stop(ctx, 2));
        SQCC_TRY(timer_start(ctx, 3));
        SQCC_CUDA(cudaMemsetAsync(d_Vxc, 0, sizeof(double) * nspin * n * n, ctx->stream));
        if (nspin == 1) SQCC_TRY(launch_vxc<1>(ctx, wv, d_Vxc));
        else SQCC_TRY(launch_vxc<2>(ctx, wv, d_Vxc));
        SQCC_TRY(timer_stop(ctx, 3));
        return SQCC_OK;
    }
    SQCC_TRY(timer_start(ctx, 2));
    SQCC_TRY(rho_gemm(ctx, d_D, nspin, rho));
    lda_pointwise_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(rho, ctx->d_w, G, nspin, c_pot, c_en, wv, d_Exc);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    SQCC_TRY(timer_stop(ctx, 2));
    SQCC_TRY(timer_start(ctx, 3));
    SQCC_TRY(vxc_gemm(ctx, wv, nspin, d_Vxc));
    SQCC_TRY(timer_stop(ctx, 3));
    return SQCC_OK;
}

}  // namespace sqcc
