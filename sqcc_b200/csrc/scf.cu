// Device-resident SCF iteration: everything between two Fock builds stays in HBM.
//
// The reference does, per iteration, on the host (hf_ksdft.py): Fock assembly :498-504 / :527-548, electronic energy
// :552-587, X^T F X + eigh :97-123 / :604-611, C = X C' :616-621 and the density :150-156 / :178-199 / :625.  With the
// J/K build at ~1-10 ms those O(N^3) host steps plus the D / J / K transfers dominate the SCF wall time for N >= 200.
// Here the density, J/K (and V_xc), Fock matrices, orbitals and the new density never leave the device; one double
// (the electronic energy, needed for the reference's |dE| < 1e-9 test) crosses PCIe per iteration.
//
//   sqcc_scf_fock_energy : J/K (+ LDA V_xc) for the resident density -> F (per spin) and E_elec, same formulas as the
//                          reference; with a peer group attached J/K are the full matrices (fused exchange, comm.cu)
//   sqcc_scf_update      : F' = X^T F X (two DMMA GEMMs), eigh(F') with cuSOLVER Dsyevd (library call, not a hot loop),
//                          C^T = V X (GEMM), D = occ * C_occ C_occ^T (GEMM + scale)
// cuSOLVER is resolved with dlopen at first use so that libsqcc_b200.so itself has no link-time dependency on it.
#include <cusolverDn.h>
#include <dlfcn.h>

#include "gemm_f64.cuh"

namespace sqcc {

struct CusolverApi {
    void *lib = nullptr;
    cusolverStatus_t (*create)(cusolverDnHandle_t *) = nullptr;
    cusolverStatus_t (*destroy)(cusolverDnHandle_t) = nullptr;
    cusolverStatus_t (*set_stream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    cusolverStatus_t (*syevd_bufsize)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, const double *, int,
                                      const double *, int *) = nullptr;
    cusolverStatus_t (*syevd)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, double *, int, double *, double *,
                              int, int *) = nullptr;
};

static CusolverApi g_cusolver;

static int load_cusolver()
{
    if (g_cusolver.lib) return SQCC_OK;
    const char *names[] = {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"};
    void *h = nullptr;
    for (const char *nm : names)
        if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) {
        set_error("device SCF needs cuSOLVER (libcusolver.so not found)");
        return SQCC_ERR_UNSUPPORTED;
    }
    CusolverApi a;
    a.lib = h;
    a.create = reinterpret_cast<decltype(a.create)>(dlsym(h, "cusolverDnCreate"));
    a.destroy = reinterpret_cast<decltype(a.destroy)>(dlsym(h, "cusolverDnDestroy"));
    a.set_stream = reinterpret_cast<decltype(a.set_stream)>(dlsym(h, "cusolverDnSetStream"));
    a.syevd_bufsize = reinterpret_cast<decltype(a.syevd_bufsize)>(dlsym(h, "cusolverDnDsyevd_bufferSize"));
    a.syevd = reinterpret_cast<decltype(a.syevd)>(dlsym(h, "cusolverDnDsyevd"));
    if (!a.create || !a.destroy || !a.set_stream || !a.syevd_bufsize || !a.syevd) {
        set_error("cuSOLVER symbols missing");
        return SQCC_ERR_UNSUPPORTED;
    }
    g_cusolver = a;
    return SQCC_OK;
}

struct ScfState {
    int n = 0, nspin = 1, nocc[2] = {0, 0}, ksdft = 0;
    double *buf = nullptr;   // one allocation, carved below
    double *H, *X, *D, *JK, *VXC, *F, *T, *FP, *CT, *W, *E, *EP;   // E: [0] electronic energy, [1] E_xc; EP: block partials
    double *work = nullptr;
    int lwork = 0;
    int *info = nullptr;
    cusolverDnHandle_t solver = nullptr;
};

void scf_free(sqcc_ctx *ctx)
{
    ScfState *s = ctx->scf;
    if (!s) return;
    if (s->solver && g_cusolver.destroy) g_cusolver.destroy(s->solver);
    cudaFree(s->buf);
    cudaFree(s->work);
    cudaFree(s->info);
    delete s;
    ctx->scf = nullptr;
}

// F (per spin) and the electronic energy, formulas of hf_ksdft.py:498-504, :527-548, :552-587.
//   restricted HF : F = H + J - K/2            E = 1/2 sum D (H + F)
//   restricted KS : F = H + J + Vxc            E = 1/2 sum D (2H + J) + Exc
//   unrestricted HF : F_s = H + J - K_s        E = 1/2 sum (Da+Db) H + 1/2 sum Da Fa + 1/2 sum Db Fb
//   unrestricted KS : F_s = H + J + Vxc_s      E = 1/2 sum (Da+Db) (2H + J) + Exc
__global__ void __launch_bounds__(256) scf_fock_energy_kernel(const double *__restrict__ H, const double *__restrict__ D,
                                                              const double *__restrict__ J, const double *__restrict__ K,
                                                              const double *__restrict__ V, int nspin, int ksdft, size_t nn,
                                                              double *__restrict__ F, double *__restrict__ E)
{
    double part = 0.0;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nn; idx += (size_t)gridDim.x * blockDim.x) {
        const double h = H[idx], j = J[idx];
        if (nspin == 1) {
            const double d = D[idx];
            if (ksdft) {
                F[idx] = h + j + V[idx];
                part += 0.5 * d * (2.0 * h + j);
            } else {
                const double f = h + j - 0.5 * K[idx];
                F[idx] = f;
                part += 0.5 * d * (h + f);
            }
        } else {
            const double da = D[idx], db = D[nn + idx];
            if (ksdft) {
                F[idx] = h + j + V[idx];
                F[nn + idx] = h + j + V[nn + idx];
                part += 0.5 * (da + db) * (2.0 * h + j);
            } else {
                const double fa = h + j - K[idx], fb = h + j - K[nn + idx];
                F[idx] = fa;
                F[nn + idx] = fb;
                part += 0.5 * (da + db) * h + 0.5 * da * fa + 0.5 * db * fb;
            }
        }
    }
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        E[blockIdx.x] = s;   // block partial; summed in a fixed order below
    }
}

// Fixed-order sum of the block partials: every rank of a multi-GPU run must get the bit-identical energy (the ranks take
// the |dE| < 1e-9 decision independently and have to agree on it).
__global__ void scf_energy_sum_kernel(const double *__restrict__ part, int nparts, double *__restrict__ E)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < nparts; ++b) s += part[b];
        E[0] = s;
    }
}

__global__ void scf_scale_kernel(double *x, size_t n, double a)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= a;
}

static int nn_gemm(sqcc_ctx *ctx, const double *A, bool a_transposed, const double *B, double *C, int m, int n, int k, int ld)
{
    // C[m,n] = op(A)[m,k] B[k,n]; all matrices row-major with leading dimension ld
    GemmArgs g{};
    g.A = A;
    g.B = B;
    g.C = C;
    g.M = m;
    g.N = n;
    g.K = k;
    g.sa_m = a_transposed ? 1 : ld;
    g.sa_k = a_transposed ? ld : 1;
    g.sb_k = ld;
    g.sc_m = ld;
    g.batch = 1;
    g.splitk = 1;
    return gemm_f64(ctx, g, EPI_STORE);
}

}  // namespace sqcc

using namespace sqcc;

extern "C" {

int sqcc_scf_setup(sqcc_ctx *ctx, const double *h_H, const double *h_X, int nspin, int n_occ_a, int n_occ_b, int ksdft)
{
    SQCC_REQUIRE(ctx && ctx->d_eri && h_H && h_X, "no ERI attached / null matrices");
    SQCC_REQUIRE(nspin == 1 || nspin == 2, "nspin must be 1 (restricted) or 2 (unrestricted)");
    const int n = ctx->n;
    SQCC_REQUIRE(n_occ_a >= 0 && n_occ_a <= n && n_occ_b >= 0 && n_occ_b <= n, "occupation out of range");
    SQCC_REQUIRE(!ksdft || (ctx->d_phi && ctx->grid_n == n), "KS-DFT needs the grid attached (sqcc_grid_attach)");
    SQCC_TRY(load_cusolver());
    SQCC_CUDA(cudaSetDevice(ctx->device));
    scf_free(ctx);
    ScfState *s = new ScfState();
    ctx->scf = s;
    s->n = n;
    s->nspin = nspin;
    s->nocc[0] = n_occ_a;
    s->nocc[1] = n_occ_b;
    s->ksdft = ksdft;
    const size_t nn = (size_t)n * n;
    // H, X, D[2], JK[3], VXC[2], F[2], T, FP, CT[2], W[2 n], E[2]
    const size_t total = nn * (1 + 1 + 2 + 3 + 2 + 2 + 1 + 1 + 2) + 2 * (size_t)n + 2 + 1024;
    SQCC_CUDA(cudaMalloc(&s->buf, total * sizeof(double)));
    SQCC_CUDA(cudaMemsetAsync(s->buf, 0, total * sizeof(double), ctx->stream));
    double *p = s->buf;
    s->H = p; p += nn;
    s->X = p; p += nn;
    s->D = p; p += 2 * nn;
    s->JK = p; p += 3 * nn;
    s->VXC = p; p += 2 * nn;
    s->F = p; p += 2 * nn;
    s->T = p; p += nn;
    s->FP = p; p += nn;
    s->CT = p; p += 2 * nn;
    s->W = p; p += 2 * (size_t)n;
    s->E = p; p += 2;
    s->EP = p;
    SQCC_CUDA(cudaMemcpyAsync(s->H, h_H, nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SQCC_CUDA(cudaMemcpyAsync(s->X, h_X, nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SQCC_CUDA(cudaMalloc(&s->info, sizeof(int)));
    SQCC_CUDA(cudaMemset(s->info, 0, sizeof(int)));
    if (g_cusolver.create(&s->solver) != CUSOLVER_STATUS_SUCCESS) {
        set_error("cusolverDnCreate failed");
        return SQCC_ERR_CUDA;
    }
    g_cusolver.set_stream(s->solver, ctx->stream);
    if (g_cusolver.syevd_bufsize(s->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, s->FP, n, s->W, &s->lwork) !=
        CUSOLVER_STATUS_SUCCESS) {
        set_error("cusolverDnDsyevd_bufferSize failed");
        return SQCC_ERR_CUDA;
    }
    SQCC_CUDA(cudaMalloc(&s->work, sizeof(double) * (size_t)(s->lwork > 0 ? s->lwork : 1)));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    return SQCC_OK;
}

int sqcc_scf_set_density(sqcc_ctx *ctx, const double *h_D)
{
    SQCC_REQUIRE(ctx && ctx->scf && h_D, "sqcc_scf_setup first");
    ScfState *s = ctx->scf;
    const size_t nn = (size_t)s->n * s->n;
    SQCC_CUDA(cudaMemcpyAsync(s->D, h_D, s->nspin * nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    return SQCC_OK;
}

// The resident density's J/K (and LDA V_xc) into the state's [J; K...] buffer.  On a sharded tensor whose partials
// are NOT summed by the fused exchange (no peer group, or sqcc_comm_set_enabled(0)) the buffer holds this rank's
// PARTIAL matrices: the caller must all-reduce *d_jk (n_doubles doubles) before sqcc_scf_finish.
int sqcc_scf_jk(sqcc_ctx *ctx, void **d_jk, int64_t *n_doubles)
{
    SQCC_REQUIRE(ctx && ctx->scf, "sqcc_scf_setup first");
    ScfState *s = ctx->scf;
    const size_t nn = (size_t)s->n * s->n;
    SQCC_CUDA(cudaMemsetAsync(s->E, 0, 2 * sizeof(double), ctx->stream));
    if (s->ksdft) SQCC_TRY(lda_vxc(ctx, s->D, s->nspin, s->VXC, s->E + 1, nullptr));
    SQCC_TRY(jk_build(ctx, s->D, s->nspin, s->ksdft ? 0 : 1, s->JK, s->ksdft ? nullptr : s->JK + nn));
    if (d_jk) *d_jk = s->JK;
    if (n_doubles) *n_doubles = (int64_t)((s->ksdft ? 1 : 1 + s->nspin) * nn);
    return SQCC_OK;
}

// Fock matrices + electronic energy from the (complete) J/K in the state's buffer.
int sqcc_scf_finish(sqcc_ctx *ctx, double *h_e_elec)
{
    SQCC_REQUIRE(ctx && ctx->scf && h_e_elec, "sqcc_scf_setup first");
    ScfState *s = ctx->scf;
    const int n = s->n;
    const size_t nn = (size_t)n * n;
    const int nparts = ctx->sm_count < 1024 ? ctx->sm_count : 1024;
    scf_fock_energy_kernel<<<nparts, 256, 0, ctx->stream>>>(s->H, s->D, s->JK, s->JK + nn, s->VXC, s->nspin, s->ksdft, nn,
                                                           s->F, s->EP);
    scf_energy_sum_kernel<<<1, 32, 0, ctx->stream>>>(s->EP, nparts, s->E);
    ctx->launches += 2;
    SQCC_CUDA(cudaGetLastError());
    double e[2];
    SQCC_CUDA(cudaMemcpyAsync(e, s->E, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    int info = 0;   // the eigensolver status of the previous update is visible now
    SQCC_CUDA(cudaMemcpy(&info, s->info, sizeof(int), cudaMemcpyDeviceToHost));
    SQCC_REQUIRE(info == 0, "cusolverDnDsyevd did not converge");
    *h_e_elec = e[0] + (s->ksdft ? e[1] : 0.0);
    if (ctx->comm.attached && ctx->comm.h_status && *ctx->comm.h_status != 0) {
        set_error("peer barrier timed out in the fused J/K exchange");
        return SQCC_ERR_CUDA;
    }
    return SQCC_OK;
}

int sqcc_scf_fock_energy(sqcc_ctx *ctx, double *h_e_elec)
{
    SQCC_REQUIRE(ctx && ctx->scf && h_e_elec, "sqcc_scf_setup first");
    // a shard's partial J/K must never reach the Fock matrix: without the fused exchange the caller has to sum the
    // partials itself (sqcc_scf_jk -> all-reduce -> sqcc_scf_finish)
    const bool summed = ctx->whole || (ctx->comm.attached && ctx->comm.enabled && ctx->comm.world > 1 && !ctx->comm.same_device);
    SQCC_REQUIRE(summed, "sharded ERI without the fused peer exchange: use sqcc_scf_jk + all-reduce + sqcc_scf_finish");
    SQCC_TRY(sqcc_scf_jk(ctx, nullptr, nullptr));
    return sqcc_scf_finish(ctx, h_e_elec);
}

int sqcc_scf_update(sqcc_ctx *ctx)
{
    SQCC_REQUIRE(ctx && ctx->scf, "sqcc_scf_setup first");
    ScfState *s = ctx->scf;
    const int n = s->n;
    const size_t nn = (size_t)n * n;
    SQCC_TRY(timer_start(ctx, 6));
    for (int sp = 0; sp < s->nspin; ++sp) {
        const double *F = s->F + sp * nn;
        double *CT = s->CT + sp * nn, *D = s->D + sp * nn, *W = s->W + (size_t)sp * n;
        // F' = X^T (F X)
        SQCC_TRY(nn_gemm(ctx, F, false, s->X, s->T, n, n, n, n));
        SQCC_TRY(nn_gemm(ctx, s->X, true, s->T, s->FP, n, n, n, n));
        // eigh: cuSOLVER is column-major, F' is symmetric; on return row k of FP (row-major view) is eigenvector k
        if (g_cusolver.syevd(s->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, s->FP, n, W, s->work, s->lwork,
                             s->info) != CUSOLVER_STATUS_SUCCESS) {
            set_error("cusolverDnDsyevd failed");
            return SQCC_ERR_CUDA;
        }
        // C^T[k, p] = sum_q V[k, q] X[q, p]   (C = X C', hf_ksdft.py:616-621)
        SQCC_TRY(nn_gemm(ctx, s->FP, false, s->X, CT, n, n, n, n));
        // D[p, q] = occ * sum_{k < n_occ} C^T[k, p] C^T[k, q]   (hf_ksdft.py:150-156, :178-199)
        const int nocc = s->nocc[sp];
        if (nocc > 0) {
            SQCC_TRY(nn_gemm(ctx, CT, true, CT, D, n, n, nocc, n));
            if (s->nspin == 1) {
                scf_scale_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(D, nn, 2.0);
                ctx->launches++;
            }
        } else {
            SQCC_CUDA(cudaMemsetAsync(D, 0, nn * sizeof(double), ctx->stream));
        }
    }
    SQCC_CUDA(cudaGetLastError());
    SQCC_TRY(timer_stop(ctx, 6));
    return SQCC_OK;
}

int sqcc_scf_get(sqcc_ctx *ctx, double *h_D, double *h_Ct, double *h_eps)
{
    SQCC_REQUIRE(ctx && ctx->scf, "sqcc_scf_setup first");
    ScfState *s = ctx->scf;
    const size_t nn = (size_t)s->n * s->n;
    if (h_D) SQCC_CUDA(cudaMemcpyAsync(h_D, s->D, s->nspin * nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (h_Ct) SQCC_CUDA(cudaMemcpyAsync(h_Ct, s->CT, s->nspin * nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (h_eps) SQCC_CUDA(cudaMemcpyAsync(h_eps, s->W, s->nspin * (size_t)s->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    return SQCC_OK;
}

}  // extern "C"
