// Fused LDA grid kernels for N <= 64 AOs (N even): persistent CTAs, cp.async double-buffered phi tiles, FP64 DMMA.
//
//   grid_rho_lda_kernel : rho_n = phi_n D phi_n^T per spin, v_x / E_x point functions fused in the epilogue
//                         (hf_ksdft.py:426-440, :634-646; ksdft_functional.py:41-177).  The (128 x N)(N x N) product
//                         of a tile stays in registers and is dotted with the SAME shared-memory phi tile.
//   grid_vxc_kernel     : V_pq = sum_n phi_np (w_n v_n) phi_nq (hf_ksdft.py:456-476).  A (= phi^T scaled) and B (= phi)
//                         fragments come from one shared tile; each CTA keeps its 64 x 64 partial in registers over
//                         its whole grid-point range and flushes it once with red.global.add.f64.
// Shared tiles are [rows][68] doubles: stride 68 == 4 (mod 16) makes every m8n8k4 fragment load conflict free.
// Larger / odd N use the generic tiled GEMM path (gemm_f64.cuh).
#pragma once
#include "gemm_f64.cuh"

namespace sqcc {

constexpr int GF_TM = 128;      // grid points per tile
constexpr int GF_LD = 68;       // shared row stride (doubles)
constexpr int GF_WARPS = 8;
constexpr int GF_RHO_WARPS = 16;   // rho kernel: 128-point tiles, 8 rows per warp

__device__ __forceinline__ void gf_cp16(void *dst, const void *src, bool on)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(on ? 16u : 0u) : "memory");
}

// Copy tile `t` (rows t*128 .. +128 of phi[G][n], n even) into At[128][68]; rows past G are zero-filled.
template <int TM = GF_TM>
__device__ __forceinline__ void gf_load_tile(double *At, const double *__restrict__ phi, int64_t G, int n, int64_t t)
{
    const int half = n >> 1;
    const int chunks = TM * half;
    const int64_t g0 = t * TM;
    for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
        const int r = c / half, cc = c - r * half;
        const bool on = g0 + r < G;
        gf_cp16(At + r * GF_LD + 2 * cc, phi + (on ? (g0 + r) * n + 2 * cc : 0), on);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int NSPIN>
__global__ void __launch_bounds__(GF_RHO_WARPS * 32, 1)
grid_rho_lda_kernel(const double *__restrict__ phi, const double *__restrict__ w, const double *__restrict__ D, int64_t G,
                    int n, double c_pot, double c_en, double *__restrict__ rho_out, double *__restrict__ wv,
                    double *__restrict__ exc)
{
    extern __shared__ __align__(128) double gsm[];
    double *Ds = gsm;                               // [NSPIN][64][68]
    double *At = Ds + NSPIN * 64 * GF_LD;           // [2][128][68]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // zero everything once (pad columns / rows must stay finite), then stage D
    for (int e = tid; e < NSPIN * 64 * GF_LD + 2 * GF_TM * GF_LD; e += GF_RHO_WARPS * 32) gsm[e] = 0.0;
    __syncthreads();
    // rho_p = sum_mn phi_pm D_mn phi_pn = sum_{m <= n} phi_pm U_mn phi_pn with the upper-triangular U_mn = D_mn + D_nm
    // (m < n), U_mm = D_mm: the product T = phi U needs only the k-steps with k <= n of every column tile -- 72 of the 128
    // (k-step, column tile) pairs at N = 64 -- and is exact for any D, symmetric or not.
    for (int e = tid; e < NSPIN * n * n; e += GF_RHO_WARPS * 32) {
        const int s = e / (n * n), r = (e / n) % n, c = e % n;
        Ds[(s * 64 + r) * GF_LD + c] = r < c ? D[e] + D[(size_t)s * n * n + (size_t)c * n + r] : (r == c ? D[e] : 0.0);
    }
    const int64_t ntiles = (G + GF_TM - 1) / GF_TM;
    const int ksteps = (n + 3) >> 2;
    double exc_part = 0.0;
    int buf = 0;
    if ((int64_t)blockIdx.x < ntiles) gf_load_tile(At, phi, G, n, blockIdx.x);
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t tn = t + gridDim.x;
        __syncthreads();   // everyone is done with the buffer that is refilled next (and D is staged)
        if (tn < ntiles) {
            gf_load_tile(At + (buf ^ 1) * GF_TM * GF_LD, phi, G, n, tn);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double *A = At + buf * GF_TM * GF_LD + (warp * 8) * GF_LD;   // this warp's 8 rows (16 warps: four per scheduler
                                                                            // hide the fragment loads and the epilogue)
#pragma unroll
        for (int s = 0; s < NSPIN; ++s) {
            double acc[8][2];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
            const double *B = Ds + s * 64 * GF_LD;
#pragma unroll
            for (int jm = 0; jm < 8; ++jm) {       // k-steps 2 jm, 2 jm + 1 (rows 8 jm .. 8 jm + 7 of U) meet column tiles j >= jm
                if (2 * jm >= ksteps) break;       // (warp-uniform)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ks = 2 * jm + h;
                    if (ks >= ksteps) break;
                    const int kq = 4 * ks + (lane & 3);
                    const double a0 = A[(lane >> 2) * GF_LD + kq];
#pragma unroll
                    for (int j = jm; j < 8; ++j) {
                        if (8 * j >= n) break;     // column tiles past the basis
                        const double b = B[kq * GF_LD + 8 * j + (lane >> 2)];
                        dmma_m8n8k4(acc[j][0], acc[j][1], a0, b);
                    }
                }
            }
            // rho = rowdot(T, phi): lane holds T[row = lane/4][col = 8j + 2(lane%4) + {0,1}]
            double rsum;
            {
                const int row = lane >> 2;
                double part = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const double2 x = *reinterpret_cast<const double2 *>(A + row * GF_LD + 8 * j + 2 * (lane & 3));
                    part = fma(acc[j][0], x.x, fma(acc[j][1], x.y, part));
                }
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);
                rsum = part;
            }
            // point functions: every lane of a quad holds the row sum; lane%4 == 0 finishes row lane/4.
            // rho^(4/3) = rho * rho^(1/3): same NaN / zero behaviour as NumPy's pow, ~1 ulp from pow(rho, 4/3).
            {
                const int64_t g = t * GF_TM + warp * 8 + (lane >> 2);
                if ((lane & 3) == 0 && g < G) {
                    const double r = rsum;
                    const double wg = w[g];
                    const double cb = pow(r, 1.0 / 3.0);   // pow like NumPy: NaN for rho < 0
                    if (rho_out) rho_out[s * G + g] = r;
                    wv[s * G + g] = wg * (c_pot * cb);
                    exc_part += wg * (r * cb);
                }
            }
        }
        buf ^= 1;
    }
    // E_x: block reduction, one atomic per CTA
    __shared__ double red[GF_RHO_WARPS];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) exc_part += __shfl_xor_sync(0xffffffffu, exc_part, o);
    if (lane == 0) red[warp] = exc_part;
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int i = 0; i < GF_RHO_WARPS; ++i) sum += red[i];
        atomicAdd(exc, c_en * sum);
    }
}

template <int NSPIN, int TMV>
__global__ void __launch_bounds__(GF_WARPS * 32)
grid_vxc_kernel(const double *__restrict__ phi, const double *__restrict__ wv, int64_t G, int n, double *__restrict__ V)
{
    extern __shared__ __align__(128) double gsm[];
    double *At = gsm;                               // [2][128][68]
    double *Ws = At + 2 * TMV * GF_LD;            // [2][NSPIN][128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < 2 * TMV * GF_LD + 2 * NSPIN * TMV; e += GF_WARPS * 32) gsm[e] = 0.0;
    const int64_t ntiles = (G + TMV - 1) / TMV;
    // V is symmetric: only the 36 m8n8 tiles (pt <= qt) of the 64 x 64 output are computed (the flush mirrors them).  Rows
    // r and 7 - r hold 9 of them; warp 2r takes (r, r .. r+4), warp 2r + 1 the other 3 - r tiles of row r and the r + 1
    // tiles of row 7 - r: five / four DMMAs per k-step and warp instead of eight, one A fragment per row, and the two
    // warps of a scheduler (w, w + 4) carry 5 + 4 tiles.
    const int wr = warp >> 1, second = warp & 1;
    const int prow0 = wr, prow1 = second ? 7 - wr : wr;   // p-tile of the first `nfirst` tiles / of the rest
    const int nfirst = second ? 3 - wr : 5;
    const int ntile = second ? 4 : 5;
    int qt[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) qt[i] = !second ? wr + i : (i < nfirst ? wr + 5 + i : (7 - wr) + (i - nfirst));
    if (second) qt[4] = qt[3];   // (unused slot: any valid column)
    double acc[NSPIN][5][2];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s)
#pragma unroll
        for (int i = 0; i < 5; ++i) acc[s][i][0] = acc[s][i][1] = 0.0;
    auto load_w = [&](int b, int64_t t) {
        for (int e = tid; e < NSPIN * TMV; e += GF_WARPS * 32) {
            const int s = e / TMV, r = e % TMV;
            const int64_t g = t * TMV + r;
            Ws[(b * NSPIN + s) * TMV + r] = g < G ? wv[s * G + g] : 0.0;
        }
    };
    int buf = 0;
    __syncthreads();
    if ((int64_t)blockIdx.x < ntiles) {
        gf_load_tile<TMV>(At, phi, G, n, blockIdx.x);
        load_w(0, blockIdx.x);
    }
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t tn = t + gridDim.x;
        __syncthreads();
        if (tn < ntiles) {
            gf_load_tile<TMV>(At + (buf ^ 1) * TMV * GF_LD, phi, G, n, tn);
            load_w(buf ^ 1, tn);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double *A = At + buf * TMV * GF_LD;
        const double *Wt = Ws + buf * NSPIN * TMV;
#pragma unroll 2
        for (int ks = 0; ks < TMV / 4; ++ks) {
            const int gq = 4 * ks + (lane & 3);           // grid point (k index) of this lane's fragment element
            const double *row = A + gq * GF_LD + (lane >> 2);
            double fb[5];
            const double fa0 = row[8 * prow0], fa1 = row[8 * prow1];
#pragma unroll
            for (int i = 0; i < 5; ++i) fb[i] = row[8 * qt[i]];
#pragma unroll
            for (int s = 0; s < NSPIN; ++s) {
                const double sc = Wt[s * TMV + gq];
                const double a0 = fa0 * sc, a1 = fa1 * sc;
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma_m8n8k4(acc[s][i][0], acc[s][i][1], i < nfirst ? a0 : a1, fb[i]);
                if (!second) dmma_m8n8k4(acc[s][4][0], acc[s][4][1], a0, fb[4]);   // (warp-uniform branch)
            }
        }
        buf ^= 1;
    }
    // flush the CTA's partial: lane holds V[p = 8 pt + lane/4][q = 8 qt + 2(lane%4) + {0,1}]; tiles above the diagonal
    // go to both triangles
#pragma unroll
    for (int s = 0; s < NSPIN; ++s)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (i >= ntile) continue;
            const int pt = i < nfirst ? prow0 : prow1;
            const int pr = 8 * pt + (lane >> 2), q = 8 * qt[i] + 2 * (lane & 3);
            double *Vs = V + (size_t)s * n * n;
            if (pr < n) {
                if (q < n) atomicAdd(Vs + (size_t)pr * n + q, acc[s][i][0]);
                if (q + 1 < n) atomicAdd(Vs + (size_t)pr * n + q + 1, acc[s][i][1]);
                if (pt != qt[i]) {
                    if (q < n) atomicAdd(Vs + (size_t)q * n + pr, acc[s][i][0]);
                    if (q + 1 < n) atomicAdd(Vs + (size_t)(q + 1) * n + pr, acc[s][i][1]);
                }
            }
        }
}

inline bool grid_fused_ok(const sqcc_ctx *ctx)
{
    return ctx->grid_n <= 64 && (ctx->grid_n & 1) == 0 && (reinterpret_cast<uintptr_t>(ctx->d_phi) & 15) == 0;
}

template <int NSPIN>
static int launch_rho_lda(sqcc_ctx *ctx, const double *d_D, double c_pot, double c_en, double *d_rho, double *d_wv,
                          double *d_exc)
{
    const size_t smem = sizeof(double) * (NSPIN * 64 * GF_LD + 2 * GF_TM * GF_LD);
    auto kern = grid_rho_lda_kernel<NSPIN>;
    SQCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (ctx->G + GF_TM - 1) / GF_TM;
    const int grid = (int)(ntiles < ctx->sm_count ? ntiles : ctx->sm_count);
    kern<<<grid, GF_RHO_WARPS * 32, smem, ctx->stream>>>(ctx->d_phi, ctx->d_w, d_D, ctx->G, ctx->grid_n, c_pot, c_en, d_rho,
                                                     d_wv, d_exc);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

template <int NSPIN>
static int launch_vxc(sqcc_ctx *ctx, const double *d_wv, double *d_V)
{
    constexpr int TMV = 64;   // 64-point tiles: 70 KB per CTA, three CTAs per SM overlap each other's tile loads
    const size_t smem = sizeof(double) * (2 * TMV * GF_LD + 2 * NSPIN * TMV);
    auto kern = grid_vxc_kernel<NSPIN, TMV>;
    SQCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SQCC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GF_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    static const int vxc_ctas = getenv("SQCC_VXC_CTAS") ? atoi(getenv("SQCC_VXC_CTAS")) : 3;
    if (per_sm > vxc_ctas) per_sm = vxc_ctas;
    const int64_t ntiles = (ctx->G + TMV - 1) / TMV;
    const int grid = (int)(ntiles < (int64_t)ctx->sm_count * per_sm ? ntiles : (int64_t)ctx->sm_count * per_sm);
    kern<<<grid, GF_WARPS * 32, smem, ctx->stream>>>(ctx->d_phi, d_wv, ctx->G, ctx->grid_n, d_V);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

}  // namespace sqcc
