// FP64 tensor-core GEMM kernels and their dispatch (declarations and the DMMA helper live in gemm_f64.cuh).
#include "gemm_f64.cuh"

namespace sqcc {

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

// ---- cp.async-pipelined variant (EPI_STORE, no scale): the AO->MO quarter transforms ---------------------------------
// CTA = 4 warps laid out WM x WN, every warp a (8 MT) x 32 sub-tile (MT x 4 m8n8 accumulators, MT + 4 fragment loads
// per 4 MT DMMAs), CTA tile BM x BN x 16 with BM = 8 MT WM, BN = 32 WN.  Operands go global -> shared with 8-byte
// cp.async (any double alignment; out-of-range elements zero-filled) through a PSTAGES-deep ring, one commit group per
// k-tile, so PSTAGES - 1 tiles are in flight while the tensor pipe works (plus the other CTAs of the SM).  <1,4,MT> (BM = 8 MT <= 32, BN = 128) serves
// the skinny products with M = n_occ; <2,2,4> (64 x 64) the rest.
// (Tried and dropped: several batch entries per CTA with the ring running across the batch boundaries -- the flattened
// loop was 20-45 % slower on the M = 21, K = 222 first quarter than one CTA per product.)

__device__ __forceinline__ void cp_async8_zfill(double *smem_dst, const double *src, bool on)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned nbytes = on ? 8u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
}

template <int WM, int WN, int MT, int PSTAGES, int BK>
constexpr int gemm_pipe_smem_bytes()
{
    return PSTAGES * BK * ((8 * MT * WM + 4) + (32 * WN + 4) + 1) * (int)sizeof(double);
}

// EPI / SCALE (the grid quadrature for N > 64): EPI_ATOMIC = split-K partial sums (blockIdx.z = batch * splitk + split) with
// the per-k scale w v folded into the A fragments; EPI_ROWDOT = out[m] += sum_n acc[m,n] X[m,n] (rho on the grid).
template <bool A_MFAST, int BMODE, int WM, int WN, int MT, int PSTAGES, int BK, int EPI = EPI_STORE, bool SCALE = false>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_f64_pipe_kernel(const GemmArgs g)
{
    static_assert(WM * WN == 4, "four warps");
    constexpr int BM = 8 * MT * WM, BN = 32 * WN;
    constexpr int LDA_S = BM + 4, LDB_S = BN + 4;          // == 4 (mod 8): conflict-free fragment loads
    constexpr int EA = (BM * BK + GEMM_THREADS - 1) / GEMM_THREADS, EB = BN * BK / GEMM_THREADS;
    extern __shared__ __align__(128) double gemm_sm[];
    double *As = gemm_sm;                                      // [PSTAGES][BK][LDA_S]
    double *Bs = gemm_sm + PSTAGES * BK * LDA_S;           // [PSTAGES][BK][LDB_S]
    double *Ss = Bs + PSTAGES * BK * LDB_S;                // [PSTAGES][BK] per-k scale
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp / WN) * (8 * MT), wn = (warp % WN) * 32;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (EPI == EPI_ATOMIC && BM == BN && g.symmetric && blockIdx.x < blockIdx.y) return;   // mirrored by the tile (y, x)
    const bool mirror = EPI == EPI_ATOMIC && BM == BN && g.symmetric && blockIdx.x != blockIdx.y;
    const int bz = EPI == EPI_STORE ? blockIdx.z : blockIdx.z / g.splitk;
    const int sz = EPI == EPI_STORE ? 0 : blockIdx.z % g.splitk;
    const double *A = g.A + (int64_t)bz * g.batch_a;
    const double *B = g.B + (int64_t)bz * g.batch_b;
    const double *scale = SCALE ? g.scale + (int64_t)bz * g.batch_scale : nullptr;
    const int64_t *koff = BMODE == B_KOFF ? g.koff + (int64_t)bz * g.batch_koff : nullptr;
    // k-tiles of this CTA (split-K: contiguous ranges of whole tiles)
    const int Keff = g.ktri && n0 + BN < g.K ? n0 + BN : g.K;
    const int ktiles_all = (Keff + BK - 1) / BK;
    const int per = EPI == EPI_STORE ? ktiles_all : (ktiles_all + g.splitk - 1) / g.splitk;
    const int kt_lo = sz * per;
    const int ktiles = kt_lo + per < ktiles_all ? per : (ktiles_all > kt_lo ? ktiles_all - kt_lo : 0);

    auto issue_tile = [&](int kt, int stage) {
        const int kbase = (kt_lo + kt) * BK;
        double *as = As + stage * BK * LDA_S, *bs = Bs + stage * BK * LDB_S;
        if (SCALE && tid < BK) cp_async8_zfill(Ss + stage * BK + tid, kbase + tid < g.K ? scale + kbase + tid : scale, kbase + tid < g.K);
#pragma unroll
        for (int e = 0; e < EA; ++e) {
            const int idx = tid + e * GEMM_THREADS;
            if (BM * BK % GEMM_THREADS != 0 && idx >= BM * BK) break;
            int m, k;
            if (A_MFAST) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
            const int gm = m0 + m, gk = kbase + k;
            const bool on = gm < g.M && gk < g.K;
            cp_async8_zfill(as + k * LDA_S + m, on ? A + (int64_t)gm * g.sa_m + (int64_t)gk * g.sa_k : A, on);
        }
#pragma unroll
        for (int e = 0; e < EB; ++e) {
            const int idx = tid + e * GEMM_THREADS;
            const int nn = idx % BN, kb = idx / BN;
            const int gn = n0 + nn, gk = kbase + kb;
            const bool on = gn < g.N && gk < g.K;
            const double *src = B;
            if (on) {
                if (BMODE == B_ROWS) src = B + (int64_t)gk * g.sb_k + gn;
                else if (BMODE == B_KOFF) src = B + __ldg(koff + gk) + gn;
                else if (BMODE == B_PAIRQ) {
                    const int64_t hi = gk > bz ? gk : bz, lo = gk > bz ? bz : gk;
                    src = g.B + (hi * (hi + 1) / 2 + lo) * g.sb_k + gn;
                } else {
                    const int64_t hi = gk > gn ? gk : gn, lo = gk > gn ? gn : gk;
                    src = B + hi * (hi + 1) / 2 + lo;
                }
            }
            cp_async8_zfill(bs + kb * LDB_S + nn, src, on);
        }
    };

    double acc[MT][4][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < PSTAGES - 1; ++s) {
        if (s < ktiles) issue_tile(s, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kt = 0; kt < ktiles; ++kt) {
        asm volatile("cp.async.wait_group %0;" ::"n"(PSTAGES - 2) : "memory");   // tile kt has landed (this thread's part)
        __syncthreads();                                                          // ... everyone's; stage (kt-1)%P is free
        if (kt + PSTAGES - 1 < ktiles) issue_tile(kt + PSTAGES - 1, (kt + PSTAGES - 1) % PSTAGES);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double *as = As + (kt % PSTAGES) * BK * LDA_S, *bs = Bs + (kt % PSTAGES) * BK * LDB_S;
        const double *ss = Ss + (kt % PSTAGES) * BK;
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            double af[MT], bf[4];
#pragma unroll
            for (int i = 0; i < MT; ++i) af[i] = as[(ks + (lane & 3)) * LDA_S + wm + i * 8 + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = bs[(ks + (lane & 3)) * LDB_S + wn + j * 8 + (lane >> 2)];
            if (SCALE) {   // the B fragment is the cheaper side to scale (4 values per lane against MT)
                const double sc = ss[ks + (lane & 3)];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] *= sc;
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    double *C = g.C + (int64_t)bz * g.batch_c;
    if (EPI == EPI_ROWDOT) {
        // out[m] += sum_n acc[m,n] X[m,n]: the lane's columns, then the four lanes of a quad, one atomic per row and warp
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int gm = m0 + wm + i * 8 + (lane >> 2);
            double part = 0.0;
            if (gm < g.M) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int gn = n0 + wn + j * 8 + 2 * (lane & 3);
                    const double *x = g.X + (int64_t)gm * g.sc_m + gn;
                    if (gn < g.N) part = fma(acc[i][j][0], __ldg(x), part);
                    if (gn + 1 < g.N) part = fma(acc[i][j][1], __ldg(x + 1), part);
                }
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if ((lane & 3) == 0 && gm < g.M) atomicAdd(C + gm, part);
        }
    } else {
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
                    if (mirror) {
                        if (gn < g.N) atomicAdd(C + (int64_t)gn * g.sc_m + gm, acc[i][j][0]);
                        if (gn + 1 < g.N) atomicAdd(C + (int64_t)(gn + 1) * g.sc_m + gm, acc[i][j][1]);
                    }
                } else {
                    if (gn < g.N) dst[0] = acc[i][j][0];
                    if (gn + 1 < g.N) dst[1] = acc[i][j][1];
                }
            }
        }
    }
}

template <bool MF, int BMD, int WM, int WN, int MT, int PSTAGES, int BK, int EPI = EPI_STORE, bool SCALE = false>
inline int gemm_pipe_launch_st(sqcc_ctx *ctx, const GemmArgs &g)
{
    constexpr int BM = 8 * MT * WM, BN = 32 * WN;
    constexpr int smem = gemm_pipe_smem_bytes<WM, WN, MT, PSTAGES, BK>();
    auto kern = gemm_f64_pipe_kernel<MF, BMD, WM, WN, MT, PSTAGES, BK, EPI, SCALE>;
    static bool attr_done = false;   // one attribute call per instantiation
    if (!attr_done) {
        SQCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = true;
    }
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.batch * (EPI == EPI_STORE ? 1 : g.splitk));
    SQCC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "GEMM grid too large");
    kern<<<grid, GEMM_THREADS, smem, ctx->stream>>>(g);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

template <bool MF, int BMD, int WM, int WN, int MT>
inline int gemm_pipe_launch(sqcc_ctx *ctx, const GemmArgs &g)
{
    // BK = 16 with a 2-deep ring: at 41 KB per CTA five CTAs share an SM, and on the short-K (K = N <= 680) quarter
    // transforms that occupancy beats a deeper ring (N = 222 MP2 transform: 2 stages 5.9 ms, 3: 6.3, 4: 6.9)
    return gemm_pipe_launch_st<MF, BMD, WM, WN, MT, 2, 16>(ctx, g);
}

template <bool MF, int BMD>
inline int gemm_pipe_shape(sqcc_ctx *ctx, const GemmArgs &g)
{
    if (g.M <= 8) return gemm_pipe_launch<MF, BMD, 1, 4, 1>(ctx, g);
    if (g.M <= 16) return gemm_pipe_launch<MF, BMD, 1, 4, 2>(ctx, g);
    if (g.M <= 24) return gemm_pipe_launch<MF, BMD, 1, 4, 3>(ctx, g);
    if (g.M <= 32) return gemm_pipe_launch<MF, BMD, 1, 4, 4>(ctx, g);
    return gemm_pipe_launch<MF, BMD, 2, 2, 4>(ctx, g);
}

// EPI_STORE without a per-k scale and without split-K: the pipelined kernel (all B modes, both A layouts).
inline int gemm_f64_store(sqcc_ctx *ctx, const GemmArgs &g, int bmode)
{
    const bool mfast = g.sa_m == 1 && g.sa_k != 1;
    if (mfast) {
        if (bmode == B_ROWS) return gemm_pipe_shape<true, B_ROWS>(ctx, g);
        if (bmode == B_KOFF) return gemm_pipe_shape<true, B_KOFF>(ctx, g);
        if (bmode == B_PAIRQ) return gemm_pipe_shape<true, B_PAIRQ>(ctx, g);
        return gemm_pipe_shape<true, B_SYMPK>(ctx, g);
    }
    SQCC_REQUIRE(bmode == B_ROWS, "packed B operands need the A^T form");
    return gemm_pipe_shape<false, B_ROWS>(ctx, g);
}

// bmode / small-M tiles are only instantiated for the A-transposed STORE form the AO->MO transforms use.
int gemm_f64(sqcc_ctx *ctx, const GemmArgs &g, int epi, int bmode)
{
    SQCC_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.batch > 0 && g.splitk > 0, "bad GEMM shape");
    SQCC_REQUIRE(g.sa_m == 1 || g.sa_k == 1, "GEMM A operand needs a unit stride");
    SQCC_REQUIRE(epi != EPI_STORE || g.splitk == 1, "split-K needs an accumulating epilogue");
    if (epi == EPI_STORE && !g.scale) return gemm_f64_store(ctx, g, bmode);
    const bool mfast = g.sa_m == 1 && g.sa_k != 1;
    // the grid quadrature (generic N): pipelined kernel, 64 x 64 tiles, three ring stages (K is long or the grid is huge)
    if (bmode == B_ROWS && epi == EPI_ATOMIC && g.scale && mfast)
        return gemm_pipe_launch_st<true, B_ROWS, 2, 2, 4, 3, 16, EPI_ATOMIC, true>(ctx, g);
    if (bmode == B_ROWS && epi == EPI_ROWDOT && !g.scale && !mfast)
        return gemm_pipe_launch_st<false, B_ROWS, 2, 2, 4, 3, 16, EPI_ROWDOT, false>(ctx, g);
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
