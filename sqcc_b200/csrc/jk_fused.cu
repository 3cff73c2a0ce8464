// Fused Coulomb/exchange build over the native-v1 packed ERI shard (one pass, FP64).
//
// Replaces the loop nests hf_ksdft.py:483-487 (J), :491-495 (K), and the UHF/UKS variants :510-515,
// :522-526, :531-535.  For every stored v = (ij|kl) * f  (SURVEY.md A.1, D symmetric, D2 = 2*(Da+Db)):
//     J~[i,j] += v * D2[k,l]          J~[k,l] += v * D2[i,j]
//     K~[i,k] += v * D[j,l]           K~[i,l] += v * D[j,k]
//     K~[j,k] += v * D[i,l]           K~[j,l] += v * D[i,k]           (per spin density)
//     J = J~ + J~^T,  K = K~ + K~^T.
//
// Kernel shape (HBM-bound streaming kernel, no tensor cores):
//   * a WARP is the unit of work.  A warp task is (k-range of JK_KN sub-rows) x (l-chunk of 64 columns,
//     two per lane -> one 128-bit load per lane per row) x (a list of row-blocks of NI x NJ rows (i,j)).
//   * everything an integral touches lives in registers: D[j,l], D[i,l] lane operands, the K~[i,l], K~[j,l]
//     and J~[i,j] accumulators; D[i,k], D[j,k], D2[i,j] are warp-uniform; K~[i,k], K~[j,k] are reduced
//     across lanes with a multi-value shuffle butterfly once per sub-row.
//   * J~[k,l] is accumulated across ALL row-blocks of the task in a per-warp shared-memory slab next to the
//     TMA-free staged D2[k,l] tile (exclusive ownership -> plain read-modify-write, no shared atomics).
//   * flushes to the global J~/K~ work matrices use red.global.add.f64; they amount to ~2% of the
//     streamed elements.
//   * persistent grid (SM count x resident CTAs), tasks pulled from an atomic counter; adjacent task ids
//     are adjacent l-chunks of the same rows, so concurrently running warps stream adjacent addresses.
#include "common.cuh"

namespace sqcc {

struct JKParams {
    const double *P;
    const int64_t *rowoff;
    const int64_t *ioff;   // [n+1] offset of row (i,0) from the start of the full native tensor
    int64_t shard_base;    // offset of this shard's first row in the full native tensor
    int64_t ij_begin, ij_end;
    int n, ldd;
    const double *d2;   // [n][ldd]   2*(Da+Db)
    const double *dsp;  // [nspin][n][ldd]
    double *jt;         // [n][ldd]
    double *kt;         // [nspin][n][ldd]
    const JKTask *tasks;
    const RowBlock *rowblocks;
    int ntasks;
    unsigned int *counter;
};

__device__ __forceinline__ double2 ldg_stream(const double *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ void red_add(double *p, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// Reduce V (power of two, <= 32) per-lane values across the warp.  Afterwards every lane of group
// g = lane / (32 / V) holds the total of vals[g] in vals[0].
template <int V>
__device__ __forceinline__ void warp_reduce_multi(double (&vals)[V], int lane)
{
    int width = 16;
#pragma unroll
    for (int cnt = V; cnt > 1; cnt >>= 1) {
        const int half = cnt >> 1;
        const bool upper = (lane & width) != 0;
#pragma unroll
        for (int t = 0; t < half; ++t) {
            double send = upper ? vals[t] : vals[t + half];
            double keep = upper ? vals[t + half] : vals[t];
            vals[t] = keep + __shfl_xor_sync(0xffffffffu, send, width);
        }
        width >>= 1;
    }
#pragma unroll
    for (; width >= 1; width >>= 1) vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], width);
}

constexpr int next_pow2(int x) { return x <= 1 ? 1 : (x <= 2 ? 2 : (x <= 4 ? 4 : (x <= 8 ? 8 : (x <= 16 ? 16 : 32)))); }

template <int NI, int NJ, int NSPIN, bool WANTK, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) jk_tiled_kernel(const JKParams p)
{
    extern __shared__ double2 smem2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 *s_jt = smem2 + (size_t)warp * (2 * JK_KN * 32);
    double2 *s_d2 = s_jt + JK_KN * 32;
    const int n = p.n, ldd = p.ldd;
    const size_t nl = (size_t)n * ldd;
    constexpr int NS = WANTK ? NSPIN : 1;  // sizes of K-related register arrays (unused when !WANTK)
    constexpr int VCD = next_pow2(NI + NJ);
    constexpr int VJ = next_pow2(NI * NJ);

    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(p.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned)p.ntasks) break;
        const JKTask task = p.tasks[t];
        const int k0 = task.k0;
        const int cbase = task.chunk * JK_CHUNK;
        const int l0 = cbase + 2 * lane;
        const bool lane_in = l0 < ldd;
        const int kk_lo = cbase > k0 ? cbase - k0 : 0;                    // sub-rows k < cbase do not reach the chunk
        const int kk_hi = (k0 + JK_KN <= n) ? JK_KN : (n - k0);
        const bool full_chunk = k0 >= cbase + JK_CHUNK - 1;               // every sub-row covers the whole chunk

        // stage the D2[k, l-chunk] tile and clear the J~[k, l-chunk] slab (warp-private)
#pragma unroll 4
        for (int kk = 0; kk < JK_KN; ++kk) {
            const int k = k0 + kk;
            s_jt[kk * 32 + lane] = make_double2(0.0, 0.0);
            s_d2[kk * 32 + lane] = (k < n && lane_in) ? ldg2(p.d2 + (size_t)k * ldd + l0) : make_double2(0.0, 0.0);
        }
        __syncwarp();

        int cur_i0 = -1;
        double2 accA[NS][NI], di[NS][NI];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int a = 0; a < NI; ++a) accA[s][a] = di[s][a] = make_double2(0.0, 0.0);

        auto flushA = [&]() {
            if (WANTK && cur_i0 >= 0 && lane_in) {
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int a = 0; a < NI; ++a)
                        if (cur_i0 + a < n) {
                            double *dst = p.kt + s * nl + (size_t)(cur_i0 + a) * ldd + l0;
                            red_add(dst, accA[s][a].x);
                            red_add(dst + 1, accA[s][a].y);
                        }
            }
        };

        for (int rb = task.rb_begin; rb < task.rb_begin + task.rb_count; ++rb) {
            const RowBlock blk = p.rowblocks[rb];
            const int i0 = blk.i0, j0 = blk.j0;
            if (i0 != cur_i0) {
                flushA();
                cur_i0 = i0;
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int a = 0; a < NI; ++a) {
                        accA[s][a] = make_double2(0.0, 0.0);
                        if (WANTK)
                            di[s][a] = (i0 + a < n && lane_in) ? ldg2(p.dsp + s * nl + (size_t)(i0 + a) * ldd + l0)
                                                                : make_double2(0.0, 0.0);
                    }
            }
            double2 dj[NS][NJ], accB[NS][NJ];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int b = 0; b < NJ; ++b) {
                    accB[s][b] = make_double2(0.0, 0.0);
                    dj[s][b] = make_double2(0.0, 0.0);
                    if (WANTK && j0 + b < n && lane_in) dj[s][b] = ldg2(p.dsp + s * nl + (size_t)(j0 + b) * ldd + l0);
                }
            double accJ[NI][NJ], d2ij[NI][NJ];
            const double *rp[NI][NJ];
            bool rvalid[NI][NJ];
#pragma unroll
            for (int a = 0; a < NI; ++a)
#pragma unroll
                for (int b = 0; b < NJ; ++b) {
                    const int i = i0 + a, j = j0 + b;
                    const int64_t ij = pair_index(i, j);
                    rvalid[a][b] = (i < n) && (j <= i) && ij >= p.ij_begin && ij < p.ij_end;
                    accJ[a][b] = 0.0;
                    d2ij[a][b] = rvalid[a][b] ? __ldg(p.d2 + (size_t)i * ldd + j) : 0.0;
                    rp[a][b] = rvalid[a][b] ? p.P + p.rowoff[ij - p.ij_begin] + l0 : p.P;
                }
            // interior row-block: every (i,j) is a valid row and every k of the range is < i
            bool fast = full_chunk && (i0 >= k0 + JK_KN);
#pragma unroll
            for (int a = 0; a < NI; ++a)
#pragma unroll
                for (int b = 0; b < NJ; ++b) fast = fast && rvalid[a][b];

            auto body = [&](auto masked_tag) {
                constexpr bool MASKED = decltype(masked_tag)::value;
                int64_t te = subrow_offset(k0 + kk_lo);
#pragma unroll 2
                for (int kk = kk_lo; kk < kk_hi; ++kk) {
                    const int k = k0 + kk;
                    double2 v[NI][NJ];
#pragma unroll
                    for (int a = 0; a < NI; ++a)
#pragma unroll
                        for (int b = 0; b < NJ; ++b) {
                            if (MASKED) {
                                const int i = i0 + a;
                                const int lmax = !rvalid[a][b] ? -1 : (k < i ? k : (k == i ? j0 + b : -1));
                                v[a][b] = (l0 <= lmax) ? ldg_stream(rp[a][b] + te) : make_double2(0.0, 0.0);
                            } else {
                                v[a][b] = ldg_stream(rp[a][b] + te);
                            }
                        }
                    te += pad_sub(k + 1);
                    const double2 d2 = s_d2[kk * 32 + lane];
                    double2 jt2 = make_double2(0.0, 0.0);
                    double cd[VCD];
#pragma unroll
                    for (int q = 0; q < VCD; ++q) cd[q] = 0.0;
                    double ui[NS][NI], uj[NS][NJ];
                    if (WANTK) {
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
#pragma unroll
                            for (int a = 0; a < NI; ++a)
                                ui[s][a] = (!MASKED || i0 + a < n) ? __ldg(p.dsp + s * nl + (size_t)(i0 + a) * ldd + k) : 0.0;
#pragma unroll
                            for (int b = 0; b < NJ; ++b)
                                uj[s][b] = (!MASKED || j0 + b < n) ? __ldg(p.dsp + s * nl + (size_t)(j0 + b) * ldd + k) : 0.0;
                        }
                    }
#pragma unroll
                    for (int a = 0; a < NI; ++a)
#pragma unroll
                        for (int b = 0; b < NJ; ++b) {
                            const double2 x = v[a][b];
                            accJ[a][b] = fma(x.x, d2.x, fma(x.y, d2.y, accJ[a][b]));
                            jt2.x = fma(x.x, d2ij[a][b], jt2.x);
                            jt2.y = fma(x.y, d2ij[a][b], jt2.y);
                            if (WANTK) {
#pragma unroll
                                for (int s = 0; s < NS; ++s) {
                                    accA[s][a].x = fma(x.x, uj[s][b], accA[s][a].x);   // K~[i,l] += v D[j,k]
                                    accA[s][a].y = fma(x.y, uj[s][b], accA[s][a].y);
                                    accB[s][b].x = fma(x.x, ui[s][a], accB[s][b].x);   // K~[j,l] += v D[i,k]
                                    accB[s][b].y = fma(x.y, ui[s][a], accB[s][b].y);
                                }
                            }
                        }
                    {
                        double2 cur = s_jt[kk * 32 + lane];
                        cur.x += jt2.x;
                        cur.y += jt2.y;
                        s_jt[kk * 32 + lane] = cur;
                    }
                    if (WANTK) {
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
#pragma unroll
                            for (int q = 0; q < VCD; ++q) cd[q] = 0.0;
#pragma unroll
                            for (int a = 0; a < NI; ++a)
#pragma unroll
                                for (int b = 0; b < NJ; ++b) {
                                    const double2 x = v[a][b];
                                    cd[a] = fma(x.x, dj[s][b].x, fma(x.y, dj[s][b].y, cd[a]));            // K~[i,k] += v D[j,l]
                                    cd[NI + b] = fma(x.x, di[s][a].x, fma(x.y, di[s][a].y, cd[NI + b]));  // K~[j,k] += v D[i,l]
                                }
                            warp_reduce_multi<VCD>(cd, lane);
                            constexpr int GROUP = 32 / VCD;
                            const int g = lane / GROUP;
                            if ((lane % GROUP) == 0 && g < NI + NJ) {
                                const int row = g < NI ? i0 + g : j0 + (g - NI);
                                if (row < n) red_add(p.kt + s * nl + (size_t)row * ldd + k, cd[0]);
                            }
                        }
                    }
                }
            };
            if (fast)
                body(std::false_type{});
            else
                body(std::true_type{});

            // flush the per-row-block accumulators
            if (WANTK && lane_in) {
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int b = 0; b < NJ; ++b)
                        if (j0 + b < n) {
                            double *dst = p.kt + s * nl + (size_t)(j0 + b) * ldd + l0;
                            red_add(dst, accB[s][b].x);
                            red_add(dst + 1, accB[s][b].y);
                        }
            }
            {
                double jv[VJ];
#pragma unroll
                for (int q = 0; q < VJ; ++q) jv[q] = 0.0;
#pragma unroll
                for (int a = 0; a < NI; ++a)
#pragma unroll
                    for (int b = 0; b < NJ; ++b) jv[a * NJ + b] = accJ[a][b];
                warp_reduce_multi<VJ>(jv, lane);
                constexpr int GROUP = 32 / VJ;
                const int g = lane / GROUP;
                if ((lane % GROUP) == 0 && g < NI * NJ) {
                    const int i = i0 + g / NJ, j = j0 + g % NJ;
                    if (i < n && j <= i) red_add(p.jt + (size_t)i * ldd + j, jv[0]);  // rows outside the shard add 0
                }
            }
        }
        flushA();
        // flush the J~[k, l-chunk] slab
        __syncwarp();
        if (lane_in) {
            for (int kk = kk_lo; kk < kk_hi; ++kk) {
                const int k = k0 + kk;
                const double2 cur = s_jt[kk * 32 + lane];
                if (l0 <= k) red_add(p.jt + (size_t)k * ldd + l0, cur.x);
                if (l0 + 1 <= k) red_add(p.jt + (size_t)k * ldd + l0 + 1, cur.y);
            }
        }
        __syncwarp();
    }
}

}  // namespace sqcc
#include "jk_pipe.cuh"
namespace sqcc {

// ---- simple cross-check kernel: one CTA per row, one global atomic per contribution ------------------
template <int NSPIN, bool WANTK>
__global__ void __launch_bounds__(256) jk_simple_kernel(const JKParams p, int64_t nrows, int i_begin)
{
    const int n = p.n, ldd = p.ldd;
    const size_t nl = (size_t)n * ldd;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
        // (i, j) of this row: linear search from the shard's first i (rows are few per block)
        int i = i_begin;
        int64_t ij = p.ij_begin + row;
        while (pair_index(i + 1, 0) <= ij) ++i;
        const int j = (int)(ij - pair_index(i, 0));
        const double *base = p.P + p.rowoff[row];
        const double d2ij = p.d2[(size_t)i * ldd + j];
        double jrow = 0.0;
        for (int k = warp; k <= i; k += nwarps) {
            const int lmax = k < i ? k : j;
            const double *sub = base + subrow_offset(k);
            for (int l = lane; l <= lmax; l += 32) {
                const double v = sub[l];
                jrow += v * p.d2[(size_t)k * ldd + l];
                atomicAdd(p.jt + (size_t)k * ldd + l, v * d2ij);
                if (WANTK) {
#pragma unroll
                    for (int s = 0; s < NSPIN; ++s) {
                        const double *d = p.dsp + s * nl;
                        double *kt = p.kt + s * nl;
                        atomicAdd(kt + (size_t)i * ldd + k, v * d[(size_t)j * ldd + l]);
                        atomicAdd(kt + (size_t)i * ldd + l, v * d[(size_t)j * ldd + k]);
                        atomicAdd(kt + (size_t)j * ldd + k, v * d[(size_t)i * ldd + l]);
                        atomicAdd(kt + (size_t)j * ldd + l, v * d[(size_t)i * ldd + k]);
                    }
                }
            }
        }
        atomicAdd(p.jt + (size_t)i * ldd + j, jrow);
    }
}

// ---- prep / finalize ---------------------------------------------------------------------------------
__global__ void jk_prep_kernel(const double *__restrict__ D, int nspin, int n, int ldd, double *__restrict__ dwork,
                               double *__restrict__ acc, unsigned int *counter)
{
    const size_t nl = (size_t)n * ldd;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) *counter = 0u;
    if (idx >= nl) return;
    const int r = (int)(idx / ldd), c = (int)(idx % ldd);
    double da = 0.0, db = 0.0;
    if (c < n) {
        da = D[(size_t)r * n + c];
        if (nspin == 2) db = D[(size_t)n * n + (size_t)r * n + c];
    }
    dwork[idx] = 2.0 * (da + db);
    dwork[nl + idx] = da;
    dwork[2 * nl + idx] = db;
    acc[idx] = 0.0;
    acc[nl + idx] = 0.0;
    acc[2 * nl + idx] = 0.0;
}

__global__ void jk_finalize_kernel(const double *__restrict__ acc, int nspin, int want_k, int n, int ldd,
                                   double *__restrict__ J, double *__restrict__ K)
{
    const size_t nl = (size_t)n * ldd;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int pr = (int)(idx / n), q = (int)(idx % n);
    const int hi = pr >= q ? pr : q, lo = pr >= q ? q : pr;
    const double jt = acc[(size_t)hi * ldd + lo];
    J[idx] = pr == q ? 2.0 * jt : jt;
    if (want_k) {
        for (int s = 0; s < nspin; ++s) {
            const double *kt = acc + (1 + s) * nl;
            K[(size_t)s * n * n + idx] = kt[(size_t)pr * ldd + q] + kt[(size_t)q * ldd + pr];
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------
static int build_one_schedule(sqcc_ctx *ctx, JKSchedule &sc, int ni, int nj)
{
    sc.ni = ni;
    sc.nj = nj;
    sc.h_rowblocks.clear();
    sc.h_tasks.clear();
    const int n = ctx->n;
    // row-blocks: i0 descending (long rows first), j0 ascending
    if (ctx->ij_end > ctx->ij_begin) {
        int i_hi = ((ctx->i_end - 1) / ni) * ni, i_lo = (ctx->i_begin / ni) * ni;
        for (int i0 = i_hi; i0 >= i_lo; i0 -= ni) {
            int imax = i0 + ni - 1 < n - 1 ? i0 + ni - 1 : n - 1;
            for (int j0 = 0; j0 <= imax; j0 += nj) {
                // keep the row-block only if at least one of its rows lies in the shard
                int64_t first = pair_index(i0, j0), last = pair_index(imax, (j0 + nj - 1) < imax ? (j0 + nj - 1) : imax);
                if (last < ctx->ij_begin || first >= ctx->ij_end) continue;
                sc.h_rowblocks.push_back({i0, j0});
            }
        }
    }
    const int nrb = (int)sc.h_rowblocks.size();
    // group size: aim for >= ~16 tasks per resident warp, at least 4 and at most 32 row-blocks per task
    int64_t warps = (int64_t)ctx->sm_count * 12;
    int nkr = (ctx->i_end + JK_KN - 1) / JK_KN;
    int nch = (ctx->i_end + JK_CHUNK - 1) / JK_CHUNK;
    int64_t cells = (int64_t)nrb * nkr * (nch + 1) / 2 / 2;  // rough count of (row-block, k-range, chunk) cells
    int group = (int)(cells / (warps * 16));
    if (group < 4) group = 4;
    if (group > 32) group = 32;
    for (int g0 = 0; g0 < nrb; g0 += group) {
        int g1 = g0 + group < nrb ? g0 + group : nrb;
        // largest i in this group (list is i0-descending): first entry
        int imax = sc.h_rowblocks[g0].i0 + ni - 1;
        if (imax > n - 1) imax = n - 1;
        for (int k0 = 0; k0 <= imax; k0 += JK_KN) {
            // restrict to row-blocks of the group that contain some i >= k0 (a prefix of the group)
            int cnt = 0;
            while (g0 + cnt < g1 && sc.h_rowblocks[g0 + cnt].i0 + ni - 1 >= k0) ++cnt;
            if (cnt == 0) continue;
            int kmax = k0 + JK_KN - 1 < imax ? k0 + JK_KN - 1 : imax;
            for (int c = 0; c * JK_CHUNK <= kmax; ++c) sc.h_tasks.push_back({k0, c, g0, cnt});
        }
    }
    sc.n_tasks = (int)sc.h_tasks.size();
    if (sc.d_rowblocks) cudaFree(sc.d_rowblocks);
    if (sc.d_tasks) cudaFree(sc.d_tasks);
    sc.d_rowblocks = nullptr;
    sc.d_tasks = nullptr;
    SQCC_CUDA(cudaMalloc(&sc.d_rowblocks, sizeof(RowBlock) * (nrb > 0 ? nrb : 1)));
    SQCC_CUDA(cudaMalloc(&sc.d_tasks, sizeof(JKTask) * (sc.n_tasks > 0 ? sc.n_tasks : 1)));
    SQCC_CUDA(cudaMemcpyAsync(sc.d_rowblocks, sc.h_rowblocks.data(), sizeof(RowBlock) * nrb, cudaMemcpyHostToDevice,
                              ctx->stream));
    SQCC_CUDA(cudaMemcpyAsync(sc.d_tasks, sc.h_tasks.data(), sizeof(JKTask) * sc.n_tasks, cudaMemcpyHostToDevice,
                              ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    return SQCC_OK;
}

int jk_build_schedule(sqcc_ctx *ctx)
{
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[0], JK_NI, JK_NJ));
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[1], 2, 2));
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[2], 4, 4));
    const size_t nl = (size_t)ctx->n * ctx->ldd;
    if (ctx->d_dwork) cudaFree(ctx->d_dwork);
    if (ctx->d_acc) cudaFree(ctx->d_acc);
    ctx->d_dwork = ctx->d_acc = nullptr;
    SQCC_CUDA(cudaMalloc(&ctx->d_dwork, sizeof(double) * 3 * nl));
    SQCC_CUDA(cudaMalloc(&ctx->d_acc, sizeof(double) * 3 * nl));
    if (!ctx->d_counter) SQCC_CUDA(cudaMalloc(&ctx->d_counter, sizeof(unsigned int)));
    // row (i,0) offsets for arithmetic row addressing in the TMA kernel
    std::vector<int64_t> ioff(ctx->n + 1);
    ioff[0] = 0;
    for (int i = 0; i < ctx->n; ++i) ioff[i + 1] = ioff[i] + (int64_t)(i + 1) * subrow_offset(i) + subrow_offset(i + 1);
    if (ctx->d_ioff) cudaFree(ctx->d_ioff);
    ctx->d_ioff = nullptr;
    SQCC_CUDA(cudaMalloc(&ctx->d_ioff, sizeof(int64_t) * (ctx->n + 1)));
    SQCC_CUDA(cudaMemcpyAsync(ctx->d_ioff, ioff.data(), sizeof(int64_t) * (ctx->n + 1), cudaMemcpyHostToDevice, ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    {
        int i = ctx->i_begin;
        int64_t j = ctx->ij_begin - pair_index(i, 0);
        ctx->shard_base = ioff[i] + j * subrow_offset(i) + subrow_offset(j);
    }
    return SQCC_OK;
}

template <int NI, int NJ, int NSPIN, bool WANTK, int WARPS, int STAGES, bool SLAB, bool RING>
static int launch_pipe_cfg(sqcc_ctx *ctx, const JKParams &p);

// (WARPS, STAGES) of the pipelined kernel: tuning knob jk_cfg (0 = default).
template <int NI, int NJ, int NSPIN, bool WANTK>
static int launch_pipe(sqcc_ctx *ctx, const JKParams &p)
{
    switch (ctx->jk_cfg) {
    case 1: return launch_pipe_cfg<NI, NJ, NSPIN, WANTK, 12, 3, false, true>(ctx, p);
    case 2: return launch_pipe_cfg<NI, NJ, NSPIN, WANTK, 8, 3, true, true>(ctx, p);
    case 3: return launch_pipe_cfg<NI, NJ, NSPIN, WANTK, 8, 1, true, false>(ctx, p);
    default: return launch_pipe_cfg<NI, NJ, NSPIN, WANTK, 11, 2, true, true>(ctx, p);
    }
}

template <int NI, int NJ, int NSPIN, bool WANTK, int WARPS, int STAGES, bool SLAB, bool RING>
static int launch_pipe_cfg(sqcc_ctx *ctx, const JKParams &p)
{
    constexpr int NVP = ((WANTK ? NSPIN : 1) * (NI + NJ) + 7) / 8;
    const size_t smem = (size_t)WARPS * jk_pipe_warp_bytes<NI, NJ, STAGES, SLAB, RING, NVP>();
    auto kern = jk_pipe_kernel<NI, NJ, NSPIN, WANTK, WARPS, STAGES, SLAB, RING>;
    SQCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SQCC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = ctx->sm_count * per_sm;
    int max_useful = (p.ntasks + WARPS - 1) / WARPS;
    if (grid > max_useful) grid = max_useful > 0 ? max_useful : 1;
    kern<<<grid, WARPS * 32, smem, ctx->stream>>>(p);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

template <int NI, int NJ, int NSPIN, bool WANTK>
static int launch_tiled(sqcc_ctx *ctx, const JKParams &p)
{
    constexpr int WARPS = 8;
    const size_t smem = (size_t)WARPS * 2 * JK_KN * 32 * sizeof(double2);
    auto kern = jk_tiled_kernel<NI, NJ, NSPIN, WANTK, WARPS>;
    SQCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SQCC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = ctx->sm_count * per_sm;
    int max_useful = (p.ntasks + WARPS - 1) / WARPS;
    if (grid > max_useful) grid = max_useful > 0 ? max_useful : 1;
    kern<<<grid, WARPS * 32, smem, ctx->stream>>>(p);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

int jk_build(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k, double *d_J, double *d_K)
{
    SQCC_REQUIRE(ctx->d_eri, "no ERI attached (sqcc_eri_attach)");
    SQCC_REQUIRE(nspin == 1 || nspin == 2, "nspin must be 1 or 2");
    SQCC_REQUIRE(d_D && d_J && (!want_k || d_K), "null D/J/K pointer");
    const int n = ctx->n, ldd = ctx->ldd;
    const size_t nl = (size_t)n * ldd;
    jk_prep_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, ctx->stream>>>(d_D, nspin, n, ldd, ctx->d_dwork, ctx->d_acc,
                                                                         ctx->d_counter);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());

    // default: 4x4 row-blocks (RHF, J-only), 2x2 for the dual-density UHF pass; cfg 1..3 select 4x2 tuning variants
    const bool wide = ctx->jk_variant == 0 && (ctx->jk_cfg == 0 || ctx->jk_cfg == 4) && !(nspin == 2 && want_k);
    const bool uhf = nspin == 2 && want_k;
    const bool uhf_wide = uhf && ctx->jk_variant == 0 && ctx->jk_cfg != 5;   // 4x2 row-blocks, two reduction passes
    const JKSchedule &sc = ctx->sched[uhf ? (uhf_wide ? 0 : 1) : (wide ? 2 : 0)];
    JKParams p;
    p.P = ctx->d_eri;
    p.rowoff = ctx->d_rowoff;
    p.ioff = ctx->d_ioff;
    p.shard_base = ctx->shard_base;
    p.ij_begin = ctx->ij_begin;
    p.ij_end = ctx->ij_end;
    p.n = n;
    p.ldd = ldd;
    p.d2 = ctx->d_dwork;
    p.dsp = ctx->d_dwork + nl;
    p.jt = ctx->d_acc;
    p.kt = ctx->d_acc + nl;
    p.tasks = sc.d_tasks;
    p.rowblocks = sc.d_rowblocks;
    p.ntasks = sc.n_tasks;
    p.counter = ctx->d_counter;

    SQCC_TRY(timer_start(ctx, 0));
    if (ctx->jk_variant == 1) {
        int64_t nrows = ctx->ij_end - ctx->ij_begin;
        int grid = (int)(nrows < (int64_t)ctx->sm_count * 8 ? nrows : (int64_t)ctx->sm_count * 8);
        if (grid < 1) grid = 1;
        if (!want_k)
            jk_simple_kernel<1, false><<<grid, 256, 0, ctx->stream>>>(p, nrows, ctx->i_begin);
        else if (nspin == 1)
            jk_simple_kernel<1, true><<<grid, 256, 0, ctx->stream>>>(p, nrows, ctx->i_begin);
        else
            jk_simple_kernel<2, true><<<grid, 256, 0, ctx->stream>>>(p, nrows, ctx->i_begin);
        ctx->launches++;
        SQCC_CUDA(cudaGetLastError());
    } else if (uhf_wide && p.ntasks > 0) {
        SQCC_TRY((launch_pipe_cfg<4, 2, 2, true, 8, 2, true, true>(ctx, p)));
    } else if (wide && p.ntasks > 0) {
        if (!want_k)
            SQCC_TRY((launch_pipe_cfg<4, 4, 1, false, 8, 2, true, true>(ctx, p)));
        else
            SQCC_TRY((launch_pipe_cfg<4, 4, 1, true, 8, 2, true, true>(ctx, p)));
    } else if (ctx->jk_variant == 0 && p.ntasks > 0) {
        if (!want_k)
            SQCC_TRY((launch_pipe<JK_NI, JK_NJ, 1, false>(ctx, p)));
        else if (nspin == 1)
            SQCC_TRY((launch_pipe<JK_NI, JK_NJ, 1, true>(ctx, p)));
        else
            SQCC_TRY((launch_pipe<2, 2, 2, true>(ctx, p)));
    } else if (p.ntasks > 0) {
        if (!want_k)
            SQCC_TRY((launch_tiled<JK_NI, JK_NJ, 1, false>(ctx, p)));
        else if (nspin == 1)
            SQCC_TRY((launch_tiled<JK_NI, JK_NJ, 1, true>(ctx, p)));
        else
            SQCC_TRY((launch_tiled<2, 2, 2, true>(ctx, p)));
    }
    SQCC_TRY(timer_stop(ctx, 0));

    SQCC_TRY(timer_start(ctx, 1));
    jk_finalize_kernel<<<(unsigned)(((size_t)n * n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_acc, nspin, want_k, n,
                                                                                         ldd, d_J, d_K);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    SQCC_TRY(timer_stop(ctx, 1));
    return SQCC_OK;
}

}  // namespace sqcc
