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
#include <stdlib.h>

#include "common.cuh"

namespace sqcc {

struct JKParams {
    const double *P;
    const int64_t *rowoff;  // [nrows + 1] shard-relative offsets of the held rows (simple kernel)
    const int64_t *rowij;   // [nrows] pair index of the held rows (simple kernel)
    const int64_t *ioff;    // [n] shard-relative: row (i, j) starts at ioff[i] + j Te(i) + Te(j) (held rows only)
    int n, ldd;
    const double *d2;   // [n][ldd]   2*(Da+Db)
    const double *dsp;  // [nspin][n][ldd]
    double *jt;         // [n][ldd]
    double *kt;         // [nspin][n][ldd]
    const JKTask *tasks;
    const RowBlock *rowblocks;
    int ntasks;
    unsigned int *counter;
    unsigned long long *trace;  // optional (tuning): per warp {first task start, last task end, tasks done} in ns
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

}  // namespace sqcc
#include "jk_pipe.cuh"
namespace sqcc {

// ---- simple cross-check kernel: one CTA per row, one global atomic per contribution ------------------
template <int NSPIN, bool WANTK>
__global__ void __launch_bounds__(256) jk_simple_kernel(const JKParams p, int64_t nrows)
{
    const int n = p.n, ldd = p.ldd;
    const size_t nl = (size_t)n * ldd;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int64_t ij = p.rowij[row];
        int i = (int)((sqrt(8.0 * (double)ij + 1.0) - 1.0) * 0.5);
        while (pair_index(i, 0) > ij) --i;
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
    // row-blocks: i-blocks anchored at the top of the i range (the longest rows sit in full blocks; the clipped block
    // is the one at i = 0), i0 descending (long rows first), j0 ascending; blocks without a held row are dropped
    for (int i0 = n - ni; i0 > -ni; i0 -= ni) {
        const int ilo = i0 < 0 ? 0 : i0;   // the lowest block is stored as i0 = 0 with its upper rows masked off
        const int ihi = i0 + ni;           // one past the last i of the block
        if (ihi <= ctx->i_begin || ilo >= ctx->i_end) continue;
        for (int j0 = 0; j0 <= ihi - 1; j0 += nj) {
            unsigned int mask = 0;
            for (int a = 0; a < ni; ++a)
                for (int b = 0; b < nj; ++b) {
                    const int i = ilo + a, j = j0 + b;
                    if (i < ihi && i < n && j >= ctx->jlo[i] && j < ctx->jhi[i]) mask |= 1u << (a * nj + b);
                }
            if (mask) sc.h_rowblocks.push_back({ilo, j0, mask, 0});
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
    if (const char *e = getenv("SQCC_JK_GROUP")) {   // tuning knob: fixed row-blocks per task
        int g = atoi(e);
        if (g >= 1 && g <= 32) group = g;
    }
    // Guided grouping: a task lasts ~(row-blocks in its group) x 16 sub-rows, and nearly all (k-range, chunk) tasks of a
    // group are full-size, so with a fixed group the last, partially filled wave of tasks idles most warps for one task
    // duration (~85 us of a 0.53 ms kernel at N = 222, ~90 us of the 1.26 ms shards at 8 GPUs).  The group therefore
    // shrinks towards the end of the list: never more row-blocks than would leave `tail_waves` tasks per resident warp.
    std::vector<int64_t> cells_after(nrb + 1, 0);   // (k-range, chunk) cells of row-blocks g .. nrb-1
    for (int g = nrb - 1; g >= 0; --g) {
        int imax = sc.h_rowblocks[g].i0 + ni - 1;
        if (imax > n - 1) imax = n - 1;
        int64_t t = 0;
        for (int k0 = 0; k0 <= imax; k0 += JK_KN) t += (k0 + JK_KN - 1 < imax ? k0 + JK_KN - 1 : imax) / JK_CHUNK + 1;
        cells_after[g] = cells_after[g + 1] + t;
    }
    const int64_t resident = (int64_t)ctx->sm_count * 8;
    const double tail_waves = getenv("SQCC_JK_TAIL") ? atof(getenv("SQCC_JK_TAIL")) : 2.0;
    for (int g0 = 0; g0 < nrb;) {
        int gsz = group;
        if (tail_waves > 0.0) {
            // cells_after / gsz tasks remain if the rest is cut in groups of gsz: keep that above tail_waves x resident
            const int64_t lim = (int64_t)((double)cells_after[g0] / (tail_waves * (double)resident));
            if (lim < gsz) gsz = (int)(lim < 1 ? 1 : lim);
        }
        int g1 = g0 + gsz < nrb ? g0 + gsz : nrb;
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
        g0 = g1;
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
    ctx->d_dwork = nullptr;
    SQCC_CUDA(cudaMalloc(&ctx->d_dwork, sizeof(double) * 3 * nl));
    SQCC_TRY(comm_alloc(ctx));   // accumulators live in the peer-mappable exchange region
    if (!ctx->d_counter) SQCC_CUDA(cudaMalloc(&ctx->d_counter, sizeof(unsigned int)));
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
    ctx->trace_warps = grid * WARPS;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

// prep + the fused kernel: leaves this rank's partial J~/K~ in the context's accumulators
int jk_partial(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k)
{
    SQCC_REQUIRE(ctx->d_eri, "no ERI attached (sqcc_eri_attach)");
    SQCC_REQUIRE(nspin == 1 || nspin == 2, "nspin must be 1 or 2");
    SQCC_REQUIRE(d_D, "null D pointer");
    const int n = ctx->n, ldd = ctx->ldd;
    const size_t nl = (size_t)n * ldd;
    jk_prep_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, ctx->stream>>>(d_D, nspin, n, ldd, ctx->d_dwork, ctx->d_acc,
                                                                         ctx->d_counter);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());

    // default: 4x4 row-blocks (RHF, J-only), 2x2 for the dual-density UHF pass; cfg 1..3 select 4x2 tuning variants
    if (ctx->jk_variant == 2) ctx->jk_cfg = 3;   // direct 128-bit global loads (no shared-memory ring)
    const bool wide = ctx->jk_variant == 0 && (ctx->jk_cfg == 0 || ctx->jk_cfg == 4) && !(nspin == 2 && want_k);
    const bool uhf = nspin == 2 && want_k;
    const bool uhf_wide = uhf && ctx->jk_variant == 0 && ctx->jk_cfg != 5;   // 4x2 row-blocks, two reduction passes
    const JKSchedule &sc = ctx->sched[uhf ? (uhf_wide ? 0 : 1) : (wide ? 2 : 0)];
    JKParams p;
    p.P = ctx->d_eri;
    p.rowoff = ctx->d_rowoff;
    p.rowij = ctx->d_rowij;
    p.ioff = ctx->d_ioff;
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
    p.trace = ctx->d_trace;

    SQCC_TRY(timer_start(ctx, 0));
    if (ctx->jk_variant == 1) {
        int64_t nrows = ctx->nrows;
        int grid = (int)(nrows < (int64_t)ctx->sm_count * 8 ? nrows : (int64_t)ctx->sm_count * 8);
        if (grid < 1) grid = 1;
        if (!want_k)
            jk_simple_kernel<1, false><<<grid, 256, 0, ctx->stream>>>(p, nrows);
        else if (nspin == 1)
            jk_simple_kernel<1, true><<<grid, 256, 0, ctx->stream>>>(p, nrows);
        else
            jk_simple_kernel<2, true><<<grid, 256, 0, ctx->stream>>>(p, nrows);
        ctx->launches++;
        SQCC_CUDA(cudaGetLastError());
    } else if (uhf_wide && p.ntasks > 0) {
        SQCC_TRY((launch_pipe_cfg<4, 2, 2, true, 8, 2, true, true>(ctx, p)));   // a deeper ring changes nothing (profiles/)
    } else if (wide && p.ntasks > 0) {
        if (!want_k)
            SQCC_TRY((launch_pipe_cfg<4, 4, 1, false, 8, 2, true, true>(ctx, p)));
        else
            SQCC_TRY((launch_pipe_cfg<4, 4, 1, true, 8, 2, true, true>(ctx, p)));
    } else if (p.ntasks > 0) {
        if (!want_k)
            SQCC_TRY((launch_pipe<JK_NI, JK_NJ, 1, false>(ctx, p)));
        else if (nspin == 1)
            SQCC_TRY((launch_pipe<JK_NI, JK_NJ, 1, true>(ctx, p)));
        else
            SQCC_TRY((launch_pipe<2, 2, 2, true>(ctx, p)));
    }
    SQCC_TRY(timer_stop(ctx, 0));
    return SQCC_OK;
}

int jk_build(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k, double *d_J, double *d_K)
{
    SQCC_REQUIRE(d_J && (!want_k || d_K), "null J/K pointer");
    SQCC_TRY(jk_partial(ctx, d_D, nspin, want_k));
    const int n = ctx->n, ldd = ctx->ldd;
    SQCC_TRY(timer_start(ctx, 1));
    if (ctx->comm.attached && ctx->comm.enabled && ctx->comm.world > 1 && !ctx->comm.same_device) {
        // ranks on distinct devices: exchange + finalize in one kernel over NVLink peer memory (comm.cu)
        SQCC_TRY(jk_reduce(ctx, nspin, want_k, d_J, d_K));
        SQCC_TRY(timer_stop(ctx, 1));
        return SQCC_OK;
    }
    jk_finalize_kernel<<<(unsigned)(((size_t)n * n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_acc, nspin, want_k, n,
                                                                                         ldd, d_J, d_K);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    SQCC_TRY(timer_stop(ctx, 1));
    return SQCC_OK;
}

}  // namespace sqcc
