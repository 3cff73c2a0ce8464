// Fused Coulomb/exchange build over the native v2 packed ERI shard (one pass, FP64).
//
// Replaces the loop nests hf_ksdft.py:483-487 (J), :491-495 (K), and the UHF/UKS variants :510-515,
// :522-526, :531-535.  For every stored v = (ij|kl) * f  (SURVEY.md A.1, D symmetric, D2 = 2*(Da+Db)):
//     J~[i,j] += v * D2[k,l]          J~[k,l] += v * D2[i,j]
//     K~[i,k] += v * D[j,l]           K~[i,l] += v * D[j,k]
//     K~[j,k] += v * D[i,l]           K~[j,l] += v * D[i,k]           (per spin density)
//     J = J~ + J~^T,  K = K~ + K~^T.
// The streaming kernel is in jk_stream.cuh (HBM-bound, no tensor cores: 1.5 flop per byte); this file holds the
// schedule, the prep / finalize kernels, the cross-check kernel and the launch logic.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace sqcc {

struct JKParams {
    const double *P;    // this rank's shard (tile range) of the native v2 tensor
    int n, ldd;
    const double *d2;   // [n][ldd]   2*(Da+Db)
    const double *dsp;  // [nspin][n][ldd]
    double *jt;         // [n][ldd]
    double *kt;         // [nspin][n][ldd]
    const JKTask *tasks;
    const RowBlock *rowblocks;
    int ntasks;
    unsigned int *counter;
    double fx_scale;            // deterministic mode: 2^shift of the fixed-point accumulators
    int rotate;                 // each task starts its row-block list at a task-dependent offset (see jk_stream.cuh)
    int overlap_prev;           // host side only: launch with programmatic stream serialization (see launch_stream)
};

// 16-byte read-only load that does not allocate in L1: with 227 KB of the SM's 256 KB carved out as shared memory the
// remaining L1 is left to the few spilled registers (a local-memory reload that misses L1 is an L2 round trip)
__device__ __forceinline__ double2 ldg2(const double *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

}  // namespace sqcc
#include "jk_stream.cuh"
namespace sqcc {

// ---- cross-check kernel: storage order, one global atomic per contribution ---------------------------------
template <int NSPIN, bool WANTK>
__global__ void __launch_bounds__(256) jk_simple_kernel(const JKParams p, const int64_t *__restrict__ ibase,
                                                        const int64_t *__restrict__ tfirst, int64_t t_lo, int64_t t_hi,
                                                        int64_t shard_base)
{
    const int n = p.n, ldd = p.ldd;
    const size_t nl = (size_t)n * ldd;
    const int c = blockIdx.y;
    int ib = 0;
    for (int64_t t = t_lo + blockIdx.x; t < t_hi; t += gridDim.x) {
        while (tfirst[ib + 1] <= t) ++ib;
        const int ihi = lay_ihi(n, ib), imax = ihi - 1, jb = (int)(t - tfirst[ib]);
        if (LAY_CHUNK * c > imax) continue;
        const double *tile = p.P + (ibase[ib] + (int64_t)jb * lay_tile_len(imax) - shard_base);
        for (int k = LAY_CHUNK * c; k <= imax; ++k) {
            const int w = lay_w(k - LAY_CHUNK * c);
            const double *piece = tile + lay_piece_offset(imax, c, k);
            for (int e = threadIdx.x; e < LAY_ROWS * w; e += blockDim.x) {
                const int r = e / w, x = e - r * w;
                const int i = ihi - 4 + (r & 3), j = 4 * jb + (r >> 2), l = LAY_CHUNK * c + x;
                if (!(i >= 0 && j <= i && k <= i && l <= (k < i ? k : j))) continue;
                const double v = piece[e];
                atomicAdd(p.jt + (size_t)i * ldd + j, v * p.d2[(size_t)k * ldd + l]);
                atomicAdd(p.jt + (size_t)k * ldd + l, v * p.d2[(size_t)i * ldd + j]);
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
    }
}

// ---- prep / finalize ---------------------------------------------------------------------------------
// dsum (optional, deterministic mode): [3] sums of |D2|, |Da|, |Db| accumulated with integer atomics on the raw
// bit patterns is not possible, so they are ordinary float atomics -- they only size the fixed-point scale, which is
// then rounded DOWN to a power of two with two guard bits, so last-bit differences of the sums never change it.
__global__ void jk_prep_kernel(const double *__restrict__ D, int nspin, int n, int ldd, double *__restrict__ dwork,
                               double *__restrict__ acc, unsigned int *counter, double *__restrict__ dsum)
{
    const size_t nl = (size_t)n * ldd;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) counter[0] = counter[1] = 0u;
    double da = 0.0, db = 0.0;
    if (idx < nl) {
        const int r = (int)(idx / ldd), c = (int)(idx % ldd);
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
    if (dsum) {
        double s2 = fabs(2.0 * (da + db)), sa = fabs(da), sb = fabs(db);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(dsum, s2);
            atomicAdd(dsum + 1, sa);
            atomicAdd(dsum + 2, sb);
        }
    }
}

// J = J~ + J~^T, K = K~ + K~^T.  inv_scale != 0: the accumulators are 64-bit fixed point (deterministic mode).
__global__ void jk_finalize_kernel(const double *__restrict__ acc, int nspin, int want_k, int n, int ldd,
                                   double *__restrict__ J, double *__restrict__ K, double inv_scale)
{
    const size_t nl = (size_t)n * ldd;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int pr = (int)(idx / n), q = (int)(idx % n);
    const int hi = pr >= q ? pr : q, lo = pr >= q ? q : pr;
    if (inv_scale != 0.0) {
        const long long *a = reinterpret_cast<const long long *>(acc);
        const long long jt = a[(size_t)hi * ldd + lo];
        J[idx] = (double)(pr == q ? 2 * jt : jt) * inv_scale;
        if (want_k)
            for (int s = 0; s < nspin; ++s)
                K[(size_t)s * n * n + idx] = (double)(a[(1 + s) * nl + (size_t)pr * ldd + q] + a[(1 + s) * nl + (size_t)q * ldd + pr]) * inv_scale;
        return;
    }
    const double jt = acc[(size_t)hi * ldd + lo];
    J[idx] = pr == q ? 2.0 * jt : jt;
    if (want_k) {
        for (int s = 0; s < nspin; ++s) {
            const double *kt = acc + (1 + s) * nl;
            K[(size_t)s * n * n + idx] = kt[(size_t)pr * ldd + q] + kt[(size_t)q * ldd + pr];
        }
    }
}

// max |stored value| of the shard (deterministic mode, once per attached tensor)
__global__ void __launch_bounds__(256) eri_absmax_kernel(const double *__restrict__ P, int64_t len, unsigned long long *out)
{
    double m = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < len; t += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, fabs(P[t]));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));   // non-negative doubles order like integers
}

// ---- host side -----------------------------------------------------------------------------------------
// cw: columns per task chunk (64; 32 for the spin-split UHF pass, whose tasks count 32-column chunks)
// part: 0 all tasks, 1 only tasks whose pieces are 48 / 64 columns wide, 2 only the narrow ones (16 / 32 columns: the first
//       two k-ranges of every chunk column), which the narrow-piece pass runs two sub-rows per iteration
static int build_one_schedule(sqcc_ctx *ctx, JKSchedule &sc, int nb, int cw = JK_CHUNK, int part = 0)
{
    sc.nb = nb;
    sc.h_rowblocks.clear();
    sc.h_tasks.clear();
    const int n = ctx->n;
    const int nib = lay_nblocks(n);
    // row-blocks of the held tiles: i-blocks descending (long rows first), j ascending.  The 4 x 2 shape makes two
    // row-blocks out of a tile (its halves b < 2 / b >= 2); a half whose rows all lie above the diagonal is dropped.
    for (int ib = nib - 1; ib >= 0; --ib) {
        const int imax = lay_ihi(n, ib) - 1;
        const int64_t t0 = ctx->h_tfirst[ib], t1 = ctx->h_tfirst[ib + 1];
        const int64_t a = t0 > ctx->t_begin ? t0 : ctx->t_begin, b = t1 < ctx->t_end ? t1 : ctx->t_end;
        for (int64_t t = a; t < b; ++t) {
            const int jb = (int)(t - t0);
            const int64_t base = ctx->h_ibase[ib] + (int64_t)jb * lay_tile_len(imax) - ctx->shard_base;
            for (int j0 = 4 * jb; j0 < 4 * jb + 4 && j0 <= imax; j0 += nb) sc.h_rowblocks.push_back({base, imax, j0});
        }
    }
    const int nrb = (int)sc.h_rowblocks.size();
    // group size: aim for >= ~16 tasks per resident warp, at least 8 and at most 32 row-blocks per task
    int64_t warps = (int64_t)ctx->sm_count * 12;
    int nkr = (ctx->i_end + JK_KN - 1) / JK_KN;
    int nch = (ctx->i_end + cw - 1) / cw;
    int64_t cells = (int64_t)nrb * nkr * (nch + 1) / 2 / 2;  // rough count of (row-block, k-range, chunk) cells
    int group = (int)(cells / (warps * 16));
    if (group < 8) group = 8;   // (N = 222 / 264: 8-16 row-blocks per task are 1-2 % faster than 4, the guided tail keeps the balance)
    if (group > 32) group = 32;
    if (const char *e = getenv("SQCC_JK_GROUP")) {   // tuning knob: fixed row-blocks per task
        int g = atoi(e);
        if (g >= 1 && g <= 32) group = g;
    }
    // Guided grouping: a task lasts ~(row-blocks in its group) x 16 sub-rows, and nearly all (k-range, chunk) tasks of a
    // group are full-size, so with a fixed group the last, partially filled wave of tasks idles most warps for one task
    // duration.  The group therefore shrinks towards the end of the list: never more row-blocks than would leave
    // `tail_waves` tasks per resident warp.
    auto cells_of = [&](int imax) {   // (k-range, chunk) tasks of one row-block: chunks c with 64 c <= k0
        int64_t t = 0;
        for (int k0 = 0; k0 <= imax; k0 += JK_KN) t += k0 / cw + 1;
        return t;
    };
    std::vector<int64_t> cells_after(nrb + 1, 0);
    for (int g = nrb - 1; g >= 0; --g) cells_after[g] = cells_after[g + 1] + cells_of(sc.h_rowblocks[g].imax);
    const int64_t resident = (int64_t)ctx->sm_count * 8;
    const double tail_waves = getenv("SQCC_JK_TAIL") ? atof(getenv("SQCC_JK_TAIL")) : 2.0;
    for (int g0 = 0; g0 < nrb;) {
        int gsz = group;
        if (tail_waves > 0.0) {
            const int64_t lim = (int64_t)((double)cells_after[g0] / (tail_waves * (double)resident));
            if (lim < gsz) gsz = (int)(lim < 1 ? 1 : lim);
        }
        int g1 = g0 + gsz < nrb ? g0 + gsz : nrb;
        const int imax = sc.h_rowblocks[g0].imax;   // largest i of the group (list is i-descending)
        for (int k0 = 0; k0 <= imax; k0 += JK_KN) {
            // row-blocks of the group that have sub-rows k >= k0 (a prefix of the group)
            int cnt = 0;
            while (g0 + cnt < g1 && sc.h_rowblocks[g0 + cnt].imax >= k0) ++cnt;
            if (cnt == 0) continue;
            for (int c = 0; c * cw <= k0; ++c) {
                const bool narrow = k0 - c * cw < 32;
                if (part == 0 || (part == 1 && !narrow) || (part == 2 && narrow)) sc.h_tasks.push_back({k0, c, g0, cnt});
            }
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
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[0], 4));
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[1], 2));
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[2], 4, 32));
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[3], 4, JK_CHUNK, 1));
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[4], 4, JK_CHUNK, 2));
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[5], 2, JK_CHUNK, 1));
    SQCC_TRY(build_one_schedule(ctx, ctx->sched[6], 2, JK_CHUNK, 2));
    const size_t nl = (size_t)ctx->n * ctx->ldd;
    if (ctx->d_dwork) cudaFree(ctx->d_dwork);
    ctx->d_dwork = nullptr;
    // [3][N][ldd] + one chunk of slack (the density tile loads of the last rows never leave the allocation) + [4] sums
    SQCC_CUDA(cudaMalloc(&ctx->d_dwork, sizeof(double) * (3 * nl + JK_CHUNK + 4)));
    SQCC_CUDA(cudaMemsetAsync(ctx->d_dwork, 0, sizeof(double) * (3 * nl + JK_CHUNK + 4), ctx->stream));
    SQCC_TRY(comm_alloc(ctx));   // accumulators live in the peer-mappable exchange region
    if (!ctx->d_counter) SQCC_CUDA(cudaMalloc(&ctx->d_counter, sizeof(unsigned long long) * 2));
    ctx->eri_absmax = -1.0;
    return SQCC_OK;
}

template <int NB, int NSPIN, bool WANTK, int WARPS, int STAGES, bool TMA, bool FIXED, bool HINT, int HM = 0>
static int launch_stream(sqcc_ctx *ctx, const JKParams &p)
{
    using SM = JKSmem<NB, (WANTK && HM != 1) ? NSPIN : 1, STAGES, HM>;
    const size_t smem = (size_t)WARPS * SM::WARP_BYTES;
    static_assert((size_t)WARPS * SM::WARP_BYTES <= 232448, "shared memory per CTA");
    auto kern = jk_stream_kernel<NB, NSPIN, WANTK, WARPS, STAGES, TMA, FIXED, HINT, HM>;
    SQCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = ctx->sm_count;
    int max_useful = (p.ntasks + WARPS - 1) / WARPS;
    if (grid > max_useful) grid = max_useful > 0 ? max_useful : 1;
    if (p.overlap_prev) {
        // programmatic dependent launch: the CTAs of this kernel may start as soon as the previous kernel's CTAs have
        // called griddepcontrol.launch_dependents and free their SM -- the narrow-piece pass and this kernel only add
        // into the same accumulators, so the tail of the one overlaps the ramp-up of the other
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(WARPS * 32);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SQCC_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    } else {
        kern<<<grid, WARPS * 32, smem, ctx->stream>>>(p);
    }
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

// STAGES counts ring slots (the R rows of one iteration + their density tile row).
template <int NB, int NSPIN, bool WANTK, int STAGES>
static int launch_stream_cfg(sqcc_ctx *ctx, const JKParams &p)
{
    // cfg 0: TMA ring, L2 cache-policy hints (default; adds the narrow-piece pass, see jk_partial)
    //     1: cp.async ring for every piece (A/B of the TMA path)   2: no L2 hints
    //     6: spin-split UHF   7: one pass over all tasks (no narrow-piece pass)
    // (retired after their A/B, logs in profiles/: 3 = no J~ slab + one more ring stage, 4 = half-piece ring slots)
    if (ctx->jk_fixed) return launch_stream<NB, NSPIN, WANTK, 8, STAGES, true, true, true>(ctx, p);
    switch (ctx->jk_cfg) {
    case 1: return launch_stream<NB, NSPIN, WANTK, 8, STAGES, false, false, true>(ctx, p);
    case 2: return launch_stream<NB, NSPIN, WANTK, 8, STAGES, true, false, false>(ctx, p);
    default: return launch_stream<NB, NSPIN, WANTK, 8, STAGES, true, false, true>(ctx, p);
    }
}

// deterministic mode: fixed-point scale 2^s with |every partial sum| * 2^s < 2^61
static int fixed_scale(sqcc_ctx *ctx, int nspin, double *scale)
{
    const size_t nl = (size_t)ctx->n * ctx->ldd;
    if (ctx->eri_absmax < 0.0) {
        unsigned long long *d_m = reinterpret_cast<unsigned long long *>(ctx->d_counter) + 1;
        SQCC_CUDA(cudaMemsetAsync(d_m, 0, sizeof(unsigned long long), ctx->stream));
        eri_absmax_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->native_len, d_m);
        ctx->launches++;
        unsigned long long bits = 0;
        SQCC_CUDA(cudaMemcpyAsync(&bits, d_m, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
        SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
        double m;
        memcpy(&m, &bits, sizeof(m));
        ctx->eri_absmax = m;
    }
    double sums[3];
    SQCC_CUDA(cudaMemcpyAsync(sums, ctx->d_dwork + 3 * nl + JK_CHUNK, sizeof(sums), cudaMemcpyDeviceToHost, ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    double dmax = sums[0] > sums[1] ? sums[0] : sums[1];
    if (nspin == 2 && sums[2] > dmax) dmax = sums[2];
    const double bound = ctx->eri_absmax * dmax;
    SQCC_REQUIRE(isfinite(bound), "non-finite operands in deterministic J/K mode");
    int e = 0;
    if (bound > 0.0) frexp(bound, &e);   // bound < 2^e
    // two guard bits: the sums come from float atomics and may differ in the last bits between runs and ranks; the
    // exponent only changes when a sum sits on a power of two, which the guard bits absorb on the safe side
    int shift = 59 - e;
    if (shift > 1000) shift = 1000;
    *scale = ldexp(1.0, shift);
    return SQCC_OK;
}

// prep + the fused kernel: leaves this rank's partial J~/K~ in the context's accumulators
int jk_partial(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k)
{
    SQCC_REQUIRE(ctx->d_eri, "no ERI attached (sqcc_eri_attach_tiles)");
    SQCC_REQUIRE(nspin == 1 || nspin == 2, "nspin must be 1 or 2");
    SQCC_REQUIRE(d_D, "null D pointer");
    const int n = ctx->n, ldd = ctx->ldd;
    const size_t nl = (size_t)n * ldd;
    double *dsum = ctx->jk_fixed ? ctx->d_dwork + 3 * nl + JK_CHUNK : nullptr;
    if (dsum) SQCC_CUDA(cudaMemsetAsync(dsum, 0, 4 * sizeof(double), ctx->stream));
    jk_prep_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, ctx->stream>>>(d_D, nspin, n, ldd, ctx->d_dwork, ctx->d_acc,
                                                                         ctx->d_counter, dsum);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());

    const bool uhf = nspin == 2 && want_k;
    // UHF: jk_cfg 6 = spin-split half-warps over 4 x 4 row-blocks and 32-column chunks (sched[2]); else 4 x 2 row-blocks
    const bool uhf_split = uhf && ctx->jk_variant == 0 && ctx->jk_cfg == 6 && !ctx->jk_fixed;
    // 4 x 4 shapes, default configuration: the tasks with 48 / 64-column pieces go through the TMA kernel, the narrow ones
    // (16 / 32 columns) through the narrow-piece pass (two sub-rows per iteration, one per half-warp)
    const bool two_pass = ctx->jk_variant == 0 && ctx->jk_cfg == 0 && !ctx->jk_fixed;
    const JKSchedule &sc = ctx->sched[uhf ? (uhf_split ? 2 : (two_pass ? 5 : 1)) : (two_pass ? 3 : 0)];
    const JKSchedule &sc_narrow = ctx->sched[uhf ? 6 : 4];
    JKParams p;
    p.P = ctx->d_eri;
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
    {
        const char *e = getenv("SQCC_JK_ROTATE");   // tuning knob: 0 = every task walks its row-block list from the start
        p.rotate = e ? atoi(e) : 1;
    }
    p.overlap_prev = 0;
    p.fx_scale = 0.0;
    ctx->fx_scale = 0.0;
    if (ctx->jk_fixed) {
        SQCC_REQUIRE(ctx->jk_variant == 0, "the deterministic mode uses the streaming kernel (variant 0)");
        SQCC_TRY(fixed_scale(ctx, nspin, &p.fx_scale));
        ctx->fx_scale = p.fx_scale;
    }

    SQCC_TRY(timer_start(ctx, 0));
    if (ctx->jk_variant == 1) {
        const int64_t tiles = ctx->t_end - ctx->t_begin;
        const dim3 grid((unsigned)(tiles < (int64_t)ctx->sm_count * 4 ? (tiles > 0 ? tiles : 1) : (int64_t)ctx->sm_count * 4),
                        (unsigned)((n - 1) / LAY_CHUNK + 1));
        if (!want_k)
            jk_simple_kernel<1, false><<<grid, 256, 0, ctx->stream>>>(p, ctx->d_ibase, ctx->d_tfirst, ctx->t_begin, ctx->t_end, ctx->shard_base);
        else if (nspin == 1)
            jk_simple_kernel<1, true><<<grid, 256, 0, ctx->stream>>>(p, ctx->d_ibase, ctx->d_tfirst, ctx->t_begin, ctx->t_end, ctx->shard_base);
        else
            jk_simple_kernel<2, true><<<grid, 256, 0, ctx->stream>>>(p, ctx->d_ibase, ctx->d_tfirst, ctx->t_begin, ctx->t_end, ctx->shard_base);
        ctx->launches++;
        SQCC_CUDA(cudaGetLastError());
    } else if (p.ntasks > 0 || two_pass) {
        if (two_pass && sc_narrow.n_tasks > 0) {   // narrow pieces first (short tasks; the long ones fill the tail)
            JKParams pn = p;
            pn.tasks = sc_narrow.d_tasks;
            pn.rowblocks = sc_narrow.d_rowblocks;
            pn.ntasks = sc_narrow.n_tasks;
            pn.counter = ctx->d_counter + 1;
            if (uhf)
                SQCC_TRY((launch_stream<2, 2, true, 8, 3, false, false, true, 2>(ctx, pn)));
            else if (!want_k)
                SQCC_TRY((launch_stream<4, 1, false, 8, 2, false, false, true, 2>(ctx, pn)));
            else
                SQCC_TRY((launch_stream<4, 1, true, 8, 2, false, false, true, 2>(ctx, pn)));
        }
        {
            const char *e = getenv("SQCC_JK_PDL");   // tuning knob: 0 = plain stream order between the two passes
            p.overlap_prev = (two_pass && sc_narrow.n_tasks > 0 && p.ntasks > 0 && (!e || atoi(e) != 0)) ? 1 : 0;
        }
        if (uhf_split)
            SQCC_TRY((launch_stream<4, 2, true, 8, 3, false, false, true, 1>(ctx, p)));
        else if (uhf)
            SQCC_TRY((launch_stream_cfg<2, 2, true, 3>(ctx, p)));
        else if (p.ntasks == 0)
            ;
        else if (!want_k)
            SQCC_TRY((launch_stream_cfg<4, 1, false, 2>(ctx, p)));
        else
            SQCC_TRY((launch_stream_cfg<4, 1, true, 2>(ctx, p)));
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
    jk_finalize_kernel<<<(unsigned)(((size_t)n * n + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->d_acc, nspin, want_k, n, ldd, d_J, d_K, ctx->fx_scale != 0.0 ? 1.0 / ctx->fx_scale : 0.0);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    SQCC_TRY(timer_stop(ctx, 1));
    return SQCC_OK;
}

}  // namespace sqcc
