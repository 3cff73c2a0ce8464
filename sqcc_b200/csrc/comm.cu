// Cross-rank sum of the partial J~/K~ work matrices over NVLink peer memory, fused with the J/K finalize step.
//
// Pair-block sharding leaves every rank with partial accumulators J~, K~a, K~b ([3][N][ldd], at most
// 5.9 MB at N = 494).  The path's one exchange step sums them over the ranks.  Instead of finalize -> NCCL all-reduce
// (two more launches plus the collective's own latency) ONE kernel per rank does the whole exchange on mapped peer
// memory (CUDA IPC between the per-GPU processes, plain pointers inside one process):
//
//   barrier A   every rank's J/K kernel has finished (release/acquire flags written straight into the peers' memory)
//   phase 1     rank r owns slice r of the stacked planes: it loads that slice of EVERY rank's accumulators over
//               NVLink (128-bit loads), adds them in rank order 0..W-1 and stores the sum into every rank's result
//               buffer (reduce-scatter + all-gather in one pass; each element crosses NVLink once per direction)
//   barrier B   all slices have landed everywhere
//   phase 2     local: J = J~ + J~^T, K = K~ + K~^T from the local result buffer into the caller's J/K
//
// Every element is summed by exactly one rank in a fixed order, so all ranks end with bit-identical J/K.
// Spins are bounded (COMM_TIMEOUT_NS).  A rank whose wait times out does NOT carry on with incomplete data: it skips
// its phase-1 stores, releases barrier B with COMM_FAIL_BIT set in the flag value (so that every peer learns of the
// failure in the same epoch), every rank that saw a failure fills its J/K with NaN, and the sticky status word (mapped
// host memory, sqcc_comm_status / sqcc_comm_reset_status) is raised on each of them.
#include "common.cuh"

namespace sqcc {

constexpr unsigned long long COMM_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
constexpr unsigned long long COMM_FAIL_BIT = 1ull << 63;   // flag value = epoch | COMM_FAIL_BIT: the sender gave up

struct CommRank {   // one rank's view of the group (device copy in sqcc_ctx::comm.d_view)
    const double *acc[COMM_MAXW];           // every rank's accumulators [3][N][ldd]
    double *res[COMM_MAXW];                 // every rank's result buffer [3][N][ldd]
    unsigned long long *flags[COMM_MAXW];   // every rank's flag block [2][COMM_MAXW]
    double *J, *K;                          // this rank's outputs
    unsigned int *gridbar;                  // this rank's grid barrier counter
    int *status;                            // mapped host word: != 0 after a barrier timeout
    unsigned long long *stamps;             // [8] globaltimer at the phase boundaries of the last launch (CTA 0)
    int rank;
};

__device__ __forceinline__ unsigned long long comm_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_peer2(const double *p)   // system-scope load: never served from a stale L1 line
{
    double2 r;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_peer2(double *p, double2 v)
{
    asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
// Returns false when the wait timed out or the sender flagged a failure for this epoch.
__device__ __forceinline__ bool wait_flag(const unsigned long long *p, unsigned long long epoch, int *status)
{
    const unsigned long long t0 = comm_now();
    for (;;) {
        const unsigned long long v = ld_acquire_sys(p);
        if ((v & ~COMM_FAIL_BIT) >= epoch) {
            if ((v & COMM_FAIL_BIT) && (v & ~COMM_FAIL_BIT) == epoch) {
                *status = 3;   // a peer reported a failed exchange
                return false;
            }
            return true;
        }
        __nanosleep(200);   // the GPUs run at their power cap: a waiting rank should not burn it polling
        if (comm_now() - t0 > COMM_TIMEOUT_NS) {
            *status = 1;
            __threadfence_system();
            return false;
        }
    }
}

// grid = (CTAs per rank, ranks emulated by this launch); cooperative launch (CTA 0 of a rank waits for its siblings).
// inv_scale != 0: the accumulators are 64-bit fixed point (deterministic mode): integer sums, converted in phase 2.
__global__ void __launch_bounds__(256) jk_reduce_kernel(const CommRank *views, int world, int nspin, int want_k, int n,
                                                        int ldd, unsigned long long epoch, unsigned int bar_target,
                                                        double inv_scale)
{
    const CommRank &cr = views[blockIdx.y];
    const int me = cr.rank;
    const int tid = threadIdx.x;
    const size_t nl = (size_t)n * ldd;
    const int nplanes = want_k ? 1 + nspin : 1;

    const bool stamp = blockIdx.x == 0 && tid == 0;
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;
    __syncthreads();
    if (stamp) cr.stamps[0] = comm_now();
    // ---- barrier A: my accumulators are complete (the J/K kernel precedes this launch on the stream)
    if (blockIdx.x == 0 && tid < world) {
        __threadfence_system();
        st_release_sys(cr.flags[tid] + me, epoch);
    }
    if (tid < world && !wait_flag(cr.flags[me] + tid, epoch, cr.status)) s_fail = 1;
    __syncthreads();
    const bool failed_a = s_fail != 0;
    if (stamp) cr.stamps[1] = comm_now();

    // ---- phase 1: reduce slice `me` of every rank's accumulators, broadcast the sums (skipped after a failed barrier:
    // the peers' accumulators may be incomplete, nothing derived from them is published)
    if (!failed_a) {
        const size_t total2 = (size_t)nplanes * nl / 2;   // 16-byte units (ldd is a multiple of 4)
        const size_t s0 = total2 * (size_t)me / world, s1 = total2 * (size_t)(me + 1) / world;
        for (size_t t = s0 + (size_t)blockIdx.x * blockDim.x + tid; t < s1; t += (size_t)gridDim.x * blockDim.x) {
            double2 sum = make_double2(0.0, 0.0);
            long long ix = 0, iy = 0;
            for (int g0 = 0; g0 < world; g0 += 8) {   // eight peer loads in flight, summed in rank order
                double2 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = g0 + u < world ? ld_peer2(cr.acc[g0 + u] + 2 * t) : make_double2(0.0, 0.0);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    sum.x += v[u].x;
                    sum.y += v[u].y;
                    ix += __double_as_longlong(v[u].x);   // (zero bits are integer zero as well)
                    iy += __double_as_longlong(v[u].y);
                }
            }
            if (inv_scale != 0.0) sum = make_double2(__longlong_as_double(ix), __longlong_as_double(iy));
#pragma unroll 8
            for (int g = 0; g < world; ++g) st_peer2(cr.res[g] + 2 * t, sum);
        }
    }

    // ---- barrier B: every CTA of this rank has stored its sums; then every rank has
    __syncthreads();
    if (stamp) cr.stamps[2] = comm_now();
    if (tid == 0) {
        __threadfence_system();
        atomicAdd(cr.gridbar, 1u);
    }
    if (blockIdx.x == 0) {
        if (tid == 0) {
            const unsigned long long t0 = comm_now();
            while ((int)(ld_acquire_gpu(cr.gridbar) - bar_target) < 0) {
                __nanosleep(100);
                if (comm_now() - t0 > COMM_TIMEOUT_NS) {
                    *cr.status = 2;
                    s_fail = 1;
                    break;
                }
            }
            __threadfence_system();
            cr.stamps[3] = comm_now();
        }
        __syncthreads();
        // release barrier B; a rank that gave up says so in the flag, so every peer poisons its result in this epoch
        if (tid < world) st_release_sys(cr.flags[tid] + COMM_MAXW + me, s_fail ? (epoch | COMM_FAIL_BIT) : epoch);
    }
    if (tid < world && !wait_flag(cr.flags[me] + COMM_MAXW + tid, epoch, cr.status)) s_fail = 1;
    __syncthreads();
    const bool failed = s_fail != 0;
    if (stamp) cr.stamps[4] = comm_now();

    // ---- phase 2 (local): symmetrise into the caller's J / K.  Four elements per thread with all loads issued before
    // the first store (the L2 round trips overlap instead of chaining)
    {
        const double *res = cr.res[me];
        const size_t nn = (size_t)n * n;
        const size_t stride = (size_t)gridDim.x * blockDim.x;
        for (size_t base = (size_t)blockIdx.x * blockDim.x + tid; base < nn; base += 4 * stride) {
            double jv[4], ka[4], kb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const size_t idx = base + u * stride;
                jv[u] = ka[u] = kb[u] = 0.0;
                if (idx < nn) {
                    const int pr = (int)(idx / n), q = (int)(idx % n);
                    const int hi = pr >= q ? pr : q, lo = pr >= q ? q : pr;
                    if (inv_scale != 0.0) {
                        const long long *ri = reinterpret_cast<const long long *>(res);
                        const long long jt = __ldcg(ri + (size_t)hi * ldd + lo);
                        jv[u] = (double)(pr == q ? 2 * jt : jt) * inv_scale;
                        if (want_k) {
                            ka[u] = (double)(__ldcg(ri + nl + (size_t)pr * ldd + q) + __ldcg(ri + nl + (size_t)q * ldd + pr)) * inv_scale;
                            if (nspin == 2)
                                kb[u] = (double)(__ldcg(ri + 2 * nl + (size_t)pr * ldd + q) + __ldcg(ri + 2 * nl + (size_t)q * ldd + pr)) * inv_scale;
                        }
                    } else {
                        const double jt = __ldcg(res + (size_t)hi * ldd + lo);
                        jv[u] = pr == q ? 2.0 * jt : jt;
                        if (want_k) {
                            ka[u] = __ldcg(res + nl + (size_t)pr * ldd + q) + __ldcg(res + nl + (size_t)q * ldd + pr);
                            if (nspin == 2)
                                kb[u] = __ldcg(res + 2 * nl + (size_t)pr * ldd + q) + __ldcg(res + 2 * nl + (size_t)q * ldd + pr);
                        }
                    }
                }
            }
            const double poison = __longlong_as_double(0x7ff8000000000000ll);   // NaN: a failed exchange never looks valid
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const size_t idx = base + u * stride;
                if (idx < nn) {
                    cr.J[idx] = failed ? poison : jv[u];
                    if (want_k) {
                        cr.K[idx] = failed ? poison : ka[u];
                        if (nspin == 2) cr.K[nn + idx] = failed ? poison : kb[u];
                    }
                }
            }
        }
    }
    if (stamp) cr.stamps[5] = comm_now();
}

// ---------------------------------------------------------------- host side
static size_t sym_doubles(const sqcc_ctx *ctx) { return 6 * (size_t)ctx->n * ctx->ldd + 2 * COMM_MAXW; }

static void comm_close_peers(sqcc_ctx *ctx)
{
    Comm &c = ctx->comm;
    if (c.ipc)
        for (int g = 0; g < c.world; ++g)
            if (g != c.rank && c.peer_base[g]) cudaIpcCloseMemHandle(c.peer_base[g]);
    for (int g = 0; g < COMM_MAXW; ++g) c.peer_base[g] = nullptr;
    c.view_valid = false;
    c.attached = false;
    c.enabled = true;   // a toggle set for an old mapping must not outlive it (sqcc_comm_set_enabled)
    c.ipc = false;
    c.world = 1;
    c.rank = 0;
}

// (Re)allocate the symmetric region [acc 3 nl][res 3 nl][flags]; called whenever an ERI shard is attached.
int comm_alloc(sqcc_ctx *ctx)
{
    Comm &c = ctx->comm;
    comm_close_peers(ctx);   // a new region invalidates every mapping of the old one
    if (c.d_sym) cudaFree(c.d_sym);
    c.d_sym = nullptr;
    const size_t nd = sym_doubles(ctx);
    SQCC_CUDA(cudaMalloc(&c.d_sym, nd * sizeof(double)));
    SQCC_CUDA(cudaMemsetAsync(c.d_sym, 0, nd * sizeof(double), ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->d_acc = c.d_sym;
    c.epoch = 0;
    if (!c.d_gridbar) {
        SQCC_CUDA(cudaMalloc(&c.d_gridbar, sizeof(unsigned int)));
        SQCC_CUDA(cudaMalloc(&c.d_stamps, sizeof(unsigned long long) * 8));
        SQCC_CUDA(cudaMemset(c.d_stamps, 0, sizeof(unsigned long long) * 8));
        SQCC_CUDA(cudaMalloc(&c.d_view, sizeof(CommRank) * COMM_MAXW));
        SQCC_CUDA(cudaHostAlloc(&c.h_status, sizeof(int), cudaHostAllocMapped));
        *c.h_status = 0;
        SQCC_CUDA(cudaHostGetDevicePointer(&c.d_status, c.h_status, 0));
    }
    SQCC_CUDA(cudaMemset(c.d_gridbar, 0, sizeof(unsigned int)));
    c.bar_count = 0;
    return SQCC_OK;
}

void comm_free(sqcc_ctx *ctx)
{
    Comm &c = ctx->comm;
    comm_close_peers(ctx);
    cudaFree(c.d_sym);
    cudaFree(c.d_gridbar);
    cudaFree(c.d_stamps);
    cudaFree(c.d_view);
    if (c.h_status) cudaFreeHost(c.h_status);
    c = Comm();
    ctx->d_acc = nullptr;
}

static void fill_view(const sqcc_ctx *ctx, CommRank &v, double *J, double *K)
{
    const Comm &c = ctx->comm;
    const size_t nl = (size_t)ctx->n * ctx->ldd;
    for (int g = 0; g < COMM_MAXW; ++g) {
        double *base = g < c.world ? static_cast<double *>(c.peer_base[g]) : nullptr;
        v.acc[g] = base;
        v.res[g] = base ? base + 3 * nl : nullptr;
        v.flags[g] = base ? reinterpret_cast<unsigned long long *>(base + 6 * nl) : nullptr;
    }
    v.J = J;
    v.K = K;
    v.gridbar = c.d_gridbar;
    v.status = c.d_status;
    v.stamps = c.d_stamps;
    v.rank = c.rank;
}

static int launch_reduce(sqcc_ctx *ctx, const CommRank *d_views, int ranks_here, int gx, int nspin, int want_k,
                         unsigned long long epoch, unsigned int bar_target)
{
    int world = ctx->comm.world, n = ctx->n, ldd = ctx->ldd;
    double inv_scale = ctx->fx_scale != 0.0 ? 1.0 / ctx->fx_scale : 0.0;
    void *args[] = {(void *)&d_views, &world, &nspin, &want_k, &n, &ldd, &epoch, &bar_target, &inv_scale};
    SQCC_CUDA(cudaLaunchCooperativeKernel((const void *)jk_reduce_kernel, dim3(gx, ranks_here), dim3(256), args, 0,
                                          ctx->stream));
    ctx->launches++;
    return SQCC_OK;
}

// One rank's exchange + finalize (ranks on distinct devices: every rank launches its own kernel).
int jk_reduce(sqcc_ctx *ctx, int nspin, int want_k, double *d_J, double *d_K)
{
    Comm &c = ctx->comm;
    SQCC_REQUIRE(c.attached && c.world > 1, "no peer group attached (sqcc_comm_attach)");
    SQCC_REQUIRE(!c.same_device, "ranks sharing one device must be reduced with sqcc_jk_reduce_local");
    if (!c.view_valid || c.view_J != d_J || c.view_K != d_K) {   // the device copy of the view changes only with J/K
        CommRank v;
        fill_view(ctx, v, d_J, d_K);
        SQCC_CUDA(cudaMemcpyAsync(c.d_view, &v, sizeof(CommRank), cudaMemcpyHostToDevice, ctx->stream));
        c.view_valid = true;
        c.view_J = d_J;
        c.view_K = d_K;
    }
    const int gx = ctx->sm_count < 148 ? ctx->sm_count : 148;   // one small CTA per SM: the exchange is latency-bound
    c.epoch += 1;
    c.bar_count += (unsigned int)gx;
    return launch_reduce(ctx, c.d_view, 1, gx, nspin, want_k, c.epoch, c.bar_count);
}

}  // namespace sqcc

using namespace sqcc;

extern "C" {

int sqcc_comm_export(sqcc_ctx *ctx, void *handle)
{
    SQCC_REQUIRE(ctx && handle && ctx->comm.d_sym, "no ERI attached (the exchange region is sized by sqcc_eri_attach)");
    static_assert(sizeof(cudaIpcMemHandle_t) == SQCC_COMM_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    SQCC_CUDA(cudaSetDevice(ctx->device));
    SQCC_CUDA(cudaIpcGetMemHandle(&h, ctx->comm.d_sym));
    memcpy(handle, &h, sizeof(h));
    return SQCC_OK;
}

int sqcc_comm_attach(sqcc_ctx *ctx, int world, int rank, const void *handles)
{
    SQCC_REQUIRE(ctx && handles && ctx->comm.d_sym, "no ERI attached / null handles");
    SQCC_REQUIRE(world >= 1 && world <= COMM_MAXW && rank >= 0 && rank < world, "bad world / rank");
    SQCC_CUDA(cudaSetDevice(ctx->device));
    Comm &c = ctx->comm;
    comm_close_peers(ctx);
    c.world = world;
    c.rank = rank;
    c.ipc = true;
    c.same_device = false;
    for (int g = 0; g < world; ++g) {
        if (g == rank) {
            c.peer_base[g] = c.d_sym;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)g * SQCC_COMM_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            comm_close_peers(ctx);
            set_error(std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(g) + "): " + cudaGetErrorString(e));
            return SQCC_ERR_CUDA;
        }
        c.peer_base[g] = p;
    }
    c.attached = true;
    return SQCC_OK;
}

int sqcc_comm_attach_local(sqcc_ctx **ctxs, int world)
{
    SQCC_REQUIRE(ctxs && world >= 1 && world <= COMM_MAXW, "bad group");
    bool same = true;
    for (int g = 0; g < world; ++g) {
        SQCC_REQUIRE(ctxs[g] && ctxs[g]->comm.d_sym, "every context needs an ERI shard attached");
        SQCC_REQUIRE(ctxs[g]->n == ctxs[0]->n, "contexts disagree on N");
        if (ctxs[g]->device != ctxs[0]->device) same = false;
    }
    for (int g = 0; g < world; ++g) {
        sqcc_ctx *ctx = ctxs[g];
        comm_close_peers(ctx);
        SQCC_CUDA(cudaSetDevice(ctx->device));
        for (int h = 0; h < world; ++h) {
            if (ctxs[h]->device != ctx->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[h]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) {
                    cudaGetLastError();
                } else if (e != cudaSuccess) {
                    set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                    return SQCC_ERR_CUDA;
                }
            }
            ctx->comm.peer_base[h] = ctxs[h]->comm.d_sym;
        }
        ctx->comm.world = world;
        ctx->comm.rank = g;
        ctx->comm.same_device = same;
        ctx->comm.attached = true;
    }
    return SQCC_OK;
}

int sqcc_comm_detach(sqcc_ctx *ctx)
{
    SQCC_REQUIRE(ctx, "null ctx");
    comm_close_peers(ctx);
    return SQCC_OK;
}

int sqcc_comm_stamps(sqcc_ctx *ctx, uint64_t *out8)
{
    SQCC_REQUIRE(ctx && out8 && ctx->comm.d_stamps, "no exchange region");
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    SQCC_CUDA(cudaMemcpy(out8, ctx->comm.d_stamps, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost));
    return SQCC_OK;
}

int sqcc_comm_set_enabled(sqcc_ctx *ctx, int enabled)
{
    SQCC_REQUIRE(ctx, "null ctx");
    ctx->comm.enabled = enabled != 0;
    return SQCC_OK;
}

int sqcc_comm_status(sqcc_ctx *ctx)
{
    SQCC_REQUIRE(ctx, "null ctx");
    if (ctx->comm.h_status && *ctx->comm.h_status != 0) {
        set_error("peer barrier timed out in the fused J/K exchange (a rank did not arrive)");
        return SQCC_ERR_CUDA;
    }
    return ctx->comm.attached ? ctx->comm.world : 0;
}

int sqcc_comm_reset_status(sqcc_ctx *ctx)
{
    SQCC_REQUIRE(ctx, "null ctx");
    if (ctx->comm.h_status) *ctx->comm.h_status = 0;
    return SQCC_OK;
}

int sqcc_jk_partial(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k)
{
    SQCC_REQUIRE(ctx, "null ctx");
    return jk_partial(ctx, d_D, nspin, want_k);
}

// All ranks of an in-process group that shares ONE device, in a single launch (blockIdx.y = rank): kernels that wait
// for one another must not be separate launches on one GPU.  Used by the single-GPU tests of the exchange.
int sqcc_jk_reduce_local(sqcc_ctx **ctxs, int world, int nspin, int want_k, double **d_J, double **d_K)
{
    SQCC_REQUIRE(ctxs && d_J && world >= 1 && world <= COMM_MAXW, "bad group");
    SQCC_REQUIRE(nspin == 1 || nspin == 2, "nspin must be 1 or 2");
    std::vector<CommRank> views(world);
    const int gx = 8;
    for (int g = 0; g < world; ++g) {
        sqcc_ctx *ctx = ctxs[g];
        SQCC_REQUIRE(ctx && ctx->comm.attached && ctx->comm.same_device && ctx->comm.world == world && ctx->comm.rank == g,
                     "contexts are not a same-device local group (sqcc_comm_attach_local)");
        SQCC_REQUIRE(d_J[g] && (!want_k || (d_K && d_K[g])), "null output");
        SQCC_REQUIRE(ctx->stream == ctxs[0]->stream, "a same-device group must share one stream");
        fill_view(ctx, views[g], d_J[g], want_k ? d_K[g] : nullptr);
        ctx->comm.epoch += 1;
        ctx->comm.bar_count += (unsigned int)gx;
        SQCC_REQUIRE(ctx->comm.epoch == ctxs[0]->comm.epoch && ctx->comm.bar_count == ctxs[0]->comm.bar_count,
                     "group members are out of step");
    }
    sqcc_ctx *c0 = ctxs[0];
    SQCC_CUDA(cudaSetDevice(c0->device));
    c0->comm.view_valid = false;
    SQCC_CUDA(cudaMemcpyAsync(c0->comm.d_view, views.data(), sizeof(CommRank) * world, cudaMemcpyHostToDevice, c0->stream));
    return launch_reduce(c0, c0->comm.d_view, world, gx, nspin, want_k, c0->comm.epoch, c0->comm.bar_count);
}

}  // extern "C"
