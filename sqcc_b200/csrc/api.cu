// C ABI of libsqcc_b200.so (see include/sqcc_b200.h).  Thin: argument checks, staging, dispatch.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace sqcc {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }

int ensure_io(sqcc_ctx *ctx, size_t bytes)
{
    if (ctx->io_bytes >= bytes) return SQCC_OK;
    if (ctx->d_io) cudaFree(ctx->d_io);
    ctx->d_io = nullptr;
    ctx->io_bytes = 0;
    SQCC_CUDA(cudaMalloc(&ctx->d_io, bytes));
    ctx->io_bytes = bytes;
    return SQCC_OK;
}

int ensure_pin(sqcc_ctx *ctx, size_t bytes)
{
    if (ctx->pin_bytes >= bytes) return SQCC_OK;
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr;
    ctx->pin_bytes = 0;
    SQCC_CUDA(cudaMallocHost(&ctx->h_pin, bytes));
    ctx->pin_bytes = bytes;
    return SQCC_OK;
}

int timer_start(sqcc_ctx *ctx, int slot)
{
    Timer &t = ctx->timers[slot];
    if (!t.a[t.next]) {
        SQCC_CUDA(cudaEventCreate(&t.a[t.next]));
        SQCC_CUDA(cudaEventCreate(&t.b[t.next]));
    }
    SQCC_CUDA(cudaEventRecord(t.a[t.next], ctx->stream));
    return SQCC_OK;
}

int timer_stop(sqcc_ctx *ctx, int slot)
{
    Timer &t = ctx->timers[slot];
    SQCC_CUDA(cudaEventRecord(t.b[t.next], ctx->stream));
    t.used = true;
    t.next = (t.next + 1) % TIMER_RING;
    if (t.count < TIMER_RING) ++t.count;
    return SQCC_OK;
}

// Tile index at which shard g of `world` starts: equal-COST cut with 4 x 4 tile granularity.  The J/K kernel spends the
// same time on every piece (one loop iteration: the 12 x 16 DFMAs run on the stored zeros of a narrow piece as well),
// so the cost of a tile is its number of pieces, sum over chunks c of (imax + 1 - 64 c), plus SQCC_SHARD_ALPHA (default
// 1) iterations' worth of per-sweep overhead for every 16 sub-rows (row-block prologue / epilogue).
static double shard_alpha()
{
    if (const char *e = getenv("SQCC_SHARD_ALPHA")) return atof(e);
    return 1.0;
}

static double tile_cost(int imax, double alpha)
{
    double pieces = 0.0, sweeps = 0.0;
    for (int c = 0; LAY_CHUNK * c <= imax; ++c) {
        pieces += imax + 1 - LAY_CHUNK * c;
        sweeps += (imax + 1 - LAY_CHUNK * c + JK_KN - 1) / JK_KN;
    }
    return pieces + alpha * sweeps;
}

// `weights` (optional, [world]): relative speed of the ranks -- rank r gets the share weights[r] / sum(weights) of the
// streaming cost instead of 1 / world (GPUs of one box settle at different clocks under their power caps).
static int64_t tile_cut(int n, int world, int g, const double *weights = nullptr)
{
    const int64_t nt = tile_count(n);
    if (g <= 0) return 0;
    if (g >= world) return nt;
    // cumulative cost by tile (tile order: i-block ascending, j-block ascending)
    std::vector<double> cum;
    cum.reserve(nt + 1);
    cum.push_back(0.0);
    const double alpha = shard_alpha();
    for (int ib = 0; ib < lay_nblocks(n); ++ib) {
        const int imax = lay_ihi(n, ib) - 1;
        const double cost = tile_cost(imax, alpha);
        for (int jb = 0; jb < lay_njb(imax); ++jb) cum.push_back(cum.back() + cost);
    }
    double frac = (double)g / (double)world;
    if (weights) {
        double tot = 0.0, part = 0.0;
        for (int r = 0; r < world; ++r) tot += weights[r];
        for (int r = 0; r < g; ++r) part += weights[r];
        frac = part / tot;
    }
    const double target = cum.back() * frac;
    int64_t lo = 0, hi = nt;
    while (lo < hi) {  // first tile whose start cost >= target
        int64_t mid = (lo + hi) / 2;
        if (cum[mid] < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static int attach_common(sqcc_ctx *ctx, double *d_native, int n)
{
    SQCC_REQUIRE((reinterpret_cast<uintptr_t>(d_native) & 15) == 0, "ERI buffer must be 16-byte aligned");
    SQCC_CUDA(cudaSetDevice(ctx->device));
    ctx->d_eri = d_native;
    ctx->n = n;
    ctx->ldd = (n + 3) & ~3;
    SQCC_TRY(eri_build_tables(ctx));
    SQCC_TRY(jk_build_schedule(ctx));
    return SQCC_OK;
}

}  // namespace sqcc

using namespace sqcc;

extern "C" {

int sqcc_version(void) { return 100; }
const char *sqcc_last_error(void) { return g_error.c_str(); }

int sqcc_device_count(int *count)
{
    SQCC_REQUIRE(count, "null count");
    SQCC_CUDA(cudaGetDeviceCount(count));
    return SQCC_OK;
}

int sqcc_ctx_create(int device, sqcc_ctx **out)
{
    SQCC_REQUIRE(out, "null out");
    SQCC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SQCC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("libsqcc_b200 is built for sm_100a only; device is ") + prop.name);
        return SQCC_ERR_UNSUPPORTED;
    }
    sqcc_ctx *ctx = new sqcc_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return SQCC_OK;
}

int sqcc_ctx_destroy(sqcc_ctx *ctx)
{
    if (!ctx) return SQCC_OK;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_ibase);
    for (auto &sc : ctx->sched) {
        cudaFree(sc.d_rowblocks);
        cudaFree(sc.d_tasks);
    }
    cudaFree(ctx->d_dwork);
    scf_free(ctx);
    comm_free(ctx);
    cudaFree(ctx->d_counter);
    cudaFree(ctx->d_io);
    cudaFree(ctx->d_gridwork);
    cudaFree(ctx->d_mp2work);
    cudaFree(ctx->d_mp2out);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    for (auto &t : ctx->timers)
        for (int e = 0; e < TIMER_RING; ++e) {
            if (t.a[e]) cudaEventDestroy(t.a[e]);
            if (t.b[e]) cudaEventDestroy(t.b[e]);
        }
    delete ctx;
    return SQCC_OK;
}

int sqcc_ctx_set_stream(sqcc_ctx *ctx, void *cuda_stream)
{
    SQCC_REQUIRE(ctx, "null ctx");
    ctx->stream = (cudaStream_t)cuda_stream;
    return SQCC_OK;
}

int sqcc_ctx_synchronize(sqcc_ctx *ctx)
{
    SQCC_REQUIRE(ctx, "null ctx");
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    return SQCC_OK;
}

int64_t sqcc_npair(int n) { return (int64_t)n * (n + 1) / 2; }
int64_t sqcc_nquad(int n)
{
    int64_t p = sqcc_npair(n);
    return p * (p + 1) / 2;
}

int64_t sqcc_tile_count(int n) { return n > 0 ? tile_count(n) : -1; }

int sqcc_shard_tiles(int n, int world, int rank, int64_t *t_begin, int64_t *t_end)
{
    SQCC_REQUIRE(n > 0 && world > 0 && rank >= 0 && rank < world && t_begin && t_end, "bad shard arguments");
    int64_t a = tile_cut(n, world, rank), b = tile_cut(n, world, rank + 1);
    if (b < a) b = a;
    *t_begin = a;
    *t_end = b;
    return SQCC_OK;
}

int sqcc_shard_tiles_weighted(int n, int world, int rank, const double *weights, int64_t *t_begin, int64_t *t_end)
{
    SQCC_REQUIRE(n > 0 && world > 0 && rank >= 0 && rank < world && weights && t_begin && t_end, "bad shard arguments");
    for (int r = 0; r < world; ++r) SQCC_REQUIRE(weights[r] > 0.0 && isfinite(weights[r]), "weights must be positive");
    int64_t a = tile_cut(n, world, rank, weights), b = tile_cut(n, world, rank + 1, weights);
    if (b < a) b = a;
    *t_begin = a;
    *t_end = b;
    return SQCC_OK;
}

int sqcc_tile_rows(int n, int64_t t_begin, int64_t t_end, int *jlo, int *jhi)
{
    SQCC_REQUIRE(n > 0 && t_begin >= 0 && t_end >= t_begin && t_end <= tile_count(n) && jlo && jhi, "bad tile range");
    std::vector<int> a, b;
    tile_intervals(n, t_begin, t_end, a, b);
    for (int i = 0; i < n; ++i) {
        jlo[i] = a[i];
        jhi[i] = b[i];
    }
    return SQCC_OK;
}

int64_t sqcc_native_len_tiles(int n, int64_t t_begin, int64_t t_end)
{
    if (n <= 0 || t_begin < 0 || t_end < t_begin || t_end > tile_count(n)) return -1;
    return tile_offset(n, t_end) - tile_offset(n, t_begin);
}

int64_t sqcc_native_offset(int n, int64_t t_begin, int i, int j, int k, int l)
{
    if (n <= 0 || i < 0 || i >= n || j < 0 || j > i || k < 0 || l < 0 || l > k) return -1;
    if (pair_index(k, l) > pair_index(i, j)) return -1;
    std::vector<int64_t> ibase, tfirst;
    tile_tables(n, ibase, tfirst);
    if (tfirst[lay_iblock(n, i)] + (j >> 2) < t_begin) return -1;
    return lay_offset(ibase.data(), n, i, j, k, l) - tile_offset(n, t_begin);
}

int sqcc_eri_attach_tiles(sqcc_ctx *ctx, double *d_native, int n, int64_t t_begin, int64_t t_end)
{
    SQCC_REQUIRE(ctx && d_native, "null ctx / buffer");
    SQCC_REQUIRE(n > 0 && n <= 4096, "N out of range (1..4096)");
    SQCC_REQUIRE(t_begin >= 0 && t_end >= t_begin && t_end <= tile_count(n), "bad tile range");
    ctx->t_begin = t_begin;
    ctx->t_end = t_end;
    return attach_common(ctx, d_native, n);
}

int sqcc_eri_synth(sqcc_ctx *ctx, uint64_t seed)
{
    SQCC_REQUIRE(ctx && ctx->d_eri, "no ERI attached");
    ctx->eri_absmax = -1.0;
    return eri_synth(ctx, seed);   // writes every cell of the shard (zeros outside the canonical range)
}

int sqcc_eri_pack_full_slab(sqcc_ctx *ctx, const double *d_slab, int p0, int np)
{
    SQCC_REQUIRE(ctx && ctx->d_eri && d_slab, "no ERI attached / null slab");
    SQCC_REQUIRE(p0 >= 0 && np > 0 && p0 + np <= ctx->n, "slab range out of bounds");
    return eri_pack_full_slab(ctx, d_slab, p0, np);
}

int sqcc_eri_pack_full_host(sqcc_ctx *ctx, const double *h_eri_full)
{
    SQCC_REQUIRE(ctx && ctx->d_eri && h_eri_full, "no ERI attached / null tensor");
    const int n = ctx->n;
    const size_t slab = (size_t)n * n * n;  // doubles per p
    // slabs of up to ~256 MB, double-buffered through pinned memory
    int np = (int)((size_t)(256u << 20) / (slab * sizeof(double)));
    if (np < 1) np = 1;
    if (np > n) np = n;
    SQCC_TRY(ensure_pin(ctx, 2 * slab * np * sizeof(double)));
    SQCC_TRY(ensure_io(ctx, 2 * slab * np * sizeof(double)));
    cudaEvent_t done[2];
    SQCC_CUDA(cudaEventCreate(&done[0]));
    SQCC_CUDA(cudaEventCreate(&done[1]));
    int buf = 0, issued = 0;
    for (int p0 = ctx->i_begin; p0 < ctx->i_end; p0 += np, buf ^= 1, ++issued) {   // slabs without held rows pack nothing
        int cnt = p0 + np <= ctx->i_end ? np : ctx->i_end - p0;
        double *hp = ctx->h_pin + (size_t)buf * slab * np;
        double *dp = ctx->d_io + (size_t)buf * slab * np;
        if (issued >= 2) SQCC_CUDA(cudaEventSynchronize(done[buf]));
        memcpy(hp, h_eri_full + (size_t)p0 * slab, (size_t)cnt * slab * sizeof(double));
        SQCC_CUDA(cudaMemcpyAsync(dp, hp, (size_t)cnt * slab * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        SQCC_TRY(eri_pack_full_slab(ctx, dp, p0, cnt));
        SQCC_CUDA(cudaEventRecord(done[buf], ctx->stream));
    }
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaEventDestroy(done[0]);
    cudaEventDestroy(done[1]);
    return SQCC_OK;
}

int sqcc_eri_pack_canonical(sqcc_ctx *ctx, const double *d_rows, int64_t ij0, int64_t nrows)
{
    SQCC_REQUIRE(ctx && ctx->d_eri && d_rows, "no ERI attached / null rows");
    SQCC_REQUIRE(ij0 >= 0 && nrows > 0 && ij0 + nrows <= sqcc_npair(ctx->n), "row range out of bounds");
    return eri_pack_canonical(ctx, d_rows, ij0, nrows);
}

int sqcc_eri_clear(sqcc_ctx *ctx)
{
    SQCC_REQUIRE(ctx && ctx->d_eri, "no ERI attached");
    SQCC_CUDA(cudaMemsetAsync(ctx->d_eri, 0, sizeof(double) * ctx->native_len, ctx->stream));
    return SQCC_OK;
}

int sqcc_eri_pack_blocks(sqcc_ctx *ctx, const double *d_values, const int64_t *d_desc, int nblk)
{
    SQCC_REQUIRE(ctx && ctx->d_eri && d_values && d_desc && nblk >= 0, "no ERI attached / null block buffers");
    return eri_pack_blocks(ctx, d_values, d_desc, nblk);
}

int sqcc_eri_unpack_full(sqcc_ctx *ctx, double *d_eri_full)
{
    SQCC_REQUIRE(ctx && d_eri_full, "null ctx / output");
    return eri_unpack_full(ctx, d_eri_full);
}

int sqcc_jk_set_variant(sqcc_ctx *ctx, int variant)
{
    SQCC_REQUIRE(ctx && variant >= 0 && (variant <= 1 || (variant >= 16 && variant <= 23)),
                 "variant must be 0 (streaming kernel), 1 (simple atomics kernel) or 16+cfg (streaming kernel, tuning cfg)");
    ctx->jk_variant = variant < 16 ? variant : 0;
    ctx->jk_cfg = variant >= 16 ? variant - 16 : 0;
    return SQCC_OK;
}

int sqcc_jk_set_deterministic(sqcc_ctx *ctx, int on)
{
    SQCC_REQUIRE(ctx, "null ctx");
    ctx->jk_fixed = on != 0;
    return SQCC_OK;
}

int sqcc_eri_absmax(sqcc_ctx *ctx, double *absmax, int set)
{
    SQCC_REQUIRE(ctx && absmax, "null ctx / value");
    if (set) {
        SQCC_REQUIRE(*absmax >= 0.0, "negative bound");
        ctx->eri_absmax = *absmax;
    } else {
        *absmax = ctx->eri_absmax;
    }
    return SQCC_OK;
}

int sqcc_jk_build(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k, double *d_J, double *d_K)
{
    SQCC_REQUIRE(ctx, "null ctx");
    return jk_build(ctx, d_D, nspin, want_k, d_J, d_K);
}

// true when `p` is page-locked host memory the device can copy from / to directly (cudaHostAlloc, cudaHostRegister,
// torch pin_memory): such buffers skip the library's pinned staging copy
static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int sqcc_jk_build_host(sqcc_ctx *ctx, const double *h_D, int nspin, int want_k, double *h_J, double *h_K)
{
    SQCC_REQUIRE(ctx && h_D && h_J && (!want_k || h_K), "null ctx / buffers");
    SQCC_REQUIRE(nspin == 1 || nspin == 2, "nspin must be 1 or 2");
    const size_t nn = (size_t)ctx->n * ctx->n;
    // staging layout: [D nspin*nn][J nn][K nspin*nn]
    const size_t total = (2 * (size_t)nspin + 1) * nn;
    SQCC_TRY(ensure_io(ctx, total * sizeof(double)));
    SQCC_TRY(ensure_pin(ctx, total * sizeof(double)));
    double *dD = ctx->d_io, *dJ = dD + nspin * nn, *dK = dJ + nn;
    const bool pin_in = is_pinned_host(h_D), pin_j = is_pinned_host(h_J), pin_k = want_k && is_pinned_host(h_K);
    const double *src = h_D;
    if (!pin_in) {
        memcpy(ctx->h_pin, h_D, nspin * nn * sizeof(double));
        src = ctx->h_pin;
    }
    SQCC_CUDA(cudaMemcpyAsync(dD, src, nspin * nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SQCC_TRY(jk_build(ctx, dD, nspin, want_k, dJ, dK));
    double *stage_j = ctx->h_pin + nspin * nn, *stage_k = stage_j + nn;
    SQCC_CUDA(cudaMemcpyAsync(pin_j ? h_J : stage_j, dJ, nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (want_k)
        SQCC_CUDA(cudaMemcpyAsync(pin_k ? h_K : stage_k, dK, nspin * nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!pin_j) memcpy(h_J, stage_j, nn * sizeof(double));
    if (want_k && !pin_k) memcpy(h_K, stage_k, nspin * nn * sizeof(double));
    return SQCC_OK;
}

int sqcc_grid_attach(sqcc_ctx *ctx, const double *d_phi, const double *d_w, int64_t G, int n)
{
    SQCC_REQUIRE(ctx && d_phi && d_w && G > 0 && n > 0, "bad grid arguments");
    ctx->d_phi = d_phi;
    ctx->d_w = d_w;
    ctx->G = G;
    ctx->grid_n = n;
    return SQCC_OK;
}

int sqcc_lda_vxc(sqcc_ctx *ctx, const double *d_D, int nspin, double *d_Vxc, double *d_Exc, double *d_rho)
{
    SQCC_REQUIRE(ctx && ctx->d_phi, "no grid attached (sqcc_grid_attach)");
    SQCC_REQUIRE(d_D && d_Vxc && d_Exc && (nspin == 1 || nspin == 2), "bad lda_vxc arguments");
    return lda_vxc(ctx, d_D, nspin, d_Vxc, d_Exc, d_rho);
}

int sqcc_lda_vxc_host(sqcc_ctx *ctx, const double *h_D, int nspin, double *h_Vxc, double *h_Exc, double *h_rho)
{
    SQCC_REQUIRE(ctx && ctx->d_phi, "no grid attached (sqcc_grid_attach)");
    SQCC_REQUIRE(h_D && h_Vxc && h_Exc && (nspin == 1 || nspin == 2), "bad lda_vxc arguments");
    const size_t nn = (size_t)ctx->grid_n * ctx->grid_n;
    const size_t g = (size_t)ctx->G;
    // staging: [D nspin*nn][Vxc nspin*nn][Exc 1 (+1 pad)][rho nspin*G]
    const size_t small = 2 * nspin * nn + 2;
    SQCC_TRY(ensure_io(ctx, (small + nspin * g) * sizeof(double)));
    SQCC_TRY(ensure_pin(ctx, (small + (h_rho ? nspin * g : 0)) * sizeof(double)));
    double *dD = ctx->d_io, *dV = dD + nspin * nn, *dE = dV + nspin * nn, *dR = dE + 2;
    memcpy(ctx->h_pin, h_D, nspin * nn * sizeof(double));
    SQCC_CUDA(cudaMemcpyAsync(dD, ctx->h_pin, nspin * nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SQCC_TRY(lda_vxc(ctx, dD, nspin, dV, dE, dR));
    size_t out = nspin * nn + 2 + (h_rho ? nspin * g : 0);
    SQCC_CUDA(cudaMemcpyAsync(ctx->h_pin + nspin * nn, dV, out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(h_Vxc, ctx->h_pin + nspin * nn, nspin * nn * sizeof(double));
    *h_Exc = ctx->h_pin[2 * nspin * nn];
    if (h_rho) memcpy(h_rho, ctx->h_pin + small, nspin * g * sizeof(double));
    return SQCC_OK;
}

int sqcc_grid_quadrature(sqcc_ctx *ctx, const double *d_v, double *d_V)
{
    SQCC_REQUIRE(ctx && ctx->d_phi && d_v && d_V, "no grid attached / null pointers");
    return grid_quadrature(ctx, d_v, d_V);
}

int sqcc_grid_rho(sqcc_ctx *ctx, const double *d_D, double *d_rho)
{
    SQCC_REQUIRE(ctx && ctx->d_phi && d_D && d_rho, "no grid attached / null pointers");
    return grid_rho(ctx, d_D, d_rho);
}

int sqcc_ao2mo(sqcc_ctx *ctx, const double *d_C1, int n1, const double *d_C2, int n2, const double *d_C3, int n3,
               const double *d_C4, int n4, double *d_out)
{
    SQCC_REQUIRE(ctx && ctx->d_eri && d_C1 && d_C2 && d_C3 && d_C4 && d_out, "no ERI attached / null pointers");
    SQCC_REQUIRE(n1 > 0 && n2 > 0 && n3 > 0 && n4 > 0, "bad MO block sizes");
    return ao2mo(ctx, d_C1, n1, d_C2, n2, d_C3, n3, d_C4, n4, d_out);
}

int sqcc_mp2(sqcc_ctx *ctx, const double *d_C, const double *d_eps, int n_occ, double *d_ecorr, double *d_iajb)
{
    SQCC_REQUIRE(ctx && ctx->d_eri && d_C && d_eps && d_ecorr, "no ERI attached / null pointers");
    SQCC_REQUIRE(n_occ > 0 && n_occ < ctx->n, "n_occ out of range");
    return mp2(ctx, d_C, d_eps, n_occ, d_ecorr, d_iajb);
}

int sqcc_mp2_host(sqcc_ctx *ctx, const double *h_C, const double *h_eps, int n_occ, double *h_ecorr, double *h_iajb)
{
    SQCC_REQUIRE(ctx && ctx->d_eri && h_C && h_eps && h_ecorr, "no ERI attached / null pointers");
    SQCC_REQUIRE(n_occ > 0 && n_occ < ctx->n, "n_occ out of range");
    const size_t n = ctx->n, nn = n * n, nv = n - n_occ;
    const size_t nout = h_iajb ? (size_t)n_occ * nv * n_occ * nv : 0;
    SQCC_TRY(ensure_io(ctx, (nn + n + 2 + nout) * sizeof(double)));
    SQCC_TRY(ensure_pin(ctx, (nn + n + 2 + nout) * sizeof(double)));
    double *dC = ctx->d_io, *dE = dC + nn, *dX = dE + n, *dO = dX + 2;
    memcpy(ctx->h_pin, h_C, nn * sizeof(double));
    memcpy(ctx->h_pin + nn, h_eps, n * sizeof(double));
    SQCC_CUDA(cudaMemcpyAsync(dC, ctx->h_pin, (nn + n) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SQCC_TRY(mp2(ctx, dC, dE, n_occ, dX, h_iajb ? dO : nullptr));
    SQCC_CUDA(cudaMemcpyAsync(ctx->h_pin + nn + n, dX, (2 + nout) * sizeof(double), cudaMemcpyDeviceToHost,
                              ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_ecorr = ctx->h_pin[nn + n];
    if (h_iajb) memcpy(h_iajb, ctx->h_pin + nn + n + 2, nout * sizeof(double));
    return SQCC_OK;
}

int sqcc_get_timings(sqcc_ctx *ctx, double *ms, int n)
{
    SQCC_REQUIRE(ctx && ms && n > 0, "bad timing arguments");
    int m = n < 8 ? n : 8;
    for (int s = 0; s < m; ++s) {
        ms[s] = -1.0;
        sqcc::Timer &t = ctx->timers[s];
        if (t.used && t.count > 0) {   // mean over the launches since the last query
            double sum = 0.0;
            for (int e = 0; e < t.count; ++e) {
                const int idx = (t.next - 1 - e + 2 * TIMER_RING) % TIMER_RING;
                SQCC_CUDA(cudaEventSynchronize(t.b[idx]));
                float f = 0.f;
                SQCC_CUDA(cudaEventElapsedTime(&f, t.a[idx], t.b[idx]));
                sum += f;
            }
            ms[s] = sum / t.count;
            t.count = 0;
        }
    }
    return m;
}

int64_t sqcc_launch_count(sqcc_ctx *ctx) { return ctx ? ctx->launches : -1; }

int sqcc_eri_bytes(sqcc_ctx *ctx, int64_t *native_bytes, int64_t *algorithmic_bytes)
{
    SQCC_REQUIRE(ctx && ctx->d_eri, "no ERI attached");
    if (native_bytes) *native_bytes = ctx->native_len * 8;
    if (algorithmic_bytes) *algorithmic_bytes = ctx->n_unique * 8;
    return SQCC_OK;
}

}  // extern "C"
