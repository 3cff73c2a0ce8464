// ERI ingest kernels for the native v2 packed layout (layout.cuh): shard geometry, synthetic fill, pack (full tensor
// slabs / canonical rows / dense blocks) and unpack.
// Consumes what interface_psi4.py:177-190 (ao_electron_repulsion_integral, full [N,N,N,N]) produces.
#include <math.h>

#include "common.cuh"

namespace sqcc {

// (i, j) from a pair index, exact for ij < 2^52.
__host__ __device__ __forceinline__ void pair_unrank(int64_t ij, int &i, int &j)
{
    int64_t t = (int64_t)((sqrt(8.0 * (double)ij + 1.0) - 1.0) * 0.5);
    while (t * (t + 1) / 2 > ij) --t;
    while ((t + 1) * (t + 2) / 2 <= ij) ++t;
    i = (int)t;
    j = (int)(ij - t * (t + 1) / 2);
}

void pair_unrank_host(int64_t ij, int &i, int &j) { pair_unrank(ij, i, j); }

// ---- tile geometry (host) ----------------------------------------------------------------------------------
int64_t tile_count(int n)
{
    int64_t t = 0;
    for (int ib = 0; ib < lay_nblocks(n); ++ib) t += lay_njb(lay_ihi(n, ib) - 1);
    return t;
}

// ibase[ib] = offset (doubles) of the first tile of i-block ib in the full tensor, tfirst[ib] = its tile index
void tile_tables(int n, std::vector<int64_t> &ibase, std::vector<int64_t> &tfirst)
{
    const int nb = lay_nblocks(n);
    ibase.assign(nb + 1, 0);
    tfirst.assign(nb + 1, 0);
    for (int ib = 0; ib < nb; ++ib) {
        const int imax = lay_ihi(n, ib) - 1;
        ibase[ib + 1] = ibase[ib] + (int64_t)lay_njb(imax) * lay_tile_len(imax);
        tfirst[ib + 1] = tfirst[ib] + lay_njb(imax);
    }
}

// offset (doubles) of tile t from the start of the full tensor (t == tile_count: the total length)
int64_t tile_offset(int n, int64_t t)
{
    std::vector<int64_t> ibase, tfirst;
    tile_tables(n, ibase, tfirst);
    const int nb = lay_nblocks(n);
    if (t >= tfirst[nb]) return ibase[nb];
    int ib = 0;
    while (tfirst[ib + 1] <= t) ++ib;
    return ibase[ib] + (t - tfirst[ib]) * lay_tile_len(lay_ihi(n, ib) - 1);
}

// rows held by the tile range, per i the interval [jlo, jhi) of j
int tile_intervals(int n, int64_t t_begin, int64_t t_end, std::vector<int> &jlo, std::vector<int> &jhi)
{
    jlo.assign(n, 0);
    jhi.assign(n, 0);
    int64_t t0 = 0;
    for (int ib = 0; ib < lay_nblocks(n); ++ib) {
        const int hi = lay_ihi(n, ib), lo = hi - 4 < 0 ? 0 : hi - 4;
        const int64_t cnt = lay_njb(hi - 1);
        const int64_t a = t_begin > t0 ? t_begin - t0 : 0, c = (t_end < t0 + cnt ? t_end : t0 + cnt) - t0;
        if (c > a) {
            for (int i = lo; i < hi; ++i) {
                jlo[i] = (int)(4 * a < i + 1 ? 4 * a : i + 1);
                jhi[i] = (int)(4 * c < i + 1 ? 4 * c : i + 1);
            }
        }
        t0 += cnt;
    }
    return SQCC_OK;
}

int eri_build_tables(sqcc_ctx *ctx)
{
    const int n = ctx->n, nb = lay_nblocks(n);
    tile_tables(n, ctx->h_ibase, ctx->h_tfirst);
    tile_intervals(n, ctx->t_begin, ctx->t_end, ctx->jlo, ctx->jhi);
    ctx->shard_base = tile_offset(n, ctx->t_begin);
    ctx->native_len = tile_offset(n, ctx->t_end) - ctx->shard_base;
    ctx->whole = ctx->t_begin == 0 && ctx->t_end == ctx->h_tfirst[nb];
    ctx->nrows = 0;
    ctx->n_unique = 0;
    for (int i = 0; i < n; ++i)
        if (ctx->jhi[i] > ctx->jlo[i]) {
            ctx->nrows += ctx->jhi[i] - ctx->jlo[i];
            for (int j = ctx->jlo[i]; j < ctx->jhi[i]; ++j) ctx->n_unique += pair_index(i, j) + 1;
        }
    // i range of the i-blocks the held tiles belong to (whole blocks: a fill must visit every row slot of a held tile)
    ctx->i_begin = ctx->i_end = 0;
    if (ctx->t_end > ctx->t_begin) {
        int ib0 = 0, ib1 = 0;
        while (ctx->h_tfirst[ib0 + 1] <= ctx->t_begin) ++ib0;
        while (ctx->h_tfirst[ib1 + 1] <= ctx->t_end - 1) ++ib1;
        ctx->i_begin = lay_ihi(n, ib0) - 4 < 0 ? 0 : lay_ihi(n, ib0) - 4;
        ctx->i_end = lay_ihi(n, ib1);
    }
    if (ctx->d_ibase) cudaFree(ctx->d_ibase);
    ctx->d_ibase = ctx->d_tfirst = nullptr;
    SQCC_CUDA(cudaMalloc(&ctx->d_ibase, sizeof(int64_t) * 2 * (nb + 1)));
    ctx->d_tfirst = ctx->d_ibase + (nb + 1);
    SQCC_CUDA(cudaMemcpyAsync(ctx->d_ibase, ctx->h_ibase.data(), sizeof(int64_t) * (nb + 1), cudaMemcpyHostToDevice, ctx->stream));
    SQCC_CUDA(cudaMemcpyAsync(ctx->d_tfirst, ctx->h_tfirst.data(), sizeof(int64_t) * (nb + 1), cudaMemcpyHostToDevice, ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    return SQCC_OK;
}

// ---- synthetic generator: bit-identical to oracle/synth.py and oracle/csrc/jk_oracle.c --------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double synth_value(uint64_t npair, uint64_t seed, int i, int j, int k, int l)
{
    uint64_t ij = (uint64_t)pair_index(i, j), kl = (uint64_t)pair_index(k, l);
    uint64_t z = splitmix64(seed + 0x9E3779B97F4A7C15ULL * (ij * npair + kl + 1ULL));
    double u = __dmul_rn((double)(z >> 11), 0x1.0p-53);
    double t = __dadd_rn(__dmul_rn(2.0, u), -1.0);
    double d = (double)((i - j) + (k - l));
    double s = __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(0.25, d)));
    return ij == kl ? __dmul_rn(s, __dadd_rn(fabs(t), 1.0)) : __dmul_rn(s, t);
}

enum { SRC_SYNTH = 0, SRC_FULL_SLAB = 1, SRC_CANONICAL = 2 };

// Writes the shard in storage order: blockIdx.x strides over the held tiles, blockIdx.y is the chunk column, the CTA
// walks the pieces k = 64 c .. imax of that column (coalesced stores; the source is read along l).  Cells outside the
// canonical range are written as 0.0, so a fill needs no memset.  Rows (i, .) outside [i_lo, i_hi) (slab / row-range
// sources deliver the tensor in parts) are left alone.
template <int SRC>
__global__ void __launch_bounds__(256) fill_tiles_kernel(double *__restrict__ P, const int64_t *__restrict__ ibase,
                                                         const int64_t *__restrict__ tfirst, int64_t t_lo, int64_t t_hi,
                                                         int64_t shard_base, int n, uint64_t seed,
                                                         const double *__restrict__ src, int i_lo, int i_hi,
                                                         int64_t ij_lo, int64_t ij_hi, int64_t canon_base)
{
    const uint64_t np_ = (uint64_t)n * (uint64_t)(n + 1) / 2;
    const int c = blockIdx.y;
    int ib = 0;
    for (int64_t t = t_lo + blockIdx.x; t < t_hi; t += gridDim.x) {
        while (tfirst[ib + 1] <= t) ++ib;
        const int ihi = lay_ihi(n, ib), imax = ihi - 1, jb = (int)(t - tfirst[ib]);
        if (LAY_CHUNK * c > imax) continue;
        double *tile = P + (ibase[ib] + (int64_t)jb * lay_tile_len(imax) - shard_base);
        for (int k = LAY_CHUNK * c; k <= imax; ++k) {
            const int w = lay_w(k - LAY_CHUNK * c);
            double *piece = tile + lay_piece_offset(imax, c, k);
            for (int e = threadIdx.x; e < LAY_ROWS * w; e += blockDim.x) {
                const int r = e / w, x = e - r * w;
                const int i = ihi - 4 + (r & 3), j = 4 * jb + (r >> 2), l = LAY_CHUNK * c + x;
                if (i >= 0 && (i < i_lo || i >= i_hi)) continue;
                double v = 0.0;
                if (i >= 0 && j <= i && k <= i && l <= (k < i ? k : j)) {
                    const int64_t ij = pair_index(i, j);
                    if (SRC == SRC_SYNTH) {
                        v = synth_value(np_, seed, i, j, k, l);
                    } else if (SRC == SRC_FULL_SLAB) {
                        v = src[(((int64_t)(i - i_lo) * n + j) * n + k) * n + l];
                    } else {
                        if (ij < ij_lo || ij >= ij_hi) continue;
                        v = src[ij * (ij + 1) / 2 - canon_base + pair_index(k, l)];
                    }
                    v *= lay_degeneracy(i, j, k, l);
                } else if (SRC == SRC_CANONICAL && i >= 0 && j <= i) {
                    const int64_t ij = pair_index(i, j);
                    if (ij < ij_lo || ij >= ij_hi) continue;
                }
                piece[e] = v;
            }
        }
    }
}

// tile range (clipped to the shard) that contains rows with i in [i_lo, i_hi)
static void tiles_of_irange(const sqcc_ctx *ctx, int i_lo, int i_hi, int64_t &t_lo, int64_t &t_hi)
{
    const int n = ctx->n;
    if (i_hi > n) i_hi = n;
    if (i_lo < 0) i_lo = 0;
    if (i_hi <= i_lo) {
        t_lo = t_hi = 0;
        return;
    }
    t_lo = ctx->h_tfirst[lay_iblock(n, i_lo)];
    t_hi = ctx->h_tfirst[lay_iblock(n, i_hi - 1) + 1];
    if (t_lo < ctx->t_begin) t_lo = ctx->t_begin;
    if (t_hi > ctx->t_end) t_hi = ctx->t_end;
    if (t_hi < t_lo) t_hi = t_lo;
}

static int launch_fill(sqcc_ctx *ctx, int srckind, int64_t t_lo, int64_t t_hi, uint64_t seed, const double *src, int i_lo,
                       int i_hi, int64_t ij_lo, int64_t ij_hi, int64_t canon_base)
{
    if (t_hi <= t_lo) return SQCC_OK;
    const int64_t tiles = t_hi - t_lo;
    const dim3 grid((unsigned)(tiles < (int64_t)ctx->sm_count * 4 ? tiles : (int64_t)ctx->sm_count * 4),
                    (unsigned)((ctx->n - 1) / LAY_CHUNK + 1));
    double *P = ctx->d_eri;
    const int64_t *ib = ctx->d_ibase, *tf = ctx->d_tfirst;
    const int64_t sb = ctx->shard_base;
    if (srckind == SRC_SYNTH)
        fill_tiles_kernel<SRC_SYNTH><<<grid, 256, 0, ctx->stream>>>(P, ib, tf, t_lo, t_hi, sb, ctx->n, seed, src, i_lo, i_hi, ij_lo, ij_hi, canon_base);
    else if (srckind == SRC_FULL_SLAB)
        fill_tiles_kernel<SRC_FULL_SLAB><<<grid, 256, 0, ctx->stream>>>(P, ib, tf, t_lo, t_hi, sb, ctx->n, seed, src, i_lo, i_hi, ij_lo, ij_hi, canon_base);
    else
        fill_tiles_kernel<SRC_CANONICAL><<<grid, 256, 0, ctx->stream>>>(P, ib, tf, t_lo, t_hi, sb, ctx->n, seed, src, i_lo, i_hi, ij_lo, ij_hi, canon_base);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

int eri_synth(sqcc_ctx *ctx, uint64_t seed)
{
    return launch_fill(ctx, SRC_SYNTH, ctx->t_begin, ctx->t_end, seed, nullptr, 0, ctx->n, 0, 0, 0);
}

int eri_pack_full_slab(sqcc_ctx *ctx, const double *d_slab, int p0, int np)
{
    // rows (i, *) with i in [p0, p0+np) intersected with the shard
    int64_t t_lo, t_hi;
    tiles_of_irange(ctx, p0, p0 + np, t_lo, t_hi);
    return launch_fill(ctx, SRC_FULL_SLAB, t_lo, t_hi, 0, d_slab, p0, p0 + np, 0, 0, 0);
}

int eri_pack_canonical(sqcc_ctx *ctx, const double *d_rows, int64_t ij0, int64_t nrows)
{
    int i0, j0, i1, j1;
    pair_unrank(ij0, i0, j0);
    pair_unrank(ij0 + nrows - 1, i1, j1);
    int64_t t_lo, t_hi;
    tiles_of_irange(ctx, i0, i1 + 1, t_lo, t_hi);
    return launch_fill(ctx, SRC_CANONICAL, t_lo, t_hi, 0, d_rows, 0, ctx->n, ij0, ij0 + nrows, ij0 * (ij0 + 1) / 2);
}

// ---- block ingestion: dense (i-range x j-range x k-range x l-range) blocks, e.g. shell quartets -------------------
// the integral provider hands over one dense block per (canonical) shell quartet -- Psi4:
// MintsHelper.ao_eri_shell(M, N, P, Q) -- and the full [N,N,N,N] host tensor of interface_psi4.py:177-190 (19 GB at
// N = 222, 476 GB at N = 494) never exists.  Every element is moved to its canonical position (i >= j, k >= l,
// ij >= kl) of the native layout if this shard holds its tile; symmetric images inside a block carry the same
// integral, so the duplicate stores are benign.
struct EriBlockDesc {
    int64_t offset;              // first element of the block in the value buffer
    int64_t i0, ni, j0, nj, k0, nk, l0, nl;
};

__global__ void __launch_bounds__(256) pack_blocks_kernel(double *__restrict__ P, const int64_t *__restrict__ ibase,
                                                          const int64_t *__restrict__ tfirst, int64_t t_begin,
                                                          int64_t t_end, int64_t shard_base,
                                                          const double *__restrict__ vals,
                                                          const EriBlockDesc *__restrict__ desc, int nblk, int n)
{
    for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
        const EriBlockDesc d = desc[b];
        const int64_t total = d.ni * d.nj * d.nk * d.nl;
        for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
            int l = (int)(d.l0 + e % d.nl);
            int k = (int)(d.k0 + (e / d.nl) % d.nk);
            int j = (int)(d.j0 + (e / (d.nl * d.nk)) % d.nj);
            int i = (int)(d.i0 + e / (d.nl * d.nk * d.nj));
            if (i >= n || j >= n || k >= n || l >= n) continue;
            if (i < j) { const int t = i; i = j; j = t; }
            if (k < l) { const int t = k; k = l; l = t; }
            if (pair_index(i, j) < pair_index(k, l)) {
                int t = i; i = k; k = t;
                t = j; j = l; l = t;
            }
            const int64_t tile = tfirst[lay_iblock(n, i)] + (j >> 2);
            if (tile < t_begin || tile >= t_end) continue;   // the tile of row (i, j) belongs to another rank's shard
            P[lay_offset(ibase, n, i, j, k, l) - shard_base] = vals[d.offset + e] * lay_degeneracy(i, j, k, l);
        }
    }
}

int eri_pack_blocks(sqcc_ctx *ctx, const double *d_values, const int64_t *d_desc, int nblk)
{
    if (nblk <= 0) return SQCC_OK;
    static_assert(sizeof(EriBlockDesc) == 9 * sizeof(int64_t), "descriptor is nine int64");
    const int grid = nblk < ctx->sm_count * 8 ? nblk : ctx->sm_count * 8;
    pack_blocks_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->d_ibase, ctx->d_tfirst, ctx->t_begin, ctx->t_end,
                                                     ctx->shard_base, d_values,
                                                     reinterpret_cast<const EriBlockDesc *>(d_desc), nblk, ctx->n);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

// ---- unpack: native (unsharded) -> full [N,N,N,N], plain integral values ---------------------------
__global__ void __launch_bounds__(256) unpack_full_kernel(const double *__restrict__ P, const int64_t *__restrict__ ibase,
                                                          int n, double *__restrict__ out)
{
    // blockIdx.x enumerates (p, q, r); threads run over s (coalesced writes).
    int64_t pqr = blockIdx.x;
    int r = (int)(pqr % n);
    int q = (int)((pqr / n) % n);
    int p = (int)(pqr / ((int64_t)n * n));
    int i = p >= q ? p : q, j = p >= q ? q : p;
    int64_t ij = pair_index(i, j);
    for (int s = threadIdx.x; s < n; s += blockDim.x) {
        int k = r >= s ? r : s, l = r >= s ? s : r;
        int64_t kl = pair_index(k, l);
        int a = i, b = j, c = k, d = l;
        if (kl > ij) { a = k; b = l; c = i; d = j; }
        double v = P[lay_offset(ibase, n, a, b, c, d)];
        out[pqr * n + s] = v / lay_degeneracy(a, b, c, d);
    }
}

int eri_unpack_full(sqcc_ctx *ctx, double *d_full)
{
    SQCC_REQUIRE(ctx->d_eri && ctx->whole, "sqcc_eri_unpack_full needs an unsharded ERI attached");
    int64_t blocks = (int64_t)ctx->n * ctx->n * ctx->n;
    unpack_full_kernel<<<(unsigned)blocks, ctx->n < 256 ? ((ctx->n + 31) / 32) * 32 : 256, 0, ctx->stream>>>(
        ctx->d_eri, ctx->d_ibase, ctx->n, d_full);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

}  // namespace sqcc
