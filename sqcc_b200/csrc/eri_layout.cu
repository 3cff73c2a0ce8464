// ERI ingest kernels: native-v1 packed layout tables, synthetic fill, pack (full / canonical) and unpack.
// Consumes what interface_psi4.py:177-190 (ao_electron_repulsion_integral, full [N,N,N,N]) produces.
#include <math.h>

#include "common.cuh"

namespace sqcc {

int64_t iblock_offset(int64_t i)
{
    int64_t off = 0;
    for (int64_t t = 0; t < i; ++t) off += (t + 1) * subrow_offset(t) + subrow_offset(t + 1);
    return off;
}

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

// ---- shard geometry (host) ---------------------------------------------------------------------------
// Tiles ("pair-blocks"): 4 x 4 blocks of rows (i, j).  The i-blocks are anchored at the TOP of the i range
// (block b covers i in [n - 4 (NB - b), n - 4 (NB - b - 1)) clipped to >= 0), so that the longest rows always sit in
// full blocks; j-blocks start at multiples of 4.  Global tile order: i-block ascending, j-block ascending.
static inline void iblock_range(int n, int b, int &lo, int &hi)
{
    const int nb = (n + 3) / 4;
    lo = n - 4 * (nb - b);
    hi = lo + 4;
    if (lo < 0) lo = 0;
}

int64_t tile_count(int n)
{
    const int nb = (n + 3) / 4;
    int64_t t = 0;
    for (int b = 0; b < nb; ++b) {
        int lo, hi;
        iblock_range(n, b, lo, hi);
        t += (hi - 1) / 4 + 1;
    }
    return t;
}

int tile_intervals(int n, int64_t t_begin, int64_t t_end, std::vector<int> &jlo, std::vector<int> &jhi)
{
    jlo.assign(n, 0);
    jhi.assign(n, 0);
    const int nb = (n + 3) / 4;
    int64_t t0 = 0;
    for (int b = 0; b < nb; ++b) {
        int lo, hi;
        iblock_range(n, b, lo, hi);
        const int64_t cnt = (hi - 1) / 4 + 1;
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

int row_intervals(int n, int64_t ij_begin, int64_t ij_end, std::vector<int> &jlo, std::vector<int> &jhi)
{
    jlo.assign(n, 0);
    jhi.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        const int64_t r0 = pair_index(i, 0), r1 = r0 + i + 1;
        const int64_t a = ij_begin > r0 ? ij_begin : r0, c = ij_end < r1 ? ij_end : r1;
        if (c > a) {
            jlo[i] = (int)(a - r0);
            jhi[i] = (int)(c - r0);
        }
    }
    return SQCC_OK;
}

int64_t intervals_native_len(int n, const std::vector<int> &jlo, const std::vector<int> &jhi)
{
    int64_t len = 0;
    for (int i = 0; i < n; ++i)
        if (jhi[i] > jlo[i]) len += (int64_t)(jhi[i] - jlo[i]) * subrow_offset(i) + subrow_offset(jhi[i]) - subrow_offset(jlo[i]);
    return len;
}

int eri_build_tables(sqcc_ctx *ctx)
{
    const int n = ctx->n;
    ctx->h_rowptr.assign(n + 1, 0);
    ctx->h_ioff.assign(n, 0);
    ctx->i_begin = n;
    ctx->i_end = 0;
    for (int i = 0; i < n; ++i) {
        ctx->h_rowptr[i + 1] = ctx->h_rowptr[i] + (ctx->jhi[i] - ctx->jlo[i]);
        if (ctx->jhi[i] > ctx->jlo[i]) {
            if (i < ctx->i_begin) ctx->i_begin = i;
            ctx->i_end = i + 1;
        }
    }
    if (ctx->i_end == 0) ctx->i_begin = 0;
    const int64_t nrows = ctx->h_rowptr[n];
    ctx->nrows = nrows;
    ctx->whole = nrows == pair_index(n, 0);
    std::vector<int64_t> h(nrows + 1), hij(nrows > 0 ? nrows : 1);
    int64_t off = 0, uniq = 0, r = 0;
    for (int i = 0; i < n; ++i) {
        if (ctx->jhi[i] > ctx->jlo[i])
            ctx->h_ioff[i] = off - ((int64_t)ctx->jlo[i] * subrow_offset(i) + subrow_offset(ctx->jlo[i]));
        for (int j = ctx->jlo[i]; j < ctx->jhi[i]; ++j, ++r) {
            h[r] = off;
            hij[r] = pair_index(i, j);
            off += row_length(i, j);
            uniq += pair_index(i, j) + 1;
        }
    }
    h[nrows] = off;
    ctx->native_len = off;
    ctx->n_unique = uniq;
    if (ctx->d_rowoff) cudaFree(ctx->d_rowoff);
    if (ctx->d_rowij) cudaFree(ctx->d_rowij);
    if (ctx->d_ioff) cudaFree(ctx->d_ioff);
    if (ctx->d_jlo) cudaFree(ctx->d_jlo);
    ctx->d_rowoff = ctx->d_rowij = ctx->d_ioff = nullptr;
    ctx->d_jlo = ctx->d_jhi = nullptr;
    SQCC_CUDA(cudaMalloc(&ctx->d_jlo, sizeof(int) * 2 * n));
    ctx->d_jhi = ctx->d_jlo + n;
    SQCC_CUDA(cudaMemcpyAsync(ctx->d_jlo, ctx->jlo.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    SQCC_CUDA(cudaMemcpyAsync(ctx->d_jhi, ctx->jhi.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    SQCC_CUDA(cudaMalloc(&ctx->d_rowoff, sizeof(int64_t) * (nrows + 1)));
    SQCC_CUDA(cudaMalloc(&ctx->d_rowij, sizeof(int64_t) * hij.size()));
    SQCC_CUDA(cudaMalloc(&ctx->d_ioff, sizeof(int64_t) * n));
    SQCC_CUDA(cudaMemcpyAsync(ctx->d_rowoff, h.data(), sizeof(int64_t) * (nrows + 1), cudaMemcpyHostToDevice,
                              ctx->stream));
    SQCC_CUDA(cudaMemcpyAsync(ctx->d_rowij, hij.data(), sizeof(int64_t) * hij.size(), cudaMemcpyHostToDevice,
                              ctx->stream));
    SQCC_CUDA(cudaMemcpyAsync(ctx->d_ioff, ctx->h_ioff.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    SQCC_CUDA(cudaStreamSynchronize(ctx->stream));
    return SQCC_OK;
}

// local row range [lo, hi) of the held rows whose pair index lies in [ij_lo, ij_hi)
static void local_row_range(const sqcc_ctx *ctx, int64_t ij_lo, int64_t ij_hi, int64_t &lo, int64_t &hi)
{
    auto count_before = [&](int64_t ij) {   // held rows with pair index < ij
        if (ij <= 0) return (int64_t)0;
        if (ij >= pair_index(ctx->n, 0)) return ctx->nrows;
        int i, j;
        pair_unrank(ij, i, j);
        int64_t c = ctx->h_rowptr[i];
        const int a = ctx->jlo[i], b = ctx->jhi[i];
        if (j > a) c += (j < b ? j : b) - a;
        return c;
    };
    lo = count_before(ij_lo);
    hi = count_before(ij_hi);
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

// One CTA per row (i,j); warps stride over sub-rows k, lanes over l (coalesced writes / reads).
template <int SRC>
__global__ void __launch_bounds__(256) fill_rows_kernel(double *__restrict__ P, const int64_t *__restrict__ rowoff,
                                                        const int64_t *__restrict__ rowij, int64_t row_lo, int64_t row_hi, int n,
                                                        uint64_t seed, const double *__restrict__ src, int p0,
                                                        int64_t canon_base)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint64_t np_ = (uint64_t)n * (uint64_t)(n + 1) / 2;
    for (int64_t row = row_lo + blockIdx.x; row < row_hi; row += gridDim.x) {
        int64_t ij = rowij[row];
        int i, j;
        pair_unrank(ij, i, j);
        double *base = P + rowoff[row];
        for (int k = warp; k <= i; k += nwarps) {
            int lmax = k < i ? k : j;
            double *sub = base + subrow_offset(k);
            for (int l = lane; l <= lmax; l += 32) {
                double v;
                if (SRC == SRC_SYNTH) {
                    v = synth_value(np_, seed, i, j, k, l);
                } else if (SRC == SRC_FULL_SLAB) {
                    v = src[(((int64_t)(i - p0) * n + j) * n + k) * n + l];
                } else {
                    v = src[ij * (ij + 1) / 2 - canon_base + pair_index(k, l)];
                }
                sub[l] = v * degeneracy(i, j, k, l);
            }
            for (int l = lmax + 1 + lane; l < (int)pad_sub(lmax + 1); l += 32) sub[l] = 0.0;  // pad cells
        }
    }
}

static int launch_fill(sqcc_ctx *ctx, int srckind, int64_t row_lo, int64_t row_hi, uint64_t seed, const double *src,
                       int p0, int64_t canon_base)
{
    if (row_hi <= row_lo) return SQCC_OK;
    int64_t rows = row_hi - row_lo;
    int grid = (int)(rows < (int64_t)ctx->sm_count * 16 ? rows : (int64_t)ctx->sm_count * 16);
    if (srckind == SRC_SYNTH)
        fill_rows_kernel<SRC_SYNTH><<<grid, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->d_rowoff, ctx->d_rowij, row_lo,
                                                                   row_hi, ctx->n, seed, src, p0, canon_base);
    else if (srckind == SRC_FULL_SLAB)
        fill_rows_kernel<SRC_FULL_SLAB><<<grid, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->d_rowoff, ctx->d_rowij,
                                                                       row_lo, row_hi, ctx->n, seed, src, p0,
                                                                       canon_base);
    else
        fill_rows_kernel<SRC_CANONICAL><<<grid, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->d_rowoff, ctx->d_rowij,
                                                                       row_lo, row_hi, ctx->n, seed, src, p0,
                                                                       canon_base);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

int eri_synth(sqcc_ctx *ctx, uint64_t seed)
{
    return launch_fill(ctx, SRC_SYNTH, 0, ctx->nrows, seed, nullptr, 0, 0);
}

int eri_pack_full_slab(sqcc_ctx *ctx, const double *d_slab, int p0, int np)
{
    // rows (i, *) with i in [p0, p0+np) intersected with the shard
    int64_t row_lo, row_hi;
    local_row_range(ctx, pair_index(p0, 0), pair_index(p0 + np, 0), row_lo, row_hi);
    if (row_hi <= row_lo) return SQCC_OK;
    return launch_fill(ctx, SRC_FULL_SLAB, row_lo, row_hi, 0, d_slab, p0, 0);
}

int eri_pack_canonical(sqcc_ctx *ctx, const double *d_rows, int64_t ij0, int64_t nrows)
{
    int64_t lo, hi;
    local_row_range(ctx, ij0, ij0 + nrows, lo, hi);
    if (hi <= lo) return SQCC_OK;
    return launch_fill(ctx, SRC_CANONICAL, lo, hi, 0, d_rows, 0, ij0 * (ij0 + 1) / 2);
}

// ---- block ingestion: dense (i-range x j-range x k-range x l-range) blocks, e.g. shell quartets -------------------
// the integral provider hands over one dense block per (canonical) shell quartet -- Psi4:
// MintsHelper.ao_eri_shell(M, N, P, Q) -- and the full [N,N,N,N] host tensor of interface_psi4.py:177-190 (19 GB at
// N = 222, 476 GB at N = 494) never exists.  Every element is moved to its canonical position (i >= j, k >= l,
// ij >= kl) of the native layout if this shard holds row (i, j); symmetric images inside a block carry the same
// integral, so the duplicate stores are benign.
struct EriBlockDesc {
    int64_t offset;              // first element of the block in the value buffer
    int64_t i0, ni, j0, nj, k0, nk, l0, nl;
};

__global__ void __launch_bounds__(256) pack_blocks_kernel(double *__restrict__ P, const int64_t *__restrict__ ioff,
                                                          const int *__restrict__ jlo, const int *__restrict__ jhi,
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
            if (j < jlo[i] || j >= jhi[i]) continue;   // row (i, j) belongs to another rank's shard
            const int64_t off = ioff[i] + (int64_t)j * subrow_offset(i) + subrow_offset(j) + subrow_offset(k) + l;
            P[off] = vals[d.offset + e] * degeneracy(i, j, k, l);
        }
    }
}

int eri_pack_blocks(sqcc_ctx *ctx, const double *d_values, const int64_t *d_desc, int nblk)
{
    if (nblk <= 0) return SQCC_OK;
    static_assert(sizeof(EriBlockDesc) == 9 * sizeof(int64_t), "descriptor is nine int64");
    const int grid = nblk < ctx->sm_count * 8 ? nblk : ctx->sm_count * 8;
    pack_blocks_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->d_ioff, ctx->d_jlo, ctx->d_jhi, d_values,
                                                     reinterpret_cast<const EriBlockDesc *>(d_desc), nblk, ctx->n);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

// ---- unpack: native (unsharded) -> full [N,N,N,N], plain integral values ---------------------------
__global__ void __launch_bounds__(256) unpack_full_kernel(const double *__restrict__ P,
                                                          const int64_t *__restrict__ rowoff, int n,
                                                          double *__restrict__ out)
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
        double v = P[rowoff[pair_index(a, b)] + subrow_offset(c) + d];
        out[pqr * n + s] = v / degeneracy(a, b, c, d);
    }
}

int eri_unpack_full(sqcc_ctx *ctx, double *d_full)
{
    SQCC_REQUIRE(ctx->d_eri && ctx->whole, "sqcc_eri_unpack_full needs an unsharded ERI attached");
    int64_t blocks = (int64_t)ctx->n * ctx->n * ctx->n;
    unpack_full_kernel<<<(unsigned)blocks, ctx->n < 256 ? ((ctx->n + 31) / 32) * 32 : 256, 0, ctx->stream>>>(
        ctx->d_eri, ctx->d_rowoff, ctx->n, d_full);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

}  // namespace sqcc
