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

int eri_build_tables(sqcc_ctx *ctx)
{
    int64_t nrows = ctx->ij_end - ctx->ij_begin;
    std::vector<int64_t> h(nrows + 1);
    int i, j;
    pair_unrank(ctx->ij_begin, i, j);
    int64_t off = 0, uniq = 0;
    for (int64_t r = 0; r < nrows; ++r) {
        h[r] = off;
        off += row_length(i, j);
        uniq += pair_index(i, j) + 1;
        if (++j > i) { ++i; j = 0; }
    }
    h[nrows] = off;
    ctx->native_len = off;
    ctx->n_unique = uniq;
    if (ctx->d_rowoff) cudaFree(ctx->d_rowoff);
    SQCC_CUDA(cudaMalloc(&ctx->d_rowoff, sizeof(int64_t) * (nrows + 1)));
    SQCC_CUDA(cudaMemcpyAsync(ctx->d_rowoff, h.data(), sizeof(int64_t) * (nrows + 1), cudaMemcpyHostToDevice,
                              ctx->stream));
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

// One CTA per row (i,j); warps stride over sub-rows k, lanes over l (coalesced writes / reads).
template <int SRC>
__global__ void __launch_bounds__(256) fill_rows_kernel(double *__restrict__ P, const int64_t *__restrict__ rowoff,
                                                        int64_t ij_begin, int64_t row_lo, int64_t row_hi, int n,
                                                        uint64_t seed, const double *__restrict__ src, int p0,
                                                        int64_t canon_base)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint64_t np_ = (uint64_t)n * (uint64_t)(n + 1) / 2;
    for (int64_t row = row_lo + blockIdx.x; row < row_hi; row += gridDim.x) {
        int64_t ij = ij_begin + row;
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
        fill_rows_kernel<SRC_SYNTH><<<grid, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->d_rowoff, ctx->ij_begin, row_lo,
                                                                   row_hi, ctx->n, seed, src, p0, canon_base);
    else if (srckind == SRC_FULL_SLAB)
        fill_rows_kernel<SRC_FULL_SLAB><<<grid, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->d_rowoff, ctx->ij_begin,
                                                                       row_lo, row_hi, ctx->n, seed, src, p0,
                                                                       canon_base);
    else
        fill_rows_kernel<SRC_CANONICAL><<<grid, 256, 0, ctx->stream>>>(ctx->d_eri, ctx->d_rowoff, ctx->ij_begin,
                                                                       row_lo, row_hi, ctx->n, seed, src, p0,
                                                                       canon_base);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

int eri_synth(sqcc_ctx *ctx, uint64_t seed)
{
    return launch_fill(ctx, SRC_SYNTH, 0, ctx->ij_end - ctx->ij_begin, seed, nullptr, 0, 0);
}

int eri_pack_full_slab(sqcc_ctx *ctx, const double *d_slab, int p0, int np)
{
    // rows (i, *) with i in [p0, p0+np) intersected with the shard
    int64_t row_lo = pair_index(p0, 0), row_hi = pair_index(p0 + np, 0);
    if (row_lo < ctx->ij_begin) row_lo = ctx->ij_begin;
    if (row_hi > ctx->ij_end) row_hi = ctx->ij_end;
    if (row_hi <= row_lo) return SQCC_OK;
    row_lo -= ctx->ij_begin;
    row_hi -= ctx->ij_begin;
    return launch_fill(ctx, SRC_FULL_SLAB, row_lo, row_hi, 0, d_slab, p0, 0);
}

int eri_pack_canonical(sqcc_ctx *ctx, const double *d_rows, int64_t ij0, int64_t nrows)
{
    int64_t lo = ij0 > ctx->ij_begin ? ij0 : ctx->ij_begin;
    int64_t hi = (ij0 + nrows) < ctx->ij_end ? (ij0 + nrows) : ctx->ij_end;
    if (hi <= lo) return SQCC_OK;
    return launch_fill(ctx, SRC_CANONICAL, lo - ctx->ij_begin, hi - ctx->ij_begin, 0, d_rows, 0,
                       ij0 * (ij0 + 1) / 2);
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
    SQCC_REQUIRE(ctx->d_eri && ctx->ij_begin == 0 && ctx->ij_end == pair_index(ctx->n, 0),
                 "sqcc_eri_unpack_full needs an unsharded ERI attached");
    int64_t blocks = (int64_t)ctx->n * ctx->n * ctx->n;
    unpack_full_kernel<<<(unsigned)blocks, ctx->n < 256 ? ((ctx->n + 31) / 32) * 32 : 256, 0, ctx->stream>>>(
        ctx->d_eri, ctx->d_rowoff, ctx->n, d_full);
    ctx->launches++;
    SQCC_CUDA(cudaGetLastError());
    return SQCC_OK;
}

}  // namespace sqcc
