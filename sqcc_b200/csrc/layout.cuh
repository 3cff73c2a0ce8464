// Native v2 packed ERI layout: tile-interleaved, piece-contiguous (host + device arithmetic).
//
// 8-fold symmetry: canonical quads i >= j, k >= l, ij >= kl with ij = i(i+1)/2 + j.  The stored value is
// (ij|kl) * 0.5^[i==j] * 0.5^[k==l] * 0.5^[ij==kl] (exact), so the accumulation formulas need no special cases.
//
//   tile   a 4 x 4 block of rows (i, j) ("pair-block").  i-blocks are anchored at the TOP of the i range: block ib covers
//          i in [ihi - 4, ihi), ihi = n - 4 (nb - 1 - ib), nb = ceil(n / 4) (the lowest block is clipped at i = 0);
//          j-blocks start at multiples of 4.  Global tile order: ib ascending, jb ascending.  A shard is a contiguous
//          tile range, and so a contiguous byte range.
//   piece  inside a tile, for a sub-row k <= imax = ihi - 1 and a 64-column chunk c <= k / 64: the 16 rows of the tile
//          side by side, [16][w] doubles with w = min(64, 16 ((k - 64 c) / 16 + 1)) (the triangular edge l <= k is cut
//          at multiples of 16 columns).  Row slot r = 4 b + a holds row (i, j) = (ihi - 4 + a, 4 jb + b), so the two
//          4 x 2 halves of a tile (b < 2 / b >= 2) are the two contiguous halves of every piece.  Cells outside the
//          canonical range (j > i, k > i, l > k, l > j when k == i, i < 0) are stored as 0.0: consumers are mask free.
//   order  pieces of a tile are stored chunk-major: for c: for k = 64 c .. imax.  The 16 sub-rows x 64 columns x 16 rows
//          a warp task of the J/K kernel sweeps are therefore ONE contiguous run of 8 KB pieces: one cp.async.bulk per
//          pipeline stage, no per-row address arithmetic, no masks.
//
// All tiles of an i-block have the same length, so a tile's offset is ibase[ib] + jb * tile_len(imax); everything
// inside a tile is closed-form.  Size: N = 494 63.5 GB (+6.2 % over the 59.80 GB of unique integrals).
#pragma once
#include <stdint.h>

namespace sqcc {

constexpr int LAY_CHUNK = 64;   // columns per chunk
constexpr int LAY_ROWS = 16;    // rows per tile (4 x 4)

__host__ __device__ __forceinline__ int lay_nblocks(int n) { return (n + 3) >> 2; }
// one past the last i of i-block ib
__host__ __device__ __forceinline__ int lay_ihi(int n, int ib) { return n - 4 * (lay_nblocks(n) - 1 - ib); }
__host__ __device__ __forceinline__ int lay_iblock(int n, int i) { return lay_nblocks(n) - 1 - ((n - 1 - i) >> 2); }
// number of j-blocks (tiles) of the i-block whose last row is imax
__host__ __device__ __forceinline__ int lay_njb(int imax) { return (imax >> 2) + 1; }
// columns of the piece m = k - 64 c sub-rows into its chunk
__host__ __device__ __forceinline__ int lay_w(int m) { return m >= 48 ? 64 : 16 * ((m >> 4) + 1); }
// S(m) = sum_{m' < m} w(m')
__host__ __device__ __forceinline__ int64_t lay_S(int64_t m)
{
    if (m >= 48) return 64 * (m - 24);
    const int64_t q = m >> 4, r = m & 15;
    return 16 * (q + 1) * (8 * q + r);
}
// offset (doubles) of piece (k, c) from the start of its tile; also valid one past the last piece of chunk c
__host__ __device__ __forceinline__ int64_t lay_piece_offset(int imax, int c, int k)
{
    return LAY_ROWS * (64 * ((int64_t)c * (imax - 23) - 32 * (int64_t)c * (c - 1)) + lay_S(k - LAY_CHUNK * c));
}
__host__ __device__ __forceinline__ int64_t lay_tile_len(int imax) { return lay_piece_offset(imax, imax >> 6, imax + 1); }
__host__ __device__ __forceinline__ double lay_degeneracy(int i, int j, int k, int l)
{
    double f = 1.0;
    if (i == j) f *= 0.5;
    if (k == l) f *= 0.5;
    if (i == k && j == l) f *= 0.5;
    return f;
}
// (row slot, valid) of row (i, j) inside its tile
__host__ __device__ __forceinline__ int lay_slot(int ihi, int i, int j) { return 4 * (j & 3) + (i - (ihi - 4)); }

// Offset (doubles) of the canonical element (i, j, k, l) from the start of the FULL tensor, given ibase[] (offset of the
// first tile of every i-block, [nb + 1]).
__host__ __device__ __forceinline__ int64_t lay_offset(const int64_t *ibase, int n, int i, int j, int k, int l)
{
    const int ib = lay_iblock(n, i), ihi = lay_ihi(n, ib), imax = ihi - 1;
    const int c = l >> 6, m = k - LAY_CHUNK * c;
    return ibase[ib] + (int64_t)(j >> 2) * lay_tile_len(imax) + lay_piece_offset(imax, c, k) +
           (int64_t)lay_slot(ihi, i, j) * lay_w(m) + (l - LAY_CHUNK * c);
}

}  // namespace sqcc
