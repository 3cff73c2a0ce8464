"""Index helpers for the 8-fold packed ERI layouts (oracle side, numpy).  TEST INFRASTRUCTURE.

Two layouts are defined.

canonical  (SURVEY.md A.1): pair index ``ij = i(i+1)/2 + j`` (i >= j), quads ``ij >= kl`` stored at
           ``ij(ij+1)/2 + kl``; holds the plain integral (ij|kl) as produced by
           ``interface_psi4.py:177-190`` (``ERI[i,j,k,l]``).

native v2  (what the CUDA kernels stream; sqcc_b200/csrc/layout.cuh): tile-interleaved.  The 16 rows of a 4 x 4 block
           of rows (i, j) lie side by side for every (sub-row k, 64-column chunk) as one contiguous "piece"; cells
           outside the canonical range are 0.0.  Stored value is (ij|kl) * f with the degeneracy factor
           f = 0.5^[i==j] * 0.5^[k==l] * 0.5^[ij==kl]   (exact powers of two).
"""
import numpy as np


def npair(n):
    return n * (n + 1) // 2


def nquad(n):
    p = npair(n)
    return p * (p + 1) // 2


def pair_index(i, j):
    """ij for i >= j (arrays ok)."""
    return i * (i + 1) // 2 + j


def pair_unrank(n):
    """Arrays (I, J) with I[ij] >= J[ij] for ij in 0..npair-1."""
    i, j = np.tril_indices(n)
    return i.astype(np.int64), j.astype(np.int64)


def canonical_offset(ij, kl):
    """Offset of (ij|kl) in the canonical packed array; requires ij >= kl."""
    return ij * (ij + 1) // 2 + kl


def pack_canonical(eri_full):
    """Full [N,N,N,N] tensor -> canonical packed vector (takes the canonical element, no averaging)."""
    n = eri_full.shape[0]
    p = npair(n)
    I, J = pair_unrank(n)
    m = eri_full[I[:, None], J[:, None], I[None, :], J[None, :]]  # [npair, npair]
    r, c = np.tril_indices(p)
    return np.ascontiguousarray(m[r, c])


def unpack_canonical(packed, n):
    """Canonical packed vector -> full 8-fold symmetric [N,N,N,N] tensor."""
    p = npair(n)
    m = np.zeros((p, p))
    r, c = np.tril_indices(p)
    m[r, c] = packed
    m[c, r] = packed
    idx = np.zeros((n, n), dtype=np.int64)
    I, J = pair_unrank(n)
    idx[I, J] = np.arange(p)
    idx[J, I] = np.arange(p)
    return np.ascontiguousarray(m[idx[:, :, None, None], idx[None, None, :, :]])


# ---------------------------------------------------------------- native v2 layout (sqcc_b200/csrc/layout.cuh)
# Tile-interleaved: a tile is a 4 x 4 block of rows (i, j); i-blocks are anchored at the top of the i range (block ib
# covers i in [ihi - 4, ihi), ihi = n - 4 (nb - 1 - ib)); j-blocks start at multiples of 4; tile order ib ascending, jb
# ascending.  Inside a tile, chunk-major: for c: for k = 64 c .. imax: a "piece" [16][w] with w = min(64, 16 ((k - 64 c)
# // 16 + 1)); row slot r = 4 (j % 4) + (i - (ihi - 4)).  Cells outside the canonical range hold 0.0.
LAY_CHUNK = 64
LAY_ROWS = 16


def lay_nblocks(n):
    return (n + 3) // 4


def lay_ihi(n, ib):
    return n - 4 * (lay_nblocks(n) - 1 - ib)


def lay_iblock(n, i):
    return lay_nblocks(n) - 1 - (n - 1 - i) // 4


def lay_njb(imax):
    return imax // 4 + 1


def lay_w(m):
    return 64 if m >= 48 else 16 * (m // 16 + 1)


def lay_S(m):
    """sum of lay_w(m') for m' < m."""
    if m >= 48:
        return 64 * (m - 24)
    q, r = m // 16, m % 16
    return 16 * (q + 1) * (8 * q + r)


def lay_piece_offset(imax, c, k):
    return LAY_ROWS * (64 * (c * (imax - 23) - 32 * c * (c - 1)) + lay_S(k - LAY_CHUNK * c))


def lay_tile_len(imax):
    return lay_piece_offset(imax, imax // 64, imax + 1)


def tile_tables(n):
    """(ibase, tfirst): offset (doubles) and tile index of the first tile of every i-block, length nb + 1."""
    nb = lay_nblocks(n)
    ibase, tfirst = [0], [0]
    for ib in range(nb):
        imax = lay_ihi(n, ib) - 1
        ibase.append(ibase[-1] + lay_njb(imax) * lay_tile_len(imax))
        tfirst.append(tfirst[-1] + lay_njb(imax))
    return ibase, tfirst


def tile_count(n):
    return tile_tables(n)[1][-1]


def tile_offset(n, t):
    ibase, tfirst = tile_tables(n)
    nb = lay_nblocks(n)
    if t >= tfirst[nb]:
        return ibase[nb]
    ib = max(b for b in range(nb) if tfirst[b] <= t)
    return ibase[ib] + (t - tfirst[ib]) * lay_tile_len(lay_ihi(n, ib) - 1)


def native_len(n, t_begin=0, t_end=None):
    if t_end is None:
        t_end = tile_count(n)
    return tile_offset(n, t_end) - tile_offset(n, t_begin)


def native_offset(n, i, j, k, l):
    """Offset (doubles, from the start of the full tensor) of the canonical element (i, j, k, l)."""
    ibase, _ = tile_tables(n)
    ib = lay_iblock(n, i)
    ihi = lay_ihi(n, ib)
    imax = ihi - 1
    c = l // 64
    m = k - 64 * c
    slot = 4 * (j % 4) + (i - (ihi - 4))
    return ibase[ib] + (j // 4) * lay_tile_len(imax) + lay_piece_offset(imax, c, k) + slot * lay_w(m) + (l - 64 * c)


def degeneracy_factor(i, j, k, l):
    f = np.where(i == j, 0.5, 1.0) * np.where(k == l, 0.5, 1.0)
    return f * np.where((i == k) & (j == l), 0.5, 1.0)


def canonical_to_native(packed, n, t_begin=0, t_end=None):
    """Canonical packed vector -> native v2 shard holding the tiles [t_begin, t_end)."""
    ibase, tfirst = tile_tables(n)
    if t_end is None:
        t_end = tfirst[-1]
    base = tile_offset(n, t_begin)
    out = np.zeros(tile_offset(n, t_end) - base)
    for ib in range(lay_nblocks(n)):
        ihi = lay_ihi(n, ib)
        imax = ihi - 1
        tl = lay_tile_len(imax)
        for jb in range(lay_njb(imax)):
            t = tfirst[ib] + jb
            if t < t_begin or t >= t_end:
                continue
            tile = ibase[ib] + jb * tl - base
            for a in range(4):
                i = ihi - 4 + a
                if i < 0:
                    continue
                for b in range(4):
                    j = 4 * jb + b
                    if j > i:
                        continue
                    ij = pair_index(i, j)
                    for k in range(i + 1):
                        lmax = k if k < i else j
                        l = np.arange(lmax + 1)
                        v = packed[canonical_offset(ij, pair_index(k, l))] * degeneracy_factor(i, j, k, l)
                        c = l // 64
                        m = k - 64 * c
                        w = np.where(m >= 48, 64, 16 * (m // 16 + 1))
                        po = np.array([lay_piece_offset(imax, int(cc), k) for cc in range(int(c.max()) + 1)])
                        out[tile + po[c] + (4 * b + a) * w + (l - 64 * c)] = v
    return out


def native_to_canonical(native, n):
    """Inverse of canonical_to_native for a full (unsharded) tensor."""
    p = npair(n)
    I, J = pair_unrank(n)
    out = np.zeros(nquad(n))
    for ij in range(p):
        i, j = int(I[ij]), int(J[ij])
        for kl in range(ij + 1):
            k, l = int(I[kl]), int(J[kl])
            out[canonical_offset(ij, kl)] = native[native_offset(n, i, j, k, l)] / float(degeneracy_factor(i, j, k, l))
    return out
