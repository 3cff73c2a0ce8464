"""Index helpers for the 8-fold packed ERI layouts (oracle side, numpy).  TEST INFRASTRUCTURE.

Two layouts are defined.

canonical  (SURVEY.md A.1): pair index ``ij = i(i+1)/2 + j`` (i >= j), quads ``ij >= kl`` stored at
           ``ij(ij+1)/2 + kl``; holds the plain integral (ij|kl) as produced by
           ``interface_psi4.py:177-190`` (``ERI[i,j,k,l]``).

native v1  (what the CUDA kernels stream): same rows ``ij`` in the same order; inside a row the
           sub-row for ``k`` (l = 0..k, or l = 0..j when k == i) is padded to a multiple of 16
           doubles so that every sub-row starts 128-byte aligned (one L1 line per 8-lane quarter of a 128-bit load).  Pad cells are 0.0.
           Stored value is (ij|kl) * f with the degeneracy factor
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


# ---------------------------------------------------------------- native v1 layout
LAYOUT_ALIGN = 16  # doubles; must equal SUBROW_ALIGN in sqcc_b200/csrc/common.cuh


def pad_sub(x):
    return (x + (LAYOUT_ALIGN - 1)) // LAYOUT_ALIGN * LAYOUT_ALIGN


def subrow_offset(k):
    """Te(k) = sum_{k' < k} pad_sub(k' + 1) = A (M + 1) (A M / 2 + r), A = LAYOUT_ALIGN, M = k // A, r = k % A."""
    k = np.asarray(k, dtype=np.int64)
    m, r = k // LAYOUT_ALIGN, k % LAYOUT_ALIGN
    return LAYOUT_ALIGN * (m + 1) * ((LAYOUT_ALIGN // 2) * m + r)


def row_length(i, j):
    """Padded length (doubles) of row (i, j): sub-rows k < i are full, sub-row k == i holds l <= j."""
    return subrow_offset(i) + pad_sub(np.asarray(j, dtype=np.int64) + 1)


def row_offsets(n, ij_begin=0, ij_end=None):
    """int64[npair_local + 1] offsets (doubles) of rows ij_begin..ij_end-1 relative to the shard start."""
    if ij_end is None:
        ij_end = npair(n)
    I, J = pair_unrank(n)
    lens = row_length(I[ij_begin:ij_end], J[ij_begin:ij_end])
    out = np.zeros(ij_end - ij_begin + 1, dtype=np.int64)
    np.cumsum(lens, out=out[1:])
    return out


def degeneracy_factor(i, j, k, l):
    f = np.where(i == j, 0.5, 1.0) * np.where(k == l, 0.5, 1.0)
    return f * np.where((i == k) & (j == l), 0.5, 1.0)


def canonical_to_native(packed, n, ij_begin=0, ij_end=None):
    """Canonical packed vector -> native v1 shard (rows ij_begin..ij_end-1)."""
    if ij_end is None:
        ij_end = npair(n)
    I, J = pair_unrank(n)
    ro = row_offsets(n, ij_begin, ij_end)
    out = np.zeros(ro[-1])
    K, L = I, J  # same unranking for kl
    te = subrow_offset(K)
    for ij in range(ij_begin, ij_end):
        i, j = int(I[ij]), int(J[ij])
        kl = np.arange(ij + 1)
        k, l = K[kl], L[kl]
        v = packed[canonical_offset(ij, kl)] * degeneracy_factor(i, j, k, l)
        out[ro[ij - ij_begin] + te[kl] + l] = v
    return out


def native_to_canonical(native, n):
    """Inverse of canonical_to_native for a full (unsharded) tensor."""
    p = npair(n)
    I, J = pair_unrank(n)
    ro = row_offsets(n)
    out = np.zeros(nquad(n))
    te = subrow_offset(I)
    for ij in range(p):
        i, j = int(I[ij]), int(J[ij])
        kl = np.arange(ij + 1)
        k, l = I[kl], J[kl]
        out[canonical_offset(ij, kl)] = native[ro[ij] + te[kl] + l] / degeneracy_factor(i, j, k, l)
    return out


def shard_row_cuts(n, world, align=1):
    """Equal-AREA cuts of the row range [0, npair) into `world` contiguous shards.

    Returns int64[world + 1] pair-index boundaries.  `align` rounds interior cuts to i-block boundaries:
    a cut is always placed at a row with j == 0 and i % align == 0 so kernels can tile rows by i-block.
    """
    I, J = pair_unrank(n)
    ro = row_offsets(n)
    total = ro[-1]
    cuts = [0]
    # candidate cut rows: start of i-blocks
    cand_i = np.arange(0, n, align, dtype=np.int64)
    cand_ij = cand_i * (cand_i + 1) // 2
    cand_off = ro[cand_ij]
    for g in range(1, world):
        target = total * g / world
        c = int(np.argmin(np.abs(cand_off - target)))
        cuts.append(int(max(cand_ij[c], cuts[-1])))
    cuts.append(npair(n))
    return np.asarray(cuts, dtype=np.int64)
