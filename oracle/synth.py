"""Counter-based synthetic inputs.  TEST INFRASTRUCTURE.

The same generator is implemented on the device (``sqcc_b200/csrc/eri_layout.cu``: ``synth_value``);
every step is an exact IEEE-754 operation (integer hash, int->double, one divide, one multiply),
so the CPU and GPU tensors are BIT-IDENTICAL.

  value(ij, kl) for canonical ij >= kl, pairs ij=(i,j), kl=(k,l):
      z   = splitmix64(seed + GOLDEN * (ij * npair + kl + 1))
      u   = (z >> 11) * 2^-53                     in [0, 1)
      t   = 2u - 1                                 in [-1, 1)
      s   = 1 / (1 + 0.25 * (|i-j| + |k-l|))       decay away from the pair diagonals
      val = s * (|t| + 1)  if ij == kl  else  s * t
"""
import numpy as np

from . import packing as pk

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x.copy()
        z = (z ^ (z >> np.uint64(30))) * M1
        z = (z ^ (z >> np.uint64(27))) * M2
        z = z ^ (z >> np.uint64(31))
    return z


def synth_values(n, seed, ij, kl):
    """Synthetic (ij|kl) for arrays of canonical pair indices ij >= kl."""
    p = np.uint64(pk.npair(n))
    ij = np.asarray(ij, dtype=np.int64)
    kl = np.asarray(kl, dtype=np.int64)
    I, J = pk.pair_unrank(n)
    with np.errstate(over="ignore"):
        ctr = ij.astype(np.uint64) * p + kl.astype(np.uint64) + np.uint64(1)
        z = splitmix64(np.uint64(seed) + GOLDEN * ctr)
    u = (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
    t = 2.0 * u - 1.0
    d = (np.abs(I[ij] - J[ij]) + np.abs(I[kl] - J[kl])).astype(np.float64)
    s = 1.0 / (1.0 + 0.25 * d)
    return np.where(ij == kl, s * (np.abs(t) + 1.0), s * t)


def synth_eri_canonical(n, seed):
    """Whole canonical packed vector (use only for N where nquad fits in host memory)."""
    p = pk.npair(n)
    r, c = np.tril_indices(p)
    return synth_values(n, seed, r, c)


def synth_eri_full(n, seed):
    """Full [N,N,N,N] 8-fold-symmetric tensor, as ``interface_psi4.py:177-190`` would return it."""
    return pk.unpack_canonical(synth_eri_canonical(n, seed), n)


def synth_eri_slab_pq(n, seed, p, q):
    """ERI[p, q, :, :] ([N, N]) generated lazily -- the operand of ``hf_ksdft.py:486-487``."""
    r, s = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return _gather(n, seed, np.full_like(r, p), np.full_like(r, q), r, s)


def synth_eri_slab_pxqx(n, seed, p, q):
    """ERI[p, :, q, :] ([N, N]) generated lazily -- the operand of ``hf_ksdft.py:494-495``."""
    r, s = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return _gather(n, seed, np.full_like(r, p), r, np.full_like(r, q), s)


def _gather(n, seed, a, b, c, d):
    ab = pk.pair_index(np.maximum(a, b), np.minimum(a, b))
    cd = pk.pair_index(np.maximum(c, d), np.minimum(c, d))
    hi = np.maximum(ab, cd)
    lo = np.minimum(ab, cd)
    return synth_values(n, seed, hi, lo)


def synth_orbitals(n, seed):
    """Orthonormal C [N,N] from QR of a seeded Gaussian matrix."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return q


def synth_density_rhf(n, n_occ, seed):
    """D = 2 C_occ C_occ^T  (hf_ksdft.py:150-156)."""
    c = synth_orbitals(n, seed)[:, :n_occ]
    return 2.0 * c @ c.T


def synth_density_uhf(n, n_alpha, n_beta, seed):
    """D[0] = Ca Ca^T, D[1] = Cb Cb^T  (hf_ksdft.py:185-199)."""
    ca = synth_orbitals(n, seed)[:, :n_alpha]
    cb = synth_orbitals(n, seed + 1)[:, :n_beta]
    return np.stack([ca @ ca.T, cb @ cb.T])


def synth_grid(g, n, seed):
    """phi [G,N] and w [G]: phi = N(0,1) * exp(-U(0,8)) per row, w = U(0, 1e-2)."""
    rng = np.random.default_rng(seed)
    phi = rng.standard_normal((g, n)) * np.exp(-rng.uniform(0.0, 8.0, size=(g, 1)))
    w = rng.uniform(0.0, 1e-2, size=g)
    return np.ascontiguousarray(phi), w


def synth_mo(n, n_occ, seed, gap=0.3):
    """Orthonormal C and sorted orbital energies with a >= gap Eh HOMO-LUMO gap."""
    rng = np.random.default_rng(seed)
    c = synth_orbitals(n, seed)
    eps = np.sort(rng.uniform(-2.0, -0.4, size=n_occ)).tolist() + \
        np.sort(rng.uniform(-0.4 + gap, 3.0, size=n - n_occ)).tolist()
    return c, np.asarray(eps)
