"""Coulomb / exchange contraction oracle (numpy).  TEST INFRASTRUCTURE.

Three restatements of the same maths, from most literal to fastest:

* ``jk_loop``    -- the reference's own loop form, statement for statement
                    (hf_ksdft.py:483-487 J, :491-495 K; UHF :510-515, :522-526, :531-535).
* ``jk_einsum``  -- einsum over the full tensor (validated against ``jk_loop``).
* ``jk_packed``  -- the 8-fold accumulation formula on the canonical packed vector
                    (SURVEY.md A.1), i.e. the algorithm the CUDA kernel implements.
"""
import numpy as np

from . import packing as pk


def j_loop(eri, d):
    """hf_ksdft.py:483-487 (RHF/RKS) and :511-515 (UHF/UKS with total density)."""
    n = d.shape[0]
    out = np.zeros((n, n))
    for p in range(n):
        for q in range(n):
            out[p, q] = np.sum(eri[p, q, :, :] * d)
    return out


def k_loop(eri, d):
    """hf_ksdft.py:491-495 (RHF), :522-526 (alpha), :531-535 (beta)."""
    n = d.shape[0]
    out = np.zeros((n, n))
    for p in range(n):
        for q in range(n):
            out[p, q] = np.sum(eri[p, :, q, :] * d)
    return out


def jk_loop(eri, d, want_k=True):
    """Reference semantics.  d [N,N] -> (J(d), K(d)); d [2,N,N] -> (J(d0+d1), [K(d0), K(d1)])."""
    if d.ndim == 2:
        return j_loop(eri, d), (k_loop(eri, d) if want_k else None)
    total = d[0] + d[1]                                   # hf_ksdft.py:510
    j = j_loop(eri, total)
    k = np.stack([k_loop(eri, d[0]), k_loop(eri, d[1])]) if want_k else None
    return j, k


def jk_einsum(eri, d, want_k=True):
    if d.ndim == 2:
        j = np.einsum("pqrs,rs->pq", eri, d)
        k = np.einsum("prqs,rs->pq", eri, d) if want_k else None
        return j, k
    j = np.einsum("pqrs,rs->pq", eri, d[0] + d[1])
    k = np.stack([np.einsum("prqs,rs->pq", eri, d[s]) for s in range(2)]) if want_k else None
    return j, k


def jk_packed(packed, n, d, want_k=True):
    """SURVEY.md A.1 on the canonical packed vector (vectorised per row ij)."""
    I, J = pk.pair_unrank(n)
    dj = d if d.ndim == 2 else d[0] + d[1]
    ds = [d] if d.ndim == 2 else [d[0], d[1]]
    jt = np.zeros((n, n))
    kt = [np.zeros((n, n)) for _ in ds]
    for ij in range(pk.npair(n)):
        i, j = int(I[ij]), int(J[ij])
        kl = np.arange(ij + 1)
        k, l = I[kl], J[kl]
        v = packed[pk.canonical_offset(ij, kl)] * pk.degeneracy_factor(i, j, k, l)
        jt[i, j] += np.dot(v, dj[k, l] + dj[l, k])
        np.add.at(jt, (k, l), v * (dj[i, j] + dj[j, i]))
        if want_k:
            for s, dd in enumerate(ds):
                np.add.at(kt[s], (np.full_like(k, i), k), v * dd[j, l])
                np.add.at(kt[s], (np.full_like(k, i), l), v * dd[j, k])
                np.add.at(kt[s], (np.full_like(k, j), k), v * dd[i, l])
                np.add.at(kt[s], (np.full_like(k, j), l), v * dd[i, k])
    jm = jt + jt.T
    if not want_k:
        return jm, None
    km = [x + x.T for x in kt]
    return jm, (km[0] if d.ndim == 2 else np.stack(km))


def jk_sampled(n, seed, d, entries, want_k=True):
    """J/K at selected (p, q) entries with lazily generated ERI slabs (N=494: no full tensor).

    Follows hf_ksdft.py:486-487 / :494-495 per entry.  Returns (J[len], K[nspin, len] or None).
    """
    from . import synth
    dj = d if d.ndim == 2 else d[0] + d[1]
    ds = [d] if d.ndim == 2 else [d[0], d[1]]
    jv = np.zeros(len(entries))
    kv = np.zeros((len(ds), len(entries)))
    for t, (p, q) in enumerate(entries):
        jv[t] = np.sum(synth.synth_eri_slab_pq(n, seed, p, q) * dj)
        if want_k:
            slab = synth.synth_eri_slab_pxqx(n, seed, p, q)
            for s, dd in enumerate(ds):
                kv[s, t] = np.sum(slab * dd)
    return jv, (kv if want_k else None)
