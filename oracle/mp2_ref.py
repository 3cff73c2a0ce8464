"""Closed-shell MP2 oracle (numpy).  TEST INFRASTRUCTURE.

``mp2_loop`` follows mp2.py:89-99 (denominators), :123-170 (four quarter transforms) and :240-254
(energy sum) statement for statement (the reference's second, identical transform :185-228 yields the
same array, so ``ibja[i,b,j,a] == iajb[i,b,j,a]``).  ``mp2_einsum`` is the fast restatement.
"""
import numpy as np


def transform_loop(eri, c_occ, c_virt):
    n = eri.shape[0]
    no, nv = c_occ.shape[1], c_virt.shape[1]
    t1 = np.zeros((no, n, n, n))
    for i in range(no):
        for q in range(n):
            for r in range(n):
                for s in range(n):
                    t1[i, q, r, s] = np.dot(c_occ[:, i], eri[:, q, r, s])
    t2 = np.zeros((no, n, no, n))
    for i in range(no):
        for q in range(n):
            for j in range(no):
                for s in range(n):
                    t2[i, q, j, s] = np.dot(c_occ[:, j], t1[i, q, :, s])
    t3 = np.zeros((no, nv, no, n))
    for i in range(no):
        for a in range(nv):
            for j in range(no):
                for s in range(n):
                    t3[i, a, j, s] = np.dot(c_virt[:, a], t2[i, :, j, s])
    iajb = np.zeros((no, nv, no, nv))
    for i in range(no):
        for a in range(nv):
            for j in range(no):
                for b in range(nv):
                    iajb[i, a, j, b] = np.dot(c_virt[:, b], t3[i, a, j, :])
    return iajb


def transform_einsum(eri, c_occ, c_virt):
    t1 = np.einsum("pi,pqrs->iqrs", c_occ, eri, optimize=True)
    t2 = np.einsum("rj,iqrs->iqjs", c_occ, t1, optimize=True)
    t3 = np.einsum("qa,iqjs->iajs", c_virt, t2, optimize=True)
    return np.einsum("sb,iajs->iajb", c_virt, t3, optimize=True)


def energy_loop(iajb, e_occ, e_virt):
    no, nv = len(e_occ), len(e_virt)
    e = 0.0
    for i in range(no):
        for j in range(no):
            for a in range(nv):
                for b in range(nv):
                    x = iajb[i, a, j, b]
                    direct = 2.0 * x * x
                    exch = -x * iajb[i, b, j, a]
                    e += (direct + exch) / (e_occ[i] + e_occ[j] - e_virt[a] - e_virt[b])
    return e


def energy_einsum(iajb, e_occ, e_virt):
    den = (e_occ[:, None, None, None] - e_virt[None, :, None, None]
           + e_occ[None, None, :, None] - e_virt[None, None, None, :])
    return float(np.sum((2.0 * iajb * iajb - iajb * iajb.transpose(0, 3, 2, 1)) / den))


def mp2_corr(eri, c, eps, n_occ, loops=False):
    """E_corr for AO ERIs `eri`, AOxMO coefficients `c`, orbital energies `eps` (mp2.py:52-64 split)."""
    c_occ, c_virt = c[:, :n_occ], c[:, n_occ:]
    e_occ, e_virt = eps[:n_occ], eps[n_occ:]
    if loops:
        iajb = transform_loop(eri, c_occ, c_virt)
        return energy_loop(iajb, e_occ, e_virt), iajb
    iajb = transform_einsum(eri, c_occ, c_virt)
    return energy_einsum(iajb, e_occ, e_virt), iajb


def mp2_corr_synth_chunked(n, seed, c, eps, n_occ, chunk=512):
    """Config-size restatement (N = 222 needs no 19 GB tensor): the four quarter transforms of mp2.py:123-170 as GEMMs
    over chunks of canonical (r, s) pairs of the lazily generated synthetic tensor, then the energy sum of :240-254.
    Returns (E_corr, iajb [n_occ, n_virt, n_occ, n_virt])."""
    from . import cjk
    c_occ, c_virt = np.ascontiguousarray(c[:, :n_occ]), np.ascontiguousarray(c[:, n_occ:])
    nv = n - n_occ
    npair = n * (n + 1) // 2
    y = np.empty((npair, n_occ, nv))                    # (ia|rs) for canonical rs
    for rs0 in range(0, npair, chunk):
        rs1 = min(npair, rs0 + chunk)
        x = cjk.synth_slabs_rs(n, seed, rs0, rs1)       # [c, p, q]
        t1 = np.matmul(c_occ.T[None], x)                # [c, i, q]    temp1 (:123-130)
        y[rs0:rs1] = np.matmul(t1, c_virt[None])        # [c, i, a]    (contract q with C_virt, :149-156)
    r, s = np.tril_indices(n)
    iajb = np.empty((n_occ, nv, n_occ, nv))
    for i in range(n_occ):
        z = np.zeros((nv, n, n))
        z[:, r, s] = y[:, i, :].T
        z[:, s, r] = y[:, i, :].T
        w = np.matmul(c_occ.T[None], z)                 # [a, j, s]    (:136-143)
        iajb[i] = np.matmul(w, c_virt[None])            # [a, j, b]    (:162-170)
    return energy_einsum(iajb, eps[:n_occ], eps[n_occ:]), iajb
