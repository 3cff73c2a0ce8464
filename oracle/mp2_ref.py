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
