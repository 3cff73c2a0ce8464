"""LDA-exchange grid quadrature oracle (numpy).  TEST INFRASTRUCTURE.

Loop forms follow the reference statement for statement; GEMM forms are validated against them.
  rho   : hf_ksdft.py:426-428 / :634-636 (RKS), :434-440 / :640-646 (UKS)
  v_x   : ksdft_functional.py:41-65 (unpolarised), :98-134 (spin-polarised)
  E_x   : ksdft_functional.py:68-95, :137-177
  V_xc  : hf_ksdft.py:456-461 (RKS), :469-476 (UKS)
"""
import numpy as np


def rho_loop(phi, d):
    g = phi.shape[0]
    rho = np.zeros(g)
    for n in range(g):
        temp = np.matmul(phi[n, :], d)
        rho[n] = np.dot(temp, phi[n, :])
    return rho


def rho_gemm(phi, d):
    return np.einsum("np,np->n", phi @ d, phi)


def lda_potential(rho):
    c = -(3.0 / np.pi) ** (1.0 / 3.0)
    return c * (rho ** (1.0 / 3.0))


def lda_energy(rho, w):
    c = (-3.0 / 4.0) * ((3.0 / np.pi) ** (1.0 / 3.0))
    return c * np.sum(w * rho ** (4.0 / 3.0))


def lda_potential_spinpol(rho_a, rho_b):
    c = -(2.0 ** (1.0 / 3.0)) * (3.0 / np.pi) ** (1.0 / 3.0)
    return c * (rho_a ** (1.0 / 3.0)), c * (rho_b ** (1.0 / 3.0))


def lda_energy_spinpol(rho_a, rho_b, w):
    c = (-3.0 / 4.0) * ((3.0 / np.pi) ** (1.0 / 3.0)) * (2.0 ** (1.0 / 3.0))
    return c * (np.sum(w * rho_a ** (4.0 / 3.0)) + np.sum(w * rho_b ** (4.0 / 3.0)))


def vxc_loop(phi, w, v):
    n = phi.shape[1]
    out = np.zeros((n, n))
    for p in range(n):
        for q in range(n):
            out[p, q] = np.sum(w * phi[:, p] * v * phi[:, q])
    return out


def vxc_gemm(phi, w, v):
    return phi.T @ (phi * (w * v)[:, None])


def lda_vxc(phi, w, d, loops=False):
    """Rows G+H+I fused, reference semantics.

    d [N,N]   -> (Vxc [N,N],   Exc, rho [G])      RKS
    d [2,N,N] -> (Vxc [2,N,N], Exc, rho [2,G])    UKS
    """
    f_rho = rho_loop if loops else rho_gemm
    f_v = vxc_loop if loops else vxc_gemm
    if d.ndim == 2:
        rho = f_rho(phi, d)
        return f_v(phi, w, lda_potential(rho)), lda_energy(rho, w), rho
    ra, rb = f_rho(phi, d[0]), f_rho(phi, d[1])
    va, vb = lda_potential_spinpol(ra, rb)
    return (np.stack([f_v(phi, w, va), f_v(phi, w, vb)]),
            lda_energy_spinpol(ra, rb, w), np.stack([ra, rb]))
