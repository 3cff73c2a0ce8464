"""Slater (LDA) exchange, host-side point functions with the reference's names (ksdft_functional.py:16-209).

The SCF driver does NOT call these: rho, v_x, E_x and V_xc are produced by one fused device pass
(``FockEngine.lda_vxc`` -> ``sqcc_lda_vxc``).  They are kept so that code written against the reference module
keeps working (the kernels are unused by the reference itself).  Exchange only, no correlation.
"""
import numpy as np

_CX = (3.0 / np.pi) ** (1.0 / 3.0)
_TWO13 = 2.0 ** (1.0 / 3.0)


def lda_potential(electron_density_at_grids):
    return -_CX * electron_density_at_grids ** (1.0 / 3.0)


def lda_energy(electron_density_at_grids, grid_weights):
    return (-3.0 / 4.0) * _CX * np.sum(grid_weights * electron_density_at_grids ** (4.0 / 3.0))


def lda_potential_spinpol(rho_alpha, rho_beta):
    c = -_TWO13 * _CX
    return c * rho_alpha ** (1.0 / 3.0), c * rho_beta ** (1.0 / 3.0)


def lda_energy_spinpol(rho_alpha, rho_beta, grid_weights):
    c = (-3.0 / 4.0) * _CX * _TWO13
    return c * (np.sum(grid_weights * rho_alpha ** (4.0 / 3.0)) + np.sum(grid_weights * rho_beta ** (4.0 / 3.0)))


def lda_kernel(electron_density_at_grids):
    return (-3.0 / 4.0) * _CX * electron_density_at_grids ** (1.0 / 3.0)


def lda_kernel_spinpol(rho_alpha, rho_beta):
    c = (-3.0 / 4.0) * _TWO13 * _CX
    return c * rho_alpha ** (1.0 / 3.0), c * rho_beta ** (1.0 / 3.0)
