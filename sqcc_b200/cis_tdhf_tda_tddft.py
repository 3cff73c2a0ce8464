"""Closed-shell CIS with the reference's Calculator API (cis_tdhf_tda_tddft.py:19-265).

H[ia,jb] = delta_ij delta_ab (e_a - e_i) + 2 (ia|jb) - (ij|ab)   (cis_tdhf_tda_tddft.py:234-246)

The two AO->MO transforms (:128-165 and :184-211) run on the device (``FockEngine.ao2mo`` -> ``sqcc_ao2mo``, the
same DMMA quarter-transform kernels as MP2); the dense eigensolve stays on the host (np.linalg.eigh, :256) as in
the reference.  The n_occ*n_virt <= 1000 guard (:75-77) is lifted.
"""
import numpy as np

from .backend import FockEngine

AU_TO_EV = 27.21162


class Calculator:
    def __init__(self, scf_object):
        self.scf = scf_object

    def cis(self):
        scf = self.scf
        if scf.spin_multiplicity != 1:
            raise NotImplementedError("Current CIS only supports the singlet state.")
        n_occ = int(scf.num_electrons / 2)
        c_occ, c_virt = scf.mo_coefficients[:, :n_occ], scf.mo_coefficients[:, n_occ:]
        e_occ, e_virt = scf.mo_energies[:n_occ], scf.mo_energies[n_occ:]
        n_virt = c_virt.shape[1]
        engine = getattr(scf, "engine", None)
        if engine is None:
            if getattr(scf, "ao_electron_repulsion_integral", None) is None:
                raise RuntimeError("the SCF object carries neither an engine nor the AO ERI tensor "
                                   "(direct_ingest on several ranks keeps only shards): nothing to transform")
            engine = FockEngine()
            engine.attach_eri_full(scf.ao_electron_repulsion_integral)
        print("Transforming AO integrals to MO basis for (ia|jb) and (ij|ab) on the GPU...")
        iajb = engine.ao2mo(c_occ, c_virt, c_occ, c_virt)     # [i,a,j,b]
        ijab = engine.ao2mo(c_occ, c_occ, c_virt, c_virt)     # [i,j,a,b]
        dim = n_occ * n_virt
        hamiltonian = 2.0 * iajb.reshape(dim, dim) - ijab.transpose(0, 2, 1, 3).reshape(dim, dim)
        hamiltonian[np.diag_indices(dim)] += (e_virt[None, :] - e_occ[:, None]).reshape(dim)
        print("Diagonalizing CIS Hamiltonian matrix...")
        excitation_energies, _ = np.linalg.eigh(hamiltonian)
        print("")
        print("The number of excited states is %s." % len(excitation_energies))
        print("CIS excitation energies (eV):")
        print(*excitation_energies * AU_TO_EV)
        return excitation_energies
