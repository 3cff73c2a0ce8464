"""AO-integral provider: the same surface as the reference's interface_psi4.Psi4Interface (interface_psi4.py:65-406).

Psi4 still supplies the AO integrals, DFT grids and AO-on-grid values (BASELINE.json north_star); it is imported
lazily so that the package loads on machines without it (this image has no Psi4 -- tests inject a provider with
the same surface through ``hf_ksdft.ipsi4``).  Only what the SCF driver consumes is wrapped.
"""
import numpy as np

ANG_TO_BOHR = 1 / 0.52917721067


def _psi4():
    try:
        import psi4
    except ImportError as err:
        raise ImportError("Psi4 is required to generate AO integrals (conda install psi4); "
                          "alternatively assign a provider module to sqcc_b200.hf_ksdft.ipsi4") from err
    return psi4


def get_psi4_version():
    return _psi4().__version__


class Psi4Interface:
    def __init__(self, mol_xyz, nuclear_numbers, geom_coordinates, basis_set_name, ksdft_functional_name=None):
        psi4 = _psi4()
        rows = mol_xyz.strip().splitlines()
        if rows and rows[0].strip().isdigit():
            rows = rows[2:]           # full XYZ text: drop the count and comment lines
        # keep the input frame: no recentring / reorientation (the grids and MM charges live in it)
        self.mol = psi4.geometry("nocom\nnoreorient\n" + "\n".join(rows) + "\n")
        self.mol.fix_orientation(True)
        self.mol.fix_com(True)
        self.nuclear_numbers = nuclear_numbers
        self.geom_coordinates = geom_coordinates
        self.basis_set_name = basis_set_name
        self.ksdft_functional_name = ksdft_functional_name
        psi4.set_options({"basis": basis_set_name, "dft_radial_points": 230, "dft_spherical_points": 770})
        self.wfn = psi4.core.Wavefunction.build(self.mol, psi4.core.get_global_option("BASIS"))
        self.psi4_basis_set = self.wfn.basisset()
        self.mints = psi4.core.MintsHelper(self.psi4_basis_set)
        self.psi4_dft_functional = psi4.driver.dft.build_superfunctional("svwn", True)[0]

    def _as_array(self, matrix):
        return np.asarray(matrix, dtype="float64")

    def ao_kinetic_integral(self):
        return self._as_array(self.mints.ao_kinetic())

    def ao_nuclear_attraction_integral(self):
        return self._as_array(self.mints.ao_potential())

    def ao_electron_repulsion_integral(self):
        return self._as_array(self.mints.ao_eri())

    def iter_eri_shell_quartets(self):
        """Canonical shell quartets (M >= N, P >= Q, MN >= PQ) as dense blocks (i0, j0, k0, l0, values[nM,nN,nP,nQ])
        from MintsHelper.ao_eri_shell: feeds FockEngine.attach_eri_blocks, so the full [N,N,N,N] tensor of
        ao_electron_repulsion_integral() (interface_psi4.py:177-190) is never materialised.
        Not exercised in this repository's tests (no Psi4 in the image); the provider-independent part is."""
        bs = self.psi4_basis_set
        ns = bs.nshell()
        first = [bs.shell(s).function_index for s in range(ns)]
        size = [bs.shell(s).nfunction for s in range(ns)]
        for m in range(ns):
            for n in range(m + 1):
                for p in range(m + 1):
                    for q in range((p if p < m else n) + 1):
                        block = np.asarray(self.mints.ao_eri_shell(m, n, p, q), dtype="float64")
                        yield first[m], first[n], first[p], first[q], block.reshape(size[m], size[n], size[p], size[q])

    def ao_overlap_integral(self):
        return self._as_array(self.mints.ao_overlap())

    def generate_numerical_integration_grids_and_weights(self, mm_coords=None, mm_charges=None):
        psi4 = _psi4()
        vbase = psi4.core.VBase.build(self.psi4_basis_set, self.psi4_dft_functional, "RV")
        vbase.initialize()
        x, y, z, w = vbase.get_np_xyzw()
        return np.stack([x, y, z], axis=1), w

    def generate_ao_values_at_grids(self, real_space_grids):
        out = np.zeros((len(real_space_grids), self.psi4_basis_set.nbf()))
        for n, point in enumerate(real_space_grids):
            out[n] = self.psi4_basis_set.compute_phi(*point)
        return out

    def get_basis_atomic_affiliation(self):
        bs = self.psi4_basis_set
        return np.array([bs.function_to_center(f) for f in range(bs.nbf())])

    def mm_potential_at_grids(self, mm_coords, mm_charges, real_space_grids):
        """Point-charge potential v(r) = sum_A -q_A / |r - R_A| on the grid (Bohr), skipping |r - R_A| <= 1e-10
        like interface_psi4.py:391-396; the quadrature phi^T (w v) phi of :399-403 runs on the device."""
        v = np.zeros(len(real_space_grids))
        for xyz, q in zip(np.asarray(mm_coords) * ANG_TO_BOHR, mm_charges):
            r = np.linalg.norm(real_space_grids - xyz, axis=1)
            v += np.where(r > 1e-10, -q / np.where(r > 1e-10, r, 1.0), 0.0)
        return v

    def compute_external_potential_integrals(self, mm_coords, mm_charges, grids=None, grid_weights=None):
        """Host form kept for API compatibility (interface_psi4.py:316-406); the SCF driver uses the device path."""
        if grids is None or grid_weights is None:
            grids, grid_weights = self.generate_numerical_integration_grids_and_weights()
        phi = self.generate_ao_values_at_grids(grids)
        v = self.mm_potential_at_grids(mm_coords, mm_charges, grids)
        return phi.T @ (phi * (grid_weights * v)[:, None])
