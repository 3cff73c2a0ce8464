"""RHF / UHF / RKS-LDA / UKS-LDA SCF driver with the reference's Calculator API (hf_ksdft.py:16-821).

Same constructor signature, ``scf()`` / ``save_mo_data_to_json()`` methods, result attributes, printed lines
("SCF step k:  E Hartree", "SCF converged!") and numerics conventions as the reference: plain Roothaan iteration,
core-Hamiltonian guess, symmetric orthogonalisation, |dE| < 1e-9 after the first step, at most 1000 steps,
F = H + J - K/2 (RHF, D carries the factor 2), F_s = H + J(Da+Db) - K(D_s) (UHF), F = H + J + V_xc (KS, no exact
exchange), Angstrom -> Bohr with 1/0.52917721067.

What changed is where the per-iteration contractions run: the loop nests of hf_ksdft.py:426-440, :456-476,
:483-495, :510-535 and :634-646 are single calls into the B200 engine (``FockEngine.jk`` / ``.lda_vxc``).  The
O(N^3) host steps (eigh, X^T F X, density) stay in NumPy exactly as in the reference, so only the rounding of
J/K/V_xc differs (<= 1e-11) and iteration counts are reproduced.  The reference's size guards for KS-DFT
(num_ao > 120, hf_ksdft.py:326-329) are lifted: they only protected its Python loops.
"""
import json
import os

import numpy as np

from . import interface_psi4 as ipsi4
from .backend import FockEngine

ANG_TO_BOHR = 1 / 0.52917721067
MAX_SCF_ITER = 1000
ENERGY_TOL = 1.0e-9


class Calculator:
    def __init__(self, mol_xyz, nuclear_numbers, geom_coordinates, basis_set_name, ksdft_functional_name,
                 molecular_charge, spin_multiplicity, spin_orbital_treatment, mm_coords=None, mm_charges=None,
                 engine=None, device_scf=None, direct_ingest=None):
        self.mol_xyz = mol_xyz
        self.nuclear_numbers = nuclear_numbers
        self.geom_coordinates = geom_coordinates
        self.basis_set_name = basis_set_name
        self.ksdft_functional_name = ksdft_functional_name
        self.num_electrons = np.sum(nuclear_numbers) - molecular_charge
        self.spin_multiplicity = spin_multiplicity
        self.spin_orbital_treatment = spin_orbital_treatment.lower()
        if self.spin_multiplicity > 1 and self.spin_orbital_treatment == "restricted":
            raise ValueError(
                f"Cannot use restricted calculation (spin_orbital_treatment='restricted') for open-shell system "
                f"with spin_multiplicity={self.spin_multiplicity}. Use spin_orbital_treatment='unrestricted'.")
        self.mm_coords = mm_coords
        self.mm_charges = mm_charges
        self.flag_qmmm = mm_coords is not None and mm_charges is not None
        if self.flag_qmmm:
            print("=" * 60)
            print("QM/MM calculation enabled")
            print(f"Number of MM point charges: {len(self.mm_charges)}")
            print(f"Total MM charge: {np.sum(self.mm_charges):.6f}")
            print("Embedding method: Numerical integration on DFT grids")
            print("=" * 60)
            print()
        name = None if ksdft_functional_name is None else str(ksdft_functional_name).lower()
        if name in (None, "none"):
            self.flag_ksdft = False
        elif name == "lda":
            self.flag_ksdft = True
        else:
            raise NotImplementedError("Only the LDA exchange functional is implemented.")
        self.engine = engine          # optional pre-built FockEngine (multi-GPU runs pass their rank's engine)
        # device_scf: keep D / J / K / F / C on the device between Fock builds (csrc/scf.cu, cuSOLVER eigh) instead of
        # the reference's host NumPy steps; default off (exact host path), or SQCC_DEVICE_SCF=1 in the environment
        self.device_scf = (os.environ.get("SQCC_DEVICE_SCF", "0") == "1") if device_scf is None else bool(device_scf)
        # direct_ingest: take the ERIs shell quartet by shell quartet (provider.iter_eri_shell_quartets) straight into
        # the packed device shard instead of through the full [N,N,N,N] host tensor; default off, or SQCC_DIRECT_INGEST=1
        self.direct_ingest = ((os.environ.get("SQCC_DIRECT_INGEST", "0") == "1") if direct_ingest is None
                              else bool(direct_ingest))
        self._eri_host = None
        self.scf_iterations = 0
        self.scf_converged = False
        self.scf_energy_trace = []

    @property
    def ao_electron_repulsion_integral(self):
        """Full [N,N,N,N] tensor (result attribute of hf_ksdft.py:650-667).  With direct ingestion it only exists
        packed on the device and is expanded on first access."""
        if self._eri_host is None and self.engine is not None and self.engine.world == 1 and self.engine.n:
            self._eri_host = self.engine.eri_full()
        return self._eri_host

    @ao_electron_repulsion_integral.setter
    def ao_electron_repulsion_integral(self, value):
        self._eri_host = value

    # ------------------------------------------------------------------ host steps (unchanged maths)
    @staticmethod
    def solve_one_electron_problem(orthogonalizer, fock_matrix):
        """eigh(X^H F X)  (hf_ksdft.py:97-123)."""
        transformed = np.matmul(np.matmul(np.transpose(orthogonalizer.conjugate()), fock_matrix), orthogonalizer)
        return np.linalg.eigh(transformed)

    def _occupations(self):
        """(n_alpha, n_beta) from the multiplicity (hf_ksdft.py:178-181)."""
        excess = int(self.spin_multiplicity - 1)
        n_beta = (self.num_electrons - excess) // 2
        return n_beta + excess, n_beta

    def calc_density_matrix_in_ao_basis(self, mo_coefficients):
        """D = 2 C_occ C_occ^T (restricted) or [Ca Ca^T, Cb Cb^T] (unrestricted)  (hf_ksdft.py:126-201)."""
        print(self.spin_orbital_treatment)
        if self.spin_orbital_treatment == "restricted":
            if self.num_electrons % 2 != 0:
                raise ValueError(
                    f"Restricted calculation requires even number of electrons, but got {self.num_electrons}. "
                    f"Use spin_multiplicity > 1 for systems with odd number of electrons.")
            occ = mo_coefficients[:, :int(self.num_electrons / 2)]
            return 2.0 * np.matmul(occ, np.transpose(occ))
        if (self.num_electrons % 2) != (self.spin_multiplicity + 1) % 2:
            raise ValueError(
                f"Inconsistent spin_multiplicity={self.spin_multiplicity} for {self.num_electrons} electrons. "
                f"Even electrons require odd spin_multiplicity (1,3,5,...), odd electrons require even "
                f"spin_multiplicity (2,4,6,...).")
        n_alpha, n_beta = self._occupations()
        num_ao = mo_coefficients.shape[1]
        density = np.zeros((2, num_ao, num_ao))
        ca, cb = mo_coefficients[0, :, :n_alpha], mo_coefficients[1, :, :n_beta]
        density[0] = np.matmul(ca, np.transpose(ca))
        density[1] = np.matmul(cb, np.transpose(cb))
        return density

    @staticmethod
    def calc_nuclei_repulsion_energy(coordinates, charges):
        """sum_{i<j} Z_i Z_j / r_ij, coordinates in Angstrom (hf_ksdft.py:203-236)."""
        energy = 0.0
        for i in range(len(coordinates)):
            for j in range(i + 1, len(coordinates)):
                energy += charges[i] * charges[j] / np.linalg.norm((coordinates[i] - coordinates[j]) * ANG_TO_BOHR)
        print(f"Nuclear repulsion energy: {energy:.12f} Hartree")
        return energy

    @staticmethod
    def calc_qmmm_nuclear_interaction(qm_coordinates, qm_charges, mm_coordinates, mm_charges):
        """sum_iA Z_i q_A / r_iA (hf_ksdft.py:238-272)."""
        energy = 0.0
        for i in range(len(qm_coordinates)):
            for a in range(len(mm_coordinates)):
                energy += qm_charges[i] * mm_charges[a] / np.linalg.norm(
                    (qm_coordinates[i] - mm_coordinates[a]) * ANG_TO_BOHR)
        return energy

    def _agree(self, value):
        """Multi-GPU runs execute this driver once per rank.  Every rank must take the same |dE| < 1e-9 decision
        (a rank that stops alone would leave the others waiting in the J/K exchange), so the energy of rank 0 is the
        one all ranks use: LDA quantities are accumulated with floating-point atomics and may differ in the last
        bits between ranks."""
        eng = self.engine
        if eng is None or eng.world == 1:
            return value
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return value
        t = torch.tensor([float(value)], dtype=torch.float64, device=eng.device)
        dist.broadcast(t, 0, group=eng.group)
        return float(t.item())

    # ------------------------------------------------------------------ SCF
    def scf(self):
        restricted = self.spin_orbital_treatment == "restricted"
        proc = ipsi4.Psi4Interface(self.mol_xyz, self.nuclear_numbers, self.geom_coordinates, self.basis_set_name,
                                   self.ksdft_functional_name)
        self.psi4_interface = proc
        kinetic = proc.ao_kinetic_integral()
        nuclear = proc.ao_nuclear_attraction_integral()
        overlap = proc.ao_overlap_integral()
        num_ao = len(overlap)

        # the packed ERI tensor goes to HBM once, before the loop (replaces the host-resident N^4 operand)
        engine = self.engine if self.engine is not None else FockEngine()
        self.engine = engine
        if self.direct_ingest and hasattr(proc, "iter_eri_shell_quartets"):
            eri = None       # never built on the host; ao_electron_repulsion_integral expands the device copy on demand
            engine.attach_eri_blocks(num_ao, proc.iter_eri_shell_quartets())
        else:
            eri = proc.ao_electron_repulsion_integral()
            engine.attach_eri_full(eri)

        grids = weights = None
        if self.flag_ksdft or self.flag_qmmm:
            grids, weights = proc.generate_numerical_integration_grids_and_weights()
            engine.attach_grid(proc.generate_ao_values_at_grids(grids), weights)

        core_hamiltonian = kinetic + nuclear
        if self.flag_qmmm:
            # V_ext[p,q] = sum_n w_n phi_np v_MM(r_n) phi_nq  (interface_psi4.py:387-403) on the device
            core_hamiltonian += engine.grid_quadrature(proc.mm_potential_at_grids(self.mm_coords, self.mm_charges, grids))
            print("External potential added to core Hamiltonian")
            print()

        # S^-1/2 = U s^-1/2 U^T  (hf_ksdft.py:362-378)
        s_val, s_vec = np.linalg.eigh(overlap)
        for i in range(len(s_val)):
            s_val[i] = s_val[i] ** (-0.5)
        orthogonalizer = np.matmul(np.matmul(s_vec, np.diag(s_val)), np.transpose(s_vec.conjugate()))

        # core guess (hf_ksdft.py:384-399)
        orbital_energies, c_prime = Calculator.solve_one_electron_problem(orthogonalizer, core_hamiltonian)
        if restricted:
            mo_coefficients = np.matmul(orthogonalizer, c_prime)
        else:
            mo_coefficients = np.zeros((2, num_ao, num_ao))
            mo_coefficients[0] = np.matmul(orthogonalizer, c_prime)
            mo_coefficients[1] = np.matmul(orthogonalizer, c_prime)
        density = self.calc_density_matrix_in_ao_basis(mo_coefficients)

        nuclear_repulsion = Calculator.calc_nuclei_repulsion_energy(self.geom_coordinates, self.nuclear_numbers)
        if self.flag_qmmm:
            e_qmmm = Calculator.calc_qmmm_nuclear_interaction(self.geom_coordinates, self.nuclear_numbers,
                                                              self.mm_coords, self.mm_charges)
            nuclear_repulsion += e_qmmm
            print(f"QM/MM nuclear interaction energy: {e_qmmm:.10f} Hartree")
            print()

        self.scf_energy_trace = []
        self.scf_converged = False
        alpha_energies = beta_energies = None
        total_energy = old_total_energy = None
        if self.device_scf:
            # same loop with every per-iteration matrix resident in HBM: only E_elec comes back each step
            n_a, n_b = (int(self.num_electrons / 2),) * 2 if restricted else self._occupations()
            engine.scf_setup(core_hamiltonian, orthogonalizer, 1 if restricted else 2, n_a, n_b, self.flag_ksdft)
            engine.scf_set_density(density)
            for step in range(MAX_SCF_ITER):
                total_energy = self._agree(engine.scf_fock_energy() + nuclear_repulsion)
                self.scf_energy_trace.append(float(total_energy))
                print("SCF step %s: " % str(step + 1), total_energy, "Hartree")
                if step > 0 and abs(old_total_energy - total_energy) < ENERGY_TOL:
                    print("SCF converged!")
                    self.scf_converged = True
                    break
                old_total_energy = total_energy
                engine.scf_update()
            d_dev, c_dev, eps_dev = engine.scf_get()
            if restricted:
                density, mo_coefficients, orbital_energies = d_dev[0], c_dev[0], eps_dev[0]
            else:
                density, mo_coefficients = d_dev, c_dev
                alpha_energies, beta_energies = eps_dev[0], eps_dev[1]
        for step in range(0 if self.device_scf else MAX_SCF_ITER):
            # --- device: V_xc / E_xc (KS) and J (/K) for the current density
            xc_energy = 0.0
            vxc = None
            if self.flag_ksdft:
                vxc, xc_energy, _ = engine.lda_vxc(density)
            coulomb, exchange = engine.jk(density, want_k=not self.flag_ksdft)

            # --- host: Fock matrices and electronic energy (hf_ksdft.py:498-504, :527-548, :552-587)
            if restricted:
                if self.flag_ksdft:
                    fock = core_hamiltonian + coulomb + vxc
                    electronic = 0.5 * np.sum(density * (2.0 * core_hamiltonian + coulomb)) + xc_energy
                else:
                    fock = core_hamiltonian + coulomb - 0.5 * exchange
                    electronic = 0.5 * np.sum(density * (core_hamiltonian + fock))
            else:
                if self.flag_ksdft:
                    fock_a = core_hamiltonian + coulomb + vxc[0]
                    fock_b = core_hamiltonian + coulomb + vxc[1]
                    electronic = 0.5 * np.sum((density[0] + density[1]) * (2.0 * core_hamiltonian + coulomb))
                    electronic += xc_energy
                else:
                    fock_a = core_hamiltonian + coulomb - exchange[0]
                    fock_b = core_hamiltonian + coulomb - exchange[1]
                    electronic = 0.5 * np.sum((density[0] + density[1]) * core_hamiltonian)
                    electronic += 0.5 * np.sum(density[0] * fock_a)
                    electronic += 0.5 * np.sum(density[1] * fock_b)
            total_energy = self._agree(electronic + nuclear_repulsion)
            self.scf_energy_trace.append(float(total_energy))
            print("SCF step %s: " % str(step + 1), total_energy, "Hartree")
            if step > 0 and abs(old_total_energy - total_energy) < ENERGY_TOL:
                print("SCF converged!")
                self.scf_converged = True
                break
            old_total_energy = total_energy

            # --- host: new orbitals and density (hf_ksdft.py:604-625)
            if restricted:
                orbital_energies, c_prime = Calculator.solve_one_electron_problem(orthogonalizer, fock)
                mo_coefficients = np.matmul(orthogonalizer, c_prime)
            else:
                alpha_energies, ca = Calculator.solve_one_electron_problem(orthogonalizer, fock_a)
                beta_energies, cb = Calculator.solve_one_electron_problem(orthogonalizer, fock_b)
                mo_coefficients = np.zeros((2, num_ao, num_ao))
                mo_coefficients[0] = np.matmul(orthogonalizer, ca)
                mo_coefficients[1] = np.matmul(orthogonalizer, cb)
            density = self.calc_density_matrix_in_ao_basis(mo_coefficients)
        self.scf_iterations = len(self.scf_energy_trace)

        # result hand-off (hf_ksdft.py:650-667)
        self.density_matrix_in_ao_basis = density
        self.num_ao = num_ao
        if restricted:
            self.mo_energies = orbital_energies
        else:
            if alpha_energies is None:   # converged before the first diagonalisation cannot happen (step > 0)
                alpha_energies = beta_energies = orbital_energies
            self.mo_energies = np.array([alpha_energies, beta_energies])
        self.mo_coefficients = mo_coefficients
        self.scf_energy = total_energy
        self.ao_electron_repulsion_integral = eri
        self.basis_set_object = proc.psi4_basis_set
        self._print_mulliken(density, overlap, proc.get_basis_atomic_affiliation())

    def _print_mulliken(self, density, overlap, affiliation):
        """Mulliken charges from diag(P S) (hf_ksdft.py:670-721); printed only."""
        total = density if self.spin_orbital_treatment == "restricted" else density[0] + density[1]
        charges = np.array(self.nuclear_numbers, dtype=float)
        np.subtract.at(charges, affiliation, np.diag(np.matmul(total, overlap)))
        print("Mulliken atomic charges (|e|):", *charges)
        if self.spin_multiplicity != 1:
            spin = np.zeros(len(self.nuclear_numbers))
            np.add.at(spin, affiliation, np.diag(np.matmul(density[0] - density[1], overlap)))
            print("Mulliken atomic spin charges (|e|):", *spin)

    # ------------------------------------------------------------------ mo_data.json (hf_ksdft.py:724-821)
    def save_mo_data_to_json(self, output_file="mo_data.json"):
        basis = self.basis_set_object
        ao_info = []
        for shell_idx in range(basis.nshell()):
            shell = basis.shell(shell_idx)
            for func_idx in range(shell.nfunction):
                ao_info.append({
                    "atom_index": int(shell.ncenter),
                    "atom_number": int(self.nuclear_numbers[shell.ncenter]),
                    "atom_coords": self.geom_coordinates[shell.ncenter].tolist(),
                    "shell_type": "SPDFGHI"[shell.am],
                    "angular_momentum": int(shell.am),
                    "function_index": int(func_idx),
                })
        common = {
            "num_electrons": int(self.num_electrons), "num_ao": int(self.num_ao), "num_mo": int(self.num_ao),
            "basis_set": self.basis_set_name,
            "functional": self.ksdft_functional_name if self.ksdft_functional_name else "HF",
            "ao_info": ao_info,
        }
        if self.spin_orbital_treatment == "restricted":
            n_occ = self.num_electrons // 2
            data = {"calculation_type": "restricted", **common,
                    "mo_energies": self.mo_energies.tolist(), "mo_coefficients": self.mo_coefficients.tolist(),
                    "occupation_numbers": [2.0] * n_occ + [0.0] * (self.num_ao - n_occ)}
            order = ["calculation_type", "num_electrons", "num_ao", "num_mo", "basis_set", "functional", "ao_info",
                     "mo_energies", "mo_coefficients", "occupation_numbers"]
        else:
            n_alpha, n_beta = self._occupations()
            data = {"calculation_type": "unrestricted", **common, "num_alpha": int(n_alpha), "num_beta": int(n_beta)}
            for tag, s, n_s in (("alpha", 0, n_alpha), ("beta", 1, n_beta)):
                data[tag] = {"mo_energies": self.mo_energies[s].tolist(),
                             "mo_coefficients": self.mo_coefficients[s].tolist(),
                             "occupation_numbers": [1.0] * n_s + [0.0] * (self.num_ao - n_s)}
            order = ["calculation_type", "num_electrons", "num_alpha", "num_beta", "num_ao", "num_mo", "basis_set",
                     "functional", "ao_info", "alpha", "beta"]
        with open(output_file, "w") as fh:
            json.dump({k: data[k] for k in order}, fh, indent=2)
        print(f"MO data saved to {output_file}")
