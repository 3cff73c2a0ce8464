"""Generate tests/golden/* by EXECUTING the unmodified reference (build container only).

  python -m oracle.make_golden

* ref_kernels.npz -- J/K (RHF, UHF), rho / V_xc / E_xc (RKS, UKS) and the MP2 correlation energy produced by the
  reference's own statements: the loop nests are read from /root/reference/code/python3/hf_ksdft.py at run time
  (line ranges below) and exec'd on seeded inputs, nothing is copied into this repository; MP2 runs the
  reference's mp2.Calculator on a duck-typed SCF object.
* ref_scf.json -- total energies per SCF iteration, iteration counts and MP2/CIS results of the reference's
  Calculator.scf() driven through the s-Gaussian stub interface (oracle/sgauss.py) for a set of small systems.
"""
import json
import os
import textwrap
import types

import numpy as np

from . import reference_loader as rl
from . import sgauss, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
HF = os.path.join(rl.REFERENCE_DIR, "hf_ksdft.py")


def _exec_lines(ranges, ns):
    src = open(HF).read().splitlines()
    for lo, hi in ranges:
        code = textwrap.dedent("\n".join(src[lo - 1:hi]))
        exec(compile(code, f"{HF}:{lo}-{hi}", "exec"), ns)
    return ns


def kernels():
    n, g = 7, 300
    eri = synth.synth_eri_full(n, 4242)
    d = synth.synth_density_rhf(n, 3, 1)
    du = synth.synth_density_uhf(n, 4, 2, 2)
    phi, w = synth.synth_grid(g, n, 3)
    out = dict(eri=eri, d=d, du=du, phi=phi, w=w)
    with rl.reference_modules() as mods:
        kdf = mods["ksdft_functional"]
        base = dict(np=np, kdf=kdf, num_ao=n, num_grids=g, ao_electron_repulsion_integral=eri,
                    ao_values_at_grids=phi, weights_grids=w)
        # RHF J :483-487, K :491-495
        ns = _exec_lines([(483, 487), (491, 495)], dict(base, density_matrix_in_ao_basis=d))
        out["j_rhf"], out["k_rhf"] = ns["electron_repulsion_in_Fock_matrix"], ns["exchange_in_Fock_matrix"]
        # UHF J :510-515, Ka :522-526, Kb :531-535
        ns = _exec_lines([(510, 515), (522, 526), (531, 535)], dict(base, density_matrix_in_ao_basis=du))
        out["j_uhf"] = ns["electron_repulsion_in_Fock_matrix"]
        out["k_uhf"] = np.stack([ns["alpha_exchange_in_Fock_matrix"], ns["beta_exchange_in_Fock_matrix"]])
        # RKS rho :422-428, v/Vxc :452-461, Exc :561-562
        ns = _exec_lines([(422, 428), (452, 461), (561, 562)], dict(base, density_matrix_in_ao_basis=d))
        out["rho_rks"], out["vxc_rks"] = ns["electron_density_at_grids"], ns["exchange_correlation_potential_in_Fock_matrix"]
        out["exc_rks"] = ns["exchange_correlation_energy"]
        # UKS rho :432-440, Vxc :465-476, Exc :581-582
        ns = _exec_lines([(432, 440), (465, 476), (581, 582)], dict(base, density_matrix_in_ao_basis=du))
        out["rho_uks"] = np.stack([ns["alpha_density_at_grids"], ns["beta_density_at_grids"]])
        out["vxc_uks"] = np.stack([ns["alpha_xc_matrix"], ns["beta_xc_matrix"]])
        out["exc_uks"] = ns["exchange_correlation_energy"]
        # MP2 through the reference's own class (mp2.py:30-265)
        nm, no = 6, 2
        meri = synth.synth_eri_full(nm, 77)
        c, eps = synth.synth_mo(nm, no, 5)
        scf_like = types.SimpleNamespace(spin_multiplicity=1, num_electrons=2 * no, mo_coefficients=c, mo_energies=eps,
                                         scf_energy=0.0, ao_electron_repulsion_integral=meri)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            out["mp2_ecorr"] = mods["mp2"].Calculator(scf_like).mp2()
        out.update(mp2_eri=meri, mp2_c=c, mp2_eps=eps, mp2_nocc=no)
    np.savez_compressed(os.path.join(OUT, "ref_kernels.npz"), **out)


def h2_stack(nx, ny, nz, bond=0.74, sep=3.0):
    """nx x ny x nz H2 molecules (bond along z): closed shells with a large gap, converges without damping."""
    zs, co = [], []
    for a in range(nx):
        for b in range(ny):
            for c in range(nz):
                zs += [1, 1]
                co += [[a * sep, b * sep, c * (sep + bond)], [a * sep, b * sep, c * (sep + bond) + bond]]
    return sgauss.cluster(zs, co)


# name -> (builder, basis, functional, charge, multiplicity, treatment, post[, (mm_coords_ang, mm_charges)])
def _systems():
    r14 = 1.4 * 0.52917721067
    sq = 0.9
    mm1 = (np.array([[4.0, 0.0, 0.0]]), np.array([-1.1]))          # the point charge of tests/hf_mm/n2_singlet
    mm2 = (np.array([[4.0, 0.0, 0.0], [0.0, 3.0, 2.0]]), np.array([-1.1, 0.6]))
    return {
        # QM/MM (hf_ksdft.py:345-358, :408-414; interface_psi4.py:387-403), mirrors tests/hf_mm/*
        "h6_rhf_mm": (sgauss.h_chain(6), "sto-3g", None, 0, 1, "restricted", (), mm2),
        "h4sq_rks_mm": (sgauss.cluster([1, 1, 1, 1], [[0, 0, 0], [0, 0, 0.74], [0, 2.5, 0], [0, 2.5, 0.74]]), "dz-s",
                        "lda", 0, 1, "restricted", (), mm1),
        "h4_uhf_triplet_mm": (sgauss.h_chain(4), "sto-3g", None, 0, 3, "unrestricted", (), mm1),
        # 64 - 72 basis functions: the production J/K schedule (several tasks, 4 x 4 row-blocks, full and narrow pieces)
        # and the generic N > 64 LDA path under the iteration-count gate
        # (plain Roothaan iteration is fragile on bigger clusters: (H2)_24 and planar 4 x 4 arrangements run into the
        # 1000-iteration cap or converge to a saddle after 75 steps in the reference; these stacks converge in < 20)
        "h2x16_dz_rhf": (h2_stack(2, 4, 2), "dz-s", None, 0, 1, "restricted", ()),
        "h2x12_tz_rhf": (h2_stack(2, 3, 2), "tz-s", None, 0, 1, "restricted", ("mp2",)),
        "h2x12_tz_rks": (h2_stack(2, 3, 2), "tz-s", "lda", 0, 1, "restricted", ()),
        "h2_rhf": (sgauss.cluster([1, 1], [[0, 0, 0], [0, 0, r14]]), "sto-3g", None, 0, 1, "restricted", ()),
        "h6_rhf": (sgauss.h_chain(6), "sto-3g", None, 0, 1, "restricted", ("mp2", "cis")),
        "h6_uhf_singlet": (sgauss.h_chain(6), "sto-3g", None, 0, 1, "unrestricted", ()),
        "h4_uhf_triplet": (sgauss.h_chain(4), "sto-3g", None, 0, 3, "unrestricted", ()),
        "h10_rhf": (sgauss.h_chain(10), "sto-3g", None, 0, 1, "restricted", ("mp2",)),
        "h8_dz_rhf": (sgauss.h_chain(8), "dz-s", None, 0, 1, "restricted", ("mp2",)),
        "heh_plus_rhf": (sgauss.cluster([2, 1], [[0, 0, 0], [0, 0, 0.7743]]), "sto-3g", None, 1, 1, "restricted", ()),
        "h2_rks": (sgauss.cluster([1, 1], [[0, 0, 0], [0, 0, r14]]), "sto-3g", "lda", 0, 1, "restricted", ()),
        "h2_uks_singlet": (sgauss.cluster([1, 1], [[0, 0, 0], [0, 0, r14]]), "sto-3g", "lda", 0, 1, "unrestricted", ()),
        "h4sq_rks": (sgauss.cluster([1, 1, 1, 1], [[0, 0, 0], [0, 0, 0.74], [0, 2.5, 0], [0, 2.5, 0.74]]), "dz-s",
                     "lda", 0, 1, "restricted", ()),
        "h3_uks_doublet": (sgauss.cluster([1, 1, 1], [[0, 0, 0], [0, 0, sq], [0, 0, 2 * sq]]), "dz-s", "lda", 0, 2,
                           "unrestricted", ()),
        "h5_uhf_doublet_cation2": (sgauss.h_chain(5), "dz-s", None, 0, 2, "unrestricted", ()),
    }


def scf_runs():
    res = {}
    import time
    only = os.environ.get("GOLDEN_ONLY")          # comma-separated names: regenerate just these, keep the others
    path = os.path.join(OUT, "ref_scf.json")
    if only and os.path.exists(path):
        res = json.load(open(path))
    for name, spec in _systems().items():
        if only and name not in only.split(","):
            continue
        mol, basis, func, charge, mult, treat, post = spec[:7]
        mm = spec[7] if len(spec) > 7 else None
        t0 = time.perf_counter()
        calc, energies = rl.run_reference_scf(mol, basis, func, charge, mult, treat, mm=mm)
        entry = dict(basis=basis, functional=func, charge=charge, multiplicity=mult, treatment=treat,
                     zs=[int(z) for z in mol[1]], coords=np.asarray(mol[2]).tolist(), mol_xyz=mol[0],
                     energies=energies, iterations=len(energies), converged=bool(calc.converged),
                     scf_energy=float(calc.scf_energy), num_ao=int(calc.num_ao),
                     # wall time of the unmodified reference's Calculator(...).scf() in the build container (8 vCPU Xeon,
                     # NumPy 2.3 / OpenBLAS), integrals from the s-Gaussian stub included: bench.py's SCF wall-time leg
                     reference_scf_seconds=time.perf_counter() - t0)
        if mm is not None:
            entry["mm_coords"] = np.asarray(mm[0]).tolist()
            entry["mm_charges"] = np.asarray(mm[1]).tolist()
        if "mp2" in post:
            entry["mp2_energy"] = float(rl.run_reference_mp2(calc))
        if "cis" in post:
            entry["cis_energies"] = np.asarray(rl.run_reference_cis(calc)).tolist()
        res[name] = entry
        print(name, entry["iterations"], entry["converged"], entry["scf_energy"])
    with open(path, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if not os.environ.get("GOLDEN_ONLY"):
        kernels()
    scf_runs()
