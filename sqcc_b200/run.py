"""run.py -- command-line driver with the reference's entry points (run.py:26, :77, :107, :137, :241).

  cd <directory holding sqc.conf>;  python -m sqcc_b200.run        (or: python /path/to/sqcc_b200/run.py)

Reads ./sqc.conf exactly like the reference, runs HF/KS-DFT and, on request, MP2 / CIS, logging the same
``run_* elapsed_time`` lines.  The electron-density plots of the reference (visualizer.py, matplotlib) are out of
this package's scope: requesting them raises a clear error instead of silently skipping work.
"""
import logging

import numpy as np
import os
import sys
import time

if __package__ in (None, ""):   # executed as a script: make the package importable
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    __package__ = "sqcc_b200"

from . import cis_tdhf_tda_tddft, hf_ksdft, mp2  # noqa: E402
from . import input as conf  # noqa: E402
from . import interface_psi4 as ipsi4  # noqa: E402

logging.basicConfig(level=logging.INFO, format="%(message)s")


def _timed(label, fn):
    start = time.perf_counter()
    result = fn()
    logging.info(f"{label} elapsed_time: {time.perf_counter() - start:.6f} [sec]\n")
    return result


def run_hf_ksdft():
    """(scf_object, flag_mp2, flag_cis), like run.py:26-75."""
    flags = {}

    def work():
        (mol_xyz, zs, coords, basis, functional, charge, mult, treatment,
         flags["cis"], flags["mp2"], flag_qmmm, mm_file) = conf.get_calc_params()
        mm_coords, mm_charges = conf.read_mm_charges(mm_file) if flag_qmmm else (None, None)
        scf = hf_ksdft.Calculator(mol_xyz, zs, coords, basis, functional, charge, mult, treatment,
                                  mm_coords, mm_charges)
        scf.scf()
        scf.save_mo_data_to_json("mo_data.json")
        return scf

    scf = _timed("run_hf_ksdft", work)
    return scf, flags["mp2"], flags["cis"]


def run_cis(scf_object):
    return _timed("run_cis", lambda: cis_tdhf_tda_tddft.Calculator(scf_object).cis())


def run_mp2(scf_object):
    return _timed("run_mp2", lambda: mp2.Calculator(scf_object).mp2())


def run_analysis(scf_object, analysis_params):
    """run.py:137-237.  The numerical part of the density analysis -- rho on the 1000 x 1000 sampling planes,
    visualizer.py:49-87 -- runs on the device (sqcc_b200/visualizer.py) and is written to density_plane_<axis>.npz;
    the matplotlib figures themselves (visualizer.py:200-618) and the density-difference mode that re-reads two
    mo_data.json files are not part of the Fock-build path."""
    if analysis_params.get("electron_density_difference"):
        raise NotImplementedError(
            "sqcc_b200 accelerates the SCF/MP2/CIS path and the density sampling only; run the reference visualizer on "
            "the two mo_data.json files for: electron_density_difference")
    if analysis_params.get("electron_density") and scf_object is not None:
        from .visualizer import DensityVisualizer

        def work():
            vis = DensityVisualizer(scf_object)
            for axis in "xyz":
                rho, gu, gv = vis.density_on_plane(axis=axis, num_points=1000)
                np.savez_compressed(f"density_plane_{axis}.npz", rho=rho, u=gu, v=gv)
        _timed("run_analysis", work)


def main():
    print("=" * 60)
    print("SQCC: Simple Quantum Chemistry Calculator (B200 Fock-build backend)")
    print("=" * 60)
    print()
    print("Psi4 version: %s" % ipsi4.get_psi4_version())
    print()
    analysis_params = conf.get_analysis_params()
    if analysis_params.get("electron_density_difference") and not analysis_params.get("electron_density"):
        run_analysis(None, analysis_params)
        return
    scf_result, flag_mp2, flag_cis = run_hf_ksdft()
    if flag_mp2:
        run_mp2(scf_result)
    if flag_cis:
        run_cis(scf_result)
    if any(analysis_params.values()):
        run_analysis(scf_result, analysis_params)


if __name__ == "__main__":
    main()
