"""Import the UNMODIFIED reference modules from /root/reference with a stub interface_psi4.

TEST INFRASTRUCTURE, and only usable in the build container: ``/root/reference`` does not exist on the
GPU box.  Used by ``oracle/make_golden.py`` (to generate ``tests/golden``) and by the CPU tests that
are skipped when the reference tree is absent.  Nothing is copied: the reference sources are executed
where they lie (hf_ksdft.py imports ``interface_psi4`` at :12, which imports psi4 -- not installed --
so a stub module is registered first).
"""
import contextlib
import importlib
import io
import os
import sys

REFERENCE_DIR = "/root/reference/code/python3"
_REF_MODULES = ("hf_ksdft", "ksdft_functional", "mp2", "cis_tdhf_tda_tddft", "interface_psi4")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "hf_ksdft.py"))


@contextlib.contextmanager
def reference_modules(stub=None):
    """Context manager yielding a dict of freshly imported reference modules."""
    from . import sgauss
    if not available():
        raise RuntimeError("reference tree not present")
    saved = {k: sys.modules.get(k) for k in _REF_MODULES}
    old_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True          # the reference dir is read-only
    sys.path.insert(0, REFERENCE_DIR)
    try:
        for k in _REF_MODULES:
            sys.modules.pop(k, None)
        sys.modules["interface_psi4"] = stub if stub is not None else sgauss.stub_module()
        mods = {k: importlib.import_module(k) for k in ("hf_ksdft", "ksdft_functional", "mp2",
                                                        "cis_tdhf_tda_tddft")}
        yield mods
    finally:
        sys.path.remove(REFERENCE_DIR)
        sys.dont_write_bytecode = old_flag
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_reference_scf(mol, basis, functional, charge, mult, treatment, mm=None, quiet=True):
    """Run the unmodified reference SCF; returns (calculator, per-iteration energies)."""
    mol_xyz, zs, coords = mol
    with reference_modules() as mods:
        calc = mods["hf_ksdft"].Calculator(mol_xyz, zs, coords, basis, functional, charge, mult,
                                           treatment, *(mm or (None, None)))
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            calc.scf()
        out = buf.getvalue()
    energies = [float(line.split()[3]) for line in out.splitlines() if line.startswith("SCF step")]
    if not quiet:
        print(out)
    calc.converged = "SCF converged!" in out
    return calc, energies


def run_reference_mp2(scf_calc):
    with reference_modules() as mods:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            e = mods["mp2"].Calculator(scf_calc).mp2()
    return e


def run_reference_cis(scf_calc):
    with reference_modules() as mods:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            e = mods["cis_tdhf_tda_tddft"].Calculator(scf_calc).cis()
    return e
