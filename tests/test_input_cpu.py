"""Host-side driver pieces that need no GPU: sqc.conf / xyz / MM parsers (same tuples as the reference's input.py)."""
import glob
import os
import sys
import types

import numpy as np
import pytest

from sqcc_b200 import input as conf

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _in_dir(path, fn):
    cwd = os.getcwd()
    os.chdir(path)
    try:
        return fn()
    finally:
        os.chdir(cwd)


def test_authored_n2_sto3g_config():
    """BASELINE.json configs[0] (N2 RHF STO-3G): the reference ships no such sqc.conf."""
    p = _in_dir(os.path.join(HERE, "data", "n2_rhf_sto3g"), conf.get_calc_params)
    mol_xyz, zs, xyz, basis, func, charge, mult, treat, cis, mp2, qmmm, mm = p
    assert list(zs) == [7, 7] and xyz.shape == (2, 3) and basis == "sto-3g" and func == "None"
    assert (charge, mult, treat, cis, mp2, qmmm, mm) == (0, 1, "restricted", False, True, False, None)
    assert mol_xyz.count("\n") == 2


def test_authored_baseline_configs():
    """Inputs for the other BASELINE.json configs the reference ships no sqc.conf for: O2 triplet
    UHF and UKS-LDA def2-TZVP, benzene RHF+MP2 def2-TZVP.  They need Psi4 to run; here they must parse like any
    reference input (unrestricted default for multiplicity 3, 16 / 42 electrons)."""
    for name, func in (("o2_triplet_uhf_def2tzvp", "None"), ("o2_triplet_uks_lda_def2tzvp", "lda")):
        p = _in_dir(os.path.join(HERE, "data", name), conf.get_calc_params)
        assert list(p[1]) == [8, 8] and p[3] == "def2-tzvp" and p[4] == func
        assert (p[5], p[6], p[7]) == (0, 3, "unrestricted") and not p[8] and not p[9]
    p = _in_dir(os.path.join(HERE, "data", "benzene_rhf_mp2_def2tzvp"), conf.get_calc_params)
    assert list(p[1]) == [6] * 6 + [1] * 6 and p[2].shape == (12, 3) and p[7] == "restricted" and p[9] is True
    cc = np.linalg.norm(p[2][0] - p[2][1])
    assert abs(cc - 1.397) < 1e-3


def test_error_behaviour(tmp_path):
    with pytest.raises(FileNotFoundError):
        _in_dir(str(tmp_path), conf.get_calc_params)
    (tmp_path / "sqc.conf").write_text("[calc]\ngeom_xyz = a.xyz\n")
    with pytest.raises(KeyError):
        _in_dir(str(tmp_path), conf.get_calc_params)
    (tmp_path / "bad.xyz").write_text("2\n\nXx 0 0 0\nH 0 0 1\n")
    with pytest.raises(ValueError):
        conf.read_xyz(str(tmp_path / "bad.xyz"))
    (tmp_path / "short.xyz").write_text("3\n\nH 0 0 0\nH 0 0 1\n")
    with pytest.raises(ValueError):
        conf.read_xyz(str(tmp_path / "short.xyz"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_same_tuples_as_reference_parser():
    """Every example directory of the reference parses to the same 12-tuple / dict / MM arrays."""
    bse = types.ModuleType("basis_set_exchange")
    bse.lut = types.SimpleNamespace(element_Z_from_sym=lambda s: conf._Z_OF[s.lower()])
    saved = {k: sys.modules.get(k) for k in ("basis_set_exchange", "input")}
    sys.modules["basis_set_exchange"] = bse
    sys.path.insert(0, os.path.join(REF, "code", "python3"))
    flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        sys.modules.pop("input", None)
        import input as ref_input
        dirs = sorted({os.path.dirname(p) for p in glob.glob(os.path.join(REF, "tests", "**", "sqc.conf"), recursive=True)})
        assert len(dirs) >= 15
        checked = 0
        for d in dirs:
            a = _in_dir(d, conf.get_analysis_params)
            b = _in_dir(d, ref_input.get_analysis_params)
            assert a == b
            if "[calc]" not in open(os.path.join(d, "sqc.conf")).read():
                continue
            mine = _in_dir(d, conf.get_calc_params)
            ref = _in_dir(d, ref_input.get_calc_params)
            for x, y in zip(mine, ref):
                if isinstance(y, np.ndarray):
                    assert np.array_equal(x, y)
                else:
                    assert x == y
            if mine[-1]:
                ma, mb = _in_dir(d, lambda: conf.read_mm_charges(mine[-1]))
                ra, rb = _in_dir(d, lambda: ref_input.read_mm_charges(ref[-1]))
                assert np.array_equal(ma, ra) and np.array_equal(mb, rb)
            checked += 1
        assert checked >= 15
    finally:
        sys.dont_write_bytecode = flag
        sys.path.remove(os.path.join(REF, "code", "python3"))
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
