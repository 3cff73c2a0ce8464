"""sqc.conf / xyz / MM-charge readers with the reference's API (input.py:18, :124, :260, :335).

Same function names, arguments, return tuples, file formats and exception types as the reference so that
``run.py`` is a drop-in; the implementation is independent (one INI loader shared by both getters, an
internal periodic table instead of basis_set_exchange when that package is absent).
"""
import configparser
import os

import numpy as np

_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
            "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb "
            "Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()
_Z_OF = {s.lower(): z for z, s in enumerate(_SYMBOLS, start=1)}


def element_number(token):
    """Atomic number from 'C', 'c' or '6' (input.py:97-105 accepts both spellings)."""
    try:
        return int(token)
    except ValueError:
        pass
    try:
        from basis_set_exchange import lut  # same lookup the reference uses, when installed
        return lut.element_Z_from_sym(token)
    except ImportError:
        try:
            return _Z_OF[token.lower()]
        except KeyError:
            raise ValueError(f"Unknown element symbol: {token}")
    except KeyError:
        raise ValueError(f"Unknown element symbol: {token}")


def read_xyz(xyz_file_name):
    """(mol_xyz, nuclear_numbers[int], coordinates[N,3] in Angstrom) from a standard XYZ file."""
    if not os.path.isfile(xyz_file_name):
        raise FileNotFoundError(f"XYZ file '{xyz_file_name}' not found.")
    with open(xyz_file_name, "r") as fh:
        lines = fh.readlines()
    if len(lines) < 3:
        raise ValueError("XYZ file must have at least 3 lines (number of atoms, comment, atom data)")
    try:
        expected = int(lines[0].strip())
    except ValueError:
        raise ValueError(f"First line must be an integer (number of atoms), got: {lines[0].strip()}")
    zs, xyz, kept = [], [], []
    for number, line in enumerate(lines[2:], start=3):
        if not line.strip():
            break  # a blank line ends the atom block
        fields = line.split()
        if len(fields) < 4:
            raise ValueError(f"Line {number} must have at least 4 elements (element x y z), got: {line.strip()}")
        zs.append(element_number(fields[0]))
        try:
            xyz.append([float(v) for v in fields[1:4]])
        except ValueError as err:
            raise ValueError(f"Invalid coordinate format on line {number}: {err}")
        kept.append(line)
    if len(zs) != expected:
        raise ValueError(f"Number of atoms mismatch: expected {expected}, got {len(zs)}")
    return "".join(kept), np.array(zs, dtype=int), np.array(xyz, dtype=float)


def _load_conf(flag_multiple_mols, another_conf_file_name):
    name = another_conf_file_name if flag_multiple_mols else "sqc.conf"
    if not os.path.isfile(name):
        raise FileNotFoundError(f"Configuration file '{name}' does not exist.")
    conf = configparser.ConfigParser()
    try:
        conf.read(name)
    except configparser.Error as err:
        raise ValueError(f"Error parsing config file: {err}")
    return conf


def get_calc_params(flag_multiple_mols=False, another_conf_file_name="sqc2.conf"):
    """The reference's 12-tuple (input.py:244-257):
    (mol_xyz, nuclear_numbers, geom_coordinates, basis_set_name, ksdft_functional_name, molecular_charge,
     spin_multiplicity, spin_orbital_treatment, flag_cis, flag_mp2, flag_qmmm, mm_charges_file)."""
    conf = _load_conf(flag_multiple_mols, another_conf_file_name)
    if "calc" not in conf:
        raise KeyError("Config file must have a [calc] section")
    calc = conf["calc"]
    try:
        xyz_file = calc["geom_xyz"]
        basis = calc["gauss_basis_set"]
        charge = int(calc["molecular_charge"])
        mult = int(calc["spin_multiplicity"])
    except KeyError as err:
        raise KeyError(f"Missing required parameter in config file: {err}")
    except ValueError as err:
        raise ValueError(f"Invalid integer value in config file: {err}")
    functional = calc.get("ksdft_functional", None)
    treatment = calc.get("spin_orbital_treatment")
    if treatment:
        treatment = treatment.lower()
        if treatment not in ("restricted", "unrestricted"):
            raise ValueError(f"spin_orbital_treatment must be 'restricted' or 'unrestricted', got '{treatment}'")
    else:
        treatment = "restricted" if mult == 1 else "unrestricted"
        print(f"spin_orbital_treatment not specified, using automatic default: '{treatment}'")
    mol_xyz, zs, coords = read_xyz(xyz_file)
    flag_cis = calc.get("excited_state", "").lower() == "cis"
    flag_mp2 = calc.get("post_hartree-fock", "").lower() == "mp2"
    flag_qmmm = calc.get("qmmm", "false").lower() == "true"
    mm_file = None
    if flag_qmmm:
        mm_file = calc.get("mm_charges", None)
        if mm_file is None:
            raise ValueError("QM/MM is enabled but 'mm_charges' parameter is missing in config file")
    return (mol_xyz, zs, coords, basis, functional, charge, mult, treatment, flag_cis, flag_mp2, flag_qmmm, mm_file)


def get_analysis_params(flag_multiple_mols=False, another_conf_file_name="sqc2.conf"):
    """[analysis] switches (input.py:306-330)."""
    conf = _load_conf(flag_multiple_mols, another_conf_file_name)
    params = {"electron_density": False, "electron_density_difference": False, "mo_file1": None, "mo_file2": None}
    if "analysis" in conf:
        sec = conf["analysis"]
        params["electron_density"] = sec.get("electron_density", "false").lower() == "true"
        params["electron_density_difference"] = sec.get("electron_density_difference", "false").lower() == "true"
        if params["electron_density_difference"]:
            params["mo_file1"] = sec.get("mo_file1", None)
            params["mo_file2"] = sec.get("mo_file2", None)
    return params


def read_mm_charges(file_path):
    """(mm_coords[M,3] in Angstrom, mm_charges[M]) from 'count / blank / q x y z' (input.py:368-390)."""
    if not os.path.isfile(file_path):
        raise FileNotFoundError(f"MM charges file '{file_path}' not found.")
    with open(file_path, "r") as fh:
        lines = fh.readlines()
    if len(lines) < 3:
        raise ValueError("MM charges file must have at least 3 lines")
    try:
        count = int(lines[0].strip())
    except ValueError:
        raise ValueError(f"First line must be an integer (number of charges), got: {lines[0].strip()}")
    body = lines[2:2 + count]
    if len(body) < count:
        raise ValueError(f"Expected {count} charges, but file has fewer lines")
    table = np.zeros((count, 4))
    for row, line in enumerate(body):
        fields = line.split()
        if len(fields) < 4:
            raise ValueError(f"Line {row + 3} must have 4 values (charge x y z), got: {line.strip()}")
        try:
            table[row] = [float(v) for v in fields[:4]]
        except ValueError as err:
            raise ValueError(f"Invalid number format on line {row + 3}: {err}")
    return table[:, 1:4].copy(), table[:, 0].copy()
