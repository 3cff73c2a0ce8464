"""Closed-shell MP2 with the reference's Calculator API (mp2.py:20-265).

``Calculator(scf_object).mp2()`` returns E_SCF + E_corr and prints the same two result lines.  The four AO->MO
quarter transforms (mp2.py:123-170) and the energy sum (:89-99, :240-254) run on the device as FP64 tensor-core
GEMMs (``FockEngine.mp2_corr`` -> ``sqcc_mp2``); the reference's second, identical transform (:185-228) is not
repeated -- (ib|ja) is the same array read transposed.  The n_occ*n_virt <= 1000 guard (mp2.py:68-74) only
protected the reference's Python loops and is lifted.
"""
from .backend import FockEngine


class Calculator:
    def __init__(self, scf_object):
        self.scf = scf_object

    def mp2(self):
        scf = self.scf
        if scf.spin_multiplicity != 1:
            raise NotImplementedError("Current MP2 only supports the closed-shell singlet state.")
        n_occ = int(scf.num_electrons / 2)
        engine = getattr(scf, "engine", None)
        if engine is None:   # an SCF object that did not come from sqcc_b200.hf_ksdft (e.g. loaded results)
            if getattr(scf, "ao_electron_repulsion_integral", None) is None:
                raise RuntimeError("the SCF object carries neither an engine nor the AO ERI tensor "
                                   "(direct_ingest on several ranks keeps only shards): nothing to transform")
            engine = FockEngine()
            engine.attach_eri_full(scf.ao_electron_repulsion_integral)
        if engine.world != 1:
            raise NotImplementedError("MP2 needs the unsharded ERI tensor on one GPU")
        print("Transforming AO integrals to MO basis (four quarter transforms on the GPU)...")
        correlation = engine.mp2_corr(scf.mo_coefficients, scf.mo_energies, n_occ)
        self.mp2_correlation_energy = correlation
        total = scf.scf_energy + correlation
        print("")
        print("MP2 energy (Hartree):", total)
        print("MP2 correlation energy (Hartree):", correlation)
        return total
