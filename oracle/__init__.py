"""CPU oracle for the SQCC Fock-build hot path.  TEST INFRASTRUCTURE ONLY.

Everything in this package is a CPU restatement (numpy / plain C) of the reference
algorithms in ``/root/reference/code/python3`` (``hf_ksdft.py``, ``ksdft_functional.py``,
``mp2.py``), each function citing the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``sqcc_b200``) never
imports it: the product path is hand-written CUDA behind ``libsqcc_b200.so`` and fails
loudly if that library is missing.

Pinning status: the reference ships NO golden vectors / expected outputs for this path
(SURVEY.md section 4, 8c).  The oracle is therefore pinned against outputs of the
*unmodified reference itself*, executed in the build container with a stub
``interface_psi4`` (Psi4 is not installed): ``oracle/make_golden.py`` generated
``tests/golden/*.npz|json`` and ``tests/test_oracle_*.py`` re-check the restatements
against them.  Third-party arithmetic not under /root/reference (Psi4 1.9.1 integrals,
README.md:154) is outside the replaced path and remains "parity unpinned".
"""
