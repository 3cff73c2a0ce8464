"""sqcc_b200 -- B200-native SCF Fock-build hot path of SQCC (J/K, LDA grid quadrature, MP2 AO->MO).

Hand-written sm_100a CUDA behind a C ABI (include/sqcc_b200.h, sqcc_b200/lib/libsqcc_b200.so) with a thin
ctypes layer; the driver modules (hf_ksdft, mp2, cis_tdhf_tda_tddft, input, run) keep the reference's API.
"""
__version__ = "0.1.0"
