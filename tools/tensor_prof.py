"""One LDA quadrature (O2 def2-TZVP size, UKS) and one MP2 transform (benzene size) for ncu captures of the DMMA kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

rng = np.random.default_rng(1)
eng = FockEngine(device=0)
g, n = 354200, 62
phi = rng.standard_normal((g, n)) * np.exp(-rng.uniform(0, 8, size=(g, 1)))
eng.attach_grid(phi, rng.uniform(0, 1e-2, size=g))
qa, _ = np.linalg.qr(rng.standard_normal((n, n)))
d = np.stack([qa[:, :9] @ qa[:, :9].T, qa[:, :7] @ qa[:, :7].T])
for _ in range(2):
    eng.lda_vxc(d)
n, no = 222, 21
eng.attach_eri_synth(n, 20261019)
q, _ = np.linalg.qr(rng.standard_normal((n, n)))
eps = np.concatenate([np.sort(rng.uniform(-2, -0.4, no)), np.sort(rng.uniform(0.0, 3.0, n - no))])
for _ in range(2):
    eng.mp2_corr(q, eps, no)
print("done", eng.timings())
