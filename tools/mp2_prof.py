"""python tools/mp2_prof.py N NOCC : one MP2 call on a synthetic tensor (for ncu launch lists)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

n, no = int(sys.argv[1]), int(sys.argv[2])
eng = FockEngine(device=0)
eng.attach_eri_synth(n, 20261019)
rng = np.random.default_rng(1)
q, _ = np.linalg.qr(rng.standard_normal((n, n)))
eps = np.concatenate([np.sort(rng.uniform(-2, -0.4, no)), np.sort(rng.uniform(0.0, 3.0, n - no))])
for _ in range(2):
    e = eng.mp2_corr(q, eps, no)
print(e, eng.timings())
