"""Short J/K run for ncu: python tools/prof_jk.py N rhf|uhf|jonly [variant] [builds]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

n, kind = int(sys.argv[1]), sys.argv[2]
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
builds = int(sys.argv[4]) if len(sys.argv) > 4 else 4
nocc = {62: 7, 222: 21, 264: 21, 494: 47}.get(n, max(1, n // 10))
rng = np.random.default_rng(7)
q, _ = np.linalg.qr(rng.standard_normal((n, n)))
qb, _ = np.linalg.qr(rng.standard_normal((n, n)))
d = 2.0 * q[:, :nocc] @ q[:, :nocc].T
if kind == "uhf":
    d = np.stack([q[:, :nocc + 1] @ q[:, :nocc + 1].T, qb[:, :nocc - 1] @ qb[:, :nocc - 1].T])
dd = torch.from_numpy(d).cuda()
eng = FockEngine(device=0)
eng.attach_eri_synth(n, 20261019)
eng.set_jk_variant(variant)
for _ in range(builds):
    eng.jk_device(dd, kind != "jonly")
torch.cuda.synchronize()
print("kernel ms", eng.timings()["jk_kernel"])
eng.close()
