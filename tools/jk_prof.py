"""Tiny driver for ncu captures: a few J/K builds on a synthetic packed ERI.  python tools/jk_prof.py NBF VARIANT NSPIN WANTK"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

nbf, variant, nspin, wantk = (int(x) for x in (sys.argv[1:5] + ["222", "0", "1", "1"][len(sys.argv) - 1:]))
eng = FockEngine(device=0)
eng.attach_eri_synth(nbf, 20261019)
eng.set_jk_variant(variant)
q, _ = np.linalg.qr(np.random.default_rng(7).standard_normal((nbf, nbf)))
d = 2.0 * q[:, :nbf // 10] @ q[:, :nbf // 10].T
if nspin == 2:
    d = np.stack([0.5 * d, 0.4 * d])
dd = torch.from_numpy(d).to(eng.device)
ts = []
for _ in range(4):
    eng.jk_device(dd, bool(wantk))
    ts.append(eng.timings()["jk_kernel"])
_, ab = eng.eri_bytes()
print(f"nbf {nbf} variant {variant} nspin {nspin} wantk {wantk}: kernel ms {ts}  GB/s {ab / (min(ts) * 1e-3) / 1e9:.1f}")
