"""Kernel-level timing of the LDA grid quadrature: python tools/perf_lda.py [G N nspin]... (CUDA-event times of the rho and
V_xc kernels, TFLOP/s = 2 G N^2 nspin per kernel)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(354200, 62, 2), (200000, 120, 1), (200000, 222, 1)]
eng = FockEngine(device=0)
for g, n, nspin in cases:
    rng = np.random.default_rng(3)
    phi = rng.standard_normal((g, n)) * 0.1
    w = rng.random(g) * 1e-2
    eng.attach_grid(phi, w)
    a = rng.standard_normal((n, n))
    d = a @ a.T / n
    dd = d if nspin == 1 else np.stack([0.6 * d, 0.4 * d])
    for _ in range(3):
        eng.lda_vxc(dd)
    rho, vxc = [], []
    for _ in range(10):
        eng.lda_vxc(dd)
        t = eng.timings()
        rho.append(t["rho"])
        vxc.append(t["vxc"])
    fl = 2.0 * g * n * n * nspin
    r, v = float(np.median(rho)), float(np.median(vxc))
    print(f"G={g} N={n} nspin={nspin}: rho {r:.3f} ms = {fl / r / 1e9:.1f} TFLOP/s   vxc {v:.3f} ms = {fl / v / 1e9:.1f} TFLOP/s", flush=True)
