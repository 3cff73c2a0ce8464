"""Small end-to-end exercise of every kernel family (for compute-sanitizer memcheck): python tools/sanity_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine, FockGroup  # noqa: E402

rng = np.random.default_rng(0)


def dens(n, no):
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return 2.0 * q[:, :no] @ q[:, :no].T


eng = FockEngine(device=0)
for n in (5, 37, 70):
    eng.attach_eri_synth(n, 11)
    d = dens(n, max(1, n // 4))
    j, k = eng.jk(d)
    ju, ku = eng.jk(np.stack([0.5 * d, 0.5 * d]))
    jo, _ = eng.jk(d, want_k=False)
    for v in (1, 17, 18, 22, 23):
        eng.set_jk_variant(v)
        j2, k2 = eng.jk(d)
        assert np.abs(j2 - j).max() < 1e-10 and np.abs(k2 - k).max() < 1e-10
    eng.set_jk_variant(0)
    assert np.abs(ju - j).max() < 1e-10 and np.abs(jo - j).max() < 1e-10
n = 22
eng.attach_eri_synth(n, 3)
q, _ = np.linalg.qr(rng.standard_normal((n, n)))
eps = np.concatenate([np.sort(rng.uniform(-2, -0.4, 5)), np.sort(rng.uniform(0.0, 3.0, n - 5))])
e = eng.mp2_corr(q, eps, 5)
phi = rng.standard_normal((3001, n)) * np.exp(-rng.uniform(0, 8, size=(3001, 1)))
eng.attach_grid(phi, rng.uniform(0, 1e-2, size=3001))
d = dens(n, 5)
v, exc, rho = eng.lda_vxc(np.stack([0.6 * d, 0.4 * d]), want_rho=True)
h = rng.standard_normal((n, n))
h = 0.5 * (h + h.T) - 3 * np.eye(n)
for ks in (False, True):
    eng.scf_setup(h, np.eye(n), 2, 6, 4, ks)
    eng.scf_set_density(np.stack([0.6 * d, 0.4 * d]))
    for _ in range(3):
        eng.scf_fock_energy()
        eng.scf_update()
    eng.scf_get()
eng.close()
grp = FockGroup([0] * 3).attach_eri_synth(37, 5)
for _ in range(2):
    outs = grp.jk_device(dens(37, 9))
grp.close()
print("sanity ok", e, exc)
