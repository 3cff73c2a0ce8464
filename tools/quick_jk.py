"""Quick GPU check of the J/K kernels against the C oracle: python tools/quick_jk.py [n ...] (prints as it goes)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cjk, synth  # noqa: E402  (checker only)
from sqcc_b200.backend import FockEngine  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [5, 37, 70, 131]
eng = FockEngine(device=0)
for n in sizes:
    seed = 11 + n
    eng.attach_eri_synth(n, seed)
    eng.synchronize()
    print(f"n={n} attached", flush=True)
    packed = cjk.synth_canonical(n, seed)
    d = synth.synth_density_rhf(n, max(1, n // 4), 3)
    du = synth.synth_density_uhf(n, max(1, n // 3), max(1, n // 4), 5)
    for variant in [int(v) for v in os.environ.get('QUICK_VARIANTS', '1,0,17,18,22').split(',')]:
        eng.set_jk_variant(variant)
        for tag, dd, wk in (("rhf", d, True), ("uhf", du, True), ("jonly", d, False)):
            t0 = time.time()
            j, k = eng.jk(dd, want_k=wk)
            jr, kr = cjk.jk_packed(packed, n, dd, want_k=wk)
            ej = np.abs(j - jr).max()
            ek = np.abs(k - kr).max() if wk else 0.0
            print(f"  variant {variant:2d} {tag:5s} max|dJ| {ej:.2e} max|dK| {ek:.2e}  ({time.time() - t0:.2f}s)", flush=True)
    eng.set_jk_variant(0)
    eng.set_deterministic(True)
    a = eng.jk(du)
    b = eng.jk(du)
    jr, kr = cjk.jk_packed(packed, n, du)
    print(f"  deterministic: identical {np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])} "
          f"max|dJ| {np.abs(a[0] - jr).max():.2e} max|dK| {np.abs(a[1] - kr).max():.2e}", flush=True)
    eng.set_deterministic(False)
eng.close()
print("quick_jk done")
