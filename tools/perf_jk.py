"""Kernel-level timing of the J/K build: python tools/perf_jk.py [N] [variants...]  (CUDA events around back-to-back builds
after a warm-up long enough to reach the power-capped steady state; GB/s are algorithmic bytes 8 * nquad / time)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 494
variants = [int(a) for a in sys.argv[2:]] or [0, 17, 18]
nocc = {62: 7, 222: 21, 264: 21, 494: 47}.get(n, max(1, n // 10))
rng = np.random.default_rng(7)
q, _ = np.linalg.qr(rng.standard_normal((n, n)))
d = torch.from_numpy(2.0 * q[:, :nocc] @ q[:, :nocc].T).cuda()
qb, _ = np.linalg.qr(rng.standard_normal((n, n)))
du = torch.from_numpy(np.stack([q[:, :nocc + 1] @ q[:, :nocc + 1].T, qb[:, :nocc - 1] @ qb[:, :nocc - 1].T])).cuda()
eng = FockEngine(device=0)
eng.attach_eri_synth(n, 20261019)
eng.synchronize()
native, algo = eng.eri_bytes()
print(f"N={n} native {native / 1e9:.2f} GB algorithmic {algo / 1e9:.2f} GB", flush=True)
rounds = int(os.environ.get("PERF_ROUNDS", "3"))
kinds = os.environ.get("PERF_KINDS", "rhf,uhf,jonly").split(",")
res = {}
for rnd in range(rounds):      # variants interleaved: the power controller drifts over seconds
    for tag, dd, wk in (("rhf", d, True), ("uhf", du, True), ("jonly", d, False)):
        if tag not in kinds:
            continue
        for variant in variants:
            eng.set_jk_variant(variant)
            t0 = time.time()
            while time.time() - t0 < (1.0 if rnd == 0 else 0.3):     # warm-up into the power-capped steady state
                for _ in range(10):
                    eng.jk_device(dd, wk)
                torch.cuda.synchronize()
            steps = 40
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                eng.jk_device(dd, wk)
            e1.record()
            torch.cuda.synchronize()
            res.setdefault((tag, variant), []).append(e0.elapsed_time(e1) / steps)
for tag in kinds:
    for variant in variants:
        ms = res[(tag, variant)]
        m = float(np.median(ms))
        print(f"  {tag:5s} variant {variant:2d}  step ms {' '.join(f'{x:7.3f}' for x in ms)}  median {m:7.3f}  "
              f"{algo / (m * 1e-3) / 1e9:7.1f} GB/s algorithmic ({native / (m * 1e-3) / 1e9:7.1f} native)  {1e3 / m:7.2f} builds/s", flush=True)
eng.close()
