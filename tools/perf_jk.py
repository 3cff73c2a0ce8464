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
for variant in variants:
    eng.set_jk_variant(variant)
    for tag, dd, wk in (("rhf", d, True), ("uhf", du, True), ("jonly", d, False)):
        out = torch.empty((1 + (dd.dim() - 1 if wk else 0) * (2 if dd.dim() == 3 else 1) + (0 if not wk or dd.dim() == 3 else 0), n, n),
                          dtype=torch.float64, device="cuda") if False else None
        t0 = time.time()
        reps = 0
        while time.time() - t0 < 1.2:     # warm-up into the power-capped steady state
            for _ in range(10):
                eng.jk_device(dd, wk)
            torch.cuda.synchronize()
            reps += 10
        steps = max(10, min(200, reps // 2))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            eng.jk_device(dd, wk)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kms = eng.timings()["jk_kernel"]
        print(f"  variant {variant:2d} {tag:5s} step {ms:8.3f} ms  kernel(last) {kms:8.3f} ms  {algo / (ms * 1e-3) / 1e9:7.1f} GB/s algorithmic "
              f"({native / (ms * 1e-3) / 1e9:7.1f} native)  {1e3 / ms:7.2f} builds/s", flush=True)
eng.close()
