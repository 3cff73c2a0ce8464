"""Per-shard J/K kernel time on ONE GPU: python tools/shard_prof.py NBF WORLD [VARIANT]
Attaches every rank's pair-block shard of the synthetic tensor in turn and times the fused kernel on it, so shard
imbalance and small-shard efficiency can be read without a multi-GPU box."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

nbf, world = int(sys.argv[1]), int(sys.argv[2])
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
q, _ = np.linalg.qr(np.random.default_rng(7).standard_normal((nbf, nbf)))
d = 2.0 * q[:, :nbf // 10] @ q[:, :nbf // 10].T
ranks = [int(x) for x in os.environ["SQCC_RANKS"].split(",")] if os.environ.get("SQCC_RANKS") else range(world)
for r in ranks:
    eng = FockEngine(device=0, rank=r, world=world)
    eng.attach_eri_synth(nbf, 20261019)
    eng.set_jk_variant(variant)
    dd = torch.from_numpy(d).to(eng.device)
    ts = []
    for it in range(8):
        eng.jk_device(dd, True)
        if it >= 3:
            ts.append(eng.timings()["jk_kernel"])
    nb, ab = eng.eri_bytes()
    t = float(np.mean(ts))
    print(f"rank {r}/{world}: tiles [{eng.plan.t_begin},{eng.plan.t_end}) i [{eng.plan.i_begin},{eng.plan.i_end}) "
          f"native {nb / 1e9:.3f} GB algo {ab / 1e9:.3f} GB kernel {t:.4f} ms (min {min(ts):.4f})  {ab / t / 1e6:.0f} GB/s algo", flush=True)
    eng.close()
    del eng
    torch.cuda.empty_cache()
