"""Kernel-time sweep: python tools/jk_modes.py NBF [uhf|rhf|jonly] (env knobs select the configuration)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

nbf = int(sys.argv[1])
kind = sys.argv[2] if len(sys.argv) > 2 else "rhf"
eng = FockEngine(device=0)
eng.attach_eri_synth(nbf, 20261019)
q, _ = np.linalg.qr(np.random.default_rng(7).standard_normal((nbf, nbf)))
no = max(1, nbf // 10)
d = 2.0 * q[:, :no] @ q[:, :no].T
if kind == "uhf":
    d = np.stack([0.5 * d, 0.4 * d])
dd = torch.from_numpy(d).to(eng.device)
ts = []
reps = int(os.environ.get("REPS", "30"))
for it in range(reps):
    eng.jk_device(dd, kind != "jonly")
    ts.append(eng.timings()["jk_kernel"])
_, ab = eng.eri_bytes()
# sustained: back-to-back builds
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(reps):
    eng.jk_device(dd, kind != "jonly")
e1.record()
torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / reps
env = {k: v for k, v in os.environ.items() if k.startswith("SQCC_")}
print(f"nbf {nbf} {kind} {env}: kernel min {min(ts):.4f} median {np.median(ts):.4f} ms -> {ab / (np.median(ts) * 1e-3) / 1e9:.0f} GB/s; "
      f"sustained step {sus:.4f} ms -> {ab / (sus * 1e-3) / 1e9:.0f} GB/s", flush=True)
