"""Secondary measurements: LDA grid quadrature (O2-size UKS) and MP2 AO->MO transforms (N2 / benzene size) on one B200,
CUDA-event kernel times from the library, FP64 GEMM peak from torch.matmul(float64).  python tools/bench_lda_mp2.py [--mp2-n 222]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

mp2_sizes = [(62, 7)]
if "--mp2-n" in sys.argv:
    n = int(sys.argv[sys.argv.index("--mp2-n") + 1])
    mp2_sizes.append((n, 21))
res = {}
# FP64 GEMM peak (cuBLAS DGEMM via torch), best of 5
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
res["dgemm_8192_tflops"] = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
del a, b

eng = FockEngine(device=0)
rng = np.random.default_rng(1)
g, n = 354200, 62
phi = rng.standard_normal((g, n)) * np.exp(-rng.uniform(0, 8, size=(g, 1)))
w = rng.uniform(0, 1e-2, size=g)
q, _ = np.linalg.qr(rng.standard_normal((n, n)))
d2 = np.stack([q[:, :9] @ q[:, :9].T, q[:, :7] @ q[:, :7].T])
eng.attach_grid(phi, w)
for tag, d in (("rks", 2.0 * d2[1]), ("uks", d2)):
    ts = []
    for it in range(6):
        t0 = time.perf_counter(); eng.lda_vxc(d); t1 = time.perf_counter()
        tm = eng.timings()
        if it >= 2:
            ts.append((tm["rho"], tm["vxc"], (t1 - t0) * 1e3))
    rho_ms, vxc_ms, wall = np.mean(ts, axis=0)
    nspin = 1 if d.ndim == 2 else 2
    fl = 2.0 * g * n * n * nspin
    res[f"lda_{tag}_G{g}_N{n}"] = {"rho_ms": rho_ms, "vxc_ms": vxc_ms, "host_call_ms": wall,
                                  "rho_tflops": fl / (rho_ms * 1e-3) / 1e12, "vxc_tflops": fl / (vxc_ms * 1e-3) / 1e12,
                                  "phi_gbs_rho": g * n * 8 / (rho_ms * 1e-3) / 1e9}
for n, no in mp2_sizes:
    nv = n - no
    eng.attach_eri_synth(n, 20261019)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    eps = np.concatenate([np.sort(rng.uniform(-2, -0.4, no)), np.sort(rng.uniform(0.0, 3.0, nv))])
    ts = []
    for it in range(3):
        t0 = time.perf_counter(); e = eng.mp2_corr(q, eps, no); t1 = time.perf_counter()
        tm = eng.timings()
        if it >= 1:
            ts.append((tm["mp2_transform"], tm["mp2_energy"], (t1 - t0) * 1e3))
    tr, en, wall = np.mean(ts, axis=0)
    fl = 2.0 * no * n ** 4 + 2.0 * no * no * n ** 3 + 2.0 * no * no * nv * n * n + 2.0 * no * no * nv * nv * n
    res[f"mp2_N{n}_nocc{no}"] = {"transform_ms": tr, "energy_ms": en, "host_call_ms": wall, "flops": fl,
                                 "tflops": fl / (tr * 1e-3) / 1e12, "e_corr": e}
print(json.dumps(res, indent=1))
