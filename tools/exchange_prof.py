"""Phase timing of the fused peer-memory exchange: torchrun --nproc-per-node W tools/exchange_prof.py [NBF]"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqcc_b200.backend import FockEngine  # noqa: E402

nbf = int(sys.argv[1]) if len(sys.argv) > 1 else 494
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = FockEngine(device=local, rank=rank, world=world)
eng.attach_eri_synth(nbf, 20261019)
q, _ = np.linalg.qr(np.random.default_rng(7).standard_normal((nbf, nbf)))
d = torch.from_numpy(2.0 * q[:, :nbf // 10] @ q[:, :nbf // 10].T).to(eng.device)
out = torch.empty((2, nbf, nbf), dtype=torch.float64, device=eng.device)
rounds = int(os.environ.get("ROUNDS", "1"))
steps = int(os.environ.get("STEPS", "20"))
for mode in ("p2p", "nccl", "p2p") if rounds == 1 else ("p2p", "nccl") * rounds:
    eng.set_exchange(mode)
    for _ in range(3):
        eng.jk_device(d, True, out)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.jk_device(d, True, out)
    e1.record()
    torch.cuda.synchronize()
    tm = eng.timings()
    msg = f"rank {rank} {mode}: step {e0.elapsed_time(e1) / steps:.4f} ms  jk_kernel {tm['jk_kernel']:.4f} finalize/exchange {tm['jk_finalize']:.4f}"
    if mode == "p2p":
        st = (ctypes.c_uint64 * 8)()
        eng.lib.sqcc_comm_stamps(eng.ctx, st)
        s = [int(x) for x in st[:6]]
        msg += "  phases us: A %.1f reduce %.1f gridbar %.1f B %.1f local %.1f" % tuple((s[i + 1] - s[i]) / 1e3 for i in range(5))
    print(msg, flush=True)
    dist.barrier()
eng.close()
dist.destroy_process_group()
