"""Multi-process correctness check of the fused peer-memory exchange (CUDA IPC path):
torchrun --nproc-per-node W tools/check_multigpu.py   -- every rank must end with the oracle's J/K, bit-identical
across ranks, for RHF / UHF / J-only, repeated builds and re-attached tensors; the NCCL path must agree to 1e-11; the
deterministic mode must be bit-reproducible and identical on all ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cjk, synth  # noqa: E402  (checker only)
from sqcc_b200.backend import FockEngine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = FockEngine(device=local, rank=rank, world=world)
TOL = 1e-11
worst = 0.0
for n, seed in ((70, 5), (37, 9), (130, 2)):
    eng.attach_eri_synth(n, seed)
    assert eng._p2p, "peer mapping failed"
    packed = cjk.synth_canonical(n, seed)
    for kind in ("rhf", "uhf", "jonly", "rhf"):
        d = synth.synth_density_uhf(n, n // 3, n // 4, 3) if kind == "uhf" else synth.synth_density_rhf(n, n // 4, 3)
        jr, kr = cjk.jk_packed(packed, n, d, want_k=kind != "jonly")
        for rep in range(2):
            j, k = eng.jk(d, want_k=kind != "jonly")
            worst = max(worst, np.abs(j - jr).max(), 0.0 if k is None else np.abs(k - kr).max())
            # bit-identical on every rank
            t = torch.from_numpy(np.concatenate([j.ravel(), [] if k is None else k.ravel()])).to(eng.device)
            ref = t.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(t, ref), f"rank {rank}: result differs from rank 0 ({kind})"
        eng.set_exchange("nccl")
        j2, k2 = eng.jk(d, want_k=kind != "jonly")
        worst = max(worst, np.abs(j2 - jr).max(), 0.0 if k2 is None else np.abs(k2 - kr).max())
        eng.set_exchange("p2p")
    # deterministic mode (64-bit fixed-point accumulation, integer sums across the ranks): two builds bit-identical,
    # every rank identical, still within the parity gate
    eng.set_deterministic(True)
    d = synth.synth_density_uhf(n, n // 3, n // 4, 3)
    jr, kr = cjk.jk_packed(packed, n, d)
    ja, ka = eng.jk(d)
    jb, kb = eng.jk(d)
    assert np.array_equal(ja, jb) and np.array_equal(ka, kb), f"rank {rank}: deterministic mode not reproducible"
    worst = max(worst, np.abs(ja - jr).max(), np.abs(ka - kr).max())
    t = torch.from_numpy(np.concatenate([ja.ravel(), ka.ravel()])).to(eng.device)
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref), f"rank {rank}: deterministic result differs from rank 0"
    eng.set_deterministic(False)
    # device-resident SCF step on the sharded tensor (replicated linear algebra, fused exchange inside)
    h = np.diag(-np.arange(1.0, n + 1.0))
    eng.scf_setup(h, np.eye(n), 1, n // 4, n // 4, False)
    eng.scf_set_density(synth.synth_density_rhf(n, n // 4, 3))
    e1 = eng.scf_fock_energy()
    eng.scf_update()
    e2 = eng.scf_fock_energy()
    et = torch.tensor([e1, e2], dtype=torch.float64, device=eng.device)
    e0 = et.clone()
    dist.broadcast(e0, 0)
    assert torch.equal(et, e0), f"SCF energies differ across ranks: {et.tolist()} vs {e0.tolist()}"
    # the same two steps with the partials summed by NCCL between the two halves of the step (sqcc_scf_jk ->
    # all-reduce -> sqcc_scf_finish): the Fock matrix must never be built from a rank's partial J/K
    eng.set_exchange("nccl")
    eng.scf_set_density(synth.synth_density_rhf(n, n // 4, 3))
    f1 = eng.scf_fock_energy()
    eng.scf_update()
    f2 = eng.scf_fock_energy()
    assert abs(f1 - e1) <= 1e-9 and abs(f2 - e2) <= 1e-9, (rank, e1, f1, e2, f2)
    # ... and against the host-side energy of the oracle's J/K for the same density (hf_ksdft.py:498-499, :556-557)
    d0 = synth.synth_density_rhf(n, n // 4, 3)
    jr, kr = cjk.jk_packed(packed, n, d0)
    e_ref = 0.5 * np.sum(d0 * (2.0 * h + jr - 0.5 * kr))
    assert abs(e1 - e_ref) <= 1e-9 * max(1.0, abs(e_ref)), (e1, e_ref)
    eng.set_exchange("p2p")
assert worst <= TOL, worst
eng.close()
dist.barrier()
if rank == 0:
    print(f"multi-GPU check ok on {world} ranks: max |J,K - oracle| = {worst:.2e}")
dist.destroy_process_group()
