#!/usr/bin/env python
"""Benchmark of the fused J/K Fock build (BASELINE.json metric) -- one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nbf 494]

Workload: one RHF Fock build (J and K, hf_ksdft.py:483-495) over the synthetic 8-fold packed ERI tensor of
the anthracene-size configuration (N = 494 basis functions, 59.8 GB packed, BASELINE.json configs[4] --
the configuration the north_star target is quoted on; it fits one B200).  At N GPUs the SAME tensor is
sharded by pair-block rows over the ranks (strong scaling): each rank streams its shard once, the
partial [J; K] is summed with one NCCL all-reduce.  A "step" is one Fock build.

  value        builds/s, density and outputs resident in HBM (device pointers), CUDA-event timed, max over ranks
  e2e          builds/s through the host-buffer API (numpy D in, numpy J/K out; H2D + D2H inside the timed region)
  roofline     dominant kernel (jk_pipe_kernel): algorithmic bytes 8*nquad(shard) / CUDA-event kernel time
  cpu_baseline the oracle's multi-core C port of the same pass, timed on a bounded row sample of the same tensor
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261019
NOCC = {62: 7, 222: 21, 264: 21, 494: 47}
WARM_MIN = 25        # untimed builds before every timed loop (>= --warmup) ...
WARM_SECONDS = 1.5   # ... and at least this long: every GPU reaches its 1000 W cap within ~0.3 s of streaming and the power
                     # controller then oscillates for ~1 s (8-GPU steps measured 1.36-1.74 ms inside that window, 1.6 ms after)


def nquad(n):
    p = n * (n + 1) // 2
    return p * (p + 1) // 2


def density_rhf(n, n_occ, seed):
    """D = 2 C_occ C_occ^T from a seeded orthonormal C (same recipe as oracle/synth.py)."""
    q, _ = np.linalg.qr(np.random.default_rng(seed).standard_normal((n, n)))
    return 2.0 * q[:, :n_occ] @ q[:, :n_occ].T


def density_uhf(n, n_a, n_b, seed):
    qa, _ = np.linalg.qr(np.random.default_rng(seed).standard_normal((n, n)))
    qb, _ = np.linalg.qr(np.random.default_rng(seed + 1).standard_normal((n, n)))
    return np.stack([qa[:, :n_a] @ qa[:, :n_a].T, qb[:, :n_b] @ qb[:, :n_b].T])


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nbf", type=int, default=494, help="basis functions of the synthetic packed ERI")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary (N=222/264, UHF, LDA, MP2) measurements")
    ap.add_argument("--variant", type=int, default=0, help="J/K kernel: 0 TMA pipeline, 2 direct-load, 1 simple")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: fused peer-memory exchange kernel (default) or finalize + NCCL all-reduce")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline sample time")
    return ap.parse_args()


def profiled_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full`
    capture (profiles/r01_traffic.json), per launch; None when no capture exists for this size."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(f"jk_pipe_rhf_n{n}", {}).get("dram_bytes")
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = get(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_port_baseline(n, d, seconds):
    """Multi-core C port (oracle) of the one-pass 8-fold J/K accumulation on a bounded sample of rows of the SAME
    synthetic tensor: the last rows (longest) that fit ~1.5 GB; builds/s = sample elements/s / nquad."""
    from oracle import cjk
    npair = n * (n + 1) // 2
    budget = int(1.5e9 // 8)
    ij0 = npair
    while ij0 > 0 and npair * (npair + 1) // 2 - (ij0 - 1) * ij0 // 2 <= budget:
        ij0 -= 1
    # coarse: shrink the sample further so one pass takes at most ~seconds/3
    rows = cjk.synth_rows(n, SEED, ij0, npair)
    elems = rows.size
    cjk.jk_packed_rows(rows, n, ij0, npair, d)  # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        cjk.jk_packed_rows(rows, n, ij0, npair, d)
        reps += 1
        el = time.perf_counter() - t0
        if el > seconds or reps >= 400:
            break
    per_elem = el / (reps * elems)
    builds = 1.0 / (per_elem * nquad(n))
    return {"value": builds, "unit": "builds/s", "cores": cjk.num_threads(), "kind": "port",
            "sample": f"rows ij in [{ij0},{npair}) of the N={n} synthetic packed tensor ({elems * 8 / 1e9:.2f} GB, "
                      f"{reps} passes, {el:.1f} s), oracle/csrc/jk_oracle.c one-pass 8-fold J/K with OpenMP; "
                      f"extrapolated by element count to the whole tensor",
            "gbs": 8.0 / per_elem / 1e9}


def reference_arm(args):
    """--impl reference: the reference's own CPU form of the path (hf_ksdft.py:483-495: per (p,q) an elementwise
    multiply + np.sum over an N x N slab, J then K), numpy, on a bounded sample of (p,q) entries per step."""
    from oracle import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.nbf
    d = synth.synth_density_rhf(n, NOCC.get(n, max(1, n // 10)), 7)
    rng = np.random.default_rng(1)
    m = 24  # sampled (p, q) entries per step
    pq = [(int(a), int(b)) for a, b in rng.integers(0, n, size=(m, 2))]
    slabs_j = [synth.synth_eri_slab_pq(n, SEED, p, q) for p, q in pq]      # ERI[p,q,:,:]
    slabs_k = [synth.synth_eri_slab_pxqx(n, SEED, p, q) for p, q in pq]    # ERI[p,:,q,:] (contiguous copy)

    def step():
        acc = 0.0
        for sj, sk in zip(slabs_j, slabs_k):
            acc += np.sum(sj * d)      # hf_ksdft.py:486-487
            acc += np.sum(sk * d)      # hf_ksdft.py:494-495
        return acc

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    per_entry = el / (args.steps * m)
    ms_per_build = per_entry * n * n * 1e3
    value = 1e3 / ms_per_build
    sample = (f"{m} random (p,q) entries per step of the N={n} synthetic tensor (slabs pre-generated outside the timed "
              f"region); one build = N^2 = {n * n} entries, extrapolated")
    line = {"impl": "reference", "metric": "J/K Fock builds/sec", "value": value, "unit": "builds/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_build,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"RHF J+K Fock build, synthetic 8-fold symmetric ERI, N={n} basis functions",
                       "nbf": n, "nquad": nquad(n)},
            "cpu_baseline": {"value": value, "unit": "builds/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "builds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from sqcc_b200.backend import FockEngine   # the product path; oracle/ is only used by the CPU-baseline legs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.nbf
    eng = FockEngine(device=local, rank=rank, world=world, exchange=args.exchange)
    eng.attach_eri_synth(n, SEED)
    eng.set_jk_variant(args.variant)
    eng.synchronize()
    d_host = density_rhf(n, NOCC.get(n, max(1, n // 10)), 7)
    d_dev = torch.from_numpy(d_host).to(eng.device)
    out = torch.empty((2, n, n), dtype=torch.float64, device=eng.device)
    native_bytes, algo_bytes = eng.eri_bytes()
    peak, peak_src = peaks()

    # ---- device-resident throughput (value).  Untimed warm-up: at least WARM_MIN builds -- the first ~20 builds after
    # the ERI fill run 3-5 % slower (clock / power-state ramp, seen on every rank of the 8-GPU runs)
    def warm_up(step):
        """max(W, WARM_MIN) untimed steps, continued until WARM_SECONDS have passed (same count on every rank)."""
        n_done, t_start = 0, time.perf_counter()
        while True:
            for _ in range(max(args.warmup, WARM_MIN)):
                step()
            n_done += max(args.warmup, WARM_MIN)
            torch.cuda.synchronize()
            go = torch.tensor([1.0 if time.perf_counter() - t_start < WARM_SECONDS else 0.0], device=eng.device)
            if world > 1:
                dist.all_reduce(go, op=dist.ReduceOp.MAX)
            if go.item() == 0.0:
                return n_done

    warm = warm_up(lambda: eng.jk_device(d_dev, True, out))
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.jk_device(d_dev, True, out)
    e1.record()
    barrier()
    clocks = sampler.result()
    launches = eng.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps

    # ---- dominant-kernel time, per launch, CUDA events on the launching stream (roofline)
    ktimes = []
    for _ in range(max(3, min(args.steps, 10))):
        eng.jk_device(d_dev, True, out)
        ktimes.append(eng.timings()["jk_kernel"])
    k_ms = float(np.mean(ktimes))
    kt = torch.tensor([k_ms], dtype=torch.float64, device=eng.device)
    ab = torch.tensor([float(algo_bytes)], dtype=torch.float64, device=eng.device)
    rank_ms = [k_ms]
    if world > 1:
        gathered = [torch.zeros_like(kt) for _ in range(world)]
        dist.all_gather(gathered, kt)
        rank_ms = [float(x.item()) for x in gathered]
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    achieved = float(ab.item()) / (float(kt.item()) * 1e-3) / 1e9  # GB/s on the slowest rank's shard

    # ---- the same steps with the other exchange (N > 1): finalize + NCCL all-reduce vs the fused peer-memory kernel
    other = None
    used_exchange = eng.exchange       # "nccl" if the peer mapping was refused and every rank fell back
    if world > 1 and (used_exchange == "p2p" or args.exchange == "nccl"):
        alt = "nccl" if used_exchange == "p2p" else "p2p"
        eng.set_exchange(alt)
        warm_up(lambda: eng.jk_device(d_dev, True, out))
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            eng.jk_device(d_dev, True, out)
        a1.record()
        barrier()
        ams = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=eng.device)
        dist.all_reduce(ams, op=dist.ReduceOp.MAX)
        other = {"exchange": alt, "ms_per_step": float(ams.item()) / args.steps,
                 "builds_per_s": 1e3 * args.steps / float(ams.item())}
        eng.set_exchange(used_exchange)

    # ---- end to end through host buffers: pinned host D in, pinned host J/K out, H2D + D2H inside every step
    d_pin = eng.pinned(n, n)
    d_pin[...] = d_host
    jk_pin = (eng.pinned(n, n), eng.pinned(1, n, n))
    warm_up(lambda: eng.jk(d_pin, out=jk_pin))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        j_h, k_h = eng.jk(d_pin, out=jk_pin)
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_builds = args.steps / float(t_e2e.item())

    if rank == 0:
        line = {
            "metric": "J/K Fock builds/sec", "value": 1e3 / ms_per_step, "unit": "builds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "warmup_actual": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"RHF J+K Fock build, synthetic 8-fold packed ERI, N={n} basis functions "
                                   f"(anthracene def2-TZVP size), pair-block sharded over {world} GPU(s)",
                       "nbf": n, "nquad": nquad(n), "packed_gb_algorithmic": nquad(n) * 8 / 1e9,
                       "native_gb_rank0": native_bytes / 1e9,
                       "parallelism": f"pair-block (4x4 tile) shard x{world}" + ("" if world == 1 else
                                      (" + fused peer-memory exchange kernel (NVLink)" if used_exchange == "p2p"
                                       else " + NCCL all-reduce")),
                       "l2_policy": "inputs larger than L2 (ERI shard >> 126 MB streamed every step)"},
            "eri_gbs_aggregate": nquad(n) * 8 / (ms_per_step * 1e-3) / 1e9,
            "frac_of_8tbs_per_gpu": nquad(n) * 8 / (ms_per_step * 1e-3) / 1e9 / world / 8000.0,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profiled_traffic(n) if world == 1 else None, "kernel": "jk_pipe_kernel", "kernel_ms": float(kt.item()),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": float(ab.item())},
            "e2e": {"value": e2e_builds, "unit": "builds/s", "h2d_bytes_per_step": int(d_host.nbytes),
                    "d2h_bytes_per_step": int(2 * n * n * 8)},
            "gpu_launches": int(launches), "clocks": clocks, "rank_kernel_ms": rank_ms,
        }
        if other:
            line["exchange_alternative"] = other
        if world == 1:
            line["cpu_baseline"] = cpu_port_baseline(n, d_host, args.cpu_seconds)
        if world == 1 and not args.no_extras:
            try:
                line["extras"] = extras(eng, torch)
            except Exception as exc:  # extras never invalidate the headline line
                line["extras"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras(eng, torch):
    """Secondary measurements at 1 GPU (CUDA-event kernel times): other basis sizes, UHF, J-only."""
    res = {}
    for n in (264, 222):
        eng.attach_eri_synth(n, SEED)
        _, ab = eng.eri_bytes()
        for tag, d, wk in (("rhf", density_rhf(n, NOCC[n], 7), True),
                           ("uhf", density_uhf(n, NOCC[n] + 1, NOCC[n] - 1, 7), True),
                           ("j_only", density_rhf(n, NOCC[n], 7), False)):
            dd = torch.from_numpy(d).to(eng.device)
            ts = []
            for it in range(8):
                eng.jk_device(dd, wk)
                if it >= 3:
                    ts.append(eng.timings()["jk_kernel"])
            t = float(np.mean(ts))
            res[f"n{n}_{tag}"] = {"kernel_ms": t, "gbs": ab / (t * 1e-3) / 1e9, "builds_per_s_kernel": 1e3 / t}
    # FP64 tensor-pipe work: LDA grid quadrature at the O2 def2-TZVP size and the MP2 AO->MO transforms at benzene size,
    # against cuBLAS DGEMM (torch.matmul float64 8192^3) measured in the same run
    a = torch.randn(8192, 8192, dtype=torch.float64, device=eng.device)
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a
    peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    res["fp64_dgemm_tflops_measured"] = peak
    rng = np.random.default_rng(1)
    g, n = 354200, 62
    phi = rng.standard_normal((g, n)) * np.exp(-rng.uniform(0, 8, size=(g, 1)))
    eng.attach_grid(phi, rng.uniform(0, 1e-2, size=g))
    d = density_uhf(n, 9, 7, 3)
    ts = []
    for it in range(5):
        eng.lda_vxc(d)
        tm = eng.timings()
        if it >= 2:
            ts.append((tm["rho"], tm["vxc"]))
    rho_ms, vxc_ms = np.mean(ts, axis=0)
    fl = 2.0 * g * n * n * 2
    res["lda_uks_G354200_N62"] = {"rho_ms": float(rho_ms), "vxc_ms": float(vxc_ms),
                                  "rho_tflops": fl / (rho_ms * 1e-3) / 1e12, "vxc_tflops": fl / (vxc_ms * 1e-3) / 1e12,
                                  "vxc_frac_of_dgemm": fl / (vxc_ms * 1e-3) / 1e12 / peak}
    n, no = 222, 21
    eng.attach_eri_synth(n, SEED)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    eps = np.concatenate([np.sort(rng.uniform(-2, -0.4, no)), np.sort(rng.uniform(0.0, 3.0, n - no))])
    ts = []
    for it in range(3):
        eng.mp2_corr(q, eps, no)
        if it >= 1:
            ts.append(eng.timings()["mp2_transform"])
    t = float(np.mean(ts))
    nv = n - no
    fl = 2.0 * no * n ** 4 + 2.0 * no * no * n ** 3 + 2.0 * no * no * nv * n * n + 2.0 * no * no * nv * nv * n
    res["mp2_N222_nocc21"] = {"transform_ms": t, "reference_flops": fl, "tflops_reference_equivalent": fl / (t * 1e-3) / 1e12}
    res["scf_iteration"] = scf_iteration_times(eng)
    return res


def scf_iteration_times(eng, sizes=(62, 222, 494), iters=6):
    """Wall time of ONE SCF iteration (BASELINE metric part 3) around the J/K build, synthetic operands (random symmetric
    H_core, X = I, synthetic ERI: the loop does not converge, the per-iteration work is what is timed):
      host_la   the reference's structure -- J/K through host buffers, then NumPy Fock / energy / X^T F X / eigh /
                C = X C' / density (hf_ksdft.py:498-499, :556-557, :97-123, :616-625) -- what Calculator.scf() does;
      device_la the device-resident iteration (csrc/scf.cu): one double crosses PCIe per iteration."""
    from sqcc_b200.hf_ksdft import Calculator
    out = {}
    for n in sizes:
        eng.attach_eri_synth(n, SEED)
        rng = np.random.default_rng(n)
        h = rng.standard_normal((n, n))
        h = 0.5 * (h + h.T) - 3.0 * np.eye(n)
        x = np.eye(n)
        nocc = NOCC.get(n, max(1, n // 10))
        d0 = density_rhf(n, nocc, 7)
        # host linear algebra
        d = d0.copy()
        t_jk = t_la = 0.0
        for it in range(iters + 1):
            t0 = time.perf_counter()
            j, k = eng.jk(d)
            t1 = time.perf_counter()
            fock = h + j - 0.5 * k
            _e = 0.5 * np.sum(d * (h + fock))
            _w, cp = Calculator.solve_one_electron_problem(x, fock)
            c = np.matmul(x, cp)
            d = 2.0 * np.matmul(c[:, :nocc], c[:, :nocc].T)
            t2 = time.perf_counter()
            if it > 0:
                t_jk += t1 - t0
                t_la += t2 - t1
        # device-resident
        eng.scf_setup(h, x, 1, nocc, nocc, False)
        eng.scf_set_density(d0)
        eng.scf_fock_energy()
        eng.scf_update()
        eng.synchronize()
        t0 = time.perf_counter()
        for it in range(iters):
            eng.scf_fock_energy()
            eng.scf_update()
        eng.synchronize()
        t_dev = time.perf_counter() - t0
        out[f"n{n}"] = {"host_la_ms": 1e3 * (t_jk + t_la) / iters, "host_la_jk_ms": 1e3 * t_jk / iters,
                        "host_la_numpy_ms": 1e3 * t_la / iters, "device_la_ms": 1e3 * t_dev / iters,
                        "device_update_kernel_ms": eng.timings()["scf_update"]}
    return out


if __name__ == "__main__":
    main()
