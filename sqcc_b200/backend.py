"""Host-side engine over libsqcc_b200.so: device buffers (torch), pair-block sharding, NCCL all-reduce.

One process drives one GPU.  With ``world > 1`` (torchrun) every rank holds the contiguous pair-block
shard ``[ij_begin, ij_end)`` of the packed ERI tensor, builds partial J/K with one pass over its shard,
and the partial matrices are summed with ONE ``all_reduce`` on the stacked ``[J; K]`` buffer
(SURVEY.md 8e).  PyTorch is used only for device memory, streams and ``torch.distributed``.
"""
import ctypes

import numpy as np
import torch

from . import _lib


class ShardPlan:
    """Equal-area cut of the packed tensor over `world` ranks at 4 x 4 pair-block ("tile") granularity
    (include/sqcc_b200.h: sqcc_shard_tiles; pure host arithmetic, no GPU needed).

    Rank r holds the tile range [t_begin, t_end); `jlo`/`jhi` give, per i, the interval of j whose rows (i, j)
    it stores (pair-index order inside the shard buffer)."""

    def __init__(self, n, world=1, rank=0, weights=None):
        """weights: optional relative speeds of the ranks (share of rank r = weights[r] / sum)."""
        lib = _lib.load()
        self.n, self.world, self.rank = int(n), int(world), int(rank)
        self.weights = None if weights is None else [float(w) for w in weights]
        b, e = ctypes.c_int64(), ctypes.c_int64()
        if self.weights is None:
            _lib.check(lib.sqcc_shard_tiles(self.n, self.world, self.rank, ctypes.byref(b), ctypes.byref(e)))
        else:
            assert len(self.weights) == self.world
            w = (ctypes.c_double * self.world)(*self.weights)
            _lib.check(lib.sqcc_shard_tiles_weighted(self.n, self.world, self.rank, w, ctypes.byref(b), ctypes.byref(e)))
        self.t_begin, self.t_end = b.value, e.value
        self.native_len = lib.sqcc_native_len_tiles(self.n, self.t_begin, self.t_end)
        if self.native_len < 0:
            raise _lib.SqccError("invalid shard")
        lo, hi = (ctypes.c_int * self.n)(), (ctypes.c_int * self.n)()
        _lib.check(lib.sqcc_tile_rows(self.n, self.t_begin, self.t_end, lo, hi))
        self.jlo, self.jhi = np.array(lo[:], dtype=np.int64), np.array(hi[:], dtype=np.int64)
        held = np.nonzero(self.jhi > self.jlo)[0]
        # first / one-past-last i with held rows
        self.i_begin, self.i_end = (int(held[0]), int(held[-1]) + 1) if held.size else (0, 0)
        self.nrows = int((self.jhi - self.jlo).sum())

    def segments(self):
        """Maximal runs [ij_lo, ij_hi) of consecutive pair indices held by this rank (canonical row ranges)."""
        segs = []
        for i in range(self.n):
            if self.jhi[i] > self.jlo[i]:
                a = i * (i + 1) // 2 + int(self.jlo[i])
                b = i * (i + 1) // 2 + int(self.jhi[i])
                if segs and segs[-1][1] == a:
                    segs[-1][1] = b
                else:
                    segs.append([a, b])
        return [tuple(x) for x in segs]

    @staticmethod
    def all_cuts(n, world, weights=None):
        return [ShardPlan(n, world, r, weights) for r in range(world)]


def reduce_partials(buf, group=None):
    """Sum the stacked partial [J; K] buffer over ranks in place (the path's only exchange step)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _RawDeviceBuffer:
    """A library-owned float64 device buffer exposed through __cuda_array_interface__ (no copy, no ownership)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _device_view(ptr, count, device):
    with torch.cuda.device(device):
        return torch.as_tensor(_RawDeviceBuffer(ptr, count), device=device)


class FockEngine:
    """J/K, LDA grid quadrature and MP2 transforms on one B200; mirrors the loop nests of hf_ksdft.py/mp2.py."""

    def __init__(self, device=None, rank=0, world=1, group=None, exchange="p2p"):
        """exchange: how the partial [J; K] of the ranks are summed when a process group is initialised --
        "p2p"  one fused kernel per rank over NVLink peer memory (csrc/comm.cu; the default),
        "nccl" finalize + ``dist.all_reduce`` (the baseline the fused exchange is measured against)."""
        self.lib = _lib.load()
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.exchange = exchange
        self._p2p = False      # the fused exchange is active: sqcc_jk_build returns the FULL J/K
        self._mapped = False   # the peers' regions are mapped (stays true while the exchange is toggled off)
        if not torch.cuda.is_available():
            raise _lib.SqccError("sqcc_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.rank, self.world, self.group = rank, world, group
        h = ctypes.c_void_p()
        _lib.check(self.lib.sqcc_ctx_create(self.device.index, ctypes.byref(h)))
        self.ctx = h
        self.n = 0
        self.plan = None
        self.eri = None
        self._phi = self._w = None
        self.shard_weights = None    # relative rank speeds of the current shard cut (None: equal shares)
        self._refill = None          # re-creates the tensor on a new shard cut (rebalance)
        self._use_stream()

    def _use_stream(self):
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream().cuda_stream
        _lib.check(self.lib.sqcc_ctx_set_stream(self.ctx, ctypes.c_void_p(s)))

    def _detach_peers(self, collective=True):
        """Unmap the peers' regions; nobody may free its region before every rank has unmapped it (barrier).
        collective=False (garbage collection / interpreter exit) skips the barrier: a destructor must not block on
        ranks that may never arrive."""
        if self._mapped:
            self._p2p = self._mapped = False
            self.lib.sqcc_comm_detach(self.ctx)
            import torch.distributed as dist
            if collective and dist.is_available() and dist.is_initialized():
                dist.barrier(group=self.group)

    def close(self, collective=True):
        """Release the context.  With a peer group attached this is a collective call (all ranks close together)."""
        if getattr(self, "ctx", None):
            self._detach_peers(collective)
            self.lib.sqcc_ctx_destroy(self.ctx)
            self.ctx = None
            self.eri = self._phi = self._w = None     # give the device memory back now, not at garbage collection
            self._refill = None                       # (the closure refers back to this engine)

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:
            pass

    # ------------------------------------------------------------------ ERI ingest
    def _alloc_shard(self, n):
        self._detach_peers()
        self.n = int(n)
        self.plan = ShardPlan(n, self.world, self.rank, self.shard_weights)
        # + 4 KB of slack: the zero-fill cp.async of the lanes past a narrow piece carry (unread) addresses up to 368
        # bytes beyond the last piece
        self.eri = torch.empty(self.plan.native_len + 512, dtype=torch.float64, device=self.device)
        self._use_stream()
        _lib.check(self.lib.sqcc_eri_attach_tiles(self.ctx, _ptr(self.eri), self.n, self.plan.t_begin, self.plan.t_end))
        self._p2p = False
        if self.exchange == "p2p" and self._group_size() > 1:
            self._attach_peers()

    def _group_size(self):
        import torch.distributed as dist
        if self.world > 1 and dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def _attach_peers(self):
        """Exchange the CUDA IPC handles of the ranks' accumulator regions (torch.distributed is only the
        messenger) and map them: afterwards sqcc_jk_build returns the full J/K on every rank.  If any rank cannot
        map its peers (no peer access / IPC disabled) ALL ranks fall back to finalize + NCCL all-reduce together."""
        import warnings
        import torch.distributed as dist
        nb = 64  # SQCC_COMM_HANDLE_BYTES
        mine = (ctypes.c_ubyte * nb)()
        ok = self.lib.sqcc_comm_export(self.ctx, mine) == 0
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
        gathered = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(gathered, t, group=self.group)
        err = ""
        if ok:
            blob = bytes(torch.cat(gathered).cpu().numpy().tobytes())
            ok = self.lib.sqcc_comm_attach(self.ctx, self.world, self.rank, blob) == 0
            if not ok:
                err = self.lib.sqcc_last_error().decode()
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self._p2p = self._mapped = True
            return
        self.lib.sqcc_comm_detach(self.ctx)
        dist.barrier(group=self.group)
        self._p2p = self._mapped = False
        self.exchange = "nccl"
        if self.rank == 0:
            warnings.warn("sqcc_b200: peer-memory exchange unavailable (" + (err or "a rank failed to map its peers") +
                          "); using finalize + NCCL all-reduce")

    def attach_eri_synth(self, n, seed):
        """Counter-based synthetic packed ERI, generated on the device shard by shard."""
        self._alloc_shard(n)
        _lib.check(self.lib.sqcc_eri_synth(self.ctx, ctypes.c_uint64(seed)))
        self._refill = lambda: self.attach_eri_synth(n, seed)
        return self

    def attach_eri_full(self, eri_full):
        """Pack the full [N,N,N,N] host tensor of interface_psi4.py:177-190 into this rank's shard."""
        eri_full = np.ascontiguousarray(eri_full, dtype=np.float64)
        n = eri_full.shape[0]
        assert eri_full.shape == (n, n, n, n)
        self._alloc_shard(n)
        _lib.check(self.lib.sqcc_eri_pack_full_host(self.ctx, eri_full.ctypes.data_as(ctypes.c_void_p)))
        self._refill = (lambda: self.attach_eri_full(eri_full)) if self.world > 1 else None   # keeps the host tensor alive
        return self

    def attach_eri_canonical(self, packed, n, rows_per_chunk=None):
        """Pack a canonical packed host vector (ij(ij+1)/2+kl) into this rank's shard."""
        packed = np.ascontiguousarray(packed, dtype=np.float64)
        self._alloc_shard(n)
        budget = 64 << 20  # doubles per staged chunk
        for lo, hi in self.plan.segments():
            ij = lo
            while ij < hi:
                e = ij
                while e < hi and (e + 1) * (e + 2) // 2 - ij * (ij + 1) // 2 <= budget:
                    e += 1
                e = max(e, ij + 1)
                seg = torch.from_numpy(packed[ij * (ij + 1) // 2: e * (e + 1) // 2]).to(self.device)
                _lib.check(self.lib.sqcc_eri_pack_canonical(self.ctx, _ptr(seg), ij, e - ij))
                torch.cuda.current_stream().synchronize()
                ij = e
        return self

    def attach_eri_blocks(self, n, blocks, batch_doubles=8 << 20):
        """Direct ingestion: `blocks` yields (i0, j0, k0, l0, values[ni,nj,nk,nl]) -- e.g. one
        dense block per canonical shell quartet -- and every integral goes straight to its place in this rank's packed
        shard; the full [N,N,N,N] host tensor is never built.  Blocks are staged through pinned memory in batches."""
        self._alloc_shard(n)
        _lib.check(self.lib.sqcc_eri_clear(self.ctx))
        stage = torch.empty(batch_doubles, dtype=torch.float64, pin_memory=True)
        stage_np = stage.numpy()
        dev = torch.empty(batch_doubles, dtype=torch.float64, device=self.device)
        desc, used = [], 0

        def flush():
            nonlocal desc, used
            if not desc:
                return
            dev[:used].copy_(stage[:used], non_blocking=True)
            dd = torch.tensor(desc, dtype=torch.int64).to(self.device)
            _lib.check(self.lib.sqcc_eri_pack_blocks(self.ctx, _ptr(dev), _ptr(dd), len(desc)))
            torch.cuda.current_stream().synchronize()
            desc, used = [], 0

        for i0, j0, k0, l0, vals in blocks:
            vals = np.ascontiguousarray(vals, dtype=np.float64)
            ni, nj, nk, nl = vals.shape
            if vals.size > batch_doubles:
                raise ValueError("block larger than the staging batch")
            if used + vals.size > batch_doubles:
                flush()
            stage_np[used:used + vals.size] = vals.ravel()
            desc.append([used, i0, ni, j0, nj, k0, nk, l0, nl])
            used += vals.size
        flush()
        return self

    def rebalance(self, steps=24, rounds=2, damping=1.0):
        """Re-cut the shards from MEASURED kernel times (collective; needs a tensor that can be re-created: synthetic or
        a host tensor).  Under their 1000 W caps the GPUs of one box sustain different SM clocks (1.27-1.55 ms for equal
        shards of N = 494 on 8 GPUs), and the exchange waits for the slowest rank: rank r's share becomes proportional to
        its measured streaming speed (share / kernel time).  Returns the relative speeds used for the final cut."""
        import torch.distributed as dist
        if self._group_size() <= 1 or self._refill is None:
            return self.shard_weights
        n = self.n
        rng = np.random.default_rng(5)
        d = rng.standard_normal((n, n))
        d = torch.from_numpy(d + d.T).to(self.device)
        for _ in range(rounds):
            for _ in range(8):
                self.jk_device(d, True)
            torch.cuda.synchronize()
            self.timings()
            for _ in range(steps):
                self.jk_device(d, True)
            torch.cuda.synchronize()
            ms = float(self.timings()["jk_kernel"])
            t = torch.tensor([ms], dtype=torch.float64, device=self.device)
            allt = [torch.zeros_like(t) for _ in range(self.world)]
            dist.all_gather(allt, t, group=self.group)
            ms_r = np.array([float(x[0]) for x in allt])
            w_old = np.ones(self.world) if self.shard_weights is None else np.asarray(self.shard_weights)
            speed = w_old / np.maximum(ms_r, 1e-6)     # cost share of the current cut over the time it took
            speed = speed / speed.mean()
            if self.shard_weights is not None and damping < 1.0:
                speed = damping * speed + (1.0 - damping) * np.asarray(self.shard_weights)
            self.shard_weights = [float(v) for v in speed]
            self._refill()
        return self.shard_weights

    # ------------------------------------------------------------------ on-disk cache of the packed shard
    ERI_FILE_MAGIC = "sqcc-b200 native-v2 ERI shard"

    def save_eri(self, path):
        """Write this rank's packed shard (native v2 layout, exactly the bytes the kernels stream) to `path` (+ `.json`
        header).  One file per rank: a later run of the same (n, world, rank, shard cut) attaches it with attach_eri_file
        instead of asking the integral provider again (the reference recomputes the N^4 tensor on every run,
        interface_psi4.py:187-190)."""
        import json
        hdr = {"magic": self.ERI_FILE_MAGIC, "n": self.n, "world": self.world, "rank": self.rank,
               "t_begin": self.plan.t_begin, "t_end": self.plan.t_end, "native_len": self.plan.native_len,
               "shard_weights": self.shard_weights, "dtype": "<f8"}
        out = np.lib.format.open_memmap(path, mode="w+", dtype=np.float64, shape=(self.plan.native_len,))
        chunk = 32 << 20
        stage = torch.empty(chunk, dtype=torch.float64, pin_memory=True)
        for lo in range(0, self.plan.native_len, chunk):
            hi = min(self.plan.native_len, lo + chunk)
            stage[:hi - lo].copy_(self.eri[lo:hi])
            torch.cuda.current_stream().synchronize()
            out[lo:hi] = stage[:hi - lo].numpy()
        out.flush()
        del out
        with open(path + ".json", "w") as f:
            json.dump(hdr, f)

    def attach_eri_file(self, path):
        """Attach a shard written by save_eri (same n / world / rank / shard cut): memory-mapped, streamed to the device
        through a pinned staging buffer."""
        import json
        hdr = json.load(open(path + ".json"))
        if hdr.get("magic") != self.ERI_FILE_MAGIC:
            raise _lib.SqccError(f"{path}: not a sqcc-b200 packed ERI shard")
        if (hdr["world"], hdr["rank"]) != (self.world, self.rank):
            raise _lib.SqccError(f"{path}: written by rank {hdr['rank']} of {hdr['world']}, this engine is rank {self.rank} of {self.world}")
        self.shard_weights = hdr.get("shard_weights")
        self._alloc_shard(hdr["n"])
        if (self.plan.t_begin, self.plan.t_end, self.plan.native_len) != (hdr["t_begin"], hdr["t_end"], hdr["native_len"]):
            raise _lib.SqccError(f"{path}: shard cut differs from this library's (layout / cost model changed)")
        src = np.load(path, mmap_mode="r")
        chunk = 32 << 20
        stage = torch.empty(chunk, dtype=torch.float64, pin_memory=True)
        for lo in range(0, self.plan.native_len, chunk):
            hi = min(self.plan.native_len, lo + chunk)
            stage[:hi - lo].numpy()[...] = src[lo:hi]
            self.eri[lo:hi].copy_(stage[:hi - lo], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        self._refill = None
        return self

    def eri_bytes(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.sqcc_eri_bytes(self.ctx, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def eri_full(self):
        """Expand the (unsharded) packed tensor back to the dense [N,N,N,N] host array."""
        n = self.n
        out = torch.empty((n, n, n, n), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sqcc_eri_unpack_full(self.ctx, _ptr(out)))
        return out.cpu().numpy()

    def set_jk_variant(self, variant):
        _lib.check(self.lib.sqcc_jk_set_variant(self.ctx, int(variant)))

    def set_deterministic(self, on=True):
        """Deterministic J/K (include/sqcc_b200.h: sqcc_jk_set_deterministic): fixed-point integer accumulation, results
        bit-for-bit reproducible from run to run and identical on every rank of a peer group.  With several ranks the
        bound on the stored integrals is agreed once (max over the shards); call again after attaching a new tensor."""
        _lib.check(self.lib.sqcc_jk_set_deterministic(self.ctx, int(bool(on))))
        self._deterministic = bool(on)
        if on and self._group_size() > 1 and self.eri is not None:
            import torch.distributed as dist
            m = torch.tensor([float(self.eri.abs().max().item())], dtype=torch.float64, device=self.device)
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
            v = ctypes.c_double(float(m.item()))
            _lib.check(self.lib.sqcc_eri_absmax(self.ctx, ctypes.byref(v), 1))

    # ------------------------------------------------------------------ J/K
    def jk_device(self, d, want_k=True, out=None):
        """d: cuda float64 [N,N] or [2,N,N].  Returns the stacked device buffer [1+nspin, N, N] = [J; K...]."""
        nspin = 1 if d.dim() == 2 else 2
        n = self.n
        if out is None:
            out = torch.empty((1 + (nspin if want_k else 0), n, n), dtype=torch.float64, device=self.device)
        self._use_stream()
        if self._p2p:
            self.comm_check()   # mapped host word, no sync: a failed exchange of an EARLIER build surfaces here
        _lib.check(self.lib.sqcc_jk_build(self.ctx, _ptr(d), nspin, int(want_k), _ptr(out[0]),
                                          _ptr(out[1]) if want_k else None))
        if not self._p2p:
            reduce_partials(out, self.group)
        return out

    def set_exchange(self, exchange):
        """Switch between the fused peer-memory exchange ("p2p") and finalize + NCCL all-reduce ("nccl") without
        re-attaching the shard (A/B measurements)."""
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        if exchange == "p2p" and not self._p2p and self._group_size() > 1:
            if self._mapped:
                _lib.check(self.lib.sqcc_comm_set_enabled(self.ctx, 1))
                self._p2p = True
            else:
                self._attach_peers()
        elif exchange == "nccl" and self._p2p:
            _lib.check(self.lib.sqcc_comm_set_enabled(self.ctx, 0))
            self._p2p = False
        self.exchange = exchange

    def comm_check(self):
        """Raises if a peer barrier of the fused exchange timed out (reads a mapped host word; no sync).  After a
        failure every rank's J/K of that build are NaN; comm_reset() clears the sticky status."""
        return _lib.check(self.lib.sqcc_comm_status(self.ctx))

    def comm_reset(self):
        _lib.check(self.lib.sqcc_comm_reset_status(self.ctx))

    def pinned(self, *shape):
        """A page-locked float64 numpy array (torch owns the allocation): buffers handed to jk() that are pinned
        are copied from / to directly, without the library's staging copy."""
        t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        a = t.numpy()
        self._pinned_keep = getattr(self, "_pinned_keep", []) + [t]
        return a

    def jk(self, d, want_k=True, out=None):
        """Host in / host out, reference semantics (hf_ksdft.py:483-495, :510-535).

        d [N,N] -> (J, K);  d [2,N,N] -> (J(d0+d1), K[2,N,N]);  K is None when want_k is False.
        out: optional (J [N,N], K [nspin,N,N]) float64 arrays to fill (e.g. from pinned()).
        """
        d = np.ascontiguousarray(d, dtype=np.float64)
        nspin = 1 if d.ndim == 2 else 2
        n = self.n
        if self.world == 1 or self._p2p:
            j = np.empty((n, n)) if out is None else out[0]
            k = (np.empty((nspin, n, n)) if out is None else out[1]) if want_k else None
            self._use_stream()
            _lib.check(self.lib.sqcc_jk_build_host(
                self.ctx, d.ctypes.data_as(ctypes.c_void_p), nspin, int(want_k),
                j.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p) if want_k else None))
            if self._p2p:
                self.comm_check()
        else:
            res = self.jk_device(torch.from_numpy(d).to(self.device), want_k).cpu().numpy()
            j, k = res[0], (res[1:] if want_k else None)
            if out is not None:   # fill the caller's buffers on this branch too
                out[0][...] = j
                j = out[0]
                if want_k:
                    out[1][...] = k
                    k = out[1]
        if want_k and d.ndim == 2:
            k = k[0]
        return j, k

    # ------------------------------------------------------------------ device-resident SCF iteration (csrc/scf.cu)
    def scf_setup(self, h_core, orthogonalizer, nspin, n_occ_a, n_occ_b, ksdft):
        h = np.ascontiguousarray(h_core, dtype=np.float64)
        x = np.ascontiguousarray(orthogonalizer, dtype=np.float64)
        self._use_stream()
        _lib.check(self.lib.sqcc_scf_setup(self.ctx, h.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p),
                                           int(nspin), int(n_occ_a), int(n_occ_b), int(bool(ksdft))))
        self._scf_nspin = int(nspin)

    def scf_set_density(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        _lib.check(self.lib.sqcc_scf_set_density(self.ctx, d.ctypes.data_as(ctypes.c_void_p)))

    def scf_fock_energy(self):
        """Fock build for the resident density; returns the electronic energy (the only per-iteration D2H).
        On a sharded tensor without the fused peer exchange the partial [J; K] are all-reduced here (NCCL) between the
        two halves of the step, so the Fock matrix never sees a rank's partial matrices."""
        e = ctypes.c_double()
        self._use_stream()
        if self._group_size() > 1 and not self._p2p:
            ptr, cnt = ctypes.c_void_p(), ctypes.c_int64()
            _lib.check(self.lib.sqcc_scf_jk(self.ctx, ctypes.byref(ptr), ctypes.byref(cnt)))
            reduce_partials(_device_view(ptr.value, cnt.value, self.device), self.group)
            _lib.check(self.lib.sqcc_scf_finish(self.ctx, ctypes.byref(e)))
        else:
            _lib.check(self.lib.sqcc_scf_fock_energy(self.ctx, ctypes.byref(e)))
        return e.value

    def scf_update(self):
        self._use_stream()
        _lib.check(self.lib.sqcc_scf_update(self.ctx))

    def scf_get(self):
        """(density [nspin,N,N], C [nspin,N,N] AO x MO, eps [nspin,N]) from the device."""
        n, ns = self.n, self._scf_nspin
        d, ct, eps = np.empty((ns, n, n)), np.empty((ns, n, n)), np.empty((ns, n))
        _lib.check(self.lib.sqcc_scf_get(self.ctx, d.ctypes.data_as(ctypes.c_void_p), ct.ctypes.data_as(ctypes.c_void_p),
                                         eps.ctypes.data_as(ctypes.c_void_p)))
        return d, np.ascontiguousarray(np.transpose(ct, (0, 2, 1))), eps

    # ------------------------------------------------------------------ LDA grid
    def attach_grid(self, phi, w):
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        self._phi = torch.from_numpy(phi).to(self.device)
        self._w = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float64)).to(self.device)
        _lib.check(self.lib.sqcc_grid_attach(self.ctx, _ptr(self._phi), _ptr(self._w), phi.shape[0], phi.shape[1]))
        return self

    def lda_vxc(self, d, want_rho=False):
        """Rows G+H+I fused: d [N,N] -> (Vxc [N,N], Exc, rho [G]|None); d [2,N,N] -> ([2,N,N], Exc, [2,G]|None)."""
        d = np.ascontiguousarray(d, dtype=np.float64)
        nspin = 1 if d.ndim == 2 else 2
        n, g = self._phi.shape[1], self._phi.shape[0]
        v = np.empty((nspin, n, n))
        exc = ctypes.c_double()
        rho = np.empty((nspin, g)) if want_rho else None
        self._use_stream()
        _lib.check(self.lib.sqcc_lda_vxc_host(self.ctx, d.ctypes.data_as(ctypes.c_void_p), nspin,
                                              v.ctypes.data_as(ctypes.c_void_p), ctypes.byref(exc),
                                              rho.ctypes.data_as(ctypes.c_void_p) if want_rho else None))
        if d.ndim == 2:
            return v[0], exc.value, (rho[0] if want_rho else None)
        return v, exc.value, rho

    def grid_quadrature(self, v):
        """V_pq = sum_n phi_np (w_n v_n) phi_nq for a caller-supplied potential on the attached points."""
        n = self._phi.shape[1]
        vt = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(self.device)
        out = torch.empty((n, n), dtype=torch.float64, device=self.device)
        self._use_stream()
        _lib.check(self.lib.sqcc_grid_quadrature(self.ctx, _ptr(vt), _ptr(out)))
        return out.cpu().numpy()

    def grid_rho(self, d):
        dt = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float64)).to(self.device)
        out = torch.empty(self._phi.shape[0], dtype=torch.float64, device=self.device)
        self._use_stream()
        _lib.check(self.lib.sqcc_grid_rho(self.ctx, _ptr(dt), _ptr(out)))
        return out.cpu().numpy()

    # ------------------------------------------------------------------ MP2 / AO->MO
    def mp2_corr(self, c, eps, n_occ, want_iajb=False):
        c = np.ascontiguousarray(c, dtype=np.float64)
        eps = np.ascontiguousarray(eps, dtype=np.float64)
        n = self.n
        nv = n - n_occ
        e = ctypes.c_double()
        iajb = np.empty((n_occ, nv, n_occ, nv)) if want_iajb else None
        self._use_stream()
        _lib.check(self.lib.sqcc_mp2_host(self.ctx, c.ctypes.data_as(ctypes.c_void_p),
                                          eps.ctypes.data_as(ctypes.c_void_p), int(n_occ), ctypes.byref(e),
                                          iajb.ctypes.data_as(ctypes.c_void_p) if want_iajb else None))
        return (e.value, iajb) if want_iajb else e.value

    def ao2mo(self, c1, c2, c3, c4):
        """(pq|rs) -> (ab|cd) for four column blocks of one [N,N] AO x MO matrix given as (matrix, lo, hi)."""
        n = self.n
        mats = []
        for c in (c1, c2, c3, c4):
            m = np.zeros((n, n))
            m[:, :c.shape[1]] = c
            mats.append(torch.from_numpy(m).to(self.device))
        dims = [c.shape[1] for c in (c1, c2, c3, c4)]
        out = torch.empty(dims, dtype=torch.float64, device=self.device)
        self._use_stream()
        _lib.check(self.lib.sqcc_ao2mo(self.ctx, _ptr(mats[0]), dims[0], _ptr(mats[1]), dims[1], _ptr(mats[2]),
                                       dims[2], _ptr(mats[3]), dims[3], _ptr(out)))
        return out.cpu().numpy()

    # ------------------------------------------------------------------ instrumentation
    def timings(self):
        buf = (ctypes.c_double * 8)()
        _lib.check(self.lib.sqcc_get_timings(self.ctx, buf, 8))
        names = ["jk_kernel", "jk_finalize", "rho", "vxc", "mp2_transform", "mp2_energy", "scf_update", "t7"]
        return {k: buf[i] for i, k in enumerate(names)}

    def launch_count(self):
        return int(self.lib.sqcc_launch_count(self.ctx))

    def synchronize(self):
        _lib.check(self.lib.sqcc_ctx_synchronize(self.ctx))
        if self._p2p:
            self.comm_check()


class FockGroup:
    """Several shard contexts driven by ONE process: one engine per rank, on distinct devices
    or -- for single-GPU tests of the exchange -- all on the same device.  The partial J/K of the ranks are summed by
    the fused peer-memory kernel (csrc/comm.cu); with a shared device that is a single launch covering all ranks."""

    def __init__(self, devices):
        self.engines = [FockEngine(device=d, rank=r, world=len(devices), exchange="nccl") for r, d in enumerate(devices)]
        self.world = len(devices)
        self.same_device = len(set(int(d) for d in devices)) == 1
        self.lib = self.engines[0].lib

    def attach_eri_synth(self, n, seed):
        for e in self.engines:
            e.attach_eri_synth(n, seed)
        arr = (ctypes.c_void_p * self.world)(*[e.ctx for e in self.engines])
        _lib.check(self.lib.sqcc_comm_attach_local(arr, self.world))
        return self

    def jk_device(self, d_host, want_k=True):
        """d_host: numpy [N,N] or [2,N,N].  Returns one stacked device buffer [J; K...] per rank (all identical)."""
        d_host = np.ascontiguousarray(d_host, dtype=np.float64)
        nspin = 1 if d_host.ndim == 2 else 2
        n = self.engines[0].n
        outs, ds = [], []
        for e in self.engines:
            with torch.cuda.device(e.device):
                ds.append(torch.from_numpy(d_host).to(e.device))
                outs.append(torch.empty((1 + (nspin if want_k else 0), n, n), dtype=torch.float64, device=e.device))
        if self.same_device:
            for e, d in zip(self.engines, ds):
                e._use_stream()
                _lib.check(self.lib.sqcc_jk_partial(e.ctx, _ptr(d), nspin, int(want_k)))
            ctxs = (ctypes.c_void_p * self.world)(*[e.ctx for e in self.engines])
            js = (ctypes.c_void_p * self.world)(*[o[0].data_ptr() for o in outs])
            ks = (ctypes.c_void_p * self.world)(*[(o[1].data_ptr() if want_k else None) for o in outs])
            _lib.check(self.lib.sqcc_jk_reduce_local(ctxs, self.world, nspin, int(want_k), js, ks))
        else:
            for e, d, o in zip(self.engines, ds, outs):
                with torch.cuda.device(e.device):
                    e._use_stream()
                    _lib.check(self.lib.sqcc_jk_build(e.ctx, _ptr(d), nspin, int(want_k), _ptr(o[0]),
                                                      _ptr(o[1]) if want_k else None))
        for e in self.engines:
            e.synchronize()
            e.comm_check()
        return outs

    def close(self):
        for e in self.engines:
            self.lib.sqcc_comm_detach(e.ctx)
        for e in self.engines:
            e.close()
