"""GPU parity of the fused J/K build against the oracle (hf_ksdft.py:483-495, :510-535).

Tolerance from BASELINE.json north_star: J/K within 1e-11 max-abs.
"""
import numpy as np
import pytest

from oracle import cjk, jk_ref, packing as pk, synth

pytestmark = pytest.mark.gpu
TOL = 1e-11


def _check(eng, n, seed, d, want_k=True, variant=0):
    eng.attach_eri_synth(n, seed)
    eng.set_jk_variant(variant)
    j, k = eng.jk(d, want_k=want_k)
    eng.set_jk_variant(0)
    packed = cjk.synth_canonical(n, seed)
    jr, kr = cjk.jk_entries(cjk.SRC_PACKED, packed, n, seed, d, want_k=want_k)
    assert np.abs(j - jr).max() <= TOL, np.abs(j - jr).max()
    if want_k:
        assert np.abs(k - kr).max() <= TOL, np.abs(k - kr).max()
    else:
        assert k is None
    return j, k


@pytest.mark.parametrize("n", [1, 2, 3, 5, 6, 10, 17, 36, 62, 65, 100])
def test_rhf_jk_matches_oracle(engine, n):
    d = synth.synth_density_rhf(n, max(1, n // 4), 11 + n)
    _check(engine, n, 20261019 + n, d)


@pytest.mark.parametrize("n", [2, 7, 10, 36, 62, 67])
def test_uhf_jk_matches_oracle(engine, n):
    d = synth.synth_density_uhf(n, max(1, n // 3), max(1, n // 4), 5 + n)
    _check(engine, n, 20261019 + n, d)


@pytest.mark.parametrize("n", [4, 10, 43, 62])
def test_j_only_matches_oracle(engine, n):
    d = synth.synth_density_rhf(n, max(1, n // 4), 3 + n)
    _check(engine, n, 77 + n, d, want_k=False)
    du = synth.synth_density_uhf(n, max(1, n // 3), max(1, n // 4), 5 + n)
    _check(engine, n, 77 + n, du, want_k=False)


@pytest.mark.parametrize("variant", [1, 17, 18, 22, 23])
@pytest.mark.parametrize("n", [6, 36, 70, 131])
def test_other_kernel_variants_match_oracle(engine, n, variant):
    """variant 1: simple atomics kernel; 16+c: tuning configurations of the streaming kernel kept for A/B runs --
    1 cp.async ring for every piece (no TMA), 2 no L2 cache-policy hints, 6 spin-split UHF pass, 7 single pass (no
    narrow-piece pass)."""
    d = synth.synth_density_rhf(n, max(1, n // 4), 1)
    _check(engine, n, 5, d, variant=variant)
    du = synth.synth_density_uhf(n, max(1, n // 3), max(1, n // 4), 2)
    _check(engine, n, 5, du, variant=variant)
    _check(engine, n, 5, d, want_k=False, variant=variant)


@pytest.mark.parametrize("n", [10, 70, 150])
def test_deterministic_mode_is_bitwise_reproducible(engine, n):
    """with fixed-point integer accumulation two builds of the same J/K are bit-identical
    (the floating-point atomics of the default mode are not, once several warps flush into one element) and the result
    still meets the parity gate."""
    seed = 31 + n
    engine.attach_eri_synth(n, seed)
    packed = cjk.synth_canonical(n, seed)
    for d in (synth.synth_density_rhf(n, max(1, n // 4), 2), synth.synth_density_uhf(n, max(1, n // 3), max(1, n // 4), 2)):
        jr, kr = cjk.jk_packed(packed, n, d)
        engine.set_deterministic(True)
        runs = [engine.jk(d) for _ in range(3)]
        engine.set_deterministic(False)
        for j, k in runs[1:]:
            assert np.array_equal(j, runs[0][0]) and np.array_equal(k, runs[0][1])
        assert np.abs(runs[0][0] - jr).max() <= TOL and np.abs(runs[0][1] - kr).max() <= TOL
        jo, _ = engine.jk(d)
        assert np.abs(jo - jr).max() <= TOL


def test_reference_loop_form_small(engine):
    """Literal loop form of the reference (numpy) on the full tensor, N=10 (N2/STO-3G size)."""
    n, seed = 10, 99
    eri = synth.synth_eri_full(n, seed)
    d = synth.synth_density_rhf(n, 7, 4)
    jr, kr = jk_ref.jk_loop(eri, d)
    engine.attach_eri_full(eri)
    j, k = engine.jk(d)
    assert np.abs(j - jr).max() <= TOL and np.abs(k - kr).max() <= TOL
    du = synth.synth_density_uhf(n, 8, 6, 4)
    jr, kr = jk_ref.jk_loop(eri, du)
    j, k = engine.jk(du)
    assert np.abs(j - jr).max() <= TOL and np.abs(k - kr).max() <= TOL


def test_pack_paths_agree(engine):
    """Full-tensor pack, canonical pack and on-device synthesis give the same native tensor (bit-exact)."""
    n, seed = 21, 1234
    packed = synth.synth_eri_canonical(n, seed)
    native_ref = pk.canonical_to_native(packed, n)
    engine.attach_eri_synth(n, seed)
    a = engine.eri.cpu().numpy()[: len(native_ref)]
    engine.attach_eri_canonical(packed, n)
    b = engine.eri.cpu().numpy()[: len(native_ref)]
    engine.attach_eri_full(pk.unpack_canonical(packed, n))
    c = engine.eri.cpu().numpy()[: len(native_ref)]
    assert np.array_equal(a, native_ref) and np.array_equal(b, native_ref) and np.array_equal(c, native_ref)
    full = engine.eri_full()
    assert np.array_equal(full, pk.unpack_canonical(packed, n))


def test_benzene_size_222(engine):
    """configs[1] size: N=222 (benzene def2-TZVP), full J/K vs the C oracle on the packed tensor."""
    n, seed = 222, 20261019
    d = synth.synth_density_rhf(n, 21, 7)
    engine.attach_eri_synth(n, seed)
    j, k = engine.jk(d)
    packed = cjk.synth_canonical(n, seed)
    jr, kr = cjk.jk_packed(packed, n, d)
    assert np.abs(j - jr).max() <= TOL and np.abs(k - kr).max() <= TOL
    # literal reference loop nest on sampled entries incl. degeneracy classes
    ent = [(0, 0), (221, 221), (221, 0), (100, 100), (37, 150), (150, 37), (1, 0), (221, 220)]
    je, ke = cjk.jk_entries(cjk.SRC_PACKED, packed, n, seed, d, ent)
    for t, (p, q) in enumerate(ent):
        assert abs(j[p, q] - je[t]) <= TOL and abs(k[p, q] - ke[t]) <= TOL
    # size-independent properties: symmetry and linearity
    assert np.abs(j - j.T).max() <= TOL and np.abs(k - k.T).max() <= TOL
    j2, k2 = engine.jk(2.5 * d)
    assert np.abs(j2 - 2.5 * j).max() <= 5 * TOL and np.abs(k2 - 2.5 * k).max() <= 5 * TOL
    # UHF with Da = Db = D/2 reproduces RHF: J(D), K(D/2) = K(D)/2 (tests/hf/n2_singlet vs .../unrestricted)
    ju, ku = engine.jk(np.stack([0.5 * d, 0.5 * d]))
    assert np.abs(ju - j).max() <= TOL and np.abs(2.0 * ku[0] - k).max() <= TOL
    assert np.abs(ku[0] - ku[1]).max() <= TOL
    # dual-density pass with genuinely different spin densities (hf_ksdft.py:510-535), whole matrices
    du = synth.synth_density_uhf(n, 23, 19, 3)
    ju, ku = engine.jk(du)
    jr, kr = cjk.jk_packed(packed, n, du)
    assert np.abs(ju - jr).max() <= TOL and np.abs(ku - kr).max() <= TOL
    je, ke = cjk.jk_entries(cjk.SRC_PACKED, packed, n, seed, du, ent)
    for t, (p, q) in enumerate(ent):
        assert abs(ju[p, q] - je[t]) <= TOL and abs(ku[0][p, q] - ke[0][t]) <= TOL and abs(ku[1][p, q] - ke[1][t]) <= TOL
    jo, ko = engine.jk(du, want_k=False)
    assert ko is None and np.abs(jo - jr).max() <= TOL


def test_benzene_size_264(engine):
    """The basis size BASELINE.json literally names for configs[1] (~264): RHF and UHF (Da != Db) whole matrices vs the
    C oracle on the canonical packed tensor, plus reference-loop-nest entries."""
    n, seed = 264, 20261019
    d = synth.synth_density_rhf(n, 21, 7)
    du = synth.synth_density_uhf(n, 22, 20, 9)
    engine.attach_eri_synth(n, seed)
    j, k = engine.jk(d)
    ju, ku = engine.jk(du)
    packed = cjk.synth_canonical(n, seed)
    jr, kr = cjk.jk_packed(packed, n, d)
    assert np.abs(j - jr).max() <= TOL and np.abs(k - kr).max() <= TOL
    jr, kr = cjk.jk_packed(packed, n, du)
    assert np.abs(ju - jr).max() <= TOL and np.abs(ku - kr).max() <= TOL
    ent = [(0, 0), (263, 263), (263, 0), (131, 132), (17, 250), (250, 17), (1, 0), (263, 262)]
    je, ke = cjk.jk_entries(cjk.SRC_PACKED, packed, n, seed, d, ent)
    for t, (p, q) in enumerate(ent):
        assert abs(j[p, q] - je[t]) <= TOL and abs(k[p, q] - ke[t]) <= TOL
    engine.attach_eri_synth(4, 1)


@pytest.mark.parametrize("n,world", [(10, 2), (37, 3), (70, 8), (5, 8)])
def test_sharded_partials_sum_to_full(n, world):
    """Pair-block sharding on ONE device: every rank's engine holds its tile range, builds its partial [J;K];
    the partials must add up to the unsharded result (what the NCCL all-reduce produces on N GPUs)."""
    from sqcc_b200.backend import FockEngine
    seed = 99 + n
    d = synth.synth_density_uhf(n, max(1, n // 3), max(1, n // 4), 5)
    packed = cjk.synth_canonical(n, seed)
    jr, kr = cjk.jk_packed(packed, n, d)
    jsum, ksum = np.zeros((n, n)), np.zeros((2, n, n))
    for rank in range(world):
        eng = FockEngine(rank=rank, world=world)     # no process group: reduce_partials is the identity
        eng.attach_eri_synth(n, seed)
        out = eng.jk_device(__import__("torch").from_numpy(d).to(eng.device)).cpu().numpy()
        jsum += out[0]
        ksum += out[1:]
        # each shard's partial equals the oracle's partial for the same rows
        jp, kp = cjk.jk_packed_segments(n, seed, eng.plan.segments(), d)
        assert np.abs(out[0] - jp).max() <= TOL and np.abs(out[1:] - kp).max() <= TOL
        eng.close()
    assert np.abs(jsum - jr).max() <= TOL and np.abs(ksum - kr).max() <= TOL


@pytest.mark.parametrize("n,world", [(13, 1), (30, 1), (23, 3)])
def test_direct_block_ingest_equals_full_tensor_pack(n, world):
    """dense blocks per canonical (ragged, fake) shell quartet go straight into the packed shard;
    the resulting native buffer must be bit-identical to the one packed from the full [N,N,N,N] tensor."""
    from oracle import packing as pk
    from sqcc_b200.backend import FockEngine
    full = pk.unpack_canonical(cjk.synth_canonical(n, 5 + n), n)
    sizes, pat, t = [], [1, 3, 2, 5], 0
    while sum(sizes) < n:
        sizes.append(min(pat[t % 4], n - sum(sizes)))
        t += 1
    first = np.concatenate([[0], np.cumsum(sizes)]).astype(int)

    def blocks():
        ns = len(sizes)
        for m in range(ns):
            for nn in range(m + 1):
                for p in range(m + 1):
                    for q in range((p if p < m else nn) + 1):
                        yield (first[m], first[nn], first[p], first[q],
                               full[first[m]:first[m + 1], first[nn]:first[nn + 1], first[p]:first[p + 1], first[q]:first[q + 1]])

    for rank in range(world):
        a, b = FockEngine(rank=rank, world=world), FockEngine(rank=rank, world=world)
        a.attach_eri_full(full)
        b.attach_eri_blocks(n, blocks(), batch_doubles=4096)     # small batches: several flushes
        assert a.plan.native_len == b.plan.native_len     # (the buffers carry a few KB of unread slack behind the shard)
        m = a.plan.native_len
        assert np.array_equal(a.eri[:m].cpu().numpy(), b.eri[:m].cpu().numpy())
        a.close()
        b.close()


@pytest.mark.parametrize("n,world", [(23, 3), (41, 5), (9, 4)])
def test_sharded_ingest_from_host_tensors(n, world):
    """Row A on a sharded engine: every rank packs ITS tile-range shard from the full [N,N,N,N] host tensor
    (interface_psi4.py:177-190 output; slabs through pinned staging) and from the canonical packed vector; the shard's
    partial J/K must equal the oracle's partial for the same rows, and the partials must add up to the full J/K."""
    from oracle import packing as pk
    from sqcc_b200.backend import FockEngine
    seed = 17 + n
    packed = cjk.synth_canonical(n, seed)
    full = pk.unpack_canonical(packed, n)
    d = synth.synth_density_uhf(n, max(1, n // 3), max(1, n // 4), 3)
    jr, kr = cjk.jk_packed(packed, n, d)
    for mode in ("full", "canonical"):
        jsum, ksum = np.zeros((n, n)), np.zeros((2, n, n))
        for rank in range(world):
            eng = FockEngine(rank=rank, world=world)
            if mode == "full":
                eng.attach_eri_full(full)
            else:
                eng.attach_eri_canonical(packed, n)
            out = eng.jk_device(__import__("torch").from_numpy(d).to(eng.device)).cpu().numpy()
            jp, kp = cjk.jk_packed_segments(n, seed, eng.plan.segments(), d)
            assert np.abs(out[0] - jp).max() <= TOL and np.abs(out[1:] - kp).max() <= TOL, (mode, rank)
            jsum += out[0]
            ksum += out[1:]
            eng.close()
        assert np.abs(jsum - jr).max() <= TOL and np.abs(ksum - kr).max() <= TOL


@pytest.mark.parametrize("n,world,uhf", [(10, 2, False), (37, 3, True), (70, 8, True), (62, 8, False), (5, 8, False), (222, 4, False)])
def test_fused_peer_exchange_single_launch(n, world, uhf):
    """The fused exchange kernel (csrc/comm.cu: flag barriers in peer memory, slice-wise reduce + broadcast, local
    symmetrisation) with all ranks' contexts on ONE device, emulated as a single cooperative launch (blockIdx.y =
    rank).  Every rank must end with the full J/K, bit-identical across ranks, equal to the oracle."""
    from sqcc_b200.backend import FockGroup
    seed = 7 + n
    d = synth.synth_density_uhf(n, max(1, n // 3), max(1, n // 4), 5) if uhf else synth.synth_density_rhf(n, max(1, n // 3), 5)
    if n <= 80:
        jr, kr = cjk.jk_packed(cjk.synth_canonical(n, seed), n, d)
    grp = FockGroup([0] * world).attach_eri_synth(n, seed)
    for rep in range(3):   # repeated builds exercise the epoch-numbered barriers and buffer reuse
        outs = [o.cpu().numpy() for o in grp.jk_device(d if rep != 1 else 0.5 * d)]
        scale = 0.5 if rep == 1 else 1.0
        for o in outs[1:]:
            assert np.array_equal(o, outs[0])
        if n <= 80:
            k = outs[0][1:] if uhf else outs[0][1]
            assert np.abs(outs[0][0] - scale * jr).max() <= TOL and np.abs(k - scale * kr).max() <= TOL
    if n > 80:   # sampled entries by the reference loop nest
        ent = [(0, 0), (n - 1, n - 1), (n - 1, 0), (n // 2, n // 3), (1, 0), (n - 1, n - 2), (17, 150), (150, 17)]
        je, ke = cjk.jk_entries(cjk.SRC_SYNTH, None, n, seed, d, ent)
        for t, (p, q) in enumerate(ent):
            assert abs(outs[0][0][p, q] - je[t]) <= TOL and abs(outs[0][1][p, q] - ke[t]) <= TOL
    jo = grp.jk_device(d, want_k=False)[0].cpu().numpy()
    assert np.abs(jo[0] - outs[0][0]).max() <= TOL
    grp.close()


def test_packed_shard_file_cache(tmp_path):
    """the packed shard written to disk and attached again gives the same tensor and the same J/K
    (unsharded and as rank 1 of 3)."""
    from sqcc_b200.backend import FockEngine
    n, seed = 45, 8
    d = synth.synth_density_uhf(n, 11, 9, 2)
    for rank, world in ((0, 1), (1, 3)):
        a = FockEngine(rank=rank, world=world)
        a.attach_eri_synth(n, seed)
        ref = a.jk_device(__import__("torch").from_numpy(d).to(a.device)).cpu().numpy()
        path = str(tmp_path / f"eri_{rank}_{world}.npy")
        a.save_eri(path)
        b = FockEngine(rank=rank, world=world)
        b.attach_eri_file(path)
        m = a.plan.native_len
        assert np.array_equal(a.eri[:m].cpu().numpy(), b.eri[:m].cpu().numpy())
        got = b.jk_device(__import__("torch").from_numpy(d).to(b.device)).cpu().numpy()
        assert np.abs(got - ref).max() <= TOL
        with pytest.raises(Exception):
            FockEngine(rank=0, world=2).attach_eri_file(path)
        a.close()
        b.close()


def test_large_basis_700_sharded_on_one_gpu():
    """N = 700 (round 1 refused N > 680: 32-bit row offsets): the 250 GB native tensor does not fit one B200, so its four
    tile-range shards are attached one after the other on the same device and their partial J/K summed -- what four
    ranks would all-reduce -- then checked on sampled entries against the reference loop nest on the lazily generated
    synthetic tensor."""
    import torch
    from sqcc_b200.backend import FockEngine
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~65 GB of device memory")
    n, seed, world = 700, 20261019, 4
    d = synth.synth_density_rhf(n, 60, 3)
    dd = torch.from_numpy(d).cuda()
    total = torch.zeros((2, n, n), dtype=torch.float64, device="cuda")
    for rank in range(world):
        eng = FockEngine(rank=rank, world=world)
        eng.attach_eri_synth(n, seed)
        total += eng.jk_device(dd)
        eng.close()
        del eng
        torch.cuda.empty_cache()
    res = total.cpu().numpy()
    rng = np.random.default_rng(1)
    ent = [(0, 0), (699, 699), (699, 0), (350, 351), (681, 690), (1, 0)] + [(int(a), int(b)) for a, b in rng.integers(0, n, size=(6, 2))]
    je, ke = cjk.jk_entries(cjk.SRC_SYNTH, None, n, seed, d, ent)
    for t, (p, q) in enumerate(ent):
        assert abs(res[0][p, q] - je[t]) <= TOL and abs(res[1][p, q] - ke[t]) <= TOL, (p, q)
    assert np.abs(res[0] - res[0].T).max() <= TOL and np.abs(res[1] - res[1].T).max() <= TOL


def test_anthracene_size_494(engine):
    """configs[4] size: N=494 (59.8 GB of unique integrals).  The full tensor cannot exist on the host, so parity
    is checked on sampled (p,q) entries evaluated with the reference loop nest on the lazily generated synthetic
    tensor (incl. all degeneracy classes), plus size-independent properties: symmetry, linearity, UHF(D/2,D/2)=RHF."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~65 GB of device memory")
    n, seed = 494, 20261019
    d = synth.synth_density_rhf(n, 47, 7)
    engine.attach_eri_synth(n, seed)
    j, k = engine.jk(d)
    rng = np.random.default_rng(0)
    ent = [(0, 0), (493, 493), (493, 0), (0, 493), (247, 247), (100, 300), (300, 100), (493, 492), (1, 0), (2, 2)]
    ent += [(int(a), int(b)) for a, b in rng.integers(0, n, size=(22, 2))]
    je, ke = cjk.jk_entries(cjk.SRC_SYNTH, None, n, seed, d, ent)
    for t, (p, q) in enumerate(ent):
        assert abs(j[p, q] - je[t]) <= TOL and abs(k[p, q] - ke[t]) <= TOL, (p, q)
    assert np.abs(j - j.T).max() <= TOL and np.abs(k - k.T).max() <= TOL
    j2, k2 = engine.jk(-0.5 * d)
    assert np.abs(j2 + 0.5 * j).max() <= TOL and np.abs(k2 + 0.5 * k).max() <= TOL
    ju, ku = engine.jk(np.stack([0.5 * d, 0.5 * d]))
    assert np.abs(ju - j).max() <= TOL and np.abs(2.0 * ku[0] - k).max() <= TOL and np.abs(ku[0] - ku[1]).max() <= TOL
    jo, ko = engine.jk(d, want_k=False)
    assert ko is None and np.abs(jo - j).max() <= TOL
    # dual-density pass with Da != Db (hf_ksdft.py:510-535): sampled entries by the reference loop nest
    du = synth.synth_density_uhf(n, 49, 45, 11)
    ju, ku = engine.jk(du)
    je, ke = cjk.jk_entries(cjk.SRC_SYNTH, None, n, seed, du, ent)
    for t, (p, q) in enumerate(ent):
        assert abs(ju[p, q] - je[t]) <= TOL, (p, q)
        assert abs(ku[0][p, q] - ke[0][t]) <= TOL and abs(ku[1][p, q] - ke[1][t]) <= TOL, (p, q)
    assert np.abs(ju - ju.T).max() <= TOL and np.abs(ku - ku.transpose(0, 2, 1)).max() <= TOL
    assert np.abs(ku[0] - ku[1]).max() > 1e-3   # the two spin planes really differ
    engine.attach_eri_synth(4, 1)   # release the 62 GB shard for the following tests
