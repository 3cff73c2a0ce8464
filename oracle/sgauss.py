"""Psi4-free integral fixture: contracted s-type Gaussians.  TEST INFRASTRUCTURE.

Psi4 is not installed in this image, so the reference's ``interface_psi4.Psi4Interface``
(interface_psi4.py:65-406) cannot be constructed.  ``SGaussInterface`` exposes the same surface
(method names, argument meaning, array shapes, Angstrom input converted with 1/0.52917721067 like
interface_psi4.py:371 / hf_ksdft.py:221) backed by closed-form integrals over contracted s Gaussians
(Szabo & Ostlund, appendix A): S, T, V, (ab|cd) with the Boys function F0.  ``stub_module()``
returns a module object that can be registered as ``sys.modules["interface_psi4"]`` so the
UNMODIFIED reference ``hf_ksdft.Calculator.scf()`` runs on it (``oracle/reference_loader.py``), and
the same object can be handed to ``sqcc_b200.hf_ksdft`` so both sides see identical integrals.
"""
import types

import numpy as np
from scipy.special import erf

ANG_TO_BOHR = 1 / 0.52917721067

_D_STO3G = (0.15432897, 0.53532814, 0.44463454)
BASES = {
    "sto-3g": {
        1: [[(3.42525091, _D_STO3G[0]), (0.62391373, _D_STO3G[1]), (0.16885540, _D_STO3G[2])]],
        2: [[(6.36242139, _D_STO3G[0]), (1.15892300, _D_STO3G[1]), (0.31364979, _D_STO3G[2])]],
    },
    # split-valence s-only basis (two functions per atom): contracted pair + one diffuse primitive
    "dz-s": {
        1: [[(3.42525091, 0.2769344), (0.62391373, 0.8178421)], [(0.16885540, 1.0)]],
        2: [[(6.36242139, 0.2769344), (1.15892300, 0.8178421)], [(0.31364979, 1.0)]],
    },
    # three s functions per atom
    "tz-s": {
        1: [[(3.42525091, 1.0)], [(0.62391373, 1.0)], [(0.16885540, 1.0)]],
        2: [[(6.36242139, 1.0)], [(1.15892300, 1.0)], [(0.31364979, 1.0)]],
    },
}


def boys_f0(t):
    t = np.asarray(t, dtype=np.float64)
    small = t < 1e-8
    ts = np.where(small, 1.0, t)
    val = 0.5 * np.sqrt(np.pi / ts) * erf(np.sqrt(ts))
    return np.where(small, 1.0 - t / 3.0, val)


class _FakeBasisSet:
    """Minimal stand-in for psi4.core.BasisSet as used by hf_ksdft.py:737-750 (save_mo_data_to_json)."""

    class _Shell:
        def __init__(self, center):
            self.ncenter = center
            self.am = 0
            self.nfunction = 1

    def __init__(self, centers):
        self._shells = [self._Shell(c) for c in centers]

    def nshell(self):
        return len(self._shells)

    def shell(self, idx):
        return self._shells[idx]

    def nbf(self):
        return len(self._shells)


class SGaussInterface:
    """Same constructor and methods as interface_psi4.Psi4Interface (interface_psi4.py:65-406)."""

    grid_radial = 40
    grid_theta = 10

    def __init__(self, mol_xyz, nuclear_numbers, geom_coordinates, basis_set_name,
                 ksdft_functional_name=None):
        self.nuclear_numbers = np.asarray(nuclear_numbers)
        self.geom_coordinates = np.asarray(geom_coordinates, dtype=np.float64)
        self.basis_set_name = basis_set_name
        self.ksdft_functional_name = ksdft_functional_name
        basis = BASES[str(basis_set_name).lower()]
        self.centers_bohr = self.geom_coordinates * ANG_TO_BOHR
        alphas, coefs, prim_center, prim_ao, ao_center = [], [], [], [], []
        for a, z in enumerate(self.nuclear_numbers):
            for shell in basis[int(z)]:
                ao = len(ao_center)
                ao_center.append(a)
                # normalise the contraction so that <chi|chi> = 1
                al = np.array([p[0] for p in shell])
                d = np.array([p[1] for p in shell]) * (2.0 * al / np.pi) ** 0.75
                pp = al[:, None] + al[None, :]
                s = np.sum(d[:, None] * d[None, :] * (np.pi / pp) ** 1.5)
                d = d / np.sqrt(s)
                for x, c in zip(al, d):
                    alphas.append(x)
                    coefs.append(c)
                    prim_center.append(a)
                    prim_ao.append(ao)
        self.alpha = np.array(alphas)
        self.coef = np.array(coefs)
        self.prim_center = np.array(prim_center)
        self.prim_ao = np.array(prim_ao)
        self.ao_center = np.array(ao_center)
        self.num_ao = len(ao_center)
        self.psi4_basis_set = _FakeBasisSet(ao_center)
        # contraction matrix [nprim, nao]
        self._cm = np.zeros((len(alphas), self.num_ao))
        self._cm[np.arange(len(alphas)), self.prim_ao] = self.coef
        self._pair_cache = None

    # -- primitive pair quantities
    def _pairs(self):
        if self._pair_cache is None:
            a = self.alpha
            r = self.centers_bohr[self.prim_center]
            p = a[:, None] + a[None, :]
            mu = a[:, None] * a[None, :] / p
            ab2 = np.sum((r[:, None, :] - r[None, :, :]) ** 2, axis=-1)
            kab = np.exp(-mu * ab2)
            pc = (a[:, None, None] * r[:, None, :] + a[None, :, None] * r[None, :, :]) / p[:, :, None]
            self._pair_cache = (p, mu, ab2, kab, pc)
        return self._pair_cache

    def _contract2(self, m):
        return self._cm.T @ m @ self._cm

    def ao_overlap_integral(self):
        p, mu, ab2, kab, _ = self._pairs()
        return self._contract2((np.pi / p) ** 1.5 * kab)

    def ao_kinetic_integral(self):
        p, mu, ab2, kab, _ = self._pairs()
        return self._contract2(mu * (3.0 - 2.0 * mu * ab2) * (np.pi / p) ** 1.5 * kab)

    def ao_nuclear_attraction_integral(self):
        p, mu, ab2, kab, pc = self._pairs()
        v = np.zeros_like(p)
        for z, c in zip(self.nuclear_numbers, self.centers_bohr):
            pc2 = np.sum((pc - c) ** 2, axis=-1)
            v += -float(z) * (2.0 * np.pi / p) * kab * boys_f0(p * pc2)
        return self._contract2(v)

    # Big systems (N ~ 64-96, 32-48 atoms) take tens of seconds per tensor / grid in numpy; the same geometry is asked
    # for several times in one test session (host-path and device-path SCF tests), so results are memoised per process.
    _memo = {}

    def _key(self, what):
        return (what, str(self.basis_set_name).lower(), self.nuclear_numbers.tobytes(), self.geom_coordinates.tobytes())

    def ao_electron_repulsion_integral(self):
        k = self._key("eri")
        if k not in SGaussInterface._memo:
            if len(SGaussInterface._memo) > 6:
                SGaussInterface._memo.clear()
            SGaussInterface._memo[k] = self._eri_uncached()
        return SGaussInterface._memo[k].copy()

    def _eri_uncached(self):
        p, mu, ab2, kab, pc = self._pairs()
        npr = len(self.alpha)
        pf = p.reshape(-1)
        kf = kab.reshape(-1)
        pcf = pc.reshape(-1, 3)
        cm2 = (self._cm[:, None, :, None] * self._cm[None, :, None, :]).reshape(npr * npr, -1)
        n = self.num_ao
        out = np.zeros((n * n, n * n))
        chunk = max(1, 2_000_000 // (npr * npr))
        for s in range(0, npr * npr, chunk):
            e = min(npr * npr, s + chunk)
            pq2 = np.sum((pcf[s:e, None, :] - pcf[None, :, :]) ** 2, axis=-1)
            pp, qq = pf[s:e, None], pf[None, :]
            g = 2.0 * np.pi ** 2.5 / (pp * qq * np.sqrt(pp + qq)) * kf[s:e, None] * kf[None, :] \
                * boys_f0(pp * qq / (pp + qq) * pq2)
            out += cm2[s:e].T @ (g @ cm2)
        eri = out.reshape(n, n, n, n)
        # enforce exact 8-fold symmetry (the closed form is symmetric up to rounding)
        eri = 0.5 * (eri + eri.transpose(1, 0, 2, 3))
        eri = 0.5 * (eri + eri.transpose(0, 1, 3, 2))
        eri = 0.5 * (eri + eri.transpose(2, 3, 0, 1))
        return np.ascontiguousarray(eri, dtype="float64")

    def iter_eri_shell_quartets(self, shell_sizes=None):
        """Canonical shell quartets (M >= N, P >= Q, MN >= PQ) as dense blocks (i0, j0, k0, l0, values), the surface
        of sqcc_b200.interface_psi4.Psi4Interface.iter_eri_shell_quartets.  `shell_sizes` groups consecutive AOs
        into fake shells (default: ragged 1,3,2,... pattern) so that multi-function blocks are exercised."""
        eri = self.ao_electron_repulsion_integral()
        n = eri.shape[0]
        if shell_sizes is None:
            shell_sizes, pat, t = [], [1, 3, 2, 5], 0
            while sum(shell_sizes) < n:
                shell_sizes.append(min(pat[t % len(pat)], n - sum(shell_sizes)))
                t += 1
        first = np.concatenate([[0], np.cumsum(shell_sizes)])
        ns = len(shell_sizes)
        for m in range(ns):
            for nn in range(m + 1):
                for p in range(m + 1):
                    for q in range((p if p < m else nn) + 1):
                        yield (int(first[m]), int(first[nn]), int(first[p]), int(first[q]),
                               eri[first[m]:first[m + 1], first[nn]:first[nn + 1], first[p]:first[p + 1],
                                   first[q]:first[q + 1]])

    def get_basis_atomic_affiliation(self):
        return self.ao_center.copy()

    # -- KS-DFT surface: Becke-partitioned atom-centred product grids
    def generate_numerical_integration_grids_and_weights(self, mm_coords=None, mm_charges=None):
        k = self._key("grid")
        if k not in SGaussInterface._memo:
            if len(SGaussInterface._memo) > 6:
                SGaussInterface._memo.clear()
            SGaussInterface._memo[k] = self._grids_uncached()
        g, w = SGaussInterface._memo[k]
        return g.copy(), w.copy()

    def _grids_uncached(self):
        nr, nt = self.grid_radial, self.grid_theta
        # Gauss-Chebyshev (2nd kind) radial points mapped with r = (1+x)/(1-x)
        i = np.arange(1, nr + 1)
        x = np.cos(i * np.pi / (nr + 1))
        wr = np.pi / (nr + 1) * np.sin(i * np.pi / (nr + 1)) ** 2 / np.sqrt(1.0 - x * x)
        r = (1.0 + x) / (1.0 - x)
        wr = wr * 2.0 / (1.0 - x) ** 2 * r * r
        ct, wt = np.polynomial.legendre.leggauss(nt)
        nphi = 2 * nt
        ph = (np.arange(nphi) + 0.5) * 2.0 * np.pi / nphi
        st = np.sqrt(1.0 - ct * ct)
        ux = (st[:, None] * np.cos(ph)[None, :]).reshape(-1)
        uy = (st[:, None] * np.sin(ph)[None, :]).reshape(-1)
        uz = np.repeat(ct, nphi)
        wa = np.repeat(wt, nphi) * (2.0 * np.pi / nphi)
        pts, wts = [], []
        centers = self.centers_bohr
        for a, ca in enumerate(centers):
            xyz = ca[None, :] + (r[:, None, None] * np.stack([ux, uy, uz], axis=-1)[None, :, :]).reshape(-1, 3)
            w = (wr[:, None] * wa[None, :]).reshape(-1)
            # Becke fuzzy-cell weights.  Vectorised over b with c as the (ascending) outer loop: for every b the factors
            # are multiplied in the same order, with the same scalar operations, as the plain double loop over (b, c != b)
            # -- bit-identical weights (the committed goldens depend on them), ~10x faster for 30+ atoms.
            nc = len(centers)
            cell = np.ones((nc, len(xyz)))
            dist = np.linalg.norm(xyz[None, :, :] - centers[:, None, :], axis=-1)
            for c in range(nc):
                rbc = np.array([np.linalg.norm(centers[b] - centers[c]) if b != c else 1.0 for b in range(nc)])
                m = (dist - dist[c][None, :]) / rbc[:, None]
                for _ in range(3):
                    m = 1.5 * m - 0.5 * m ** 3
                f = 0.5 * (1.0 - m)
                f[c] = 1.0
                cell *= f
            w = w * cell[a] / np.sum(cell, axis=0)
            pts.append(xyz)
            wts.append(w)
        return np.concatenate(pts), np.concatenate(wts)

    def generate_ao_values_at_grids(self, real_space_grids):
        g = np.asarray(real_space_grids)
        r = self.centers_bohr[self.prim_center]
        d2 = np.sum((g[:, None, :] - r[None, :, :]) ** 2, axis=-1)
        prim = np.exp(-self.alpha[None, :] * d2)
        return np.ascontiguousarray(prim @ self._cm)

    def compute_external_potential_integrals(self, mm_coords, mm_charges, real_space_grids, weights_grids):
        """Grid quadrature of the MM point-charge potential, form of interface_psi4.py:387-403."""
        phi = self.generate_ao_values_at_grids(real_space_grids)
        v = self.mm_potential_at_grids(mm_coords, mm_charges, real_space_grids)
        return phi.T @ (phi * (weights_grids * v)[:, None])

    def mm_potential_at_grids(self, mm_coords, mm_charges, real_space_grids):
        v = np.zeros(len(real_space_grids))
        for xyz, q in zip(np.asarray(mm_coords) * ANG_TO_BOHR, mm_charges):
            r = np.linalg.norm(real_space_grids - xyz, axis=-1)
            v += np.where(r > 1e-10, -q / np.where(r > 1e-10, r, 1.0), 0.0)
        return v


def stub_module(name="interface_psi4"):
    m = types.ModuleType(name)
    m.Psi4Interface = SGaussInterface
    m.get_psi4_version = lambda: "stub-sgauss (psi4 not installed)"
    return m


def h_chain(n_atoms, spacing_ang=0.9, z=1):
    """Linear chain along z; returns (mol_xyz, nuclear_numbers, coords_ang) like input.read_xyz."""
    coords = np.zeros((n_atoms, 3))
    coords[:, 2] = spacing_ang * np.arange(n_atoms)
    zs = np.full(n_atoms, z, dtype=int)
    sym = {1: "H", 2: "He"}[z]
    xyz = f"{n_atoms}\nchain\n" + "\n".join(f"{sym} 0.0 0.0 {c:.10f}" for c in coords[:, 2]) + "\n"
    return xyz, zs, coords


def cluster(zs, coords_ang):
    zs = np.asarray(zs, dtype=int)
    coords = np.asarray(coords_ang, dtype=np.float64)
    sym = {1: "H", 2: "He"}
    xyz = f"{len(zs)}\ncluster\n" + "\n".join(
        f"{sym[int(z)]} {c[0]:.10f} {c[1]:.10f} {c[2]:.10f}" for z, c in zip(zs, coords)) + "\n"
    return xyz, zs, coords
