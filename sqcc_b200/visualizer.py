"""Electron density on arbitrary points: the rho = phi D phi^T loop of the reference's
visualizer (visualizer.py:49-87, one Python iteration per point; 10^6 points for its 1000 x 1000 planes, :202) on the
device -- the same DMMA kernel as rho on the integration grid (csrc/grid_lda.cu, sqcc_grid_rho).

Only the numerical part of the reference's DensityVisualizer is provided: `calculate_density_at_points` and the plane
sampling around it.  Plotting (matplotlib contour / slice figures, visualizer.py:200-618) is not part of the Fock-build
path: feed the returned arrays to the reference's plotting code or any other.
"""
import numpy as np

ANG_TO_BOHR = 1.0 / 0.52917721067   # the reference's conversion constant (hf_ksdft.py:221, interface_psi4.py:371)


class DensityVisualizer:
    """Same constructor and `calculate_density_at_points` as visualizer.DensityVisualizer (visualizer.py:23-87)."""

    def __init__(self, scf_calculator, engine=None, points_per_pass=262144):
        self.scf = scf_calculator
        self.density_matrix = scf_calculator.density_matrix_in_ao_basis
        self.num_ao = scf_calculator.num_ao
        self.nuclear_numbers = scf_calculator.nuclear_numbers
        self.geom_coordinates = scf_calculator.geom_coordinates
        self.psi4_interface = scf_calculator.psi4_interface
        self.engine = engine if engine is not None else getattr(scf_calculator, "engine", None)
        if self.engine is None:
            from .backend import FockEngine
            self.engine = FockEngine()
        self.points_per_pass = int(points_per_pass)

    def calculate_density_at_points(self, points):
        """points [P, 3] in Bohr -> rho [P].  Restricted: D; otherwise the total density D[0] + D[1] (the reference
        decides by spin_multiplicity == 1, visualizer.py:66 -- kept, including its behaviour for unrestricted singlets,
        SURVEY.md A.3)."""
        points = np.ascontiguousarray(points, dtype=np.float64)
        if self.scf.spin_multiplicity == 1:
            density = self.density_matrix
        else:
            density = self.density_matrix[0] + self.density_matrix[1]
        density = np.ascontiguousarray(density, dtype=np.float64)
        out = np.empty(len(points))
        for lo in range(0, len(points), self.points_per_pass):
            hi = min(len(points), lo + self.points_per_pass)
            phi = self.psi4_interface.generate_ao_values_at_grids(points[lo:hi])     # provider side (Psi4 / stub)
            self.engine.attach_grid(phi, np.ones(hi - lo))
            out[lo:hi] = self.engine.grid_rho(density)
        return out

    def plane_points(self, axis="z", offset=0.0, num_points=1000, margin=3.0):
        """The square sampling plane of the reference's contour plots (num_points x num_points, visualizer.py:200-260):
        perpendicular to `axis` at `offset` (Bohr), spanning the molecule plus `margin` Bohr."""
        coords = np.asarray(self.geom_coordinates, dtype=np.float64) * ANG_TO_BOHR
        ax = "xyz".index(axis)
        u, v = [a for a in range(3) if a != ax]
        lo = coords.min(axis=0) - margin
        hi = coords.max(axis=0) + margin
        gu = np.linspace(lo[u], hi[u], num_points)
        gv = np.linspace(lo[v], hi[v], num_points)
        uu, vv = np.meshgrid(gu, gv, indexing="ij")
        pts = np.zeros((num_points * num_points, 3))
        pts[:, u], pts[:, v], pts[:, ax] = uu.ravel(), vv.ravel(), offset
        return pts, (gu, gv)

    def density_on_plane(self, axis="z", offset=0.0, num_points=1000, margin=3.0):
        pts, (gu, gv) = self.plane_points(axis, offset, num_points, margin)
        return self.calculate_density_at_points(pts).reshape(num_points, num_points), gu, gv
