"""Synthetic benchmark molecules: (H2O)n clusters and all-trans alkanes.  Deterministic:
the same call always returns the same coordinates, so the CPU baseline and the GPU run see one input."""
import numpy as np

from .molecule import ANGSTROM_TO_BOHR, Molecule


def water_cluster_xyz(n, seed=20240607, lattice=3.10, jitter=0.15, min_dist=1.5):
    """n rigid water molecules on the first n sites of a 4x4x4 simple-cubic lattice (constant `lattice` A),
    r(OH) = 0.9572 A, angle HOH = 104.52 deg, uniformly random orientation, +-`jitter` A uniform displacement;
    a molecule is redrawn until every intermolecular distance is >= `min_dist` A.  Returns (Z, xyz in A)."""
    if n > 64:
        raise ValueError("at most 64 molecules")
    rng = np.random.default_rng(seed)
    roh, ang = 0.9572, np.deg2rad(104.52)
    local = np.array([[0.0, 0.0, 0.0], [roh * np.sin(ang / 2), 0.0, roh * np.cos(ang / 2)], [-roh * np.sin(ang / 2), 0.0, roh * np.cos(ang / 2)]])
    placed = []
    for site in range(n):
        ix, iy, iz = site % 4, (site // 4) % 4, site // 16
        base = np.array([ix, iy, iz], dtype=float) * lattice
        while True:
            q = rng.standard_normal(4)
            q /= np.linalg.norm(q)
            a, b, c, d = q
            R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                          [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                          [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])
            mol = local @ R.T + base + rng.uniform(-jitter, jitter, 3)
            if all(np.min(np.linalg.norm(mol[:, None, :] - other[None, :, :], axis=2)) >= min_dist for other in placed):
                placed.append(mol)
                break
    xyz = np.concatenate(placed, axis=0)
    Z = np.tile(np.array([8, 1, 1], dtype=np.int32), n)
    return Z, xyz


def alkane_xyz(nc):
    """All-trans C_n H_{2n+2}: C-C 1.54 A, angle CCC 112 deg, C-H 1.09 A, tetrahedral H.  Returns (Z, xyz in A)."""
    cc, ch = 1.54, 1.09
    half = np.deg2rad(112.0) / 2
    dx, dy = cc * np.sin(half), cc * np.cos(half)
    C = np.array([[i * dx, (i % 2) * dy, 0.0] for i in range(nc)])
    tet = np.deg2rad(109.4712206) / 2
    H = []
    for i in range(nc):
        up = 1.0 if i % 2 else -1.0  # zig-zag: substituents point away from the chain's bend
        # two H in the plane perpendicular to the chain plane
        for s in (+1.0, -1.0):
            H.append(C[i] + ch * np.array([0.0, up * np.cos(tet), s * np.sin(tet)]))
    # terminal hydrogens along the extended chain direction
    H.append(C[0] + ch * np.array([-np.sin(half), np.cos(half), 0.0]) * np.array([1.0, 1.0, 1.0]))
    last_up = 1.0 if (nc - 1) % 2 else -1.0
    H.append(C[-1] + ch * np.array([np.sin(half), -last_up * np.cos(half), 0.0]))
    xyz = np.concatenate([C, np.array(H)], axis=0)
    Z = np.array([6] * nc + [1] * len(H), dtype=np.int32)
    return Z, xyz


def water_cluster(n, basis="p631"):
    Z, xyz = water_cluster_xyz(n)
    return Molecule(Z, xyz * ANGSTROM_TO_BOHR, basis=basis, name="(H2O)%d" % n)


def alkane(nc, basis="p631"):
    Z, xyz = alkane_xyz(nc)
    return Molecule(Z, xyz * ANGSTROM_TO_BOHR, basis=basis, name="C%dH%d" % (nc, 2 * nc + 2))


def synthetic_density(mol, seed=20240607):
    """Symmetric positive semi-definite P = C C^T, C in R^{nb x nelec/2} i.i.d. N(0,1)/sqrt(nb)."""
    rng = np.random.default_rng(seed)
    nocc = max(1, mol.nelec // 2)
    Cm = rng.standard_normal((mol.nbf, nocc)) / np.sqrt(mol.nbf)
    return Cm @ Cm.T


WORKLOADS = {
    # name: (factory, (radial_points, lebedev_order, lmax))
    "h2o32": (lambda: water_cluster(32), (20, 8, 10)),
    "h2o64": (lambda: water_cluster(64), (20, 8, 10)),
    "h2o8": (lambda: water_cluster(8), (20, 8, 10)),
    "c40h82": (lambda: alkane(40), (422, 10, 11)),
    "c40h82_fine": (lambda: alkane(40), (20, 8, 10)),
    "c10h22": (lambda: alkane(10), (20, 8, 10)),
}
