"""Deterministic synthetic molecular datasets.

The reference's datasets are not available offline, so every
measurement and every golden vector in this repo uses geometries generated here.

A dataset is a dict with the keys the reference's ``GDMLTrain.create_task`` reads
(``/root/reference/src/sGDML/sgdml/train.py:370-438``): ``name, theory, z, R[T,N,3],
E[T], F[T,N,3]``.  Geometry = fixed base structure + iid Gaussian displacements;
labels come from a harmonic pair potential ``E = 1/2 sum_{a<b} (r_ab - r0_ab)^2`` with
``F = -grad E`` so energies and forces are consistent.
"""
import numpy as np

_ETHANOL_BASE = np.array(
    [
        # C, C, O, H x 6 (Angstrom) -- an idealised staggered ethanol
        [0.000, 0.000, 0.000],
        [1.520, 0.000, 0.000],
        [2.050, 1.320, 0.000],
        [-0.390, 1.020, 0.000],
        [-0.390, -0.510, 0.885],
        [-0.390, -0.510, -0.885],
        [1.900, -0.520, 0.885],
        [1.900, -0.520, -0.885],
        [3.010, 1.270, 0.000],
    ]
)
_ETHANOL_Z = np.array([6, 6, 8, 1, 1, 1, 1, 1, 1])


def _jittered_grid(n_atoms, spacing, rng):
    """n_atoms sites of a cubic grid (spacing in Angstrom) with a fixed small jitter."""
    side = int(np.ceil(n_atoms ** (1.0 / 3.0)))
    g = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing='ij'), -1).reshape(-1, 3)
    base = g[:n_atoms].astype(float) * spacing
    base += rng.uniform(-0.15, 0.15, size=base.shape)
    return base


def base_structure(kind):
    """Return (R0[N,3], z[N], sigma_R) for 'ethanol' (9), 'aspirin' (21), 'nanotube' (370)
    or 'grid<N>' (N atoms on a jittered grid)."""
    rng = np.random.default_rng(12345)
    if kind == 'ethanol':
        return _ETHANOL_BASE.copy(), _ETHANOL_Z.copy(), 0.1
    if kind == 'aspirin':
        return _jittered_grid(21, 1.45, rng), np.array([6] * 9 + [8] * 4 + [1] * 8), 0.1
    if kind == 'nanotube':
        return _jittered_grid(370, 1.5, rng), np.full(370, 6), 0.05
    if kind.startswith('grid'):
        n = int(kind[4:])
        return _jittered_grid(n, 1.5, rng), np.full(n, 6), 0.08
    raise ValueError('unknown synthetic structure: %s' % kind)


def make_dataset(kind, n_geometries, seed=0):
    """Synthetic dataset dict (keys as the reference's .npz datasets)."""
    R0, z, sigma_R = base_structure(kind)
    n_atoms = R0.shape[0]
    rng = np.random.default_rng(seed)
    R = R0[None] + sigma_R * rng.standard_normal((n_geometries, n_atoms, 3))

    a, b = np.tril_indices(n_atoms, k=-1)
    d0 = np.linalg.norm(R0[a] - R0[b], axis=-1)
    diff = R[:, a, :] - R[:, b, :]
    dist = np.linalg.norm(diff, axis=-1)
    E = 0.5 * np.sum((dist - d0) ** 2, axis=1)
    g = ((dist - d0) / dist)[..., None] * diff  # dE/dr_a for pair (a,b)
    F = np.zeros_like(R)
    np.add.at(F, (slice(None), a), -g)
    np.add.at(F, (slice(None), b), g)
    return {
        'name': np.array('synthetic_' + kind),
        'theory': np.array('harmonic_pairs'),
        'type': np.array('d'),
        'z': z,
        'R': R,
        'E': E,
        'F': F,
    }


def ethanol_perms():
    """A closed permutation group of the 9-atom ethanol labelling above (identity first):
    cyclic rotations of the methyl H's (3,4,5) times the swap of the methylene H's (6,7) -> S=6."""
    perms = []
    for rot in ([3, 4, 5], [4, 5, 3], [5, 3, 4]):
        for sw in ([6, 7], [7, 6]):
            perms.append([0, 1, 2] + rot + sw + [8])
    return np.array(perms, dtype=np.int64)
