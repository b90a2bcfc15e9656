"""Host-side (numpy) descriptor preparation -- one-off work that stays in Python.

Mirrors the parts of the reference's ``Desc`` that feed the solve step
(``/root/reference/src/sGDML/sgdml/utils/desc.py:237-462``): the inverse-distance descriptor
``x_d = 1/|r_a - r_b|`` over pairs ``d <-> (a_d, b_d)`` enumerated like ``np.tril_indices(N, -1)``
(a_d > b_d), its compressed Jacobian ``g_d = (r_a - r_b)/|r_a - r_b|^3`` and the conversion
of atom permutations to descriptor permutations.  The per-matvec pieces (J.v, J^T.f) run on
the GPU (csrc/), not here.
"""
import numpy as np


class Desc(object):
    """Same constructor/attributes as the reference's ``Desc`` (desc.py:238-290), minus the
    periodic-boundary and cut-off options, which the solve path never uses."""

    def __init__(self, n_atoms, interact_cut_off=None, max_processes=None):
        if interact_cut_off is not None:
            raise NotImplementedError('interact_cut_off is outside the hot-path contract')
        self.n_atoms = n_atoms
        self.dim_i = 3 * n_atoms
        self.dim = (n_atoms * (n_atoms - 1)) // 2
        self.tril_indices = np.tril_indices(n_atoms, k=-1)
        self.max_processes = max_processes

    def from_R(self, R, lat_and_inv=None, callback=None):
        """(R_desc[M,D], R_d_desc[M,D,3]) for geometries R[M,3N]  (desc.py:292-358, :112-200)."""
        if lat_and_inv is not None:
            raise NotImplementedError('lattices are outside the hot-path contract')
        R = np.asarray(R, dtype=np.float64)
        if R.ndim == 1:
            R = R[None, :]
        R = R.reshape(R.shape[0], -1, 3)
        a, b = self.tril_indices
        pdiff = R[:, a, :] - R[:, b, :]
        pdist = np.sqrt(np.einsum('mdc,mdc->md', pdiff, pdiff))
        R_desc = 1.0 / pdist
        R_d_desc = pdiff / (pdist ** 3)[..., None]
        if callback is not None:
            callback(R.shape[0], R.shape[0])
        return R_desc, R_d_desc

    def perm(self, perm):
        """Descriptor permutation of an atom permutation (desc.py:360-389)."""
        n = len(perm)
        rest = np.zeros((n, n))
        rest[np.tril_indices(n, -1)] = list(range((n ** 2 - n) // 2))
        rest = rest + rest.T
        rest = rest[perm, :]
        rest = rest[:, perm]
        return rest[np.tril_indices(n, -1)].astype(int)

    # host copies of the J.v / J^T.f helpers, for callers that still want numpy
    def d_desc_dot_vec(self, R_d_desc, vecs):
        """J.v (desc.py:394-405)."""
        if R_d_desc.ndim == 2:
            R_d_desc = R_d_desc[None, ...]
        if vecs.ndim == 1:
            vecs = vecs[None, ...]
        i, j = self.tril_indices
        vecs = vecs.reshape(vecs.shape[0], -1, 3)
        return np.einsum('kji,kji->kj', R_d_desc, vecs[:, j, :] - vecs[:, i, :])


def tril_perms_lin_from_perms(perms, desc=None):
    """``tril_perms_lin[d*S + p] = pi_p(d) + p*D``  (train.py:783-790)."""
    perms = np.asarray(perms)
    n_perms, n_atoms = perms.shape
    if desc is None:
        desc = Desc(n_atoms)
    tril_perms = np.array([desc.perm(p) for p in perms])
    perm_offsets = np.arange(n_perms)[:, None] * desc.dim
    return (tril_perms + perm_offsets).flatten('F')


def desc_perms_from_tril_perms_lin(tril_perms_lin, dim_d):
    """Inverse of the linearisation above: ``pi[p, d]`` as an int32 ``[S, D]`` table."""
    tril_perms_lin = np.asarray(tril_perms_lin)
    n_perms = len(tril_perms_lin) // dim_d
    pi = tril_perms_lin.reshape(dim_d, n_perms).T - (np.arange(n_perms) * dim_d)[:, None]
    assert pi.min() >= 0 and pi.max() < dim_d, 'malformed tril_perms_lin'
    return np.ascontiguousarray(pi, dtype=np.int32)


def n_atoms_from_dim_d(dim_d):
    """N from D = N(N-1)/2  (train.py:128)."""
    return int((1 + np.sqrt(8 * dim_d + 1)) / 2)
