"""Device-backed ``GDMLPredict``: energies and forces of arbitrary geometries from a trained model
(``/root/reference/src/sGDML/sgdml/predict.py:238-449, 997-1110``; torch twin ``torchtools.py:41-326``).

Same constructor arguments and the same ``predict`` / ``set_alphas`` contract as the reference class; descriptors of the
query geometries (``Desc.from_R``), the pair sums, the energy contraction and ``J^T f`` all run in libmlffpc.so
(mlffpc_desc_from_r / mlffpc_predict).  The model's ``R_d_desc_alpha`` (beta_j = J_j alpha_j) is the only coefficient
table a prediction needs, so a model loaded from an ``.npz`` file works without the training Jacobians.
"""
import numpy as np
import torch

from .desc import Desc
from .engine import Engine


def _is_none(x):
    """None, or the 0-d object array numpy makes of it when a model is stored as .npz."""
    if x is None:
        return True
    x = np.asarray(x)
    return x.dtype == object and x.size == 1 and x.reshape(-1)[0] is None


def _scalar(x):
    return float(np.asarray(x).reshape(-1)[0])


class GDMLPredict(object):
    def __init__(self, model, batch_size=None, num_workers=1, max_processes=None, use_torch=False):
        if 'type' not in model or not (model['type'] == 'm' or model['type'] == b'm' or str(model['type']) == 'm'):
            raise ValueError('The provided data structure is not a valid model.')  # the reference logs and exits
        if 'alphas_E' in model:
            raise NotImplementedError('models trained with use_E_cstr=True are outside the hot-path contract')
        if 'lattice' in model or not _is_none(model.get('interact_cut_off', None)):
            raise NotImplementedError('lattices / interaction cut-offs are outside the hot-path contract')
        self.n_atoms = np.asarray(model['z']).shape[0]
        self.desc = Desc(self.n_atoms, max_processes=max_processes)
        self.std = _scalar(model['std']) if 'std' in model else 1.0
        self.c = _scalar(model['c'])
        self.sig = _scalar(model['sig'])
        self.use_torch = use_torch
        self.batch_size, self.num_workers = batch_size, num_workers
        R_desc = np.ascontiguousarray(np.asarray(model['R_desc']).T, dtype=np.float64)  # stored [D, M] (train.py:664)
        self.n_train = R_desc.shape[0]
        self.engine = Engine(R_desc, None, np.asarray(model['tril_perms_lin']), self.sig,
                             perms=np.asarray(model['perms']))
        self._beta = torch.as_tensor(np.ascontiguousarray(model['R_d_desc_alpha'], dtype=np.float64),
                                     device=self.engine.device)
        assert tuple(self._beta.shape) == (self.n_train, self.desc.dim)

    # ---- reference API -------------------------------------------------------------------------------
    def set_alphas(self, R_d_desc, alphas, alphas_E=None):
        """Re-target the model to new coefficients (predict.py:400-449): beta = J alphas."""
        if alphas_E is not None:
            raise NotImplementedError('use_E_cstr=True is outside the hot-path contract')
        beta = self.desc.d_desc_dot_vec(np.asarray(R_d_desc), np.asarray(alphas).reshape(-1, 3 * self.n_atoms))
        self._beta = torch.as_tensor(np.ascontiguousarray(beta), device=self.engine.device)

    def set_beta_device(self, beta):
        """Device-resident variant of ``set_alphas`` (beta[M, D] CUDA tensor, e.g. ``Engine.d_desc_dot_vec``)."""
        self._beta = beta.contiguous()

    def prepare_parallel(self, n_bulk=1, n_reps=1, return_is_from_cache=False):
        """CPU tuning in the reference (predict.py:624-893); nothing to tune here.  Returns geometries per second 0."""
        return (0, False) if return_is_from_cache else 0

    def get_GPU_batch(self):
        return self.n_train

    def predict_device(self, R=None, R_desc=None, R_d_desc=None, want_E=True):
        """(E[B], F[B, 3N]) as CUDA tensors, scaled by ``std`` and shifted by ``c`` like predict.py:1106-1108."""
        eng = self.engine
        if R_desc is not None and R_d_desc is not None:  # training mode: descriptors are already known
            xd = torch.as_tensor(R_desc, device=eng.device, dtype=torch.float64).contiguous()
            gd = torch.as_tensor(R_d_desc, device=eng.device, dtype=torch.float64).contiguous()
        else:
            xd, gd = eng.desc_from_R(R)
        E, F = eng.predict(xd, gd, beta=self._beta, want_E=want_E)
        F = F * self.std
        if want_E:
            E = E * self.std + self.c
        return E, F

    def predict(self, R, R_desc=None, R_d_desc=None):
        """Energies [B] and forces [B, 3N] for geometries R[B, 3N] (predict.py:997-1110), host arrays."""
        R = np.asarray(R, dtype=np.float64)
        if R.ndim == 1:
            R = R[None, :]
        E, F = self.predict_device(R.reshape(R.shape[0], -1, 3), R_desc, R_d_desc)
        return E.cpu().numpy(), F.cpu().numpy().reshape(R.shape[0], -1)
