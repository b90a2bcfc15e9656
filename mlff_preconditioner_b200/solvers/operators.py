"""Device-backed operators with the ``scipy.sparse.linalg.LinearOperator`` calling convention the
reference uses as its plugin seam (``iterative_solver.py:322,445``; ``iterative_cholesky.py:150,239``):
``op.shape``, ``op.dtype``, ``op.matvec(v)`` on host numpy vectors; ``op.device_apply(t)`` keeps the
vector on the GPU.  Single-GPU objects; the sharded solve drives the engine directly."""
import numpy as np
import torch


class _DeviceOperator(object):
    dtype = np.dtype('float64')

    def __init__(self, engine):
        self.engine = engine
        self.shape = (engine.n, engine.n)

    def matvec(self, v):
        if self.engine.world != 1:
            raise NotImplementedError('host matvec is single-GPU only')
        t = torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64).ravel(), device=self.engine.device)
        return self.device_apply(t).cpu().numpy()

    __call__ = matvec

    def dot(self, v):
        return self.matvec(v)


class KernelOperator(_DeviceOperator):
    """``v -> sign * (K v - lam v)``.  sign = +1 is the reference's ``K_op`` (iterative_solver.py:416-443),
    sign = -1 the ``-K_op`` handed to CG and to the column oracle (:789, :996).  Unlike the reference
    there is no "priming" first call (:418-421): that quirk only absorbs scipy's dtype probe."""

    def __init__(self, engine, lam, sign=1.0, K_local=None):
        super().__init__(engine)
        self.lam, self.sign, self.K_local = float(lam), float(sign), K_local

    def device_apply(self, v):
        e = self.engine
        if self.K_local is not None and self.K_local.dim() == 1:  # symmetric tile storage
            return e.symop_apply(self.K_local, v, alpha=self.sign, shift=-self.sign * self.lam)
        if self.K_local is not None:
            return e.gemv(self.K_local, v, alpha=self.sign, shift=-self.sign * self.lam, x_off=e.row0)
        return e.matvec_free(v, alpha=self.sign, shift=-self.sign * self.lam)

    def __neg__(self):
        return KernelOperator(self.engine, self.lam, -self.sign, self.K_local)


class LowRankPreconditioner(_DeviceOperator):
    """``a -> sign * (a - T^T (T a)) / lam`` with T[k, n_local] on the device.
    sign = +1: pivoted-Cholesky Woodbury inverse (iterative_cholesky.py:145-148);
    sign = -1: Nystroem operators (iterative_solver.py:315-318, :376-379).
    With ``Mk`` the same inverse is held in an orthonormal basis (``Engine.orthonormal_factor_``):
    ``a -> sign * ((a - T^T T a)/lam + T^T Mk T a)``; with ``E = T T^T - I`` as well
    (``Engine.projected_factor_``) the complement uses the exact projector to first order:
    ``a -> sign * ((a - T^T (w - E w))/lam + T^T Mk w)``, ``w = T a``."""

    def __init__(self, engine, T, lam, sign, Mk=None, E=None):
        super().__init__(engine)
        self.T, self.lam, self.sign, self.Mk, self.E = T, float(lam), float(sign), Mk, E

    def device_apply(self, a):
        return self.engine.precon_apply(self.T, self.lam, self.sign, a, Mk=self.Mk, E=self.E)
