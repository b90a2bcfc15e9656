"""Pivoted-Cholesky preconditioner on the device; mirrors the reference's ``IterativeCholesky``
(``/root/reference/src/sGDML/sgdml/solvers/iterative_cholesky.py``)."""
import numpy as np

from . import incomplete_cholesky as ichol
from .operators import KernelOperator, LowRankPreconditioner
from ..engine import Engine


class Iterative(object):
    def __init__(self, gdml_train, desc, task, callback=None, max_processes=None, use_torch=False):
        # same constructor as iterative_cholesky.py:27-51; gdml_train / desc / use_torch are accepted
        # for signature parity and unused: everything runs through the device engine.
        self.gdml_train = gdml_train
        self.task = task
        self.desc = desc
        self.K_op = None
        self.engine = None
        n_train, n_atoms = task['R_train'].shape[:2]
        self.n = 3 * n_atoms * n_train
        self.callback = callback

    def _engine(self, R_desc, R_d_desc, tril_perms_lin, sig):
        if self.engine is None:
            self.engine = Engine(R_desc, R_d_desc, tril_perms_lin, sig, perms=self.task.get('perms'))
        return self.engine

    def _assemble_kernel_mat_diag(self, tril_perms_lin, sig, R_desc, R_d_desc, n, use_E_cstr=False,
                                  cols_m_limit=None):
        """``-diag(K)`` as a CUDA tensor (iterative_cholesky.py:241-373)."""
        if use_E_cstr:
            assert False, 'not implemented yet'  # iterative_cholesky.py:352
        eng = self._engine(R_desc, R_d_desc, tril_perms_lin, sig)
        assert eng.n == n, 'incorrect dimensions'
        return eng.kernel_diag()

    def _init_precon_operator(self, diag_K, K_op, lam_regularization, break_percentage=0.1):
        """``(P_op, info_cholesky)`` (iterative_cholesky.py:115-150): rank ``k = int(break_percentage*n)``
        pivoted Cholesky of ``K_op`` (= -K + lam I) and its Woodbury inverse."""
        assert isinstance(K_op, KernelOperator) and K_op.sign < 0, 'K_op must be the negated device kernel operator'
        eng = K_op.engine
        assert diag_K.shape[0] == eng.n_local, 'incorrect dimensions'
        self.K_op = K_op
        k = int(break_percentage * self.n)
        L, _, info_cholesky = ichol.pivoted_cholesky(get_col=ichol.KernelColumns(eng), diagonal=diag_K, max_rank=k)
        if k == 0:
            return LowRankPreconditioner(eng, None, lam_regularization, 1.0), info_cholesky
        T = eng.woodbury_factor_(L.t(), lam_regularization)  # in place: the factor's storage becomes T
        return LowRankPreconditioner(eng, T, lam_regularization, 1.0), info_cholesky

    def _get_col_K(self, i):
        """One column of ``self.K_op`` (iterative_cholesky.py:152-156) -- generated directly, not by a matvec."""
        col = ichol.KernelColumns(self.K_op.engine)(i)
        col[i] += self.K_op.lam  # the reference's column of -K + lam I
        return col
