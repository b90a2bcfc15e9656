"""Host-side caller of the solve step: the ``solver == 'cg'`` branch of the reference's
``GDMLTrain.train`` (``/root/reference/src/sGDML/sgdml/train.py:707-970``) and the explicit-kernel entry
point ``_assemble_kernel_mat`` (``train.py:1121-1308``), with the reference's signatures.

Dataset sampling, the permutational-symmetry search and the integration constant are one-off host
work outside the hot path; ``create_task`` here takes the training
points and permutations as given.
"""
import numpy as np

from .desc import Desc, tril_perms_lin_from_perms
from .engine import Engine
from .solvers import iterative_solver


class GDMLTrain(object):
    def __init__(self, max_processes=None, use_torch=False, return_K=None):
        self._max_processes = max_processes
        self._use_torch = use_torch
        self.return_K = False if return_K is None else return_K
        self.last_solver = None

    def create_task(self, train_dataset, n_train, valid_dataset=None, n_valid=0, sig=10, lam=1e-15,
                    use_sym=False, use_E=True, use_E_cstr=False, use_cprsn=False, solver='cg',
                    solver_tol=1e-4, n_inducing_pts_init=25, interact_cut_off=None, callback=None,
                    perms=None, idxs_train=None):
        """Task dict with the keys of train.py:431-453.  ``perms`` (atom permutations, identity first)
        must be supplied when symmetries are wanted; the data-dependent search (utils/perm.py) is not
        part of this package."""
        if use_E_cstr:
            raise NotImplementedError('use_E_cstr=True is outside the hot-path contract')
        if use_sym and perms is None:
            raise NotImplementedError('symmetry search is host prep outside this package: pass perms=')
        n_atoms = train_dataset['R'].shape[1]
        if idxs_train is None:
            idxs_train = np.arange(int(n_train))
        if perms is None:
            perms = np.arange(n_atoms)[None, :]
        return {
            'type': 't',
            'dataset_name': np.asarray(train_dataset['name']).astype(str),
            'dataset_theory': np.asarray(train_dataset['theory']).astype(str),
            'z': train_dataset['z'],
            'R_train': train_dataset['R'][idxs_train, :, :],
            'F_train': train_dataset['F'][idxs_train, :, :],
            'E_train': train_dataset['E'][idxs_train] if use_E else None,
            'idxs_train': idxs_train,
            'sig': sig,
            'lam': lam,
            'use_E': use_E,
            'use_E_cstr': use_E_cstr,
            'use_sym': use_sym,
            'use_cprsn': use_cprsn,
            'solver_name': solver,
            'solver_tol': solver_tol,
            'n_inducing_pts_init': n_inducing_pts_init,
            'interact_cut_off': interact_cut_off,
            'perms': np.asarray(perms),
        }

    def train(self, task, cprsn_callback=None, save_progr_callback=None, callback=None, break_percentage=0.1,
              n_columns=None, str_preconditioner='', flag_eigvals=False):
        """Model dict; mirrors train.py:770-950 for ``solver_name == 'cg'``."""
        task = dict(task)
        solver = task['solver_name']
        if solver != 'cg':
            raise NotImplementedError("only solver='cg' is on the hot path (analytic / cg_cholesky are out of scope)")
        n_train, n_atoms = task['R_train'].shape[:2]
        desc = Desc(n_atoms, interact_cut_off=task['interact_cut_off'], max_processes=self._max_processes)
        tril_perms_lin = tril_perms_lin_from_perms(task['perms'], desc)
        R = task['R_train'].reshape(n_train, -1)
        R_desc, R_d_desc = desc.from_R(R, callback=callback)
        y = task['F_train'].ravel().copy()
        y_std = np.std(y)
        y /= y_std
        if n_columns is not None:
            break_percentage = n_columns / len(y)
        assert 0 <= break_percentage <= 1, 'break_percentage is too large'
        task['lam'] = 1e-10  # train.py:866
        iterative = iterative_solver.Iterative(self, desc, callback=callback, max_processes=self._max_processes,
                                               use_torch=self._use_torch)
        self.last_solver = iterative
        (alphas, num_iters, resid, train_rmse, inducing_pts_idxs, is_conv, info_solver) = iterative.solve(
            task, R_desc, R_d_desc, tril_perms_lin, y, y_std, save_progr_callback=save_progr_callback,
            break_percentage=break_percentage, str_preconditioner=str_preconditioner, flag_eigvals=flag_eigvals)
        model = {
            'type': 'm',
            'dataset_name': task['dataset_name'],
            'solver_name': solver,
            'solver_tol': task['solver_tol'],
            'norm_y_train': np.linalg.norm(y),
            'z': task['z'],
            'idxs_train': task['idxs_train'],
            'R_desc': R_desc.T,
            'std': y_std,
            'sig': task['sig'],
            'lam': task['lam'],
            'alphas_F': alphas,
            'perms': task['perms'],
            'tril_perms_lin': tril_perms_lin,
            'use_E': task['use_E'],
            'solver_resid': resid,
            'solver_iters': num_iters,
            'inducing_pts_idxs': inducing_pts_idxs,
            'c': 0.0,
        }
        model.update(info_solver)
        return model

    def _assemble_kernel_mat(self, R_desc, R_d_desc, tril_perms_lin, sig, desc, use_E_cstr=False,
                             col_idxs=np.s_[:], callback=None):
        """Explicit kernel as a CUDA tensor of shape (n, n_cols) (train.py:1121-1308).  ``col_idxs`` is a
        full slice or a sorted unique index list; the panel is a transposed view of the device's
        ``[n_cols, n]`` layout."""
        if use_E_cstr:
            raise NotImplementedError('use_E_cstr=True is outside the hot-path contract')
        eng = Engine(R_desc, R_d_desc, tril_perms_lin, sig)
        if isinstance(col_idxs, slice):
            if col_idxs == np.s_[:]:
                return eng.kernel_assemble()
            col_idxs = np.arange(eng.n)[col_idxs]
        col_idxs = np.asarray(col_idxs)
        assert len(col_idxs) == len(set(col_idxs.tolist()))  # train.py:1197
        assert np.array_equal(col_idxs, np.sort(col_idxs))  # train.py:1201
        if len(col_idxs) > eng.n:
            raise ValueError('Columns indexed beyond range.')
        return eng.kernel_columns(col_idxs).t()
