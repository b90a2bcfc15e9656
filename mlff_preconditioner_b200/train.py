"""Host-side caller of the solve step: the ``solver == 'cg'`` branch of the reference's
``GDMLTrain.train`` (``/root/reference/src/sGDML/sgdml/train.py:707-970``), ``create_model`` (``:597-702``),
``_recov_int_const`` (``:972-1119``) and the explicit-kernel entry point ``_assemble_kernel_mat``
(``:1121-1308``), with the reference's signatures.

Everything of size O(M D) or larger runs on the device: descriptors and Jacobians of the training geometries
(mlffpc_desc_from_r instead of the fork pool of ``Desc.from_R``), the solve, ``R_d_desc_alpha = J alpha``
(mlffpc_d_desc_dot_vec) and the energy pass of the integration constant (mlffpc_predict with the training
geometries as queries -- an O(M^2 S D) sweep the reference does on the CPU after every training).  Dataset sampling
and the permutational-symmetry search stay outside; ``create_task`` takes the
training points and permutations as given.
"""
import logging

import numpy as np
import torch

from . import __version__
from .desc import Desc, tril_perms_lin_from_perms
from .engine import Engine, descriptors_on_device
from .solvers import iterative_solver


class GDMLTrain(object):
    _mlffpc_native = True  # create_model takes engine= (the reference's does not)

    def __init__(self, max_processes=None, use_torch=False, return_K=None):
        self._max_processes = max_processes
        self._use_torch = use_torch
        self.return_K = False if return_K is None else return_K
        self.last_solver = None
        self.log = logging.getLogger(__name__)

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
            'code_version': __version__,
            'dataset_name': np.asarray(train_dataset['name']).astype(str),
            'dataset_theory': np.asarray(train_dataset['theory']).astype(str),
            'z': train_dataset['z'],
            'R_train': train_dataset['R'][idxs_train, :, :],
            'F_train': train_dataset['F'][idxs_train, :, :],
            'E_train': train_dataset['E'][idxs_train] if use_E else None,
            'idxs_train': idxs_train,
            'md5_train': train_dataset.get('md5', 'n/a') if hasattr(train_dataset, 'get') else 'n/a',
            'idxs_valid': np.arange(0),
            'md5_valid': 'n/a',
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

    def create_model(self, task, solver, R_desc, R_d_desc, tril_perms_lin, std, alphas_F, alphas_E=None,
                     solver_resid=None, solver_iters=None, norm_y_train=None, inducing_pts_idxs=None,
                     engine=None):
        """Model dict with the reference's keys (train.py:597-702).  ``R_desc`` / ``R_d_desc`` may be host arrays or
        CUDA tensors; ``R_d_desc_alpha`` is computed on the device when an ``engine`` bound to these geometries is
        passed (the reference's einsum, :640-645, otherwise)."""
        if 'cprsn_keep_atoms_idxs' in task:
            raise NotImplementedError('compressed models (use_cprsn) are outside the hot-path contract')
        if alphas_E is not None:
            raise NotImplementedError('use_E_cstr=True is outside the hot-path contract')
        alphas_F = np.asarray(alphas_F, dtype=np.float64)
        if engine is not None and getattr(engine, '_R_d_desc', None) is not None:
            a_dev = torch.as_tensor(np.ascontiguousarray(alphas_F.ravel()), device=engine.device)
            r_d_desc_alpha = engine.d_desc_dot_vec(a_dev).cpu().numpy()
        else:
            Rdd = R_d_desc.cpu().numpy() if torch.is_tensor(R_d_desc) else np.asarray(R_d_desc)
            n_atoms = task['R_train'].shape[1]
            r_d_desc_alpha = Desc(n_atoms).d_desc_dot_vec(Rdd, alphas_F.reshape(-1, 3 * n_atoms))
        R_desc_host = R_desc.cpu().numpy() if torch.is_tensor(R_desc) else np.asarray(R_desc)
        model = {
            'type': 'm',
            'code_version': __version__,
            'dataset_name': task['dataset_name'],
            'dataset_theory': task.get('dataset_theory', 'unknown'),
            'solver_name': solver,
            'solver_tol': task['solver_tol'],
            'norm_y_train': norm_y_train,
            'n_inducing_pts_init': task.get('n_inducing_pts_init', 25),
            'z': task['z'],
            'idxs_train': task['idxs_train'],
            'md5_train': task.get('md5_train', 'n/a'),
            'idxs_valid': task.get('idxs_valid', np.arange(0)),
            'md5_valid': task.get('md5_valid', 'n/a'),
            'n_test': 0,
            'md5_test': None,
            'f_err': {'mae': np.nan, 'rmse': np.nan},
            'R_desc': R_desc_host.T,
            'R_d_desc_alpha': r_d_desc_alpha,
            'interact_cut_off': task.get('interact_cut_off', None),
            'c': 0.0,
            'std': std,
            'sig': task['sig'],
            'lam': task['lam'],
            'alphas_F': alphas_F,
            'perms': task['perms'],
            'tril_perms_lin': tril_perms_lin,
            'use_E': task['use_E'],
            'use_cprsn': task.get('use_cprsn', False),
        }
        if solver_resid is not None:
            model['solver_resid'] = solver_resid
        if solver_iters is not None:
            model['solver_iters'] = solver_iters
        if inducing_pts_idxs is not None:
            model['inducing_pts_idxs'] = inducing_pts_idxs
        if task['use_E']:
            model['e_err'] = {'mae': np.nan, 'rmse': np.nan}
        for key in ('lattice', 'r_unit', 'e_unit'):
            if key in task:
                model[key] = task[key]
        return model

    def create_task_from_model(self, model, dataset):
        """Task that re-creates a model's training run (train.py:537-594; used by ``sgdml resume``).  Faithful to the
        reference incl. ``inducing_pts_idxs``, which its own ``Iterative.solve`` then rejects (iterative_solver.py:680);
        ``io.resume_task`` is the variant that can actually continue a run."""
        idxs_train = model['idxs_train']
        task = {
            'type': 't',
            'code_version': __version__,
            'dataset_name': model['dataset_name'],
            'dataset_theory': model['dataset_theory'],
            'z': model['z'],
            'R_train': dataset['R'][idxs_train, :, :],
            'F_train': dataset['F'][idxs_train, :, :],
            'idxs_train': idxs_train,
            'md5_train': model['md5_train'],
            'idxs_valid': model['idxs_valid'],
            'md5_valid': model['md5_valid'],
            'sig': model['sig'],
            'lam': model['lam'],
            'use_E': model['use_E'],
            'use_E_cstr': 'alphas_E' in model,
            'use_sym': model['perms'].shape[0] > 1,
            'perms': model['perms'],
            'use_cprsn': model['use_cprsn'],
            'solver_name': model['solver_name'],
            'solver_tol': model['solver_tol'],
            'n_inducing_pts_init': model['n_inducing_pts_init'],
            'interact_cut_off': None,
        }
        if 'e_err' in model:
            task['E_train'] = dataset['E'][idxs_train]
        else:
            task['E_train'] = None
        for key in ('lattice', 'r_unit', 'e_unit'):
            if key in model:
                task[key] = model[key]
        if 'alphas_F' in model:
            task['alphas0_F'] = model['alphas_F']
        if 'solver_iters' in model:
            task['solver_iters'] = model['solver_iters']
        if 'inducing_pts_idxs' in model:
            task['inducing_pts_idxs'] = model['inducing_pts_idxs']
        return task

    def train(self, task, cprsn_callback=None, save_progr_callback=None, callback=None, break_percentage=0.1,
              n_columns=None, str_preconditioner='', flag_eigvals=False):
        """Model dict; mirrors train.py:770-970 for ``solver_name == 'cg'``."""
        task = dict(task)
        solver = task['solver_name']
        if solver != 'cg':
            raise NotImplementedError("only solver='cg' is on the hot path (analytic / cg_cholesky are out of scope)")
        if task.get('use_E_cstr', False):
            raise NotImplementedError('use_E_cstr=True is outside the hot-path contract')
        if 'lattice' in task:
            raise NotImplementedError('lattices are outside the hot-path contract')
        n_train, n_atoms = task['R_train'].shape[:2]
        desc = Desc(n_atoms, interact_cut_off=task['interact_cut_off'], max_processes=self._max_processes)
        tril_perms_lin = tril_perms_lin_from_perms(task['perms'], desc)
        # descriptors and Jacobians on the device (train.py:813-819); they stay there for the solve
        R_desc, R_d_desc = descriptors_on_device(task['R_train'])
        if callback is not None:
            callback(n_train, n_train, disp_str='Generating descriptors and their Jacobians')
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
        if not is_conv:
            self.log.warning('Iterative solver did not converge! The optimization problem underlying this force field '
                             'reconstruction task seems to be highly ill-conditioned. We will continue with this '
                             'unconverged model, but its accuracy will likely be very bad.')
        model = self.create_model(task, solver, R_desc, R_d_desc, tril_perms_lin, y_std, alphas, solver_resid=resid,
                                  solver_iters=num_iters, norm_y_train=np.linalg.norm(y),
                                  inducing_pts_idxs=inducing_pts_idxs, engine=iterative.engine)
        model.update(info_solver)
        if model['use_E']:
            c = self._recov_int_const(model, task, R_desc=R_desc, R_d_desc=R_d_desc, engine=iterative.engine)
            if c is None:
                model['use_E'] = False  # train.py:957-959
            else:
                model['c'] = c
        return model

    def _recov_int_const(self, model, task, R_desc=None, R_d_desc=None, engine=None):
        """Least-squares integration constant and the reference's sanity checks on the labels (train.py:972-1119).
        The energy predictions for the M training geometries (an O(M^2 S D) pass) come from the device."""
        n_train = task['E_train'].shape[0]
        if engine is None or getattr(engine, '_R_d_desc', None) is None:
            from .predict import GDMLPredict

            gdml = GDMLPredict(model, max_processes=self._max_processes)
            E_pred, _ = gdml.predict(task['R_train'].reshape(n_train, -1), R_desc=R_desc, R_d_desc=R_d_desc)
        else:
            a_dev = torch.as_tensor(np.ascontiguousarray(np.asarray(model['alphas_F'], dtype=np.float64).ravel()),
                                    device=engine.device)
            E_raw, _ = engine.predict(engine._R_desc, engine._R_d_desc, alphas=a_dev, want_E=True)
            E_pred = E_raw.cpu().numpy() * float(model['std']) + float(model['c'])
        E_ref = np.squeeze(task['E_train'])
        e_fact = np.linalg.lstsq(np.column_stack((E_pred, np.ones(E_ref.shape))), E_ref, rcond=-1)[0][0]
        corrcoef = np.corrcoef(E_ref, E_pred)[0, 1]
        if np.sign(e_fact) == -1:
            self.log.warning('The provided dataset contains gradients instead of force labels (flipped sign). '
                             'Please correct!')
            return None
        if corrcoef < 0.95:
            self.log.warning('Inconsistent energy labels detected! The predicted energies for the training data are only '
                             'weakly correlated with the reference labels (correlation coefficient {:.2f}).'.format(corrcoef))
            return None
        if np.abs(e_fact - 1) > 1e-1:
            self.log.warning('Different scales in energy vs. force labels detected! The integrated forces differ from '
                             'the energy labels by factor ~{:.2f}.'.format(e_fact))
            return None
        return np.sum(E_ref - E_pred) / E_ref.shape[0]

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
