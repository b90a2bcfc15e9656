"""Mirror of the reference's experiment glue ``src/tools/create_data.py`` around the device solver: same function
names, arguments and -- most importantly -- the same result dictionary / pickle layout that the reference's plotting
and analysis code reads (``create_data.py:100-170``), so result files of both implementations are interchangeable.

Datasets: the reference loads ``<parent of script>/data/<name>_dft.npz`` (``:23-39``); those files are not shipped, so
``create_task`` also accepts a dataset dict (e.g. ``synthetic.make_dataset``) directly.
"""
import os
import pickle
import platform
from datetime import datetime
from pathlib import Path

import numpy as np

from .. import synthetic
from ..io import save_model  # noqa: F401  (re-exported for drivers)
from ..train import GDMLTrain

lam = 1E-9
info_keys = ['dataset_name', 'sig', 'lam', 'solver_tol']


def get_dataset(path_to_script, name_dataset):
    """``np.load`` of the reference's dataset files (create_data.py:23-39); ``synthetic:<kind>:<T>`` generates one."""
    if name_dataset.startswith('synthetic:'):
        _, kind, T = name_dataset.split(':')
        return synthetic.make_dataset(kind, int(T), seed=0)
    available = ['aspirin', 'ethanol', 'paracetamol', 'benzene', 'uracil', 'azobenzene', 'toluene']
    if path_to_script.startswith('.'):
        path_to_script = os.path.abspath(path_to_script)
    folder = Path(path_to_script).parent / 'data'
    assert folder.exists(), 'Data folder does not exists.'
    if name_dataset in available:
        file_name = f'{name_dataset}_dft.npz'
    elif name_dataset == 'nanotube':
        file_name = f'larger_aims_{name_dataset}.npz'
    elif name_dataset == 'catcher':
        file_name = f'aims_{name_dataset}.npz'
    else:
        assert False, f'incorrect input dataset: {name_dataset}'
    return np.load(str(folder / file_name))


def get_number_of_atoms(dataset_name):
    table = {'aspirin': 21, 'ethanol': 9, 'uracil': 12, 'benzene': 12, 'toluene': 15, 'azobenzene': 24,
             'azobenzene_new': 24, 'catcher': 88, 'aims_catcher': 88, 'nanotube': 370, 'larger_aims_nanotube': 370}
    if dataset_name not in table:
        raise NotImplementedError(f'dataset_name = {dataset_name} is not specified. ')
    return table[dataset_name]


def normalize_to_aspirin(n_datapoints, dataset_name):
    """Rescale n_datapoints per basis of aspirin (create_data.py:75-79)."""
    return max(int(n_datapoints * 21 / get_number_of_atoms(dataset_name)), 2)


def calculate_kernel_size(n_datapoints, dataset_name):
    return 3 * get_number_of_atoms(dataset_name) * n_datapoints


def create_task(n_datapoints, path_to_script='', name_dataset='aspirin', dataset=None, perms=None):
    """(task, gdml_train) like create_data.py:88-97: sig = 10, lam = 1e-15, solver = 'cg'.  The reference also draws
    1000 validation points and searches permutations here (host prep, out of scope): the first ``n_datapoints``
    geometries are the training set and ``perms`` defaults to the identity."""
    if dataset is None:
        dataset = get_dataset(path_to_script=path_to_script, name_dataset=name_dataset)
    gdml_train = GDMLTrain(use_torch=True)
    task = gdml_train.create_task(dataset, int(n_datapoints), valid_dataset=dataset, n_valid=0, sig=10, lam=1e-15,
                                  solver='cg', perms=perms)
    return task, gdml_train


def cg_steps(task, gdml_train, n_datapoints, preconditioner_strength, preconditioner, flag_eigvals=False,
             path_to_script='', write=True):
    """One training run and its result pickle (create_data.py:100-170).  Returns the result dict (the reference
    returns None; the file it writes is the same)."""
    name_dataset = str(task['dataset_name'])
    task['truncated_cholesky'] = 1500
    task['str_preconditioner'] = preconditioner

    def callback(*args, **kwargs):
        pass

    dic_cg_steps = {}
    model = gdml_train.train(task=task, break_percentage=preconditioner_strength, callback=callback,
                             str_preconditioner=preconditioner, flag_eigvals=flag_eigvals)
    actual_preconditioner_size = len(model['inducing_pts_idxs']) / len(model['alphas_F'])
    n = len(model['alphas_F'])
    k = int(actual_preconditioner_size * n)
    if preconditioner == 'cholesky':
        t = model['time_cholesky']
        t_begin = np.median(t[:20])
        t_end = np.median(t[20:])
        dic_cg_steps['t_cholesky'] = t
        dic_cg_steps['time_cg_step'] = model['total_time_cg'] / model['solver_iters']
        dic_cg_steps['chol_t_begin'] = t_begin
        dic_cg_steps['chol_t_end'] = t_end
        dic_cg_steps['chol_t_correction'] = t_end / t_begin - 1
    if model['is_conv'] is False and flag_eigvals is False:
        raise RuntimeError('Solver is not converged.')
    dic_cg_steps[f'{preconditioner}_percentage'] = actual_preconditioner_size
    dic_cg_steps[f'{preconditioner}_cgsteps'] = model['solver_iters']
    dic_cg_steps['K.shape'] = (len(model['alphas_F']), len(model['alphas_F']))
    dic_cg_steps['n_kernel'] = n
    dic_cg_steps['k'] = k
    dic_cg_steps['total_time_preconditioner'] = model['total_time_preconditioner']
    dic_cg_steps['total_time_solve'] = model['total_time_solve']
    dic_cg_steps['total_time_cg'] = model['total_time_cg']
    dic_cg_steps['task'] = task
    for label in info_keys:
        dic_cg_steps[label] = task[label]
    dic_cg_steps['platform'] = platform.uname()
    dic_cg_steps['n_datapoints'] = n_datapoints
    if write:
        now = datetime.now()
        folder = Path(os.path.abspath(path_to_script)) / 'data_new' / name_dataset / preconditioner / f'n = {n_datapoints}'
        file_name = f"{now.date()}_{now.strftime('%H%M')}_k = {k}"
        if flag_eigvals is True:
            file_name += '_eigvals'
        folder.mkdir(exist_ok=True, parents=True)
        with open(folder / (file_name + '.pickle'), 'wb') as file:
            pickle.dump(dic_cg_steps, file)
        dic_cg_steps['_file'] = str(folder / (file_name + '.pickle'))
    return dic_cg_steps, model
