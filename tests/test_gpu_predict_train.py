"""GPU: the caller side of the solve step on the device against goldens of the
unmodified reference (tests/golden/predict_*.npz): ``Desc.from_R``, ``GDMLPredict.predict`` for geometries outside the
training set, ``create_model``'s ``R_d_desc_alpha``, the integration constant of ``_recov_int_const``, the model file
format, the checkpoint/resume protocol and the ``create_data.cg_steps`` result dictionary."""
import os
import pickle

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10
CASES = ['predict_eth_s6_m24', 'predict_asp_s1_m10']


@pytest.fixture(scope='module')
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    return torch


def _model_from_golden(g):
    return {'type': 'm', 'z': g['z'], 'R_desc': g['R_desc_T'], 'R_d_desc_alpha': g['R_d_desc_alpha'], 'sig': int(g['sig']),
            'std': float(g['std']), 'c': float(g['c']), 'perms': g['perms'], 'tril_perms_lin': g['tril_perms_lin'],
            'alphas_F': g['alphas_F'], 'interact_cut_off': None}


def _task_from_golden(g, tol=None):
    M = int(g['M'])
    return {'type': 't', 'dataset_name': 'synthetic_' + str(g['kind']), 'dataset_theory': 'harmonic_pairs', 'z': g['z'],
            'R_train': g['R_train'], 'F_train': g['F_train'], 'E_train': g['E_train'], 'idxs_train': np.arange(M),
            'md5_train': 'synthetic', 'idxs_valid': np.arange(M, M + 1), 'md5_valid': 'synthetic', 'sig': int(g['sig']),
            'lam': 1e-15, 'use_E': True, 'use_E_cstr': False, 'use_sym': g['perms'].shape[0] > 1, 'use_cprsn': False,
            'solver_name': 'cg', 'solver_tol': float(g['tol']) if tol is None else tol, 'n_inducing_pts_init': 25,
            'interact_cut_off': None, 'perms': g['perms'], 'truncated_cholesky': 1500}


@pytest.mark.parametrize('case', CASES)
def test_descriptors_on_device(torch_cuda, golden, case):
    from mlff_preconditioner_b200.desc import Desc
    from mlff_preconditioner_b200.engine import descriptors_on_device

    g = golden(case)
    M, N = int(g['M']), int(g['N'])
    xd, gd = descriptors_on_device(g['R_train'])
    assert relerr(xd.cpu().numpy(), g['R_desc_T'].T) < 1e-15        # the reference's own descriptors
    xh, gh = Desc(N).from_R(g['R_train'].reshape(M, -1))
    assert np.abs(xd.cpu().numpy() - xh).max() <= 4e-16 * np.abs(xh).max()
    assert np.abs(gd.cpu().numpy() - gh).max() <= 1e-14 * np.abs(gh).max()   # r^3 by two products vs numpy's power


@pytest.mark.parametrize('case', CASES)
def test_predict_matches_reference(torch_cuda, golden, case):
    from mlff_preconditioner_b200.predict import GDMLPredict

    g = golden(case)
    gp = GDMLPredict(_model_from_golden(g))
    E, F = gp.predict(g['R_query'])
    assert E.shape == g['E_query'].shape and F.shape == g['F_query'].shape
    assert relerr(E, g['E_query']) < TOL and relerr(F, g['F_query']) < TOL
    # one geometry, 1-D input (predict.py:1041-1042)
    E1, F1 = gp.predict(g['R_query'][3])
    assert relerr(E1, g['E_query'][3:4]) < TOL and relerr(F1, g['F_query'][3:4]) < TOL
    # training mode: descriptors supplied (predict.py:1038)
    from mlff_preconditioner_b200.engine import descriptors_on_device
    xd, gd = descriptors_on_device(g['R_train'])
    M = int(g['M'])
    E_tr, F_tr = gp.predict(g['R_train'].reshape(M, -1), R_desc=xd, R_d_desc=gd)
    assert relerr(E_tr, g['E_train_pred']) < TOL and relerr(F_tr, g['F_train_pred']) < TOL
    # batching of the queries does not change anything
    Eb, Fb = gp.engine.predict(*gp.engine.desc_from_R(g['R_query'].reshape(-1, int(g['N']), 3)), beta=gp._beta, max_batch=5)
    assert relerr(Eb.cpu().numpy() * gp.std + gp.c, g['E_query']) < TOL
    # set_alphas re-targets the model (predict.py:400-449): twice the coefficients, twice the forces
    gp.set_alphas(gd.cpu().numpy(), 2.0 * g['alphas_F'])
    E2, F2 = gp.predict(g['R_query'])
    assert relerr(F2, 2.0 * g['F_query']) < TOL


@pytest.mark.parametrize('case', CASES)
def test_train_model_dict_and_integration_constant(torch_cuda, golden, case, tmp_path):
    """Our GDMLTrain.train on the golden's training set: the reference's model keys, R_d_desc_alpha = J alpha, the
    integration constant and -- through the saved file -- predictions equal to the reference's."""
    from mlff_preconditioner_b200 import io as mio
    from mlff_preconditioner_b200.predict import GDMLPredict
    from mlff_preconditioner_b200.train import GDMLTrain
    from oracle import sgdml_oracle as orc

    g = golden(case)
    M = int(g['M'])
    gt = GDMLTrain(use_torch=True)
    np.random.seed(0)
    model = gt.train(_task_from_golden(g), break_percentage=float(g['frac']), str_preconditioner='cholesky')
    ref_keys = set(str(k) for k in g['model_keys'])
    assert ref_keys <= set(model.keys()) | {'eigvals', 'eigvals_K'}, sorted(ref_keys - set(model.keys()))
    assert model['is_conv'] and model['use_E']
    assert relerr(model['alphas_F'], g['alphas_F']) < 1e-3                       # both solved to tol 1e-6
    assert abs(model['solver_iters'] - int(g['solver_iters'])) <= max(1, int(0.05 * int(g['solver_iters'])))
    _, gd = orc.desc_from_R(g['R_train'])
    assert relerr(model['R_d_desc_alpha'], orc.d_desc_dot_vec(gd, model['alphas_F'].reshape(M, -1))) < 1e-12
    assert abs(model['c'] - float(g['c'])) <= 1e-4 * abs(float(g['c']))
    assert relerr(model['R_desc'], g['R_desc_T']) < 1e-15 and model['std'] == pytest.approx(float(g['std']), rel=1e-14)
    # file format: flat .npz like the reference CLI writes; loads back into a predictor
    path = mio.save_model(str(tmp_path / 'model.npz'), model)
    loaded = mio.load_model(path)
    assert mio.is_valid_model(loaded)
    E, F = GDMLPredict(loaded).predict(g['R_query'])
    assert relerr(F, g['F_query']) < 1e-4 and relerr(E, g['E_query']) < 1e-4
    # labels with flipped sign: energies are switched off like train.py:957-959
    bad = _task_from_golden(g)
    bad['E_train'] = -bad['E_train'] + 2 * float(np.mean(g['E_train']))
    m2 = gt.train(bad, break_percentage=float(g['frac']), str_preconditioner='cholesky')
    assert m2['use_E'] is False and m2['c'] == 0.0


def test_checkpoint_segments_and_resume(torch_cuda, golden):
    """save_progr_callback receives unconverged models while ONE uninterrupted CG recurrence runs (identical iteration
    count and coefficients to the run without callback); an unconverged model seeds a later run through alphas0_F."""
    from mlff_preconditioner_b200 import io as mio
    from mlff_preconditioner_b200.train import GDMLTrain

    g = golden('predict_eth_s6_m24')
    gt = GDMLTrain(use_torch=True)
    task = _task_from_golden(g)
    plain = gt.train(dict(task), break_percentage=float(g['frac']), str_preconditioner='cholesky')
    seen = []
    t2 = dict(task)
    t2['_checkpoint_seconds'] = 1e-4       # a segment boundary every few iterations
    t2['_checkpoint_first_iters'] = 7
    seg = gt.train(t2, save_progr_callback=seen.append, break_percentage=float(g['frac']), str_preconditioner='cholesky')
    assert len(seen) >= 2
    assert seg['solver_iters'] == plain['solver_iters']
    assert relerr(seg['alphas_F'], plain['alphas_F']) < 1e-13
    its = [m['solver_iters'] for m in seen]
    assert its == sorted(its) and its[-1] <= seg['solver_iters']
    for m in seen:
        assert m['type'] == 'm' and m['alphas_F'].shape == plain['alphas_F'].shape and np.isfinite(m['c'])
    # resume: an early unconverged model continues to the same solution, iteration counts add up
    early = seen[0]
    t3 = mio.resume_task(task, early)
    res = gt.train(t3, break_percentage=float(g['frac']), str_preconditioner='cholesky')
    assert res['is_conv'] and relerr(res['alphas_F'], plain['alphas_F']) < 1e-3
    assert res['solver_iters'] > early['solver_iters']
    # the reference's own create_task_from_model carries inducing_pts_idxs, which Iterative.solve rejects (:680)
    ds = {'R': g['R_train'], 'F': g['F_train'], 'E': g['E_train']}
    t4 = gt.create_task_from_model(plain, ds)
    assert 'alphas0_F' in t4 and 'inducing_pts_idxs' in t4 and t4['solver_iters'] == plain['solver_iters']
    with pytest.raises(AssertionError):
        gt.train(t4, break_percentage=float(g['frac']), str_preconditioner='cholesky')


def test_cg_steps_result_dictionary(torch_cuda, tmp_path):
    """The create_data.cg_steps mirror writes the pickle the reference's analysis code reads (create_data.py:117-169)."""
    from mlff_preconditioner_b200 import synthetic
    from mlff_preconditioner_b200.tools import create_data
    from mlff_preconditioner_b200.tools.rule_of_thumb import default_break_percentage

    ds = synthetic.make_dataset('ethanol', 40, seed=1)
    task, gt = create_data.create_task(30, dataset=ds)
    assert task['solver_name'] == 'cg' and task['sig'] == 10 and task['lam'] == 1e-15
    n = 3 * 9 * 30
    frac = default_break_percentage(task['dataset_name'], n)
    res, model = create_data.cg_steps(task, gt, 30, frac, 'cholesky', path_to_script=str(tmp_path))
    want = {'t_cholesky', 'time_cg_step', 'chol_t_begin', 'chol_t_end', 'chol_t_correction', 'cholesky_percentage',
            'cholesky_cgsteps', 'K.shape', 'n_kernel', 'k', 'total_time_preconditioner', 'total_time_solve',
            'total_time_cg', 'task', 'dataset_name', 'sig', 'lam', 'solver_tol', 'platform', 'n_datapoints'}
    assert want <= set(res.keys())
    assert res['K.shape'] == (n, n) and res['n_kernel'] == n and res['k'] == int(frac * n)
    assert isinstance(res['cholesky_cgsteps'], int) and res['t_cholesky'].shape == (res['k'],)
    assert os.path.exists(res['_file'])
    with open(res['_file'], 'rb') as f:
        back = pickle.load(f)
    assert back['cholesky_cgsteps'] == res['cholesky_cgsteps'] and back['task']['str_preconditioner'] == 'cholesky'
    res2, _ = create_data.cg_steps(task, gt, 30, 0.1, 'random_scores', path_to_script=str(tmp_path))
    assert 'random_scores_cgsteps' in res2 and 't_cholesky' not in res2
    assert create_data.normalize_to_aspirin(1000, 'ethanol') == 2333 and create_data.calculate_kernel_size(10, 'aspirin') == 630
