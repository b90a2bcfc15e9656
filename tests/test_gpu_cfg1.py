"""GPU parity at the full size of BASELINE.json configs[0] -- the one configuration the unmodified reference runs end
to end in the build container (nanotube-size N = 370, M = 9, n = 9990, k = 1954, tol 1e-6; golden
tests/golden/cfg1_nanotube_m9.npz written by tests/golden/make_golden.py cfg1: 7 minutes of the reference on 8 cores).

This is where north_star's "CG iteration counts within +-1" is well defined: the reference needs 119 CG iterations
(num_iters = 120 in its callback convention); the device must reproduce that with the reference's formula, in every
operator mode, and may not need more with the projected form.  It also covers the N = 370 kernels (D = 68 265: the
per-block assembly fallback, 64-column look-ahead panels of 9990 rows) that the small goldens never reach."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope='module')
def cfg1(golden):
    import torch

    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    from mlff_preconditioner_b200.desc import Desc, tril_perms_lin_from_perms
    from mlff_preconditioner_b200.engine import Engine

    g = golden('cfg1_nanotube_m9')
    M, N = int(g['M']), int(g['N'])
    desc = Desc(N)
    tpl = tril_perms_lin_from_perms(g['perms'], desc)
    R_desc, R_d_desc = desc.from_R(g['R_train'].reshape(M, -1))
    eng = Engine(R_desc, R_d_desc, tpl, int(g['sig']), perms=g['perms'])
    return dict(g=g, eng=eng, R_desc=R_desc, R_d_desc=R_d_desc, tpl=tpl, torch=torch)


def test_kernel_entries_diag_matvec(cfg1):
    g, eng, torch = cfg1['g'], cfg1['eng'], cfg1['torch']
    lam = float(g['lam'])
    assert eng.n == 9990
    assert relerr(eng.kernel_diag().cpu().numpy(), g['diag']) < TOL
    panel = eng.kernel_columns(g['panel_cols']).t().cpu().numpy()
    assert relerr(panel, g['K_panel']) < TOL
    assert np.abs(panel - g['K_panel']).max() <= TOL * np.abs(g['K_panel']).max()
    v = torch.as_tensor(g['v'], device=eng.device)
    assert relerr(eng.matvec_free(v, alpha=1.0, shift=-lam).cpu().numpy(), g['K_op_v']) < TOL
    K = eng.kernel_assemble()
    assert relerr(K[:, torch.as_tensor(g['panel_cols'], device=eng.device)].cpu().numpy(), g['K_panel']) < TOL
    assert relerr(eng.gemv(K, v, alpha=1.0, shift=-lam).cpu().numpy(), g['K_op_v']) < TOL
    Ksym = eng.symop_assemble()
    assert relerr(eng.symop_apply(Ksym, v, alpha=1.0, shift=-lam).cpu().numpy(), g['K_op_v']) < TOL


def test_pivots_factor_and_woodbury_apply(cfg1):
    g, eng, torch = cfg1['g'], cfg1['eng'], cfg1['torch']
    k = int(g['chol_k'])
    Lt, idx, _, _ = eng.pchol_build(k)
    assert np.array_equal(idx.cpu().numpy(), g['index_columns'])          # all 9990 entries, bit-exact
    L = Lt.t()
    rows = torch.as_tensor(g['L_rows'], device=eng.device)
    cols = torch.as_tensor(g['L_cols'], device=eng.device)
    assert relerr(L[rows, :].cpu().numpy(), g['L_sample']) < TOL
    assert relerr(L[:, cols].cpu().numpy(), g['L_colsample']) < TOL
    assert relerr(torch.linalg.norm(L, dim=0).cpu().numpy(), g['L_colnorm']) < TOL
    # plain (no look-ahead panel) build: same pivots, same factor
    eng.set_option('pchol_lookahead', 0)
    Lt2, idx2, _, _ = eng.pchol_build(k)
    eng.set_option('pchol_lookahead', 1)
    assert np.array_equal(idx2.cpu().numpy(), g['index_columns'])
    assert relerr(Lt2.cpu().numpy(), Lt.cpu().numpy()) < 1e-12
    lam = float(g['lam'])
    a = torch.as_tensor(g['a'], device=eng.device)
    T = eng.woodbury_factor_(Lt, lam)
    assert relerr(eng.precon_apply(T, lam, 1.0, a).cpu().numpy(), g['P_chol_a']) < TOL
    Qt, Mk, E = eng.projected_factor_(Lt2, lam)
    assert relerr(eng.precon_apply(Qt, lam, 1.0, a, Mk=Mk, E=E).cpu().numpy(), g['P_chol_a']) < TOL


@pytest.mark.parametrize('mode', ['assembled_sym', 'assembled', 'matrix_free'])
def test_iteration_count_within_one_of_the_reference(cfg1, mode):
    """The reference: 119 CG iterations to 1e-6 (golden num_iters = 120 counts the legacy driver's extra callback)."""
    from mlff_preconditioner_b200.solvers.iterative_solver import Iterative

    g = cfg1['g']
    ref_iters = int(g['num_iters'])
    assert ref_iters == 120
    task = {'R_train': g['R_train'], 'F_train': g['F_train'], 'sig': int(g['sig']), 'lam': float(g['lam']),
            'perms': g['perms'], 'use_E_cstr': False, 'solver_tol': float(g['tol']), 'n_inducing_pts_init': 25,
            'truncated_cholesky': 1500, 'kernel_mode': mode, '_want_hist': True}
    counts = {}
    for form in ('woodbury', 'projected'):
        task['precon_form'] = form
        it = Iterative(None, None)
        alphas, num_iters, resid, rmse, idxs, is_conv, info = it.solve(
            task, cfg1['R_desc'], cfg1['R_d_desc'], cfg1['tpl'], g['y'], float(g['y_std']),
            break_percentage=float(g['frac']), str_preconditioner='cholesky')
        assert is_conv and info['precon_form'] == form
        assert np.array_equal(info['index_columns'], g['index_columns'])
        assert relerr(alphas, g['alphas']) < 1e-4, (form, relerr(alphas, g['alphas']))
        counts[form] = num_iters
        if form == 'woodbury':
            assert abs(num_iters - ref_iters) <= 1, (mode, num_iters, ref_iters)
            # the residual curve itself follows the reference's (same operator, same preconditioner, same recurrences)
            h = it.timings['resid_hist_rel'] * np.linalg.norm(g['y'])
            hr = g['resid_hist']
            m = min(len(hr), len(h) - 1, 100)
            assert np.abs(np.log10(h[1:m + 1] / hr[:m])).max() < 0.15, np.abs(np.log10(h[1:m + 1] / hr[:m])).max()
        it.engine.close()
    assert counts['projected'] <= counts['woodbury'] + 1, counts
