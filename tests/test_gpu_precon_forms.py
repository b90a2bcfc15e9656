"""GPU parity of the three evaluations of the pivoted-Cholesky preconditioner (L L^T + lam I)^{-1} and of the
extended-precision Gram they rest on (csrc/gramdd.cu, csrc/precon.cu):

  'woodbury'   the reference's formula (iterative_cholesky.py:141-148) -- library default
  'projected'  orthonormal basis + k x k inverse + defect-corrected complement projector -- what bench.py times
  'orthonormal' + option precon_reorth   the four-pass cross-check of the projected form

Checked against the numpy oracle restatements (oracle.woodbury_*, orthonormal_*, projected_apply, gram_defect)
on the golden geometries and on a seeded n = 5400 case; tolerances 1e-10 relative (north_star) unless a comment
says why not."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10
OP_CASES = ['eth_s1_m12', 'eth_s6_m6', 'asp_s1_m4', 'grid40_s1_m3']


@pytest.fixture(scope='module')
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    return torch


def _engine(g):
    from mlff_preconditioner_b200.engine import Engine

    return Engine(g['R_desc'], g['R_d_desc'], g['tril_perms_lin'], int(g['sig']), perms=g['perms'])


def _seeded_engine(M=200, kind='ethanol', seed=21):
    from mlff_preconditioner_b200 import synthetic
    from mlff_preconditioner_b200.desc import Desc, tril_perms_lin_from_perms
    from mlff_preconditioner_b200.engine import Engine

    ds = synthetic.make_dataset(kind, M, seed=seed)
    N = ds['R'].shape[1]
    perms = np.arange(N)[None]
    desc = Desc(N)
    tpl = tril_perms_lin_from_perms(perms, desc)
    R_desc, R_d_desc = desc.from_R(ds['R'].reshape(M, -1))
    y = ds['F'].ravel().copy()
    y /= np.std(y)
    return Engine(R_desc, R_d_desc, tpl, 10, perms=perms), dict(R_desc=R_desc, R_d_desc=R_d_desc, tpl=tpl, y=y, ds=ds,
                                                                 perms=perms)


def _longdouble_gram(X):
    Xl = X.astype(np.longdouble)
    return Xl @ Xl.T


@pytest.mark.parametrize('m,n', [(5, 40), (70, 1000), (131, 5003), (333, 2049), (64, 30000)])
def test_gram_extended_accumulation(torch_cuda, m, n):
    """mlffpc_syrk_rows (default gram_mode = 1) against an extended-precision Gram: the (hi, lo) accumulation must beat
    a single fp64 running sum by orders of magnitude; split-K slices (small m) and ragged edges included."""
    torch = torch_cuda
    from conftest import load_golden

    eng = _engine(load_golden('eth_s6_m6'))
    rng = np.random.default_rng(m * 7 + n)
    X = rng.standard_normal((m, n))
    Xt = torch.as_tensor(X, device=eng.device)
    ref = _longdouble_gram(X)
    W = eng.syrk_rows(Xt, shift=0.25).cpu().numpy()
    refs = np.asarray(ref + 0.25 * np.eye(m, dtype=np.longdouble), dtype=float)
    assert np.array_equal(W, W.T)
    # one rounding of the exact sum (the final hi + lo) -- not sqrt(n) of them
    assert np.abs(W - refs).max() <= 4e-16 * np.abs(refs).max()
    eng.set_option('gram_mode', 0)
    W0 = eng.syrk_rows(Xt, shift=0.25).cpu().numpy()
    eng.set_option('gram_mode', 1)
    assert relerr(W0, refs) < 1e-13  # the plain kernel is still a correct Gram


def test_gram_defect_exact_small(torch_cuda):
    """k = 24, n = 300: E = Q Q^T - I in exact rational arithmetic.  The vector-pipe kernel (defect_mode 2: TwoProduct +
    TwoSum) must return the correctly rounded answer up to one ulp; the DMMA kernel (mode 1) is exact up to the rounding
    inside each 16-term k-tile product."""
    torch = torch_cuda
    from fractions import Fraction
    from conftest import load_golden

    eng = _engine(load_golden('eth_s6_m6'))
    k, n = 24, 300
    Q = np.linalg.qr(np.random.default_rng(0).standard_normal((n, k)))[0].T.copy()
    Qf = [[Fraction(float(v)) for v in row] for row in Q]
    E_exact = np.array([[float(sum(a * b for a, b in zip(Qf[i], Qf[j])) - (1 if i == j else 0)) for j in range(k)]
                        for i in range(k)])
    scale = np.abs(E_exact).max()
    assert 1e-17 < scale < 1e-14
    Qt = torch.as_tensor(Q, device=eng.device)
    eng.set_option('defect_mode', 2)
    E2 = eng.gram_defect(Qt).cpu().numpy()
    eng.set_option('defect_mode', 1)
    E1 = eng.gram_defect(Qt).cpu().numpy()
    assert np.abs(E2 - E_exact).max() <= 2.3e-16 * scale + 1e-30, np.abs(E2 - E_exact).max()
    # DMMA k-tile: 16 terms of size ~1/n summed with one fp64 rounding each, then exact: ~eps * 16/n * sqrt(16 * n/16)
    assert np.abs(E1 - E_exact).max() <= 2e-16, np.abs(E1 - E_exact).max()


@pytest.mark.parametrize('k,n', [(200, 20000), (130, 4097), (70, 100000)])
def test_gram_defect_dmma_vs_exact_kernel(torch_cuda, k, n):
    """Larger shapes (split-K slices, ragged edges, cfg2-length rows): the DMMA kernel against the exact kernel, and the
    exact kernel against an extended-precision numpy Gram (whose own error is ~n * 5e-20)."""
    torch = torch_cuda
    from conftest import load_golden

    eng = _engine(load_golden('eth_s6_m6'))
    Q = np.linalg.qr(np.random.default_rng(k + n).standard_normal((n, k)))[0].T.copy()
    Qt = torch.as_tensor(Q, device=eng.device)
    eng.set_option('defect_mode', 2)
    E2 = eng.gram_defect(Qt).cpu().numpy()
    eng.set_option('defect_mode', 1)
    E1 = eng.gram_defect(Qt).cpu().numpy()
    Ql = Q.astype(np.longdouble)
    E_ld = np.asarray(Ql @ Ql.T - np.eye(k, dtype=np.longdouble), dtype=float)
    assert np.array_equal(E1, E1.T) and np.array_equal(E2, E2.T)
    assert np.abs(E2 - E_ld).max() <= 3e-20 * n
    # the k-tile products carry eps * |16-term partial| ~ eps * 16 / n per rounding; the (hi, lo) sum adds nothing
    assert np.abs(E1 - E2).max() <= 1.1e-16 * (16.0 / n) * 4 * np.sqrt(n), np.abs(E1 - E2).max()
    assert np.abs(E1 - E2).max() < 0.2 * np.abs((Q @ Q.T - np.eye(k)) - E2).max()


@pytest.mark.parametrize('case', OP_CASES)
def test_forms_apply_vs_oracle(torch_cuda, golden, case):
    """All three forms are the same operator: on the golden geometries the device applies agree with the oracle's
    restatement fed with the device factors (kernel path) and with each other (factor path)."""
    torch = torch_cuda
    from oracle import sgdml_oracle as orc

    g = golden(case)
    eng = _engine(g)
    lam = float(g['lam'])
    k = int(g['chol_k'])
    a = torch.as_tensor(g['a'], device=eng.device)
    Lt0, _, _, _ = eng.pchol_build(k)
    # reference formula against the reference's own output
    T = eng.woodbury_factor_(Lt0.clone(), lam)
    z_w = eng.precon_apply(T, lam, 1.0, a).cpu().numpy()
    assert relerr(z_w, g['P_chol_a']) < TOL
    # projected form: factors from the device, formula from the oracle
    Qt, Mk, E = eng.projected_factor_(Lt0.clone(), lam)
    Qn, Mn, En = Qt.cpu().numpy(), Mk.cpu().numpy(), E.cpu().numpy()
    assert np.abs(Qn @ Qn.T - np.eye(k)).max() < 1e-13
    assert np.abs(En - orc.gram_defect(Qn)).max() < 1e-16   # k-tile rounding of the DMMA kernel (see the defect tests)
    z_p = eng.precon_apply(Qt, lam, 1.0, a, Mk=Mk, E=E).cpu().numpy()
    assert relerr(z_p, orc.projected_apply(Qn, Mn, En, lam, g['a'])) < 1e-12
    assert relerr(eng.precon_apply(Qt, lam, -1.0, a, Mk=Mk, E=E).cpu().numpy(), -z_p) < 1e-14
    # ... and the same operator as the reference's formula (a random vector lives mostly in the complement, where all
    # forms are well conditioned)
    assert relerr(z_p, g['P_chol_a']) < TOL
    # four-pass cross-check
    eng.set_option('precon_reorth', 1)
    z_r = eng.precon_apply(Qt, lam, 1.0, a, Mk=Mk).cpu().numpy()
    eng.set_option('precon_reorth', 0)
    assert relerr(z_r, orc.orthonormal_apply_reorth(Qn, Mn, lam, g['a'])) < 1e-12
    assert relerr(z_r, z_p) < TOL
    # range(L): the part the Woodbury subtraction cancels; truth = L (L^T L + lam I)^{-1} c in extended precision
    L = Lt0.t().cpu().numpy()
    c = np.random.default_rng(5).standard_normal(k)
    a_r = L @ c
    Ll = L.astype(np.longdouble)
    G = np.asarray(Ll.T @ Ll, dtype=float)
    truth = L @ np.linalg.solve(G + lam * np.eye(k), c)
    ar_t = torch.as_tensor(a_r, device=eng.device)
    err_p = relerr(eng.precon_apply(Qt, lam, 1.0, ar_t, Mk=Mk, E=E).cpu().numpy(), truth)
    err_w = relerr(eng.precon_apply(T, lam, 1.0, ar_t).cpu().numpy(), truth)
    assert err_p < 1e-5 and err_p <= err_w * 1.01 + 1e-12, (err_p, err_w)


def test_forms_on_a_seeded_system(torch_cuda):
    """n = 5400, k = 540: apply of every form against the oracle restatement, and the projected form's defect
    correction against the four-pass projection."""
    torch = torch_cuda
    from oracle import sgdml_oracle as orc

    eng, d = _seeded_engine(M=200)
    lam, k = 1e-10, 540
    rng = np.random.default_rng(2)
    a_np = rng.standard_normal(eng.n)
    a = torch.as_tensor(a_np, device=eng.device)
    Lt0, _, _, _ = eng.pchol_build(k)
    L = Lt0.t().cpu().numpy()
    T_ref = orc.woodbury_factor(L, lam)
    T = eng.woodbury_factor_(Lt0.clone(), lam)
    assert relerr(eng.precon_apply(T, lam, 1.0, a).cpu().numpy(), orc.woodbury_apply(T_ref, lam, a_np)) < TOL
    Qt, Mk, E = eng.projected_factor_(Lt0.clone(), lam)
    Qn, Mn, En = Qt.cpu().numpy(), Mk.cpu().numpy(), E.cpu().numpy()
    assert np.abs(En - orc.gram_defect(Qn)).max() < 1e-16
    z_p = eng.precon_apply(Qt, lam, 1.0, a, Mk=Mk, E=E).cpu().numpy()
    assert relerr(z_p, orc.projected_apply(Qn, Mn, En, lam, a_np)) < 1e-12
    assert relerr(z_p, orc.woodbury_apply(T_ref, lam, a_np)) < TOL
    # four-pass cross-check of the projected form
    eng.set_option('precon_reorth', 1)
    z_re = eng.precon_apply(Qt, lam, 1.0, a, Mk=Mk).cpu().numpy()
    eng.set_option('precon_reorth', 0)
    assert relerr(z_re, z_p) < TOL


def _solve(d, eng_kwargs, form, mode, k_frac, tol, options=None):
    from mlff_preconditioner_b200.solvers.iterative_solver import Iterative

    task = {'R_train': d['ds']['R'], 'F_train': d['ds']['F'], 'sig': 10, 'lam': 1e-10, 'perms': d['perms'],
            'use_E_cstr': False, 'solver_tol': tol, 'n_inducing_pts_init': 25, 'truncated_cholesky': 1500,
            'kernel_mode': mode, 'precon_form': form, '_options': options or {}}
    it = Iterative(None, None)
    out = it.solve(task, d['R_desc'], d['R_d_desc'], d['tpl'], d['y'], 1.0, break_percentage=k_frac,
                   str_preconditioner='cholesky')
    return out, it


@pytest.mark.parametrize('mode', ['assembled', 'assembled_sym', 'matrix_free'])
def test_solve_every_form_every_operator(torch_cuda, mode):
    """n = 5400, tol 1e-6: every form converges to the same coefficients in every operator mode; the projected form
    needs no more iterations than the reference formula and as many as its four-pass cross-check (+-2)."""
    eng, d = _seeded_engine(M=200)
    eng.close()
    res = {}
    for form, opts in (('woodbury', None), ('projected', None), ('orthonormal', {'precon_reorth': 1})):
        (alphas, iters, resid, rmse, idxs, conv, info), it = _solve(d, {}, form, mode, 0.1, 1e-6, opts)
        assert conv and info['precon_form'] == form
        res[form] = (alphas, iters)
        it.engine.set_option('precon_reorth', 0)
    a_w, it_w = res['woodbury']
    a_p, it_p = res['projected']
    a_r, it_r = res['orthonormal']
    assert relerr(a_p, a_w) < 1e-4 and relerr(a_r, a_w) < 1e-4
    assert it_p <= it_w + 1, (it_p, it_w)
    assert abs(it_p - it_r) <= max(2, int(0.02 * it_r)), (it_p, it_r)


def test_default_form_is_the_reference_formula(torch_cuda):
    eng, d = _seeded_engine(M=40)
    eng.close()
    from mlff_preconditioner_b200.solvers.iterative_solver import Iterative

    task = {'R_train': d['ds']['R'], 'F_train': d['ds']['F'], 'sig': 10, 'lam': 1e-10, 'perms': d['perms'],
            'use_E_cstr': False, 'solver_tol': 1e-4, 'n_inducing_pts_init': 25, 'truncated_cholesky': 1500}
    it = Iterative(None, None)
    out = it.solve(task, d['R_desc'], d['R_d_desc'], d['tpl'], d['y'], 1.0, break_percentage=0.1,
                   str_preconditioner='cholesky')
    assert out[6]['precon_form'] == 'woodbury' and it.last_P_op.Mk is None and it.last_P_op.E is None
    with pytest.raises(ValueError):
        task['precon_form'] = 'nonsense'
        Iterative(None, None).solve(task, d['R_desc'], d['R_d_desc'], d['tpl'], d['y'], 1.0, break_percentage=0.1,
                                    str_preconditioner='cholesky')


@pytest.mark.parametrize('k', [5, 33, 540, 1000])
def test_tma_row_strip_gemv_equals_register_gemv(torch_cuda, k):
    """'T r' of the apply on the TMA row-strip kernel (option tma_rows = 1, default) against the register-staged GEMV
    (tma_rows = 0) and numpy: few rows (several CTAs per 32-row strip), rows not a multiple of 32, columns not a
    multiple of 256 (n_local = 5400), and a padded leading dimension."""
    torch = torch_cuda
    eng, d = _seeded_engine(M=200)
    lam = 1e-10
    rng = np.random.default_rng(k)
    Tn = rng.standard_normal((k, eng.n_local)) / np.sqrt(eng.n_local)
    a_np = rng.standard_normal(eng.n_local)
    a = torch.as_tensor(a_np, device=eng.device)
    ref = (a_np - Tn.T @ (Tn @ a_np)) / lam
    for ld in (eng.n_local, eng.n_local + 6):
        buf = torch.zeros((k, ld), dtype=torch.float64, device=eng.device)
        buf[:, :eng.n_local] = torch.as_tensor(Tn, device=eng.device)
        T = buf[:, :eng.n_local]
        eng.set_option('tma_rows', 1)
        z1 = eng.precon_apply(T, lam, 1.0, a).cpu().numpy()
        z1b = eng.precon_apply(T, lam, 1.0, a).cpu().numpy()
        eng.set_option('tma_rows', 0)
        z0 = eng.precon_apply(T, lam, 1.0, a).cpu().numpy()
        eng.set_option('tma_rows', 1)
        assert np.array_equal(z1, z1b)                     # deterministic
        assert relerr(z1, ref) < 1e-11 and relerr(z0, ref) < 1e-11 and relerr(z1, z0) < 1e-12
