"""GPU: the device-driven PCG loop (csrc/pcg.cu) against the oracle's restatement of the legacy scipy loop
(oracle.pcg, reference call site iterative_solver.py:995-1005) on well-conditioned systems where both must agree
step for step: iteration counts exactly, iterates to 1e-10, residual histories entry by entry -- including the
corner cases of the stopping logic (iteration cap, convergence at the first iteration, x0 probe, x0 restart)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def system(golden):
    import torch

    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    from mlff_preconditioner_b200.engine import Engine
    from oracle import sgdml_oracle as orc

    g = golden('eth_s1_m12')
    eng = Engine(g['R_desc'], g['R_d_desc'], g['tril_perms_lin'], int(g['sig']), perms=g['perms'])
    K = eng.kernel_assemble()
    k = int(g['chol_k'])
    Lt, _, _, _ = eng.pchol_build(k)
    return dict(g=g, eng=eng, K=K, Lt=Lt, torch=torch, orc=orc, k=k)


def _run_both(system, lam, tol, maxiter, precon, x0=None, mode='assembled'):
    g, eng, torch, orc = system['g'], system['eng'], system['torch'], system['orc']
    A = -g['K'] + lam * np.eye(eng.n)
    b = g['y']
    T = None
    psolve = lambda r: r.copy()  # noqa: E731
    if precon:
        T = eng.woodbury_factor_(system['Lt'].clone(), lam)
        T_ref = orc.woodbury_factor(g['L'], lam)
        psolve = lambda r: orc.woodbury_apply(T_ref, lam, r)  # noqa: E731
    hist_ref = []

    def mv(v):
        return A @ v

    x_ref, it_ref, res_ref, info_ref = orc.pcg(mv, b, psolve, tol, maxiter, x0=x0)
    eng.set_option('symmetric_gemv', 1 if mode == 'assembled_sym' else 0)
    Kdev = {'assembled': system['K'], 'assembled_sym': None, 'matrix_free': None}[mode]
    if mode == 'assembled_sym':
        Kdev = eng.symop_assemble()
    x0_t = None if x0 is None else torch.as_tensor(x0, device=eng.device)
    x, it, resid, info, bnrm2, hist = eng.pcg(torch.as_tensor(b, device=eng.device), lam, tol, maxiter, K_local=Kdev, T=T,
                                              precon_sign=1.0, x0=x0_t, want_hist=True)
    eng.set_option('symmetric_gemv', 0)
    return (x.cpu().numpy(), it, resid, info, hist), (x_ref, it_ref, res_ref, info_ref)


@pytest.mark.parametrize('mode', ['assembled', 'assembled_sym', 'matrix_free'])
@pytest.mark.parametrize('precon,lam', [(False, 1.0), (True, 1e-2), (True, 1e-3)])
def test_step_for_step_agreement(system, mode, precon, lam):
    """Well-conditioned cases where rounding cannot move the count: -K + I without a preconditioner (condition number
    ~2; with lam = 1e-2 the six-fold eigenvalue lam of the rigid-body null space of K makes the unpreconditioned count
    depend on the summation order: 27/28 on the device, 30 in numpy), and the preconditioned solves."""
    (x, it, resid, info, hist), (x_ref, it_ref, res_ref, info_ref) = _run_both(system, lam, 1e-8, 5000, precon, mode=mode)
    assert info == 0 and info_ref == 0
    assert it == it_ref, (it, it_ref)
    assert relerr(x, x_ref) < 1e-9
    assert abs(resid - res_ref) <= 1e-3 * res_ref + 1e-14   # the last residual is the most rounding-sensitive number
    assert len(hist) == it + 1 and np.all(np.isfinite(hist)) and hist[-1] == resid


def test_iteration_cap_and_batched_state(system):
    """Stop by maxiter at counts that fall on and off the loop's batch boundaries: x after m iterations equals the
    oracle's x after m iterations, info != 0."""
    for m in (1, 2, 3, 4, 5, 7, 8, 9, 13):
        (x, it, resid, info, hist), (x_ref, it_ref, res_ref, info_ref) = _run_both(system, 1e-2, 1e-30, m, True)
        assert it == m and it_ref == m and info != 0 and info_ref != 0
        assert relerr(x, x_ref) < 1e-9, m
        assert abs(resid - res_ref) <= 1e-6 * res_ref, m
        assert len(hist) == m + 1


def test_convergence_at_first_iteration_and_probe(system):
    g, eng, torch = system['g'], system['eng'], system['torch']
    # so loose that the first iterate passes (legacy: no true-residual recheck at it == 1)
    (x, it, resid, info, hist), (x_ref, it_ref, res_ref, info_ref) = _run_both(system, 1e-2, 0.999, 100, True)
    assert (it, info) == (it_ref, info_ref) and it >= 1
    assert relerr(x, x_ref) < 1e-9
    # x0 = the solution: the legacy probe ||A x0 - b|| <= tol returns before the first iteration
    lam = 1e-2
    A = -g['K'] + lam * np.eye(eng.n)
    sol = np.linalg.solve(A, g['y'])
    (x, it, resid, info, hist), (x_ref, it_ref, res_ref, info_ref) = _run_both(system, lam, 1e-6, 100, True, x0=sol)
    assert it == 0 and it_ref == 0 and info == 0 and relerr(x, sol) < 1e-12
    # restart from a perturbed solution
    x0 = sol + 1e-3 * np.random.default_rng(0).standard_normal(eng.n)
    (x, it, resid, info, hist), (x_ref, it_ref, res_ref, info_ref) = _run_both(system, lam, 1e-10, 1000, True, x0=x0)
    assert it == it_ref and info == 0 and relerr(x, x_ref) < 1e-9


def test_nan_is_reported_not_looped(system):
    eng, torch = system['eng'], system['torch']
    b = torch.full((eng.n,), float('nan'), dtype=torch.float64, device=eng.device)
    with pytest.raises(Exception):
        eng.pcg(b, 1e-2, 1e-8, 50, K_local=system['K'])
