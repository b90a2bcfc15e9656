"""GPU, BASELINE.json's full single-GPU size (cfg2: N = 9, M = 4000, n = 108 000, K = 93.3 GB):
size-independent properties, since the oracle cannot run at this size.

 * the assembled GEMV and the matrix-free operator (independent kernels) agree to 1e-10;
 * K is symmetric (probe: u.(K v) == v.(K u));
 * partial Cholesky: index_columns is a permutation, pivots are unique, the residual diagonal is >= 0 and the
   factor reproduces A exactly on pivot columns (L L^T e_pi = A e_pi);
 * Woodbury: P (L L^T + lam I) v == v;
 * PCG: the returned solution satisfies ||A x - b|| <= tol ||b|| when re-checked with the *other* operator.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M_FULL = 4000
K_RANK = 600
LAM = 1e-10


@pytest.fixture(scope='module')
def full():
    import torch

    assert torch.cuda.is_available()
    free, total = torch.cuda.mem_get_info()
    if free < 110e9:
        pytest.skip('needs ~100 GB of free HBM (B200)')
    from bench import make_inputs
    from mlff_preconditioner_b200.engine import Engine

    inp = make_inputs('cfg2')
    assert inp['n'] == 108000
    eng = Engine(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10, perms=inp['perms'])
    K = eng.kernel_assemble()
    return dict(inp=inp, eng=eng, K=K, torch=torch)


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def test_gemv_vs_matrix_free_and_symmetry(full):
    torch, eng, K = full['torch'], full['eng'], full['K']
    gen = torch.Generator(device=eng.device).manual_seed(1)
    u = torch.randn(eng.n, dtype=torch.float64, device=eng.device, generator=gen)
    v = torch.randn(eng.n, dtype=torch.float64, device=eng.device, generator=gen)
    Kv = eng.gemv(K, v)
    Kv_mf = eng.matvec_free(v)
    assert _rel(Kv, Kv_mf) < 1e-10
    assert _rel(eng.symv(K, v), Kv) < 1e-12      # lower-triangle-only matvec
    Ku = eng.gemv(K, u)
    a, b = float(u @ Kv), float(v @ Ku)
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b))
    # operator form used by CG: A = -K + lam I
    Av = eng.gemv(K, v, alpha=-1.0, shift=LAM, x_off=0)
    assert _rel(Av, -Kv + LAM * v) < 1e-13
    # diagonal of the assembled matrix == the closed-form diagonal kernel
    assert _rel(-torch.diagonal(K), eng.kernel_diag()) < 1e-10
    # a column panel == the corresponding columns of the assembled matrix
    cols = torch.tensor([0, 13, 53999, 54000, 107999], device=eng.device)
    assert _rel(eng.kernel_columns(cols), K[:, cols].t().contiguous()) < 1e-10


def test_partial_cholesky_properties(full):
    torch, eng = full['torch'], full['eng']
    diag0 = eng.kernel_diag()
    Lt, idx, diag_res, step_s = eng.pchol_build(K_RANK, diag=diag0)
    idx_h = idx.cpu().numpy()
    assert np.array_equal(np.sort(idx_h), np.arange(eng.n))          # a permutation of all rows
    piv = idx_h[:K_RANK]
    assert len(set(piv.tolist())) == K_RANK
    assert float(diag_res.min()) > -1e-12 * float(diag0.max())       # Schur complement stays PSD
    assert float(diag_res.sum()) < float(diag0.sum())
    assert step_s.shape == (K_RANK,) and (step_s > 0).all()
    # pivots are chosen in order of decreasing residual diagonal at selection time: L[pi_m, m]^2 non-increasing
    lpp = torch.stack([Lt[m, piv[m]] for m in range(0, K_RANK, 37)])
    assert bool((lpp[:-1] >= lpp[1:] * (1 - 1e-12)).all())
    # exactness on pivot columns: (L L^T)[:, pi] == A[:, pi]
    for m in (0, 1, 17, K_RANK // 2, K_RANK - 1):
        pi = int(piv[m])
        col = eng.kernel_columns(np.array([pi]), scale=-1.0)[0]
        rec = Lt.t() @ Lt[:, pi]
        assert _rel(rec, col) < 1e-9, m
    full['Lt'] = Lt


def test_lookahead_and_plain_pivoted_cholesky_agree(full):
    """The blocked (candidate-panel) build and the plain left-looking build choose the same pivots and give the
    same factor up to summation order."""
    torch, eng = full['torch'], full['eng']
    diag0 = eng.kernel_diag()
    eng.set_option('pchol_lookahead', 0)
    Lt0, idx0, _, _ = eng.pchol_build(K_RANK, diag=diag0, want_times=False)
    eng.set_option('pchol_lookahead', 1)
    Lt1, idx1, _, _ = eng.pchol_build(K_RANK, diag=diag0, want_times=False)
    assert torch.equal(idx0, idx1)
    assert float((Lt0 - Lt1).abs().max()) <= 1e-11 * float(Lt0.abs().max())


def test_woodbury_inverse_property(full):
    torch, eng = full['torch'], full['eng']
    Lt = full.get('Lt')
    if Lt is None:
        Lt = eng.pchol_build(K_RANK)[0]
    gen = torch.Generator(device=eng.device).manual_seed(2)
    v = torch.randn(eng.n, dtype=torch.float64, device=eng.device, generator=gen)
    # The identity P (L L^T + lam I) v == v holds for every lam; the formula (a - T^T T a)/lam amplifies
    # rounding by ||L L^T|| / lam, so the tight check uses lam = 1e-3 (error ~ 1e-16 * 1e2 / 1e-3) ...
    lam_t = 1e-3
    T_t = eng.woodbury_factor_(Lt.clone(), lam_t)
    w = Lt.t() @ (Lt @ v) + lam_t * v                  # (L L^T + lam I) v
    assert _rel(eng.precon_apply(T_t, lam_t, 1.0, w), v) < 1e-9
    del T_t
    # ... and at the solver's lam = 1e-10 the same round trip is only accurate to ~ eps * ||L L^T|| / lam
    w = Lt.t() @ (Lt @ v) + LAM * v
    T = eng.woodbury_factor_(Lt, LAM)
    assert _rel(eng.precon_apply(T, LAM, 1.0, w), v) < 1e-2
    # exact complement: vectors orthogonal to range(L) are scaled by 1/lam.  T T^T = I - lam W^{-1} ~ I
    u = torch.randn(K_RANK, dtype=torch.float64, device=eng.device, generator=gen)
    TTt_u = T @ (T.t() @ u)
    assert _rel(TTt_u, u) < 1e-4
    full['T'] = T


def test_symmetric_storage_is_half_and_matches(full):
    """The symmetric tile storage (packed bands, TMA kernel) holds about half of K and gives the same matvec."""
    torch, eng, K = full['torch'], full['eng'], full['K']
    assert eng.symop_storage_elems() < 0.51 * eng.n * eng.n
    gen = torch.Generator(device=eng.device).manual_seed(3)
    v = torch.randn(eng.n, dtype=torch.float64, device=eng.device, generator=gen)
    ref = eng.gemv(K, v, alpha=-1.0, shift=LAM)
    Ksym = eng.symop_assemble()
    out = eng.symop_apply(Ksym, v, alpha=-1.0, shift=LAM)
    assert _rel(out, ref) < 1e-12
    assert torch.equal(out, eng.symop_apply(Ksym, v, alpha=-1.0, shift=LAM))     # deterministic
    del Ksym


def test_one_launch_tile_pass_equals_tile_by_tile(full):
    """Sharded symmetric operator (emulated ranks of an 8- and a 2-rank run at full size): the pass over all tiles of a
    rank in ONE persistent launch is deterministic and agrees with the tile-by-tile passes to rounding (the strips are
    cut between CTAs at different units, so the row sums are associated differently), and the ranks' partials sum to
    K v."""
    torch, eng, K = full['torch'], full['eng'], full['K']
    gen = torch.Generator(device=eng.device).manual_seed(5)
    v = torch.randn(eng.n, dtype=torch.float64, device=eng.device, generator=gen)
    try:
        for world, ranks in ((8, (0, 3, 5, 7)), (2, (0, 1))):
            total = torch.zeros(eng.n, dtype=torch.float64, device=eng.device) if len(ranks) == world else None
            for rank in ranks:
                eng.set_layout(rank, world)
                Ksym = eng.symop_assemble()
                eng.set_option('symop_multi', 1)
                p1 = eng.symop_apply(Ksym, v, partial=True)
                again = eng.symop_apply(Ksym, v, partial=True)
                eng.set_option('symop_multi', 0)
                p0 = eng.symop_apply(Ksym, v, partial=True)
                assert _rel(p1, p0) < 1e-14, (world, rank)
                assert torch.equal(p1, again), (world, rank)
                if total is not None:
                    total += p1
                del Ksym, p0, p1, again
            if total is not None:
                assert _rel(total, eng.gemv(K, v, alpha=1.0, shift=0.0)) < 1e-12
    finally:
        eng.set_option('symop_multi', 1)
        eng.set_layout(0, 1)


def test_woodbury_at_bench_rank(full):
    """Same Woodbury identity at the rank the benchmark uses (k = 4839: 76 POTRF panels, 19 outer TRSM panels)."""
    torch, eng, inp = full['torch'], full['eng'], full['inp']
    k = inp['k']
    Lt = eng.pchol_build(k, want_times=False)[0]
    gen = torch.Generator(device=eng.device).manual_seed(4)
    v = torch.randn(eng.n, dtype=torch.float64, device=eng.device, generator=gen)
    lam_t = 1e-3
    w = Lt.t() @ (Lt @ v) + lam_t * v
    T = eng.woodbury_factor_(Lt, lam_t)
    assert _rel(eng.precon_apply(T, lam_t, 1.0, w), v) < 1e-9
    del T, Lt


def test_pcg_solution_checked_with_other_operator(full):
    torch, eng, K, inp = full['torch'], full['eng'], full['K'], full['inp']
    T = full.get('T')
    if T is None:
        T = eng.woodbury_factor_(eng.pchol_build(K_RANK)[0], LAM)
    y = torch.as_tensor(inp['y'], device=eng.device)
    tol = 1e-3
    x, iters, resid, info, bnrm2 = eng.pcg(y, LAM, tol, 5 * eng.n, K_local=K, T=T, precon_sign=1.0)
    assert info == 0 and iters > 1
    r = y - eng.matvec_free(x, alpha=-1.0, shift=LAM)              # residual with the matrix-free operator
    assert float(r.norm()) <= 1.001 * tol * float(y.norm())
    assert abs(float(r.norm()) - resid) <= 1e-6 * float(y.norm())
    x2, iters2, _, info2, _ = eng.pcg(y, LAM, tol, 5 * eng.n, K_local=None, T=T, precon_sign=1.0)
    assert info2 == 0 and abs(iters2 - iters) <= max(1, int(0.05 * iters))
