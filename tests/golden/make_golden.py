"""Generate the frozen golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shims.py) on seeded synthetic inputs.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

Every array written here is an output of reference code: ``GDMLTrain._assemble_kernel_mat``
(train.py:1121), ``IterativeCholesky._assemble_kernel_mat_diag`` (iterative_cholesky.py:241),
``Iterative._init_kernel_operator`` (iterative_solver.py:383, torch-CPU path),
``incomplete_cholesky.pivoted_cholesky`` (incomplete_cholesky.py:24),
``IterativeCholesky._init_precon_operator`` (iterative_cholesky.py:115),
``Iterative._init_precon_operator`` / ``_lev_scores`` / ``solve`` (iterative_solver.py:95,447,620).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from mlff_preconditioner_b200 import synthetic  # noqa: E402

sgdml = ref_shims.load_reference()
from sgdml.train import GDMLTrain  # noqa: E402
from sgdml.utils.desc import Desc  # noqa: E402
from sgdml.solvers.iterative_solver import Iterative  # noqa: E402
from sgdml.solvers.iterative_cholesky import Iterative as IterativeCholesky  # noqa: E402
from sgdml.solvers import incomplete_cholesky as ichol  # noqa: E402

SIG = 10
LAM = 1e-10  # train.py:866
GT = GDMLTrain(use_torch=True)


def noop(*a, **k):
    pass


def prep(kind, M, perms, seed):
    ds = synthetic.make_dataset(kind, M + 2, seed=seed)
    task = ref_shims.make_task(sgdml, ds, M, perms, sig=SIG)
    task['lam'] = LAM
    N = ds['R'].shape[1]
    desc = Desc(N, max_processes=1)
    tril_perms = np.array([desc.perm(p) for p in perms])
    tril_perms_lin = (tril_perms + np.arange(perms.shape[0])[:, None] * desc.dim).flatten('F')
    R = task['R_train'].reshape(M, -1)
    R_desc, R_d_desc = desc.from_R(R, callback=noop)
    y = task['F_train'].ravel().copy()
    y_std = np.std(y)
    y /= y_std
    return task, desc, tril_perms_lin, R_desc, R_d_desc, y, y_std


def case_operators(name, kind, M, perms, seed, frac=0.1, solve_strs=(), solve_frac=0.25, tol=1e-4,
                   store_K=True):
    task, desc, tpl, R_desc, R_d_desc, y, y_std = prep(kind, M, perms, seed)
    N = task['R_train'].shape[1]
    n = 3 * N * M
    rng = np.random.default_rng(seed + 100)
    out = dict(kind=kind, M=M, N=N, perms=perms, seed=seed, sig=SIG, lam=LAM,
               R_train=task['R_train'], F_train=task['F_train'], R_desc=R_desc, R_d_desc=R_d_desc,
               tril_perms_lin=tpl, y=y, y_std=y_std)

    # a1: explicit kernel (full and a sorted column subset)
    K = GT._assemble_kernel_mat(R_desc, R_d_desc, tpl, SIG, desc, callback=noop).copy()
    cols = np.sort(rng.choice(n, size=max(3, n // 7), replace=False))
    K_panel = GT._assemble_kernel_mat(R_desc, R_d_desc, tpl, SIG, desc, col_idxs=cols, callback=noop).copy()
    assert np.abs(K_panel - K[:, cols]).max() < 1e-13 * np.abs(K).max()
    if store_K:
        out['K'] = K
    out['panel_cols'] = cols
    out['K_panel'] = K_panel
    out['K_fro'] = np.linalg.norm(K)
    out['K_rowsum'] = K.sum(axis=1)

    # a2: -diag(K)
    itc = IterativeCholesky(gdml_train=GT, desc=desc, task=task, callback=noop, use_torch=True)
    diag = itc._assemble_kernel_mat_diag(tril_perms_lin=tpl, sig=SIG, R_desc=R_desc, R_d_desc=R_d_desc, n=n)
    out['diag'] = diag

    # a3: matrix-free operator K v - lam v (torch-CPU path of the reference)
    it = Iterative(GT, desc, callback=noop, use_torch=True)
    K_op = it._init_kernel_operator(task, R_desc, R_d_desc, tpl, LAM, n, callback=noop)
    v = rng.standard_normal(n)
    out['v'] = v
    out['K_op_v'] = K_op.matvec(v)

    # a6: pivoted partial Cholesky of -K_op
    k = int(frac * n)
    itc.K_op = -K_op
    L, index_columns, info = ichol.pivoted_cholesky(get_col=itc._get_col_K, diagonal=diag, max_rank=k)
    out['chol_k'] = k
    out['L'] = L
    out['index_columns'] = index_columns

    # a7: Woodbury operator
    P_op, _ = itc._init_precon_operator(diag, -K_op, lam_regularization=LAM, break_percentage=frac)
    a = rng.standard_normal(n)
    out['a'] = a
    out['P_chol_a'] = P_op.matvec(a)

    # a8: Nystroem operator on random sorted columns, leverage scores
    idxs = np.sort(rng.choice(n, size=k, replace=False))
    P_nys = it._init_precon_operator(task, R_desc, R_d_desc, tpl, idxs, callback=noop)
    out['nys_idxs'] = idxs
    out['P_nys_a'] = P_nys.matvec(a)
    P_sb = it._init_precon_operator_sb(task, R_desc, R_d_desc, tpl, idxs, callback=noop)
    out['P_sb_a'] = P_sb.matvec(a)
    n_ind = min(M, int(max(np.ceil(frac * M), 1)))
    np.random.seed(7)
    scores, order = it._lev_scores(R_desc, R_d_desc, tpl, SIG, LAM, False, n_ind, callback=noop)
    out['lev_n_inducing'] = n_ind
    out['lev_scores'] = scores

    # a9: full solves (np.random.seed(0) immediately before each)
    task['solver_tol'] = tol
    for s in solve_strs:
        np.random.seed(0)
        it2 = Iterative(GT, desc, callback=noop, use_torch=True)
        alphas, num_iters, resid, rmse, ind, is_conv, info = it2.solve(
            task, R_desc, R_d_desc, tpl, y, y_std, break_percentage=solve_frac, str_preconditioner=s)
        out['solve_%s_alphas' % s] = alphas
        out['solve_%s_iters' % s] = num_iters
        out['solve_%s_resid' % s] = resid
        out['solve_%s_idxs' % s] = ind
        out['solve_%s_conv' % s] = is_conv
        print('   ', name, s, 'iters', num_iters, 'resid %.3e' % resid, 'conv', is_conv)
    out['solve_frac'] = solve_frac
    out['solve_tol'] = tol
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('wrote', name, 'n =', n, 'k =', k)


def case_large_cholesky(name, kind, M, seed, frac, tol):
    """n ~ 2k: pivot sequence, sampled L, iteration counts; K itself is not stored."""
    perms = np.arange(synthetic.base_structure(kind)[0].shape[0])[None]
    task, desc, tpl, R_desc, R_d_desc, y, y_std = prep(kind, M, perms, seed)
    N = task['R_train'].shape[1]
    n = 3 * N * M
    task['solver_tol'] = tol
    it = Iterative(GT, desc, callback=noop, use_torch=True)
    np.random.seed(0)
    alphas, num_iters, resid, rmse, ind, is_conv, info = it.solve(
        task, R_desc, R_d_desc, tpl, y, y_std, break_percentage=frac, str_preconditioner='cholesky')
    out = dict(kind=kind, M=M, N=N, perms=perms, seed=seed, sig=SIG, lam=LAM, R_train=task['R_train'],
               F_train=task['F_train'], y=y, y_std=y_std, frac=frac, tol=tol, alphas=alphas,
               num_iters=num_iters, resid=resid, is_conv=is_conv, index_columns=info['index_columns'],
               chol_k=int(frac * n))
    # the factor: rerun the factorisation to keep L (solve() does not return it)
    itc = IterativeCholesky(gdml_train=GT, desc=desc, task=task, callback=noop, use_torch=True)
    diag = itc._assemble_kernel_mat_diag(tril_perms_lin=tpl, sig=SIG, R_desc=R_desc, R_d_desc=R_d_desc, n=n)
    K_op = it._init_kernel_operator(task, R_desc, R_d_desc, tpl, LAM, n, callback=noop)
    itc.K_op = -K_op
    L, index_columns, _ = ichol.pivoted_cholesky(get_col=itc._get_col_K, diagonal=diag,
                                                 max_rank=int(frac * n))
    assert np.array_equal(index_columns, info['index_columns'])
    rng = np.random.default_rng(seed + 5)
    rows = np.sort(rng.choice(n, size=64, replace=False))
    out['diag'] = diag
    out['L_rows'] = rows
    out['L_sample'] = L[rows, :]
    out['L_colnorm'] = np.linalg.norm(L, axis=0)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('wrote', name, 'n =', n, 'iters', num_iters, 'resid %.3e' % resid, 'conv', is_conv)


def case_cfg1_solve(name='cfg1_nanotube_m9', frac_k=None):
    """BASELINE.json configs[0] at full size: nanotube-size N = 370, M = 9, n = 9990, k = 1954 (rule of thumb),
    tol 1e-6, one run of the unmodified ``Iterative.solve(str_preconditioner='cholesky')`` (about 7 minutes on 8
    cores).  The factor, the preconditioner operator and the residual history are captured by wrapping the
    reference's own callables at run time (nothing is modified)."""
    from bench import make_inputs
    inp = make_inputs('cfg1')
    n, k, M, N = inp['n'], inp['k'], inp['M'], inp['N']
    ds = synthetic.make_dataset(inp['kind'], M + 2, seed=0)
    perms = inp['perms']
    task = ref_shims.make_task(sgdml, ds, M, perms, sig=SIG, solver_tol=inp['tol'])
    task['lam'] = LAM
    desc = Desc(N, max_processes=1)
    R_desc, R_d_desc, tpl, y, y_std = inp['R_desc'], inp['R_d_desc'], inp['tpl'], inp['y'], inp['y_std']
    # the repo's host Desc against the reference's (the GPU test regenerates the descriptors from R_train)
    Rd_ref, Rdd_ref = desc.from_R(task['R_train'].reshape(M, -1), callback=noop)
    assert np.abs(Rd_ref - R_desc).max() <= 1e-15 * np.abs(Rd_ref).max()
    assert np.abs(Rdd_ref - R_d_desc).max() <= 1e-14 * np.abs(Rdd_ref).max()
    R_desc, R_d_desc = Rd_ref, Rdd_ref
    captured = {}
    orig_pc = ichol.pivoted_cholesky

    def pc_capture(*a, **kw):
        res = orig_pc(*a, **kw)
        captured['L'], captured['index_columns'] = res[0], res[1]
        return res

    orig_init = IterativeCholesky._init_precon_operator

    def init_capture(self, *a, **kw):
        res = orig_init(self, *a, **kw)
        captured['P_op'] = res[0]
        return res

    ichol.pivoted_cholesky = pc_capture
    IterativeCholesky._init_precon_operator = init_capture
    ref_shims.RESID_HISTORY = []
    try:
        it = Iterative(GT, desc, callback=noop, use_torch=True)
        frac = (k + 0.5) / n
        alphas, num_iters, resid, rmse, ind, is_conv, info = it.solve(
            task, R_desc, R_d_desc, tpl, y, y_std, break_percentage=frac, str_preconditioner='cholesky')
    finally:
        ichol.pivoted_cholesky = orig_pc
        IterativeCholesky._init_precon_operator = orig_init
    hist = np.array(ref_shims.RESID_HISTORY)
    ref_shims.RESID_HISTORY = None
    L = captured['L']
    assert L.shape == (n, k) and np.array_equal(captured['index_columns'], info['index_columns'])
    rng = np.random.default_rng(11)
    rows = np.sort(rng.choice(n, size=48, replace=False))
    colsel = np.array([0, 1, 2, k // 4, k // 2, k - 2, k - 1])
    a = rng.standard_normal(n)
    v = rng.standard_normal(n)
    K_op = it._init_kernel_operator(task, R_desc, R_d_desc, tpl, LAM, n, callback=noop)
    K_op_v = K_op.matvec(v)
    itc = IterativeCholesky(gdml_train=GT, desc=desc, task=task, callback=noop, use_torch=True)
    diag = itc._assemble_kernel_mat_diag(tril_perms_lin=tpl, sig=SIG, R_desc=R_desc, R_d_desc=R_d_desc, n=n)
    pcols = np.sort(rng.choice(n, size=24, replace=False))
    K_panel = GT._assemble_kernel_mat(R_desc, R_d_desc, tpl, SIG, desc, col_idxs=pcols, callback=noop).copy()
    out = dict(kind=inp['kind'], M=M, N=N, perms=perms, sig=SIG, lam=LAM, R_train=task['R_train'],
               F_train=task['F_train'], y=y, y_std=y_std, chol_k=k, frac=frac, tol=inp['tol'], alphas=alphas,
               num_iters=num_iters, resid=resid, is_conv=is_conv, index_columns=info['index_columns'],
               resid_hist=hist, L_rows=rows, L_sample=L[rows, :], L_cols=colsel, L_colsample=L[:, colsel],
               L_colnorm=np.linalg.norm(L, axis=0), a=a, P_chol_a=captured['P_op'].matvec(a), v=v, K_op_v=K_op_v,
               diag=diag, panel_cols=pcols, K_panel=K_panel)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('wrote', name, 'n =', n, 'k =', k, 'iters', num_iters, 'resid %.3e' % resid, 'conv', is_conv)


def case_predict(name, kind, M, perms, seed, n_query=17, frac=0.2, tol=1e-6):
    """Reference ``GDMLTrain.train`` (model dict incl. R_d_desc_alpha and the integration constant of
    ``_recov_int_const``, train.py:597-702, :972-1119) followed by ``GDMLPredict.predict`` on geometries that are not
    in the training set, both CPU routes of the reference (numpy worker predict.py:72-234 and the torch module
    torchtools.py:172-272)."""
    from sgdml.predict import GDMLPredict

    ds = synthetic.make_dataset(kind, M + n_query, seed=seed)
    task = ref_shims.make_task(sgdml, ds, M, perms, sig=SIG, solver_tol=tol)
    np.random.seed(0)
    model = GT.train(task, callback=noop, break_percentage=frac, str_preconditioner='cholesky')
    assert model['is_conv']
    Rq = ds['R'][M:M + n_query].reshape(n_query, -1)
    E_np, F_np = GDMLPredict(model, use_torch=False).predict(Rq)
    E_t, F_t = GDMLPredict(model, use_torch=True).predict(Rq)
    assert np.abs(E_np - E_t).max() < 1e-9 * np.abs(E_np).max() and np.abs(F_np - F_t).max() < 1e-9 * np.abs(F_np).max()
    # training-mode prediction (descriptors given), as _recov_int_const uses it
    N = ds['R'].shape[1]
    desc = Desc(N, max_processes=1)
    Rtr = task['R_train'].reshape(M, -1)
    R_desc, R_d_desc = desc.from_R(Rtr, callback=noop)
    E_tr, F_tr = GDMLPredict(model, use_torch=False).predict(Rtr, R_desc, R_d_desc)
    out = dict(kind=kind, M=M, N=N, perms=perms, seed=seed, sig=SIG, frac=frac, tol=tol,
               R_train=task['R_train'], F_train=task['F_train'], E_train=task['E_train'], z=ds['z'],
               alphas_F=model['alphas_F'], R_d_desc_alpha=model['R_d_desc_alpha'], R_desc_T=model['R_desc'],
               tril_perms_lin=model['tril_perms_lin'], c=model['c'], std=model['std'], lam=model['lam'],
               solver_iters=model['solver_iters'], use_E=model['use_E'],
               R_query=Rq, E_query=E_np, F_query=F_np, E_train_pred=E_tr, F_train_pred=F_tr,
               model_keys=np.array(sorted(model.keys())))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('wrote', name, 'iters', model['solver_iters'], 'c', model['c'], 'use_E', model['use_E'])


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'predict':
        case_predict('predict_eth_s6_m24', 'ethanol', 24, synthetic.ethanol_perms(), seed=6)
        case_predict('predict_asp_s1_m10', 'aspirin', 10, np.arange(21)[None], seed=7)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'cfg1':
        case_cfg1_solve()
        sys.exit(0)
    id9 = np.arange(9)[None]
    all_strs = ('cholesky', 'random_scores', 'lev_scores', 'inverse_lev', 'lev_random',
                'truncated_cholesky', 'truncated_cholesky_custom')
    case_operators('eth_s1_m12', 'ethanol', 12, id9, seed=0, frac=0.1, solve_strs=all_strs)
    case_operators('eth_s6_m6', 'ethanol', 6, synthetic.ethanol_perms(), seed=1, frac=0.15,
                   solve_strs=('cholesky', 'random_scores'))
    case_operators('asp_s1_m4', 'aspirin', 4, np.arange(21)[None], seed=2, frac=0.1,
                   solve_strs=('cholesky',))
    case_operators('grid40_s1_m3', 'grid40', 3, np.arange(40)[None], seed=3, frac=0.1,
                   solve_strs=('cholesky',))
    case_large_cholesky('eth_s1_m80_chol', 'ethanol', 80, seed=4, frac=0.1, tol=1e-4)
