"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference in this container.

Imports ``/root/reference/src/sGDML/sgdml`` under python 3.12 / numpy 2 / scipy 1.18 with the four
compatibility shims listed in SURVEY.md section 8c.  Nothing is written to /root/reference and
no reference source is copied.  Used by ``tests/golden/make_golden.py`` to generate the frozen
golden vectors and by ``oracle/check_against_reference.py``; it cannot run on the GPU box
(/root/reference does not exist there) and no product code imports it.
"""
import sys
import types

import numpy as np

REFERENCE_SGDML = '/root/reference/src/sGDML'
# when a list is assigned here, legacy_cg appends ||r|| after every iteration (golden generation only)
RESID_HISTORY = None


def legacy_cg(A, b, x0=None, tol=1e-5, maxiter=None, M=None, callback=None, atol=None):
    """Stand-in for scipy-1.7.3 ``scipy.sparse.linalg.cg(tol=, atol=None)`` (shim 3).

    scipy >= 1.12 dropped ``tol``/legacy ``atol`` and the reference's ``_cg_status`` reads the local
    variable ``resid`` of *this* frame (iterative_solver.py:884), so the function keeps a local of
    that name.  Semantics restated from the legacy revcom driver: x0 = 0 unless given, stop when
    ||r|| <= tol*||b||; on first apparent convergence (after iteration > 1) the residual is recomputed
    as b - A x and re-tested; ``callback(x)`` fires at the start of every iteration and once more on
    exit (so the reference's ``num_iters`` = CG iterations + 1).
    """
    import scipy.sparse.linalg as spla

    A = spla.aslinearoperator(A)
    n = len(b)
    if maxiter is None:
        maxiter = n * 10
    psolve = (lambda r: r) if M is None else spla.aslinearoperator(M).matvec
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=float)
    bnrm2 = float(np.linalg.norm(b))
    resid = float(np.linalg.norm(A.matvec(x) - b))  # legacy _get_atol probe
    if resid <= tol:
        return x, 0
    atol_eff = tol if bnrm2 == 0 else tol * bnrm2
    resid = atol_eff
    r = b - A.matvec(x)
    rho_prev, p = None, None
    info = maxiter
    it = 0
    while it < maxiter:
        it += 1
        if callback is not None:
            callback(x)
        z = psolve(r)
        rho = float(np.dot(r, z))
        if it == 1:
            p = z.copy()
        else:
            p = z + (rho / rho_prev) * p
        q = A.matvec(p)
        alpha = rho / float(np.dot(p, q))
        x = x + alpha * p
        r = r - alpha * q
        rho_prev = rho
        resid = float(np.linalg.norm(r))
        if resid <= atol_eff and it > 1:
            r = b - A.matvec(x)
            resid = float(np.linalg.norm(r))
        if RESID_HISTORY is not None:
            RESID_HISTORY.append(resid)
        if resid <= atol_eff:
            info = 0
            break
    if callback is not None:
        callback(x)
    return x, info


def load_reference():
    """Install the shims and return the reference's ``sgdml`` package."""
    # shim 1: matplotlib is not installed (iterative_solver.py:29, dev_utils.py:1)
    if 'matplotlib' not in sys.modules:
        mpl = types.ModuleType('matplotlib')
        plt = types.ModuleType('matplotlib.pyplot')
        mpl.pyplot = plt
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = plt
    # shim 2: np.int was removed in numpy 1.24 (desc.py:260)
    if not hasattr(np, 'int'):
        np.int = int
    if REFERENCE_SGDML not in sys.path:
        sys.path.insert(0, REFERENCE_SGDML)

    import scipy.linalg
    import scipy.sparse.linalg

    # shim 3: legacy cg signature + frame-local 'resid'
    scipy.sparse.linalg.cg = legacy_cg

    # shim 4: eigh(eigvals=(lo,hi)) -> subset_by_index (iterative_solver.py:577)
    if not getattr(scipy.linalg.eigh, '_mlffpc_shim', False):
        _eigh = scipy.linalg.eigh

        def eigh(a, *args, eigvals=None, **kw):
            if eigvals is not None:
                kw['subset_by_index'] = list(eigvals)
            return _eigh(a, *args, **kw)

        eigh._mlffpc_shim = True
        scipy.linalg.eigh = eigh

    import sgdml  # noqa: E402  (the reference package)
    import sgdml.train  # noqa: F401
    import sgdml.solvers.iterative_solver  # noqa: F401

    return sgdml


def make_task(sgdml, dataset, n_train, perms, sig=10, solver_tol=1e-4):
    """A minimal ``task`` dict with the keys the solve path reads (train.py:431-453), built
    without ``create_task`` (which needs 1000 validation points and a random symmetry search)."""
    R_train = dataset['R'][:n_train]
    return {
        'type': 't',
        'code_version': sgdml.__version__,
        'dataset_name': dataset['name'].astype(str),
        'dataset_theory': dataset['theory'].astype(str),
        'z': dataset['z'],
        'R_train': R_train,
        'F_train': dataset['F'][:n_train],
        'E_train': dataset['E'][:n_train],
        'idxs_train': np.arange(n_train),
        'md5_train': 'synthetic',
        'idxs_valid': np.arange(n_train, n_train + 1),
        'md5_valid': 'synthetic',
        'sig': sig,
        'lam': 1e-15,
        'use_E': True,
        'use_E_cstr': False,
        'use_sym': perms.shape[0] > 1,
        'use_cprsn': False,
        'solver_name': 'cg',
        'solver_tol': solver_tol,
        'n_inducing_pts_init': 25,
        'interact_cut_off': None,
        'perms': perms,
        'truncated_cholesky': 1500,
    }
