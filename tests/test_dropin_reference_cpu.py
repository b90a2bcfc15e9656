"""Build-container only: the drop-in executed for real, once.

The UNMODIFIED reference drives the run -- ``src/tools/create_data.cg_steps`` -> ``GDMLTrain.train(solver='cg')``
(train.py:859-908) -> ``iterative_solver.Iterative(...).solve(...)`` -- with ``mlff_preconditioner_b200.patch.install()``
swapping in this package's ``Iterative``.  There is no GPU here, so the package's ``Engine`` is replaced BY THE TEST with
a stand-in that answers the same calls from the numpy oracle (test infrastructure; the product has no such path and
raises without CUDA).  What is checked is everything around the device calls: constructor and ``solve`` signatures as
the reference calls them, the 7-tuple, the ``info`` keys ``train``/``create_model``/``cg_steps`` consume, the model
the reference builds from our coefficients (its own ``_recov_int_const`` runs on them), and the result pickle
(create_data.py:117-169).  The same task through the unpatched reference gives the comparison run."""
import glob
import os
import pickle
import sys

import numpy as np
import pytest

REF_ROOT = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, 'src', 'tools')),
                                reason='the reference is only mounted in the build container')


class OracleEngine(object):
    """Stand-in for mlff_preconditioner_b200.engine.Engine on a box without a GPU: same attributes and call signatures
    (only those the 'cholesky' / assembled path of Iterative.solve uses), numpy oracle underneath."""

    def __init__(self, R_desc, R_d_desc, tril_perms_lin, sig, perms=None, device=None, rank=0, world=1, init_comm=None,
                 R=None):
        import torch
        from oracle import sgdml_oracle as orc

        self.torch, self.orc = torch, orc
        self.R_desc, self.R_d_desc, self.tpl, self.sig = np.asarray(R_desc), np.asarray(R_d_desc), tril_perms_lin, sig
        self.M, self.D = self.R_desc.shape
        self.N = orc.n_atoms_from_dim_d(self.D)
        self.dim_i = 3 * self.N
        self.n = self.n_local = self.M * self.dim_i
        self.row0, self.rank, self.world = 0, 0, 1
        self.device = torch.device('cpu')
        self.h2d_bytes = 0
        self._K = None
        self.options = {}
        self.last_pcg_stats = {'op_ms': 0.0, 'op_calls': 0, 'precon_ms': 0.0}

    def set_option(self, name, value):
        self.options[name] = value

    def close(self):
        pass

    def _kernel(self):
        if self._K is None:
            self._K = self.orc.assemble_kernel_mat(self.R_desc, self.R_d_desc, self.tpl, self.sig)
        return self._K

    def kernel_assemble(self, out=None):
        return self.torch.from_numpy(self._kernel())

    def pchol_build(self, k, diag=None, forced_pivots=None, want_times=True):
        K = self._kernel()
        d = self.orc.kernel_mat_diag(self.R_desc, self.R_d_desc, self.tpl, self.sig)
        L, idx = self.orc.pivoted_cholesky(lambda i: -K[:, i], d, int(k))
        return (self.torch.from_numpy(np.ascontiguousarray(L.T)), self.torch.from_numpy(idx.astype(np.int64)), None,
                np.full(int(k), 1e-3))

    def woodbury_factor_(self, Lt, lam):
        return self.torch.from_numpy(self.orc.woodbury_factor(Lt.numpy().T, lam))

    def pcg(self, b, lam, tol, maxiter, K_local=None, T=None, precon_sign=1.0, x0=None, want_hist=False, Mk=None, E=None,
            resume_iters=0, x_inout=None):
        A = -K_local.numpy() + lam * np.eye(self.n)
        Tn = T.numpy()
        x, it, resid, info = self.orc.pcg(lambda v: A @ v, b.numpy(), lambda a: precon_sign * self.orc.woodbury_apply(Tn, lam, a),
                                          tol, maxiter, x0=None if x0 is None else x0.numpy())
        self.last_pcg_stats = {'op_ms': 1.0, 'op_calls': it, 'precon_ms': 1.0}
        return self.torch.from_numpy(x), it, resid, info, float(np.linalg.norm(b.numpy()))


@pytest.fixture()
def reference_modules(monkeypatch):
    import torch

    from oracle import ref_shims

    sgdml = ref_shims.load_reference()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from src.tools import create_data  # loads the trainer a second time as src.sGDML.sgdml.* 
    import src.sGDML.sgdml.solvers.iterative_solver as ref_solver_mod

    from mlff_preconditioner_b200 import patch
    from mlff_preconditioner_b200.solvers import iterative_solver as ours

    monkeypatch.setattr(ours, 'Engine', OracleEngine)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    yield sgdml, create_data, ref_solver_mod, patch
    patch.uninstall()


def _task(sgdml, M=16, tol=1e-6):
    from mlff_preconditioner_b200 import synthetic
    from oracle import ref_shims

    ds = synthetic.make_dataset('ethanol', M + 2, seed=3)
    task = ref_shims.make_task(sgdml, ds, M, np.arange(9)[None], sig=10, solver_tol=tol)
    task['kernel_mode'] = 'assembled'   # 'auto' asks the CUDA driver for free memory
    return task


def test_cg_steps_through_the_patched_reference(reference_modules, tmp_path):
    sgdml, create_data, ref_solver_mod, patch = reference_modules
    from mlff_preconditioner_b200.solvers.iterative_solver import Iterative as Ours

    # comparison run: the reference end to end, nothing patched
    gt_ref = create_data.train.GDMLTrain(use_torch=True)
    ref_dir = tmp_path / 'ref'
    create_data.cg_steps(_task(sgdml), gt_ref, 16, 0.25, 'cholesky', path_to_script=str(ref_dir))
    ref_file = glob.glob(str(ref_dir / 'data_new' / '*' / 'cholesky' / 'n = 16' / '*.pickle'))
    assert len(ref_file) == 1
    with open(ref_file[0], 'rb') as f:
        ref = pickle.load(f)

    patched = patch.install()
    assert 'src.sGDML.sgdml.solvers.iterative_solver' in patched and ref_solver_mod.Iterative is Ours
    captured = {}
    orig_train = create_data.train.GDMLTrain.train

    def spy(self, *a, **kw):
        captured['model'] = orig_train(self, *a, **kw)
        return captured['model']

    create_data.train.GDMLTrain.train = spy
    try:
        our_dir = tmp_path / 'ours'
        # the reference's trainer is a process-wide singleton (train.py:267-273): the same instance serves both runs
        create_data.cg_steps(_task(sgdml), gt_ref, 16, 0.25, 'cholesky', path_to_script=str(our_dir))
    finally:
        create_data.train.GDMLTrain.train = orig_train
    our_file = glob.glob(str(our_dir / 'data_new' / '*' / 'cholesky' / 'n = 16' / '*.pickle'))
    assert len(our_file) == 1
    with open(our_file[0], 'rb') as f:
        res = pickle.load(f)

    # same result dictionary: keys, types, shapes (create_data.py:117-157)
    assert set(res.keys()) == set(ref.keys())
    for key in ref:
        assert type(res[key]) is type(ref[key]), key
    assert res['K.shape'] == ref['K.shape'] and res['k'] == ref['k'] and res['n_kernel'] == ref['n_kernel']
    assert res['t_cholesky'].shape == ref['t_cholesky'].shape and res['cholesky_percentage'] == ref['cholesky_percentage']
    # n = 432 with lam = 1e-10 is rounding-dominated: the stand-in multiplies with the dense matrix, the reference with
    # its matrix-free torch operator (the reference's own two operators differ by as much, tests/test_oracle_golden.py)
    assert abs(res['cholesky_cgsteps'] - ref['cholesky_cgsteps']) <= max(1, int(0.02 * ref['cholesky_cgsteps']))
    for key in ('total_time_preconditioner', 'total_time_solve', 'total_time_cg', 'time_cg_step'):
        assert isinstance(res[key], float) and res[key] >= 0.0
    # the model the REFERENCE built from our solve: its own create_model / _recov_int_const ran on our coefficients
    m = captured['model']
    assert m['is_conv'] is True and m['solver_iters'] == res['cholesky_cgsteps']
    for key in ('alphas_F', 'R_d_desc_alpha', 'c', 'std', 'inducing_pts_idxs', 'index_columns', 'L.shape', 'time_cholesky',
                'total_time_cholesky', 'solver_resid'):
        assert key in m, key
    assert m['use_E'] and np.isfinite(m['c']) and m['inducing_pts_idxs'].shape == (res['k'],)
    assert m['precon_form'] == 'woodbury'     # the default evaluates the reference's formula
