"""Build-container only: bench.py's CPU arm times the oracle's torch-CPU kernel matvec.  It must be the SAME operator
as the reference's ``K_op.matvec`` (bit-identical here: same op sequence) and must not cost more -- a slower port
would inflate every GPU-over-CPU ratio the bench prints (round-1 verdict: the old port was 4-10x slower).
Skipped where /root/reference is not mounted (the GPU box)."""
import os
import time

import numpy as np
import pytest

REF = '/root/reference/src/sGDML/sgdml'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='the reference is only mounted in the build container')


def _best(f, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = f()
        ts.append(time.perf_counter() - t0)
    return out, min(ts)


def test_port_matvec_costs_what_the_reference_costs():
    import torch

    from bench import WORKLOADS, _cpu_threads, make_inputs
    from mlff_preconditioner_b200 import synthetic
    from oracle import ref_shims
    from oracle import sgdml_oracle as orc

    M = 1000
    WORKLOADS['t1000'] = ('ethanol', M, 1e-6)
    inp = make_inputs('t1000')
    n = inp['n']
    _cpu_threads()
    sgdml = ref_shims.load_reference()
    from sgdml.solvers.iterative_solver import Iterative
    from sgdml.train import GDMLTrain
    from sgdml.utils.desc import Desc

    ds = synthetic.make_dataset(inp['kind'], M + 2, seed=0)
    task = ref_shims.make_task(sgdml, ds, M, inp['perms'], sig=10, solver_tol=1e-6)
    task['lam'] = 1e-10
    noop = lambda *a, **k: None  # noqa: E731
    it = Iterative(GDMLTrain(use_torch=True), Desc(inp['N'], max_processes=1), callback=noop, use_torch=True)
    K_op = it._init_kernel_operator(task, inp['R_desc'], inp['R_d_desc'], inp['tpl'], 1e-10, n, callback=noop)
    v = np.random.default_rng(0).standard_normal(n)
    K_op.matvec(v)  # priming call returns v (iterative_solver.py:418-421)
    K_op.matvec(v)
    ref, t_ref = _best(lambda: K_op.matvec(v), 5)
    Rs_t = torch.from_numpy(np.ascontiguousarray(inp['task']['R_train']))
    Xp_t = torch.from_numpy(np.ascontiguousarray(orc.permuted_rows(inp['R_desc'], inp['tpl']).reshape(-1, 36)))
    orc.kernel_matvec_torch_cpu(Rs_t, Xp_t, inp['R_d_desc'], inp['tpl'], 10, v)
    out, t_port = _best(lambda: orc.kernel_matvec_torch_cpu(Rs_t, Xp_t, inp['R_d_desc'], inp['tpl'], 10, v), 5)
    assert np.abs(out - 1e-10 * v - ref).max() <= 1e-14 * np.abs(ref).max()
    assert t_port <= 1.3 * t_ref, (t_port, t_ref)
