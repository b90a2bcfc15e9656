#!/usr/bin/env python
"""Build-container only: time the UNMODIFIED reference solver (Iterative.solve, 'cholesky' preconditioner, torch on
the CPU because no GPU is visible here) on a synthetic system, through oracle/ref_shims.py.  Nothing of the reference
is copied; /root/reference must be mounted.

    python scripts/run_reference_cpu.py cfg1            # nanotube-size N = 370, M = 9, n = 9990, k = 1954, tol 1e-6
    python scripts/run_reference_cpu.py small [tol]     # ethanol-size M = 300, n = 8100
Prints one JSON line (wall times from the reference's own info dict)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import make_inputs
    from oracle import ref_shims

    workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg1'
    tol = float(sys.argv[2]) if len(sys.argv) > 2 else None
    inp = make_inputs(workload, tol_override=tol)
    torch.set_num_threads(os.cpu_count())
    sgdml = ref_shims.load_reference()
    from sgdml.solvers.iterative_solver import Iterative
    from sgdml.train import GDMLTrain
    from sgdml.utils.desc import Desc
    from mlff_preconditioner_b200 import synthetic

    ds = synthetic.make_dataset(inp['kind'], inp['M'] + 2, seed=0)
    task = ref_shims.make_task(sgdml, ds, inp['M'], inp['perms'], sig=10, solver_tol=inp['tol'])
    task['lam'] = 1e-10
    desc = Desc(inp['N'], max_processes=1)
    gt = GDMLTrain(use_torch=True)
    it = Iterative(gt, desc, callback=lambda *a, **k: None, use_torch=True)
    n, k = inp['n'], inp['k']
    t0 = time.perf_counter()
    alphas, num_iters, resid, rmse, ind, is_conv, info = it.solve(
        task, inp['R_desc'], inp['R_d_desc'], inp['tpl'], inp['y'], inp['y_std'], break_percentage=(k + 0.5) / n,
        str_preconditioner='cholesky')
    wall = time.perf_counter() - t0
    print(json.dumps({'impl': 'reference (unmodified, via oracle/ref_shims.py)', 'workload': workload, 'n': n, 'k': k,
                      'tol': inp['tol'], 'cores': os.cpu_count(), 'wall_s': wall, 'num_iters': int(num_iters),
                      'is_conv': bool(is_conv), 'rel_resid': float(resid / np.linalg.norm(inp['y'])),
                      'total_time_preconditioner': float(info['total_time_preconditioner']),
                      'total_time_cg': float(info['total_time_cg']), 'total_time_solve': float(info['total_time_solve'])}))


if __name__ == '__main__':
    main()
