#!/usr/bin/env python
"""cfg3 (BASELINE.json configs[2]): preconditioner rank sweep on the cfg2 operator -- build time, apply time,
CG iterations and time-to-solution per (variant, k).  One JSON line per point.

    python scripts/sweep_cfg3.py --M 4000 --ks 100,500,2000,5000 --variants cholesky,random_scores --mode assembled
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--M', type=int, default=4000)
    ap.add_argument('--kind', default='ethanol')
    ap.add_argument('--ks', default='100,200,500,1000,2000,5000')
    ap.add_argument('--variants', default='cholesky,random_scores,lev_random,truncated_cholesky')
    ap.add_argument('--mode', default='assembled', choices=['assembled', 'assembled_sym', 'matrix_free'])
    ap.add_argument('--tol', type=float, default=1e-6)
    ap.add_argument('--maxiter', type=int, default=None)
    args = ap.parse_args()

    import torch
    from bench import make_inputs, WORKLOADS
    from mlff_preconditioner_b200.engine import Engine
    from mlff_preconditioner_b200.solvers.iterative_solver import Iterative

    WORKLOADS['sweep'] = (args.kind, args.M, args.tol)
    inp = make_inputs('sweep')
    n = inp['n']
    eng = Engine(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10, perms=inp['perms'])
    y_t = torch.as_tensor(inp['y'], device=eng.device)
    task = dict(inp['task'])
    task['kernel_mode'] = args.mode
    if args.maxiter:
        task['_maxiter'] = args.maxiter
    if args.mode == 'assembled':
        task['_K_buffer'] = eng.empty(eng.n_local, eng.n)
    elif args.mode == 'assembled_sym':
        task['_K_buffer'] = eng.empty(eng.symop_storage_elems())
    for variant in args.variants.split(','):
        for k in [int(x) for x in args.ks.split(',')]:
            frac = (k + 0.5) / n
            n_ind = min(inp['M'], int(max(np.ceil(frac * inp['M']), 1)))
            np.random.seed(0)
            it = Iterative(None, None)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            try:
                x, iters, resid, info, idxs, _, t_pre, t_cg = it.solve_device(task, eng, y_t, frac, variant, n_ind)
            except Exception as e:  # keep sweeping; report the failure
                print(json.dumps({'variant': variant, 'k': k, 'error': '%s: %s' % (type(e).__name__, e)}), flush=True)
                continue
            torch.cuda.synchronize()
            total = time.perf_counter() - t0
            st = it.timings['pcg_stats']
            print(json.dumps({
                'variant': variant, 'k': k, 'n': n, 'mode': args.mode, 'tol': args.tol,
                'build_s': t_pre, 'pchol_s': it.timings.get('pchol_build'), 'assemble_s': it.timings['assemble'],
                'cg_s': t_cg, 'cg_iters': iters, 'converged': info == 0, 'total_s': total,
                'apply_ms': st['precon_ms'] / max(st['op_calls'], 1), 'matvec_ms': st['op_ms'] / max(st['op_calls'], 1),
                'rel_resid': resid / float(np.linalg.norm(inp['y']))}), flush=True)
            del it, x
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
