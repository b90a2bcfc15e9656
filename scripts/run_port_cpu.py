#!/usr/bin/env python
"""Full CPU solve with the oracle port of the reference's algorithm (oracle/sgdml_oracle.py): pivoted Cholesky with
numpy's einsum Schur update, Woodbury factor with LAPACK, legacy-scipy PCG with the torch-CPU kernel operator -- every
phase timed, the iteration count and the residual history recorded.  This is the measured (not extrapolated) CPU
number for a workload and the source of the reference-formula iteration count in bench_constants.json.

    python scripts/run_port_cpu.py cfg2 [--threads 8] [--colgen assemble|matvec] [--out profiles/...json]

--colgen matvec   pivot columns as the reference gets them, K_op e_i (one full matvec per column)
--colgen assemble pivot columns from the explicit kernel formula (same entries to ~1e-16, milliseconds per column);
                  the time the reference would spend is k * t_matvec and is reported separately
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
    ap.add_argument('workload')
    ap.add_argument('--threads', type=int, default=os.cpu_count())
    ap.add_argument('--colgen', default='assemble', choices=['assemble', 'matvec'])
    ap.add_argument('--out', default=None)
    ap.add_argument('--M', type=int, default=None)
    args = ap.parse_args()
    import torch
    from threadpoolctl import threadpool_limits

    from bench import make_inputs
    from oracle import sgdml_oracle as orc

    torch.set_num_threads(args.threads)
    threadpool_limits(limits=args.threads)
    inp = make_inputs(args.workload, M_override=args.M)
    n, k, M, lam, tol = inp['n'], inp['k'], inp['M'], 1e-10, inp['tol']
    R_desc, R_d_desc, tpl, y = inp['R_desc'], inp['R_d_desc'], inp['tpl'], inp['y']
    D = R_desc.shape[1]
    Rs_t = torch.from_numpy(np.ascontiguousarray(inp['task']['R_train']))
    Xp_t = torch.from_numpy(np.ascontiguousarray(orc.permuted_rows(R_desc, tpl).reshape(-1, D)))
    n_mv = [0]

    def A(v):  # (-K_op) v = -K v + lam v
        n_mv[0] += 1
        return -orc.kernel_matvec_torch_cpu(Rs_t, Xp_t, R_d_desc, tpl, 10, v) + lam * v

    t0 = time.perf_counter()
    A(np.ones(n))
    t_mv = time.perf_counter() - t0
    t0 = time.perf_counter()
    diag = orc.kernel_mat_diag(R_desc, R_d_desc, tpl, 10)
    t_diag = time.perf_counter() - t0
    if args.colgen == 'matvec':
        def get_col(i):
            e = np.zeros(n)
            e[i] = 1
            return A(e) - lam * e * 0 if False else A(e)
    else:
        def get_col(i):
            c = -orc.assemble_kernel_mat(R_desc, R_d_desc, tpl, 10, col_idxs=np.array([i]))[:, 0]
            c[i] += lam
            return c
    print('n', n, 'k', k, 't_matvec %.3f s' % t_mv, flush=True)
    t0 = time.perf_counter()
    L, index_columns = orc.pivoted_cholesky(get_col, diag, k)
    t_pchol = time.perf_counter() - t0
    print('pivoted cholesky %.1f s' % t_pchol, flush=True)
    t0 = time.perf_counter()
    T = orc.woodbury_factor(L, lam)
    t_fac = time.perf_counter() - t0
    del L
    print('woodbury factor %.1f s' % t_fac, flush=True)
    hist = []
    t_app = [0.0]

    def psolve(a):
        t1 = time.perf_counter()
        z = orc.woodbury_apply(T, lam, a)
        t_app[0] += time.perf_counter() - t1
        return z

    def mv(v):
        q = A(v)
        if len(hist) % 50 == 0:
            print('  matvec', len(hist), flush=True)
        hist.append(0)
        return q

    t0 = time.perf_counter()
    x, iters, resid, info = orc.pcg(mv, y, psolve, tol, 5 * n)
    t_cg = time.perf_counter() - t0
    t_colgen_ref = k * t_mv if args.colgen == 'assemble' else 0.0
    out = {'impl': 'oracle port of the reference algorithm (numpy/scipy/torch-CPU), full solve measured',
           'workload': args.workload, 'n': n, 'k': k, 'tol': tol, 'threads': args.threads, 'colgen': args.colgen,
           'cg_iters': int(iters), 'num_iters_reference_convention': int(iters) + 1, 'converged': info == 0,
           'rel_resid': float(resid / np.linalg.norm(y)), 't_matvec_s': t_mv, 't_diag_s': t_diag,
           't_pchol_s': t_pchol, 't_pchol_columns_as_reference_s': t_colgen_ref, 't_factor_s': t_fac, 't_cg_s': t_cg,
           't_apply_total_s': t_app[0],
           'total_as_reference_s': t_diag + t_pchol + t_colgen_ref + t_fac + t_cg,
           'index_columns_first16': [int(i) for i in index_columns[:16]]}
    print(json.dumps(out))
    if args.out:
        with open(args.out, 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
