#!/usr/bin/env python
"""CPU experiment (numpy oracle, dense K): what moves the CG iteration count of the projected-form preconditioner in the
last decade of the residual -- rounding-level noise in the OPERATOR, or the accuracy of the defect matrix E?

 (a) operator noise: y = A p is multiplied entrywise by (1 + s N(0,1)) with s = 1e-16, 3e-16, 1e-15 (several seeds);
     E in extended precision;
 (b) E noise: E (extended precision) + symmetric Gaussian noise of size 1e-18 ... 1e-17 (several seeds); exact operator.

usage: python scripts/tail_sensitivity_experiment.py M k/n tol   (e.g. 300 0.1 1e-6: n = 8100, k = 810)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_inputs, WORKLOADS
from oracle import sgdml_oracle as orc
M = int(sys.argv[1]); kfrac = float(sys.argv[2]); tol = float(sys.argv[3])
WORKLOADS['t'] = ('ethanol', M, tol)
inp = make_inputs('t'); n = inp['n']; lam = 1e-10
K = orc.assemble_kernel_mat(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10)
A = -K + lam * np.eye(n)
k = int(kfrac * n)
L, _ = orc.pivoted_cholesky(lambda i: (-K)[:, i], -np.diag(K), k)
y = inp['y']
Qt, Mk = orc.orthonormal_factor(L, lam)
Ql = Qt.astype(np.longdouble)
E0 = np.asarray(Ql @ Ql.T - np.eye(k, dtype=np.longdouble), dtype=float)
print('n', n, 'k', k, '|E|max %.2e' % np.abs(E0).max(), flush=True)
def apply_with(E):
    def f(r):
        w = Qt @ r
        return (r - Qt.T @ (w - E @ w)) / lam + Qt.T @ (Mk @ w)
    return f
def run(name, mv, psolve):
    x, it, res, info = orc.pcg(mv, y, psolve, tol, 20000)
    print('%-52s iters %5d  resid/|b| %.2e' % (name, it, res / np.linalg.norm(y)), flush=True)
    return it
run('exact operator, E extended precision', lambda v: A @ v, apply_with(E0))
for s in (1e-16, 3e-16, 1e-15):
    for seed in range(3):
        rng = np.random.default_rng(100 + seed)
        run('operator noise %.0e (seed %d), E extended' % (s, seed), lambda v: (A @ v) * (1.0 + s * rng.standard_normal(n)), apply_with(E0))
for s in (1e-18, 3e-18, 1e-17):
    for seed in range(3):
        rng = np.random.default_rng(200 + seed)
        N = rng.standard_normal((k, k)); N = (N + N.T) / np.sqrt(2)
        run('E noise %.0e (seed %d), exact operator' % (s, seed), lambda v: A @ v, apply_with(E0 + s * N))
