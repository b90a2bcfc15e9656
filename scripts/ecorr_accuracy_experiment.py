#!/usr/bin/env python
"""CPU experiment (numpy oracle, dense K): how accurately must E = Qt Qt^T - I be known for the two-pass k-space
correction  z = (r - Qt^T (w - E w)) / lam + Qt^T Mk w,  w = Qt r  to reach the iteration count of the twice-projected
(four-pass) apply?  E is formed (a) in extended precision, (b) from fp64 products of `chunk`-column slices accumulated
exactly (what a DMMA k-tile + TwoSum accumulation gives), (c) extended precision + Gaussian noise of a given size.
usage: python scripts/ecorr_accuracy_experiment.py M k/n tol"""
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
print('n', n, 'k', k, '|A|', np.linalg.norm(A, 2), flush=True)
y = inp['y']
mv = lambda v: A @ v
def run(name, psolve, maxiter=20000):
    x, it, res, info = orc.pcg(mv, y, psolve, tol, maxiter)
    print('%-44s iters %5d info %d resid/|b| %.2e' % (name, it, info, res / np.linalg.norm(y)), flush=True)
Qt, Mk = orc.orthonormal_factor(L, lam)
run('orthonormal (two passes)', lambda r: orc.orthonormal_apply(Qt, Mk, lam, r))
run('twice projected (four passes)', lambda r: orc.orthonormal_apply_reorth(Qt, Mk, lam, r))
Ql = Qt.astype(np.longdouble)
E_ld = (Ql @ Ql.T - np.eye(k, dtype=np.longdouble))
print('|E|max', float(np.abs(E_ld).max()), flush=True)
def ecorr(E):
    E = np.asarray(E, dtype=float)
    def f(r):
        w = Qt @ r
        return (r - Qt.T @ (w - E @ w)) / lam + Qt.T @ (Mk @ w)
    return f
run('E extended precision', ecorr(E_ld))
for chunk in (16, 64, 256):
    acc = np.zeros((k, k), dtype=np.longdouble)
    for c in range(0, n, chunk):
        acc += (Qt[:, c:c + chunk] @ Qt[:, c:c + chunk].T)
    E_c = acc - np.eye(k, dtype=np.longdouble)
    print('chunk %d: |E_c - E_ld|max %.2e' % (chunk, float(np.abs(E_c - E_ld).max())))
    run('E from fp64 chunk-%d products, exact sum' % chunk, ecorr(E_c))
rng = np.random.default_rng(0)
for sig in (1e-16, 1e-17, 1e-18, 1e-19):
    N = rng.standard_normal((k, k)); N = (N + N.T) / np.sqrt(2)
    run('E extended + noise %.0e' % sig, ecorr(np.asarray(E_ld, dtype=float) + sig * N))
run('E from a plain fp64 Gram', ecorr(Qt @ Qt.T - np.eye(k)))
