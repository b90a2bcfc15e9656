#!/usr/bin/env python
"""GPU diagnostic: unpreconditioned CG on the golden eth_s1_m12 kernel with lam = 1e-2 -- residual histories of
(a) mlffpc_pcg, (b) the same recurrence as a torch host loop over the device GEMV, (c) the numpy oracle.
Question: is 27/28 (device) vs 30 (numpy) iterations a property of the system (summation-order sensitive) or a bug?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from conftest import load_golden
from mlff_preconditioner_b200.engine import Engine
from oracle import sgdml_oracle as orc

g = load_golden('eth_s1_m12')
eng = Engine(g['R_desc'], g['R_d_desc'], g['tril_perms_lin'], int(g['sig']), perms=g['perms'])
K = eng.kernel_assemble()
lam, tol = 1e-2, 1e-8
b = torch.as_tensor(g['y'], device=eng.device)
x, it, resid, info, bnrm2, hist = eng.pcg(b, lam, tol, 5000, K_local=K, want_hist=True)
print('device loop: it', it, 'resid', resid)
# torch host loop, same recurrence
A = lambda v: eng.gemv(K, v, alpha=-1.0, shift=lam)
xx = torch.zeros_like(b); r = b - A(xx); p = None; rho_prev = None; h2 = [float(r.norm())]
atol = tol * float(b.norm())
for j in range(1, 200):
    z = r.clone(); rho = torch.dot(r, z)
    p = z if j == 1 else z + (rho / rho_prev) * p
    q = A(p); alpha = rho / torch.dot(p, q)
    xx = xx + alpha * p; r = r - alpha * q; rho_prev = rho
    h2.append(float(r.norm()))
    if h2[-1] <= atol:
        break
print('torch host loop: it', j)
An = -g['K'] + lam * np.eye(eng.n)
h3 = []
def mv(v):
    return An @ v
x3, it3, res3, info3 = orc.pcg(mv, g['y'], lambda r: r.copy(), tol, 5000)
print('numpy oracle: it', it3)
# numpy CG with history
xx = np.zeros(eng.n); r = g['y'] - An @ xx; h3 = [np.linalg.norm(r)]
for j in range(1, 200):
    z = r.copy(); rho = r @ z
    p = z if j == 1 else z + (rho / rho_prev) * p
    q = An @ p; alpha = rho / (p @ q)
    xx = xx + alpha * p; r = r - alpha * q; rho_prev = rho
    h3.append(np.linalg.norm(r))
    if h3[-1] <= atol:
        break
m = min(len(hist), len(h2), len(h3))
for i in range(m):
    print(i, '%.6e %.6e %.6e' % (hist[i], h2[i], h3[i]))
print('atol', atol, 'eigs of A (lowest 8):', np.linalg.eigvalsh(An)[:8], 'largest', np.linalg.eigvalsh(An)[-1])
