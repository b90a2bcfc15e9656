#!/usr/bin/env python
"""CPU experiment (numpy oracle, dense K): PCG iteration counts for different evaluations of the SAME pivoted-Cholesky
preconditioner (L L^T + lam I)^{-1}, lam = 1e-10 -- the reference Woodbury formula, the orthonormal-basis form, the
orthonormal form with the complement projected twice, an accurate QR/eigh form, and an extended-precision apply.
usage: python scripts/precon_forms_cg_experiment.py M k/n tol      (ethanol-size synthetic geometries)
Results of this round: profiles/r01v_precon_forms_cpu.txt"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.linalg
from bench import make_inputs, WORKLOADS
from oracle import sgdml_oracle as orc
M=int(sys.argv[1]); kfrac=float(sys.argv[2]); tol=float(sys.argv[3])
WORKLOADS['t']=('ethanol', M, tol)
inp=make_inputs('t'); n=inp['n']; lam=1e-10
K=orc.assemble_kernel_mat(inp['R_desc'],inp['R_d_desc'],inp['tpl'],10)
A=-K+lam*np.eye(n)
k=int(kfrac*n)
L,_=orc.pivoted_cholesky(lambda i:(-K)[:,i], -np.diag(K), k)
print('n',n,'k',k,'|LtL|',np.linalg.norm(L.T@L,2))
y=inp['y']
mv=lambda v: A@v
hist=[]
def run(name, psolve, maxiter=20000):
    x,it,res,info=orc.pcg(mv,y,psolve,tol,maxiter)
    print('%-34s iters %5d info %d resid/|b| %.2e' % (name,it,info,res/np.linalg.norm(y)), flush=True)
T=orc.woodbury_factor(L,lam)
run('woodbury (reference)', lambda r: orc.woodbury_apply(T,lam,r))
Qt,Mk=orc.orthonormal_factor(L,lam)
run('orthonormal', lambda r: orc.orthonormal_apply(Qt,Mk,lam,r))
def ortho2(r):
    w=Qt@r
    rp=r-Qt.T@w
    w2=Qt@rp
    rp=rp-Qt.T@w2          # twice is enough
    return rp/lam + Qt.T@(Mk@(w+w2))
run('orthonormal + reorthogonalised', ortho2)
# accurate reference: Householder QR + eigen-decomposition, complement projected twice
Q,R=np.linalg.qr(L)  # n x k
S=R@R.T
s,V=np.linalg.eigh(S)
U=Q@V
def exactish(r):
    w=U.T@r
    rp=r-U@w
    w2=U.T@rp
    rp=rp-U@w2
    return rp/lam + U@((w+w2)/(s+lam))
run('QR/eigh + reorthogonalised', exactish)
# extended precision apply of the orthonormal form
Ql=Qt.astype(np.longdouble); Ml=Mk.astype(np.longdouble)
def ld(r):
    rl=r.astype(np.longdouble); w=Ql@rl
    return ((rl-Ql.T@w)/np.longdouble(lam)+Ql.T@(Ml@w)).astype(float)
run('orthonormal, longdouble apply', ld)
