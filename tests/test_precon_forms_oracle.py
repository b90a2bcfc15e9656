"""CPU: the two evaluations of the pivoted-Cholesky preconditioner (L L^T + lam I)^{-1} -- the reference's Woodbury
formula (iterative_cholesky.py:141-148) and the orthonormal-basis form the device uses by default -- are the same
operator, and the second one is the accurate one at the solver's lam = 1e-10 (DESIGN.md "Woodbury accuracy").
Truth = dense solve in extended precision (np.longdouble) on a golden kernel matrix."""
import numpy as np

from conftest import load_golden
from oracle import sgdml_oracle as orc


def _truth(L, lam, a):
    Ld = L.astype(np.longdouble)
    A = Ld @ Ld.T + np.longdouble(lam) * np.eye(L.shape[0], dtype=np.longdouble)
    # Cholesky solve in extended precision (numpy.linalg does not take longdouble): plain loops, n is small
    n = A.shape[0]
    C = np.zeros_like(A)
    for j in range(n):
        C[j, j] = np.sqrt(A[j, j] - C[j, :j] @ C[j, :j])
        C[j + 1:, j] = (A[j + 1:, j] - C[j + 1:, :j] @ C[j, :j]) / C[j, j]
    y = np.zeros(n, dtype=np.longdouble)
    ad = a.astype(np.longdouble)
    for i in range(n):
        y[i] = (ad[i] - C[i, :i] @ y[:i]) / C[i, i]
    x = np.zeros(n, dtype=np.longdouble)
    for i in range(n - 1, -1, -1):
        x[i] = (y[i] - C[i + 1:, i] @ x[i + 1:]) / C[i, i]
    return x.astype(float)


def test_forms_agree_and_orthonormal_is_the_accurate_one():
    g = load_golden('eth_s1_m12')
    A = -g['K']
    n = A.shape[0]
    k = n // 4
    L, _ = orc.pivoted_cholesky(lambda i: A[:, i], g['diag'], k)
    rng = np.random.default_rng(0)
    a = rng.standard_normal(n)
    c = rng.standard_normal(k)
    a_range = L @ c                                   # a vector in range(L): the part Woodbury's subtraction cancels

    def rel(x, y):
        return np.linalg.norm(x - y) / np.linalg.norm(y)

    # moderate lam: both forms equal the dense inverse
    lam = 1e-3
    T = orc.woodbury_factor(L, lam)
    Qt, Mk = orc.orthonormal_factor(L, lam)
    assert np.abs(Qt @ Qt.T - np.eye(k)).max() < 1e-13
    ref = _truth(L, lam, a)
    assert rel(orc.woodbury_apply(T, lam, a), ref) < 1e-10
    assert rel(orc.orthonormal_apply(Qt, Mk, lam, a), ref) < 1e-10

    # the solver's lam: same operator, but only the orthonormal form keeps the range part accurate
    lam = 1e-10
    T = orc.woodbury_factor(L, lam)
    Qt, Mk = orc.orthonormal_factor(L, lam)
    ref = _truth(L, lam, a_range)
    err_w = rel(orc.woodbury_apply(T, lam, a_range), ref)
    err_o = rel(orc.orthonormal_apply(Qt, Mk, lam, a_range), ref)
    assert err_o < 1e-6 and err_o < 0.1 * err_w, (err_w, err_o)
