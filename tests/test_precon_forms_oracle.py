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
    err_r = rel(orc.orthonormal_apply_reorth(Qt, Mk, lam, a_range), ref)
    assert err_r <= err_o < err_w < 1e-6, (err_w, err_o, err_r)


def test_twice_projected_apply_restores_the_iteration_count():
    """n = 2160, k = 216, lam = 1e-10, tol 1e-6: PCG with the reference's Woodbury formula needs ~1400 iterations,
    with the twice-projected orthonormal form ~890 (the count of an accurately orthonormal basis)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import WORKLOADS, make_inputs

    WORKLOADS['t80'] = ('ethanol', 80, 1e-6)
    inp = make_inputs('t80')
    n, lam = inp['n'], 1e-10
    K = orc.assemble_kernel_mat(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10)
    A = -K + lam * np.eye(n)
    L, _ = orc.pivoted_cholesky(lambda i: (-K)[:, i], -np.diag(K), n // 10)
    T = orc.woodbury_factor(L, lam)
    Qt, Mk = orc.orthonormal_factor(L, lam)
    _, it_w, _, info_w = orc.pcg(lambda v: A @ v, inp['y'], lambda r: orc.woodbury_apply(T, lam, r), 1e-6, 5 * n)
    _, it_r, _, info_r = orc.pcg(lambda v: A @ v, inp['y'], lambda r: orc.orthonormal_apply_reorth(Qt, Mk, lam, r), 1e-6, 5 * n)
    assert info_w == 0 and info_r == 0
    assert it_r < 0.75 * it_w, (it_w, it_r)


def test_sharded_twice_projected_apply_equals_unsharded():
    """The collective sequence of csrc/precon.cu for the twice-projected apply (two k-vector all-reduces), emulated
    with row shards: w = sum_s Qt_s r_s; rp_s = r_s - Qt_s^T w; w2 = sum_s Qt_s rp_s; z_s = (rp_s - Qt_s^T w2)/lam +
    Qt_s^T Mk (w + w2)."""
    g = load_golden('eth_s1_m12')
    A = -g['K']
    n = A.shape[0]
    k = n // 4
    L, _ = orc.pivoted_cholesky(lambda i: A[:, i], g['diag'], k)
    lam = 1e-10
    Qt, Mk = orc.orthonormal_factor(L, lam)
    r = np.random.default_rng(1).standard_normal(n)
    ref = orc.orthonormal_apply_reorth(Qt, Mk, lam, r)
    for cuts in ([0, n], [0, 135, n], [0, 81, 200, n]):
        sl = [slice(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
        w = sum(Qt[:, s] @ r[s] for s in sl)
        rp = [r[s] - Qt[:, s].T @ w for s in sl]
        w2 = sum(Qt[:, s] @ p for s, p in zip(sl, rp))
        mu = Mk @ (w + w2)
        z = np.concatenate([(p - Qt[:, s].T @ w2) / lam + Qt[:, s].T @ mu for s, p in zip(sl, rp)])
        assert np.linalg.norm(z - ref) <= 1e-9 * np.linalg.norm(ref), cuts


def test_defect_from_error_free_split_is_exact_to_1e_21():
    """Design study for the next device kernel: E = Qt Qt^T - I from an error-free head/tail split
    of the rows.  The head Gram is EXACT under plain fp64 accumulation in any order (so a tensor-pipe SYRK without any
    (hi, lo) fold would do), and the result agrees with exact rational arithmetic to ~1e-22 where the fold-per-16-columns
    scheme of the current kernel is good to ~1e-18 -- the margin the projected apply needs at k ~ 5000."""
    from fractions import Fraction

    rng = np.random.default_rng(3)
    k, n = 12, 30000
    X = rng.standard_normal((k, n)) * np.exp(rng.standard_normal((k, 1)))
    for _ in range(2):          # CholeskyQR2: rows orthonormal to working precision, |E| ~ 1e-15
        X = np.linalg.solve(np.linalg.cholesky(X @ X.T), X)
    E, hb = orc.gram_defect_split(X)
    assert 2 * hb + int(np.ceil(np.log2(n))) <= 53
    _, e = np.frexp(np.abs(X).max(axis=1))
    g = np.ldexp(1.0, e - hb)[:, None]
    Qh = np.rint(X / g) * g
    assert np.array_equal(Qh + (X - Qh), X)                                              # the split is error-free
    H1 = Qh @ Qh.T
    H2 = sum(Qh[:, c:c + 7] @ Qh[:, c:c + 7].T for c in range(0, n, 7))                   # another summation order
    H3 = (Qh[:, ::-1] @ Qh[:, ::-1].T)
    assert np.array_equal(H1, H2) and np.array_equal(H1, H3)                             # exact, hence order-independent
    E16 = orc.gram_defect(X, 16)
    worst, worst16 = 0.0, 0.0
    for (i, j) in ((0, 0), (5, 5), (7, 2), (11, 0), (11, 11)):
        exact = sum(Fraction(float(a)) * Fraction(float(b)) for a, b in zip(X[i], X[j])) - (1 if i == j else 0)
        worst = max(worst, abs(float(Fraction(float(E[i, j])) - exact)))
        worst16 = max(worst16, abs(float(Fraction(float(E16[i, j])) - exact)))
    assert worst < 1e-21, worst
    assert worst16 < 3e-17       # the current scheme (emulated): fine for tests, at the edge for k ~ 5000
    assert np.abs(E).max() < 1e-14 and np.abs(E - E.T).max() == 0.0
