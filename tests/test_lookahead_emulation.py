"""CPU: the blocked ("look-ahead") pivoted Cholesky of csrc/pchol.cu restated in numpy -- candidate panel with the
Schur correction of the first m0 columns folded in by one GEMM, steps that only apply the columns chosen since,
panel rebuilt (per-rank top lists merged by the pivot rule) when the arg-max is not a candidate -- gives the
SAME pivot sequence and the same factor as the reference algorithm (oracle.pivoted_cholesky), for one rank and
for emulated row shards."""
import numpy as np

from conftest import load_golden
from oracle import sgdml_oracle as orc

LA_C, LA_LCAP = 16, 32     # small so that many rebuilds happen on a small matrix


def _top_lists(diag, pos, m, shards):
    """Per shard: rows not chosen yet with the largest residual diagonal (>= LA_C of them or all), then the
    merge by (value descending, position ascending) -- pchol_topc_kernel + pchol_merge_kernel."""
    entries = []
    for (r0, r1) in shards:
        idx = np.arange(r0, r1)
        ok = (pos[idx] >= m) & (diag[idx] > 0)
        idx = idx[ok]
        if idx.size > LA_C:
            thr = np.sort(diag[idx])[-LA_C]
            idx = idx[diag[idx] >= thr][:LA_LCAP]
        entries.extend((-diag[i], pos[i], i) for i in idx)
    entries.sort()
    return [e[2] for e in entries[:LA_C]]


def lookahead_pivoted_cholesky(A, diagonal, k, shards):
    n = A.shape[0]
    diag = np.array(diagonal, dtype=float)
    index_columns = np.arange(n)
    pos = np.arange(n)
    L = np.zeros((n, k))
    cands, panel, m0, rebuilds = [], None, 0, 0
    for m in range(k):
        live = pos >= m
        best = max(np.nonzero(live)[0], key=lambda i: (diag[i], -pos[i]))     # first maximum in permuted order
        pi = int(best)
        e, i_arg = index_columns[m], pos[pi]
        index_columns[m], index_columns[i_arg] = pi, e
        pos[pi], pos[e] = m, i_arg
        if pi not in cands:
            rebuilds += 1
            pos_sel = pos.copy()
            pos_sel[pi] = m      # the pivot still counts as "not chosen" for the selection (pos >= m)
            cands = _top_lists(diag, pos_sel, m, shards)
            assert pi in cands
            panel = A[:, cands] - L[:, :m] @ L[cands, :m].T      # one GEMM over the whole factor
            m0 = m
        col = panel[:, cands.index(pi)] - L[:, m0:m] @ L[pi, m0:m]
        piv = np.sqrt(diag[pi])
        rest = pos > m
        L[rest, m] = col[rest] / piv
        L[pi, m] = piv
        diag[rest] -= L[rest, m] ** 2
    return L, index_columns, rebuilds


def test_lookahead_equals_reference_algorithm():
    g = load_golden('eth_s1_m12')
    A = -g['K']
    n = A.shape[0]
    k = n // 3
    L_ref, idx_ref = orc.pivoted_cholesky(lambda i: A[:, i], g['diag'], k)
    for shards in ([(0, n)], [(0, n // 2), (n // 2, n)], [(0, 100), (100, 230), (230, n)]):
        L, idx, rebuilds = lookahead_pivoted_cholesky(A, g['diag'], k, shards)
        assert np.array_equal(idx, idx_ref), shards
        assert np.abs(L - L_ref).max() <= 1e-12 * np.abs(L_ref).max()
        assert 1 < rebuilds < k          # the panel is reused for several steps but not for all of them
