"""CPU, world_size = 2, gloo: the host side of the multi-GPU path.

 * the row-block partition (engine.shard_points / dist.shard_slices) tiles the rows exactly;
 * the NCCL-unique-id exchange (dist.broadcast_bytes) works through the default process group;
 * the sharded algorithms -- the exact sequence of collectives libmlffpc issues (per pivot step: an
   allgather of (value, position, index) candidates with first-position tie break + an allreduce-sum of the
   zero-padded pivot row; per CG iteration: an allgather of p and three scalar + one k-vector allreduce) --
   reproduce the single-process oracle when emulated with numpy shards over gloo.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _allreduce(x):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    dist.all_reduce(t)
    return t.numpy()


def _allgather(x):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [o.numpy() for o in outs]


def _sharded_pchol(A_rows, diag_local, row0, n, k):
    """Pivoted Cholesky on a row shard: A_rows = A[row0:row0+nl, :]."""
    nl = A_rows.shape[0]
    index_columns = np.arange(n)
    pos = np.arange(n)
    Lt = np.zeros((k, nl))
    diag = diag_local.copy()
    for m in range(k):
        cand = np.array([-1e300, 1e300, -1.0])
        for r in range(nl):
            g = row0 + r
            if pos[g] >= m:
                if diag[r] > cand[0] or (diag[r] == cand[0] and pos[g] < cand[1]):
                    cand = np.array([diag[r], float(pos[g]), float(g)])
        best = np.array([-1e300, 1e300, -1.0])
        for c in _allgather(cand):
            if c[0] > best[0] or (c[0] == best[0] and c[1] < best[1]):
                best = c
        pi = int(best[2])
        i_argmax, e = pos[pi], index_columns[m]
        index_columns[m], index_columns[i_argmax] = pi, e
        pos[pi], pos[e] = m, i_argmax
        lpiv = np.sqrt(best[0])
        lrow = np.zeros(max(m, 1))
        if row0 <= pi < row0 + nl and m > 0:
            lrow[:m] = Lt[:m, pi - row0]
        lrow = _allreduce(lrow)
        col = A_rows[:, pi]
        for r in range(nl):
            g = row0 + r
            if pos[g] > m:
                l = (col[r] - Lt[:m, r] @ lrow[:m]) / lpiv
                Lt[m, r] = l
                diag[r] -= l * l
            elif g == pi:
                Lt[m, r] = lpiv
    return Lt, index_columns


def _sharded_pcg(A_rows, b_local, T_local, lam, row0, n, tol, maxiter):
    nl = A_rows.shape[0]

    def precon(r):
        u = _allreduce(T_local @ r)
        return (r - T_local.T @ u) / lam

    def gather(v_local):
        return np.concatenate(_allgather(v_local))  # equal shards in this test

    x = np.zeros(nl)
    bnrm2 = np.sqrt(_allreduce(np.array([b_local @ b_local]))[0])
    r = b_local - A_rows @ gather(x)
    atol = tol * bnrm2
    rho_prev, p, it, info = None, None, 0, 1
    while it < maxiter:
        it += 1
        z = precon(r)
        rho = _allreduce(np.array([r @ z]))[0]
        p = z.copy() if it == 1 else z + (rho / rho_prev) * p
        q = A_rows @ gather(p)
        alpha = rho / _allreduce(np.array([p @ q]))[0]
        x += alpha * p
        r -= alpha * q
        rho_prev = rho
        resid = np.sqrt(_allreduce(np.array([r @ r]))[0])
        if resid <= atol and it > 1:
            r = b_local - A_rows @ gather(x)
            resid = np.sqrt(_allreduce(np.array([r @ r]))[0])
        if resid <= atol:
            info = 0
            break
    return x, it, info


def _sharded_pchol_merged(A_rows, diag_local, row0, n, k):
    """The round-2 step protocol of csrc/pchol.cu (host-free look-ahead stepping): ONE allgather per step carries every
    rank's best candidate together with that candidate's factor row, every rank picks the winner from the gathered
    messages, and the swap of step m is applied to index_columns / pos only at the start of step m + 1 -- the update
    of step m derives the new positions locally from the untouched arrays."""
    nl = A_rows.shape[0]
    index_columns = np.arange(n)
    pos = np.arange(n)
    Lt = np.zeros((k, nl))
    diag = diag_local.copy()
    pending = None
    for m in range(k):
        if pending is not None:                      # prepare kernel: apply the previous step's swap
            mp_, pi_ = pending
            i_argmax, e = pos[pi_], index_columns[mp_]
            index_columns[mp_], index_columns[i_argmax] = pi_, e
            pos[pi_], pos[e] = mp_, i_argmax
            pending = None
        cand = np.array([-1e300, 1e300, -1.0])
        for r in range(nl):
            g = row0 + r
            if pos[g] >= m and (diag[r] > cand[0] or (diag[r] == cand[0] and pos[g] < cand[1])):
                cand = np.array([diag[r], float(pos[g]), float(g)])
        msg = np.zeros(4 + k)
        msg[:3] = cand
        if cand[2] >= 0:
            msg[4:4 + m] = Lt[:m, int(cand[2]) - row0]
        best, lrow = None, None
        for c in _allgather(msg):                     # update kernel: winner among the ranks' messages
            if best is None or c[0] > best[0] or (c[0] == best[0] and c[1] < best[1]):
                best, lrow = c[:3], c[4:4 + m]
        pi = int(best[2])
        lpiv = np.sqrt(best[0])
        e, i_argmax = index_columns[m], pos[pi]       # arrays still hold the state before this step's swap
        col = A_rows[:, pi]
        for r in range(nl):
            g = row0 + r
            ps = i_argmax if (g == e and g != pi) else pos[g]
            if g == pi:
                Lt[m, r] = lpiv
            elif ps > m:
                l = (col[r] - Lt[:m, r] @ lrow) / lpiv
                Lt[m, r] = l
                diag[r] -= l * l
        pending = (m, pi)
    if pending is not None:                           # flush kernel
        mp_, pi_ = pending
        i_argmax, e = pos[pi_], index_columns[mp_]
        index_columns[mp_], index_columns[i_argmax] = pi_, e
        pos[pi_], pos[e] = mp_, i_argmax
    return Lt, index_columns


def _sharded_pcg_deferred(A_rows, b_local, T_local, lam, row0, n, tol, maxiter):
    """The round-2 loop of csrc/pcg.cu: ||r||^2 of iteration j travels with rho of iteration j + 1 in ONE two-element
    allreduce; the stopping test of iteration j runs at the start of the p-update of iteration j + 1 and freezes x, r, p
    (every later kernel is a no-op); on a hit after iteration 1 the true residual is recomputed once (legacy scipy)."""
    nl = A_rows.shape[0]

    def precon(r):
        u = _allreduce(T_local @ r)
        return (r - T_local.T @ u) / lam

    def gather(v_local):
        return np.concatenate(_allgather(v_local))

    x = np.zeros(nl)
    bnrm2 = np.sqrt(_allreduce(np.array([b_local @ b_local]))[0])
    r = b_local - A_rows @ gather(x)
    atol2 = (tol * bnrm2) ** 2
    red = np.zeros(4)                      # rho, rr(prev), p.q, rho_prev -- local until allreduced
    red[1] = r @ r
    p = np.zeros(nl)
    frozen, conv_iter, it = False, None, 0
    while it < maxiter + 1:                # one extra pass: the test of the last iteration
        it += 1
        z = precon(r)
        if not frozen:
            red[0] = r @ z
        red[:2] = _allreduce(red[:2])
        if it > 1 and not frozen and red[1] <= atol2:
            frozen, conv_iter = True, it - 1
        if frozen:
            # host side: recompute the true residual once (it > 1), accept or unfreeze
            if conv_iter > 1:
                r = b_local - A_rows @ gather(x)
                rr = _allreduce(np.array([r @ r]))[0]
            else:
                rr = red[1]
            if rr <= atol2:
                return x, conv_iter, 0
            frozen, it = False, conv_iter
            red[1] = r @ r
            continue
        if it > maxiter:
            break
        p = z.copy() if it == 1 else z + (red[0] / red[3]) * p
        q = A_rows @ gather(p)
        red[2] = _allreduce(np.array([p @ q]))[0]
        alpha = red[0] / red[2]
        x += alpha * p
        r -= alpha * q
        red[1] = r @ r                     # local; allreduced together with the next rho
        red[3] = red[0]
    return x, maxiter, 1


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from mlff_preconditioner_b200.dist import broadcast_bytes, dist_info, shard_slices
        from mlff_preconditioner_b200.engine import shard_points
        from oracle import sgdml_oracle as orc

        assert dist_info()[:2] == (rank, world)
        ident = bytes(range(128)) if rank == 0 else bytes(128)
        assert broadcast_bytes(ident, src=0) == bytes(range(128))

        g = load_golden('eth_s1_m12')
        M, dim_i = 12, 27
        n = M * dim_i
        pt0, pt1 = shard_points(M, rank, world)
        row0, row1 = shard_slices(M, dim_i, world)[rank]
        assert (row0, row1) == (pt0 * dim_i, pt1 * dim_i)
        lam, k = float(g['lam']), int(g['chol_k'])
        A = -g['K'] + lam * np.eye(n)
        Lt, idx = _sharded_pchol(A[row0:row1], g['diag'][row0:row1], row0, n, k)
        assert np.array_equal(idx, g['index_columns'])
        assert np.abs(Lt.T - g['L'][row0:row1]).max() < 1e-12
        # Woodbury factor from the reduced Gram, then the sharded PCG
        W = _allreduce(Lt @ Lt.T) + lam * np.eye(k)
        T_local = np.linalg.solve(np.linalg.cholesky(W), Lt)
        # round-2 protocols: merged candidate + row message with deferred swaps; deferred stopping test
        Lt2, idx2 = _sharded_pchol_merged(A[row0:row1], g['diag'][row0:row1], row0, n, k)
        assert np.array_equal(idx2, idx) and np.array_equal(Lt2, Lt)
        x2, iters2, info2 = _sharded_pcg_deferred(A[row0:row1], g['y'][row0:row1], T_local, lam, row0, n, 1e-4, 5 * n)
        x, iters, info = _sharded_pcg(A[row0:row1], g['y'][row0:row1], T_local, lam, row0, n, 1e-4, 5 * n)
        assert (iters2, info2) == (iters, info) and np.array_equal(x2, x)
        x3, iters3, info3 = _sharded_pcg_deferred(A[row0:row1], g['y'][row0:row1], T_local, lam, row0, n, 1e-30, 7)
        x4, iters4, info4 = _sharded_pcg(A[row0:row1], g['y'][row0:row1], T_local, lam, row0, n, 1e-30, 7)
        assert (iters3, info3) == (7, 1) and iters4 == 7 and info4 == 1 and np.array_equal(x3, x4)
        L_full = g['L']
        T_ref = orc.woodbury_factor(L_full, lam)
        x_ref, it_ref, _, info_ref = orc.pcg(lambda v: A @ v, g['y'], lambda a: orc.woodbury_apply(T_ref, lam, a),
                                             1e-4, 5 * n)
        assert info == 0 and info_ref == 0
        assert abs(iters - it_ref) <= max(1, int(0.05 * it_ref)), (iters, it_ref)
        err = np.linalg.norm(x - x_ref[row0:row1]) / np.linalg.norm(x_ref[row0:row1])
        assert err < 1e-3, err
        q.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, 'FAIL: %s\n%s' % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_partition_covers_rows():
    from mlff_preconditioner_b200.dist import shard_slices
    from mlff_preconditioner_b200.engine import shard_points

    for M, world in [(12, 2), (4000, 8), (10000, 8), (7, 3), (20000, 8)]:
        sl = shard_slices(M, 27, world)
        assert sl[0][0] == 0 and max(s[1] for s in sl) == M * 27
        for (a0, a1), (b0, b1) in zip(sl[:-1], sl[1:]):
            assert a1 == b0 or b0 >= M * 27
        ppr = -(-M // world)
        for r in range(world):
            if r * ppr < M:
                assert shard_points(M, r, world) == (r * ppr, min((r + 1) * ppr, M))
    with pytest.raises(ValueError):
        shard_points(2, 3, 4)
    # a world size that would leave the last rank without points is refused on EVERY rank (a rank-local error would
    # leave the others blocked in their first collective): M = 9 on 4 ranks -> 3 points per rank, rank 3 empty
    for r in range(4):
        with pytest.raises(ValueError):
            shard_points(9, r, 4)


@pytest.mark.timeout(300)
def test_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == 'ok', 'rank %d: %s' % (rank, msg)


def test_symop_plan_covers_every_block_pair_once_and_balances():
    """The symmetric tile plan (csrc/symop.cu, host mirror dist.symop_plan): every matrix entry of the
    lower-or-upper triangle is owned by exactly one rank, and the ranks read (nearly) equal shares."""
    from mlff_preconditioner_b200.dist import symop_entries_read, symop_plan

    for M, world in [(12, 1), (12, 2), (13, 2), (12, 3), (12, 4), (15, 4), (40, 8), (43, 8), (4000, 8), (10, 5), (13, 5)]:
        cover = np.zeros((M, M), dtype=np.int32)     # per point pair (i, j): how many ranks touch it (either side)
        reads = []
        for r in range(world):
            tiles = symop_plan(M, world, r)
            assert tiles[0][4] == 1 and tiles[0][0] == tiles[0][2]      # the diagonal tile comes first
            for (i0, i1, j0, j1, diag) in tiles:
                if diag:
                    blk = np.tril(np.ones((i1 - i0, i1 - i0), dtype=np.int32))
                    cover[i0:i1, j0:j1] += blk
                    cover[j0:j1, i0:i1] += np.tril(blk, -1).T
                else:
                    cover[i0:i1, j0:j1] += 1
                    cover[j0:j1, i0:i1] += 1
            reads.append(symop_entries_read(tiles, 27))
        assert (cover == 1).all(), (M, world)
        if M % world == 0 and M >= 8 * world:
            assert max(reads) <= 1.15 * min(reads), (M, world, reads)
            assert sum(reads) <= 0.56 * (27 * M) ** 2                 # about half of the full matrix


def test_symop_numpy_emulation_matches_full_matvec():
    """Sum over ranks of (tile rows A x_cols  +  tile columns A^T x_rows) == K x, on a golden kernel matrix."""
    from mlff_preconditioner_b200.dist import symop_plan

    g = load_golden('eth_s1_m12')
    K, M, di = g['K'], 12, 27
    x = np.random.default_rng(0).standard_normal(K.shape[0])
    for world in (1, 2, 3, 4, 5):
        y = np.zeros_like(x)
        for r in range(world):
            for (i0, i1, j0, j1, diag) in symop_plan(M, world, r):
                A = K[i0 * di:i1 * di, j0 * di:j1 * di]
                if diag:
                    y[i0 * di:i1 * di] += np.tril(A) @ x[j0 * di:j1 * di] + np.tril(A, -1).T @ x[i0 * di:i1 * di]
                else:
                    y[i0 * di:i1 * di] += A @ x[j0 * di:j1 * di]
                    y[j0 * di:j1 * di] += A.T @ x[i0 * di:i1 * di]
        assert np.linalg.norm(y - K @ x) <= 1e-13 * np.linalg.norm(K @ x), world
