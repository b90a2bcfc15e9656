#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py

Every rank runs the row-block sharded path (pivoted Cholesky -> Woodbury -> assembled and matrix-free PCG,
plus one Nystroem variant); every rank also runs the same problem unsharded on its own GPU and compares:
pivot permutation identical, factor shard == the corresponding rows, coefficients equal to 1e-6, iteration
counts within max(1, 5%).  Prints one 'MULTI_GPU_CHECK OK' line per rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from bench import WORKLOADS, make_inputs
    from mlff_preconditioner_b200.dist import allgather_rows, init_engine_comm
    from mlff_preconditioner_b200.engine import Engine
    from mlff_preconditioner_b200.solvers.iterative_solver import Iterative

    # default: M not divisible by the world size on purpose.  MG_M / MG_K run the same checks at another size
    # (e.g. MG_M=4000 MG_K=4839, the benchmark's system; the explicit-K comparisons are then skipped: 93 GB).
    M_pts, k = int(os.environ.get('MG_M', '203')), int(os.environ.get('MG_K', '400'))
    big = M_pts > 1000
    WORKLOADS['mg'] = ('ethanol', M_pts, 1e-5)
    inp = make_inputs('mg')
    n = inp['n']
    frac = (k + 0.5) / n
    lam = 1e-10
    y = torch.as_tensor(inp['y'], device='cuda')

    eng = Engine(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10, perms=inp['perms'], rank=rank, world=world,
                 init_comm=init_engine_comm)
    ref = Engine(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10, perms=inp['perms'])
    sl = slice(eng.row0, eng.row0 + eng.n_local)

    # kernel pieces on the shard
    assert torch.equal(eng.kernel_diag(), ref.kernel_diag()[sl])
    if not big:
        K_ref = ref.kernel_assemble()
        assert torch.equal(eng.kernel_assemble(), K_ref[sl])

    # pivoted Cholesky + Woodbury
    Lt, idx, _, _ = eng.pchol_build(k)
    Lt_ref, idx_ref, _, _ = ref.pchol_build(k)
    assert torch.equal(idx, idx_ref), 'pivot permutation differs between sharded and single-GPU runs'
    err = float((Lt - Lt_ref[:, sl]).abs().max() / Lt_ref.abs().max())
    if rank == 0:
        print('  pivots identical; max |dL| / max |L| = %.2e' % err, flush=True)
    assert err < (1e-9 if big else 1e-12), err
    T = eng.woodbury_factor_(Lt, lam)
    T_ref = ref.woodbury_factor_(Lt_ref, lam)
    a = torch.randn(n, dtype=torch.float64, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1))
    z = eng.precon_apply(T, lam, 1.0, a[sl].contiguous())
    z_ref = ref.precon_apply(T_ref, lam, 1.0, a)
    err = float((z - z_ref[sl]).norm() / z_ref[sl].norm())
    if rank == 0:
        print('  Woodbury apply sharded vs single: rel diff %.2e' % err, flush=True)
    assert err < (1e-3 if big else 1e-8), err      # the 1/lam = 1e10 amplification grows with ||L L^T||

    # projected form: the cross-rank (hi, lo) fold of the Gram kernels gives every rank the same k x k matrices, and
    # the sharded apply (one k-vector allreduce) equals the single-GPU one
    Lt2, _, _, _ = eng.pchol_build(k)
    Lt2_ref, _, _, _ = ref.pchol_build(k)
    Qt, Mk, E = eng.projected_factor_(Lt2, lam)
    Qt_ref, Mk_ref, E_ref = ref.projected_factor_(Lt2_ref, lam)
    gath = [torch.empty_like(E) for _ in range(world)]
    dist.all_gather(gath, E)
    assert all(torch.equal(g_, gath[0]) for g_ in gath), 'defect matrix differs between ranks'
    # the defect belongs to THIS factor (the sharded Gram sums differ from the single-GPU ones in the last bits, so Qt
    # and with it E differ at the 1e-16 level): gather the sharded Qt and measure its defect on one GPU
    parts = [None] * world
    dist.all_gather_object(parts, Qt.cpu())
    Q_full = torch.cat(parts, dim=1).to('cuda')
    errE = float((E - ref.gram_defect(Q_full)).abs().max())
    errQ = float((Qt - Qt_ref[:, sl]).abs().max())
    del Q_full, parts
    z = eng.precon_apply(Qt, lam, 1.0, a[sl].contiguous(), Mk=Mk, E=E)
    z_ref = ref.precon_apply(Qt_ref, lam, 1.0, a, Mk=Mk_ref, E=E_ref)
    err = float((z - z_ref[sl]).norm() / z_ref[sl].norm())
    if rank == 0:
        print('  projected form sharded vs single: max |dE| %.2e (|E| max %.2e), max |dQ| %.2e, apply rel diff %.2e'
              % (errE, float(E_ref.abs().max()), errQ, err), flush=True)
    # the two defects fold the same products in different 16-column groupings (rank 1's columns do not start on a
    # k-tile boundary): they may differ by the rounding inside the k-tile products, eps * 16/n per rounding
    tolE = 1.1e-16 * (16.0 / n) * 4 * np.sqrt(n)
    assert errE < tolE and err < (1e-3 if big else 1e-8), (errE, tolE, err)
    del Lt2, Lt2_ref, Qt, Mk, E, Qt_ref, Mk_ref, E_ref

    # matvecs
    v = torch.randn(n, dtype=torch.float64, device='cuda', generator=torch.Generator(device='cuda').manual_seed(2))
    mv = eng.matvec_free(v)
    assert float((mv - ref.matvec_free(v)[sl]).norm() / mv.norm()) < 1e-12

    del Lt, Lt_ref, T, T_ref, z, z_ref
    torch.cuda.empty_cache()
    if rank == 0:
        print('  factor checks passed (pivots, L shard, Woodbury apply)', flush=True)

    # symmetric tile operator with the real reduce-scatter
    Ksym = eng.symop_assemble()
    sv = eng.symop_apply(Ksym, v, alpha=-1.0, shift=lam)
    gv = ref.gemv(K_ref, v, alpha=-1.0, shift=lam)[sl] if not big else ref.matvec_free(v, alpha=-1.0, shift=lam)[sl]
    assert float((sv - gv).norm() / gv.norm()) < 1e-11
    del Ksym
    if big:   # the plain sharded GEMV against the matrix-free operator
        K_loc = eng.kernel_assemble()
        gv2 = eng.gemv(K_loc, v, alpha=-1.0, shift=lam, x_off=eng.row0)
        assert float((gv2 - gv).norm() / gv.norm()) < 1e-11
        del K_loc

    # full solves through the public entry point, sharded vs single GPU
    combos = (('assembled', 'cholesky', 'woodbury'), ('assembled_sym', 'cholesky', 'woodbury'),
              ('matrix_free', 'cholesky', 'woodbury'), ('assembled_sym', 'cholesky', 'projected'),
              ('matrix_free', 'cholesky', 'projected'), ('assembled', 'random_scores', 'woodbury'),
              ('matrix_free', 'lev_random', 'woodbury'))
    if big:
        combos = (('matrix_free', 'cholesky', 'projected'), ('assembled_sym', 'cholesky', 'projected'))
        torch.cuda.empty_cache()
    tol_solve = 1e-3 if big else 1e-5
    for mode, variant, form in combos:
        out = {}
        for tag, distributed in (('sharded', True), ('single', False)):
            task = dict(inp['task'])
            task.update(kernel_mode=mode, distributed=distributed, solver_tol=tol_solve, _want_hist=True,
                        precon_form=form)
            # the sharded run does NOT seed the ranks alike: rank 0's column draws are broadcast;
            # the single-GPU comparison run uses rank 0's seed on every rank
            np.random.seed(rank if distributed else 0)
            it = Iterative(None, None)
            alphas, iters, resid, rmse, idxs, conv, info = it.solve(
                task, inp['R_desc'], inp['R_d_desc'], inp['tpl'], inp['y'], inp['y_std'],
                break_percentage=frac, str_preconditioner=variant)
            assert conv
            out[tag] = (alphas, iters, idxs)
            # independent residual check of this solution with the single-GPU matrix-free operator
            xs = torch.as_tensor(-alphas, device='cuda')
            rs = y - ref.matvec_free(xs, alpha=-1.0, shift=lam)
            if rank == 0:
                hist = it.timings.get('resid_hist_rel')
                print('  %s/%s/%s %s: iters %d, ||b - A x|| / ||b|| re-checked on one GPU = %.3e, history every 100: %s'
                      % (mode, variant, form, tag, iters, float(rs.norm() / y.norm()),
                         ' '.join('%.2e' % v for v in (hist[::100] if hist is not None else []))), flush=True)
            it.engine.close()
        d = np.linalg.norm(out['sharded'][0] - out['single'][0]) / np.linalg.norm(out['single'][0])
        i1, i0 = out['sharded'][1], out['single'][1]
        if rank == 0:
            print('  %s/%s/%s: iters sharded %d single %d, |dalpha|/|alpha| = %.2e' % (mode, variant, form, i1, i0, d), flush=True)
        assert d < (1e-2 if big else 1e-4), (mode, variant, d)   # both are tol-accurate solutions (tol 1e-3 when big)
        # Nystroem preconditioners rest on a Cholesky of -K_mm +- 1e-15 I (iterative_solver.py:576-583): ill-conditioned
        # enough that the summation order of an 8-rank Gram moves the count by ~10 % (same columns, same solution)
        band = 0.05 if variant == 'cholesky' else 0.15
        assert abs(i1 - i0) <= max(1, int(band * i0)), (mode, variant, i1, i0)
        assert np.array_equal(out['sharded'][2], out['single'][2])

    # replicated-vector helper
    full = allgather_rows(eng, a[sl].contiguous())
    assert torch.equal(full, a)
    dist.barrier()
    print('MULTI_GPU_CHECK OK rank %d/%d n=%d n_local=%d peer_collectives=%s' % (rank, world, n, eng.n_local,
                                                                                eng.peer_collectives), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
