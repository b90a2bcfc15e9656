#!/usr/bin/env python
"""GPU check of the experimental option ``precon_reorth`` (twice-projected complement in the orthonormal-form
preconditioner apply) -- to be run BEFORE it becomes the library default; not collected by pytest.

    python scripts/check_precon_reorth.py                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/check_precon_reorth.py                 # sharded

Checks: (1) the device apply equals the numpy restatement (oracle.orthonormal_apply_reorth) on a 2160-row system;
(2) a full solve with the option converges to the same coefficients as without it, in fewer iterations, and its
iteration count matches the oracle PCG driven by the same apply; (3) under torchrun the sharded solve agrees with
the single-GPU one."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
    from bench import WORKLOADS, make_inputs
    from mlff_preconditioner_b200.engine import Engine
    from mlff_preconditioner_b200.solvers.iterative_solver import Iterative
    from oracle import sgdml_oracle as orc

    WORKLOADS['t80'] = ('ethanol', 80, 1e-6)
    inp = make_inputs('t80')
    n, lam, k = inp['n'], 1e-10, inp['n'] // 10

    # (1) apply vs numpy, single GPU engine on every rank
    eng = Engine(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10, perms=inp['perms'])
    Lt = eng.pchol_build(k, want_times=False)[0]
    L = Lt.t().cpu().numpy()
    Qt, Mk = eng.orthonormal_factor_(Lt, lam)
    r = np.random.default_rng(0).standard_normal(n)
    r_t = torch.as_tensor(r, device=eng.device)
    eng.set_option('precon_reorth', 1)
    z = eng.precon_apply(Qt, lam, 1.0, r_t, Mk=Mk).cpu().numpy()
    z_ref = orc.orthonormal_apply_reorth(Qt.cpu().numpy(), Mk.cpu().numpy(), lam, r)
    err = np.linalg.norm(z - z_ref) / np.linalg.norm(z_ref)
    print('[rank %d] apply vs numpy restatement: rel diff %.2e' % (rank, err), flush=True)
    assert err < 1e-7, err
    eng.close()

    # (2)/(3) solves
    out = {}
    for tag, opts, distributed in (('plain', {}, False), ('reorth', {'precon_reorth': 1}, False),
                                   ('reorth_sharded', {'precon_reorth': 1}, True)):
        if distributed and world == 1:
            continue
        task = dict(inp['task'])
        task.update(kernel_mode='matrix_free', distributed=distributed, _options=opts, solver_tol=1e-6)
        it = Iterative(None, None)
        alphas, iters, resid, rmse, idxs, conv, info = it.solve(task, inp['R_desc'], inp['R_d_desc'], inp['tpl'], inp['y'],
                                                                inp['y_std'], break_percentage=(k + 0.5) / n,
                                                                str_preconditioner='cholesky')
        assert conv, tag
        out[tag] = (alphas, iters)
        it.engine.close()
        if rank == 0:
            print('  %-15s iterations %d' % (tag, iters), flush=True)
    d = np.linalg.norm(out['reorth'][0] - out['plain'][0]) / np.linalg.norm(out['plain'][0])
    assert d < 1e-3, d
    assert out['reorth'][1] < out['plain'][1]
    # oracle PCG with the same apply (CPU, dense K)
    K = orc.assemble_kernel_mat(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10)
    A = -K + lam * np.eye(n)
    Qh, Mh = orc.orthonormal_factor(L, lam)
    _, it_ref, _, info = orc.pcg(lambda v: A @ v, inp['y'], lambda v: orc.orthonormal_apply_reorth(Qh, Mh, lam, v), 1e-6, 5 * n)
    if rank == 0:
        print('  oracle PCG with the same apply: %d iterations' % it_ref, flush=True)
    assert info == 0 and abs(out['reorth'][1] - 1 - it_ref) <= max(1, int(0.05 * it_ref)), (out['reorth'][1], it_ref)
    if 'reorth_sharded' in out:
        d = np.linalg.norm(out['reorth_sharded'][0] - out['reorth'][0]) / np.linalg.norm(out['reorth'][0])
        assert d < 1e-4 and abs(out['reorth_sharded'][1] - out['reorth'][1]) <= max(1, int(0.05 * out['reorth'][1])), d
    print('PRECON_REORTH_CHECK OK rank %d/%d' % (rank, world), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
