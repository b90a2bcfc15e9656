#!/usr/bin/env python
"""Rank sweep of the matrix-free north-star system (BASELINE.json configs[4]: aspirin-size, M = 20 000, n = 1 260 000)
on all GPUs of the node: time to solution against the preconditioner rank k.  The reference's rule of thumb
(k = 46 702) is out of reach while the k x k factorisations are replicated, so k is a free parameter here; the build
costs O(k^2 n), the iteration count falls like ~k^-0.9 (profiles/r01q_sweep_cfg3.jsonl) -- this script finds the minimum.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/cfg5_rank_sweep.py --k 6144 8192 12288 [--tol 1e-4] [--M 20000]

One JSON line per rank k (rank 0 prints)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--k', type=int, nargs='+', default=[6144, 8192, 12288])
    ap.add_argument('--tol', type=float, default=1e-4)
    ap.add_argument('--maxiter', type=int, default=10000)
    ap.add_argument('--M', type=int, default=None, help='training points (default: 20 000, the cfg5 size)')
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if args.M:
        kind, _, tol = bench.WORKLOADS['cfg5']
        bench.WORKLOADS['cfg5'] = (kind, args.M, tol)
    peak = 6549.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', peak))
    except Exception:  # noqa: BLE001
        pass
    for k in args.k:
        try:
            res = bench.north_star_solve('cfg5', 'matrix_free', 'projected', k, args.tol, args.maxiter, world, rank, dev, peak)
        except Exception as exc:  # noqa: BLE001
            res = {'k': k, 'error': '%s: %s' % (type(exc).__name__, exc)}
        torch.cuda.empty_cache()
        if rank == 0:
            print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
