#!/usr/bin/env python
"""Matrix-free operator at cfg5 size (aspirin-size N = 21, M = 20 000, n = 1 260 000): time the per-rank slice of
an 8-rank run on ONE GPU (rows of 1/8 of the points against all M points -- the operator needs no collective of
its own, the search direction is allgathered by the CG loop).  Prints one JSON line.

    python scripts/matvec_free_bench.py [--M 20000] [--world 8] [--kind aspirin] [--reps 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--M', type=int, default=20000)
    ap.add_argument('--world', type=int, default=8)
    ap.add_argument('--kind', default='aspirin')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--pairs-kernel', type=int, default=0, help="library option pairs_kernel (0 auto, 1, 2, 3)")
    args = ap.parse_args()
    from bench import WORKLOADS, make_inputs
    from mlff_preconditioner_b200.engine import Engine

    WORKLOADS['mf'] = (args.kind, args.M, 1e-6)
    inp = make_inputs('mf')
    eng = Engine(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10, perms=inp['perms'], rank=0, world=args.world,
                 init_comm=(lambda e: None))
    eng.set_option('pairs_kernel', args.pairs_kernel)
    v = torch.randn(eng.n, dtype=torch.float64, device=eng.device)
    out = eng.empty(eng.n_local)
    for _ in range(2):
        eng.matvec_free(v, alpha=-1.0, shift=1e-10, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.matvec_free(v, alpha=-1.0, shift=1e-10, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ml = eng.pt1 - eng.pt0
    flops = 8.0 * ml * eng.M * eng.S * eng.D
    ms = float(np.median(ts))
    print(json.dumps({'workload': '%s N=%d M=%d n=%d, rows of rank 0 of %d (%d points)' % (args.kind, eng.N, eng.M, eng.n, args.world, ml),
                      'matvec_ms': ms, 'all_ms': ts, 'algorithmic_flop': flops, 'tflops': flops / ms / 1e9,
                      'fp64_peak_tflops_cublas_dgemm': 35.4, 'frac_of_dgemm_rate': flops / ms / 1e9 / 35.4}))


if __name__ == '__main__':
    main()
