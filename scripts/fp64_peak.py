#!/usr/bin/env python
"""Measure the fp64 GEMM rate of this B200 two ways: cuBLAS DGEMM through torch.matmul (the practical
"measured fp64 peak", MEASURED_PEAKS.json has none) and libmlffpc's DMMA kernel (mlffpc_dgemm) on the
same shapes.  Prints one JSON line; CUDA events, best of 5 after 2 warm-ups."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def best_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    import numpy as np
    from bench import make_inputs
    from mlff_preconditioner_b200.engine import Engine

    inp = make_inputs('small')
    eng = Engine(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10, perms=inp['perms'])
    out = {}
    for (m, n, k, tb) in [(8192, 8192, 8192, False), (4864, 4864, 108000, True), (108000, 4864, 64, False)]:
        A = torch.randn(m, k, dtype=torch.float64, device='cuda')
        B = torch.randn((n, k) if tb else (k, n), dtype=torch.float64, device='cuda')
        C = torch.empty(m, n, dtype=torch.float64, device='cuda')
        fl = 2.0 * m * n * k
        t_cublas = best_ms(lambda: torch.matmul(A, B.t() if tb else B, out=C))
        t_ours = best_ms(lambda: eng.dgemm(A, B, trans_b=tb, out=C))
        out['%dx%dx%d%s' % (m, n, k, '_nt' if tb else '_nn')] = {
            'cublas_tflops': fl / t_cublas / 1e9, 'mlffpc_tflops': fl / t_ours / 1e9,
            'cublas_ms': t_cublas, 'mlffpc_ms': t_ours}
        del A, B, C
    print(json.dumps({'fp64_gemm': out}))


if __name__ == '__main__':
    main()
