#!/usr/bin/env python
"""Benchmark of the hot path: fp64 PCG time-to-solution on the assembled sGDML kernel.

One "step" = one full solve of the named workload: pivoted-Cholesky preconditioner build (on-the-fly
columns) + Woodbury factorisation + explicit kernel assembly + PCG to the relative residual `tol`.
Default workload (BASELINE.json configs[1]): synthetic ethanol-size geometries, N = 9 atoms, M = 4000
training points -> n = 108 000 (K = 93.3 GB fp64, assembled in HBM), k = rule-of-thumb rank, tol 1e-6.

  value : seconds per solve with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the reference-facing call Iterative.solve(task, R_desc, R_d_desc, ...) with HOST
          numpy buffers -- H2D of descriptors / labels and D2H of the coefficients inside the timed region
  roofline : the assembled GEMV (dominant kernel), algorithmic bytes 8*n_local*n + 8*n + 8*n_local per launch
             over its mean CUDA-event duration inside the timed solves
  cpu_baseline : the reference's CPU algorithm (oracle port, torch-CPU matvec = the reference's fastest CPU
             route) timed on this box's host cores on a bounded sample and extrapolated to the full solve

`--impl reference` prints the CPU arm alone (rank 0 only under torchrun).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (structure, M, tol)   -- BASELINE.json configs[0..4]; cfg2 is the single-GPU headline
    'cfg1': ('nanotube', 9, 1e-6),        # N = 370, n = 9 990 (the reference's own CPU-runnable case)
    'cfg2': ('ethanol', 4000, 1e-6),      # N = 9, n = 108 000, assembled on one B200
    'cfg4': ('ethanol', 10000, 1e-6),     # n = 270 000, sharded over >= 4 B200 (symmetric tile storage)
    'cfg5': ('aspirin', 20000, 1e-6),     # N = 21, n = 1 260 000, matrix-free over 8 B200
    'cfg2_mid': ('ethanol', 1500, 1e-6),
    'small': ('ethanol', 300, 1e-6),
}
# rule-of-thumb parameters (m, k_min) per molecule (reference src/tools/plot_data.py:677-706)
RULE_OF_THUMB = {'ethanol': (0.87, 10), 'aspirin': (1.14, 236), 'nanotube': (0.73, 89)}
CONSTANTS_FILE = os.path.join(ROOT, 'bench_constants.json')


def rule_of_thumb(n, k_min, m):
    """k = (k_min^m * m * n^2 / 2)^(1/(2+m))  (reference src/tools/plot_data.py:1254-1258)."""
    return int((k_min ** m * m * n ** 2 / 2) ** (1.0 / (2 + m)))


def make_inputs(workload, M_override=None, tol_override=None, k_override=None):
    from mlff_preconditioner_b200 import synthetic
    from mlff_preconditioner_b200.desc import Desc, tril_perms_lin_from_perms

    kind, M, tol = WORKLOADS[workload]
    if M_override:
        M = M_override
    ds = synthetic.make_dataset(kind, M, seed=0)
    N = ds['R'].shape[1]
    perms = np.arange(N)[None]
    desc = Desc(N)
    tpl = tril_perms_lin_from_perms(perms, desc)
    R_desc, R_d_desc = desc.from_R(ds['R'].reshape(M, -1))
    y = ds['F'].ravel().copy()
    y_std = np.std(y)
    y /= y_std
    n = 3 * N * M
    m_rt, kmin_rt = RULE_OF_THUMB.get(kind, RULE_OF_THUMB['ethanol'])
    k = min(rule_of_thumb(n, kmin_rt, m_rt), n // 4)
    if k_override:
        k = k_override
    if tol_override:
        tol = tol_override
    task = {'R_train': ds['R'], 'F_train': ds['F'], 'sig': 10, 'lam': 1e-10, 'perms': perms, 'use_E_cstr': False,
            'solver_tol': tol, 'n_inducing_pts_init': 25, 'truncated_cholesky': 1500}
    return dict(kind=kind, M=M, N=N, n=n, k=k, tol=tol, task=task, R_desc=R_desc, R_d_desc=R_d_desc, tpl=tpl,
                y=y, y_std=y_std, perms=perms)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    def __init__(self, index):
        self.index, self.samples, self.proc, self.t_mark = index, [], None, 0.0

    def start(self):
        # one sample per second is enough to see a throttle reason and keeps the driver queries away from the
        # host-synchronous parts of a step (the pivot loop reads 64 bytes of state every 8 steps)
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,' \
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
            'clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '1000'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def mark(self):
        """Start of the timed region: only samples taken after this call are reported.  The sampler itself is started
        before the warm-up steps, because nvidia-smi's start-up (NVML initialisation on an 8-GPU node) disturbs
        host-synchronous work for about half a second -- it used to land in the second timed step."""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, s in self.samples:
            if ts < self.t_mark:
                continue
            f = [x.strip() for x in s.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        busy = [c for c in sm if c > 0]
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def _cpu_threads():
    """All host threads for torch, numpy's BLAS and LAPACK -- explicitly, because torchrun exports OMP_NUM_THREADS=1."""
    import torch

    ncores = os.cpu_count() or 1
    try:
        torch.set_num_threads(ncores)
    except RuntimeError:
        pass
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=ncores)
    except Exception:
        pass
    return ncores


def cpu_reference_sample(inp, cg_iters, light=False):
    """Bounded sample of the reference's CPU algorithm on this box, extrapolated to one full solve.

    Per-operation costs are measured with the oracle port of the reference (oracle/sgdml_oracle.py); the kernel matvec
    runs the reference's torch-CPU op sequence from the Cartesian geometries (torchtools.py:172-272, one 32 GB batch)
    and costs what the reference's own K_op.matvec costs on the same cores (tests/test_cpu_arm_cost.py):
      t_mv   one K_op.matvec at full n                       (also the cost of one pivot column: get_col = K_op e_i)
      t_sch  one pivot step's Schur update at m = k/2         (incomplete_cholesky.py:66-78)
      t_fac  Woodbury factorisation at rank k' << k, scaled by (k/k')^2   (iterative_cholesky.py:141-143)
      t_app  one preconditioner apply at rank k', scaled by k/k'            (iterative_cholesky.py:145-148)
    solve = k (t_mv + t_sch) + t_fac + cg_iters (t_mv + t_app)
    """
    import torch
    from oracle import sgdml_oracle as orc

    _cpu_threads()
    n, k, M = inp['n'], inp['k'], inp['M']
    rng = np.random.default_rng(0)
    D = inp['R_desc'].shape[1]
    Rs_t = torch.from_numpy(np.ascontiguousarray(inp['task']['R_train'], dtype=np.float64))
    Xp_t = torch.from_numpy(np.ascontiguousarray(orc.permuted_rows(inp['R_desc'], inp['tpl']).reshape(-1, D)))
    v = rng.standard_normal(n)
    reps_mv = 1 if light else 3
    orc.kernel_matvec_torch_cpu(Rs_t, Xp_t, inp['R_d_desc'], inp['tpl'], 10, v)  # warm (thread pools, page faults)
    ts = []
    for _ in range(reps_mv):
        t0 = time.perf_counter()
        orc.kernel_matvec_torch_cpu(Rs_t, Xp_t, inp['R_d_desc'], inp['tpl'], 10, v)
        ts.append(time.perf_counter() - t0)
    t_mv = float(np.median(ts))

    m_half = max(1, k // 2 if not light else k // 8)
    L = np.ones((n, m_half))
    idx = np.arange(n)
    diag = np.ones(n)
    col = np.ones(n)
    t0 = time.perf_counter()
    for _ in range(1 if light else 2):
        i_pi = idx[1:]
        schur = np.einsum('c,rc->r', L[0, :m_half], L[i_pi, :m_half])
        newcol = (col[i_pi] - schur) / 2.0
        diag[i_pi] -= newcol ** 2
    t_sch = (time.perf_counter() - t0) / (1 if light else 2) * (k / 2.0) / m_half
    del L

    k_s = min(k, 256 if light else 512)
    Ls = rng.standard_normal((n, k_s))
    t0 = time.perf_counter()
    T = orc.woodbury_factor(Ls, 1e-10)
    t_fac = (time.perf_counter() - t0) * (k / k_s) ** 2
    a = rng.standard_normal(n)
    t0 = time.perf_counter()
    for _ in range(3):
        orc.woodbury_apply(T, 1e-10, a)
    t_app = (time.perf_counter() - t0) / 3 * (k / k_s)
    total = k * (t_mv + t_sch) + t_fac + cg_iters * (t_mv + t_app)
    detail = {'t_matvec_s': t_mv, 't_schur_step_s': t_sch, 't_factor_s': t_fac, 't_apply_s': t_app,
              'k': k, 'cg_iters': cg_iters}
    return total, detail


CPU_SAMPLE_TEXT = ("3 full-n kernel matvecs in the reference's torch-CPU op sequence (median), 2 Schur-update steps at m=k/2, "
                   "Woodbury factor at k'=512 (scaled (k/k')^2), 3 applies at k' (scaled k/k'); "
                   'extrapolated solve = k(t_mv+t_sch)+t_fac+iters(t_mv+t_app)')


def load_constants():
    try:
        return json.load(open(CONSTANTS_FILE))
    except Exception:
        return {}


def reference_iteration_count(workload, fallback):
    """CG iterations of the reference's own formula on this workload and where the number comes from: measured with the
    unmodified reference (cfg1), with the oracle port on the CPU (cfg2, when the run is recorded), else the device's
    count with precon_form='woodbury' (same formula, same pivots) -- never the projected form's smaller count."""
    c = load_constants().get(workload, {})
    for key, label in (('cg_iters_reference_measured', 'unmodified reference, CPU'),
                       ('cg_iters_port_measured', 'oracle port of the reference formula, CPU'),
                       ('cg_iters_gpu_woodbury_form', "device run with the reference's formula (precon_form=woodbury)")):
        if key in c:
            return int(c[key]), '%s (%s)' % (label, c.get(key + '_source', 'bench_constants.json'))
    return int(fallback), 'assumed (no recorded run of the reference formula for this workload)'


def measured_cpu_points(workload):
    """Fully measured (not extrapolated) CPU solves recorded for this workload, for context next to the sample."""
    return load_constants().get(workload, {}).get('measured_full_solves', [])


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    inp = make_inputs(args.workload, args.M, args.tol, args.k)
    cg_iters, src = reference_iteration_count(args.workload, 1000)
    ncores = _cpu_threads()
    for _ in range(args.warmup):
        cpu_reference_sample(inp, cg_iters, light=True)
    vals, detail = [], None
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        v, detail = cpu_reference_sample(inp, cg_iters)
        vals.append(v)
    wall = time.perf_counter() - t_wall
    value = float(np.mean(vals))
    sample = 'per step: %s; cg_iters=%d from %s' % (CPU_SAMPLE_TEXT, cg_iters, src)
    line = {
        'impl': 'reference', 'metric': 'pcg_time_to_solution', 'value': value, 'unit': 's', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / max(args.steps, 1),
        'higher_is_better': False, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(inp, args, max(args.gpus, 1)),
        'cpu_baseline': {'value': value, 'unit': 's', 'cores': ncores, 'kind': 'port', 'sample': sample,
                         'extrapolated': True, 'detail': detail,
                         'measured_full_solves': measured_cpu_points(args.workload)},
        'e2e': {'value': value, 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


PRECON_FORM_TEXT = {
    'projected': "projected: orthonormal basis of range(L) + k x k inverse + first-order exact complement projector "
                 "(E = Qt Qt^T - I from the extended-precision DMMA Gram), two passes over the factor per apply; same "
                 "operator as the reference's Woodbury formula, no 1/lam cancellation",
    'woodbury': "woodbury: the reference's formula (iterative_cholesky.py:141-148), Gram summed in extended precision",
    'orthonormal': 'orthonormal: projected form without the defect correction (round-1 form)',
    'reorth': 'orthonormal + complement projected twice (four passes over the factor; option precon_reorth=1)',
}


def workload_config(inp, args, world):
    storage = {'assembled': 'assembled fp64 kernel (%.1f GB) row-block sharded' % (8.0 * inp['n'] ** 2 / 1e9),
               'assembled_sym': 'assembled fp64 kernel in symmetric tile storage (%.1f GB, every entry of the lower '
                                'block triangle stored and read once) sharded' % (4.0 * inp['n'] ** 2 / 1e9),
               'matrix_free': 'matrix-free kernel operator sharded'}[args.mode]
    return {'workload': "%s: synthetic %s-size N=%d M=%d n=%d, %s over %d GPU(s), 'cholesky' (pivoted partial Cholesky) "
                        'preconditioner k=%d, tol=%g, sig=10, lam=1e-10'
                        % (args.workload, inp['kind'], inp['N'], inp['M'], inp['n'], storage, world, inp['k'], inp['tol']),
            'kernel_mode': args.mode,
            'precon_form': PRECON_FORM_TEXT[args.precon_form],
            'n': inp['n'], 'k': inp['k'],
            'tol': inp['tol'],
            'l2_policy': 'inputs larger than L2 (126 MB): every CG iteration streams this rank\'s %.1f GB of K and %.1f GB '
                         'of the preconditioner factor once; K is re-assembled every step'
                         % ({'assembled': 8.0, 'assembled_sym': 4.0, 'matrix_free': 0.0}[args.mode] * inp['n'] ** 2 / world / 1e9,
                            16.0 * inp['k'] * inp['n'] / world / 1e9)}


# ----------------------------------------------------------------------------- north-star systems (8 GPUs)
FP64_DGEMM_TFLOPS_MEASURED = 35.5   # cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/r01t_fp64_peak_16w.json)


def north_star_solve(workload, mode, form, k, tol, maxiter, world, rank, dev, hbm_peak, timed_steps=1, warm_steps=0):
    """One of BASELINE.json's multi-GPU configurations end to end on the device (inputs resident in HBM): preconditioner
    build, operator set-up, PCG.  Returns a dict for the JSON line (rank 0 prints it)."""
    import torch
    import torch.distributed as dist

    from mlff_preconditioner_b200.dist import init_engine_comm, symop_entries_read, symop_plan
    from mlff_preconditioner_b200.engine import Engine
    from mlff_preconditioner_b200.solvers.iterative_solver import Iterative

    inp = make_inputs(workload, tol_override=tol, k_override=k)
    n, k = inp['n'], inp['k']
    frac = (k + 0.5) / n
    task = dict(inp['task'])
    task.update(kernel_mode=mode, precon_form=form, _want_hist=True, _maxiter=int(maxiter), solver_tol=tol)
    eng = Engine(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10, perms=inp['perms'], rank=rank, world=world,
                 init_comm=init_engine_comm if world > 1 else None)
    y_t = torch.as_tensor(inp['y'], device=dev)
    if mode == 'assembled_sym':
        task['_K_buffer'] = eng.empty(eng.symop_storage_elems())
    n_ind = min(inp['M'], int(max(np.ceil(frac * inp['M']), 1)))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm_steps):
        Iterative(None, None).solve_device(task, eng, y_t, frac, 'cholesky', n_ind)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(timed_steps):
        it = Iterative(None, None)
        out = it.solve_device(task, eng, y_t, frac, 'cholesky', n_ind)
    e1.record()
    sync_all()
    secs = e0.elapsed_time(e1) * 1e-3 / timed_steps
    if world > 1:
        t = torch.tensor([secs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    tm = it.timings
    st = tm['pcg_stats']
    op_s = st['op_ms'] * 1e-3 / max(st['op_calls'], 1)
    res = {'workload': workload, 'n': n, 'k': k, 'tol': tol, 'kernel_mode': mode, 'precon_form': form, 'n_gpus': world,
           'value': secs, 'unit': 's', 'steps': timed_steps, 'cg_iters': int(out[1]), 'converged': out[3] == 0,
           'rel_resid': float(out[2] / np.linalg.norm(inp['y'])),
           'phases': {'preconditioner_s': tm['preconditioner'], 'pchol_build_s': tm.get('pchol_build'),
                      'assemble_s': tm['assemble'], 'cg_s': tm['cg'], 'operator_avg_ms': op_s * 1e3,
                      'precon_apply_avg_ms': st['precon_ms'] / max(st['op_calls'], 1),
                      'rel_resid_every_200_iters': [float('%.3g' % v) for v in tm['resid_hist_rel'][::200]]}}
    if mode == 'assembled_sym':
        entries = symop_entries_read(symop_plan(eng.M, world, rank), eng.dim_i)
        gbs = (8.0 * entries + 8.0 * eng.n + 8.0 * eng.n_local) / op_s / 1e9
        res['roofline'] = {'bound': 'hbm', 'kernel': 'symv_tma_kernel (+ reduce, + reduce-scatter of the partial products)',
                           'achieved_per_gpu': gbs, 'peak_per_gpu': hbm_peak, 'frac': gbs / hbm_peak,
                           'achieved_aggregate': gbs * world, 'peak_aggregate': hbm_peak * world, 'unit': 'GB/s',
                           'frac_of_nominal_8TBs': gbs / 8000.0,
                           'equivalent_full_gemv_aggregate': 8.0 * eng.n * eng.n / op_s / 1e9,
                           'note': 'rank 0; symmetric tile storage, every stored entry read once per matvec'}
    else:
        flops = 8.0 * (eng.pt1 - eng.pt0) * eng.M * eng.S * eng.D
        tf = flops / op_s / 1e12
        res['roofline'] = {'bound': 'fp64', 'kernel': 'matvec_free (pairs + DMMA GEMM)', 'achieved_per_gpu': tf,
                           'peak_per_gpu': FP64_DGEMM_TFLOPS_MEASURED, 'frac': tf / FP64_DGEMM_TFLOPS_MEASURED,
                           'achieved_aggregate': tf * world, 'unit': 'TFLOP/s',
                           'peak_source': 'cuBLAS DGEMM measured on this pool (profiles/r01t_fp64_peak_16w.json)',
                           'apply_share_of_iteration': st['precon_ms'] / max(st['precon_ms'] + st['op_ms'], 1e-9)}
    eng.close()
    del eng, it, out, task
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    # rank 0 prints exactly ONE line on stdout: anything native libraries write to fd 1 (NCCL banners) goes to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')   # both levels print a version banner on stdout
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import __graft_entry__ as g

    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from mlff_preconditioner_b200 import _lib
    from mlff_preconditioner_b200.dist import allgather_rows, init_engine_comm
    from mlff_preconditioner_b200.engine import Engine
    from mlff_preconditioner_b200.solvers.iterative_solver import Iterative

    lib = _lib.load()
    inp = make_inputs(args.workload, args.M, args.tol, args.k)
    n, k = inp['n'], inp['k']
    frac = (k + 0.5) / n  # int(frac * n) == k
    task = dict(inp['task'])
    task['kernel_mode'] = args.mode
    task['_want_hist'] = True
    task['precon_form'] = 'orthonormal' if args.precon_form == 'reorth' else args.precon_form
    task['_options'] = {'precon_reorth': 1} if args.precon_form == 'reorth' else {}
    task['_options'].update({kv.split('=')[0]: int(kv.split('=')[1]) for kv in args.opt})
    task['_maxiter'] = 100000   # safety net for the benchmark only (the solver's own limit is 5 n)
    dev = torch.device('cuda', local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: inputs in HBM before the timed region --------------------------------
    eng = Engine(inp['R_desc'], inp['R_d_desc'], inp['tpl'], 10, perms=inp['perms'], rank=rank, world=world,
                 init_comm=init_engine_comm if world > 1 else None)
    y_t = torch.as_tensor(inp['y'], device=dev)
    if args.mode == 'assembled':
        task['_K_buffer'] = eng.empty(eng.n_local, eng.n)
    elif args.mode == 'assembled_sym':
        task['_K_buffer'] = eng.empty(eng.symop_storage_elems())
    n_ind = min(inp['M'], int(max(np.ceil(frac * inp['M']), 1)))
    stats = []

    def one_step():
        it = Iterative(None, None)
        out = it.solve_device(task, eng, y_t, frac, 'cholesky', n_ind)
        return it, out

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler.mark()
    launches0 = lib.mlffpc_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev0.record()
    step_ev[0].record()
    for i in range(args.steps):
        it, out = one_step()
        stats.append((it.timings, out[1], out[2], out[3]))
        step_ev[i + 1].record()
    ev1.record()
    barrier()
    step_s = [step_ev[i].elapsed_time(step_ev[i + 1]) * 1e-3 for i in range(args.steps)]
    launches = lib.mlffpc_launch_count() - launches0
    clocks = sampler.stop()
    dev_s = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    value = dev_s / args.steps
    tm, iters, resid, info = stats[-1]
    x_last = out[0]
    alphas_dev = (-allgather_rows(eng, x_last)).cpu().numpy()

    # roofline of the dominant kernel: the operator inside the timed PCG loops
    op_ms = sum(s[0]['pcg_stats']['op_ms'] for s in stats)
    op_calls = sum(s[0]['pcg_stats']['op_calls'] for s in stats)
    pre_ms = sum(s[0]['pcg_stats']['precon_ms'] for s in stats)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    avg_op_s = op_ms * 1e-3 / max(op_calls, 1)
    if args.mode == 'assembled':
        alg_bytes = 8.0 * eng.n_local * eng.n + 8.0 * eng.n + 8.0 * eng.n_local
        achieved = alg_bytes / avg_op_s / 1e9
        roofline = {'bound': 'hbm', 'kernel': 'gemv_rows_kernel', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                    'bytes_per_launch': alg_bytes, 'avg_launch_ms': avg_op_s * 1e3, 'launches_timed': op_calls}
    elif args.mode == 'assembled_sym':
        # symmetric tile storage: the operator reads each stored entry once (8 B per entry x the
        # entries one matvec processes = this rank's tiles, diagonal tile by 32-row strips) + x, y and the
        # per-strip column partials (8 B written + 8 B read per 32 entries)
        from mlff_preconditioner_b200.dist import symop_entries_read, symop_plan
        entries = symop_entries_read(symop_plan(eng.M, world, rank), eng.dim_i)
        alg_bytes = 8.0 * entries + 8.0 * eng.n + 8.0 * eng.n_local
        achieved = alg_bytes / avg_op_s / 1e9
        roofline = {'bound': 'hbm', 'kernel': 'symv_tma_kernel (+ symv_tma_reduce_kernel)', 'achieved': achieved,
                    'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                    'bytes_per_launch': alg_bytes, 'avg_launch_ms': avg_op_s * 1e3, 'launches_timed': op_calls,
                    'note': 'symmetric storage: every entry of the lower block triangle is read once and used for '
                            'rows and columns; a full-matrix GEMV would move %.1f GB, i.e. the equivalent full-GEMV '
                            'rate is %.0f GB/s' % ((8.0 * eng.n_local * eng.n) / 1e9,
                                                   8.0 * eng.n_local * eng.n / avg_op_s / 1e9)}
    else:
        flops = 8.0 * (eng.pt1 - eng.pt0) * eng.M * eng.S * eng.D
        tf = flops / avg_op_s / 1e12
        roofline = {'bound': 'fp64', 'kernel': 'matvec_free (pairs + DMMA GEMM)', 'achieved': tf,
                    'peak': FP64_DGEMM_TFLOPS_MEASURED, 'unit': 'TFLOP/s', 'frac': tf / FP64_DGEMM_TFLOPS_MEASURED,
                    'traffic': None,
                    'peak_source': 'cuBLAS DGEMM 8192^3 measured on this pool (profiles/r01t_fp64_peak_16w.json); '
                                   'MEASURED_PEAKS.json has no fp64 entry, nominal B200 fp64 is 40 TFLOP/s',
                    'avg_launch_ms': avg_op_s * 1e3, 'launches_timed': op_calls}
    # measured DRAM traffic of the dominant kernel (ncu --set full), only valid for the configuration it was captured on
    if args.workload == 'cfg2' and world == 1 and not args.M:
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'))).get(args.mode)
            if tr:
                roofline['traffic'] = tr['bytes']
                roofline['traffic_source'] = tr['source']
        except Exception:
            pass
    line_collectives = ('none (one GPU)' if world == 1 else
                        ('NVLink peer memory, fused into the solver kernels (CUDA IPC; NCCL only for set-up, the Gram '
                         'reduction and the panel rebuilds)' if tm.get('peer_collectives') else 'NCCL'))
    phases = {'preconditioner_s': tm['preconditioner'], 'pchol_build_s': tm.get('pchol_build'),
              'assemble_s': tm['assemble'], 'cg_s': tm['cg'], 'cg_iters': iters, 'resid': resid,
              'rel_resid': resid / np.linalg.norm(inp['y']), 'converged': info == 0,
              'precon_apply_avg_ms': pre_ms / max(op_calls, 1),
              'rel_resid_every_100_iters': [float('%.3g' % v) for v in tm['resid_hist_rel'][::100]],
              'per_step': [{'device_s': float('%.4f' % step_s[i]), 'preconditioner_s': float('%.4f' % st[0]['preconditioner']),
                            'pchol_build_s': float('%.4f' % (st[0].get('pchol_build') or 0.0)),
                            'assemble_s': float('%.4f' % st[0]['assemble']), 'cg_s': float('%.4f' % st[0]['cg']),
                            'cg_iters': int(st[1])} for i, st in enumerate(stats)]}

    # ---- the same system through the matrix-free operator (one solve, reported next to the headline) ------
    del it, out
    task.pop('_K_buffer', None)
    torch.cuda.empty_cache()
    alt = None
    if args.mode != 'matrix_free' and not args.no_alt:
        task_mf = dict(task)
        task_mf['kernel_mode'] = 'matrix_free'
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        it_mf = Iterative(None, None)
        out_mf = it_mf.solve_device(task_mf, eng, y_t, frac, 'cholesky', n_ind)
        a1.record()
        barrier()
        alt = {'kernel_mode': 'matrix_free', 'value': max_over_ranks(a0.elapsed_time(a1) * 1e-3), 'unit': 's',
               'cg_iters': int(out_mf[1]), 'converged': out_mf[3] == 0,
               'matvec_avg_ms': it_mf.timings['pcg_stats']['op_ms'] / max(it_mf.timings['pcg_stats']['op_calls'], 1),
               'note': 'same solve with the on-the-fly operator instead of the assembled kernel (one run, not the headline)'}
        del it_mf, out_mf

    # ---- end-to-end arm: host numpy in, host numpy out, through Iterative.solve -----------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h2d = d2h = 0
    res = None
    for rep in range(0 if args.no_e2e else 1 + e2e_steps):  # one warm-up
        if rep == 1:
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        its = Iterative(None, None)
        res = its.solve(task, inp['R_desc'], inp['R_d_desc'], inp['tpl'], inp['y'], inp['y_std'],
                        break_percentage=frac, str_preconditioner='cholesky')
        h2d, d2h = its.timings['h2d_bytes'], its.timings['d2h_bytes']
        its.engine.close()
        del its
    if res is not None:
        e1.record()
        barrier()
        e2e_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3) / e2e_steps
        agree = float(np.linalg.norm(res[0] - alphas_dev) / np.linalg.norm(alphas_dev))
    else:
        e2e_s = agree = None

    line = {
        'metric': 'pcg_time_to_solution', 'value': value, 'unit': 's', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': value * 1e3, 'higher_is_better': False, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(inp, args, world),
        'clocks': clocks, 'collectives': line_collectives,
        'e2e': {'value': e2e_s, 'unit': 's', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'steps': e2e_steps, 'api': 'Iterative.solve(task, R_desc, R_d_desc, tril_perms_lin, y, y_std, ...)',
                'alphas_rel_diff_vs_device_arm': agree},
        'gpu_launches': int(launches),
        'roofline': roofline,
        'phases': phases,
    }
    if alt is not None:
        line['alt'] = alt

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_iters, ref_src = reference_iteration_count(args.workload, iters)
        total, detail = cpu_reference_sample(inp, ref_iters)
        line['cpu_baseline'] = {
            'value': total, 'unit': 's', 'cores': os.cpu_count(), 'kind': 'port', 'extrapolated': True,
            'sample': '%s; cg_iters=%d from %s' % (CPU_SAMPLE_TEXT, ref_iters, ref_src),
            'detail': detail, 'measured_full_solves': measured_cpu_points(args.workload)}
    # ---- BASELINE.json's multi-GPU systems in front of the driver: cfg4 (n = 270 000 assembled) and cfg5 (n = 1.26 M
    # matrix-free), one end-to-end solve each; `value` above stays on cfg2 so the strong-scaling curve is comparable
    want_ns = args.north_star == 'on' or (args.north_star == 'auto' and world == 8 and args.workload == 'cfg2')
    if want_ns:
        eng.close()
        del eng
        torch.cuda.empty_cache()
        ns = {}
        try:
            ns['cfg4'] = north_star_solve('cfg4', 'assembled_sym', args.precon_form if args.precon_form != 'reorth' else 'projected',
                                          None, 1e-6, 100000, world, rank, dev, peak, timed_steps=1, warm_steps=1)
        except Exception as exc:  # noqa: BLE001  (every rank fails alike: sizes are deterministic)
            ns['cfg4'] = {'error': '%s: %s' % (type(exc).__name__, exc)}
        try:
            ns['cfg5'] = north_star_solve('cfg5', 'matrix_free', 'projected', args.ns_cfg5_k, args.ns_cfg5_tol,
                                          args.ns_cfg5_maxiter, world, rank, dev, peak)
            ns['cfg5']['note'] = ('k = %d (%.2f %% of n) instead of the rule-of-thumb 46 702: at that rank each GPU would hold '
                                  '58.8 GB of factor and the replicated k^3 factorisation would dominate the solve; k is the '
                                  'minimum of the measured time-to-solution sweep on 8 GPUs (6144: 13.7 s, 8192: 13.4 s, 12288: '
                                  "17.2 s, 16384: 24.4 s; profiles/r02n_cfg5_sweep.jsonl); tol = %g is the reference's own "
                                  'solver_tol for its n = 5e5 runs (create_data.py:88-97)'
                                  % (args.ns_cfg5_k, 100.0 * args.ns_cfg5_k / 1260000, args.ns_cfg5_tol))
        except Exception as exc:  # noqa: BLE001
            ns['cfg5'] = {'error': '%s: %s' % (type(exc).__name__, exc)}
        line['north_star'] = ns
    if rank == 0:
        real_stdout.write(json.dumps(line) + '\n')
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--M', type=int, default=None, help='override the number of training points')
    ap.add_argument('--mode', default='assembled_sym', choices=['assembled', 'assembled_sym', 'matrix_free'],
                    help='assembled_sym (default): symmetric tile storage, every entry read once per matvec; '
                         'assembled: full row block + plain GEMV; matrix_free: on-the-fly operator')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer arm (profiling runs)')
    ap.add_argument('--no-alt', action='store_true', help='skip the extra matrix-free solve reported under "alt"')
    ap.add_argument('--tol', type=float, default=None, help='override the relative residual target (profiling runs)')
    ap.add_argument('--k', type=int, default=None, help='override the preconditioner rank')
    ap.add_argument('--precon-form', default='projected', choices=['projected', 'woodbury', 'orthonormal', 'reorth'],
                    help="evaluation of the pivoted-Cholesky preconditioner (L L^T + lam I)^-1: 'projected' (benchmark "
                         "default), 'woodbury' (the reference's formula, the library default), 'orthonormal', 'reorth'")
    ap.add_argument('--north-star', default='auto', choices=['auto', 'on', 'off'],
                    help="add one end-to-end solve of cfg4 (n = 270 000, assembled) and cfg5 (n = 1.26 M, matrix-free) to the "
                         "JSON line; 'auto' = when running cfg2 on 8 GPUs")
    ap.add_argument('--ns-cfg5-k', type=int, default=8192,
                    help='preconditioner rank of the cfg5 north-star solve (minimum of the measured sweep, profiles/r02n_cfg5_sweep.jsonl)')
    ap.add_argument('--ns-cfg5-tol', type=float, default=1e-4)
    ap.add_argument('--ns-cfg5-maxiter', type=int, default=10000)
    ap.add_argument('--opt', action='append', default=[], help='library option name=int (mlffpc_set_option), repeatable')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
