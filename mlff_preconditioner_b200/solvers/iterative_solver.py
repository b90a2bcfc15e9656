"""The solve step on the GPU: a drop-in for the reference's ``Iterative``
(``/root/reference/src/sGDML/sgdml/solvers/iterative_solver.py:75-1108``).

Same constructor, same ``solve`` signature, same 7-tuple result and ``info`` keys, same preconditioner
strings -- but every O(n) or larger operation runs in libmlffpc.so on the device: kernel diagonal,
columns and assembly from descriptors, pivoted partial Cholesky, Woodbury / Nystroem factorisation,
preconditioner apply, kernel matvec (assembled GEMV or matrix-free) and the CG recurrences.  The host
keeps what the reference also does on the host with numpy's global RNG (column draws), so
``np.random.seed(s)`` before ``solve`` selects the same columns on both sides.
"""
import timeit

import numpy as np
import torch

from .. import DONE, NOT_DONE  # noqa: F401  (callback protocol constants)
from ..dist import allgather_rows, dist_info, init_engine_comm
from ..engine import Engine
from .operators import KernelOperator, LowRankPreconditioner

NYSTROM_KEYS = ['lev_scores', 'random_scores', 'inverse_lev', 'lev_random', 'truncated_cholesky',
                'truncated_cholesky_custom']
OUT_OF_CONTRACT = ['rank_k_lev_scores', 'rank_k_lev_scores_custom', 'eigvec_precon',
                   'eigvec_precon_block_diagonal', 'eigvec_precon_atomic_interactions']


class Iterative(object):
    def __init__(self, gdml_train, desc, callback=None, max_processes=None, use_torch=False):
        self.gdml_train = gdml_train
        self.desc = desc
        self.callback = callback
        self._max_processes = max_processes
        self._use_torch = use_torch
        self.engine = None
        self.K_local = None   # explicit local rows when the operator is assembled
        self.last_P_op = None
        self.timings = {}

    # ------------------------------------------------------------------ device pieces
    def _make_engine(self, task, R_desc, R_d_desc, tril_perms_lin):
        rank, world, _ = dist_info()
        if not task.get('distributed', True):
            rank, world = 0, 1
        self.engine = Engine(R_desc, R_d_desc, tril_perms_lin, task['sig'], perms=task.get('perms'),
                             rank=rank, world=world, init_comm=init_engine_comm if world > 1 else None)
        return self.engine

    def _sync_draw(self, idxs):
        """Column draws come from numpy's global RNG like in the reference (iterative_solver.py:685, :709, :747, :472).
        In a sharded run every rank draws (so the RNG streams advance alike) but rank 0's result is the one all ranks
        use -- nothing guarantees that the ranks were seeded identically."""
        eng = self.engine
        if eng.world == 1:
            return idxs
        import torch.distributed as dist

        dev = eng.device if dist.get_backend() == 'nccl' else torch.device('cpu')
        t = torch.as_tensor(np.ascontiguousarray(idxs, dtype=np.int64), device=dev)
        dist.broadcast(t, src=0)
        return t.cpu().numpy()

    def _gather_kmm(self, Bt, idxs):
        """K_mm[a, b] = K[idxs[a], idxs[b]] from the transposed local panel Bt[b, r] = K[row0 + r, idxs[b]]."""
        eng = self.engine
        k = Bt.shape[0]
        idx_t = torch.as_tensor(np.asarray(idxs, dtype=np.int64), device=eng.device)
        loc = (idx_t >= eng.row0) & (idx_t < eng.row0 + eng.n_local)
        K_mm = torch.zeros((k, k), dtype=torch.float64, device=eng.device)
        a_pos = torch.nonzero(loc).ravel()
        if a_pos.numel():
            K_mm[a_pos, :] = Bt[:, idx_t[a_pos] - eng.row0].t()
        eng.allreduce_sum_(K_mm)
        return K_mm

    def _cho_factor_stable(self, Mat):
        """Lower Cholesky factor of ``Mat +- 1e-15 I`` (iterative_solver.py:576-583).  The reference picks
        the sign from the lowest eigenvalue (``eigh``): + if it is <= 0, - otherwise.  A matrix has a
        positive lowest eigenvalue exactly when an unshifted Cholesky succeeds, so the device probes with
        one extra factorisation instead of an eigensolver."""
        eng = self.engine
        probe = Mat.clone()
        sgn = -1.0 if eng.potrf_lower(probe, raise_on_fail=False) == 0 else 1.0
        del probe
        Mat.diagonal().add_(sgn * 1.0e-15)
        eng.potrf_lower(Mat)  # raises LinAlgError like scipy's cho_factor
        return Mat

    def _init_precon_operator(self, task, R_desc, R_d_desc, tril_perms_lin, inducing_pts_idxs, callback=None):
        """Nystroem preconditioner (iterative_solver.py:95-322): B[m, n] with P v = (B^T(B v) - v)/lam."""
        eng, lam = self.engine, task['lam']
        Bt = eng.kernel_columns(inducing_pts_idxs)                 # K_nm^T   (:112-121)
        K_mm = self._gather_kmm(Bt, inducing_pts_idxs)             #          (:124)
        K_mm.neg_()
        Lc = self._cho_factor_stable(K_mm)                         # U^T      (:218)
        eng.trsm_rows(Lc, Bt)                                      # K~^T = U^{-T} K_mn   (:219-226)
        inner = eng.syrk_rows(Bt, shift=lam)                       # K~^T K~ + lam I      (:230-231)
        Lc2 = self._cho_factor_stable(inner)                       #          (:254)
        eng.trsm_rows(Lc2, Bt)                                     # B        (:260-283)
        return LowRankPreconditioner(eng, Bt, lam, -1.0)

    def _init_precon_operator_sb(self, task, R_desc, R_d_desc, tril_perms_lin, inducing_pts_idxs, callback=None):
        """``_init_precon_operator_sb`` (iterative_solver.py:326-381): same pipeline on -K with a 1e-16 jitter."""
        eng, lam = self.engine, task['lam']
        Bt = eng.kernel_columns(inducing_pts_idxs, scale=-1.0)     # (-K_nm)^T  (:343-352)
        K_mm = self._gather_kmm(Bt, inducing_pts_idxs)
        K_mm.diagonal().add_(1e-16)
        eng.potrf_lower(K_mm)                                      # L_m      (:370)
        eng.trsm_rows(K_mm, Bt)                                    # Kbar^T   (:371)
        inner = eng.syrk_rows(Bt, shift=lam)                       #          (:372)
        eng.potrf_lower(inner)                                     # L_inner  (:373)
        eng.trsm_rows(inner, Bt)                                   # P_invers (:374)
        return LowRankPreconditioner(eng, Bt, lam, -1.0)           # -(a - P^T P a)/lam  (:376-379)

    def _lev_scores(self, R_desc, R_d_desc, tril_perms_lin, sig, lam, use_E_cstr, n_inducing_pts,
                    idxs_ordered_by_lev_score=None, callback=None):
        """Approximate ridge leverage scores (iterative_solver.py:447-552); returns host arrays
        ``(lev_scores[n], argsort)`` like the reference."""
        eng = self.engine
        n_train, dim_i = eng.M, eng.dim_i
        dim_m = np.maximum(1, n_inducing_pts // 4) * dim_i
        if idxs_ordered_by_lev_score is None:
            lev_approx_idxs = self._sync_draw(np.sort(np.random.choice(n_train * dim_i, dim_m, replace=False)))
        else:
            assert len(idxs_ordered_by_lev_score) == n_train * dim_i
            lev_approx_idxs = np.sort(idxs_ordered_by_lev_score[-dim_m:])
        B = eng.kernel_columns(lev_approx_idxs)                    # K_mn     (:489-498)
        K_mm = self._gather_kmm(B, lev_approx_idxs)
        K_mm.neg_()
        Lc = self._cho_factor_stable(K_mm)                         #          (:507)
        eng.trsm_rows(Lc, B)                                       # B = U^{-T} K_mn      (:516-533)
        B_BT_lam = eng.syrk_rows(B, shift=lam)                     #          (:535-536)
        C = self._cho_factor_stable(B_BT_lam)                      #          (:545)
        eng.trsm_rows(C, B)                                        # C^{-T} B (:546-548)
        scores_local = (B * B).sum(dim=0)                          # column sums of squares (:550)
        lev_scores = allgather_rows(eng, scores_local).cpu().numpy()
        return lev_scores, np.argsort(lev_scores)

    def _init_kernel_operator(self, task, R_desc, R_d_desc, tril_perms_lin, lam, n, callback=None):
        """``K_op``: v -> K v - lam v (iterative_solver.py:383-445), assembled or matrix-free."""
        return KernelOperator(self.engine, lam, 1.0, self.K_local)

    def _choose_kernel_mode(self, task, k):
        """'assembled_sym' (symmetric tile storage in HBM, every entry read once per iteration), 'assembled'
        (plain row block + GEMV) or 'matrix_free'.  'auto' = matrix-free when the symmetric storage does not fit next
        to the preconditioner on some rank (BASELINE.json north_star), otherwise whichever operator is FASTER on this
        system: both are applied twice to a probe vector (the assembly costs tens of milliseconds) and the ranks agree
        on the slower rank's timings.  Small descriptors favour the on-the-fly operator (cfg2: 0.5 ms vs 7.5 ms per
        matvec), large molecules with few points the stored one (cfg1)."""
        mode = task.get('kernel_mode', 'auto')
        if mode not in ('auto', 'assembled', 'assembled_sym', 'matrix_free'):
            raise ValueError("task['kernel_mode'] must be 'auto', 'assembled', 'assembled_sym' or 'matrix_free'")
        if mode != 'auto':
            return mode
        eng = self.engine
        free, _ = torch.cuda.mem_get_info(eng.device)
        need = eng.symop_storage_elems() * 8 + (2 << 30)
        fits = torch.tensor([1 if need < free else 0], dtype=torch.int32, device=eng.device)
        if eng.world > 1:  # the ranks hold different tile sets: decide together
            import torch.distributed as dist
            dist.all_reduce(fits, op=dist.ReduceOp.MIN)
        if int(fits.item()) != 1:
            return 'matrix_free'
        v = torch.ones(eng.world * (-(-eng.M // eng.world)) * eng.dim_i, dtype=torch.float64, device=eng.device)
        Ksym = eng.symop_assemble()
        times = []
        for op in (lambda: eng.symop_apply(Ksym, v), lambda: eng.matvec_free(v)):
            op()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op()
            op()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        t = torch.tensor(times, dtype=torch.float64, device=eng.device)
        if eng.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        self.timings['auto_probe_ms'] = {'assembled_sym': float(t[0]) / 2, 'matrix_free': float(t[1]) / 2}
        if float(t[0]) <= float(t[1]):
            self._probe_K = Ksym   # already assembled: the solve uses it
            return 'assembled_sym'
        del Ksym
        return 'matrix_free'

    # ------------------------------------------------------------------ device-resident solve
    def solve_device(self, task, eng, y_t, break_percentage, str_preconditioner, n_inducing_pts, x0=None,
                     on_segment=None):
        """Everything between the H2D of the inputs and the D2H of the solution: preconditioner build,
        kernel operator, PCG.  ``y_t`` is the full right-hand side on the device.  Returns
        ``(x_local, iters, resid, info, inducing_pts_idxs, info_cholesky, t_precon, t_cg)``."""
        self.engine = eng
        for name, value in (task.get('_options') or {}).items():   # library switches (diagnostics, A/B runs)
            eng.set_option(name, value)
        n, n_train, n_atoms = eng.n, eng.M, eng.N
        sig, lam = task['sig'], task['lam']
        R_desc = R_d_desc = tril_perms_lin = None  # geometry lives in the engine

        def sync():
            torch.cuda.synchronize()

        start_preconditioner = timeit.default_timer()
        info_cholesky = None
        if str_preconditioner in NYSTROM_KEYS:
            k = int(break_percentage * n)
            if str_preconditioner == 'random_scores':
                inducing_pts_idxs = self._sync_draw(np.sort(np.random.choice(np.arange(n), size=k, replace=False)))
            elif str_preconditioner in ['truncated_cholesky', 'truncated_cholesky_custom']:
                k_truncate = task['truncated_cholesky']
                k_truncate = k_truncate if k_truncate < k else k
                k_chol = int(float(k_truncate / n) * n)  # :704 -> iterative_cholesky.py:135 (may be k_truncate - 1)
                _, idx_t, _, _ = eng.pchol_build(k_chol, want_times=False)
                index_columns_cholesky = idx_t.cpu().numpy()
                inducing_pts_cholesky = index_columns_cholesky[:k_truncate]
                k_random = int(k - k_truncate) if k_truncate < k else 0
                inducing_pts_random = self._sync_draw(
                    np.random.choice(index_columns_cholesky[k_truncate:], size=k_random, replace=False))
                inducing_pts_idxs = np.sort(np.concatenate([inducing_pts_cholesky, inducing_pts_random]))
            else:  # 'lev_scores', 'inverse_lev', 'lev_random'
                lev_scores, idxs_ordered = self._lev_scores(R_desc, R_d_desc, tril_perms_lin, sig, lam, False,
                                                            n_inducing_pts)
                if str_preconditioner == 'inverse_lev':
                    inducing_pts_idxs = np.sort(idxs_ordered[:k])
                elif str_preconditioner == 'lev_scores':
                    inducing_pts_idxs = np.sort(idxs_ordered[-k:])
                else:
                    p = lev_scores / lev_scores.sum()
                    inducing_pts_idxs = self._sync_draw(np.sort(np.random.choice(np.arange(n), size=k, replace=False, p=p)))
            assert inducing_pts_idxs.shape == (k,), 'Incorrect number of inducing points.'
            if str_preconditioner == 'truncated_cholesky_custom':
                P_op = self._init_precon_operator_sb(task, R_desc, R_d_desc, tril_perms_lin, inducing_pts_idxs)
            else:
                P_op = self._init_precon_operator(task, R_desc, R_d_desc, tril_perms_lin, inducing_pts_idxs)
        elif str_preconditioner == 'cholesky':
            k = int(break_percentage * n)  # iterative_cholesky.py:135
            Lt, idx_t, _, step_s = eng.pchol_build(k)
            sync()
            self.timings['pchol_build'] = timeit.default_timer() - start_preconditioner
            info_cholesky = {'time_cholesky': step_s, 'L.shape': (n, k), 'index_columns': idx_t.cpu().numpy()}
            # (L L^T + lam I)^{-1} three ways (same operator; they differ in rounding only):
            #   'woodbury'  (default) the reference's formula (iterative_cholesky.py:141-148) with an accurately
            #               summed Gram: reproduces the reference's iteration counts where those are well defined
            #   'projected' orthonormal basis of range(L) + k x k inverse + first-order exact complement projector:
            #               no 1/lam cancellation, hence no residual plateau; iteration counts <= the reference's
            #   'orthonormal' the projected form without the defect correction (round 1; kept for A/B runs)
            form = task.get('precon_form', 'woodbury')
            if form not in ('orthonormal', 'woodbury', 'projected'):
                raise ValueError("task['precon_form'] must be 'woodbury', 'projected' or 'orthonormal'")
            Mk = E = None
            if k == 0:
                T = None
            elif form == 'woodbury':
                T = eng.woodbury_factor_(Lt, lam)
            else:
                try:
                    if form == 'projected':
                        T, Mk, E = eng.projected_factor_(Lt, lam)
                    else:
                        T, Mk = eng.orthonormal_factor_(Lt, lam)
                except np.linalg.LinAlgError:
                    # L^T L is numerically singular (pivots far below eps * max): only the shifted Gram of the
                    # reference's formula is positive definite.  Lt is untouched when the first Cholesky fails.
                    T, Mk, E, form = eng.woodbury_factor_(Lt, lam), None, None, 'woodbury'
            self.timings['precon_form'] = form
            P_op = LowRankPreconditioner(eng, T, lam, 1.0, Mk=Mk, E=E)
            inducing_pts_idxs = np.arange(int(break_percentage * n))  # :792
        else:
            raise NotImplementedError(f'str_preconditioner = {str_preconditioner}')
        sync()
        total_time_preconditioner = timeit.default_timer() - start_preconditioner
        self.last_P_op = P_op

        # kernel operator: assembled row block or matrix-free
        k_rank = 0 if P_op.T is None else P_op.T.shape[0]
        mode = self._choose_kernel_mode(task, k_rank)
        t0 = timeit.default_timer()
        self.K_local = None
        eng.set_option('symmetric_gemv', 1 if mode == 'assembled_sym' else 0)
        if mode == 'assembled':
            self.K_local = eng.kernel_assemble(out=task.get('_K_buffer'))
        elif mode == 'assembled_sym':  # symmetric tile storage: half the bytes per matvec and per rank
            self.K_local = getattr(self, '_probe_K', None)
            self._probe_K = None
            if self.K_local is None:
                self.K_local = eng.symop_assemble(out=task.get('_K_buffer'))
        sync()
        self.timings['assemble'] = timeit.default_timer() - t0
        self.timings['kernel_mode'] = mode

        maxiter = int(task.get('_maxiter', 3 * n_atoms * n_train * 5))  # :1002 (5 n; '_maxiter' caps sweeps)
        tic_start = timeit.default_timer()
        b_local = y_t[eng.row0:eng.row0 + eng.n_local].contiguous()
        kw = dict(K_local=self.K_local, T=P_op.T, precon_sign=P_op.sign, want_hist=bool(task.get('_want_hist')),
                  Mk=P_op.Mk, E=P_op.E)
        if on_segment is None:
            res = eng.pcg(b_local, lam, task['solver_tol'], maxiter, x0=x0, **kw)
        else:
            # One uninterrupted CG recurrence cut into segments: after each, the caller gets the current iterate (the
            # reference writes an unconverged model every ~2 minutes from scipy's callback, iterative_solver.py:919-954).
            period = float(task.get('_checkpoint_seconds', 120.0))
            seg = int(task.get('_checkpoint_first_iters', 50))
            done, x_io, hist0 = 0, None, None
            while True:
                cap = min(done + max(seg, 1), maxiter)
                t_seg = timeit.default_timer()
                res = eng.pcg(b_local, lam, task['solver_tol'], cap, x0=x0 if done == 0 else None,
                              resume_iters=done, x_inout=x_io, **kw)
                sync()
                x_io, it_now, info_now = res[0], res[1], res[3]
                if hist0 is None and kw['want_hist']:
                    hist0 = res[5][0]
                if info_now == 0 or it_now >= maxiter or it_now < cap:
                    break
                dt = max(timeit.default_timer() - t_seg, 1e-6)
                seg = int(np.ceil(period * (it_now - done) / dt))
                done = it_now
                on_segment(x_io, done, res[2])
            if kw['want_hist']:
                res[5][0] = hist0
        x, iters, resid, info, bnrm2 = res[:5]
        if task.get('_want_hist'):
            self.timings['resid_hist_rel'] = res[5] / bnrm2
        sync()
        total_time_cg = timeit.default_timer() - tic_start
        self.timings['peer_collectives'] = bool(getattr(eng, 'peer_collectives', False))
        self.timings.update(cg=total_time_cg, preconditioner=total_time_preconditioner, cg_iters=iters, k=k_rank,
                            pcg_stats=dict(eng.last_pcg_stats))
        return x, iters, resid, info, inducing_pts_idxs, info_cholesky, total_time_preconditioner, total_time_cg

    # ------------------------------------------------------------------ the solve step
    def solve(self, task, R_desc, R_d_desc, tril_perms_lin, y, y_std, save_progr_callback=None,
              break_percentage=None, str_preconditioner='', flag_eigvals=False):
        start_solve_routine = timeit.default_timer()
        n_train, n_atoms = task['R_train'].shape[:2]
        n = 3 * n_train * n_atoms
        sig, lam = task['sig'], task['lam']

        if task.get('use_E_cstr', False):
            raise NotImplementedError('use_E_cstr=True is outside the hot-path contract (the reference GPU path '
                                      'aborts as well, iterative_solver.py:834-836)')
        if flag_eigvals:
            raise NotImplementedError('flag_eigvals diagnostics (dev_utils.get_eigvals, O(n^3)) are out of scope')
        if str_preconditioner in OUT_OF_CONTRACT:
            raise NotImplementedError(f'str_preconditioner = {str_preconditioner} (full SVD, O(n^3)) is out of scope')
        if 'inducing_pts_idxs' in task:
            assert False, 'Nor applicable in this setting'  # iterative_solver.py:680
        alphas0_F = task['alphas0_F'] if 'alphas0_F' in task else None
        num_iters0 = task['solver_iters'] if 'solver_iters' in task else 0
        if break_percentage is None:
            n_inducing_pts_init = int(task['n_inducing_pts_init'])
        else:
            n_inducing_pts_init = int(max(np.ceil(break_percentage * n_train), 1))
        n_inducing_pts = min(n_train, n_inducing_pts_init)

        def sync():
            torch.cuda.synchronize()

        eng = self._make_engine(task, R_desc, R_d_desc, tril_perms_lin)
        assert eng.n == n
        y_t = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float64), device=eng.device)
        h2d_bytes = eng.h2d_bytes + y_t.numel() * 8
        x0 = None
        if alphas0_F is not None:
            x0_full = torch.as_tensor(-np.asarray(alphas0_F, dtype=np.float64).ravel(), device=eng.device)
            x0 = x0_full[eng.row0:eng.row0 + eng.n_local].contiguous()
        sync()
        on_segment = None
        if save_progr_callback is not None:
            def on_segment(x_local, iters_done, resid_now):
                """Unconverged model for the caller (iterative_solver.py:919-954): current coefficients, iteration count
                and -- when energies are trained -- the integration constant of the current iterate."""
                alphas_now = (-allgather_rows(eng, x_local)).cpu().numpy()
                gt = self.gdml_train
                if gt is None or not hasattr(gt, 'create_model'):
                    unconv = {'type': 'm', 'alphas_F': alphas_now, 'solver_iters': num_iters0 + iters_done + 1,
                              'solver_resid': resid_now}
                else:
                    extra = {'engine': eng} if getattr(gt, '_mlffpc_native', False) else {}  # the reference's has no such kwarg
                    unconv = gt.create_model(task, 'cg', R_desc, R_d_desc, tril_perms_lin, y_std, alphas_now,
                                             solver_resid=resid_now, solver_iters=num_iters0 + iters_done + 1,
                                             norm_y_train=np.linalg.norm(y), inducing_pts_idxs=None, **extra)
                    if task.get('E_train') is not None and task.get('use_E', False):
                        a_dev = torch.as_tensor(np.ascontiguousarray(alphas_now), device=eng.device)
                        E_raw, _ = eng.predict(eng._R_desc, eng._R_d_desc, alphas=a_dev, want_E=True)
                        E_pred = E_raw.cpu().numpy() * y_std
                        E_ref = np.squeeze(task['E_train'])
                        unconv['c'] = np.sum(E_ref - E_pred) / E_ref.shape[0]
                if eng.rank == 0:
                    save_progr_callback(unconv)

        (x, iters, resid, info, inducing_pts_idxs, info_cholesky, total_time_preconditioner,
         total_time_cg) = self.solve_device(task, eng, y_t, break_percentage, str_preconditioner, n_inducing_pts, x0,
                                            on_segment=on_segment)
        total_time_cholesky = total_time_preconditioner
        k_rank = self.timings['k']

        alphas = (-allgather_rows(eng, x)).cpu().numpy()  # :1009
        d2h_bytes = alphas.nbytes
        is_conv = info == 0
        # the legacy scipy driver calls the callback once per iteration and once more on exit (:956)
        num_iters = num_iters0 + iters + 1
        total_time_solve = timeit.default_timer() - start_solve_routine
        info_iterative_solver = {'is_conv': is_conv,
                                 'total_time_cholesky': total_time_cholesky,
                                 'total_time_cg': total_time_cg,
                                 'total_time_solve': total_time_solve,
                                 'total_time_preconditioner': total_time_preconditioner}
        if str_preconditioner == 'cholesky':
            info_iterative_solver.update(info_cholesky)
            info_iterative_solver['precon_form'] = self.timings.get('precon_form')
        self.timings.update(solve=total_time_solve, h2d_bytes=h2d_bytes, d2h_bytes=d2h_bytes)
        train_rmse = resid / np.sqrt(len(y))
        return alphas, num_iters, resid, train_rmse, inducing_pts_idxs, is_conv, info_iterative_solver
