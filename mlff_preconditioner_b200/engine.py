"""Device engine: one ``mlffpc_ctx`` + the torch CUDA tensors that back it.

PyTorch is only the buffer/stream/process-group plumbing here; all arithmetic happens in
libmlffpc.so (hand-written sm_100a CUDA) through the C ABI of include/mlffpc.h.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .desc import desc_perms_from_tril_perms_lin, n_atoms_from_dim_d


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def atom_perms_from_desc_perms(desc_perms, n_atoms):
    """Recover the atom permutations P_p from descriptor permutations pi_p (the inverse of
    ``Desc.perm``, utils/desc.py:360-389): the image of atom x is the atom common to the images of two
    pairs that contain x.  For N = 2 the identity is returned (both choices give the same pi)."""
    desc_perms = np.asarray(desc_perms)
    S, D = desc_perms.shape
    a, b = np.tril_indices(n_atoms, k=-1)
    if n_atoms < 3:
        return np.tile(np.arange(n_atoms, dtype=np.int32), (S, 1))
    pair_of = {}
    for d in range(D):
        pair_of[(a[d], b[d])] = d
        pair_of[(b[d], a[d])] = d
    out = np.zeros((S, n_atoms), dtype=np.int32)
    for p in range(S):
        for x in range(n_atoms):
            others = [y for y in range(n_atoms) if y != x][:2]
            imgs = []
            for y in others:
                e = desc_perms[p, pair_of[(x, y)]]
                imgs.append({int(a[e]), int(b[e])})
            common = imgs[0] & imgs[1]
            assert len(common) == 1, 'descriptor permutation is not induced by an atom permutation'
            out[p, x] = common.pop()
    return out


def desc_from_R_device(lib, ctx, R, n_atoms, device, stream):
    """Device ``Desc.from_R`` (utils/desc.py:292-358) through mlffpc_desc_from_r; R: host array or device tensor."""
    if not torch.is_tensor(R):
        R = torch.as_tensor(np.ascontiguousarray(R, dtype=np.float64), device=device)
    R = R.to(device=device, dtype=torch.float64).reshape(-1, n_atoms, 3).contiguous()
    Bq = R.shape[0]
    D = n_atoms * (n_atoms - 1) // 2
    R_desc = torch.empty((Bq, D), dtype=torch.float64, device=device)
    R_d_desc = torch.empty((Bq, D, 3), dtype=torch.float64, device=device)
    _lib.check(lib.mlffpc_desc_from_r(ctx, _ptr(R), Bq, int(n_atoms), _ptr(R_desc), _ptr(R_d_desc), stream))
    return R_desc, R_d_desc


def descriptors_on_device(R, device=None):
    """(R_desc[M, D], R_d_desc[M, D, 3]) as CUDA tensors for geometries R[M, N, 3] -- ``Desc.from_R`` without an
    Engine (a short-lived context; the kernel needs no geometry tables)."""
    if not torch.cuda.is_available():
        raise _lib.MlffpcError('mlff_preconditioner_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    lib = _lib.load()
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    n_atoms = (R.shape[-2] if R.ndim == 3 else None) if not torch.is_tensor(R) else (R.shape[-2] if R.dim() == 3 else None)
    if n_atoms is None:
        raise ValueError('R must have shape [M, N, 3]')
    ctx = ctypes.c_void_p()
    _lib.check(lib.mlffpc_create(ctypes.byref(ctx), dev.index))
    try:
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        out = desc_from_R_device(lib, ctx, R, n_atoms, dev, stream)
        torch.cuda.current_stream(dev).synchronize()
    finally:
        lib.mlffpc_destroy(ctx)
    return out


def shard_points(M, rank, world):
    """Row-block partition by training points: rank r owns [r*ceil(M/W), min((r+1)*ceil(M/W), M))."""
    ppr = (M + world - 1) // world
    # decided from (M, world) alone so that EVERY rank raises -- a rank-local check would leave the other ranks
    # blocked in their first collective
    if (world - 1) * ppr >= M:
        raise ValueError('%d ranks leave rank %d without training points (M = %d, %d points per rank): use at most %d '
                         'ranks' % (world, world - 1, M, ppr, (M + ppr - 1) // ppr))
    return rank * ppr, min((rank + 1) * ppr, M)


class Engine(object):
    """Geometry-bound solver context on one GPU (one rank of a row-block sharded job)."""

    def __init__(self, R_desc, R_d_desc, tril_perms_lin, sig, perms=None, device=None, rank=0, world=1,
                 init_comm=None, R=None):
        """``R_desc[M, D]`` / ``R_d_desc[M, D, 3]``: host arrays (uploaded) or CUDA tensors (used in place).
        ``R_d_desc=None`` makes a prediction-only context (no assembly / matvec with alphas).  With ``R[M, N, 3]``
        (and ``R_desc=None``) the descriptors are computed on the device (mlffpc_desc_from_r)."""
        if not torch.cuda.is_available():
            raise _lib.MlffpcError('mlff_preconditioner_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        torch.cuda.set_device(self.device)
        self.rank, self.world = rank, world
        ctx = ctypes.c_void_p()
        _lib.check(self.lib.mlffpc_create(ctypes.byref(ctx), self.device.index))
        self.ctx = ctx
        self.h2d_bytes = 0

        def to_dev(a):
            if torch.is_tensor(a):
                return a.to(device=self.device, dtype=torch.float64).contiguous()
            a = np.ascontiguousarray(a, dtype=np.float64)
            self.h2d_bytes += a.nbytes
            return torch.from_numpy(a).to(self.device)

        if R_desc is None:
            if R is None:
                raise ValueError('Engine needs descriptors (R_desc, R_d_desc) or geometries (R)')
            Rt = to_dev(R)
            n_atoms = Rt.shape[-2] if Rt.dim() == 3 else None
            if n_atoms is None:
                raise ValueError('R must have shape [M, N, 3]')
            R_desc, R_d_desc = desc_from_R_device(self.lib, self.ctx, Rt, n_atoms, self.device, self._stream())
        self._R_desc = to_dev(R_desc)
        self._R_d_desc = None if R_d_desc is None else to_dev(R_d_desc)
        self.M, self.D = self._R_desc.shape
        self.N = n_atoms_from_dim_d(self.D)
        assert self._R_d_desc is None or tuple(self._R_d_desc.shape) == (self.M, self.D, 3)
        self.dim_i = 3 * self.N
        self.n = self.M * self.dim_i
        self.sig = float(sig)
        dperms = desc_perms_from_tril_perms_lin(tril_perms_lin, self.D)
        self.S = dperms.shape[0]
        if perms is None:
            aperms = atom_perms_from_desc_perms(dperms, self.N)
        else:
            aperms = np.ascontiguousarray(perms, dtype=np.int32)
            assert aperms.shape == (self.S, self.N)
        self.pt0, self.pt1 = shard_points(self.M, rank, world)
        self.n_local = (self.pt1 - self.pt0) * self.dim_i
        self.row0 = self.pt0 * self.dim_i
        self.h2d_bytes += dperms.nbytes + aperms.nbytes
        self._dperms = torch.from_numpy(dperms).to(self.device)
        self._aperms = torch.from_numpy(aperms).to(self.device)
        if world > 1:
            if init_comm is None:
                raise ValueError('world > 1 needs init_comm (see dist.init_engine_comm)')
            init_comm(self)
        nbytes = ctypes.c_int64()
        _lib.check(self.lib.mlffpc_geometry_workspace_bytes(self.M, self.N, self.S, ctypes.byref(nbytes)))
        self._geo_ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.mlffpc_set_geometry(
            self.ctx, self.M, self.N, self.S, _ptr(self._R_desc), _ptr(self._R_d_desc), _ptr(self._dperms),
            _ptr(self._aperms), self.sig, self.pt0, self.pt1, _ptr(self._geo_ws), nbytes.value, self._stream()))
        self._ws_cache = {}
        self.peer_collectives = False
        if world > 1:
            self._peer_setup()

    def _peer_setup(self, k_max=65536):
        """Map every rank's communication buffer (CUDA IPC) so the inner loops can use NVLink peer stores / loads
        instead of one library collective per step (include/mlffpc.h, csrc/peer.cuh).  All ranks take the same
        decision: if any rank cannot export or import, every rank falls back to NCCL.  MLFFPC_PEER=0 disables it."""
        import os
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() != self.world:
            return
        want = os.environ.get('MLFFPC_PEER', '1') != '0'
        handle = (ctypes.c_char * 64)()
        ok = 1 if (want and self.lib.mlffpc_peer_export(self.ctx, int(k_max), handle) == 0) else 0
        dev = self.device if dist.get_backend() == 'nccl' else torch.device('cpu')
        mine = torch.tensor(list(handle.raw) + [ok], dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine)
        rows = [bytes(t.cpu().tolist()) for t in gathered]
        ok_all = all(r[64] == 1 for r in rows)
        if ok_all:
            blob = b''.join(r[:64] for r in rows)
            ok_all = self.lib.mlffpc_peer_import(self.ctx, blob, self.world) == 0
        flag = torch.tensor([1 if ok_all else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            self.lib.mlffpc_peer_disable(self.ctx)
            return
        self.peer_collectives = True

    # ---- plumbing -------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ws(self, key, nbytes):
        t = self._ws_cache.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws_cache[key] = t
        return t

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def close(self):
        if getattr(self, 'ctx', None) is not None and self.ctx:
            self.lib.mlffpc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- kernel entries -------------------------------------------------------------------
    def kernel_diag(self):
        """-diag(K) on the local rows (iterative_cholesky.py:241-373)."""
        out = self.empty(self.n_local)
        _lib.check(self.lib.mlffpc_kernel_diag(self.ctx, _ptr(out), self._stream()))
        return out

    def kernel_assemble(self, out=None):
        """Explicit local rows K[row0:row0+n_local, :] (train.py:1121-1308)."""
        if out is None:
            out = self.empty(self.n_local, self.n)
        assert out.shape == (self.n_local, self.n) and out.stride(1) == 1
        _lib.check(self.lib.mlffpc_kernel_assemble(self.ctx, _ptr(out), out.stride(0), self._stream()))
        return out

    def kernel_columns(self, cols, scale=1.0, out=None):
        """Transposed column panel out[c, r] = scale*K[row0 + r, cols[c]]; cols: int64 tensor/array."""
        if not torch.is_tensor(cols):
            cols = torch.as_tensor(np.asarray(cols, dtype=np.int64), device=self.device)
        cols = cols.to(device=self.device, dtype=torch.int64).contiguous()
        b = cols.numel()
        if out is None:
            out = self.empty(b, self.n_local)
        assert out.shape[0] == b and out.shape[1] == self.n_local and out.stride(1) == 1
        _lib.check(self.lib.mlffpc_kernel_columns(self.ctx, _ptr(cols), b, _ptr(out), out.stride(0), float(scale),
                                                  ctypes.c_void_p(0), 0, self._stream()))
        return out

    # ---- operators ------------------------------------------------------------------------
    def gemv(self, K, x, alpha=1.0, shift=0.0, x_off=0, out=None):
        n_rows, n_cols = K.shape
        assert K.stride(1) == 1 and x.is_contiguous() and x.numel() >= n_cols
        if out is None:
            out = self.empty(n_rows)
        _lib.check(self.lib.mlffpc_gemv(self.ctx, _ptr(K), n_rows, n_cols, K.stride(0), _ptr(x), _ptr(out),
                                        float(alpha), float(shift), int(x_off), self._stream()))
        return out

    def set_option(self, name, value):
        _lib.check(self.lib.mlffpc_set_option(self.ctx, name.encode(), int(value)))

    def symv(self, K, x, alpha=1.0, shift=0.0, out=None):
        """alpha*K x + shift*x for a symmetric assembled K (single GPU); reads only the lower triangle."""
        n = K.shape[0]
        assert K.shape == (n, n) and K.stride(1) == 1 and x.is_contiguous() and x.numel() >= n
        if out is None:
            out = self.empty(n)
        nb = ctypes.c_int64()
        _lib.check(self.lib.mlffpc_symv_workspace_bytes(n, ctypes.byref(nb)))
        ws = self._ws('symv', nb.value)
        _lib.check(self.lib.mlffpc_symv(self.ctx, _ptr(K), n, K.stride(0), _ptr(x), _ptr(out), float(alpha),
                                        float(shift), _ptr(ws), nb.value, self._stream()))
        return out

    # ---- symmetric tile operator (half the bytes, any number of ranks) ------------------------
    def set_layout(self, rank, world):
        """Override the tile partition (rank emulation on one GPU; tests only)."""
        self.set_option('layout_world', world)
        self.set_option('layout_rank', rank)
        self._lay_world = world

    def symop_tiles(self):
        """[(i_pt0, i_pt1, j_pt0, j_pt1, ld, offset, is_diag)] of this rank's tiles."""
        nt = ctypes.c_int64()
        _lib.check(self.lib.mlffpc_symop_tiles(self.ctx, ctypes.c_void_p(0), 0, ctypes.byref(nt)))
        buf = (ctypes.c_int64 * (8 * nt.value))()
        _lib.check(self.lib.mlffpc_symop_tiles(self.ctx, buf, nt.value, ctypes.byref(nt)))
        return [tuple(buf[8 * i + j] for j in range(7)) for i in range(nt.value)]

    @staticmethod
    def symop_unpack_diag(Ksym, off, nr):
        """Dense [nr, nr] view (copy) of a packed diagonal tile: band b = rows [256 b, 256 b + 256) stores the
        columns [0, 256 (b + 1)) with that pitch (csrc/symlayout.cuh); entries that are not stored come back NaN."""
        out = torch.full((nr, nr), float('nan'), dtype=torch.float64, device=Ksym.device)
        for b in range((nr + 255) // 256):
            r0, r1 = 256 * b, min(nr, 256 * b + 256)
            pitch = 256 * (b + 1)
            o = off + 65536 * (b * (b + 1) // 2)
            band = Ksym[o:o + (r1 - r0) * pitch].view(r1 - r0, pitch)
            w = min(pitch, nr)
            out[r0:r1, :w] = band[:, :w]
        return out

    def symop_storage_elems(self):
        ne = ctypes.c_int64()
        _lib.check(self.lib.mlffpc_symop_storage_elems(self.ctx, ctypes.byref(ne)))
        return ne.value

    def symop_assemble(self, out=None):
        """This rank's tiles of the symmetric storage (a flat fp64 tensor)."""
        ne = self.symop_storage_elems()
        if out is None:
            out = self.empty(ne)
        assert out.is_contiguous() and out.numel() >= ne
        _lib.check(self.lib.mlffpc_symop_assemble(self.ctx, _ptr(out), self._stream()))
        return out

    def symop_apply(self, Ksym, x_full, alpha=1.0, shift=0.0, out=None, partial=False):
        """alpha*(K x)_local + shift*x_local from the symmetric tile storage; ``partial=True`` returns this
        rank's full-length partial product (no collective) instead."""
        assert x_full.is_contiguous() and x_full.numel() >= self.n
        nb = ctypes.c_int64()
        _lib.check(self.lib.mlffpc_symop_workspace_bytes(self.ctx, ctypes.byref(nb)))
        ws = self._ws('symop', nb.value)
        if partial:
            world = self._layout_world()
            ppr = (self.M + world - 1) // world
            pout = torch.zeros(world * ppr * self.dim_i, dtype=torch.float64, device=self.device)
            _lib.check(self.lib.mlffpc_symop_apply(self.ctx, _ptr(Ksym), _ptr(x_full), ctypes.c_void_p(0), 1.0, 0.0,
                                                   _ptr(ws), nb.value, _ptr(pout), self._stream()))
            return pout[:self.n]
        if out is None:
            out = self.empty(self.n_local)
        _lib.check(self.lib.mlffpc_symop_apply(self.ctx, _ptr(Ksym), _ptr(x_full), _ptr(out), float(alpha),
                                               float(shift), _ptr(ws), nb.value, ctypes.c_void_p(0), self._stream()))
        return out

    def _layout_world(self):
        return getattr(self, '_lay_world', self.world)

    def matvec_free(self, v, alpha=1.0, shift=0.0, out=None):
        """alpha*(K v)_local + shift*v_local for the full n-vector v (predict.py:400-449,997-1052)."""
        assert v.numel() >= self.n and v.is_contiguous()
        if out is None:
            out = self.empty(self.n_local)
        nb = ctypes.c_int64()
        _lib.check(self.lib.mlffpc_matvec_free_workspace_bytes(self.ctx, ctypes.byref(nb)))
        ws = self._ws('mv', nb.value)
        _lib.check(self.lib.mlffpc_matvec_free(self.ctx, _ptr(v), _ptr(out), float(alpha), float(shift), _ptr(ws),
                                               nb.value, self._stream()))
        return out

    # ---- prediction (GDMLPredict.predict) --------------------------------------------------------
    def desc_from_R(self, R):
        """(R_desc[B, D], R_d_desc[B, D, 3]) on the device for geometries R[B, N, 3] / [B, 3N] (host or device)."""
        return desc_from_R_device(self.lib, self.ctx, R, self.N, self.device, self._stream())

    def d_desc_dot_vec(self, v):
        """beta[M, D] = J_m v_m on the device (utils/desc.py:394-405; model['R_d_desc_alpha'], train.py:640-645)."""
        assert v.numel() >= self.n and v.is_contiguous()
        out = self.empty(self.M, self.D)
        _lib.check(self.lib.mlffpc_d_desc_dot_vec(self.ctx, _ptr(v), _ptr(out), self._stream()))
        return out

    def predict(self, R_desc_q, R_d_desc_q, alphas=None, beta=None, want_E=True, max_batch=None):
        """Unscaled (E[B], F[B, 3N]) of the query geometries for coefficients ``alphas`` (full n-vector on the device)
        or ``beta[M, D] = J alpha`` (model['R_d_desc_alpha']); queries are processed in batches that keep the
        [B, 2 M S] pair tables under ~4 GB."""
        assert (alphas is None) != (beta is None), 'give exactly one of alphas and beta'
        Bq = R_desc_q.shape[0]
        F = self.empty(Bq, self.dim_i)
        E = self.empty(Bq) if want_E else None
        if max_batch is None:
            max_batch = max(64, min(32768, int((4 << 30) // max(1, 16 * self.M * self.S))))
        nb = ctypes.c_int64()
        for s0 in range(0, Bq, max_batch):
            b = min(max_batch, Bq - s0)
            _lib.check(self.lib.mlffpc_predict_workspace_bytes(self.ctx, b, ctypes.byref(nb)))
            ws = self._ws('predict', nb.value)
            xd = R_desc_q[s0:s0 + b]
            gd = R_d_desc_q[s0:s0 + b]
            assert xd.is_contiguous() and gd.is_contiguous()
            _lib.check(self.lib.mlffpc_predict(self.ctx, _ptr(xd), _ptr(gd), b, _ptr(alphas), _ptr(beta), _ptr(F[s0:s0 + b]),
                                               _ptr(E[s0:s0 + b]) if want_E else ctypes.c_void_p(0), _ptr(ws), nb.value,
                                               self._stream()))
        return E, F

    # ---- dense ----------------------------------------------------------------------------
    def dgemm(self, A, B, trans_b=False, alpha=1.0, beta=0.0, out=None):
        m, k = A.shape
        n = B.shape[0] if trans_b else B.shape[1]
        assert (B.shape[1] if trans_b else B.shape[0]) == k
        assert A.stride(1) == 1 and B.stride(1) == 1
        if out is None:
            out = self.empty(m, n)
        _lib.check(self.lib.mlffpc_dgemm(self.ctx, 1 if trans_b else 0, m, n, k, float(alpha), _ptr(A), A.stride(0),
                                         _ptr(B), B.stride(0), float(beta), _ptr(out), out.stride(0), self._stream()))
        return out

    def syrk_rows(self, X, shift=0.0):
        m, nc = X.shape
        W = self.empty(m, m)
        _lib.check(self.lib.mlffpc_syrk_rows(self.ctx, _ptr(X), m, nc, X.stride(0), float(shift), _ptr(W), m,
                                             self._stream()))
        return W

    def potrf_lower(self, W, raise_on_fail=True):
        """In-place lower Cholesky; returns LAPACK-style info (0 = ok)."""
        m = W.shape[0]
        info = ctypes.c_int()
        _lib.check(self.lib.mlffpc_potrf_lower(self.ctx, _ptr(W), m, W.stride(0), ctypes.byref(info), self._stream()))
        if info.value != 0 and raise_on_fail:
            raise np.linalg.LinAlgError('%d-th leading minor of the array is not positive definite' % info.value)
        return info.value

    def trsm_rows(self, Lf, X):
        m = Lf.shape[0]
        assert X.shape[0] == m and X.stride(1) == 1
        _lib.check(self.lib.mlffpc_trsm_rows(self.ctx, _ptr(Lf), m, Lf.stride(0), _ptr(X), X.shape[1], X.stride(0),
                                             self._stream()))
        return X

    def allreduce_sum_(self, t):
        _lib.check(self.lib.mlffpc_allreduce_sum(self.ctx, _ptr(t), t.numel(), self._stream()))
        return t

    # ---- pivoted partial Cholesky ---------------------------------------------------------
    def pchol_build(self, k, diag=None, forced_pivots=None, want_times=True):
        """(Lt[k, n_local], index_columns[n] int64, residual diag, step_seconds[k])."""
        k = int(k)
        if diag is None:
            diag = self.kernel_diag()
        else:
            diag = diag.clone()
        Lt = torch.zeros((k, self.n_local), dtype=torch.float64, device=self.device)
        idx = torch.empty(self.n, dtype=torch.int64, device=self.device)
        nb = ctypes.c_int64()
        _lib.check(self.lib.mlffpc_pchol_workspace_bytes(self.ctx, k, ctypes.byref(nb)))
        ws = self._ws('pchol', nb.value)
        times = (ctypes.c_float * max(k, 1))() if want_times else None
        fp = None
        if forced_pivots is not None:
            fp = torch.as_tensor(np.asarray(forced_pivots, dtype=np.int64), device=self.device)
        _lib.check(self.lib.mlffpc_pchol_build(self.ctx, k, _ptr(Lt), max(self.n_local, 1), _ptr(diag), _ptr(idx),
                                               _ptr(fp), times, _ptr(ws), nb.value, self._stream()))
        step_s = np.frombuffer(times, dtype=np.float32)[:k].astype(np.float64) * 1e-3 if want_times else None
        return Lt, idx, diag, step_s

    # ---- preconditioner -------------------------------------------------------------------
    def woodbury_factor_(self, Lt, lam):
        """In place: Lt -> T = chol(lam I + Lt Lt^T)^{-1} Lt (iterative_cholesky.py:141-143)."""
        k = Lt.shape[0]
        W = self.empty(k, k)
        _lib.check(self.lib.mlffpc_woodbury_factor(self.ctx, _ptr(Lt), k, Lt.stride(0), float(lam), _ptr(W),
                                                   self._stream()))
        return Lt

    def orthonormal_factor_(self, Lt, lam):
        """In place: Lt -> Qt (orthonormal rows spanning range(L)); returns (Qt, Mk) with
        Mk = (Qt L L^T Qt^T + lam I)^{-1} -- the cancellation-free form of (L L^T + lam I)^{-1}."""
        k = Lt.shape[0]
        Mk, W1, W2 = self.empty(k, k), self.empty(k, k), self.empty(k, k)
        _lib.check(self.lib.mlffpc_orthonormal_factor(self.ctx, _ptr(Lt), k, Lt.stride(0), float(lam), _ptr(Mk),
                                                      _ptr(W1), _ptr(W2), self._stream()))
        return Lt, Mk

    def projected_factor_(self, Lt, lam):
        """In place: Lt -> Qt as in ``orthonormal_factor_``; returns (Qt, Mk, E) with the defect
        E = Qt Qt^T - I from the extended-precision Gram (csrc/gramdd.cu) -- the two-pass projected form."""
        k = Lt.shape[0]
        Mk, E, W1, W2 = self.empty(k, k), self.empty(k, k), self.empty(k, k), self.empty(k, k)
        _lib.check(self.lib.mlffpc_projected_factor(self.ctx, _ptr(Lt), k, Lt.stride(0), float(lam), _ptr(Mk), _ptr(E),
                                                    _ptr(W1), _ptr(W2), self._stream()))
        return Lt, Mk, E

    def gram_defect(self, Q):
        """E = Q Q^T - I for Q[k, n_local] (summed over ranks), extended-precision accumulation."""
        k = Q.shape[0]
        E = self.empty(k, k)
        _lib.check(self.lib.mlffpc_gram_defect(self.ctx, _ptr(Q), k, Q.shape[1], Q.stride(0), _ptr(E), self._stream()))
        return E

    def precon_apply(self, T, lam, sign, r, out=None, Mk=None, E=None):
        if out is None:
            out = self.empty(self.n_local)
        k = 0 if T is None else T.shape[0]
        u = self.empty(4 * k + 8)
        _lib.check(self.lib.mlffpc_precon_apply(self.ctx, _ptr(T), k, 0 if T is None else T.stride(0), float(lam),
                                                float(sign), _ptr(r), _ptr(out), _ptr(u), _ptr(Mk), _ptr(E),
                                                self._stream()))
        return out

    # ---- PCG ------------------------------------------------------------------------------
    def pcg(self, b, lam, tol, maxiter, K_local=None, T=None, precon_sign=1.0, x0=None, want_hist=False, Mk=None,
            E=None, resume_iters=0, x_inout=None):
        """Returns (x_local, iters, resid, info, bnrm2[, hist]).  ``resume_iters`` / ``x_inout``: continue the run a
        previous call (same engine, nothing else run through ``pcg`` in between) stopped at its iteration cap."""
        k = 0 if T is None else T.shape[0]
        if x_inout is not None:
            x = x_inout
        else:
            x = torch.zeros(self.n_local, dtype=torch.float64, device=self.device) if x0 is None else x0.clone()
        nb = ctypes.c_int64()
        _lib.check(self.lib.mlffpc_pcg_workspace_bytes(self.ctx, k, 1 if K_local is None else 0, ctypes.byref(nb)))
        ws = self._ws('pcg', nb.value)
        out = (ctypes.c_double * 8)()
        hist = None
        if want_hist:
            hist = np.full(int(maxiter) + 1, np.nan)
        _lib.check(self.lib.mlffpc_pcg(
            self.ctx, _ptr(K_local), 0 if K_local is None else K_local.stride(0), float(lam), _ptr(T), k,
            0 if T is None else T.stride(0), float(precon_sign), _ptr(Mk), _ptr(E), _ptr(b), _ptr(x), float(tol),
            int(maxiter), int(resume_iters), out,
            hist.ctypes.data_as(ctypes.c_void_p) if hist is not None else ctypes.c_void_p(0), _ptr(ws), nb.value,
            self._stream()))
        self.last_pcg_stats = {'op_ms': float(out[4]), 'op_calls': int(out[5]), 'precon_ms': float(out[6])}
        res = (x, int(out[0]), float(out[1]), int(out[2]), float(out[3]))
        if want_hist:
            return res + (hist[:int(out[0]) + 1],)
        return res
