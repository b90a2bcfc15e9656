"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for rendezvous, our own NCCL
communicator (created inside libmlffpc from a broadcast unique id) for the data-path collectives.

Sharding: rows of K -- and of L/T/B, r, z, q, x, diag -- are partitioned by
contiguous blocks of training points; geometry is replicated.  Per CG iteration: one allgather of the
search direction and three scalar + one k-vector allreduce.  Per pivot step: one 32-byte allgather
(arg-max candidates) and one <= k-double allreduce (the pivot row of the factor).
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib


def dist_info():
    """(rank, world, local_rank) from torch.distributed / torchrun env, (0, 1, 0) otherwise."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size(), int(os.environ.get('LOCAL_RANK', dist.get_rank()))
    return 0, 1, 0


def broadcast_bytes(buf, src=0):
    """Broadcast a small bytes object from rank ``src`` through the default process group
    (works for gloo and nccl backends)."""
    import torch.distributed as dist

    backend = dist.get_backend()
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def init_engine_comm(engine):
    """Attach an NCCL communicator to ``engine`` (called by Engine.__init__ when world > 1)."""
    lib = engine.lib
    path = _lib.nccl_library_path().encode()
    ident = (ctypes.c_char * 128)()
    if engine.rank == 0:
        _lib.check(lib.mlffpc_comm_unique_id(path, ident))
    raw = broadcast_bytes(bytes(ident.raw), src=0)
    ident2 = (ctypes.c_char * 128).from_buffer_copy(raw)
    _lib.check(lib.mlffpc_comm_init(engine.ctx, path, ident2, engine.rank, engine.world))


def allgather_rows(engine, x_local):
    """Replicate a row-sharded vector: returns the full n-vector on every rank."""
    if engine.world == 1:
        return x_local
    ppr = (engine.M + engine.world - 1) // engine.world
    n_pad = ppr * engine.dim_i
    buf = torch.zeros(engine.world * n_pad, dtype=torch.float64, device=engine.device)
    buf[engine.row0:engine.row0 + engine.n_local] = x_local
    send = buf[engine.rank * n_pad:(engine.rank + 1) * n_pad]
    _lib.check(engine.lib.mlffpc_allgather(engine.ctx, ctypes.c_void_p(send.data_ptr()),
                                           ctypes.c_void_p(buf.data_ptr()), n_pad * 8, engine._stream()))
    return buf[:engine.n]


def shard_slices(M, dim_i, world):
    """[(row0, row1)] per rank for the ceil(M/world) point partition -- host-side helper for tests."""
    ppr = (M + world - 1) // world
    out = []
    for r in range(world):
        pt0, pt1 = r * ppr, min((r + 1) * ppr, M)
        out.append((pt0 * dim_i, max(pt0, pt1) * dim_i))
    return out


def symop_plan(M, world, rank):
    """Host mirror of the symmetric tile plan in csrc/symop.cu: [(i_pt0, i_pt1, j_pt0, j_pt1, is_diag)] for
    ``rank``.  Every unordered pair of point blocks is covered by exactly one rank and every rank reads
    about half of its row block (tests/test_dist_gloo.py checks both properties)."""
    ppr = (M + world - 1) // world

    def blk(b):
        return min(b * ppr, M), min((b + 1) * ppr, M)

    tiles = []

    def add(i0, i1, j0, j1, diag):
        if i1 > i0 and j1 > j0:
            tiles.append((i0, i1, j0, j1, diag))

    g = rank
    g0, g1 = blk(g)
    add(g0, g1, g0, g1, 1)
    for d in range(1, (world - 1) // 2 + 1):
        h0, h1 = blk((g + d) % world)
        add(g0, g1, h0, h1, 0)
    if world > 1 and world % 2 == 0:
        h0, h1 = blk((g + world // 2) % world)
        if g < world // 2:
            add(g0, g0 + (g1 - g0 + 1) // 2, h0, h1, 0)
        else:
            add(g0, g1, h0 + (h1 - h0 + 1) // 2, h1, 0)
    return tiles


def symop_entries_read(tiles, dim_i, strip_rows=32):
    """Matrix entries one symmetric matvec reads from HBM for these tiles (diagonal tiles: the lower
    triangle by ``strip_rows``-row strips including the diagonal blocks; other tiles: everything)."""
    total = 0
    for (i0, i1, j0, j1, diag) in tiles:
        nr, nc = (i1 - i0) * dim_i, (j1 - j0) * dim_i
        if not diag:
            total += nr * nc
            continue
        for r0 in range(0, nr, strip_rows):
            rows = min(strip_rows, nr - r0)
            total += rows * (r0 + rows)
    return total
