"""Device pivoted partial Cholesky behind the reference's generic signature
(``/root/reference/src/sGDML/sgdml/solvers/incomplete_cholesky.py:24-93``)."""
import numpy as np


class KernelColumns(object):
    """Column oracle of A = -K bound to a device :class:`~mlff_preconditioner_b200.engine.Engine`.

    Passing this as ``get_col`` lets :func:`pivoted_cholesky` run the whole pivot loop on the GPU with
    on-the-fly columns; calling it like the reference's ``get_col(i)`` returns one host column.
    """

    def __init__(self, engine):
        self.engine = engine

    def __call__(self, i):
        if self.engine.world != 1:
            raise NotImplementedError('host column access is single-GPU only')
        return self.engine.kernel_columns(np.array([int(i)], dtype=np.int64), scale=-1.0)[0].cpu().numpy()


def _check_inputs(get_col, diagonal, max_rank):
    """Same sanity rules as the reference (incomplete_cholesky.py:15-21), minus the wasted column."""
    assert isinstance(get_col, KernelColumns), \
        'get_col must be a KernelColumns bound to a device engine (arbitrary Python column callables are ' \
        'a CPU path; this package has none)'
    assert diagonal.ndim == 1, 'diag with more than one dimension'
    assert diagonal.shape[0] == get_col.engine.n_local, 'dimension from get_col and diag do not match'
    assert max_rank <= get_col.engine.n, f'max_rank = {max_rank} is too large'


def pivoted_cholesky(get_col, diagonal, max_rank, forced_pivots=None):
    """``(L, index_columns, info)`` like the reference.

    ``diagonal`` may be a numpy array or a CUDA tensor (this rank's rows of ``-diag(K)``).  ``L`` is a
    CUDA tensor *view* of shape ``(n_local, max_rank)`` (the device keeps the factor transposed,
    ``Lt[k, n_local]``; ``L = Lt.t()``), ``index_columns`` a numpy int64 array of length n whose first
    ``max_rank`` entries are the pivot sequence, ``info`` carries ``time_cholesky`` (seconds per step,
    CUDA-event timed), ``'L.shape'`` and ``index_columns`` (incomplete_cholesky.py:86-88).
    """
    import torch

    eng = get_col.engine
    if not torch.is_tensor(diagonal):
        diagonal = torch.as_tensor(np.asarray(diagonal, dtype=np.float64), device=eng.device)
    _check_inputs(get_col, diagonal, max_rank)
    Lt, idx, _, step_s = eng.pchol_build(int(max_rank), diag=diagonal, forced_pivots=forced_pivots)
    index_columns = idx.cpu().numpy()
    L = Lt.t()
    info = {'time_cholesky': step_s, 'L.shape': (eng.n, int(max_rank)), 'index_columns': index_columns}
    return L, index_columns, info
