"""ctypes binding of libmlffpc.so (the C ABI declared in include/mlffpc.h).

There is no CPU fallback: if the shared library is missing this module raises, and every product
entry point that needs the GPU fails loudly with it.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmlffpc.so')

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOT_PSD, ERR_LINALG, ERR_COMM, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p
c_str = ctypes.c_char_p

# name -> argtypes; every function returns int except the two noted below.  Kept in one table so the
# CPU test-suite can check that the library exports exactly what include/mlffpc.h declares.
SIGNATURES = {
    'mlffpc_create': [ctypes.POINTER(c_ptr), c_int],
    'mlffpc_destroy': [c_ptr],
    'mlffpc_comm_unique_id': [c_str, c_ptr],
    'mlffpc_comm_init': [c_ptr, c_str, c_ptr, c_int, c_int],
    'mlffpc_peer_export': [c_ptr, c_i64, c_ptr],
    'mlffpc_peer_import': [c_ptr, c_ptr, c_int],
    'mlffpc_peer_disable': [c_ptr],
    'mlffpc_allreduce_sum': [c_ptr, c_ptr, c_i64, c_ptr],
    'mlffpc_allgather': [c_ptr, c_ptr, c_ptr, c_i64, c_ptr],
    'mlffpc_geometry_workspace_bytes': [c_i64, c_int, c_int, ctypes.POINTER(c_i64)],
    'mlffpc_set_geometry': [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_i64, c_i64,
                            c_ptr, c_i64, c_ptr],
    'mlffpc_kernel_diag': [c_ptr, c_ptr, c_ptr],
    'mlffpc_kernel_assemble': [c_ptr, c_ptr, c_i64, c_ptr],
    'mlffpc_kernel_columns_workspace_bytes': [c_ptr, c_i64, ctypes.POINTER(c_i64)],
    'mlffpc_kernel_columns': [c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_ptr],
    'mlffpc_gemv': [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_dbl, c_dbl, c_i64, c_ptr],
    'mlffpc_symv_workspace_bytes': [c_i64, ctypes.POINTER(c_i64)],
    'mlffpc_symv': [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_dbl, c_dbl, c_ptr, c_i64, c_ptr],
    'mlffpc_set_option': [c_ptr, c_str, c_i64],
    'mlffpc_symop_storage_elems': [c_ptr, ctypes.POINTER(c_i64)],
    'mlffpc_symop_tiles': [c_ptr, c_ptr, c_i64, ctypes.POINTER(c_i64)],
    'mlffpc_symop_assemble': [c_ptr, c_ptr, c_ptr],
    'mlffpc_symop_workspace_bytes': [c_ptr, ctypes.POINTER(c_i64)],
    'mlffpc_symop_apply': [c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_dbl, c_ptr, c_i64, c_ptr, c_ptr],
    'mlffpc_matvec_free_workspace_bytes': [c_ptr, ctypes.POINTER(c_i64)],
    'mlffpc_matvec_free': [c_ptr, c_ptr, c_ptr, c_dbl, c_dbl, c_ptr, c_i64, c_ptr],
    'mlffpc_desc_from_r': [c_ptr, c_ptr, c_i64, c_int, c_ptr, c_ptr, c_ptr],
    'mlffpc_d_desc_dot_vec': [c_ptr, c_ptr, c_ptr, c_ptr],
    'mlffpc_predict_workspace_bytes': [c_ptr, c_i64, ctypes.POINTER(c_i64)],
    'mlffpc_predict': [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr],
    'mlffpc_dgemm': [c_ptr, c_int, c_i64, c_i64, c_i64, c_dbl, c_ptr, c_i64, c_ptr, c_i64, c_dbl, c_ptr,
                     c_i64, c_ptr],
    'mlffpc_syrk_rows': [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_dbl, c_ptr, c_i64, c_ptr],
    'mlffpc_potrf_lower': [c_ptr, c_ptr, c_i64, c_i64, ctypes.POINTER(c_int), c_ptr],
    'mlffpc_trsm_rows': [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr],
    'mlffpc_pchol_workspace_bytes': [c_ptr, c_i64, ctypes.POINTER(c_i64)],
    'mlffpc_pchol_build': [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr],
    'mlffpc_woodbury_factor': [c_ptr, c_ptr, c_i64, c_i64, c_dbl, c_ptr, c_ptr],
    'mlffpc_orthonormal_factor': [c_ptr, c_ptr, c_i64, c_i64, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr],
    'mlffpc_gram_defect': [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr],
    'mlffpc_projected_factor': [c_ptr, c_ptr, c_i64, c_i64, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    'mlffpc_precon_apply': [c_ptr, c_ptr, c_i64, c_i64, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    'mlffpc_pcg_workspace_bytes': [c_ptr, c_i64, c_int, ctypes.POINTER(c_i64)],
    'mlffpc_pcg': [c_ptr, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_i64,
                   c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr],
    'mlffpc_dot': [c_ptr, c_ptr, c_ptr, c_i64, ctypes.POINTER(c_dbl), c_ptr],
}
NON_INT_RETURNS = {'mlffpc_version': (c_int, []), 'mlffpc_last_error': (c_str, []),
                   'mlffpc_launch_count': (c_i64, [])}

_lib = None


class MlffpcError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MlffpcError(
            'libmlffpc.so is missing (%s). Build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `python mlff_preconditioner_b200/build.py`. There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    for name, (restype, argtypes) in NON_INT_RETURNS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def last_error():
    return load().mlffpc_last_error().decode('utf-8', 'replace')


def check(status):
    """Map a C status to the exception type the reference raises in the same situation
    (error conventions)."""
    if status == OK:
        return
    msg = last_error()
    if status == ERR_INVALID:
        raise ValueError(msg)
    if status == ERR_NOT_PSD:
        raise AssertionError(msg)  # incomplete_cholesky.py:62
    if status == ERR_LINALG:
        raise np.linalg.LinAlgError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise MlffpcError('libmlffpc status %d: %s' % (status, msg))


def nccl_library_path():
    """The libnccl.so.2 torch ships (so our communicator and torch's use the same NCCL build)."""
    try:
        import nvidia.nccl as _n

        cand = os.path.join(list(_n.__path__)[0], 'lib', 'libnccl.so.2')
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return 'libnccl.so.2'
