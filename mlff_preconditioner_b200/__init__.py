"""mlff_preconditioner_b200 -- B200-native fp64 preconditioned-CG solve step for sGDML Hessian-kernel
systems: a drop-in for ``Iterative.solve`` of bluecher31/mlff-preconditioner, backed by hand-written
sm_100a CUDA behind a C ABI (include/mlffpc.h).  No CPU fallback: without libmlffpc.so and a CUDA
device the solver entry points raise.
"""
__version__ = '0.1.0'

DONE = 1  # callback protocol of the reference (sgdml/__init__.py:31-32)
NOT_DONE = 0

from . import _lib  # noqa: E402,F401
