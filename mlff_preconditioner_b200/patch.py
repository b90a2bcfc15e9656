"""Drop the device solver into the reference: replace ``Iterative`` in the reference's module(s), so
``scripts/cluster_main.py`` -> ``create_data.cg_steps`` -> ``GDMLTrain.train(solver='cg')`` runs the solve
step on the GPU while everything above it (task creation, descriptors, result pickles) stays as is.

The reference is imported twice under different names in its own drivers (``src.sGDML.sgdml.*`` via
``create_data.py:8`` and top-level ``sgdml.*``), so both copies are patched when
present.
"""
import sys

from .solvers.iterative_solver import Iterative

_TARGETS = ('sgdml.solvers.iterative_solver', 'src.sGDML.sgdml.solvers.iterative_solver')


def install(modules=None):
    """Patch every already-imported copy of the reference solver module; returns the patched names."""
    patched = []
    for name in (modules or _TARGETS):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, 'Iterative', None) is not Iterative:
            mod._reference_Iterative = mod.Iterative
            mod.Iterative = Iterative
            patched.append(name)
    return patched


def uninstall():
    for name in _TARGETS:
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, '_reference_Iterative'):
            mod.Iterative = mod._reference_Iterative
            del mod._reference_Iterative
