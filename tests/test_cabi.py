"""CPU: the C-ABI library builds, loads, and exports exactly the symbols include/mlffpc.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'mlffpc.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mlffpc_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_header_symbols():
    import __graft_entry__ as g

    g.build()
    from mlff_preconditioner_b200 import _lib

    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), 'missing export: %s' % s
    bound = set(_lib.SIGNATURES) | set(_lib.NON_INT_RETURNS)
    assert bound == set(syms), (bound ^ set(syms))
    assert lib.mlffpc_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to run instead of falling back to the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import numpy as np
    from mlff_preconditioner_b200 import _lib
    from mlff_preconditioner_b200.engine import Engine

    with pytest.raises(_lib.MlffpcError):
        Engine(np.ones((2, 3)), np.ones((2, 3, 3)), np.arange(3), 10)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, 'mlff_preconditioner_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', txt, flags=re.M), os.path.join(dirpath, f)
