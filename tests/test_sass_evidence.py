"""CPU: the built library really contains the Blackwell-era instructions DESIGN.md claims (checked in SASS, where
the PTX names never appear): fp64 tensor-core MMAs (DMMA), the TMA tensor load of the symmetric operator
(UTMALDG) with its mbarrier traffic (SYNCS), and cp.async staging of the GEMM (LDGSTS).  Skipped when cuobjdump is
not installed."""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '..', 'mlff_preconditioner_b200', 'libmlffpc.so')


def _cuobjdump():
    for cand in (shutil.which('cuobjdump'), '/usr/local/cuda/bin/cuobjdump'):
        if cand and os.path.exists(cand):
            return cand
    return None


@pytest.mark.skipif(_cuobjdump() is None or not os.path.exists(LIB), reason='needs cuobjdump and the built library')
def test_sass_contains_dmma_tma_and_mbarrier():
    sass = subprocess.run([_cuobjdump(), '-sass', LIB], stdout=subprocess.PIPE, text=True).stdout
    assert 'sm_100a' in sass or 'SM100a' in sass or 'sm_100' in sass
    per_fn = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per_fn[cur] = set()
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            per_fn[cur].add(m.group(1))

    def fns_with(op):
        return [f for f, ops in per_fn.items() if op in ops]

    assert any('dgemm_kernel' in f for f in fns_with('DMMA')), 'fp64 tensor-core GEMM is missing'
    assert any('dgemm_kernel' in f for f in fns_with('LDGSTS')), 'cp.async staging is missing'
    tma = fns_with('UTMALDG')
    assert any('symv_tma_kernel' in f for f in tma), 'TMA tensor load of the symmetric operator is missing'
    assert any('symv_tma_kernel' in f for f in fns_with('SYNCS')), 'mbarrier pipeline is missing'
