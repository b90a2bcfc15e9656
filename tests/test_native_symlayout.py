"""CPU: compile (nvcc, host code only) and run the brute-force check of csrc/symlayout.cuh -- the unit enumeration the
persistent TMA kernels, the tile plan and the packed assembly all share, and the cursor with which the one-launch pass
walks the concatenated unit lists of a rank's tiles (every unit once and in order for 1 ... 296 CTAs, at most two CTAs per
strip, distinct row-sum slots)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', shutil.which('nvcc')):
        if cand and os.path.exists(cand):
            return cand
    return None


@pytest.mark.skipif(_nvcc() is None, reason='needs nvcc')
def test_symlayout_closed_forms(tmp_path):
    exe = str(tmp_path / 'symlayout_check')
    subprocess.check_call([_nvcc(), '-std=c++17', '-O1', '-o', exe, os.path.join(HERE, 'native', 'symlayout_check.cu')])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and 'SYMLAYOUT OK' in out.stdout, out.stdout
