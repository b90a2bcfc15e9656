"""CPU, build container only (needs /root/reference): ``patch.install()`` swaps the reference's ``Iterative`` for
the device solver, and the replacement has the reference's constructor / ``solve`` parameter lists, so
``GDMLTrain.train`` (train.py:868-890) can call it unchanged.  Runs in a subprocess because loading the reference
installs compatibility shims into scipy."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import inspect, sys
sys.path.insert(0, %r)
from oracle import ref_shims
sgdml = ref_shims.load_reference()
import sgdml.solvers.iterative_solver as ref_mod
RefIterative = ref_mod.Iterative
from mlff_preconditioner_b200 import patch
from mlff_preconditioner_b200.solvers.iterative_solver import Iterative
assert patch.install() == ['sgdml.solvers.iterative_solver']
assert ref_mod.Iterative is Iterative and ref_mod._reference_Iterative is RefIterative
for name in ('__init__', 'solve', '_init_precon_operator', '_init_precon_operator_sb', '_lev_scores',
             '_init_kernel_operator'):
    ref_params = list(inspect.signature(getattr(RefIterative, name)).parameters)
    our_params = list(inspect.signature(getattr(Iterative, name)).parameters)
    assert our_params[:len(ref_params)] == ref_params, (name, ref_params, our_params)
ref_solve = inspect.signature(RefIterative.solve).parameters
our_solve = inspect.signature(Iterative.solve).parameters
for k in ref_solve:
    assert ref_solve[k].default == our_solve[k].default or ref_solve[k].default is inspect._empty, k
# the trainer looks the class up through the module at call time (train.py:868)
import sgdml.train as ref_train
assert ref_train.iterative_solver.Iterative is Iterative
patch.uninstall()
assert ref_mod.Iterative is RefIterative
print('DROPIN_OK')
''' % ROOT


@pytest.mark.skipif(not os.path.isdir('/root/reference/src/sGDML/sgdml'), reason='the reference is only mounted in the build container')
def test_patch_install_swaps_the_reference_solver():
    out = subprocess.run([sys.executable, '-c', SCRIPT], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         cwd=ROOT, timeout=280)
    assert out.returncode == 0 and 'DROPIN_OK' in out.stdout, out.stderr[-3000:]
