"""CPU: the benchmark's reference arm (the only arm that runs without a GPU) prints exactly one JSON line with
the keys the driver reads, on a workload small enough for the CPU suite."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'small',
                          '--steps', '1', '--warmup', '0'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         cwd=ROOT, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'pcg_time_to_solution' and d['unit'] == 's'
    assert d['higher_is_better'] is False and d['value'] > 0 and d['dtype'] == 'f64'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_ours_arm_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--workload', 'small', '--steps', '1',
                          '--warmup', '0'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, timeout=280)
    assert out.returncode != 0          # no CPU fallback: the product arm must not produce a number without a GPU
    assert out.stdout.strip() == ''


def test_clock_sampler_reports_only_the_timed_region():
    """The sampler is started before the warm-up steps (nvidia-smi's start-up must not land in a timed step); samples
    taken before mark() are dropped, throttle reasons and the busy-clock median come from the rest."""
    sys.path.insert(0, ROOT)
    import bench

    class _Proc(object):
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    s = bench.ClockSampler(0)
    s.proc = _Proc()
    s.samples = [(10.0, '900, 1965, Not Active, Active, Not Active, Not Active'),      # warm-up: thermal flag ignored
                 (20.0, '1600, 1965, Not Active, Not Active, Not Active, Active'),
                 (21.0, '1700, 1965, Not Active, Not Active, Not Active, Active'),
                 (22.0, '0, 1965, Not Active, Not Active, Not Active, Not Active'),      # idle sample: not in the median
                 (23.0, 'garbage')]
    s.t_mark = 15.0
    out = s.stop()
    assert out['reasons'] == ['sw_power_cap'] and out['sm_mhz'] == 1650.0 and out['sm_max_mhz'] == 1965.0 and out['samples'] == 3
