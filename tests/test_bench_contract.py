"""bench.py's output contract, checked on the CPU leg (the GPU leg needs a B200): one JSON line, the keys the driver
reads, the reference arm's bookkeeping."""
import json
import os
import subprocess
import sys

import helpers as H

KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data',
        'config', 'e2e', 'cpu_baseline', 'impl', 'gpu_launches'}


def test_reference_arm_prints_one_json_line():
  env = dict(os.environ, MOE_BENCH_CPU_SAMPLE='48')
  r = subprocess.run([sys.executable, os.path.join(H.ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1', '--gpus', '1'],
                     capture_output=True, text=True, timeout=300, env=env)
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert KEYS <= set(d)
  assert d['impl'] == 'reference' and d['unit'] == 'MPix/s' and d['higher_is_better'] is True and d['vs_baseline'] is None
  assert d['value'] > 0 and d['e2e'] == {'value': d['value'], 'unit': 'MPix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
  assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and 'workload' in d['config']


def test_reference_arm_is_silent_on_other_ranks():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1', MOE_BENCH_CPU_SAMPLE='48')
  r = subprocess.run([sys.executable, os.path.join(H.ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1', '--gpus', '2'],
                     capture_output=True, text=True, timeout=120, env=env)
  assert r.returncode == 0 and r.stdout.strip() == ''


def test_native_arm_refuses_to_run_without_a_gpu():
  import torch
  if torch.cuda.is_available():
    return
  r = subprocess.run([sys.executable, os.path.join(H.ROOT, 'bench.py'), '--steps', '1'], capture_output=True, text=True, timeout=120)
  assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
