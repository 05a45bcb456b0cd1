"""bench.py's reference arm (the unmodified reference from baseline/_ref when it is installed, the C
port otherwise, on the host cores) runs without a GPU: check the one-line JSON contract the driver
parses -- keys, types, exactly one stdout line."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def test_reference_arm_prints_one_contract_line():
  res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '0',
                        '--contig-len', '20000000'], capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert res.returncode == 0, res.stderr[-2000:]
  lines = [l for l in res.stdout.split('\n') if l.strip()]
  assert len(lines) == 1, res.stdout
  d = json.loads(lines[0])
  assert d['impl'] == 'reference' and d['unit'] == 'pairs/s' and d['higher_is_better'] is True
  for key in ('metric', 'value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
    assert key in d, key
  assert d['value'] > 0 and d['steps'] == 1 and d['n_gpus'] == 1 and d['vs_baseline'] is None
  have_ref = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'mitty'))
  assert d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
  if have_ref:                                      # the C port stays as a second figure: far faster than the Python reference
    assert d['cpu_baseline']['port']['kind'] == 'port' and d['cpu_baseline']['port']['value'] > 10 * d['value']
    assert 'unmodified reference' in d['cpu_baseline']['sample']
  assert d['e2e'] == {'value': d['value'], 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
  assert 'workload' in d['config'] and 'model' not in d['config']
