"""Per-region overhead: a BED of many small regions (exome-like) through the command line."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from click.testing import CliRunner
from mitty_b200 import synth
from mitty_b200.cli import cli
n_reg = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
width = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
wl = synth.config1(contig_len=n_reg * width * 2 + 10000, names=('1',))
wl['regions'] = [('1', 5000 + 2 * width * k, 5000 + 2 * width * k + width) for k in range(n_reg)]
d = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
fa, vcf, bed = synth.write_workload(wl, os.path.join(d, 'w'))
r1, r2 = os.path.join(d, 'r1.fq'), os.path.join(d, 'r2.fq')
devs = sys.argv[3] if len(sys.argv) > 3 else None
for extra in ([], ['--corrupt']):
  if devs:
    extra = extra + ['--devices', devs]
  t0 = time.perf_counter()
  res = CliRunner().invoke(cli, ['-v', '2', 'generate-reads', fa, vcf, wl['sample'], bed, 'hiseq-X-v2.5-Garvan.pkl', '30', '7', r1, '--fastq2', r2] + extra, catch_exceptions=False)
  t1 = time.perf_counter()
  assert res.exit_code == 0, res.output
  print('\n'.join(l for l in res.output.split('\n') if 'GPU ' in l))
  pairs = sum(1 for _ in open(r1, 'rb')) // 4
  print('%d regions of %d bp %s: %.2f s = %.2f ms per region, %d pairs (%.0f pairs/s)' % (n_reg, width, ' '.join(extra), t1 - t0, 1e3 * (t1 - t0) / n_reg, pairs, pairs / (t1 - t0)))
