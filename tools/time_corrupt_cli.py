"""corrupt-reads command line on files (tmpfs): perfect FASTQ pair -> corrupted pair."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from click.testing import CliRunner
from mitty_b200 import synth
from mitty_b200.cli import cli
n_mb = int(sys.argv[1]) if len(sys.argv) > 1 else 50
d = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
wl = synth.chr1_shaped(seed=7, length=n_mb * 1000000, n_runs=max(3, n_mb // 6))
fa, vcf, bed = synth.write_workload(wl, os.path.join(d, 'w'))
r1, r2, c1, c2 = (os.path.join(d, x) for x in ('r1.fq', 'r2.fq', 'c1.fq', 'c2.fq'))
res = CliRunner().invoke(cli, ['generate-reads', fa, vcf, wl['sample'], bed, 'hiseq-X-v2.5-Garvan.pkl', '30', '7', r1, '--fastq2', r2], catch_exceptions=False)
assert res.exit_code == 0, res.output
for rep in range(2):
  t0 = time.perf_counter()
  res = CliRunner().invoke(cli, ['corrupt-reads', 'hiseq-X-v2.5-Garvan.pkl', r1, c1, '7', '--fastq2-in', r2, '--fastq2-out', c2], catch_exceptions=False)
  t1 = time.perf_counter()
  assert res.exit_code == 0, res.output
  sz = os.path.getsize(c1)
  pairs = sz / 377.0
  print('corrupt-reads: 2 x %.2f GB in %.2f s = %.2f GB/s in+out, ~%.1f M pairs/s' % (sz / 1e9, t1 - t0, 4 * sz / 1e9 / (t1 - t0), pairs / (t1 - t0) / 1e6))
