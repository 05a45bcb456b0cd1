"""Full command line to a file sink: FASTA / VCF / BED on disk -> two FASTQ files (tmpfs)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from click.testing import CliRunner
from mitty_b200 import synth
from mitty_b200.cli import cli
arg = sys.argv[1] if len(sys.argv) > 1 else '50'
wgs = arg.startswith('wgs:')
n_mb = 0 if wgs else int(arg)
threads = sys.argv[2] if len(sys.argv) > 2 else '1'
d = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
wl = synth.grch37_shaped(scale=float(arg[4:])) if wgs else synth.chr1_shaped(seed=7, length=n_mb * 1000000, n_runs=max(3, n_mb // 6))
t0 = time.perf_counter(); fa, vcf, bed = synth.write_workload(wl, os.path.join(d, 'w')); t1 = time.perf_counter()
print('wrote inputs (%d Mb FASTA, %d VCF records) in %.1f s' % (sum(len(c[1]) for c in wl['contigs']) // 1000000, sum(len(t) for t in wl['tables']), t1 - t0))
for extra in ([], ['--corrupt']):
  r1, r2 = os.path.join(d, 'r1.fq'), os.path.join(d, 'r2.fq')
  t0 = time.perf_counter()
  res = CliRunner().invoke(cli, ['-v', '4', 'generate-reads', fa, vcf, wl['sample'], bed, 'hiseq-X-v2.5-Garvan.pkl', '30', '7', r1, '--fastq2', r2, '--threads', threads] + extra, catch_exceptions=False)
  t1 = time.perf_counter()
  assert res.exit_code == 0, res.output
  print('\n'.join(l for l in res.output.split('\n') if 'Finished' in l or 'phase' in l))
  sz = os.path.getsize(r1)
  import numpy as np
  nl = 0
  with open(r1, 'rb') as fp:
    while True:
      b = fp.read(1 << 28)
      if not b:
        break
      nl += int(np.count_nonzero(np.frombuffer(b, dtype=np.uint8) == 10))
  pairs = nl // 4
  print('generate-reads %s --threads %s: %d pairs, 2 x %.2f GB to %s in %.2f s = %.2f M pairs/s' % (' '.join(extra), threads, pairs, sz / 1e9, d, t1 - t0, pairs / (t1 - t0) / 1e6))
  os.remove(r1); os.remove(r2)
