"""north_star's target run through the PRODUCT path: `generate-reads --corrupt --threads N` (readgenerate.
process_multi_threaded: N GPU worker threads, units pulled in schedule order, native sink) on the
GRCh37-shaped genome, FASTA / VCF / BED read from files, two FASTQ files written to a tmpfs directory.

    python tools/time_cli_wgs.py [threads] [scale|auto] [gzip level] [target dir]

scale 'auto': the largest genome scale whose FASTQ fits the box's memory (228 GB at scale 1)."""
import json, logging, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mitty_b200.simulation.illumina as il
import mitty_b200.simulation.readgenerate as rg
from mitty_b200 import synth
from mitty_b200.readmodels import load_model

threads = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scale = sys.argv[2] if len(sys.argv) > 2 else 'auto'
gz = int(sys.argv[3]) if len(sys.argv) > 3 else 0
target = sys.argv[4] if len(sys.argv) > 4 else '/dev/shm'
avail = 0
for ln in open('/proc/meminfo'):
  if ln.startswith('MemAvailable:'):
    avail = int(ln.split()[1]) * 1024
if scale == 'auto':
  scale = max(0.02, min(1.0, (avail - (70 << 30)) / (1.2 * 228e9 * (0.25 if gz else 1.0))))
scale = float(scale)
d = tempfile.mkdtemp(dir=target)
try:
  t0 = time.perf_counter()
  wl = synth.grch37_shaped(scale=scale, seed=7)
  fa, vcf, bed = synth.write_workload(wl, os.path.join(d, 'w'))
  t1 = time.perf_counter()
  ext = '.fq.gz' if gz else '.fq'
  r1, r2 = os.path.join(d, 'r1' + ext), os.path.join(d, 'r2' + ext)
  logging.basicConfig(level=logging.INFO)
  rg.process_multi_threaded(fa, vcf, wl['sample'], bed, il, load_model('hiseq-X-v2.5-Garvan.pkl'), 30.0, r1, r2, threads=threads, seed=7,
                            mode='philox', corrupt=True, gzip_level=gz)
  t2 = time.perf_counter()
  st = rg.last_run
  out = {'run': 'generate-reads --corrupt --threads {} (configs[3] at scale {:.3f}) -> {}'.format(threads, scale, target + ('  gzip level %d' % gz if gz else '')),
         'pairs': st['templates'], 'bytes_file1': os.path.getsize(r1), 'bytes_file2': os.path.getsize(r2),
         'seconds_units_to_files': st['seconds'], 'seconds_inputs_parsed': st.get('parse_seconds'), 'seconds_before_first_unit': st['input_seconds'], 'seconds_total': t2 - t1,
         'pairs_per_s': st['templates'] / st['seconds'], 'pairs_per_min': 60.0 * st['templates'] / st['seconds'],
         'pairs_per_min_incl_input_parsing': 60.0 * st['templates'] / (t2 - t1),
         'gbs_written': (os.path.getsize(r1) + os.path.getsize(r2)) / st['seconds'] / 1e9,
         'gpus': st['gpus'], 'writer_threads': st['writers'], 'cores': os.cpu_count(), 'mem_available_gb': avail / 2**30,
         'inputs_written_s': t1 - t0}
  print(json.dumps(out))
finally:
  shutil.rmtree(d, ignore_errors=True)
