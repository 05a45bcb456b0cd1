"""Per-region cost of an exome-style BED (many small regions) through process_multi_threaded.
python tools/time_many_regions.py [regions] [width] [workers per GPU, comma separated list to try; a trailing 'u' = unit by unit
instead of the batch path, e.g. 1,1u,4u]"""
import json, logging, os, sys, tempfile, shutil, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mitty_b200.simulation.illumina as il
import mitty_b200.simulation.readgenerate as rg
from mitty_b200 import synth
from mitty_b200.readmodels import load_model
logging.basicConfig(level=logging.ERROR)
logging.getLogger('mitty_b200.simulation.readgenerate').setLevel(logging.INFO)
n_reg = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
width = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
tries = [(int(x.rstrip('u')), not x.endswith('u')) for x in (sys.argv[3] if len(sys.argv) > 3 else '1,1u,4u').split(',')]
wl = synth.config1(contig_len=n_reg * width * 2 + 10000, names=('1',))
wl['regions'] = [('1', 5000 + 2 * width * k, 5000 + 2 * width * k + width) for k in range(n_reg)]
d = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
try:
  fa, vcf, bed = synth.write_workload(wl, os.path.join(d, 'w'))
  r1, r2 = os.path.join(d, 'r1.fq'), os.path.join(d, 'r2.fq')
  m = load_model('hiseq-X-v2.5-Garvan.pkl')
  for w, batch in tries:
    for rep in range(2):      # the second run is the warm one
      t_rep = time.time()
      rg.process_multi_threaded(fa, vcf, wl['sample'], bed, il, m, 30.0, r1, r2, threads=1, seed=7, mode='philox', corrupt=True, workers_per_gpu=w, drop_end_deletions=True,
                                batch_small=batch)
    st = rg.last_run
    st['wall'] = time.time() - t_rep
    print(json.dumps({'regions': n_reg, 'width': width, 'workers_per_gpu': w, 'batch_path': batch, 'batches': st['batches'], 'input_seconds': st['input_seconds'], 'seconds_units_to_files': st['seconds'], 'ms_per_region': 1e3 * st['seconds'] / n_reg,
                      'wall_seconds': st['wall'], 'pairs': st['templates'], 'pairs_per_s': st['templates'] / st['seconds']}))
finally:
  shutil.rmtree(d, ignore_errors=True)
