"""Timing of standalone corrupt-reads (config[1] shape: FASTQ in host memory -> corrupted FASTQ)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import mitty_b200.simulation.illumina as il
import mitty_b200.simulation.readgenerate as rg
from mitty_b200 import synth
from mitty_b200.engine import Engine, MODE_PHILOX
from mitty_b200.lib import vcfio
from mitty_b200.readmodels import load_model

n_mb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = synth.config1(contig_len=n_mb * 1000000)
m = load_model('hiseq-X-v2.5-Garvan.pkl'); rm = il.read_model_params(m, 30.0)
eng = Engine(0); eng.load_model(rm)
f1, f2 = [], []
for ri, region in enumerate(wl['regions']):
  r = vcfio.from_variant_table(wl['tables'][ri], region)
  rid = eng.load_region(np.ascontiguousarray(wl['contigs'][ri][1]), 0)
  for cpy, vl in enumerate(r['v']):
    cp = eng.build_copy(rid, vl)
    for ps in range(2):
      a, b, cnt, _, _ = rg.generate_unit(eng, il, rm, cp, region[0], cpy, 17 + ps, 'S', 0, ps, mode='philox')
      f1.append(a.copy()); f2.append(b.copy())
a1, a2 = np.concatenate(f1), np.concatenate(f2)
pairs = int((a1 == 10).sum() // 4)
eng.load_model(m)
for rep in range(3):
  eng.prof_reset()
  t0 = time.perf_counter(); o1, o2, n = eng.corrupt_fastq(a1, a2, mode=MODE_PHILOX, seed=5); t1 = time.perf_counter()
  p = eng.prof()
  print('corrupt-reads: %d pairs, %.1f MB in, wall %.1f ms (%.1f M pairs/s), k_corrupt %.2f ms (%.1f M pairs/s, %.1f GB/s read+write)' % (
    n, (a1.size + a2.size) / 1e6, 1e3 * (t1 - t0), n / (t1 - t0) / 1e6, p['emit_ms'], n / p['emit_ms'] / 1e3, 2 * (a1.size + a2.size) / p['emit_ms'] / 1e6))
