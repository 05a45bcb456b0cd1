"""Host-side time breakdown of one bench step (build_copy / generate_unit / free_copy)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import bench
import mitty_b200.simulation.illumina as il
import mitty_b200.simulation.readgenerate as rg
from mitty_b200.engine import Engine
from mitty_b200.readmodels import load_model

class A: pass
args = A(); args.contig_len = int(sys.argv[1]) if len(sys.argv) > 1 else 249250621; args.seed = 7
model = load_model(bench.MODEL); rm = il.read_model_params(model, 30.0)
t0 = time.perf_counter(); wl = bench.make_chr1(args, 0); print('make_chr1 %.2fs' % (time.perf_counter() - t0))
from mitty_b200.lib import vcfio
r = vcfio.from_variant_table(wl['tables'][0], wl['regions'][0])
eng = Engine(0); eng.load_model(rm)
ref = np.ascontiguousarray(wl['contigs'][0][1])
def T(label, fn):
  torch.cuda.synchronize(); t = time.perf_counter(); out = fn(); torch.cuda.synchronize(); print('%-28s %8.2f ms' % (label, 1e3 * (time.perf_counter() - t))); return out
for rep in range(3):
  print('--- rep', rep)
  rid = T('load_region', lambda: eng.load_region(ref, 0))
  copies = [T('build_copy %d (%d var)' % (i, len(vl)), lambda vl=vl: eng.build_copy(rid, vl)) for i, vl in enumerate(r['v'])]
  for k in range(4):
    eng.prof_reset()
    out = T('generate_unit %d' % k, lambda: rg.generate_unit(eng, il, rm, copies[k // 2], '1', k // 2, 1234 + k, 'S', 0, k, mode='philox', corrupt=(rep == 2), fetch=False))
    print('    emit kernel %.2f ms, pairs %d' % (eng.prof()['emit_ms'], out[2]))
  for cp in copies: T('free_copy', lambda cp=cp: eng.free_copy(cp))
  T('free_region', lambda: eng.free_region(rid))
