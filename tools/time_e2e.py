"""e2e unit timing: sync vs async D2H, engine-pinned vs torch-pinned buffers."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import bench
import mitty_b200.simulation.illumina as il
import mitty_b200.simulation.readgenerate as rg
from mitty_b200.engine import Engine
from mitty_b200.readmodels import load_model
class A: pass
args = A(); args.contig_len = 249250621; args.seed = 7
model = load_model(bench.MODEL); rm = il.read_model_params(model, 30.0)
wl, region, r = bench.make_workload(args, 0)
eng = Engine(0); eng.load_model(rm)
ref = np.ascontiguousarray(wl['contigs'][0][1])
rid = eng.load_region(ref, 0)
copies = [eng.build_copy(rid, vl) for vl in r['v']]
est = int((args.contig_len * rm['p'] * 1.2) * (2 * 150 + 110)) + (1 << 20)
bufs = {'engine': [(eng.pinned(est), eng.pinned(est)) for _ in range(2)],
        'torch': [(torch.empty(est, dtype=torch.uint8).pin_memory().numpy(), torch.empty(est, dtype=torch.uint8).pin_memory().numpy()) for _ in range(2)]}
for kind in ('torch', 'engine', 'torch', 'engine'):
  for wait in (True, False):
    torch.cuda.synchronize(); t0 = time.perf_counter(); tot = 0
    for k in range(4):
      _, _, cnt, _, nb = rg.generate_unit(eng, il, rm, copies[k // 2], '1', k // 2, 100 + k, 'S', 0, k, mode='philox', corrupt=True, corrupt_seed=1,
                                          out=bufs[kind][k & 1], fetch=True, wait=wait)
      tot += 2 * nb
    eng.wait_copies(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('%-7s wait=%-5s 4 units: %.1f ms, %.1f GB/s D2H-equivalent' % (kind, wait, 1e3 * dt, tot / dt / 1e9))

print('--- bench-like e2e steps (load_region + builds + 4 units + frees)')
for kind in ('engine', 'torch'):
  for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rid_ = eng.load_region(ref, 0); t1 = time.perf_counter()
    cps = [eng.build_copy(rid_, vl) for vl in r['v']]; t2 = time.perf_counter()
    for k in range(4):
      rg.generate_unit(eng, il, rm, cps[k // 2], '1', k // 2, 100 + k, 'S', 0, k, mode='philox', corrupt=True, corrupt_seed=1, out=bufs[kind][k & 1], fetch=True, wait=False)
    eng.wait_copies(); t3 = time.perf_counter()
    for cp in cps: eng.free_copy(cp)
    eng.free_region(rid_); torch.cuda.synchronize(); t4 = time.perf_counter()
    print('%-7s load_region %.1f  builds %.1f  units %.1f  frees %.1f  total %.1f ms' % (kind, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t4-t0)))
