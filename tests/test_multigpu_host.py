"""The N > 1 path on the CPU: world_size-2 gloo process group, units sharded by LPT, ordered
concatenation.  The per-unit generator here is the oracle (no GPU in this container); what is under
test is the host logic: the sharded result must be byte-identical to the single-process run."""
import os

import numpy as np
import pytest

import oracle
from mitty_b200 import multigpu, synth
from tests import helpers as H


def test_lpt_assignment():
  w = [249, 243, 198, 191, 181, 171, 159, 146, 141, 135, 135, 134, 115, 107, 102, 90, 81, 78, 59, 63, 48, 51, 155, 59]
  for world in (1, 2, 4, 8):
    parts = multigpu.assign_units(w, world)
    assert sorted(k for p in parts for k in p) == list(range(len(w)))
    loads = [sum(w[k] for k in p) for p in parts]
    assert max(loads) <= sum(w) / world * 1.10 + max(w) * (world > 4)


def _worker(rank, world, port, tmp, want_sha):
  import torch.distributed as dist
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  wl = synth.edge_workload()
  regs = H.oracle_regions(H.workload_regions(wl))
  m = H.model('hiseq-X-v2.5-Garvan.pkl')
  rm = oracle.read_model_params(m, 30.0)
  units = [(ri, cpy) for ri, r in enumerate(regs) for cpy in range(len(r['v'])) for _ in range(rm['passes'])]
  seeds, order = oracle.unit_schedule(7, len(units))
  schedule = [(units[k], int(seeds[k])) for k in order]

  def gen(ps, unit):
    (ri, cpy), seed = unit
    r = regs[ri]
    p_min, p_max, _ = oracle.node_span(r['ref'], r['region'][1] + 1, r['v'][cpy])
    ts, te, fo = oracle.templates(rm['p'], rm['rlen'], rm['cum_tlen'], p_min, p_max, seed)
    a, b, _ = oracle.generate_unit(r['ref'], r['region'][1] + 1, r['v'][cpy], rm['rlen'], ts, te, fo,
                                   '{}:0:{}'.format(wl['sample'], ps), r['region'][0], cpy)
    return a, b

  weights = [regs[u[0][0]]['ref'].size for u in schedule]
  f1, f2 = os.path.join(tmp, 'out1.fq'), os.path.join(tmp, 'out2.fq')
  gathered = multigpu.run_sharded(schedule, weights, gen, f1, f2)
  if rank == 0:
    assert all(len(g) > 0 for g in gathered)            # both ranks did work
    assert H.sha256(open(f1, 'rb').read()) == want_sha[0] and H.sha256(open(f2, 'rb').read()) == want_sha[1]
    assert not [f for f in os.listdir(tmp) if '.part' in f]
  dist.destroy_process_group()


def test_sharded_generate_world2_gloo(tmp_path):
  import torch.multiprocessing as mp
  info = H.golden()['fastq']['edge']
  port = 29500 + os.getpid() % 2000
  mp.spawn(_worker, args=(2, port, str(tmp_path), (info['r1']['sha256'], info['r2']['sha256'])), nprocs=2, join=True)


def test_batch_runs_keep_schedule_order_and_bounds(monkeypatch):
  """Host logic of the batch path (readgenerate._runs / batchable): a worker's units, ascending, are cut into
  stretches of batchable units (bounded in reference bases and units) with the other units in between, so the
  worker still walks its list in schedule order (the sink's deadlock-freedom argument rests on that)."""
  import numpy as np
  import mitty_b200.simulation.readgenerate as rg
  from mitty_b200.lib.vcfio import VariantList

  def vl(ins=0):
    if not ins:
      return VariantList([], np.zeros(0, np.uint8), [], np.zeros(0, np.uint8), [0])
    return VariantList([5], np.frombuffer(b'I', np.uint8), [ins], np.frombuffer(b'A' * (ins + 1), np.uint8), [0, ins + 1])
  # regions: 0 small, 1 large, 2 small, 3 small but with an insertion that makes its haplotype too long, 4 small
  spans = [2000, 400000, 3000, 2500, 1000]
  vcf_df = [{'region': ('c', 0, s), 'v': [vl(), vl(60000 if k == 3 else 0)]} for k, s in enumerate(spans)]
  rm = {'p': 0.03, 'rlen': 150}
  ok = rg.batchable(vcf_df, rm, 'S', 100)
  assert ok == {(0, 0), (0, 1), (2, 0), (2, 1), (3, 0), (4, 0), (4, 1)}          # (1, *) too long, (3, 1) too long with its insertion
  assert rg.batchable(vcf_df, rm, 'a-sample-name-that-is-much-too-long', 100) == set()
  schedule = [{'region_idx': r, 'region_cpy': c, 'rng_seed': 1} for r in (0, 2, 1, 4, 3, 0, 2) for c in (0, 1)]
  mine = list(range(len(schedule)))
  runs = list(rg._runs(mine, schedule, ok, vcf_df))
  flat = [k for kind, x in runs for k in (x if kind == 'batch' else [x])]
  assert flat == mine                                                             # ascending schedule order survives
  kinds = [(kind, len(x) if kind == 'batch' else 1) for kind, x in runs]
  assert kinds == [('batch', 4), ('unit', 1), ('unit', 1), ('batch', 3), ('unit', 1), ('batch', 4)]
  # bounds: a new batch starts when the reference bases or the unit count would exceed the limits
  monkeypatch.setattr(rg, 'BATCH_MAX_BASES', 4000)
  runs = list(rg._runs([0, 1, 2, 3], schedule, ok, vcf_df))
  assert [(k, list(x)) for k, x in runs] == [('batch', [0, 1]), ('batch', [2, 3])]
  monkeypatch.setattr(rg, 'BATCH_MAX_BASES', 1 << 30)
  monkeypatch.setattr(rg, 'BATCH_MAX_UNITS', 2)
  runs = list(rg._runs([0, 1, 2, 3], schedule, ok, vcf_df))
  assert [len(x) for k, x in runs] == [2, 2]                                      # cut only where a new region starts (a soft cap)
