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
