"""bench.py -- read pairs/s of the read-generation hot path (generate-reads + Illumina corruption).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[2] -- one chr1-shaped synthetic contig
(249,250,621 bp, ~10 % N in long runs, ~330 k SNP/indel records, diploid), 30x, 2x150
(hiseq-X-v2.5-Garvan model, the shipped 150-bp model), production (Philox) mode with
the corruption model fused into the emit kernel.  One STEP = the whole contig at 30x: both
chromosome copies built from the resident packed reference, then all 4 work units
(2 copies x 2 passes, ~25 M read pairs, ~18 GB of FASTQ).  With N GPUs every rank runs its own
chr1-shaped contig (different seed): units are independent, there is no collective on the data
path ("weak" scaling); torch.distributed is used only for the barrier and the max/sum of timings.

value   pairs/s with inputs resident in HBM (packed reference, variant arrays on the host side of
        mg_copy_build), outputs left in HBM; timed with CUDA events on the launch stream.
e2e     the same step through the C ABI with HOST buffers: raw reference bytes H2D + packing,
        copy builds, units, and every FASTQ byte D2H into pinned memory, inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = 'hiseq-X-v2.5-Garvan.pkl'
COVERAGE = 30.0
SLICE = 10000000  # CPU-baseline sample: one 10 Mb slice of the same contig


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--contig-len', type=int, default=249250621)
  ap.add_argument('--seed', type=int, default=7)
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--perfect', action='store_true', help='perfect reads only (no fused corruption)')
  ap.add_argument('--workload', default='chr1', choices=['chr1', 'wgs'], help="'wgs': BASELINE.json configs[3], GRCh37-shaped genome sharded by contig over the ranks (strong scaling)")
  ap.add_argument('--scale', type=float, default=1.0, help='length scale of the wgs workload')
  ap.add_argument('--soft-masked', action='store_true', help='chr1 workload with half of the bases in lower-case stretches (not the headline configuration)')
  return ap.parse_args()


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
  Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

  def __init__(self, index):
    self.index, self.rows, self.p = index, [], None

  def start(self):
    try:
      self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '25'],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
    except Exception:
      self.p = None

  def _read(self):
    for line in self.p.stdout:
      self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))

  def mark(self):
    """start of the timed region (the sampler itself is started before the warm-up: nvidia-smi takes
    longer to start than a short timed region lasts)"""
    self.t_mark = time.perf_counter()

  def stop(self):
    if self.p is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.p.terminate()
    try:
      self.p.wait(timeout=5)
    except Exception:
      self.p.kill()
    sm, mx, reasons = [], [], set()
    rows = [r for t, r in self.rows if t >= getattr(self, 't_mark', 0.0)]
    window = 'timed region'
    if not rows:   # no sample fell into a very short timed region: use the warm-up samples (same load)
      rows, window = [r for t, r in self.rows], 'warm-up + timed region'
    for r in rows:
      try:
        sm.append(float(r[1])); mx.append(float(r[2]))
        for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
          if v.lower().startswith('active'):
            reasons.add(name)
      except Exception:
        pass
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


def measured_peak():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    try:
      return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)'
    except Exception:
      pass
  return 6650.0, 'fallback (B200_PROFILING.md)'


def make_workload(args, rank):
  from mitty_b200 import synth
  from mitty_b200.lib import vcfio
  length = args.contig_len
  scale = length / 249250621.0
  wl = synth.chr1_shaped(seed=args.seed + 101 * rank, length=length, n_runs=max(3, int(39 * scale)))
  if getattr(args, 'soft_masked', False):        # repeat-masker style: alternating stretches, mean length 300
    seq = wl['contigs'][0][1]
    edges = np.cumsum(np.random.RandomState(5).geometric(1.0 / 300.0, size=2 * length // 300 + 64))
    edges = edges[edges < length]
    low = (np.searchsorted(edges, np.arange(length), side='right') & 1).astype(bool) & (seq != ord('N'))
    seq[low] |= 0x20
  region = wl['regions'][0]
  r = vcfio.from_variant_table(wl['tables'][0], region)
  return wl, region, r


# ---- CPU baseline (the oracle: plain-C port of the reference algorithm) ------------------------------

def _cpu_unit(job):
  """One oracle work unit on a slice: generate-reads + corrupt-reads (single thread)."""
  import oracle
  ref, start, vl_arrays, rm_small, seed, cum_bq = job
  cv = oracle.CopyVariants(*vl_arrays)
  p_min, p_max, _ = oracle.node_span(ref, start + 1, cv)
  t0 = time.perf_counter()
  ts, te, fo = oracle.templates(rm_small['p'], rm_small['rlen'], rm_small['cum_tlen'], p_min, p_max, seed)
  f1, f2, n = oracle.generate_unit(ref, start + 1, cv, rm_small['rlen'], ts, te, fo, 'S:0:0', '1', 0)
  c1, c2, _ = oracle.corrupt_fastq(cum_bq, seed, f1, f2)
  return n, time.perf_counter() - t0


def cpu_jobs(args, wl, n_jobs):
  import mitty_b200.simulation.illumina as il
  from mitty_b200.lib import vcfio
  from mitty_b200.readmodels import load_model
  model = load_model(MODEL)
  rm = il.read_model_params(model, COVERAGE)
  rm_small = {'p': rm['p'], 'rlen': int(rm['rlen']), 'cum_tlen': np.asarray(rm['cum_tlen'])}
  seq = wl['contigs'][0][1]
  length = seq.shape[0]
  sl = min(SLICE, length)
  jobs = []
  for k in range(n_jobs):
    # slices are taken from the non-N part of the contig, round robin
    start = int((length // 3 + k * sl) % max(1, length - sl))
    region = ('1', start, start + sl)
    r = vcfio.from_variant_table(wl['tables'][0], region)
    vl = r['v'][k % 2]
    # keep variants fully inside the slice (no deletion across the slice end)
    keep = (vl.pos > start + 1000) & (vl.pos < start + sl - 1000)
    idx = np.flatnonzero(keep)
    alts = [vl.alt_pool[vl.alt_off[i]:vl.alt_off[i + 1]].tobytes().decode() for i in idx]
    jobs.append((np.ascontiguousarray(seq[start:start + sl]), start, (vl.pos[idx], vl.op[idx], vl.oplen[idx], alts), rm_small,
                 1000 + k, np.asarray(model['cum_bq_mat'])))
  return jobs, sl


def cpu_baseline(args, wl, procs):
  import multiprocessing as mp
  jobs, sl = cpu_jobs(args, wl, procs)
  t0 = time.perf_counter()
  if procs == 1:
    res = [_cpu_unit(jobs[0])]
  else:
    with mp.get_context('fork').Pool(procs) as pool:
      res = pool.map(_cpu_unit, jobs)
  wall = time.perf_counter() - t0
  pairs = sum(r[0] for r in res)
  return pairs, wall, sl


def run_reference(args):
  """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is pure
  Python and cannot travel to the GPU box) on all host cores, on bounded samples of the workload."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import oracle
  oracle.build()
  cores = os.cpu_count() or 1
  procs = max(1, min(cores, 64))
  wl, region, r = make_workload(args, 0)
  times, pairs = [], 0
  for s in range(args.warmup + args.steps):
    n, wall, sl = cpu_baseline(args, wl, procs)
    if s >= args.warmup:
      times.append(wall); pairs += n
  value = pairs / sum(times)
  sample = '{} processes x one {} Mb-slice work unit (generate + corrupt, ~{} pairs) per step'.format(procs, sl // 1000000, pairs // max(1, args.steps))
  line = {'impl': 'reference', 'metric': 'read pairs/sec (2x150, FASTQ-formatted, corrupted)', 'value': value, 'unit': 'pairs/s',
          'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(times) / len(times),
          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
          'config': config_dict(args),
          'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': procs, 'kind': 'port', 'sample': sample},
          'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
  emit(line)


def config_dict(args):
  if getattr(args, 'workload', 'chr1') == 'wgs':
    return {'workload': 'configs[3]: GRCh37-shaped synthetic genome (24 contigs, lengths x {}, ~8% N, diploid autosomes, haploid X/Y), '
                        '30x paired 2x150, Philox mode{}, contigs dealt to the ranks by LPT'.format(args.scale, '' if args.perfect else ' + fused Illumina corruption'),
            'read_model': MODEL + ' (mean_rlen 150)', 'coverage': COVERAGE, 'seed': args.seed,
            'parallelism': 'contigs sharded over GPUs, units independent, no collective',
            'l2': 'every unit streams its FASTQ through L2 (>> 126 MB for the large contigs)'}
  return {'workload': 'configs[2]: chr1-shaped synthetic contig ({} bp, ~10% N, GIAB-density diploid VCF), 30x paired 2x150, '
                      'Philox mode{}{}'.format(args.contig_len, '' if args.perfect else ' + fused Illumina corruption',
                                               ', SOFT-MASKED variant (half of the bases lower case)' if getattr(args, 'soft_masked', False) else ''),
          'read_model': MODEL + ' (mean_rlen 150)', 'coverage': COVERAGE, 'units_per_step': 4, 'seed': args.seed,
          'parallelism': 'one contig per GPU, units independent, no collective',
          'l2': 'each unit streams ~4.6 GB of FASTQ through L2 (>> 126 MB), so nothing is re-read warm between timed units'}


# ---- the engine --------------------------------------------------------------------------------------

_REAL_STDOUT = None


def _guard_stdout():
  """The contract is ONE JSON line on stdout: anything a library prints there (NCCL's version
  banner, for one) is sent to stderr instead; emit() writes to the real stdout."""
  global _REAL_STDOUT
  if _REAL_STDOUT is None:
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
  data = (json.dumps(line) + '\n').encode()
  if _REAL_STDOUT is None:
    sys.stdout.write(data.decode()); sys.stdout.flush()
  else:
    os.write(_REAL_STDOUT, data)


def main():
  args = parse()
  _guard_stdout()
  if args.impl == 'reference':
    return run_reference(args)

  import torch
  import torch.distributed as dist
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  torch.cuda.set_device(local)
  from mitty_b200.engine import bind_host_thread_to_gpu
  numa_cores = bind_host_thread_to_gpu(local) if world > 1 else None   # pinned buffers next to the GPU
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))

  import mitty_b200.simulation.illumina as il
  import mitty_b200.simulation.readgenerate as rg
  from mitty_b200.engine import Engine
  from mitty_b200.readmodels import load_model

  model = load_model(MODEL)
  rm = il.read_model_params(model, COVERAGE)
  L = int(rm['rlen'])
  stream = torch.cuda.Stream()
  eng = Engine(local, stream=stream.cuda_stream)
  eng.load_model(rm)
  corrupt = not args.perfect
  # work items of this rank: (region, per-copy variants, pinned reference bytes)
  if args.workload == 'wgs':
    from mitty_b200 import multigpu, synth
    from mitty_b200.lib import vcfio
    mine = multigpu.assign_units([int(max(20000, n * args.scale)) for _, n in synth.GRCH37_CONTIGS], world)[rank]   # LPT by contig length
    gw = synth.grch37_shaped(scale=args.scale, seed=args.seed, only=set(mine))
    wl = gw
    items = []
    for (name, seq), vt, region in zip(gw['contigs'], gw['tables'], gw['regions']):
      items.append((region, vcfio.from_variant_table(vt, region), torch.from_numpy(np.ascontiguousarray(seq)).pin_memory().numpy()))
    max_len = max(it[0][2] for it in items)
  else:
    wl, region, r = make_workload(args, rank)
    items = [(region, r, torch.from_numpy(np.ascontiguousarray(wl['contigs'][0][1])).pin_memory().numpy())]
    max_len = args.contig_len

  def step(seed, rids, out=None, fetch=False):
    """Every item of this rank: its copies built from the (resident or just loaded) region, then
    copies x passes work units.  Returns (pairs, fastq bytes).  out: two pinned buffer pairs used
    alternately; the D2H of unit k overlaps the kernels of unit k+1."""
    pairs = nbytes = 0
    k = 0
    for (region_, r_, ref_), rid in zip(items, rids):
      if rid is None:
        rid = eng.load_region(ref_, region_[1])
      copies = [eng.build_copy(rid, vl) for vl in r_['v']]
      for cpy in range(len(copies)):
        for ps in range(rm['passes']):
          _, _, cnt, _, nb = rg.generate_unit(eng, il, rm, copies[cpy], region_[0], cpy, (seed * 7919 + k * 104729) & 0xFFFFFFFF,
                                              wl['sample'], 0, k, mode='philox', corrupt=corrupt, corrupt_seed=seed,
                                              out=out[k & 1] if out else None, fetch=fetch, wait=not fetch)
          pairs += cnt; nbytes += 2 * nb; k += 1
      if fetch:
        eng.wait_copies()
      for cp in copies:
        eng.free_copy(cp)
      if fetch:
        eng.free_region(rid)
    return pairs, nbytes

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  # ---- value: inputs resident in HBM, outputs stay in HBM
  rids = [eng.load_region(ref_, region_[1]) for region_, _, ref_ in items]
  clocks = ClockSampler(local); clocks.start()
  for w in range(args.warmup):
    step(1000 + w, rids)
  barrier()
  eng.prof_reset()
  clocks.mark()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  pairs = nbytes = 0
  with torch.cuda.stream(stream):
    ev0.record(stream)
    for s in range(args.steps):
      p_, b_ = step(2000 + s, rids)
      pairs += p_; nbytes += b_
    ev1.record(stream)
  barrier()
  ms = ev0.elapsed_time(ev1)
  clk = clocks.stop()
  prof = eng.prof()
  for rid in rids:
    eng.free_region(rid)

  # ---- e2e: host buffers in, host buffers out
  e2e = None
  if not args.no_e2e:
    est = int((max_len * rm['p'] * 1.2) * (2 * L + 110)) + (1 << 20)
    out = [(eng.pinned(est), eng.pinned(est)) for _ in range(2)]
    def e2e_step(seed):
      return step(seed, [None] * len(items), out=out, fetch=True)
    for w in range(min(args.warmup, 3)):
      e2e_step(3000 + w)
    barrier()
    t0 = time.perf_counter()
    ep = eb = 0
    for s in range(args.steps):
      p_, b_ = e2e_step(4000 + s)
      ep += p_; eb += b_
    torch.cuda.synchronize()
    e_wall = time.perf_counter() - t0
    h2d = sum(ref_.nbytes + sum(v.pos.nbytes + v.op.nbytes + v.oplen.nbytes + v.alt_pool.nbytes + v.alt_off.nbytes for v in r_['v']) for _, r_, ref_ in items)
    e2e = [ep, e_wall, h2d, eb / max(1, args.steps)]

  # ---- aggregate over ranks: max time, summed work
  if world > 1:
    t = torch.tensor([ms, e2e[1] if e2e else 0.0], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    w_ = torch.tensor([pairs, e2e[0] if e2e else 0, nbytes], dtype=torch.float64, device='cuda')
    dist.all_reduce(w_, op=dist.ReduceOp.SUM)
    ms, e_wall_max = float(t[0]), float(t[1])
    pairs_all, e_pairs_all = float(w_[0]), float(w_[1])
  else:
    pairs_all, e_pairs_all, e_wall_max = float(pairs), float(e2e[0]) if e2e else 0.0, e2e[1] if e2e else 0.0

  if rank == 0:
    peak, peak_src = measured_peak()
    # dominant kernel: k_unit_emit.  Algorithmic bytes per launch = FASTQ bytes written (both files)
    # + 2 * ceil(L/4) haplotype bytes read per pair, over the CUDA-event time of the
    # emit launches (events recorded around each launch on the launch stream inside the library).
    alg = nbytes + pairs * 2 * ((L + 3) // 4)
    ach = alg / (prof['emit_ms'] * 1e-3) / 1e9 if prof['emit_ms'] > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')   # dram bytes per launch of k_unit_emit from the committed ncu --set full capture
    if os.path.exists(tp):
      try:
        traffic = json.load(open(tp)).get('corrupt' if corrupt else 'perfect', {}).get(str(args.contig_len))
      except Exception:
        traffic = None
    line = {'metric': 'read pairs/sec (2x150, FASTQ-formatted, corrupted)' if corrupt else 'read pairs/sec (2x150, FASTQ-formatted, perfect reads)',
            'value': pairs_all / (ms * 1e-3), 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if args.workload == 'wgs' else 'weak', 'vs_baseline': None, 'dtype': 'u8',
            'data': 'synthetic', 'config': dict(config_dict(args), host_binding=('rank 0 on cores {}..{} (GPU-local, NVML)'.format(numa_cores[0], numa_cores[-1]) if numa_cores else 'none')), 'clocks': clk,
            'gpu_launches': prof['total_launches'],
            'roofline': {'bound': 'hbm', 'kernel': 'k_unit_emit', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                         'traffic': traffic, 'peak_source': peak_src, 'launches': prof['emit_launches'],
                         'avg_launch_ms': prof['emit_ms'] / max(1, prof['emit_launches']),
                         'algorithmic_bytes_per_launch': alg / max(1, prof['emit_launches']),
                         'algorithmic_bytes_per_pair': alg / max(1, pairs),
                         'other_kernels': {'k_unit_plan_avg_ms': prof['plan_ms'] / max(1, prof['emit_launches'])}}}
    if e2e:
      line['e2e'] = {'value': e_pairs_all / e_wall_max, 'unit': 'pairs/s', 'h2d_bytes_per_step': int(e2e[2]), 'd2h_bytes_per_step': int(e2e[3]),
                     'sink': 'pinned host memory'}
    if not args.no_cpu_baseline and world == 1:
      import oracle
      oracle.build()
      n, wall, sl = cpu_baseline(args, wl if args.workload == 'chr1' else {'contigs': [wl['contigs'][0]], 'tables': [wl['tables'][0]]}, 1)
      line['cpu_baseline'] = {'value': n / wall, 'unit': 'pairs/s', 'cores': 1, 'kind': 'port',
                              'sample': 'one {} Mb-slice work unit of the same contig (generate + corrupt, {} pairs), C oracle, 1 thread'.format(sl // 1000000, n)}
    emit(line)
  eng.close()
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
