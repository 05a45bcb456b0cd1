"""generate-reads: the orchestration of ``mitty/simulation/readgenerate.py`` on GPUs.

Same entry point and arguments as the reference (``process_multi_threaded``,
readgenerate.py:76-78); same work-unit schedule (``get_data_for_workers``, 129-159); same FASTQ /
qname bytes (``fastq_lines``, 222-230).  What differs is where the work happens: every unit's
template sampling, node lookup, sequence extraction, reverse complement and record formatting is
one kernel launch (``k_unit_emit``); the host only walks the schedule and appends each unit's byte
ranges to the two output files in schedule order.

Modes
  philox         production: all draws come from Philox on the device (4x32-10 for template sampling,
                 4x32-7 for the per-base corruption stream)
  deterministic  the reference's own numpy RandomState draws are made on the host and consumed on
                 the device; the output equals the reference's ``--threads 1`` FASTQ byte for byte

Multi-GPU: units are independent (each carries its own seed), so they are dealt to the GPUs by
longest-processing-time-first; there is no collective, only the ordered concatenation done here.
"""
import logging
import os
import time

import numpy as np

import mitty_b200.lib.vcfio as vio
from mitty_b200.engine import BATCH_MAX_CANDIDATES, MODE_DET, MODE_PHILOX, SEED_MAX, Engine

logger = logging.getLogger(__name__)

__qname_format__ = '@read_serial|chrom|copy|strand|pos|rlen|cigar|vs1,vs2,...|strand|pos|rlen|cigar|vs1,vs2,...'
__qname_format_details__ = """
@read_serial|chrom|copy|strand|pos|rlen|cigar|vs1,vs2,...|strand|pos|rlen|cigar|vs1,vs2,...

  read_serial  unique code for the template: sample:worker:unit:count
  chrom        chromosome the read was taken from (as in the BED file)
  copy         copy of the chromosome the read was taken from (0, 1, ...)
  strand       forward strand (0) or reverse strand (1)
  pos          position of the read, one based
  rlen         read length
  cigar        CIGAR of the read against the reference (ops: = X I D)
  vs1,vs2,...  comma separated sizes of the variants the read covers (0 SNP, +n INS, -n DEL)

The five per-read fields are repeated for the other read of the template (in file order).
chrom and pos are one based to make comparing qname info in a genome browser easier.

For reads from inside a long insertion the CIGAR has the format '>p:nI' where '>' marks a read
inside a long insertion, p is how many bases into the insertion the read starts and n is the read
length.
"""


def get_data_for_workers(model, vcf, seed):
  """The reference's work-unit schedule (readgenerate.py:129-159): one unit per (region, copy,
  pass), each with its own seed, then shuffled.  Yields dicts in consumption order."""
  seed_rng = np.random.RandomState(seed)
  shuffle_seed = seed_rng.randint(SEED_MAX)
  region_list = [
    {'region_idx': idx, 'region_cpy': cpy, 'rng_seed': seed_rng.randint(SEED_MAX)}
    for idx, v in enumerate(vcf)
    for cpy in range(len(v['v']))
    for _ in range(model['passes'])
  ]
  logger.debug('{} passes will be made'.format(len(region_list)))
  np.random.RandomState(shuffle_seed).shuffle(region_list)
  for region in region_list:
    yield region


def _without_end_crossing_deletions(vl, region, drop):
  """A deletion reaching beyond the region end makes the reference's node list end in 'D'
  (readgenerate.py:192 assumes that never happens; its p_max then overshoots the haplotype and the
  reads at the region end come out short).  That output cannot be reproduced, so by default such a
  region is an error (the message of mg_copy_build); with ``drop`` (``--drop-end-deletions``) the
  offending deletions are left out -- a documented deviation from the reference -- and counted."""
  if len(vl) == 0:
    return vl, 0
  cross = (vl.op == ord('D')) & (vl.pos + vl.oplen > region[2])
  if not cross.any():
    return vl, 0
  if not drop:
    raise ValueError('Region {}: {} deletion(s) reach beyond the region end (first at POS {}): the reference\'s node list would '
                     'end in \'D\' (readgenerate.py:192 assumes it never does). Trim the BED region or the VCF, or pass '
                     '--drop-end-deletions to leave these deletions out'.format(region, int(cross.sum()), int(vl.pos[cross][0])))
  logger.warning('Region {}: leaving out {} deletion(s) that reach beyond the region end'.format(region, int(cross.sum())))
  keep = np.flatnonzero(~cross)

  def pool(p, off):
    ln = (off[1:] - off[:-1])[keep]
    noff = np.zeros(keep.size + 1, dtype=np.int64); np.cumsum(ln, out=noff[1:])
    src = np.repeat(off[:-1][keep] - noff[:-1], ln) + np.arange(noff[-1])
    return (p[src] if src.size else np.zeros(0, dtype=np.uint8)), noff
  ap, ao = pool(vl.alt_pool, vl.alt_off)
  rp, ro = pool(vl.ref_pool, vl.ref_off) if vl.ref_pool is not None else (None, None)
  return vio.VariantList(vl.pos[keep], vl.op[keep], vl.oplen[keep], ap, ao, rp, ro), int(cross.sum())


PIN_REGION_BYTES = 4 << 20   # regions from this size on may be fetched into page-locked memory ...
PIN_MIN_REGIONS = 4           # ... by a worker that loads at least this many of them


class RegionCache(object):
  """Regions and chromosome copies resident in HBM, built on first use.  With ``expect``
  ({(region idx, copy): number of units that will ask for it}) each is released after its last
  unit; without it (units are pulled dynamically, nobody knows who gets what) the least recently
  used ones are released once ``budget`` bytes of packed sequence are resident."""

  def __init__(self, engine, vcf_df, fetch_ref, expect=None, drop_end_deletions=False, budget=24 << 30):
    self.engine, self.vcf_df, self.fetch_ref = engine, vcf_df, fetch_ref
    self.drop_end_deletions, self.dropped = drop_end_deletions, 0
    self.regions, self.copies = {}, {}
    self.left = dict(expect) if expect else None
    self.left_copies = {}
    for (r_idx, _cpy) in (expect or {}):
      self.left_copies[r_idx] = self.left_copies.get(r_idx, 0) + 1
    self.budget, self.resident, self.clock, self.used = budget, 0, 0, {}
    self._pin = None          # page-locked landing area for the reference bytes of large regions (see _fetch)
    self._large_loads = 0
    self._large_expected = len(set(r for (r, _c) in (expect or {}) if self._span(r) >= PIN_REGION_BYTES)) if expect else None

  def _span(self, r_idx):
    region = self.vcf_df[r_idx]['region']
    return region[2] - region[1]

  def _fetch(self, region):
    """The region's reference bytes.  A fetcher with ``into`` (FastaFetcher) can fill a page-locked buffer of this
    cache: the copy to the device is then one DMA instead of the driver's staged pieces, which queue behind the
    device-to-host pieces of the units being drained.  Locking the buffer costs about as much as three staged copies
    of a chromosome, so it is only done for a worker that loads many large regions (known from its unit list, or
    after the third such load).  (load_region returns after the copy: the buffer is free again.)"""
    into = getattr(self.fetch_ref, 'into', None)
    n = region[2] - region[1]
    if into is None or n < PIN_REGION_BYTES:
      return self.fetch_ref(region)
    self._large_loads += 1
    many = self._large_expected >= PIN_MIN_REGIONS if self._large_expected is not None else self._large_loads >= PIN_MIN_REGIONS
    if not many and self._pin is None:
      return self.fetch_ref(region)
    if self._pin is None or self._pin.size < n:
      self._pin = self.engine.pinned(max([n] + [r['region'][2] - r['region'][1] for r in self.vcf_df]))
    return into(region, self._pin)

  def _cost(self, r_idx):
    region = self.vcf_df[r_idx]['region']
    return (region[2] - region[1]) // 4 + 4096

  def _evict(self, keep):
    """LRU: copies first (a region can only go when none of its copies is resident)."""
    while self.resident > self.budget:
      cands = [k for k in self.copies if k != keep]
      if not cands:
        break
      k = min(cands, key=lambda c: self.used.get(c, 0))
      self.engine.free_copy(self.copies.pop(k))
      self.resident -= self._cost(k[0])
      if not any(c[0] == k[0] for c in self.copies) and k[0] != keep[0]:
        self.engine.free_region(self.regions.pop(k[0]))
        self.resident -= self._cost(k[0])

  def copy(self, r_idx, cpy):
    key = (r_idx, cpy)
    self.clock += 1
    self.used[key] = self.clock
    if key not in self.copies:
      region = self.vcf_df[r_idx]['region']
      if r_idx not in self.regions:
        self.regions[r_idx] = self.engine.load_region(self._fetch(region), region[1])
        self.resident += self._cost(r_idx)
      vl, n_drop = _without_end_crossing_deletions(self.vcf_df[r_idx]['v'][cpy], region, self.drop_end_deletions)
      self.dropped += n_drop
      self.copies[key] = self.engine.build_copy(self.regions[r_idx], vl)
      self.resident += self._cost(r_idx)
      if self.left is None:
        self._evict(key)
    return self.copies[key]

  def done(self, r_idx, cpy, built=True):
    """One unit of (region, copy) has been generated (built=False: on the batch path, which builds its own
    segments -- nothing of it is resident here)."""
    if self.left is None:
      return
    key = (r_idx, cpy)
    self.left[key] -= 1
    if self.left[key] == 0:
      if key in self.copies:
        self.engine.free_copy(self.copies.pop(key))
      del self.left[key]
      self.left_copies[r_idx] -= 1
      if self.left_copies[r_idx] == 0 and r_idx in self.regions:
        self.engine.free_region(self.regions.pop(r_idx))


class FastaFetcher(object):
  """fetch_ref of generate-reads over a FastaFile: ``f(region)`` -> the region's bytes (fasta.fetch(...),
  readgenerate.py:186); ``f.into(region, out)`` -> the same, written into ``out``."""

  def __init__(self, fasta):
    self.fasta = fasta

  def __call__(self, region):
    return self.fasta.fetch(reference=region[0], start=region[1], end=region[2])

  def into(self, region, out):
    return self.fasta.fetch(reference=region[0], start=region[1], end=region[2], out=out)


def generate_unit(engine, read_module, read_model, cp, chrom, cpy, rng_seed, sample_name, worker_id, ps,
                  mode='philox', corrupt=False, corrupt_seed=0, out=None, fetch=True, wait=True):
  """One work unit (the body of read_generating_worker's loop, readgenerate.py:183-214)
  -> (fastq1 bytes, fastq2 bytes, template count, templates that passed te < p_max, bytes per file);
  the two byte arrays are None with fetch=False (the unit stays on the device)."""
  n = int((cp.p_max - cp.p_min) * read_model['p'] * 1.2)          # illumina.py:69
  prefix = '@{}:{}:{}:'.format(sample_name, worker_id, ps)         # readgenerate.py:195, 210
  mid = '|{}|{}'.format(chrom, cpy)                                # readgenerate.py:223
  if mode == 'deterministic':
    ts, u, fo = read_module.unit_draws(read_model, cp.p_min, cp.p_max, rng_seed)
    return engine.generate_unit(cp, n, read_model['p'], MODE_DET, rng_seed, prefix, mid, ts=ts, u_tlen=u, fo=fo,
                                out=out, fetch=fetch, wait=wait)
  if not (0 <= rng_seed <= SEED_MAX):
    raise ValueError('Seed value {} is out of range 0 - {}'.format(rng_seed, SEED_MAX))
  return engine.generate_unit(cp, n, read_model['p'], MODE_PHILOX, rng_seed, prefix, mid,
                              corrupt=corrupt, corrupt_seed=corrupt_seed, out=out, fetch=fetch, wait=wait)


# ---- batches of small regions ----------------------------------------------------------------------
BATCH_MAX_BASES = 256 << 20   # reference bases of one batch (its regions are packed back to back)
BATCH_MAX_UNITS = 1 << 18


def batchable(vcf_df, read_model, sample_name, n_units):
  """-> {(region idx, copy)} whose units may take the batch path: at most BATCH_MAX_CANDIDATES template
  candidates per unit (illumina.py:69; the haplotype is at most the region plus every inserted base
  long) and qname strings that fit the unit table (32 bytes each)."""
  if len('@{}:0:{}:'.format(sample_name, max(0, n_units - 1)).encode()) > 32:
    return set()
  ok = set()
  for r_idx, r in enumerate(vcf_df):
    region = r['region']
    if len(str(region[0]).encode()) > 20:
      continue
    for cpy, vl in enumerate(r['v']):
      ins = int(vl.oplen[vl.op == ord('I')].sum()) if len(vl) else 0
      if cpy < 10**9 and int((region[2] - region[1] + ins) * read_model['p'] * 1.2) <= BATCH_MAX_CANDIDATES:
        ok.add((r_idx, cpy))
  return ok


def generate_batch(engine, read_module, read_model, units, schedule, vcf_df, fetch_ref, sample_name, mode='philox', corrupt=False,
                   corrupt_seed=0, drop_end_deletions=False, sink=None, producer=0):
  """The work units ``units`` (ascending schedule indices, all of them batchable) in ONE launch sequence:
  their regions packed back to back, one node table / haplotype over all (region, copy) segments
  (mg_batch_build), one planning CTA per unit and one emit launch (mg_batch_generate).  The bytes of
  every unit equal those of ``generate_unit``.
  -> (templates, bytes per file, per-unit bytes, per-unit templates, deletions dropped)."""
  t0 = time.perf_counter()
  reg_of, seg_of = {}, {}
  refs, bed_starts, seg_region, seg_variants, seg_chrom, seg_cpy = [], [], [], [], [], []
  unit_seg, dropped = [], 0
  for k in units:
    wd = schedule[k]
    r_idx, cpy = wd['region_idx'], wd['region_cpy']
    region = vcf_df[r_idx]['region']
    if r_idx not in reg_of:
      reg_of[r_idx] = len(refs)
      refs.append(np.asarray(fetch_ref(region), dtype=np.uint8))
      bed_starts.append(region[1])
    if (r_idx, cpy) not in seg_of:
      seg_of[(r_idx, cpy)] = len(seg_region)
      vl, n_drop = _without_end_crossing_deletions(vcf_df[r_idx]['v'][cpy], region, drop_end_deletions)
      dropped += n_drop
      seg_region.append(reg_of[r_idx]); seg_variants.append(vl); seg_chrom.append(region[0]); seg_cpy.append(cpy)
    unit_seg.append(seg_of[(r_idx, cpy)])
  t1 = time.perf_counter()
  batch = engine.build_batch(refs, bed_starts, seg_region, seg_variants)
  t2 = time.perf_counter()
  try:
    seeds = [int(schedule[k]['rng_seed']) for k in units]
    ncand = [int((batch.p_max[sg] - batch.p_min[sg]) * read_model['p'] * 1.2) for sg in unit_seg]          # illumina.py:69
    draws = None
    if mode == 'deterministic':
      per = [read_module.unit_draws(read_model, int(batch.p_min[sg]), int(batch.p_max[sg]), sd) for sg, sd in zip(unit_seg, seeds)]
      off = np.zeros(len(per) + 1, dtype=np.int64)
      np.cumsum([d[0].size for d in per], out=off[1:])
      cat = lambda i, dt: np.concatenate([d[i] for d in per]).astype(dt, copy=False) if per else np.zeros(0, dtype=dt)  # noqa: E731
      draws = (cat(0, np.int64), cat(1, np.float64), cat(2, np.int8), off)
    else:
      for sd in seeds:
        if not (0 <= sd <= SEED_MAX):
          raise ValueError('Seed value {} is out of range 0 - {}'.format(sd, SEED_MAX))
    t3 = time.perf_counter()
    nt, nb, ub, ut = engine.generate_batch(batch, unit_seg, seeds, ncand, list(units), sample_name, seg_chrom, seg_cpy, read_model['p'],
                                           MODE_DET if mode == 'deterministic' else MODE_PHILOX, draws=draws, corrupt=corrupt,
                                           corrupt_seed=corrupt_seed, sink=sink, producer=producer)
    logger.info('Batch of {} units over {} segments: inputs {:0.3f}s, build {:0.3f}s, draws {:0.3f}s, kernels {:0.3f}s'.format(
      len(units), len(seg_region), t1 - t0, t2 - t1, t3 - t2, time.perf_counter() - t3))
  finally:
    engine.free_batch(batch)
  return nt, nb, ub, ut, dropped


def _runs(my_units, schedule, ok, vcf_df):
  """my_units (ascending) cut into runs that keep the order: ('batch', [units]) for stretches of batchable
  units (bounded in reference bases and units), ('unit', k) for the others."""
  cur, bases, seen = [], 0, set()
  for k in my_units:
    wd = schedule[k]
    key = (wd['region_idx'], wd['region_cpy'])
    if key in ok:
      if wd['region_idx'] not in seen:
        region = vcf_df[wd['region_idx']]['region']
        span = region[2] - region[1]
        if cur and (bases + span > BATCH_MAX_BASES or len(cur) >= BATCH_MAX_UNITS):
          yield 'batch', cur
          cur, bases, seen = [], 0, set()
        seen.add(wd['region_idx']); bases += span
      cur.append(k)
    else:
      if cur:
        yield 'batch', cur
        cur, bases, seen = [], 0, set()
      yield 'unit', k
  if cur:
    yield 'batch', cur


last_run = {}              # statistics of the most recent process_multi_threaded call (the reference returns nothing)
CHUNK_BYTES = 64 << 20     # sink slot per file: a unit travels to the writer threads in pieces of this size
SLOTS_PER_GPU = 6          # page-locked slot pairs per GPU: the spill that lets GPUs run ahead of the files


def gpu_worker(device, producer, sink, schedule, vcf_df, fetch_ref, read_module, read_model, sample_name, mode, corrupt, corrupt_seed,
               drop_end_deletions=False, stats=None, engine=None, my_units=None, batch_small=True, batch_ok=None):
  """One host thread (or process) per GPU worker.  With ``my_units`` (ascending schedule indices: all
  units of a region on one worker, so each region / copy is built once) the worker walks its list;
  without, units are PULLED from the sink's counter one at a time, in schedule order across all
  workers (several processes sharing one sink).  Per unit: region / copy from the cache, the
  unit's kernels (its bytes stay on the device), then the context's drain thread copies it out piece
  by piece into the sink while this thread already generates the next unit.
  -> templates written by this worker."""
  own_engine = engine is None
  if own_engine:                         # (a caller that keeps an engine across calls has loaded the model)
    from mitty_b200.engine import bind_host_thread_to_gpu
    bind_host_thread_to_gpu(device)      # the worker's pinned memory traffic stays in the GPU's socket
    engine = Engine(device)
  total = 0
  t_build = t_gen = 0.0
  cache = None
  try:
    if own_engine:
      engine.load_model(read_model)
    expect = None
    if my_units is not None:
      expect = {}
      for k in my_units:
        key = (schedule[k]['region_idx'], schedule[k]['region_cpy'])
        expect[key] = expect.get(key, 0) + 1
    cache = RegionCache(engine, vcf_df, fetch_ref, expect, drop_end_deletions)
    # stretches of small units (an exome-style BED) take the batch path: one launch sequence per stretch
    ok = set()
    if my_units is not None and batch_small:
      ok = batch_ok if batch_ok is not None else batchable(vcf_df, read_model, sample_name, len(schedule))
    todo = _runs(my_units, schedule, ok, vcf_df) if my_units is not None else None
    n_batches = 0
    while True:
      kind, k = next(todo, ('end', -1)) if todo is not None else ('unit', sink.next_unit())
      if kind == 'end' or (kind == 'unit' and k < 0):
        break
      if kind == 'batch' and len(k) == 1:
        kind, k = 'unit', k[0]
      ta = time.perf_counter()
      if kind == 'batch':
        cnt, _, _, _, n_drop = generate_batch(engine, read_module, read_model, k, schedule, vcf_df, fetch_ref, sample_name, mode=mode,
                                              corrupt=corrupt, corrupt_seed=corrupt_seed, drop_end_deletions=drop_end_deletions,
                                              sink=sink, producer=producer)
        for u in k:
          cache.done(schedule[u]['region_idx'], schedule[u]['region_cpy'], built=False)
        cache.dropped += n_drop
        total += cnt
        n_batches += 1
        t_gen += time.perf_counter() - ta
        logger.debug('Batch of {} units: {} templates'.format(len(k), cnt))
        continue
      wd = schedule[k]
      r_idx, cpy = wd['region_idx'], wd['region_cpy']
      cp = cache.copy(r_idx, cpy)
      tb = time.perf_counter()
      _, _, cnt, _, nb = generate_unit(engine, read_module, read_model, cp, vcf_df[r_idx]['region'][0], cpy, int(wd['rng_seed']),
                                       sample_name, 0, k, mode=mode, corrupt=corrupt, corrupt_seed=corrupt_seed, fetch=False)
      engine.drain_async(sink, producer, k)            # announces the unit's size, then streams it
      cache.done(r_idx, cpy)
      total += cnt
      tc = time.perf_counter()
      t_build += tb - ta; t_gen += tc - tb
      logger.debug('Unit {}: {} templates'.format(k, cnt))
    engine.drain_wait()
    logger.info('GPU {}: {} templates; region/copy builds {:0.2f}s, unit kernels {:0.2f}s'.format(device, total, t_build, t_gen))
    if stats is not None:
      stats.update(build_s=t_build, gen_s=t_gen, dropped=cache.dropped, batches=n_batches)
    return total
  except BaseException as e:
    sink.abort('{}: {}'.format(type(e).__name__, e))   # wakes every producer and writer
    raise
  finally:
    if own_engine:
      engine.close()
    elif cache is not None:              # the caller's engine lives on: give the regions and copies back
      try:
        engine.drain_wait()
      except Exception:
        pass
      for cp in cache.copies.values():
        engine.free_copy(cp)
      for rid in cache.regions.values():
        engine.free_region(rid)


def process_multi_threaded(fasta_fname, vcf_fname, sample_name, bed_fname, read_module, model, coverage,
                           fastq1_fname, fastq2_fname, threads=2, seed=7, mode='philox', corrupt=False,
                           corrupt_seed=None, devices=None, drop_end_deletions=False, gzip_level=None, sink_threads=None,
                           workers_per_gpu=None, batch_small=True):
  """Same signature as the reference (readgenerate.py:76-78) plus keyword-only extras.

  ``threads`` = number of GPUs to use (capped by the GPUs present; ``devices`` overrides).  One host
  thread drives each GPU; the units are handed out in schedule order and written by the native
  output sink (writer threads inside the library) IN SCHEDULE ORDER: pwrite at the final offset for
  regular files, ordered sequential writes when the targets are FIFOs / process substitutions.
  Output order and qname serials are those of the reference's ``--threads 1`` run (worker id 0, unit
  index = schedule index), whatever the GPU count.

  workers_per_gpu: host threads (each with its own context and stream) per GPU (default 1).
  batch_small: stretches of small work units (regions of a few kb: an exome-style BED) are generated in ONE
  launch sequence each (``generate_batch``); the bytes are the same either way.
  gzip_level: 1-9 writes multi-member gzip (what the reference's ``>(gzip > r1.fq.gz)`` produces,
  Readme.md:170, without the external process); None = by file name ('.gz'), 0 = plain.
  Page-locked memory: SLOTS_PER_GPU x CHUNK_BYTES per file and GPU (768 MB per GPU for a pair).
  """
  import threading
  from mitty_b200.engine import Sink, device_count

  t_in = time.time()
  if mode not in ('philox', 'deterministic'):
    raise ValueError('mode must be "philox" or "deterministic"')
  if corrupt and mode == 'deterministic':
    # checked before any file is opened (and truncated) or any GPU work is started
    raise ValueError('fused corruption draws from Philox; for the deterministic mode run generate-reads --deterministic '
                     'and then corrupt-reads --deterministic, as the reference does')
  read_model = read_module.read_model_params(model, coverage)
  if workers_per_gpu is None:
    workers_per_gpu = 1      # (measured: several contexts on one GPU do not overlap the per-unit latencies -- the driver serialises them)
  # Start-up work that needs no parsed input runs beside the parsing: CUDA initialisation and, when the regions are large
  # enough for full-size slots, the page-locking of the sink's slots (seconds for the 6 GB of an 8-GPU run)
  warm = {}

  def warm_up():
    try:
      from mitty_b200 import _lib
      warm['n_dev'] = device_count()
      bed_span = max([r[2] - r[1] for r in vio.read_bed(bed_fname)] + [1])
      est0 = int(bed_span * 1.05 * read_model['p'] * 1.2 * (2 * int(read_model['rlen']) + 150)) + (1 << 16)
      if est0 >= CHUNK_BYTES and warm['n_dev'] > 0:
        n_workers = len(devices) if devices is not None else max(1, min(int(threads), warm['n_dev']))
        n_bufs = n_workers * max(1, int(workers_per_gpu)) * 2 * (2 if fastq2_fname is not None else 1)   # the sink locks the rest while units already travel
        warm['locked'] = _lib.lib().mg_sink_prealloc(CHUNK_BYTES, n_bufs, 4)
    except Exception as e:  # noqa: B902 -- best effort: the sink allocates what is missing
      warm['error'] = e
  warm_thread = threading.Thread(target=warm_up, daemon=True)
  warm_thread.start()
  vcf_df = vio.load_variant_file(vcf_fname, sample_name, bed_fname)
  for r in vcf_df:                       # inputs the engine rejects: found before the outputs are opened
    for vl in r['v']:
      _without_end_crossing_deletions(vl, r['region'], drop_end_deletions)
  fasta = vio.FastaFile(fasta_fname)
  fetch_ref = FastaFetcher(fasta)
  schedule = list(get_data_for_workers(read_model, vcf_df, seed))
  t_parsed = time.time()
  warm_thread.join()
  if devices is None:
    n_dev = warm.get('n_dev', device_count())
    if n_dev < 1:
      raise RuntimeError('mitty_b200: no CUDA device; the engine has no CPU fallback')
    devices = list(range(max(1, min(int(threads), n_dev))))
  devices = [d for d in devices for _ in range(max(1, int(workers_per_gpu)))]
  from mitty_b200 import multigpu
  weights = [vcf_df[wd['region_idx']]['region'][2] - vcf_df[wd['region_idx']]['region'][1] for wd in schedule]
  assign = multigpu.assign_by_region(schedule, weights, len(devices))
  if gzip_level is None:
    gzip_level = 1 if str(fastq1_fname).endswith('.gz') else 0
  cs = seed if corrupt_seed is None else corrupt_seed
  span = max([r['region'][2] - r['region'][1] for r in vcf_df] + [1])
  rlen = int(read_model['rlen'])
  est = int(span * 1.05 * read_model['p'] * 1.2 * (2 * rlen + 150)) + (1 << 16)      # bytes per file of the largest unit
  small = batchable(vcf_df, read_model, sample_name, len(schedule)) if batch_small else set()
  if small:                              # stretches of small units leave the device as one stream
    tot = sum(weights[k] for k, wd in enumerate(schedule) if (wd['region_idx'], wd['region_cpy']) in small)
    est = max(est, int(tot / len(devices) * 1.05 * read_model['p'] * 1.2 * (2 * rlen + 150)) + (1 << 16))
  chunk = max(256, min(CHUNK_BYTES, est))
  # writer threads: the cores the box has beyond one drain thread per worker (they spin on their copy events); tmpfs
  # writes keep scaling up to there (8 / 16 writers: 11.9 / 14.6 GB/s on a 16-core box), deflate needs every core
  cores = os.cpu_count() or 4
  n_writers = sink_threads or max(4, min(24, cores - len(devices) - 2) if not gzip_level else cores)
  sink = Sink(fastq1_fname, fastq2_fname, len(schedule), n_producers=len(devices), slots=SLOTS_PER_GPU, chunk_bytes=chunk,
              gzip_level=gzip_level, threads=n_writers)
  t0 = time.time()
  totals, errors = [0] * len(devices), []
  wstats = [{} for _ in devices]

  def run(i, dev):
    try:
      totals[i] = gpu_worker(dev, i, sink, schedule, vcf_df, fetch_ref, read_module, read_model, sample_name, mode, corrupt, cs,
                             drop_end_deletions, stats=wstats[i], my_units=assign[i], batch_small=batch_small, batch_ok=small)
    except BaseException as e:  # noqa: B902 -- re-raised below, in the caller's thread
      errors.append(e)

  workers = [threading.Thread(target=run, args=(i, dev), daemon=True) for i, dev in enumerate(devices)]
  try:
    for w in workers:
      w.start()
    for w in workers:
      w.join()
  finally:
    if errors:
      sink.abort(str(errors[0]))
    try:
      sink.close()
    except IOError:
      if not errors:
        raise
  if errors:
    raise errors[0]
  t1 = time.time()
  total = sum(totals)
  logger.debug('Finished: {} templates in {:0.2f}s ({:0.2f} t/s); inputs {:0.2f}s'.format(total, t1 - t0, total / max(t1 - t0, 1e-9), t0 - t_in))
  last_run.update(templates=total, seconds=t1 - t0, input_seconds=t0 - t_in, parse_seconds=t_parsed - t_in, gpus=len(devices), writers=n_writers, gzip_level=gzip_level,
                  batches=sum(w.get('batches', 0) for w in wstats), dropped=sum(w.get('dropped', 0) for w in wstats))
  return None


# ---- the qname contract's inverse (readgenerate.py:256-291) --------------------------------------

class ReadInfo(tuple):
  """What the qname says about one read: the fields of the reference's ``ri`` tuple, in its order."""
  __slots__ = ()
  _fields = ('sample', 'rid', 'chrom', 'cpy', 'strand', 'pos', 'rlen', 'cigar', 'special_cigar', 'v_list')

  def __new__(cls, *a):
    return tuple.__new__(cls, a)

  def __getattr__(self, name):
    try:
      return self[self._fields.index(name)]
    except ValueError:
      raise AttributeError(name)


def parse_qname(qname):
  """qname (without the leading '@') -> [ReadInfo of the read in file 1, ReadInfo of the read in file 2].

  Layout (``__qname_format__``): three template fields, then five fields per read.  A read taken from
  inside a long insertion carries '>p:nI' as its CIGAR: the plain CIGAR is then 'nI' and the original
  string is kept as ``special_cigar`` (readgenerate.py:277-283)."""
  f = qname.split('|')
  rid, chrom, cpy = f[0], f[1], int(f[2])
  sample = rid.split(':', 1)[0]
  out = []
  for k in range(3, len(f) - 4, 5):
    strand, pos, rlen, cigar, vs = f[k:k + 5]
    special = cigar if cigar.startswith('>') else None
    if special is not None:
      cigar = cigar.rsplit(':', 1)[1]
    out.append(ReadInfo(sample, rid, chrom, cpy, int(strand), int(pos), int(rlen), cigar, special,
                        [int(v) for v in vs.split(',') if v]))
  return out
