"""Sharding the work units of one generate-reads job over the GPUs of a box.

Units (BED region x chromosome copy x pass) are independent: each carries its own seed and only
reads per-region state (readgenerate.py:149-154).  So there is no collective on the data path:

* ``assign_units``   longest-processing-time-first assignment by haplotype length,
* every rank generates its units into a part file + an index of (schedule idx, offset, length),
* ``concatenate``    the ordered host-side concatenation: rank 0 appends the unit byte ranges in
                     schedule order, so the result is byte-identical to a one-GPU run (qname serials
                     use worker id 0 and the schedule index on every rank).

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used only for the barrier and for
gathering the small index lists.
"""
import os

import numpy as np


def assign_units(weights, world):
  """LPT: units sorted by weight descending, each to the least loaded rank.
  -> list (per rank) of unit indices, each ascending."""
  order = np.argsort(-np.asarray(weights, dtype=np.float64), kind='stable')
  load = np.zeros(world)
  out = [[] for _ in range(world)]
  for k in order.tolist():
    r = int(np.argmin(load))
    out[r].append(k)
    load[r] += weights[k]
  return [sorted(x) for x in out]


def assign_by_region(schedule, weights, world):
  """Units dealt to ``world`` workers so that all units of a BED region go to ONE worker (its packed
  reference and chromosome copies are then built once): regions by LPT on their total weight, each
  worker's units in ascending schedule order.  -> list (per worker) of unit indices."""
  tot = {}
  for k, wd in enumerate(schedule):
    tot[wd['region_idx']] = tot.get(wd['region_idx'], 0.0) + float(weights[k])
  regions = sorted(tot, key=lambda r: (-tot[r], r))
  load = [0.0] * world
  owner = {}
  for r in regions:
    w = min(range(world), key=lambda i: (load[i], i))
    owner[r] = w
    load[w] += tot[r]
  out = [[] for _ in range(world)]
  for k, wd in enumerate(schedule):
    out[owner[wd['region_idx']]].append(k)
  return out


def write_part(part_prefix, rank, unit_bytes):
  """unit_bytes: iterable of (schedule idx, bytes file1, bytes file2) -> index list."""
  index = []
  with open('{}.r{}.1'.format(part_prefix, rank), 'wb') as f1, open('{}.r{}.2'.format(part_prefix, rank), 'wb') as f2:
    off = 0
    for k, b1, b2 in unit_bytes:
      f1.write(memoryview(b1)); f2.write(memoryview(b2))
      assert len(b1) == len(b2)
      index.append((int(k), off, len(b1)))
      off += len(b1)
  return index


def concatenate(part_prefix, indices, fastq1, fastq2, chunk=1 << 26):
  """indices[rank] = [(schedule idx, offset, length)].  Sequential writes only (the targets may be
  FIFOs / process substitutions, Readme.md:170)."""
  where = {}
  for rank, idx in enumerate(indices):
    for k, off, ln in idx:
      where[k] = (rank, off, ln)
  for which, target in ((1, fastq1), (2, fastq2)):
    if target is None:
      continue
    parts = [open('{}.r{}.{}'.format(part_prefix, r, which), 'rb') for r in range(len(indices))]
    with open(target, 'wb') as out:
      for k in sorted(where):
        rank, off, ln = where[k]
        parts[rank].seek(off)
        left = ln
        while left:
          blk = parts[rank].read(min(chunk, left))
          out.write(blk)
          left -= len(blk)
    for p in parts:
      p.close()


def cleanup(part_prefix, world):
  for r in range(world):
    for which in (1, 2):
      try:
        os.remove('{}.r{}.{}'.format(part_prefix, r, which))
      except OSError:
        pass


def run_sharded(units, weights, generate_unit, fastq1, fastq2, part_prefix=None):
  """Run inside a torch.distributed process group (one process per GPU).
  units: the schedule (list, consumption order); generate_unit(schedule idx, unit) -> (bytes1, bytes2)."""
  import torch.distributed as dist
  rank, world = dist.get_rank(), dist.get_world_size()
  part_prefix = part_prefix or (fastq1 + '.part')
  mine = assign_units(weights, world)[rank]
  index = write_part(part_prefix, rank, ((k,) + tuple(generate_unit(k, units[k])) for k in mine))
  gathered = [None] * world
  dist.all_gather_object(gathered, index)
  dist.barrier()
  if rank == 0:
    concatenate(part_prefix, gathered, fastq1, fastq2)
    cleanup(part_prefix, world)
  dist.barrier()
  return gathered
