"""The Illumina read-model plugin, same contract as ``mitty/simulation/illumina.py``:

* ``read_model_params(model, diploid_coverage)``  -> dict incl. 'passes'      (illumina.py:12-40)
* ``generate_reads(model, p_min, p_max, seed)``   -> [{file_order, pos, len}] x 2 mates (43-110)
* ``corrupt_template(model, template, rng)``      -> [(qname, seq, bq_str), ...]   (113-162)

Template length inverse-CDF, the te < p_max filter and the whole corruption model run on the GPU
(``k_sample`` / ``k_corrupt`` in csrc/mg_kernels.cu).  The calls here are the *deterministic*
flavour: they consume numpy ``RandomState`` draws exactly as the reference does, so their results
equal the reference's bit for bit.  The production path (Philox draws, fused kernels) is driven by
``readgenerate.process_multi_threaded``.
"""
import numpy as np

from mitty_b200.engine import MODE_DET, SEED_MAX, default_engine


def read_model_params(model, diploid_coverage=30.0):
  """coverage -> (p, passes): p = 0.5 * cov / (2 * rlen * passes) with passes doubled until
  p <= 0.1 (illumina.py:27-31)."""
  rlen = model['mean_rlen']
  p = 1.0
  passes = 1
  while p > 0.1:
    passes *= 2
    p = 0.5 * diploid_coverage / (2 * rlen * passes)
  return {
    'diploid_coverage': diploid_coverage,
    'p': p,
    'passes': passes,
    'rlen': rlen,
    'cum_tlen': model['cum_tlen'],
    'cum_bq_mat': model['cum_bq_mat']
  }


def unit_draws(model, p_min, p_max, seed):
  """The reference's host-side draws for one work unit (illumina.py:56-58, 69-72, 93):
  -> (ts shuffled int64[N], u_tlen f64[N], fo i1[N]).  The file-order bits are drawn for all N
  candidates; the first K (K = templates with te < p_max) equal the reference's size-K draw."""
  if not (0 <= seed <= SEED_MAX):
    raise ValueError('Seed value {} is out of range 0 - {}'.format(seed, SEED_MAX))
  seed_rng = np.random.RandomState(seed)
  tloc_rng, tlen_rng, shuffle_rng, file_order_rng = [np.random.RandomState(s) for s in seed_rng.randint(SEED_MAX, size=4)]
  p = model['p']
  n = int((p_max - p_min) * p * 1.2)
  ts = tloc_rng.geometric(p=p, size=n).cumsum() + p_min + 1
  shuffle_rng.shuffle(ts)
  u = tlen_rng.rand(ts.shape[0])
  fo = file_order_rng.randint(2, size=n, dtype='i1')
  return ts.astype(np.int64, copy=False), u, fo


def generate_reads(model, p_min, p_max, seed=7, engine=None):
  """Same return value as the reference: one dict per mate with 'file_order' (int8), 'pos' (int64)
  and 'len' (uint32) arrays."""
  eng = engine or default_engine()
  ts, u, fo = unit_draws(model, p_min, p_max, seed)
  eng.load_model(model)
  n = ts.shape[0]
  ts_o, te_o, _ = eng.sample_templates(n, model['p'], MODE_DET, seed, p_min=p_min, p_max=p_max, ts=ts, u_tlen=u)
  idx = te_o >= 0
  ts_k, te_k = ts_o[idx], te_o[idx]
  rlen = model['rlen']
  r0fo = fo[:ts_k.size].copy()
  return [
    {'file_order': r0fo, 'pos': ts_k, 'len': np.full(ts_k.size, rlen, dtype=np.uint32)},
    {'file_order': 1 - r0fo, 'pos': te_k - rlen, 'len': np.full(ts_k.size, rlen, dtype=np.uint32)}
  ]


def corrupt_draws(seq_lens, corrupt_rng):
  """The reference's per-read draws (illumina.py:151-153) for reads of the given lengths, in order.
  -> (bq_rnd, call_rnd, base_rnd, draw_off)"""
  off = np.zeros(len(seq_lens) + 1, dtype=np.int64)
  np.cumsum(seq_lens, out=off[1:])
  bq, call, base = np.empty(off[-1]), np.empty(off[-1]), np.empty(off[-1], dtype=np.uint8)
  for k, rlen in enumerate(seq_lens):
    a, b = off[k], off[k + 1]
    bq[a:b] = corrupt_rng.rand(rlen)
    call[a:b] = corrupt_rng.rand(rlen)
    base[a:b] = corrupt_rng.randint(0, 3, size=rlen)
  return bq, call, base, off


def corrupt_template(model, template, corrupt_rng, engine=None):
  """template = (qname, seq[, seq]) -> [(qname, seq, bq), ...]; mate k uses cum_bq_mat[k]."""
  eng = engine or default_engine()
  qname, seqs = template[0], template[1:]
  eng.load_model(model, rlen=max(1, max(len(s) for s in seqs)))
  draws = corrupt_draws([len(s) for s in seqs], corrupt_rng)
  fq = ['@{}\n{}\n+\n{}\n'.format(qname, s, '~' * len(s)).encode() for s in seqs]
  o1, o2, _ = eng.corrupt_fastq(fq[0], fq[1] if len(fq) > 1 else None, mode=MODE_DET, draws=draws)
  out = []
  for o in (o1, o2):
    if o is None:
      continue
    lines = o.tobytes().decode().split('\n')
    out.append((qname, lines[1], lines[3]))
  return out


def describe_model(model_name, model, figfile):
  raise NotImplementedError('describe_model plots with matplotlib in the reference (illumina.py:165-210); '
                            'plotting is outside the read-generation hot path this engine covers')
