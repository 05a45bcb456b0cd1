"""CPU oracle for Mitty's read-generation hot path -- TEST INFRASTRUCTURE, not the product.

``oracle/mitty_oracle.c`` restates the reference's algorithm (mitty/simulation/rpc.py,
illumina.py, readgenerate.py, readcorrupt.py) in plain C; this module is its ctypes face.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  Nothing under ``mitty_b200/`` does.

Parity status: PINNED -- against all KATs of the reference's own tests (test_rpc.py, test_vcfio.py)
and against golden FASTQ produced by running the unmodified reference in the build container
(``tests/golden/make_golden.py``; fixtures under ``tests/golden/``).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle.so')
_lib = None

PHRED_P = 10 ** (-np.arange(100) / 10)  # illumina.py:137


def build(force=False):
  src = [os.path.join(_HERE, f) for f in ('mitty_oracle.c', 'mt19937.h')]
  if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
    subprocess.check_call(['make', '-C', _HERE, '-s', '-B'])
  return _SO


def lib():
  global _lib
  if _lib is None:
    if not os.path.exists(_SO):
      build()
    _lib = C.CDLL(_SO)
    for f in ('orc_create_node_list', 'orc_generate_read', 'orc_templates', 'orc_generate_unit',
              'orc_corrupt_fastq'):
      getattr(_lib, f).restype = C.c_int64
  return _lib


def _p(a, t):
  return a.ctypes.data_as(C.POINTER(t))


def _i64(a):
  return np.ascontiguousarray(a, dtype=np.int64)


class CopyVariants(object):
  """Variants present on one chromosome copy (the output of vcfio.parse for that copy)."""
  def __init__(self, pos, op, oplen, alts):
    self.pos = _i64(pos)
    self.op = np.frombuffer(''.join(op).encode(), dtype=np.uint8).copy() if not isinstance(op, np.ndarray) else np.ascontiguousarray(op, dtype=np.uint8)
    self.oplen = _i64(oplen)
    if isinstance(alts, tuple):  # (pool, off)
      self.alt_pool, self.alt_off = np.ascontiguousarray(alts[0], dtype=np.uint8), _i64(alts[1])
    else:
      self.alt_pool = np.frombuffer(''.join(alts).encode(), dtype=np.uint8).copy()
      self.alt_off = np.zeros(len(alts) + 1, dtype=np.int64)
      np.cumsum([len(a) for a in alts], out=self.alt_off[1:])
    if self.alt_pool.size == 0:
      self.alt_pool = np.zeros(1, dtype=np.uint8)

  def args(self):
    return (C.c_int64(self.pos.size), _p(self.pos, C.c_int64), _p(self.op, C.c_char),
            _p(self.oplen, C.c_int64), _p(self.alt_pool, C.c_char), _p(self.alt_off, C.c_int64))


def _ref(ref_seq):
  if isinstance(ref_seq, str):
    ref_seq = ref_seq.encode()
  if isinstance(ref_seq, (bytes, bytearray)):
    ref_seq = np.frombuffer(bytes(ref_seq), dtype=np.uint8)
  return np.ascontiguousarray(ref_seq, dtype=np.uint8)


# -- RNG recipes ---------------------------------------------------------------------------------

def rand(seed, n):
  out = np.empty(n, dtype=np.float64); lib().orc_rand(C.c_uint32(seed), _p(out, C.c_double), C.c_int64(n)); return out


def randint(seed, high, n):
  out = np.empty(n, dtype=np.int64); lib().orc_randint(C.c_uint32(seed), C.c_uint32(high), _p(out, C.c_int64), C.c_int64(n)); return out


def geometric(seed, p, n):
  out = np.empty(n, dtype=np.int64); lib().orc_geometric(C.c_uint32(seed), C.c_double(p), _p(out, C.c_int64), C.c_int64(n)); return out


def bits_i8(seed, n):
  out = np.empty(n, dtype=np.int8); lib().orc_bits_i8(C.c_uint32(seed), _p(out, C.c_int8), C.c_int64(n)); return out


def shuffle_i64(seed, x):
  x = _i64(x).copy(); lib().orc_shuffle_i64(C.c_uint32(seed), _p(x, C.c_int64), C.c_int64(x.size)); return x


def unit_schedule(seed, n_units):
  """a6 -> (seeds in nested (region, copy, pass) order, shuffled order of unit indices)."""
  seeds = np.empty(n_units, dtype=np.uint32); order = np.empty(n_units, dtype=np.int64)
  lib().orc_unit_schedule(C.c_uint32(seed), C.c_int64(n_units), _p(seeds, C.c_uint32), _p(order, C.c_int64))
  return seeds, order


def unit_seeds(seed):
  out = np.empty(4, dtype=np.uint32); lib().orc_unit_seeds(C.c_uint32(seed), _p(out, C.c_uint32)); return out


def corrupt_worker_seeds(seed, n):
  out = np.empty(n, dtype=np.uint32); lib().orc_corrupt_worker_seeds(C.c_uint32(seed), C.c_int(n), _p(out, C.c_uint32)); return out


# -- a7 / a11 / a12 ------------------------------------------------------------------------------

def create_node_list(ref_seq, ref_start_pos, cv):
  """-> list of (ps, pr, op, oplen, seq, v) tuples, exactly Node.tuple() of rpc.py:22-23."""
  ref = _ref(ref_seq)
  cap = 2 * cv.pos.size + 2
  pool_cap = ref.size + cv.alt_pool.size + 16
  ps, pr, oplen, v = (np.empty(cap, dtype=np.int64) for _ in range(4))
  seq_off = np.empty(cap + 1, dtype=np.int64)
  op = np.empty(cap, dtype=np.uint8); has_v = np.empty(cap, dtype=np.int8)
  pool = np.empty(pool_cap, dtype=np.uint8)
  n = lib().orc_create_node_list(_p(ref, C.c_char), C.c_int64(ref.size), C.c_int64(ref_start_pos), *cv.args(),
                                 C.c_int64(cap), _p(ps, C.c_int64), _p(pr, C.c_int64), _p(op, C.c_char),
                                 _p(oplen, C.c_int64), _p(v, C.c_int64), _p(has_v, C.c_int8),
                                 _p(seq_off, C.c_int64), _p(pool, C.c_char), C.c_int64(pool_cap))
  assert n >= 0
  return [(int(ps[i]), int(pr[i]), chr(op[i]), int(oplen[i]),
           pool[seq_off[i]:seq_off[i + 1]].tobytes().decode(), int(v[i]) if has_v[i] else None) for i in range(n)]


def generate_read(ref_seq, ref_start_pos, cv, p, l, n0=-1, n1=-1):
  """-> (pos, cigar, v_list, seq, n0, n1); n0/n1 < 0 means 'look them up' (a11)."""
  ref = _ref(ref_seq)
  cap = int(l) * 24 + ref.size + 64
  cig, vl, sq = (C.create_string_buffer(cap) for _ in range(3))
  a, b = C.c_int64(n0), C.c_int64(n1)
  pos = lib().orc_generate_read(_p(ref, C.c_char), C.c_int64(ref.size), C.c_int64(ref_start_pos), *cv.args(),
                                C.c_int64(p), C.c_int64(l), C.byref(a), C.byref(b), cig, vl, sq, C.c_int64(cap))
  v_list = [int(x) for x in vl.value.decode().split(',') if x != '']
  return int(pos), cig.value.decode(), v_list, sq.value.decode(), int(a.value), int(b.value)


# -- a9 / a10 ------------------------------------------------------------------------------------

def templates(p, rlen, cum_tlen, p_min, p_max, unit_seed):
  """-> (ts, te, fo) of kept templates, as illumina.generate_reads computes them."""
  cum_tlen = np.ascontiguousarray(cum_tlen, dtype=np.float64)
  cap = int((p_max - p_min) * p * 1.2) + 1
  ts, te = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64)
  fo = np.empty(cap, dtype=np.int8)
  n_est = C.c_int64(0)
  k = lib().orc_templates(C.c_double(p), C.c_int64(rlen), _p(cum_tlen, C.c_double), C.c_int64(cum_tlen.size),
                          C.c_int64(p_min), C.c_int64(p_max), C.c_uint32(unit_seed), C.c_int64(cap),
                          _p(ts, C.c_int64), _p(te, C.c_int64), _p(fo, C.c_int8), C.byref(n_est))
  assert k >= 0
  return ts[:k].copy(), te[:k].copy(), fo[:k].copy()


# -- a13 / a14 -----------------------------------------------------------------------------------

def generate_unit(ref_seq, ref_start_pos, cv, rlen, ts, te, fo, stub, chrom, cpy, cap=None):
  """-> (fastq1 bytes, fastq2 bytes, template count) for one work unit."""
  ref = _ref(ref_seq)
  ts, te = _i64(ts), _i64(te)
  fo = np.ascontiguousarray(fo, dtype=np.int8)
  if cap is None:
    cap = int(ts.size) * (2 * int(rlen) + 160 + len(stub) + len(chrom)) + 4096
  while True:
    o1, o2 = np.empty(cap, dtype=np.uint8), np.empty(cap, dtype=np.uint8)
    l1, l2 = C.c_int64(0), C.c_int64(0)
    n = lib().orc_generate_unit(_p(ref, C.c_char), C.c_int64(ref.size), C.c_int64(ref_start_pos), *cv.args(),
                                C.c_int64(rlen), C.c_int64(ts.size), _p(ts, C.c_int64), _p(te, C.c_int64), _p(fo, C.c_int8),
                                stub.encode(), chrom.encode(), C.c_int(cpy),
                                _p(o1, C.c_char), _p(o2, C.c_char), C.c_int64(cap), C.byref(l1), C.byref(l2))
    if n >= 0:
      return o1[:l1.value].tobytes(), o2[:l2.value].tobytes(), int(n)
    cap *= 2


# -- a16 - a18 -----------------------------------------------------------------------------------

def corrupt_fastq(cum_bq_mat, worker_seed, fq1, fq2=None):
  """corrupt-reads with one worker -> (out1 bytes, out2 bytes or None, template count)."""
  m = np.ascontiguousarray(cum_bq_mat, dtype=np.float64)
  a1 = np.frombuffer(fq1, dtype=np.uint8)
  a2 = np.frombuffer(fq2, dtype=np.uint8) if fq2 is not None else None
  cap = int(a1.size + (a2.size if a2 is not None else 0)) + 64
  o1 = np.empty(cap, dtype=np.uint8)
  o2 = np.empty(cap if a2 is not None else 1, dtype=np.uint8)
  l1, l2 = C.c_int64(0), C.c_int64(0)
  n = lib().orc_corrupt_fastq(_p(m, C.c_double), C.c_int64(m.shape[1]), C.c_int64(m.shape[2]), _p(PHRED_P, C.c_double),
                              C.c_uint32(worker_seed), _p(a1, C.c_char), C.c_int64(a1.size),
                              _p(a2, C.c_char) if a2 is not None else None, C.c_int64(a2.size if a2 is not None else 0),
                              _p(o1, C.c_char), _p(o2, C.c_char) if a2 is not None else None, C.c_int64(cap),
                              C.byref(l1), C.byref(l2))
  if n == -2:
    raise IndexError('read longer than the model (illumina.py:156)')
  assert n >= 0, n
  return o1[:l1.value].tobytes(), (o2[:l2.value].tobytes() if a2 is not None else None), int(n)


# -- whole-command restatements (threads=1 semantics) ---------------------------------------------

def node_span(ref_seq, ref_start_pos, cv):
  ref = _ref(ref_seq)
  a, b, op = C.c_int64(0), C.c_int64(0), C.c_int(0)
  lib().orc_node_span.restype = C.c_int64
  n = lib().orc_node_span(_p(ref, C.c_char), C.c_int64(ref.size), C.c_int64(ref_start_pos), *cv.args(),
                          C.byref(a), C.byref(b), C.byref(op))
  return int(a.value), int(b.value), int(n)


def read_model_params(model, diploid_coverage=30.0):
  """a5, illumina.py:12-40."""
  rlen = int(model['mean_rlen'])
  p, passes = 1.0, 1
  while p > 0.1:
    passes *= 2
    p = 0.5 * diploid_coverage / (2 * rlen * passes)
  return {'diploid_coverage': diploid_coverage, 'p': p, 'passes': passes, 'rlen': rlen,
          'cum_tlen': model['cum_tlen'], 'cum_bq_mat': model['cum_bq_mat']}


def generate_reads_cmd(regions, model, coverage, seed, sample, unit_filter=None):
  """generate-reads with --threads 1 (readgenerate.py:76-253).

  regions: [{'region': (chrom, start, end), 'ref': uint8 array of that region, 'v': [CopyVariants]}]
  -> (fastq1 bytes, fastq2 bytes, pairs)
  """
  rm = read_model_params(model, coverage)
  units = [(ri, cpy) for ri, r in enumerate(regions) for cpy in range(len(r['v'])) for _ in range(rm['passes'])]
  seeds, order = unit_schedule(seed, len(units))
  o1, o2, total = [], [], 0
  for ps, k in enumerate(order.tolist()):
    if unit_filter is not None and not unit_filter(ps):
      continue
    ri, cpy = units[k]
    r = regions[ri]
    chrom, start, _ = r['region']
    p_min, p_max, _ = node_span(r['ref'], start + 1, r['v'][cpy])
    ts, te, fo = templates(rm['p'], rm['rlen'], rm['cum_tlen'], p_min, p_max, int(seeds[k]))
    a, b, n = generate_unit(r['ref'], start + 1, r['v'][cpy], rm['rlen'], ts, te, fo,
                            '{}:{}:{}'.format(sample, 0, ps), chrom, cpy)
    o1.append(a); o2.append(b); total += n
  return b''.join(o1), b''.join(o2), total


def corrupt_reads_cmd(model, seed, fq1, fq2=None):
  """corrupt-reads with --threads 1 (readcorrupt.py:18-118)."""
  ws = int(corrupt_worker_seeds(seed, 1)[0])
  return corrupt_fastq(model['cum_bq_mat'], ws, fq1, fq2)
