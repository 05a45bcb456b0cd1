"""Numpy restatement of the engine's PRODUCTION-mode corruption (the draw layout documented at
MgCorruptCtx in mitty_b200/csrc/mg_core.cuh): Philox4x32 counters, per-cycle miscall thresholds,
Vose alias rows of the quality given a correct call / a miscall.  Test infrastructure: lets the fused GPU path be checked byte for byte, not only
statistically."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
STREAM_CORRUPT = 0x636f7272
CORRUPT_ROUNDS = 7   # MG_CORRUPT_ROUNDS: the per-base corruption stream is Philox4x32-7
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
  return philox4x32(c0, c1, c2, c3, k0, k1, 10)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
  c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
  k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
  for _ in range(rounds):
    p0, p1 = M0 * c0, M1 * c2
    h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
    c0, c1, c2, c3 = h1 ^ c1 ^ np.uint64(k0), l1, h0 ^ c3 ^ np.uint64(k1), l0
    k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
  return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def exact64_cycles(cum_bq_mat):
  """number of leading cycles whose rows (all mates) put no mass on BQ >= 64"""
  m = np.asarray(cum_bq_mat, dtype=np.float64)
  n64 = m.shape[1]
  for mi in range(m.shape[0]):
    for ci in range(m.shape[1]):
      row = m[mi, ci]
      at63 = row[63] if row.size > 63 else row[-1]
      if at63 < 1.0:
        n64 = min(n64, ci)
        break
  return n64


def _vose(q, K):
  """Vose's alias method, the same operation order as vose() in mg_api.cu (entry = prob24 << 8 | alias; prob24 is capped at 2^24 - 1, as any
  entry that fills its 32 bits is)."""
  q = list(q)
  small, large = [], []
  for i in range(K):
    q[i] *= K
    (small if q[i] < 1.0 else large).append(i)
  prob, alias = [1.0] * K, list(range(K))
  while small and large:
    s, l = small.pop(), large.pop()
    prob[s], alias[s] = q[s], l
    q[l] = (q[l] + q[s]) - 1.0
    (small if q[l] < 1.0 else large).append(l)
  out = np.zeros(K, dtype=np.uint32)
  for i in range(K):
    pr = min(max(prob[i], 0.0), 1.0)
    pq = min(int(np.floor(pr * 16777216.0 + 0.5)), (1 << 24) - 1)
    out[i] = (pq << 8) | alias[i]
  return out


def quality_tables(cum_bq_mat, phred_p, kshift, n_rows=None):
  """-> (alias u32[n_mates, n_cycles, 2, 1 << kshift], thr u32[n_mates, n_cycles]): per (mate, cycle)
  the miscall threshold floor(perr * 2^32), perr = sum_q P(q) phred_p[q], and the alias rows of the
  quality given a correct call ([0]) / given a miscall ([1]) -- build_quality_rows in mg_api.cu, same
  operation order.  Only the first n_rows cycles are built (the rest stay zero)."""
  m = np.asarray(cum_bq_mat, dtype=np.float64)
  n_mates, n_cycles, n_bq = m.shape
  K = 1 << kshift
  out = np.zeros((n_mates, n_cycles, 2, K), dtype=np.uint32)
  thr = np.zeros((n_mates, n_cycles), dtype=np.uint32)
  for mi in range(n_mates):
    for ci in range(n_cycles if n_rows is None else min(n_rows, n_cycles)):
      row = m[mi, ci]
      q = [0.0] * K
      prev = 0.0
      for b in range(n_bq):
        c = min(float(row[b]), 1.0)
        if c < prev:
          c = prev
        if min(b, 93) < K:
          q[min(b, 93)] += c - prev
        prev = c
      if min(n_bq, 93) < K:
        q[min(n_bq, 93)] += 1.0 - prev
      qe, qo, se, so = [0.0] * K, [0.0] * K, 0.0, 0.0
      for k in range(K):
        p = float(phred_p[k]) if k < 100 else 0.0
        qe[k] = q[k] * p
        qo[k] = q[k] * (1.0 - p)
        se += qe[k]
        so += qo[k]
      for k in range(K):
        qe[k] = qe[k] / se if se > 0.0 else q[k]
        qo[k] = qo[k] / so if so > 0.0 else q[k]
      out[mi, ci, 0] = _vose(qo, K)
      out[mi, ci, 1] = _vose(qe, K)
      thr[mi, ci] = 0xFFFFFFFF if se >= 1.0 else (0 if se <= 0.0 else int(np.floor(se * 4294967296.0)))
  return out, thr


def alias_distribution(alias_row, kshift):
  """P(bq) encoded by one alias row (exact rational arithmetic on the 24-bit thresholds)."""
  K = 1 << kshift
  p = np.zeros(128)
  for i in range(K):
    pq, al = int(alias_row[i]) >> 8, int(alias_row[i]) & 255
    p[i] += pq / float(1 << 24) / K
    p[al] += (1.0 - pq / float(1 << 24)) / K
  return p


def joint_distribution(alias_pair, thr, kshift):
  """(P(bq, correct), P(bq, miscall)) encoded by one cycle's two rows and its threshold."""
  pe = int(thr) / 4294967296.0
  return (1.0 - pe) * alias_distribution(alias_pair[0], kshift), pe * alias_distribution(alias_pair[1], kshift)


ROT = {ord('A'): b'CTG', ord('C'): b'ATG', ord('T'): b'ACG', ord('G'): b'ACT'}


def corrupt_file(fq, f, alias, kshift, thr, k0, k1, serials=None):
  """Corrupt a perfect FASTQ buffer (file index f) -> bytes.  alias, thr: quality_tables().
  serials: per-record template serial (default 0..n-1, the standalone kernel's numbering)."""
  lines = fq.split(b'\n')
  n_rec = (len(lines) - 1) // 4
  if serials is None:
    serials = np.arange(n_rec, dtype=np.uint64)
  out = []
  L = max(len(lines[4 * r + 1]) for r in range(n_rec)) if n_rec else 0
  nq = (L + 1) // 2
  s_grid = np.repeat(np.asarray(serials, dtype=np.uint64), nq)
  q_grid = np.tile(np.arange(nq, dtype=np.uint64), n_rec)
  r = philox4x32(s_grid & MASK, (s_grid >> np.uint64(32)) * np.uint64(2) + np.uint64(f), q_grid, STREAM_CORRUPT, k0, k1, CORRUPT_ROUNDS)
  w = np.stack(r, axis=1).reshape(n_rec, nq, 4)
  for rec in range(n_rec):
    seq = bytearray(lines[4 * rec + 1])
    Lr = len(seq)
    w_bq = w[rec, :, 0::2].reshape(-1)[:Lr].astype(np.uint64)     # cycle 2q -> r[0], 2q+1 -> r[2]
    w_call = w[rec, :, 1::2].reshape(-1)[:Lr].astype(np.uint64)
    T = thr[f, :Lr].astype(np.uint64)
    miss = w_call < T
    idx = (w_bq >> np.uint64(32 - kshift)).astype(np.int64)
    frac = ((w_bq << np.uint64(kshift)) & MASK) >> np.uint64(8)
    e = alias[f, np.arange(Lr), miss.astype(np.int64), idx].astype(np.uint64)
    bq = np.where(frac < (e >> np.uint64(8)), idx, (e & np.uint64(255)).astype(np.int64))
    rot = (w_call >= T // np.uint64(3)).astype(np.int64) + (w_call >= (np.uint64(2) * T) // np.uint64(3)).astype(np.int64)
    for n in np.flatnonzero(miss):
      seq[n] = ROT.get(seq[n], b'NNN')[rot[n]]
    out.append(lines[4 * rec] + b'\n' + bytes(seq) + b'\n+\n' + (bq + 33).astype(np.uint8).tobytes() + b'\n')
  return b''.join(out)
