"""Read POS / CIGAR: the device-backed mirror of ``mitty/simulation/rpc.py``.

``create_node_list`` builds the node table of one chromosome copy (the greedy variant walk, the
haplotype and the lookup table all on the GPU: ``k_walk_*``, ``k_hap_build``, ``k_blk_table``) and returns the same tuples as the reference's ``Node``s;
``generate_read`` evaluates one read through the emit kernel and parses the result back.  These are
convenience entry points for tests and for users of the reference's low-level API; the hot path
(``readgenerate``) drives the same kernels unit-wise.
"""
import numpy as np

from mitty_b200.engine import MODE_EXPLICIT, default_engine
from mitty_b200.lib.vcfio import VariantList


class NodeList(object):
  """Device-resident node list of one copy + the reference-compatible view of it."""

  def __init__(self, engine, region_id, copy, ref_seq):
    self.engine, self.region_id, self.copy = engine, region_id, copy
    self._ref = ref_seq

  def tuples(self):
    """[(ps, pr, op, oplen, seq, v)] exactly as rpc.Node.tuple() (rpc.py:22-23)."""
    ps, pr, op, oplen = self.engine.copy_nodes(self.copy)
    hap = self.engine.copy_haplotype(self.copy).tobytes().decode()
    out = []
    for i in range(ps.size):
      o = chr(op[i])
      a = int(ps[i]) - self.copy.p_min
      seq = '' if o == 'D' else hap[a:a + int(oplen[i])]
      v = {'=': None, 'X': 0, 'I': int(oplen[i]), 'D': -int(oplen[i])}[o]
      out.append((int(ps[i]), int(pr[i]), o, int(oplen[i]), seq, v))
    return out

  def __len__(self):
    return self.copy.n_nodes

  def free(self):
    self.engine.free_copy(self.copy)
    self.engine.free_region(self.region_id)


def create_node_list(ref_seq, ref_start_pos, vl, engine=None):
  """ref_seq: str/bytes/uint8 array of the region; ref_start_pos 1-based; vl: list of Variant-like
  objects or a VariantList (rpc.py:38-63)."""
  eng = engine or default_engine()
  if isinstance(ref_seq, str):
    ref_seq = ref_seq.encode()
  ref = np.frombuffer(ref_seq, dtype=np.uint8) if not isinstance(ref_seq, np.ndarray) else ref_seq
  if not isinstance(vl, VariantList):
    vl = VariantList.from_variants(list(vl))
  rid = eng.load_region(ref, ref_start_pos - 1)
  cp = eng.build_copy(rid, vl)
  return NodeList(eng, rid, cp, ref)


def generate_read(p, l, nodes, cum_tlen=None):
  """(pos, cigar, v_list, seq) of the read of length l starting at sample position p
  (rpc.generate_read, rpc.py:133-160, with the node lookup of rpc.py:119-130 done on the device)."""
  eng = nodes.engine
  eng.load_model({'cum_tlen': np.array([1.0]) if cum_tlen is None else cum_tlen,
                  'cum_bq_mat': np.ones((2, max(1, l), 1)), 'rlen': l})
  f1, f2, nt, _, _ = eng.generate_unit(nodes.copy, 1, 0.5, MODE_EXPLICIT, 0, '@r:0:0:', '|c|0',
                                       ts=np.array([p], dtype=np.int64), tl=np.array([l], dtype=np.int64),
                                       fo=np.zeros(1, dtype=np.int8))
  if nt == 0:
    return None
  lines = f1.tobytes().decode().split('\n')
  d = lines[0].split('|')
  return int(d[4]), d[6], [int(x) for x in d[7].split(',') if x != ''], lines[1]
