"""Host-side front half of the haplotype builder: BED / VCF / FASTA -> per-region, per-copy
variant arrays.  Mirrors ``mitty/lib/vcfio.py`` (reference) function for function:

* ``read_bed``            <- vcfio.py:45-46   (0-based half-open, file order preserved)
* ``load_variant_file``   <- vcfio.py:51-64   (records overlapping each BED region, htslib rule)
* ``split_copies``        <- vcfio.py:67-102  (ploidy = len(GT) of the FIRST record; empty -> 2)
* ``parse`` / ``Variant`` <- vcfio.py:19-33,105-126  (X / I / D classification, ValueError on
                                                      complex variants)

pysam/htslib is not a dependency: VCF text (plain or gzip/bgzip) is parsed here, and the htslib
region-overlap rule (a record at 1-based POS occupies 0-based [POS-1, POS-1+len(REF)) and is
returned iff that overlaps [start, stop)) is applied with numpy (the reference's own
test_vcfio.py:9-18 pins it: deletion at 11 is returned for [8,14), insertion at 8 is not).

The per-copy result is a ``VariantList``: flat numpy arrays ready for the C-ABI
(``mg_copy_build``), which also behaves like the reference's ``list`` of ``Variant`` objects.
"""
import gzip
import logging
import threading

import numpy as np

logger = logging.getLogger(__name__)


class Variant(object):
  """Same fields and tuple() as the reference's Variant (vcfio.py:19-33)."""
  __slots__ = ('pos', 'ref', 'alt', 'cigarop', 'oplen')

  def __init__(self, pos, ref, alt, cigarop, oplen):
    self.pos, self.ref, self.alt, self.cigarop, self.oplen = pos, ref, alt, cigarop, oplen

  def tuple(self):
    return self.pos, self.ref, self.alt, self.cigarop, self.oplen

  def __repr__(self):
    return self.tuple().__repr__()


class VariantList(object):
  """Variants on one chromosome copy, as flat arrays (pos, op, oplen, pooled REF/ALT bytes)."""
  __slots__ = ('pos', 'op', 'oplen', 'alt_pool', 'alt_off', 'ref_pool', 'ref_off')

  def __init__(self, pos, op, oplen, alt_pool, alt_off, ref_pool=None, ref_off=None):
    self.pos = np.ascontiguousarray(pos, dtype=np.int64)
    self.op = np.ascontiguousarray(op, dtype=np.uint8)
    self.oplen = np.ascontiguousarray(oplen, dtype=np.int64)
    self.alt_pool = np.ascontiguousarray(alt_pool, dtype=np.uint8)
    self.alt_off = np.ascontiguousarray(alt_off, dtype=np.int64)
    self.ref_pool, self.ref_off = ref_pool, ref_off

  @classmethod
  def from_variants(cls, vl):
    """From a list of Variant-like objects (pos, ref, alt, cigarop, oplen)."""
    alts = [v.alt.encode() for v in vl]
    refs = [v.ref.encode() for v in vl]
    alt_off = np.zeros(len(vl) + 1, dtype=np.int64); np.cumsum([len(a) for a in alts], out=alt_off[1:])
    ref_off = np.zeros(len(vl) + 1, dtype=np.int64); np.cumsum([len(a) for a in refs], out=ref_off[1:])
    return cls([v.pos for v in vl], np.frombuffer(''.join(v.cigarop for v in vl).encode(), dtype=np.uint8),
               [v.oplen for v in vl], np.frombuffer(b''.join(alts), dtype=np.uint8), alt_off,
               np.frombuffer(b''.join(refs), dtype=np.uint8), ref_off)

  def __len__(self):
    return int(self.pos.shape[0])

  def __getitem__(self, i):
    if isinstance(i, slice):
      return [self[j] for j in range(*i.indices(len(self)))]
    if i < 0:
      i += len(self)
    if not 0 <= i < len(self):
      raise IndexError(i)
    ref = '' if self.ref_pool is None else self.ref_pool[self.ref_off[i]:self.ref_off[i + 1]].tobytes().decode()
    alt = self.alt_pool[self.alt_off[i]:self.alt_off[i + 1]].tobytes().decode()
    return Variant(int(self.pos[i]), ref, alt, chr(self.op[i]), int(self.oplen[i]))

  def __iter__(self):
    return (self[i] for i in range(len(self)))


def read_bed(bed_fname):
  """BED -> [(chrom, start, end)], whitespace split, file order preserved (vcfio.py:45-46)."""
  out = []
  with open(bed_fname, 'r') as fp:
    for line in fp.readlines():
      x = line.split()
      out.append((x[0], int(x[1]), int(x[2])))
  return out


def _open_bytes(fname):
  with open(fname, 'rb') as fp:
    magic = fp.read(2)
  if magic == b'\x1f\x8b':
    with gzip.open(fname, 'rb') as fp:  # also reads bgzip (concatenated gzip members)
      return fp.read()
  with open(fname, 'rb') as fp:
    return fp.read()


class VcfTable(object):
  """All records of one sample, grouped by contig, as arrays.

  Per contig: pos int64[n], reflen int64[n], ref list[str], alleles list[tuple[str]] (REF first),
  gt list[tuple[int]] (as written for the sample; '.' -> None).
  """

  def __init__(self, fname, sample):
    if str(fname).endswith('bcf'):
      raise NotImplementedError('BCF input needs htslib; convert to VCF text (plain or gzip)')
    text = _open_bytes(fname).decode()
    self.contigs = {}
    col = None
    cur_name, cur = None, None
    for line in text.split('\n'):
      if not line or line.startswith('##'):
        continue
      if line.startswith('#'):
        hdr = line.split('\t')
        if sample not in hdr[9:]:
          raise ValueError('Sample {} not in VCF (samples: {})'.format(sample, hdr[9:]))
        col = 9 + hdr[9:].index(sample)
        continue
      f = line.split('\t')
      if len(f) <= (col or 9):
        continue
      if f[0] != cur_name:
        cur_name = f[0]
        cur = self.contigs.setdefault(cur_name, ([], [], [], []))
      fmt = f[8]
      s = f[col]
      if fmt != 'GT':
        s = s.split(':')[fmt.split(':').index('GT')]
      gt = tuple(None if g == '.' else int(g) for g in s.replace('/', '|').split('|'))
      cur[0].append(int(f[1])); cur[1].append(f[3]); cur[2].append((f[3],) + tuple(f[4].split(','))); cur[3].append(gt)
    self._arr = {}
    for name, (pos, ref, alleles, gt) in self.contigs.items():
      self._arr[name] = (np.array(pos, dtype=np.int64), np.array([len(r) for r in ref], dtype=np.int64))

  def fetch(self, contig, start, stop):
    """Indices of records overlapping 0-based [start, stop) -- htslib semantics (vcfio.py:62)."""
    if contig not in self._arr:
      return contig, np.zeros(0, dtype=np.int64)
    pos, reflen = self._arr[contig]
    p0 = pos - 1
    return contig, np.flatnonzero((p0 < stop) & (p0 + reflen > start))


def parse_copy(table, contig, idx, cpy):
  """vcfio.parse over the records ``idx`` for copy ``cpy`` -> VariantList (vcfio.py:105-126)."""
  if idx.size == 0:
    return VariantList([], np.zeros(0, np.uint8), [], np.zeros(0, np.uint8), np.zeros(1, np.int64),
                       np.zeros(0, np.uint8), np.zeros(1, np.int64))
  pos_l, ref_l, alleles_l, gt_l = table.contigs[contig]
  pos, op, oplen, alts, refs = [], [], [], [], []
  for i in idx.tolist():
    g = gt_l[i][cpy]                       # IndexError on ragged ploidy, like the reference
    if g == 0:                             # not present on this copy (vcfio.py:112)
      continue
    if g is None:
      raise ValueError('Missing GT allele at {}:{}'.format(contig, pos_l[i]))
    ref, alt = ref_l[i], alleles_l[i][g]
    l_r, l_a = len(ref), len(alt)
    if l_r == 1:
      if l_a == 1:
        o, ol = 88, 0                      # 'X'
      else:
        o, ol = 73, l_a - l_r              # 'I'
    elif l_a == 1:
      o, ol = 68, l_r - l_a                # 'D'
    else:
      raise ValueError("Complex variants present in VCF. Please filter or refactor these.")
    pos.append(pos_l[i]); op.append(o); oplen.append(ol); alts.append(alt.encode()); refs.append(ref.encode())
  alt_off = np.zeros(len(pos) + 1, dtype=np.int64); np.cumsum([len(a) for a in alts], out=alt_off[1:])
  ref_off = np.zeros(len(pos) + 1, dtype=np.int64); np.cumsum([len(a) for a in refs], out=ref_off[1:])
  return VariantList(pos, np.array(op, dtype=np.uint8), oplen, np.frombuffer(b''.join(alts), dtype=np.uint8), alt_off,
                     np.frombuffer(b''.join(refs), dtype=np.uint8), ref_off)


def split_copies(region, table, contig, idx):
  """One VariantList per chromosome copy; ploidy sniffed from the first record (vcfio.py:67-102)."""
  if idx.size == 0:
    logger.warning('Empty region ({}), assuming diploid'.format(region))
    ploidy = 2
  else:
    ploidy = len(table.contigs[contig][3][int(idx[0])])
    logger.debug('Region: {}, ploidy: {}'.format(region, ploidy))
  return {'region': region, 'v': [parse_copy(table, contig, idx, cpy) for cpy in range(ploidy)]}


def load_variant_file(fname, sample, bed_fname):
  """VCF + BED -> [{'region': (chrom, start, end), 'v': [VariantList per copy]}] (vcfio.py:51-64)."""
  table = VcfTable(fname, sample)
  return [split_copies(region, table, *table.fetch(region[0], region[1], region[2]))
          for region in read_bed(bed_fname)]


def from_variant_table(vt, region):
  """Same structure from an in-memory synthetic ``mitty_b200.synth.VariantTable`` (single ALT)."""
  chrom, start, stop = region
  p0 = vt.pos - 1
  reflen = vt.ref_off[1:] - vt.ref_off[:-1]
  altlen = vt.alt_off[1:] - vt.alt_off[:-1]
  sel = (p0 < stop) & (p0 + reflen > start)
  if ((reflen > 1) & (altlen > 1))[sel].any():
    raise ValueError("Complex variants present in VCF. Please filter or refactor these.")
  ploidy = vt.gt.shape[1] if sel.any() else 2
  op = np.where(reflen == 1, np.where(altlen == 1, 88, 73), 68).astype(np.uint8)
  oplen = np.where(reflen == 1, altlen - 1, reflen - 1)
  out = []
  for cpy in range(ploidy):
    idx = np.flatnonzero(sel & (vt.gt[:, cpy] != 0)) if sel.any() else np.zeros(0, dtype=np.int64)
    def pool(p, off):
      ln = (off[1:] - off[:-1])[idx]
      noff = np.zeros(idx.size + 1, dtype=np.int64); np.cumsum(ln, out=noff[1:])
      src = np.repeat(off[:-1][idx] - noff[:-1], ln) + np.arange(noff[-1])
      return p[src] if src.size else np.zeros(0, dtype=np.uint8), noff
    ap, ao = pool(vt.alt_pool, vt.alt_off)
    rp, ro = pool(vt.ref_pool, vt.ref_off)
    out.append(VariantList(vt.pos[idx], op[idx], oplen[idx], ap, ao, rp, ro))
  return {'region': region, 'v': out}


class FastaFile(object):
  """FASTA reader with pysam.FastaFile's fetch(reference=, start=, end=) (readgenerate.py:181,186).
  Sequences come back as uint8 arrays of the file's bytes (case and IUPAC codes preserved)."""

  def __init__(self, fname):
    data = _open_bytes(fname)
    self._seqs = {}
    a = np.frombuffer(data, dtype=np.uint8)
    hdr = np.flatnonzero(a == ord('>'))
    # only '>' at line starts are headers
    hdr = hdr[(hdr == 0) | (a[np.maximum(hdr - 1, 0)] == 10)]
    bounds = list(hdr) + [a.size]
    for k, h in enumerate(hdr.tolist()):
      eol = data.find(b'\n', h)
      if eol < 0:
        eol = len(data)
      name = data[h + 1:eol].split()[0].decode() if eol > h + 1 else ''
      self._seqs[name] = (eol + 1, bounds[k + 1])
    self._data = data
    self._cache = {}
    self._lock = threading.Lock()

  def _contig(self, name):
    with self._lock:   # one worker thread per GPU may fetch concurrently
      if name not in self._cache:
        s, e = self._seqs[name]
        self._cache[name] = np.frombuffer(self._data[s:e].translate(None, b'\r\n'), dtype=np.uint8)
      return self._cache[name]

  def fetch(self, reference=None, start=None, end=None):
    return self._contig(reference)[start:end]
