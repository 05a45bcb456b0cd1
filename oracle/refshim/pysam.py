"""Minimal pure-Python stand-in for the three pysam classes Mitty's hot path uses.

TEST INFRASTRUCTURE ONLY.  pysam/htslib is not installed in this image, and the unmodified
reference (``/root/reference/mitty``) imports it at module top in ``readgenerate.py``,
``readcorrupt.py`` and ``lib/vcfio.py``.  ``tests/golden/make_golden.py`` puts this directory on
``sys.path`` so that the *unmodified reference code* can run here and produce golden vectors.
Nothing in the product imports this file.

Only I/O semantics are provided, as the reference uses them:

* ``FastaFile(f).fetch(reference=, start=, end=)``        (readgenerate.py:181,186)
* ``VariantFile(f, mode)``, ``.subset_samples([s])``, ``.fetch(contig=, start=, stop=)`` yielding
  records with ``.pos .ref .samples[0]['GT'] .samples[0].alleles``   (vcfio.py:59-62,112-126)
  with htslib's region-overlap rule: a record at 1-based POS occupies 0-based
  ``[POS-1, POS-1+len(REF))`` and is returned iff that interval overlaps ``[start, stop)``
  (pinned by the reference's own test_vcfio.py:9-18).
* ``FastxFile(f)`` yielding ``.name .sequence .quality``  (readcorrupt.py:49-55)
* ``VariantFile(f, mode='w', header=...)`` + ``.write(record)`` as ``prepare_variant_file`` uses them
  (vcfio.py:156-164): the meta lines, a column header naming the subset sample, and each written
  record's first nine columns plus that sample's column.
"""
import gzip


def _open_text(fname):
  if str(fname).endswith('.gz'):
    return gzip.open(fname, 'rt')
  return open(fname, 'r')


class FastaFile(object):
  def __init__(self, fname):
    self._seqs = {}
    name, chunks = None, []
    with _open_text(fname) as fp:
      for line in fp:
        if line.startswith('>'):
          if name is not None:
            self._seqs[name] = ''.join(chunks)
          name, chunks = line[1:].split()[0], []
        else:
          chunks.append(line.strip())
    if name is not None:
      self._seqs[name] = ''.join(chunks)

  def fetch(self, reference=None, start=None, end=None):
    return self._seqs[reference][start:end]


class _Sample(object):
  def __init__(self, gt, alleles):
    self._gt, self.alleles = gt, alleles

  def __getitem__(self, k):
    if k == 'GT':
      return self._gt
    raise KeyError(k)


class _Samples(object):
  def __init__(self, s):
    self._s = s

  def __getitem__(self, i):
    if i == 0:
      return self._s
    raise IndexError(i)

  def values(self):
    return [self._s]


class _Record(object):
  __slots__ = ('contig', 'pos', 'ref', 'alts', 'samples', 'rlen', 'fields')

  def __init__(self, contig, pos, ref, alts, gt, fields=None):
    self.contig, self.pos, self.ref, self.alts = contig, pos, ref, alts
    self.fields = fields      # first nine columns + the subset sample's column (for VariantFile.write)
    self.rlen = len(ref)
    all_alleles = (ref,) + alts
    self.samples = _Samples(_Sample(gt, tuple(all_alleles[g] if g is not None else None for g in gt)))


class VariantFile(object):
  def __init__(self, fname, mode='r', header=None):
    self._fname = fname
    self._sample_col = None
    self._rows = []  # (contig, pos, ref, alts, [sample fields...], fmt)
    self._samples = []
    self._meta = []
    self._out = None
    if mode.startswith('w'):
      # header = the input VariantFile (its .header is itself here): meta lines + the subset sample
      import sys
      self._out = sys.stdout if fname == '-' else open(fname, 'w')
      self._out.write(''.join(header._meta))
      names = [header._samples[header._subset]] if header._subset is not None else header._samples
      self._out.write('#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t' + '\t'.join(names) + '\n')
      return
    with _open_text(fname) as fp:
      for line in fp:
        if line.startswith('##'):
          self._meta.append(line)
          continue
        f = line.rstrip('\n').split('\t')
        if line.startswith('#'):
          self._samples = f[9:]
          continue
        if len(f) < 10:
          continue
        self._rows.append(f)
    self._subset = None

  def subset_samples(self, names):
    self._subset = self._samples.index(names[0])

  @property
  def header(self):
    return self

  def write(self, rec):
    self._out.write('\t'.join(rec.fields) + '\n')

  def close(self):
    if self._out is not None:
      self._out.flush()

  def __del__(self):
    try:
      if self._out is not None:
        self._out.flush()
    except Exception:
      pass

  def fetch(self, contig=None, start=None, stop=None):
    col = 9 + (self._subset if self._subset is not None else 0)
    for f in self._rows:
      if f[0] != contig:
        continue
      pos = int(f[1])
      ref = f[3]
      p0 = pos - 1
      if not (p0 < stop and p0 + len(ref) > start):
        continue
      alts = tuple(f[4].split(','))
      fmt = f[8].split(':')
      gt_s = f[col].split(':')[fmt.index('GT')]
      gt = tuple(None if g == '.' else int(g) for g in gt_s.replace('/', '|').split('|'))
      yield _Record(contig, pos, ref, alts, gt, f[:9] + [f[col]])


class _Fastx(object):
  __slots__ = ('name', 'comment', 'sequence', 'quality')


class FastxFile(object):
  def __init__(self, fname):
    self._fp = _open_text(fname)

  def __iter__(self):
    return self

  def __next__(self):
    h = self._fp.readline()
    if not h:
      raise StopIteration
    s = self._fp.readline().rstrip('\n')
    self._fp.readline()
    q = self._fp.readline().rstrip('\n')
    r = _Fastx()
    parts = h[1:].rstrip('\n').split(None, 1)
    r.name = parts[0] if parts else ''
    r.comment = parts[1] if len(parts) > 1 else None
    r.sequence, r.quality = s, q
    return r
