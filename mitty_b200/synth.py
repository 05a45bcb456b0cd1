"""Seeded synthetic genomes, VCFs and BED files of the shapes BASELINE.json names.

There is no network for real references or call sets, so every workload the tests and
``bench.py`` run is synthesised here from ``numpy.random.RandomState`` (stream frozen by numpy
policy, so the same seed gives the same bytes on every box).  Shapes follow SURVEY.md 8(d):

* contigs of uppercase ACGT with optional long N runs (GRCh37-like gaps),
* GIAB-after-filter variant density (~740 variants per Mb), ~85 % SNP / 7.5 % INS / 7.5 % DEL,
  indel length Geom(0.3), a handful of >= 151 bp insertions so reads from inside a long insertion
  (the '>p:nI' CIGAR, reference ``mitty/simulation/rpc.py:148-154``) occur,
* variants spaced >= 40 bp apart, none within 1 kb of a region end (so no deletion crosses a
  region end, the unchecked assumption at ``mitty/simulation/readgenerate.py:192``).

Everything is returned as flat numpy arrays (positions, pooled allele bytes, GT matrix) so that a
249 Mb / 330 k-variant contig takes seconds; text writers (FASTA / VCF / BED) sit on top for the
file-based CLI path.
"""
import gzip
import io

import numpy as np

BASES = np.frombuffer(b'ACGT', dtype=np.uint8)

# GRCh37 primary-assembly contig lengths (config 4 of BASELINE.json).
GRCH37_CONTIGS = [
  ('1', 249250621), ('2', 243199373), ('3', 198022430), ('4', 191154276), ('5', 180915260),
  ('6', 171115067), ('7', 159138663), ('8', 146364022), ('9', 141213431), ('10', 135534747),
  ('11', 135006516), ('12', 133851895), ('13', 115169878), ('14', 107349540), ('15', 102531392),
  ('16', 90354753), ('17', 81195210), ('18', 78077248), ('19', 59128983), ('20', 63025520),
  ('21', 48129895), ('22', 51304566), ('X', 155270560), ('Y', 59373566)]


def synth_contig(length, seed, n_frac=0.0, n_run_min=10000, n_runs=0):
  """Random uppercase ACGT contig as a uint8 array of ASCII codes.

  :param n_frac: fraction of the contig covered by 'N' runs (GRCh37 chr1 is ~10 %)
  :param n_runs: number of N runs (each >= n_run_min); the first and last sit at the contig ends
  """
  rng = np.random.RandomState(seed)
  seq = BASES[rng.randint(0, 4, size=length, dtype=np.uint8)]
  if n_frac > 0 and n_runs > 0:
    total_n = int(length * n_frac)
    w = rng.dirichlet(np.ones(n_runs)) * max(0, total_n - n_runs * n_run_min)
    run_len = (w.astype(np.int64) + n_run_min)
    # telomere-like runs at both ends, the rest at sorted random interior positions
    starts = np.sort(rng.randint(0, length - int(run_len.max()) - 1, size=n_runs))
    starts[0] = 0
    starts[-1] = length - run_len[-1]
    for s, l in zip(starts, run_len):
      seq[s:s + l] = ord('N')
  return seq


class VariantTable(object):
  """Flat variant arrays for one contig.

  pos      int64[n]   1-based VCF POS
  ref_off  int64[n+1] offsets into ref_pool  (REF allele bytes)
  alt_off  int64[n+1] offsets into alt_pool  (single ALT allele bytes)
  gt       int8[n, ploidy]  0 = reference allele, 1 = the ALT allele
  """
  __slots__ = ('chrom', 'pos', 'ref_pool', 'ref_off', 'alt_pool', 'alt_off', 'gt')

  def __init__(self, chrom, pos, ref_pool, ref_off, alt_pool, alt_off, gt):
    self.chrom, self.pos = chrom, pos
    self.ref_pool, self.ref_off = ref_pool, ref_off
    self.alt_pool, self.alt_off = alt_pool, alt_off
    self.gt = gt

  def __len__(self):
    return int(self.pos.shape[0])

  def ref(self, i):
    return self.ref_pool[self.ref_off[i]:self.ref_off[i + 1]].tobytes().decode()

  def alt(self, i):
    return self.alt_pool[self.alt_off[i]:self.alt_off[i + 1]].tobytes().decode()


def synth_variants(chrom, seq, start, end, seed, per_mb=740.0, ploidy=2, long_ins=3,
                   min_gap=40, end_margin=1000, p_snp=0.85, p_ins=0.075):
  """Synthetic SNP/indel call set over seq[start:end] (0-based half-open, BED style).

  Variants whose REF allele touches an 'N' are dropped (callers do not call in gaps).
  """
  rng = np.random.RandomState(seed)
  span = (end - start) - 2 * end_margin
  if span <= 0:
    return _empty_table(chrom, ploidy)
  n_est = int(span * per_mb / 1e6 * 1.3) + 16
  mean_gap = 1e6 / per_mb
  kind = rng.choice(3, size=n_est, p=[p_snp, p_ins, 1.0 - p_snp - p_ins])  # 0 SNP, 1 INS, 2 DEL
  ilen = rng.geometric(0.3, size=n_est).astype(np.int64)
  if long_ins > 0:
    ins_idx = np.flatnonzero(kind == 1)
    if ins_idx.size:
      pick = ins_idx[rng.randint(0, ins_idx.size, size=min(long_ins, ins_idx.size))]
      ilen[pick] = rng.randint(160, 400, size=pick.size)
  dlen = np.where(kind == 2, ilen, 0)
  gaps = min_gap + rng.geometric(1.0 / max(1.0, mean_gap - min_gap), size=n_est).astype(np.int64)
  # a deletion consumes reference; keep the next variant min_gap clear of its end
  gaps[1:] += dlen[:-1]
  pos0 = start + end_margin + np.cumsum(gaps)          # 0-based position of the anchor base
  keep = (pos0 + dlen + 1) < (end - end_margin)
  pos0, kind, ilen, dlen = pos0[keep], kind[keep], ilen[keep], dlen[keep]
  n = pos0.size

  ref_len = np.where(kind == 2, dlen + 1, 1)
  alt_len = np.where(kind == 1, ilen + 1, 1)
  ref_off = np.zeros(n + 1, dtype=np.int64); np.cumsum(ref_len, out=ref_off[1:])
  alt_off = np.zeros(n + 1, dtype=np.int64); np.cumsum(alt_len, out=alt_off[1:])

  # REF alleles are slices of the contig
  ref_idx = np.repeat(pos0 - ref_off[:-1], ref_len) + np.arange(ref_off[-1])
  ref_pool = seq[ref_idx]
  # ALT alleles: anchor base (or a different base for SNPs) followed by random inserted bases
  alt_pool = BASES[rng.randint(0, 4, size=alt_off[-1], dtype=np.uint8)]
  anchor = seq[pos0]
  snp = (kind == 0)
  code = np.searchsorted(BASES, anchor)  # BASES is sorted (A C G T); N -> 4 (dropped below anyway)
  snp_alt = BASES[(np.clip(code, 0, 3) + rng.randint(1, 4, size=n)) % 4]
  alt_pool[alt_off[:-1]] = np.where(snp, snp_alt, anchor)

  # genotypes: every record has at least one ALT copy
  gt = rng.randint(0, 2, size=(n, ploidy)).astype(np.int8)
  none = gt.sum(axis=1) == 0
  gt[none, rng.randint(0, ploidy, size=int(none.sum()))] = 1

  # drop records whose REF touches a non-ACGT base
  bad_base = ~np.isin(ref_pool, BASES)
  bad = np.add.reduceat(bad_base.astype(np.int64), ref_off[:-1]) > 0 if n else np.zeros(0, bool)
  return _subset(VariantTable(chrom, pos0 + 1, ref_pool, ref_off, alt_pool, alt_off, gt), ~bad)


def _empty_table(chrom, ploidy):
  z = np.zeros(1, dtype=np.int64)
  e = np.zeros(0, dtype=np.uint8)
  return VariantTable(chrom, np.zeros(0, dtype=np.int64), e, z, e.copy(), z.copy(),
                      np.zeros((0, ploidy), dtype=np.int8))


def _subset(vt, mask):
  idx = np.flatnonzero(mask)
  if idx.size == len(vt):
    return vt
  def pool(p, off):
    ln = (off[1:] - off[:-1])[idx]
    noff = np.zeros(idx.size + 1, dtype=np.int64); np.cumsum(ln, out=noff[1:])
    src = np.repeat(off[:-1][idx] - noff[:-1], ln) + np.arange(noff[-1])
    return p[src], noff
  rp, ro = pool(vt.ref_pool, vt.ref_off)
  ap, ao = pool(vt.alt_pool, vt.alt_off)
  return VariantTable(vt.chrom, vt.pos[idx], rp, ro, ap, ao, vt.gt[idx])


# ---------------------------------------------------------------------------------------------
# text writers for the file-based path

def write_fasta(path, contigs, width=60):
  """contigs: list of (name, uint8 array)."""
  with open(path, 'wb') as fp:
    for name, seq in contigs:
      fp.write(b'>' + name.encode() + b'\n')
      n = seq.shape[0]
      full = (n // width) * width
      if full:
        body = np.empty((n // width, width + 1), dtype=np.uint8)
        body[:, :width] = seq[:full].reshape(-1, width)
        body[:, width] = 10
        fp.write(body.tobytes())
      if n > full:
        fp.write(seq[full:].tobytes() + b'\n')


def write_vcf(path, tables, sample, contig_lengths=None):
  """tables: list of VariantTable (one per contig, in file order). '.gz' -> gzip text."""
  buf = io.StringIO()
  buf.write('##fileformat=VCFv4.1\n')
  for t in tables:
    ln = (contig_lengths or {}).get(t.chrom)
    buf.write('##contig=<ID={}{}>\n'.format(t.chrom, '' if ln is None else ',length={}'.format(ln)))
  buf.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
  buf.write('#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t{}\n'.format(sample))
  for t in tables:
    for i in range(len(t)):
      buf.write('{}\t{}\t.\t{}\t{}\t100\tPASS\t.\tGT\t{}\n'.format(
        t.chrom, int(t.pos[i]), t.ref(i), t.alt(i), '|'.join(str(int(g)) for g in t.gt[i])))
  data = buf.getvalue().encode()
  if str(path).endswith('.gz'):
    with gzip.open(path, 'wb') as fp:
      fp.write(data)
  else:
    with open(path, 'wb') as fp:
      fp.write(data)


def write_bed(path, regions):
  with open(path, 'w') as fp:
    for chrom, s, e in regions:
      fp.write('{}\t{}\t{}\n'.format(chrom, s, e))


# ---------------------------------------------------------------------------------------------
# named workloads

def config1(contig_len=1000000, seed=7, per_mb=740.0, names=('1', '10'), ploidy=2, long_ins=3):
  """BASELINE.json configs[0]/[1]: two synthetic contigs + GIAB-density SNP/indel VCF,
  whole-contig BED, sample INTEGRATION.  Returns dict(contigs, tables, regions, sample)."""
  contigs, tables, regions = [], [], []
  for k, name in enumerate(names):
    seq = synth_contig(contig_len, seed=seed * 1000 + k)
    contigs.append((name, seq))
    tables.append(synth_variants(name, seq, 0, contig_len, seed=seed * 1000 + 500 + k,
                                 per_mb=per_mb, ploidy=ploidy, long_ins=long_ins))
    regions.append((name, 0, contig_len))
  return {'contigs': contigs, 'tables': tables, 'regions': regions, 'sample': 'INTEGRATION'}


def config5(contig_len=200000, seed=9, viral_len=9000):
  """BASELINE.json configs[4]: tumour / normal mix with a viral spike-in, as three workloads to be
  run as three generate-reads invocations and concatenated -- how the reference makes mixes (the
  sample name in the qname tells them apart, Readme.md:14-16).  All three share the reference
  contigs ('1', an X-like 'X', 'virus'): normal = diploid autosome + haploid X; tumour = triploid
  GT on both; virus = one haploid dummy variant on the viral contig.  Returns {'normal', 'tumor',
  'virus'} -> workload dicts (each with its own regions / sample), plus the shared contigs."""
  seqs = [('1', synth_contig(contig_len, seed=seed * 1000)), ('X', synth_contig(contig_len // 2, seed=seed * 1000 + 1)),
          ('virus', synth_contig(viral_len, seed=seed * 1000 + 2))]
  by = dict(seqs)

  def tables(ploidy_1, ploidy_x, k):
    return [synth_variants('1', by['1'], 0, contig_len, seed=seed * 1000 + 100 * k, per_mb=900.0, ploidy=ploidy_1, long_ins=1),
            synth_variants('X', by['X'], 0, contig_len // 2, seed=seed * 1000 + 100 * k + 1, per_mb=900.0, ploidy=ploidy_x, long_ins=1)]

  human = [('1', 0, contig_len), ('X', 0, contig_len // 2)]
  p = viral_len // 2
  ref = chr(by['virus'][p - 1])
  dummy = table_from_records('virus', [(p, ref, {'A': 'C', 'C': 'G', 'G': 'T', 'T': 'A'}[ref], (1,))], 1)
  return {'normal': {'contigs': seqs, 'tables': tables(2, 1, 1), 'regions': human, 'sample': 'NORMAL'},
          'tumor': {'contigs': seqs, 'tables': tables(3, 3, 2), 'regions': human, 'sample': 'TUMOR'},
          'virus': {'contigs': seqs, 'tables': [dummy], 'regions': [('virus', 0, viral_len)], 'sample': 'VIRUS'}}


def write_workload(wl, prefix, gz=False):
  """Write FASTA / VCF / BED for a workload dict; returns (fasta, vcf, bed) paths."""
  fasta, vcf, bed = prefix + '.fasta', prefix + ('.vcf.gz' if gz else '.vcf'), prefix + '.bed'
  write_fasta(fasta, wl['contigs'])
  write_vcf(vcf, wl['tables'], wl['sample'], {n: s.shape[0] for n, s in wl['contigs']})
  write_bed(bed, wl['regions'])
  return fasta, vcf, bed


def chr1_shaped(seed=7, length=249250621, per_mb=1330.0, n_frac=0.10, n_runs=39):
  """BASELINE.json configs[2]: one chr1-shaped contig, diploid, ~330 k variants, ~10 % N."""
  seq = synth_contig(length, seed=seed * 1000, n_frac=n_frac, n_runs=n_runs)
  vt = synth_variants('1', seq, 0, length, seed=seed * 1000 + 500, per_mb=per_mb, long_ins=20)
  return {'contigs': [('1', seq)], 'tables': [vt], 'regions': [('1', 0, length)],
          'sample': 'INTEGRATION'}


def table_from_records(chrom, records, ploidy):
  """records: [(pos, ref, alt, gt tuple)] sorted by pos -> VariantTable."""
  refs = [r[1].encode() for r in records]
  alts = [r[2].encode() for r in records]
  ref_off = np.zeros(len(records) + 1, dtype=np.int64); np.cumsum([len(x) for x in refs], out=ref_off[1:])
  alt_off = np.zeros(len(records) + 1, dtype=np.int64); np.cumsum([len(x) for x in alts], out=alt_off[1:])
  gt = np.array([r[3] for r in records], dtype=np.int8).reshape(len(records), ploidy)
  return VariantTable(chrom, np.array([r[0] for r in records], dtype=np.int64),
                      np.frombuffer(b''.join(refs), dtype=np.uint8), ref_off,
                      np.frombuffer(b''.join(alts), dtype=np.uint8), alt_off, gt)


def merge_tables(a, b):
  """Merge two VariantTables of one contig, sorted by pos (stable: a's records first on ties)."""
  recs = [(int(t.pos[i]), t.ref(i), t.alt(i), tuple(int(g) for g in t.gt[i]), k)
          for k, t in enumerate((a, b)) for i in range(len(t))]
  recs.sort(key=lambda r: (r[0], r[4]))
  return table_from_records(a.chrom, [r[:4] for r in recs], a.gt.shape[1])


def edge_workload(seed=11):
  """Small workload that exercises the reference's corner cases:

  * BED regions that do not start at 0, a deletion that starts before a region and spans into it
    (returned by the overlap fetch, then skipped by ``v.pos < ref_pos``, rpc.py:55),
  * a SNP on the first base of a region, variants overlapping an accepted deletion (greedy skip),
    SNP+INS and INS+SNP at the same POS,
  * an N run and N triplets (the ``> 2 N`` template drop, readgenerate.py:204),
  * long insertions (the '>p:nI' CIGAR), dense variants (multi-node reads),
  * triploid, haploid and variant-free (assumed diploid, vcfio.py:74-76) regions,
  * a soft-masked (lower-case) run, IUPAC codes and lower-case 'n' in the variant-free contig.
  """
  rng = np.random.RandomState(seed)

  def seq_of(n, k):
    return synth_contig(n, seed=seed * 100 + k)

  e = seq_of(30000, 0)
  e[5000:5600] = ord('N')
  e[7000:7003] = ord('N')
  e[8000] = ord('N')
  e[20000:20002] = ord('N')

  def s(a, b):
    return e[a:b].tobytes().decode()

  def other(c):
    return {'A': 'C', 'C': 'G', 'G': 'T', 'T': 'A', 'N': 'A'}[c]

  manual = [
    (995, s(994, 1005), s(994, 995), (1, 1)),          # deletion spanning the start of region 1
    (1001, s(1000, 1001), other(s(1000, 1001)), (0, 1)),  # SNP on the first base of the region
    (2000, s(1999, 2005), s(1999, 2000), (1, 1)),      # DEL len 5 ...
    (2003, s(2002, 2003), other(s(2002, 2003)), (1, 1)),  # ... SNP inside it (skipped)
    (2006, s(2005, 2006), other(s(2005, 2006)), (1, 0)),  # first base after the deletion (kept)
    (3000, s(2999, 3000), other(s(2999, 3000)), (1, 1)),  # SNP then INS at the same POS
    (3000, s(2999, 3000), s(2999, 3000) + 'GATTACA', (1, 1)),
    (3100, s(3099, 3100), s(3099, 3100) + 'TT', (1, 1)),  # INS then SNP at the same POS
    (3100, s(3099, 3100), other(s(3099, 3100)), (1, 1)),
    (16000, s(15999, 16000), s(15999, 16000) + ''.join('ACGT'[i] for i in rng.randint(0, 4, size=333)), (0, 1)),
  ]
  rnd = synth_variants('e', e, 0, 30000, seed=seed * 100 + 50, per_mb=6000.0, ploidy=2, long_ins=2,
                       min_gap=12, end_margin=40)
  # keep random variants clear of the manual ones and of the region ends (a deletion crossing a
  # region end makes the reference's node list end in 'D', readgenerate.py:192)
  man_pos = np.array([m[0] for m in manual] + [14000, 29000])
  far = np.array([np.abs(man_pos - p).min() > 400 for p in rnd.pos], dtype=bool) if len(rnd) else np.zeros(0, bool)
  te = merge_tables(table_from_records('e', manual, 2), _subset(rnd, far))

  f = seq_of(20000, 1)
  tf = synth_variants('f', f, 0, 20000, seed=seed * 100 + 51, per_mb=3000.0, ploidy=3, long_ins=1, min_gap=15, end_margin=100)
  g = seq_of(8000, 2)
  tg = synth_variants('g', g, 0, 8000, seed=seed * 100 + 52, per_mb=3000.0, ploidy=1, long_ins=0, min_gap=15, end_margin=100)
  h = seq_of(5000, 3)
  # soft-masked and IUPAC bases: copied verbatim, not complemented on the reverse strand
  # (rpc.py:21 translate table covers ATCGN only), lower-case 'n' not counted by seq.count('N')
  h[1000:1400] += 32
  h[2000:2003] = np.frombuffer(b'RYM', np.uint8)
  h[3000] = ord('K')
  h[4000:4004] = ord('n')
  th = _empty_table('h', 2)
  return {'contigs': [('e', e), ('f', f), ('g', g), ('h', h)], 'tables': [te, tf, tg, th],
          'regions': [('e', 1000, 14000), ('e', 15000, 29000), ('f', 0, 20000), ('g', 0, 8000), ('h', 0, 5000)],
          'sample': 'EDGE'}


def softmask_workload(seed=21, length=40000):
  """A soft-masked reference: about half of the contig in lower-case stretches (repeat-masker
  style, mean length ~300), with an N run and IUPAC codes inside and outside of them.  The
  reference copies lower-case bases verbatim, does not complement them on the reverse strand
  (rpc.py:21 maps ATCGN only) and turns them into N on a miscall (illumina.py:160).  The VCF is
  written in upper case, as callers do."""
  rs = np.random.RandomState(seed)
  seq = synth_contig(length, seed=seed * 100)
  upper = seq.copy()
  pos = 0
  low = False
  while pos < length:
    n = int(rs.geometric(1.0 / 300.0))
    if low:
      seq[pos:pos + n] |= 0x20
    pos += n
    low = not low
  for a, b, ch in ((5000, 5400, 'N'), (9000, 9003, 'n'), (12000, 12001, 'R'), (12500, 12502, 'y'), (20000, 20002, 'N')):
    seq[a:b] = ord(ch)
    upper[a:b] = ord(ch.upper())
  vt = synth_variants('s', upper, 0, length, seed=seed * 100 + 1, per_mb=3000.0, ploidy=2, long_ins=2, min_gap=15, end_margin=100)
  return {'contigs': [('s', seq)], 'tables': [vt], 'regions': [('s', 0, length)], 'sample': 'SOFT'}


def grch37_shaped(scale=1.0, seed=7, per_mb=1330.0, n_frac=0.08, only=None):
  """BASELINE.json configs[3]: 24 contigs with GRCh37 primary-assembly lengths (times ``scale``),
  ~4 M variants at scale 1, autosomes diploid, X / Y haploid GT (docs/preparing_vcfs.rst:10-23),
  N runs at the contig ends and inside (telomere / centromere-like gaps)."""
  contigs, tables, regions = [], [], []
  for k, (name, length) in enumerate(GRCH37_CONTIGS):
    if only is not None and k not in only:   # a rank of a sharded run synthesises its own contigs only
      continue
    n = max(20000, int(length * scale))
    runs = max(2, int(round(n / 6.0e6)))
    seq = synth_contig(n, seed=seed * 1000 + k, n_frac=n_frac, n_run_min=min(10000, max(200, n // 200)), n_runs=runs)
    ploidy = 1 if name in ('X', 'Y') else 2
    vt = synth_variants(name, seq, 0, n, seed=seed * 1000 + 500 + k, per_mb=per_mb, ploidy=ploidy,
                        long_ins=max(1, int(20 * n / 249250621)), end_margin=min(1000, n // 20))
    contigs.append((name, seq)); tables.append(vt); regions.append((name, 0, n))
  return {'contigs': contigs, 'tables': tables, 'regions': regions, 'sample': 'INTEGRATION'}
