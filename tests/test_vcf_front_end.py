"""Host front end: the vectorised VCF parser against a line-by-line restatement
of vcfio.parse (mitty/lib/vcfio.py:105-126) on records the array path does not cover."""
import gzip

import numpy as np
import pytest

from mitty_b200 import synth
from mitty_b200.lib import vcfio

HDR = '##fileformat=VCFv4.2\n##contig=<ID=1>\n##contig=<ID=3>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS0\tS1\n'


def simple_parse(text, sample, contig, start, stop, cpy):
  """pysam-free restatement: overlap fetch + vcfio.parse for one copy -> [(pos, ref, alt, op, oplen)]."""
  col, out = None, []
  for line in text.replace('\r\n', '\n').split('\n'):
    if line.startswith('#CHROM'):
      col = line.split('\t').index(sample)
    if not line or line.startswith('#'):
      continue
    f = line.split('\t')
    pos, ref, alts = int(f[1]), f[3], f[4].split(',')
    if f[0] != contig or not (pos - 1 < stop and pos - 1 + len(ref) > start):
      continue
    gts = f[col].split(':')[f[8].split(':').index('GT')].replace('/', '|').split('|')
    g = gts[cpy]
    if g == '0':
      continue
    alt = ([ref] + alts)[int(g)]
    if len(ref) == 1 and len(alt) == 1:
      out.append((pos, ref, alt, 'X', 0))
    elif len(ref) == 1:
      out.append((pos, ref, alt, 'I', len(alt) - 1))
    elif len(alt) == 1:
      out.append((pos, ref, alt, 'D', len(ref) - 1))
    else:
      raise ValueError('complex')
  return out


BODY = (
  '1\t10\t.\tA\tC\t.\tPASS\tAC=1,2;X=a:b\tGT\t0|1\t1|1\n'
  '1\t20\t.\tA\tC,G\t.\tPASS\t.\tGT\t1|2\t0|1\n'              # multi-allelic
  '1\t30\t.\tAT\tA\t.\tPASS\t.\tGT:DP\t1/0:17\t0/0:3\n'       # unphased, extra FORMAT keys
  '1\t40\t.\tA\tATTT\t.\tPASS\tDP=3\tGT\t1|1\t1|0\n'
  '1\t50\t.\tC\tG\t.\tPASS\t.\tDP:GT\t9:0|1\t1:1|1\n'         # GT not first (not valid VCF, still parsed)
  '2\t5\t.\tG\tT\t.\tPASS\t.\tGT\t1\t1\n'                     # haploid contig
  '2\t9\t.\tGCA\tG\t.\tPASS\t.\tGT\t1\t0\n'
  '1\t60\t.\tT\tA\t.\tPASS\t.\tGT\t1|0\t0|1\n'                # contig 1 again after contig 2
)


@pytest.mark.parametrize('crlf', [False, True])
@pytest.mark.parametrize('gz', [False, True])
def test_parser_matches_line_by_line(tmp_path, crlf, gz):
  text = HDR + BODY
  if crlf:
    text = text.replace('\n', '\r\n')
  path = str(tmp_path / ('a.vcf.gz' if gz else 'a.vcf'))
  with (gzip.open(path, 'wb') if gz else open(path, 'wb')) as fp:
    fp.write(text.encode())
  bed = str(tmp_path / 'a.bed')
  regions = [('1', 0, 100), ('1', 25, 45), ('2', 0, 20), ('2', 10, 20), ('3', 0, 10)]
  open(bed, 'w').write(''.join('{}\t{}\t{}\n'.format(*r) for r in regions))
  for sample in ('S0', 'S1'):
    df = vcfio.load_variant_file(path, sample, bed)
    assert [d['region'] for d in df] == regions
    for d in df:
      chrom, start, stop = d['region']
      ploidy = 1 if chrom == '2' else 2          # sniffed from the first record; no records -> 2 (vcfio.py:74-79)
      assert len(d['v']) == ploidy
      for cpy, vl in enumerate(d['v']):
        assert [v.tuple() for v in vl] == simple_parse(text, sample, chrom, start, stop, cpy)


def test_parser_errors(tmp_path):
  def load(body, sample='S0'):
    path = str(tmp_path / 'e.vcf'); open(path, 'w').write(HDR + body)
    bed = str(tmp_path / 'e.bed'); open(bed, 'w').write('1\t0\t100\n')
    return vcfio.load_variant_file(path, sample, bed)
  with pytest.raises(ValueError):
    load('1\t10\t.\tAT\tGC\t.\tPASS\t.\tGT\t1|1\t0|0\n')                     # complex variant, vcfio.py:124
  with pytest.raises(ValueError):
    load('1\t10\t.\tAT\tGC,A\t.\tPASS\t.\tGT\t1|2\t0|0\n')                   # the same through the per-record path
  with pytest.raises(ValueError):
    load('1\t10\t.\tA\tC\t.\tPASS\t.\tGT\t0|1\t0|0\n', sample='nobody')
  with pytest.raises(IndexError):
    load('1\t10\t.\tA\tC\t.\tPASS\t.\tGT\t0|1\t0|0\n1\t20\t.\tA\tC\t.\tPASS\t.\tGT\t1\t0|0\n')   # ragged ploidy
  assert [len(v) for v in load('1\t10\t.\tAT\tGC\t.\tPASS\t.\tGT\t0|0\t1|1\n')[0]['v']] == [0, 0]   # GT 0 is skipped before the check


def test_parser_matches_synthetic_table(tmp_path):
  """Same VariantLists from the VCF text as from the in-memory table the benchmark uses."""
  from mitty_b200 import synth
  wl = synth.edge_workload()
  fa, vcf, bed = synth.write_workload(wl, str(tmp_path / 'edge'), gz=True)
  df = vcfio.load_variant_file(vcf, wl['sample'], bed)
  tables = {t.chrom: t for t in wl['tables']}
  for d in df:
    ref = vcfio.from_variant_table(tables[d['region'][0]], d['region'])
    assert len(ref['v']) == len(d['v'])
    for a, b in zip(ref['v'], d['v']):
      assert np.array_equal(a.pos, b.pos) and np.array_equal(a.op, b.op) and np.array_equal(a.oplen, b.oplen)
      assert np.array_equal(a.alt_pool, b.alt_pool) and np.array_equal(a.alt_off, b.alt_off)


def test_filter_variants(tmp_path):
  """filter-variants (vcfio.py:128-169): one sample, BED-restricted, complex calls dropped; the result
  feeds generate-reads' loader without the complex-variant error."""
  from click.testing import CliRunner
  from mitty_b200.cli import cli
  body = (
    '1\t10\t.\tA\tC\t.\tPASS\t.\tGT\t0|1\t1|1\n'
    '1\t20\t.\tAT\tGC\t.\tPASS\t.\tGT\t1|0\t0|0\n'          # complex for S0, not carried by S1
    '1\t30\t.\tAT\tA,GCC\t.\tPASS\t.\tGT\t1|1\t1|2\n'       # S1 carries the complex second ALT
    '1\t40\t.\tA\tATT\t.\tPASS\t.\tGT:DP\t1|1:5\t0|1:7\n'
    '1\t500\t.\tG\tT\t.\tPASS\t.\tGT\t1|1\t1|1\n'            # outside the BED
    '2\t7\t.\tGCA\tG\t.\tPASS\t.\tGT\t1\t1\n'
  )
  vin = str(tmp_path / 'in.vcf'); open(vin, 'w').write(HDR + body)
  bed = str(tmp_path / 'r.bed'); open(bed, 'w').write('1\t0\t100\n2\t0\t50\n')
  for sample, kept in (('S0', [10, 30, 40, 7]), ('S1', [10, 20, 40, 7])):
    vout = str(tmp_path / (sample + '.vcf.gz'))
    res = CliRunner().invoke(cli, ['filter-variants', vin, sample, bed, vout], catch_exceptions=False)
    assert res.exit_code == 0, res.output
    text = gzip.open(vout, 'rt').read()
    recs = [l.split('\t') for l in text.split('\n') if l and not l.startswith('#')]
    assert [int(r[1]) for r in recs] == kept
    assert all(len(r) == 10 for r in recs)
    assert [l for l in text.split('\n') if l.startswith('#CHROM')][0].split('\t')[9:] == [sample]
    assert text.startswith('##fileformat')
    df = vcfio.load_variant_file(vout, sample, bed)                       # no complex-variant error any more
    assert [d['region'][0] for d in df] == ['1', '2']
  with pytest.raises(ValueError):
    vcfio.load_variant_file(vin, 'S0', bed)


def test_filter_variants_matches_reference(tmp_path, capsysbinary):
  """The records filter-variants keeps == what the unmodified reference's prepare_variant_file keeps
  (`_complex_variant`, mitty/lib/vcfio.py:139-146) on the same VCF/BED, sample by sample
  (tests/golden/filter_variants.json, written by tests/golden/make_filter_golden.py)."""
  import json
  import os
  from tests import helpers as H
  g = json.load(open(os.path.join(H.GOLDEN, 'filter_variants.json')))
  vin = str(tmp_path / 'in.vcf'); open(vin, 'w').write(g['vcf'])
  bed = str(tmp_path / 'r.bed'); open(bed, 'w').write(g['bed'])
  for sample, want in sorted(g['kept'].items()):
    vout = str(tmp_path / (sample + '.vcf'))
    vcfio.prepare_variant_file(vin, sample, bed, vout)
    recs = [l.split('\t') for l in open(vout).read().split('\n') if l and not l.startswith('#')]
    assert [[r[0], int(r[1]), r[3], r[4], r[9]] for r in recs] == want
    # '-' writes the same text to stdout (examples/reads/run.sh:9 pipes it into bgzip)
    capsysbinary.readouterr()
    vcfio.prepare_variant_file(vin, sample, bed, '-')
    assert capsysbinary.readouterr().out == open(vout, 'rb').read()


def test_unknown_contig_is_an_error(tmp_path):
  """A BED contig that the VCF neither declares (##contig) nor holds is pysam's
  `ValueError: invalid contig` (vcfio.py:62), e.g. 'chr1' over a VCF that says '1'; a declared contig
  without records is an empty region (diploid by default, vcfio.py:74-76)."""
  hdr = HDR.replace('##contig=<ID=3>', '##contig=<length=500,ID=7,assembly=x>')
  vin = str(tmp_path / 'c.vcf'); open(vin, 'w').write(hdr + '1\t10\t.\tA\tC\t.\tPASS\t.\tGT\t0|1\t1|1\n')
  bed = str(tmp_path / 'c.bed')
  open(bed, 'w').write('1\t0\t100\n7\t0\t100\n')
  df = vcfio.load_variant_file(vin, 'S0', bed)
  assert [len(d['v']) for d in df] == [2, 2] and len(df[1]['v'][0]) == 0
  open(bed, 'w').write('chr1\t0\t100\n')
  with pytest.raises(ValueError, match='invalid contig'):
    vcfio.load_variant_file(vin, 'S0', bed)


def _fasta_cases(tmp_path):
  rs = np.random.RandomState(11)
  def seq(n):
    s = rs.choice(np.frombuffer(b'ACGTacgtNnRY', dtype=np.uint8), size=n)
    return s.tobytes()
  def wrap(s, w, eol=b'\n', last_eol=True):
    lines = [s[i:i + w] for i in range(0, len(s), w)]
    return eol.join(lines) + (eol if last_eol and lines else b'')
  contigs = [('chr1 some description', seq(100003), 60, b'\n', True),     # the usual layout
             ('2', seq(5000), 70, b'\r\n', True),                        # CRLF
             ('empty', b'', 60, b'\n', True),
             ('one_line', seq(777), 100000, b'\n', True),
             ('exact', seq(600), 60, b'\n', True),                        # the last line is full
             ('last', seq(12345), 50, b'\n', False)]                      # no line end at the end of the file
  ragged = seq(900)
  body = b''
  want = {}
  for name, s, w, eol, le in contigs[:-1]:
    body += b'>' + name.encode() + eol + wrap(s, w, eol, le)
    want[name.split()[0]] = s
  # a contig whose lines differ in width, and one with two short lines adding up to a full one (61 bytes)
  body += b'>ragged\n' + ragged[:60] + b'\n' + ragged[60:100] + b'\n' + ragged[100:400] + b'\n\n' + ragged[400:] + b'\n'
  want['ragged'] = ragged
  tricky = seq(240)
  body += b'>tricky\n' + tricky[:60] + b'\n' + tricky[60:90] + b'\n' + tricky[90:119] + b'\n' + tricky[119:179] + b'\n' + tricky[179:239] + b'\n' + tricky[239:] + b'\n'
  want['tricky'] = tricky
  name, s, w, eol, le = contigs[-1]
  body += b'>' + name.encode() + eol + wrap(s, w, eol, le)
  want[name] = s
  path = str(tmp_path / 'ref.fa')
  with open(path, 'wb') as fp:
    fp.write(body)
  return path, want


def test_native_fasta_reader_equals_the_bytes_level_one(tmp_path):
  """mg_fasta_* (mapped file, arithmetic addressing of uniform contigs) against the plain-Python reader and
  the known sequences: line widths, CRLF, empty / one-line contigs, ragged lines, no final line end, and
  numpy-slice semantics of start / end (pysam clamps the same way)."""
  path, want = _fasta_cases(tmp_path)
  nat, ref = vcfio.FastaFile(path), vcfio.FastaFile(path, native=False)
  assert nat._h is not None and ref._h is None
  assert set(nat.references) == set(want) == set(ref.references)
  rs = np.random.RandomState(3)
  for name, s in want.items():
    n = len(s)
    assert nat.fetch(reference=name).tobytes() == s == ref.fetch(reference=name).tobytes()
    spans = [(0, n), (0, 0), (n, n + 5), (n - 1, n), (5, 3), (None, 10), (10, None), (59, 61), (60, 120), (n - 7, n + 100)]
    spans += [tuple(sorted(rs.randint(0, n + 1, size=2))) for _ in range(30)] if n else []
    for a, b in spans:
      got = nat.fetch(reference=name, start=a, end=b).tobytes()
      assert got == s[slice(a, b)] == ref.fetch(reference=name, start=a, end=b).tobytes(), (name, a, b)
  with pytest.raises(KeyError):
    nat.fetch(reference='nope', start=0, end=1)
  # into a caller's buffer (the workers' page-locked landing area)
  buf = np.zeros(200000, dtype=np.uint8)
  for fa_ in (nat, ref):
    got = fa_.fetch(reference='chr1', start=17, end=90017, out=buf)
    assert got.base is buf or got.ctypes.data == buf.ctypes.data
    assert got.tobytes() == want['chr1'][17:90017]
  with pytest.raises(ValueError):
    nat.fetch(reference='chr1', start=0, end=1000, out=np.zeros(10, dtype=np.uint8))
  nat.close()


def test_native_fasta_reader_large_region_in_threads(tmp_path):
  """A fetch of more than 8 MB is split over threads: the pieces must join without a seam."""
  rs = np.random.RandomState(5)
  s = rs.choice(np.frombuffer(b'ACGT', dtype=np.uint8), size=20000003)
  path = str(tmp_path / 'big.fa')
  with open(path, 'wb') as fp:
    fp.write(b'>big\n')
    lines = np.full((s.size + 59) // 60 * 61, 10, dtype=np.uint8).reshape(-1, 61)
    pad = np.concatenate([s, np.zeros(lines.shape[0] * 60 - s.size, dtype=np.uint8)])
    lines[:, :60] = pad.reshape(-1, 60)
    flat = lines.reshape(-1)
    fp.write(flat[:s.size + s.size // 60].tobytes())          # ends with the last base (no final line end)
  fa = vcfio.FastaFile(path)
  assert fa._h is not None
  assert np.array_equal(fa.fetch(reference='big'), s)
  assert np.array_equal(fa.fetch(reference='big', start=1234567, end=19999999), s[1234567:19999999])


def test_vcf_parsed_in_pieces_equals_one_pass(tmp_path, monkeypatch):
  """Large call sets are cut at line starts and parsed by threads: same table as the single pass, also when a
  contig continues in the next piece and when exotic records (multi-allelic, FORMAT beyond GT) fall anywhere."""
  wl = synth.config1(contig_len=400000, names=('1', '2', '3'))
  fa, vcf, bed = synth.write_workload(wl, str(tmp_path / 'w'))
  one = vcfio.VcfTable(vcf, wl['sample'])
  monkeypatch.setattr(vcfio, 'VCF_PIECE_BYTES', 1500)
  monkeypatch.setattr(vcfio.os, 'cpu_count', lambda: 7)
  many = vcfio.VcfTable(vcf, wl['sample'])
  assert list(one.contigs) == list(many.contigs) and len(one.contigs) == 3
  for name in one.contigs:
    a, b = one.contigs[name], many.contigs[name]
    for f in ('pos', 'reflen', 'rs', 're', 'as_', 'ae', 'gt', 'ploidy', 'exotic', 'ls', 'le', 'f9e', 'ss', 'se'):
      assert np.array_equal(getattr(a, f), getattr(b, f)), (name, f)
    assert a.slow == b.slow
  r1 = vcfio.load_variant_file(vcf, wl['sample'], bed)
  monkeypatch.undo()
  r0 = vcfio.load_variant_file(vcf, wl['sample'], bed)
  for x, y in zip(r0, r1):
    assert x['region'] == y['region'] and len(x['v']) == len(y['v'])
    for vx, vy in zip(x['v'], y['v']):
      assert np.array_equal(vx.pos, vy.pos) and np.array_equal(vx.op, vy.op) and np.array_equal(vx.alt_pool, vy.alt_pool)


def test_windowed_fetch_equals_the_full_scan(tmp_path):
  """VcfTable.fetch finds the records overlapping a region (htslib rule: a long REF that starts before the region
  still overlaps it) with two binary searches when the contig is sorted; same indices as the plain scan, also for
  an unsorted contig (which takes the scan)."""
  rs = np.random.RandomState(8)
  for sorted_ in (True, False):
    pos = rs.randint(1, 50000, size=3000)
    if sorted_:
      pos = np.sort(pos)
    reflen = np.where(rs.rand(pos.size) < 0.1, rs.randint(2, 400, size=pos.size), 1)
    lines = ['##fileformat=VCFv4.2', '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS']
    for p, n in zip(pos.tolist(), reflen.tolist()):
      lines.append('c\t{}\t.\t{}\t{}\t.\tPASS\t.\tGT\t0|1'.format(p, 'A' * n, 'A' if n > 1 else 'C'))
    path = str(tmp_path / ('s%d.vcf' % sorted_))
    with open(path, 'w') as fp:
      fp.write('\n'.join(lines) + '\n')
    tb = vcfio.VcfTable(path, 'S')
    c = tb.contigs['c']
    assert np.array_equal(c.pos, pos)
    for _ in range(300):
      a = int(rs.randint(0, 51000)); b = a + int(rs.randint(0, 3000))
      _, idx = tb.fetch('c', a, b)
      want = np.flatnonzero((pos - 1 < b) & (pos - 1 + reflen > a))
      assert np.array_equal(idx, want), (sorted_, a, b)
    assert c.is_sorted == sorted_
