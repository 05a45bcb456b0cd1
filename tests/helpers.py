"""Shared test fixtures: the reference's tiny example, golden-vector access, workload plumbing."""
import gzip
import hashlib
import json
import os

import numpy as np

import oracle
from mitty_b200 import synth
from mitty_b200.lib import vcfio
from mitty_b200.readmodels import load_model

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')

# The reference's tiny fixture (mitty/test/data/tiny.fasta, tiny.vcf, tiny.whole.bed, tiny.8-14.bed),
# restated inline: 25 bp of sequence and five variants of sample g0_s0.
TINY_SEQ = 'ATGACGTATCCAAGGAGGCGTTACC'
TINY_VCF = (
  '##fileformat=VCFv4.1\n'
  '##contig=<ID=1,length=23>\n'
  '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
  '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tg0_s0\n'
  '1\t5\t.\tC\tT\t100\tPASS\t.\tGT\t0|1\n'
  '1\t8\t.\tA\tATTT\t100\tPASS\t.\tGT\t0|1\n'
  '1\t11\t.\tCAA\tC\t100\tPASS\t.\tGT\t0|1\n'
  '1\t14\t.\tG\tT\t100\tPASS\t.\tGT\t1|0\n'
  '1\t20\t.\tGTTAC\tG\t100\tPASS\t.\tGT\t1|1\n')
FLAWED_TINY_VCF = (
  '##fileformat=VCFv4.1\n'
  '##contig=<ID=tiny,length=51>\n'
  '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
  '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tg0_s0\n'
  '1\t5\t.\tCGT\tCTT\t100\tPASS\t.\tGT\t0|1\n'
  '1\t10\t.\tCCA\tCC\t100\tPASS\t.\tGT\t0|1\n'
  '1\t15\t.\tGA\tGAT\t100\tPASS\t.\tGT\t0|1\n')


def write_tiny(tmp_path, gz=False):
  """-> dict of paths for the tiny fixture files."""
  p = {}
  p['fasta'] = str(tmp_path / 'tiny.fasta')
  with open(p['fasta'], 'w') as fp:
    fp.write('>1\n' + TINY_SEQ + '\n')
  for name, text in (('vcf', TINY_VCF), ('flawed', FLAWED_TINY_VCF)):
    p[name] = str(tmp_path / ('{}.vcf{}'.format(name, '.gz' if gz else '')))
    if gz:
      with gzip.open(p[name], 'wt') as fp:
        fp.write(text)
    else:
      with open(p[name], 'w') as fp:
        fp.write(text)
  p['whole_bed'] = str(tmp_path / 'tiny.whole.bed')
  with open(p['whole_bed'], 'w') as fp:
    fp.write('1\t0\t23')
  p['bed_8_14'] = str(tmp_path / 'tiny.8-14.bed')
  with open(p['bed_8_14'], 'w') as fp:
    fp.write('1\t8\t14')
  return p


_golden = None


def golden():
  global _golden
  if _golden is None:
    with open(os.path.join(GOLDEN, 'golden.json')) as fp:
      _golden = json.load(fp)
  return _golden


def golden_fastq(name):
  with gzip.open(os.path.join(GOLDEN, name), 'rb') as fp:
    return fp.read()


def sha256(b):
  return hashlib.sha256(b).hexdigest()


def oracle_cv(vl):
  """VariantList (product host arrays) -> oracle.CopyVariants (same arrays, no arithmetic)."""
  return oracle.CopyVariants(vl.pos, vl.op, vl.oplen, (vl.alt_pool, vl.alt_off))


def workload_regions(wl):
  """In-memory workload -> region dicts with reference bytes and per-copy variant arrays."""
  contigs = dict(wl['contigs'])
  tables = {t.chrom: t for t in wl['tables']}
  out = []
  for region in wl['regions']:
    chrom, s, e = region
    r = vcfio.from_variant_table(tables[chrom], region)
    r['ref'] = np.ascontiguousarray(contigs[chrom][s:e])
    out.append(r)
  return out


def oracle_regions(regions):
  return [{'region': r['region'], 'ref': r['ref'], 'v': [oracle_cv(v) for v in r['v']]} for r in regions]


def model(name):
  return load_model(name)
