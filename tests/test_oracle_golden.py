"""Pin the CPU oracle (oracle/mitty_oracle.c) before anything trusts it.

1. Every known-answer vector of the reference's own unit tests
   (mitty/test/simulation/test_rpc.py, mitty/test/lib/test_vcfio.py), restated here with their
   file:line, evaluated through the oracle and through the host VCF front half.
2. Golden vectors produced by running the unmodified reference in the build container
   (tests/golden/make_golden.py): seed schedules, template sampling, node lists, reads, and whole
   FASTQ pairs from generate-reads / corrupt-reads with --threads 1.
"""
import os

import numpy as np
import pytest

import oracle
from mitty_b200 import synth
from mitty_b200.lib import vcfio
from tests import helpers as H


@pytest.fixture()
def tiny(tmp_path):
  p = H.write_tiny(tmp_path, gz=True)
  df = vcfio.load_variant_file(p['vcf'], 'g0_s0', p['whole_bed'])
  return p, df


# ---- test_vcfio.py -------------------------------------------------------------------------------

def test_vcfio_basic(tmp_path):
  """test_vcfio.py:9-18 -- region-restricted fetch keeps the deletion that starts inside [8,14)."""
  p = H.write_tiny(tmp_path, gz=True)
  v = vcfio.load_variant_file(p['vcf'], 'g0_s0', p['bed_8_14'])
  assert v[0]['v'][1][0].tuple() == (11, 'CAA', 'C', 'D', 2), v[0]['v']
  assert v[0]['v'][0][0].tuple() == (14, 'G', 'T', 'X', 0), v[0]['v']
  assert len(v[0]['v'][0]) == 1


def test_vcfio_complex_variant_error(tmp_path):
  """test_vcfio.py:21-27"""
  p = H.write_tiny(tmp_path)
  with pytest.raises(ValueError):
    vcfio.load_variant_file(p['flawed'], 'g0_s0', p['whole_bed'])


def test_vcfio_expansion(tiny):
  """test_vcfio.py:41-62"""
  _, vcf = tiny
  v = vcf[0]['v'][1]
  assert (v[0].cigarop, v[0].oplen) == ('X', 0)
  assert (v[1].cigarop, v[1].oplen) == ('I', 3)
  assert (v[2].cigarop, v[2].oplen) == ('D', 2)


def test_vcfio_edge_golden(tmp_path):
  """Host VCF front half == reference vio.load_variant_file on the edge workload
  (triploid / haploid / empty regions, records overlapping region starts)."""
  wl = synth.edge_workload()
  fa, vcf, bed = synth.write_workload(wl, str(tmp_path / 'edge'))
  got = vcfio.load_variant_file(vcf, wl['sample'], bed)
  want = H.golden()['edge_variants']
  assert len(got) == len(want)
  for g, w in zip(got, want):
    assert list(g['region']) == w['region']
    assert [[list(v.tuple()) for v in cp] for cp in g['v']] == w['v']
  # the in-memory path used by the large workloads gives the same arrays
  mem = H.workload_regions(wl)
  for g, m in zip(got, mem):
    for a, b in zip(g['v'], m['v']):
      assert [v.tuple() for v in a] == [v.tuple() for v in b]


# ---- test_rpc.py ---------------------------------------------------------------------------------

def test_rpc_expand_sequence(tiny):
  """test_rpc.py:111-126"""
  _, vcf = tiny
  nodes = oracle.create_node_list(H.TINY_SEQ, 1, H.oracle_cv(vcf[0]['v'][1]))
  assert nodes == [
    (1, 1, '=', 4, 'ATGA', None), (5, 5, 'X', 1, 'T', 0), (6, 6, '=', 3, 'GTA', None),
    (9, 9, 'I', 3, 'TTT', 3), (12, 9, '=', 3, 'TCC', None), (14, 14, 'D', 2, '', -2),
    (15, 14, '=', 7, 'GGAGGCG', None), (21, 25, 'D', 4, '', -4), (22, 25, '=', 1, 'C', None)]


def test_rpc_single_variant_expansions(tiny):
  """test_rpc.py:24-108: snp/insertion/deletion with and without a leading M node, restated as
  create_node_list calls on sub-problems (same cursor arithmetic)."""
  _, vcf = tiny
  v = vcf[0]['v'][1]
  one = lambda i: oracle.CopyVariants([v[i].pos], [v[i].cigarop], [v[i].oplen], [v[i].alt])
  # SNP, M section (test_rpc.py:39-51)
  n = oracle.create_node_list(H.TINY_SEQ, 1, one(0))
  assert n[0] == (1, 1, '=', 4, 'ATGA', None) and n[1] == (5, 5, 'X', 1, 'T', 0)
  # SNP, no M section (test_rpc.py:24-36): region starts on the SNP
  n = oracle.create_node_list(H.TINY_SEQ[4:], 5, one(0))
  assert n[0] == (5, 5, 'X', 1, 'T', 0)
  # INS, M section (test_rpc.py:69-81) from ref_pos 6
  n = oracle.create_node_list(H.TINY_SEQ[5:], 6, one(1))
  assert n[0] == (6, 6, '=', 3, 'GTA', None) and n[1] == (9, 9, 'I', 3, 'TTT', 3)
  # DEL, M section (test_rpc.py:99-108 uses samp_pos 12 / ref_pos 9; here both cursors start at 9)
  n = oracle.create_node_list(H.TINY_SEQ[8:], 9, one(2))
  assert n[0] == (9, 9, '=', 3, 'TCC', None) and n[1] == (11, 14, 'D', 2, '', -2)


def test_rpc_begin_end_nodes(tiny):
  """test_rpc.py:129-138"""
  _, vcf = tiny
  cv = H.oracle_cv(vcf[0]['v'][1])
  n0 = [oracle.generate_read(H.TINY_SEQ, 1, cv, p, 10)[4] for p in range(1, 16)]
  n1 = [oracle.generate_read(H.TINY_SEQ, 1, cv, p, 10)[5] for p in range(1, 16)]
  assert n0 == [0, 0, 0, 0, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 6]
  assert n1 == [3, 3, 4, 4, 4, 6, 6, 6, 6, 6, 6, 6, 8, 8, 8]


RPC_KAT_CPY1 = [  # test_rpc.py:146-160
  (1, 10, 0, 3, (1, '4=1X3=2I', [0, 3], 'ATGATGTATT')),
  (2, 10, 0, 3, (2, '3=1X3=3I', [0, 3], 'TGATGTATTT')),
  (3, 10, 0, 4, (3, '2=1X3=3I1=', [0, 3], 'GATGTATTTT')),
  (4, 10, 0, 4, (4, '1=1X3=3I2=', [0, 3], 'ATGTATTTTC')),
  (5, 10, 1, 4, (5, '1X3=3I3=', [0, 3], 'TGTATTTTCC')),
  (6, 10, 2, 6, (6, '3=3I3=2D1=', [3, -2], 'GTATTTTCCG')),
  (7, 10, 2, 6, (7, '2=3I3=2D2=', [3, -2], 'TATTTTCCGG')),
  (8, 10, 2, 6, (8, '1=3I3=2D3=', [3, -2], 'ATTTTCCGGA')),
  (9, 10, 3, 6, (9, '3I3=2D4=', [3, -2], 'TTTTCCGGAG')),
  (10, 10, 3, 6, (9, '2I3=2D5=', [3, -2], 'TTTCCGGAGG')),
  (11, 10, 3, 6, (9, '1I3=2D6=', [3, -2], 'TTCCGGAGGC')),
  (12, 10, 4, 6, (9, '3=2D7=', [-2], 'TCCGGAGGCG')),
  (13, 10, 4, 8, (10, '2=2D7=4D1=', [-2, -4], 'CCGGAGGCGC')),
  (14, 9, 4, 8, (11, '1=2D7=4D1=', [-2, -4], 'CGGAGGCGC')),
  (15, 8, 6, 8, (14, '7=4D1=', [-4], 'GGAGGCGC')),
  (9, 2, 3, 3, (8, '>0:2I', [3], 'TT')),  # test_rpc.py:176-180, read from inside the insertion
]
RPC_KAT_CPY0 = [  # test_rpc.py:168-172
  (1, 10, 0, 0, (1, '10=', [], 'ATGACGTATC')),
  (2, 10, 0, 0, (2, '10=', [], 'TGACGTATCC')),
  (4, 10, 0, 0, (4, '10=', [], 'ACGTATCCAA')),
  (5, 10, 0, 1, (5, '9=1X', [0], 'CGTATCCAAT')),
  (6, 10, 0, 2, (6, '8=1X1=', [0], 'GTATCCAATG')),
]


def test_rpc_read_gen(tiny):
  """test_rpc.py:141-180: the 15 + 5 + 1 golden (pos, cigar, v_list, seq) tuples."""
  _, vcf = tiny
  for cpy, kat in ((1, RPC_KAT_CPY1), (0, RPC_KAT_CPY0)):
    cv = H.oracle_cv(vcf[0]['v'][cpy])
    for p, l, n0, n1, want in kat:
      assert oracle.generate_read(H.TINY_SEQ, 1, cv, p, l, n0, n1)[:4] == want
      # looked-up nodes agree with the hand-written ones
      assert oracle.generate_read(H.TINY_SEQ, 1, cv, p, l)[4:] == (n0, n1)


# ---- golden vectors from the reference itself ------------------------------------------------------

def test_schedule_and_seed_split():
  """a6 / a8 (readgenerate.py:129-159, illumina.py:56-58) incl. SURVEY.md 8(a) KATs."""
  g = H.golden()
  for key, shape, passes in (('schedule_seed7_2x2x2', [2, 2], 2), ('schedule_seed123456789_3-1-2x4', [3, 1, 2], 4)):
    seed = int(key.split('_')[1][4:])
    units = [(ri, c) for ri, pl in enumerate(shape) for c in range(pl) for _ in range(passes)]
    seeds, order = oracle.unit_schedule(seed, len(units))
    got = [[units[k][0], units[k][1], int(seeds[k])] for k in order]
    assert got == g[key]
  seeds, order = oracle.unit_schedule(7, 8)
  assert seeds.tolist() == [976413892, 3349725721, 1369975286, 1882953283, 4201435347, 3107259287, 1956722279, 4200432988]
  assert order.tolist() == [3, 4, 5, 2, 6, 0, 1, 7]
  for s, want in g['unit_seed_split'].items():
    assert oracle.unit_seeds(int(s)).tolist() == want
  assert oracle.unit_seeds(1882953283).tolist() == [1446531787, 3751157881, 3280394916, 1732964261]
  assert oracle.corrupt_worker_seeds(7, 2).tolist() == [327741615, 976413892]


def test_read_model_params():
  for key, want in H.golden()['read_model_params'].items():
    name, cov = key.split('@')
    got = oracle.read_model_params(H.model(name), float(cov))
    assert (got['p'], got['passes'], got['rlen']) == (want['p'], want['passes'], want['rlen'])


def test_templates_golden():
  """a9 / a10 == illumina.generate_reads of the reference (three (model, span, seed) cases)."""
  z = np.load(os.path.join(H.GOLDEN, 'templates.npz'))
  keys = sorted({k.rsplit('_', 1)[0] for k in z.files})
  assert len(keys) == 3
  for key in keys:
    name, p_min, p_max, seed = key.rsplit('_', 3)
    rm = oracle.read_model_params(H.model(name + '.pkl'), 30.0)
    ts, te, fo = oracle.templates(rm['p'], rm['rlen'], rm['cum_tlen'], int(p_min), int(p_max), int(seed))
    np.testing.assert_array_equal(ts, z[key + '_pos0'])
    np.testing.assert_array_equal(te - rm['rlen'], z[key + '_pos1'])
    np.testing.assert_array_equal(fo, z[key + '_fo0'])
    np.testing.assert_array_equal(1 - fo, z[key + '_fo1'])


def _md5_or_seq(s):
  import hashlib
  return hashlib.md5(s.encode()).hexdigest()[:8] if len(s) > 40 else s


def test_edge_nodes_and_reads():
  g = H.golden()
  regs = H.workload_regions(synth.edge_workload())
  for ri, r in enumerate(regs):
    for cpy, vl in enumerate(r['v']):
      nodes = oracle.create_node_list(r['ref'], r['region'][1] + 1, H.oracle_cv(vl))
      got = [[n[0], n[1], n[2], n[3], _md5_or_seq(n[4]), n[5]] for n in nodes]
      assert got == g['edge_nodes']['{}_{}'.format(ri, cpy)]
  r = regs[0]
  cv = H.oracle_cv(r['v'][1])
  for key, rows in g['edge_reads'].items():
    L = int(key.rsplit('L', 1)[1])
    for p, pos, cigar, v_list, seq, n0, n1 in rows:
      assert oracle.generate_read(r['ref'], r['region'][1] + 1, cv, p, L) == (pos, cigar, v_list, seq, n0, n1)


@pytest.mark.parametrize('name,wl_fn', [('edge', synth.edge_workload),
                                        ('edge250', synth.edge_workload),
                                        ('mid', lambda: synth.config1(contig_len=100000)),
                                        ('softmask', synth.softmask_workload)])
def test_fastq_golden(name, wl_fn):
  """generate-reads + corrupt-reads (--threads 1) of the oracle == the reference, byte for byte."""
  info = H.golden()['fastq'][name]
  m = H.model(info['model'])
  regs = H.oracle_regions(H.workload_regions(wl_fn()))
  wl = wl_fn()
  f1, f2, n = oracle.generate_reads_cmd(regs, m, info['coverage'], info['seed'], wl['sample'])
  assert n == info['pairs']
  assert (len(f1), H.sha256(f1)) == (info['r1']['bytes'], info['r1']['sha256'])
  assert (len(f2), H.sha256(f2)) == (info['r2']['bytes'], info['r2']['sha256'])
  c1, c2, n = oracle.corrupt_reads_cmd(m, info['seed'], f1, f2)
  assert n == info['pairs']
  assert H.sha256(c1) == info['c1']['sha256']
  assert H.sha256(c2) == info['c2']['sha256']
  if name == 'edge':
    assert f1 == H.golden_fastq('edge.r1.fq.gz') and f2 == H.golden_fastq('edge.r2.fq.gz')
    assert c1 == H.golden_fastq('edge.c1.fq.gz') and c2 == H.golden_fastq('edge.c2.fq.gz')


def test_fastq_golden_config1():
  """Full config 1 (two 1 Mb contigs, 200 k pairs): generate-reads hash; corrupt hash when present."""
  info = H.golden()['fastq']['config1']
  m = H.model(info['model'])
  wl = synth.config1()
  f1, f2, n = oracle.generate_reads_cmd(H.oracle_regions(H.workload_regions(wl)), m, info['coverage'], info['seed'], wl['sample'])
  assert n == info['pairs']
  assert H.sha256(f1) == info['r1']['sha256'] and H.sha256(f2) == info['r2']['sha256']
  if 'c1' in info:
    c1, c2, _ = oracle.corrupt_reads_cmd(m, info['seed'], f1, f2)
    assert H.sha256(c1) == info['c1']['sha256'] and H.sha256(c2) == info['c2']['sha256']


def test_corrupt_template_golden():
  """illumina.corrupt_template on single templates incl. N / lowercase / IUPAC bases and L=1."""
  m = H.model('hiseq-X-v2.5-Garvan.pkl')
  for case in H.golden()['corrupt_template']:
    q, s1, s2 = case['in']
    fq = lambda s: '@{}\n{}\n+\n{}\n'.format(q, s, '~' * len(s)).encode()
    c1, c2, _ = oracle.corrupt_fastq(m['cum_bq_mat'], case['seed'], fq(s1), fq(s2))
    want = case['out']
    assert c1.decode() == '@{}\n{}\n+\n{}\n'.format(*want[0])
    assert c2.decode() == '@{}\n{}\n+\n{}\n'.format(*want[1])
