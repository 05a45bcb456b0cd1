"""check-reads / k_roundtrip_check: the god-aligner contract verified on the device for every read."""
import os

import numpy as np
import pytest

from mitty_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _checker(eng, regs):
  from mitty_b200.engine import Checker
  chk = Checker(eng)
  handles = []
  for r in regs:
    rid = eng.load_region(r['ref'], r['region'][1])
    for cpy, vl in enumerate(r['v']):
      cp = eng.build_copy(rid, vl)
      chk.add_copy(cp, r['region'][0], cpy)
      handles.append(cp)
  return chk


@pytest.fixture(scope='module')
def eng():
  from mitty_b200.engine import Engine
  e = Engine(0)
  yield e
  e.close()


@pytest.mark.parametrize('name', ['edge', 'softmask'])
def test_reference_output_passes(eng, name):
  """The reference's own FASTQ (golden files, produced by the unmodified reference) satisfies the
  contract read by read: insertions (incl. reads inside a long one: '>p:nI'), deletions, SNPs, N runs,
  lower-case stretches, both strands."""
  if name == 'edge':
    wl, f1, f2 = synth.edge_workload(), H.golden_fastq('edge.r1.fq.gz'), H.golden_fastq('edge.r2.fq.gz')
  else:
    import mitty_b200.simulation.illumina as il  # noqa: F401
    from tests.test_gpu_parity import gpu_generate
    wl = synth.softmask_workload()
    info = H.golden()['fastq']['softmask']
    f1, f2, _ = gpu_generate(eng, wl, H.model(info['model']), info['coverage'], info['seed'], 'deterministic')
    assert H.sha256(f1) == info['r1']['sha256']            # == the reference's bytes
  chk = _checker(eng, H.workload_regions(wl))
  n, bad, rep, c1, c2 = chk.check(f1, f2)
  assert n == f1.count(b'\n') // 4 and c1 == len(f1) and c2 == len(f2)
  assert bad == 0, rep
  if name == 'edge':
    assert b'>' in f1.split(b'\n')[0] or any(b'|>' in l for l in f1.split(b'\n')[0::4])     # the special CIGAR occurs in this workload
  chk.close()


def test_tampered_reads_are_reported(eng):
  """One base changed inside an '=' segment, a shifted POS, a swapped strand, a truncated CIGAR: each is
  reported with the record it sits in, the untouched reads still pass."""
  wl = synth.edge_workload()
  f1, f2 = bytearray(H.golden_fastq('edge.r1.fq.gz')), bytearray(H.golden_fastq('edge.r2.fq.gz'))
  lines = bytes(f1).split(b'\n')
  chk = _checker(eng, H.workload_regions(wl))
  # record 3 of file 1: flip its 10th base
  off = sum(len(l) + 1 for l in lines[:4 * 3 + 1])
  f1[off + 9] = ord('A') if f1[off + 9] != ord('A') else ord('C')
  # record 7: POS of the first read + 1 (both files share the qname; file 1's read is the first)
  rec7 = sum(len(l) + 1 for l in lines[:4 * 7])
  q = lines[4 * 7].split(b'|')
  q[4] = str(int(q[4]) + 1).encode()
  newq = b'|'.join(q)
  tampered = bytes(f1[:rec7]) + newq + bytes(f1[rec7 + len(lines[4 * 7]):])
  n, bad, rep, _, _ = chk.check(tampered, bytes(f2), max_report=10)
  assert bad >= 2 and {r[1] for r in rep if r[0] == 0} >= {3, 7}, rep
  assert all(r[0] == 0 for r in rep), rep                                         # file 2 is untouched
  # strand flipped in the qname of record 11 (second read): file 2 fails there
  lines2 = bytes(f2).split(b'\n')
  q = lines2[4 * 11].split(b'|')
  q[8] = b'1' if q[8] == b'0' else b'0'
  rec11 = sum(len(l) + 1 for l in lines2[:4 * 11])
  t2 = bytes(f2[:rec11]) + b'|'.join(q) + bytes(f2[rec11 + len(lines2[4 * 11]):])
  n, bad, rep, _, _ = chk.check(H.golden_fastq('edge.r1.fq.gz'), t2, max_report=10)
  assert bad == 1 and rep[0][:2] == (1, 11), rep
  chk.close()


@pytest.mark.timeout(900)
def test_every_read_of_the_bench_unit(eng):
  """The full-size unit the benchmark is quoted on (chr1-shaped, copy 1, Philox, perfect reads): ALL
  11.3 M reads re-derive from their qnames -- not a sample."""
  import mitty_b200.simulation.illumina as il
  from mitty_b200.engine import MODE_PHILOX, Checker
  wl = synth.chr1_shaped(seed=7, length=249250621, n_runs=39)
  m = H.model('hiseq-X-v2.5-Garvan.pkl')
  rm = il.read_model_params(m, 30.0)
  eng.load_model(rm)
  r = H.workload_regions(wl)[0]
  rid = eng.load_region(r['ref'], 0)
  cp = eng.build_copy(rid, r['v'][1])
  chk = Checker(eng)
  chk.add_copy(cp, '1', 1)
  n = int((cp.p_max - cp.p_min) * rm['p'] * 1.2)
  f1, f2, cnt, _, nb = eng.generate_unit(cp, n, rm['p'], MODE_PHILOX, 4242, '@S:0:2:', '|1|1')
  total = bad = 0
  step = 1 << 30
  o1 = o2 = 0
  while o1 < f1.size:
    got, nbad, rep, c1, c2 = chk.check(f1[o1:o1 + step], f2[o2:o2 + step])
    assert got > 0
    total += got; bad += nbad
    assert nbad == 0, rep[:3]
    o1 += c1; o2 += c2
  assert total == cnt and o1 == f1.size and o2 == f2.size
  chk.close()
  eng.free_copy(cp); eng.free_region(rid)


def test_check_reads_cli(tmp_path):
  from click.testing import CliRunner
  from mitty_b200.cli import cli
  wl = synth.edge_workload()
  fa, vcf, bed = synth.write_workload(wl, str(tmp_path / 'edge'))
  r1, r2 = str(tmp_path / 'r1.fq'), str(tmp_path / 'r2.fq')
  open(r1, 'wb').write(H.golden_fastq('edge.r1.fq.gz')); open(r2, 'wb').write(H.golden_fastq('edge.r2.fq.gz'))
  res = CliRunner().invoke(cli, ['check-reads', fa, vcf, wl['sample'], bed, r1, '--fastq2', r2], catch_exceptions=False)
  assert res.exit_code == 0 and ': 0 failed' in res.output, res.output
  # corrupted reads do NOT satisfy the contract (substitutions): the command says so and exits 1
  c1, c2 = str(tmp_path / 'c1.fq'), str(tmp_path / 'c2.fq')
  open(c1, 'wb').write(H.golden_fastq('edge.c1.fq.gz')); open(c2, 'wb').write(H.golden_fastq('edge.c2.fq.gz'))
  res = CliRunner().invoke(cli, ['check-reads', fa, vcf, wl['sample'], bed, c1, '--fastq2', c2, '--max-report', '3'])
  assert res.exit_code == 1 and 'failed' in res.output and res.output.count('record') == 3, res.output
