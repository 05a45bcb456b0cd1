"""The god-aligner style round-trip checker itself, validated on the reference's golden FASTQ
(edge workload: multi-node reads, reads starting inside insertions, '>p:nI' reads, deletions)."""
from mitty_b200 import synth
from tests import helpers as H


def test_roundtrip_checker_on_reference_output():
  regs = H.workload_regions(synth.edge_workload())
  idx = {}

  def index_for(chrom, cpy):
    raise AssertionError('per-region lookup needed')

  f1, f2 = H.golden_fastq('edge.r1.fq.gz'), H.golden_fastq('edge.r2.fq.gz')
  # the edge workload has two regions on contig 'e': pick the region by position
  from mitty_b200.simulation.readgenerate import parse_qname
  errs, n, special, multi = [], 0, 0, 0
  for which, buf in enumerate((f1, f2)):
    lines = buf.decode().split('\n')
    for k in range(0, len(lines) - 1, 4):
      info = parse_qname(lines[k][1:])[which]
      r = next(r for r in regs if r['region'][0] == info.chrom and r['region'][1] < info.pos <= r['region'][2] + 1)
      key = (r['region'], info.cpy)
      if key not in idx:
        idx[key] = H.HaplotypeIndex(r['ref'], r['region'][1] + 1, H.oracle_cv(r['v'][info.cpy]))
      e = idx[key].check(info, lines[k + 1])
      n += 1
      special += info.special_cigar is not None
      multi += len(info.v_list) > 1
      if e:
        errs.append((lines[k], e))
  assert not errs, errs[:3]
  assert n > 5000 and special > 0 and multi > 100

  # and it does notice a wrong read
  lines = f1.decode().split('\n')
  info = parse_qname(lines[0][1:])[0]
  r = next(r for r in regs if r['region'][0] == info.chrom and r['region'][1] < info.pos <= r['region'][2] + 1)
  bad = lines[1][:70] + ('A' if lines[1][70] != 'A' else 'C') + lines[1][71:]
  assert idx[(r['region'], info.cpy)].check(info, bad) is not None
