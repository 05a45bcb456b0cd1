"""Golden vectors for filter-variants: runs the UNMODIFIED reference's prepare_variant_file
(mitty/lib/vcfio.py:128-169) over a small VCF that exercises its `_complex_variant` rule, with
oracle/refshim/pysam.py standing in for pysam (I/O only; the rule is the reference's own code).
Build container only (needs /root/reference).  Writes tests/golden/filter_variants.json:
the input VCF text, the BED text and, per sample, the records the reference kept.

    python tests/golden/make_filter_golden.py
"""
import json
import os
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
warnings.filterwarnings('ignore')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'refshim'))
sys.path.insert(0, '/root/reference')

import mitty.lib.vcfio as vio   # noqa: E402  (the reference)

HDR = ('##fileformat=VCFv4.1\n##contig=<ID=1,length=1000>\n##contig=<ID=2,length=1000>\n##contig=<ID=X,length=1000>\n'
       '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Depth">\n'
       '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS0\tS1\tS2\n')
BODY = (
  '1\t10\t.\tA\tC\t.\tPASS\t.\tGT\t0|1\t1|1\t0|0\n'                 # SNP
  '1\t20\t.\tAT\tGC\t.\tPASS\t.\tGT\t1|0\t0|0\t0|1\n'               # MNP: complex for carriers only
  '1\t30\t.\tAT\tA,GCC\t.\tPASS\t.\tGT\t1|1\t1|2\t2|2\n'            # multi-allelic, second ALT complex
  '1\t40\t.\tA\tATT\t.\tPASS\t.\tGT:DP\t1|1:5\t0|1:7\t0|0:3\n'      # insertion, extra FORMAT key
  '1\t50\t.\tGCA\tG\t.\tPASS\t.\tGT\t0|1\t1|0\t1|1\n'               # deletion
  '1\t60\t.\tGCA\tGCA\t.\tPASS\t.\tGT\t1|1\t0|1\t0|0\n'             # ALT equal to REF: rlen > 1, len(alt) > 1, but ref == alt
  '1\t70\t.\tCAG\tC,CAGAG\t.\tPASS\t.\tGT\t1|2\t0|2\t1|1\n'         # deletion + (complex-looking) longer ALT
  '1\t98\t.\tTTT\tT\t.\tPASS\t.\tGT\t1|1\t1|1\t1|1\n'               # deletion reaching across the region end at 100
  '1\t99\t.\tTG\tCA\t.\tPASS\t.\tGT\t0|0\t1|0\t0|0\n'               # MNP overlapping the region end
  '1\t500\t.\tG\tT\t.\tPASS\t.\tGT\t1|1\t1|1\t1|1\n'                # outside every region
  '2\t7\t.\tGCA\tG\t.\tPASS\t.\tGT\t1|1\t1|0\t0|1\n'
  '2\t15\t.\tAC\tTG\t.\tPASS\t.\tGT\t1/1\t0/0\t0/1\n'               # unphased separator
  '2\t30\t.\tA\tG\t.\tPASS\t.\tGT\t0|0\t0|0\t0|0\n'                 # nobody carries it: still written
  'X\t5\t.\tTA\tCG\t.\tPASS\t.\tGT\t1\t0\t1\n'                      # haploid
  'X\t9\t.\tT\tTAAAA\t.\tPASS\t.\tGT\t1\t1\t0\n'
)
BED = '1\t0\t100\n2\t0\t50\n2\t10\t20\nX\t0\t20\n'      # chromosome 2 twice: overlapping regions write a record twice


def main():
  tmp = tempfile.mkdtemp(prefix='fgold')
  vin, bed = os.path.join(tmp, 'in.vcf'), os.path.join(tmp, 'r.bed')
  open(vin, 'w').write(HDR + BODY)
  open(bed, 'w').write(BED)
  kept = {}
  for sample in ('S0', 'S1', 'S2'):
    vout = os.path.join(tmp, sample + '.vcf')
    vio.prepare_variant_file(vin, sample, bed, vout)
    import gc
    gc.collect()     # the reference never closes vcf_out; the shim flushes on collection
    recs = [l.split('\t') for l in open(vout).read().split('\n') if l and not l.startswith('#')]
    kept[sample] = [[r[0], int(r[1]), r[3], r[4], r[9]] for r in recs]
  with open(os.path.join(HERE, 'filter_variants.json'), 'w') as fp:
    json.dump({'vcf': HDR + BODY, 'bed': BED, 'kept': kept}, fp, indent=1, sort_keys=True)
  print(json.dumps(kept, indent=1))


if __name__ == '__main__':
  main()
