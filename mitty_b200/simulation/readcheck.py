"""check-reads: the god-aligner contract verified for every read.

Mitty's ``god-aligner`` writes each read into a BAM at the POS / CIGAR its qname states
(mitty/benchmarking/god_aligner.py:141-183 over ``parse_qname``, readgenerate.py:259-291).  This
command proves that those qnames are true for a FASTQ pair of PERFECT reads: every read is re-derived
on the GPU from (chrom, copy, strand, POS, CIGAR) + FASTA + VCF + BED (``k_roundtrip_check``) and
compared with the bases in the file.  The files are streamed in chunks like corrupt-reads' inputs."""
import logging
import time

import mitty_b200.lib.vcfio as vio
from mitty_b200.engine import CHECK_CODES, Checker, Engine
from mitty_b200.simulation.readcorrupt import _Stream

logger = logging.getLogger(__name__)


def check_fastq(fasta_fname, vcf_fname, sample_name, bed_fname, fastq1, fastq2=None, device=0, chunk_bytes=256 << 20, max_report=20,
                drop_end_deletions=False):
  """-> {'templates', 'reads', 'bad', 'examples': [(file, record, reason, qname)], 'seconds'}"""
  from mitty_b200.simulation.readgenerate import _without_end_crossing_deletions
  t0 = time.time()
  vcf_df = vio.load_variant_file(vcf_fname, sample_name, bed_fname)
  fasta = vio.FastaFile(fasta_fname)
  engine = Engine(device)
  chk = None
  try:
    chk = Checker(engine)
    for r in vcf_df:
      region = r['region']
      rid = engine.load_region(fasta.fetch(reference=region[0], start=region[1], end=region[2]), region[1])
      for cpy, vl in enumerate(r['v']):
        vl, _ = _without_end_crossing_deletions(vl, region, drop_end_deletions)
        chk.add_copy(engine.build_copy(rid, vl), region[0], cpy)
    paired = fastq2 is not None
    ins = [_Stream(fastq1, engine.pinned(chunk_bytes))] + ([_Stream(fastq2, engine.pinned(chunk_bytes))] if paired else [])
    templates = bad = 0
    examples = []
    while True:
      for st in ins:
        st.refill()
      a1 = ins[0].buf[:ins[0].fill]
      a2 = ins[1].buf[:ins[1].fill] if paired else None
      if a1.size == 0 or (paired and a2.size == 0):
        break
      n, nb, rep, c1, c2 = chk.check(a1, a2, max_report=max(0, max_report - len(examples)) or 1)
      if n == 0:
        if all(st.eof for st in ins):
          break
        raise ValueError('a FASTQ record is larger than the chunk size ({} bytes)'.format(chunk_bytes))
      for f, rec, code in rep:
        if len(examples) < max_report:
          buf = a1 if f == 0 else a2
          # the record's qname: the (4 * rec)-th line of this chunk
          pos = 0
          for _ in range(4 * rec):
            pos = int(buf[pos:].tobytes().index(b'\n')) + pos + 1
          end = int(buf[pos:].tobytes().index(b'\n')) + pos
          examples.append((f + 1, templates + rec, CHECK_CODES.get(code, str(code)), buf[pos:end].tobytes().decode(errors='replace')))
      templates += n; bad += nb
      ins[0].consume(c1)
      if paired:
        ins[1].consume(c2)
    for st in ins:
      st.close()
    return {'templates': templates, 'reads': templates * (2 if paired else 1), 'bad': bad, 'examples': examples, 'seconds': time.time() - t0}
  finally:
    if chk is not None:
      chk.close()
    engine.close()
