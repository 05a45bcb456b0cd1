"""corrupt-reads: ``mitty/simulation/readcorrupt.py`` on the GPU.

Same entry point as the reference (``multi_process``, readcorrupt.py:18).  The FASTQ pair is
indexed (newline scan), sized (exclusive scan) and corrupted on the device
(``k_nl_*``, ``k_corrupt_sizes``, ``k_corrupt_staged``; ``k_corrupt`` for the deterministic draws);
input qualities are discarded and read 1's name is written to both files, as the reference does
(readcorrupt.py:55, 113).

Page-locked memory: 2 input + 4 output buffers of ``chunk_bytes`` / 2 x ``chunk_bytes`` (2.5 GB with
the default 256 MB chunks; pass a smaller ``chunk_bytes`` on small hosts -- the result does not
depend on it).

Modes
  philox         production: Philox4x32-7 draws keyed by (seed, template index, file, cycle group);
                 one joint (quality, substitution) alias lookup per base (MgCorruptCtx, mg_core.cuh)
  deterministic  the single worker's numpy RandomState stream of the reference
                 (seed -> RandomState(seed).randint(SEED_MAX), readcorrupt.py:31,36,84) is drawn on
                 the host, read by read, and consumed on the device: byte-exact vs ``--threads 1``
"""
import logging
import time

import numpy as np

from mitty_b200.engine import MODE_DET, MODE_PHILOX, SEED_MAX, Engine

logger = logging.getLogger(__name__)


def open_fastq(fname):
  """One open() per input: the inputs may be FIFOs or /dev/fd/N process substitutions
  (examples/reads/run.sh:13-16 of the reference feeds corrupt-reads from FIFOs), where a second
  open() blocks for ever or loses the bytes already read.  Nothing is read here: see sniff_gzip."""
  import io
  return io.open(fname, 'rb', buffering=1 << 20)      # BufferedReader: peek() does not consume


def sniff_gzip(fp):
  """gzip is recognised from the first bytes of the one open stream -- at the first READ, not at
  open time: a writer that opens its two FIFOs one after the other only starts to write once BOTH
  have a reader, so blocking on the first input's bytes before opening the second would deadlock."""
  import gzip
  if fp.peek(2)[:2] == b'\x1f\x8b':
    return gzip.GzipFile(fileobj=fp, mode='rb')
  return fp


def seq_lengths(buf):
  """Lengths of the sequence lines of a 4-line-record FASTQ buffer (host side, numpy)."""
  nl = np.flatnonzero(buf == 10)
  n_rec = nl.size // 4
  return (nl[1:4 * n_rec:4] - nl[0:4 * n_rec:4] - 1).astype(np.int64)


class _Stream(object):
  """A FASTQ file read in chunks into one (pinned) buffer; the unconsumed tail is carried over."""

  def __init__(self, fname, buf):
    self.fp = open_fastq(fname)
    self.sniffed = False
    self.buf, self.fill, self.eof = buf, 0, False

  def close(self):
    self.fp.close()

  def refill(self):
    if not self.sniffed:
      self.fp, self.sniffed = sniff_gzip(self.fp), True
    mv = memoryview(self.buf)
    while self.fill < self.buf.size and not self.eof:
      n = self.fp.readinto(mv[self.fill:])
      if not n:
        self.eof = True
        if self.fill and self.buf[self.fill - 1] != 10 and self.fill < self.buf.size:
          self.buf[self.fill] = 10; self.fill += 1           # last line without a newline
      else:
        self.fill += n

  def consume(self, n):
    rest = self.fill - n
    if rest:
      self.buf[:rest] = self.buf[n:self.fill].copy()
    self.fill = rest


def multi_process(read_module, read_model, fastq1_in, fastq1_out, fastq2_in=None, fastq2_out=None, processes=2, seed=7,
                  mode='philox', device=0, chunk_bytes=256 << 20):
  """Same signature as the reference (readcorrupt.py:18) plus keyword-only extras.  ``processes``
  is accepted for command-line compatibility (one GPU does all the work).  Note the reference
  passes the RAW model dict here (cli.py:155-157), and so does the CLI of this package.

  The files are streamed in chunks through pinned buffers: the device indexes the records of each
  chunk, corrupts the complete templates and reports how many input bytes they occupied; the rest
  is carried into the next chunk.  The output does not depend on the chunk size.

  Page-locked host memory: one input buffer of ``chunk_bytes`` and two output buffers of twice that per file,
  i.e. 10 x chunk_bytes for a pair (2.5 GB at the default 256 MB; ``chunk_bytes`` scales it down)."""
  t0 = time.time()
  engine = Engine(device)
  cnt = 0
  try:
    rlen = read_model['mean_rlen'] if 'mean_rlen' in read_model else read_model['rlen']
    engine.load_model(read_model, rlen=rlen)
    paired = fastq2_in is not None
    ins = [_Stream(fastq1_in, engine.pinned(chunk_bytes))] + ([_Stream(fastq2_in, engine.pinned(chunk_bytes))] if paired else [])
    # two output buffer sets: the files of chunk k are written (one thread per file) while chunk k+1
    # is read (one thread per file) and corrupted
    outs = [(engine.pinned(2 * chunk_bytes + 64), engine.pinned(2 * chunk_bytes + 64) if paired else None) for _ in range(2)]
    pending = [[], []]
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=4)
    k = 0
    rng = None
    if mode == 'deterministic':
      worker_seed = np.random.RandomState(seed).randint(SEED_MAX)   # worker 0, readcorrupt.py:31,36
      rng = np.random.RandomState(worker_seed)
    fps = [open(fastq1_out, 'wb')] + ([open(fastq2_out, 'wb')] if paired and fastq2_out is not None else [])
    try:
      while True:
        for fu in [pool.submit(st.refill) for st in ins]:
          fu.result()
        a1 = ins[0].buf[:ins[0].fill]
        a2 = ins[1].buf[:ins[1].fill] if paired else None
        if a1.size == 0 or (paired and a2.size == 0):
          break
        draws = None
        if rng is not None:
          l1 = seq_lengths(a1)
          if paired:
            l2 = seq_lengths(a2)
            n = min(l1.size, l2.size)
            lens = np.empty(2 * n, dtype=np.int64); lens[0::2] = l1[:n]; lens[1::2] = l2[:n]
          else:
            lens = l1
          draws = read_module.corrupt_draws(lens.tolist(), rng)
        for fu in pending[k]:                                   # this buffer set's previous writes
          fu.result()
        pending[k] = []
        o1, o2, n, c1, c2 = engine.corrupt_fastq(a1, a2, mode=MODE_DET if rng is not None else MODE_PHILOX, seed=seed, draws=draws,
                                                 first_template=cnt, out=outs[k], partial=True)
        if n == 0:
          if all(st.eof for st in ins):
            break
          raise ValueError('a FASTQ record is larger than the chunk size ({} bytes)'.format(chunk_bytes))
        pending[k] = [pool.submit(fp.write, memoryview(o)) for fp, o in zip(fps, (o1, o2))]
        k ^= 1
        ins[0].consume(c1)
        if paired:
          ins[1].consume(c2)
        cnt += n
    finally:
      for fu in pending[0] + pending[1]:
        fu.result()
      pool.shutdown(wait=True)
      for fp in fps:
        fp.close()
  finally:
    engine.close()
  t1 = time.time()
  logger.debug('Processed {} templates in {:0.2f}s ({:0.2f} t/s)'.format(cnt, t1 - t0, cnt / max(t1 - t0, 1e-9)))
