"""corrupt-reads: ``mitty/simulation/readcorrupt.py`` on the GPU.

Same entry point as the reference (``multi_process``, readcorrupt.py:18).  The FASTQ pair is
indexed (newline scan), sized (exclusive scan) and corrupted on the device
(``k_nl_*``, ``k_corrupt_sizes``, ``k_corrupt``); input qualities are discarded and read 1's name
is written to both files, as the reference does (readcorrupt.py:55, 113).

Modes
  philox         production: Philox4x32-10 draws keyed by (seed, template index, file, cycle)
  deterministic  the single worker's numpy RandomState stream of the reference
                 (seed -> RandomState(seed).randint(SEED_MAX), readcorrupt.py:31,36,84) is drawn on
                 the host, read by read, and consumed on the device: byte-exact vs ``--threads 1``
"""
import logging
import time

import numpy as np

from mitty_b200.engine import MODE_DET, MODE_PHILOX, SEED_MAX, Engine

logger = logging.getLogger(__name__)


def _read(fname):
  import gzip
  with open(fname, 'rb') as fp:
    magic = fp.read(2)
  if magic == b'\x1f\x8b':
    with gzip.open(fname, 'rb') as fp:
      return np.frombuffer(fp.read(), dtype=np.uint8)
  return np.fromfile(fname, dtype=np.uint8)


def seq_lengths(buf):
  """Lengths of the sequence lines of a 4-line-record FASTQ buffer (host side, numpy)."""
  nl = np.flatnonzero(buf == 10)
  n_rec = nl.size // 4
  return (nl[1:4 * n_rec:4] - nl[0:4 * n_rec:4] - 1).astype(np.int64)


def multi_process(read_module, read_model, fastq1_in, fastq1_out, fastq2_in=None, fastq2_out=None, processes=2, seed=7,
                  mode='philox', device=0):
  """Same signature as the reference (readcorrupt.py:18) plus keyword-only extras.  ``processes``
  is accepted for command-line compatibility (one GPU does all the work).  Note the reference
  passes the RAW model dict here (cli.py:155-157), and so does the CLI of this package."""
  t0 = time.time()
  engine = Engine(device)
  try:
    a1 = _read(fastq1_in)
    a2 = _read(fastq2_in) if fastq2_in is not None else None
    if a1.size and a1[-1] != 10:
      a1 = np.concatenate([a1, np.array([10], dtype=np.uint8)])
    if a2 is not None and a2.size and a2[-1] != 10:
      a2 = np.concatenate([a2, np.array([10], dtype=np.uint8)])
    rlen = read_model['mean_rlen'] if 'mean_rlen' in read_model else read_model['rlen']
    engine.load_model(read_model, rlen=rlen)
    draws = None
    if mode == 'deterministic':
      worker_seed = np.random.RandomState(seed).randint(SEED_MAX)   # worker 0, readcorrupt.py:31,36
      l1 = seq_lengths(a1)
      if a2 is not None:
        l2 = seq_lengths(a2)
        n = min(l1.size, l2.size)
        lens = np.empty(2 * n, dtype=np.int64); lens[0::2] = l1[:n]; lens[1::2] = l2[:n]
      else:
        lens = l1
      draws = read_module.corrupt_draws(lens.tolist(), np.random.RandomState(worker_seed))
    o1, o2, cnt = engine.corrupt_fastq(a1, a2, mode=MODE_DET if mode == 'deterministic' else MODE_PHILOX, seed=seed, draws=draws)
    with open(fastq1_out, 'wb') as fp:
      fp.write(memoryview(o1))
    if fastq2_out is not None and o2 is not None:
      with open(fastq2_out, 'wb') as fp:
        fp.write(memoryview(o2))
  finally:
    engine.close()
  t1 = time.time()
  logger.debug('Processed {} templates in {:0.2f}s ({:0.2f} t/s)'.format(cnt, t1 - t0, cnt / max(t1 - t0, 1e-9)))
