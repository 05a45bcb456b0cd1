"""corrupt-reads' input streams must work on FIFOs and /dev/fd/N process substitutions: the
reference's own example feeds it from FIFOs (examples/reads/run.sh:13-16).  One open() per input,
gzip sniffed from the first bytes of that same stream (host logic; no GPU needed)."""
import gzip
import os
import threading

import numpy as np
import pytest

from mitty_b200.simulation import readcorrupt as rc

REC = b'@q%d\nACGTACGTAC\n+\n~~~~~~~~~~\n'


def _payload(n=5000):
  return b''.join(REC % k for k in range(n))


def _feed(path, data):
  with open(path, 'wb') as fp:          # blocks until the reader has opened the FIFO
    fp.write(data)


@pytest.mark.parametrize('gz', [False, True])
@pytest.mark.timeout(60)
def test_stream_reads_a_fifo_once(tmp_path, gz):
  data = _payload()
  fifo = str(tmp_path / 'in.fifo')
  os.mkfifo(fifo)
  t = threading.Thread(target=_feed, args=(fifo, gzip.compress(data) if gz else data), daemon=True)
  t.start()
  st = rc._Stream(fifo, np.zeros(4096, dtype=np.uint8))     # a second open() here would deadlock / lose bytes
  got = []
  while True:
    st.refill()
    if st.fill == 0:
      break
    got.append(st.buf[:st.fill].tobytes())
    st.consume(st.fill)
  st.close()
  t.join(timeout=10)
  assert b''.join(got) == data


@pytest.mark.timeout(60)
def test_stream_reads_dev_fd(tmp_path):
  """/dev/fd/N (what `<(cat < tf1)` expands to): re-opening a pipe's /dev/fd entry shares the read
  position with the first open, so sniffing with a separate open() would eat the first two bytes."""
  data = _payload(100)
  r, w = os.pipe()
  t = threading.Thread(target=lambda: (os.write(w, data), os.close(w)), daemon=True)
  t.start()
  st = rc._Stream('/dev/fd/{}'.format(r), np.zeros(1 << 16, dtype=np.uint8))
  st.refill()
  assert st.buf[:st.fill].tobytes() == data and st.eof
  st.close()
  os.close(r)
  t.join(timeout=10)
