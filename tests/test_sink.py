"""The native output sink (mg_sink_*, mitty_b200/csrc/mg_sink.cpp): units committed out of order by
several producers must land in the files in schedule order -- regular files (pwrite), FIFOs (ordered
sequential writes) and gzip (one member per piece; gunzips to the same bytes).  Host logic only: on a
box without a CUDA driver the slots are ordinary memory."""
import gzip
import os
import threading

import numpy as np
import pytest

from mitty_b200.engine import Sink


def _units(n, rs, chunk):
  out = []
  for k in range(n):
    size = 0 if k % 7 == 3 else int(rs.randint(1, 5 * chunk))
    out.append((rs.randint(32, 127, size=size).astype(np.uint8), rs.randint(32, 127, size=size).astype(np.uint8)))
  return out


def _produce(sink, producer, mine, units, chunk, errs):
  try:
    for k in mine:                          # increasing schedule order, size announced before the pieces
      b1, b2 = units[k]
      sink.unit_size(k, b1.size)
      for off in range(0, b1.size, chunk):
        sink.put(producer, k, off, b1[off:off + chunk], b2[off:off + chunk])
  except BaseException as e:  # noqa: B902
    errs.append(e)


@pytest.mark.parametrize('gz', [0, 1])
@pytest.mark.parametrize('paired', [True, False])
@pytest.mark.timeout(120)
def test_regular_files_in_schedule_order(tmp_path, gz, paired):
  chunk, rs = 1000, np.random.RandomState(gz)
  units = _units(40, rs, chunk)
  p1, p2 = str(tmp_path / 'a.fq'), (str(tmp_path / 'b.fq') if paired else None)
  sink = Sink(p1, p2, len(units), n_producers=3, slots=2, chunk_bytes=chunk, gzip_level=gz, threads=3)
  # three producers with interleaved units: unit 39's pieces arrive long before unit 1 is complete
  assign = [list(range(0, 40, 3)), list(range(1, 40, 3)), list(range(2, 40, 3))]
  errs = []
  ts = [threading.Thread(target=_produce, args=(sink, i, assign[i], units, chunk, errs)) for i in range(3)]
  for t in ts[::-1]:
    t.start()
  for t in ts:
    t.join()
  assert not errs, errs
  w = sink.close()
  want1, want2 = b''.join(u[0].tobytes() for u in units), b''.join(u[1].tobytes() for u in units)
  rd = (lambda p: gzip.open(p, 'rb').read()) if gz else (lambda p: open(p, 'rb').read())
  assert rd(p1) == want1
  if paired:
    assert rd(p2) == want2
  if not gz:
    assert w == (len(want1), len(want2) if paired else 0)
  else:
    assert w[0] == os.path.getsize(p1)


@pytest.mark.timeout(120)
def test_fifo_targets_get_sequential_ordered_writes(tmp_path):
  chunk, rs = 777, np.random.RandomState(5)
  units = _units(25, rs, chunk)
  p1, p2 = str(tmp_path / 'a.fifo'), str(tmp_path / 'b.fifo')
  os.mkfifo(p1); os.mkfifo(p2)
  got = {}

  def drain(name, path):
    with open(path, 'rb') as fp:
      got[name] = fp.read()

  rt = [threading.Thread(target=drain, args=(n, p)) for n, p in (('a', p1), ('b', p2))]
  for t in rt:
    t.start()
  sink = Sink(p1, p2, len(units), n_producers=2, slots=2, chunk_bytes=chunk, threads=2)
  errs = []
  ts = [threading.Thread(target=_produce, args=(sink, i, list(range(i, 25, 2)), units, chunk, errs)) for i in (1, 0)]
  for t in ts:
    t.start()
  for t in ts:
    t.join()
  assert not errs, errs
  sink.close()
  for t in rt:
    t.join()
  assert got['a'] == b''.join(u[0].tobytes() for u in units) and got['b'] == b''.join(u[1].tobytes() for u in units)


@pytest.mark.timeout(60)
def test_errors_surface(tmp_path):
  with pytest.raises(OSError):
    Sink(str(tmp_path / 'no_such_dir' / 'a.fq'), None, 1)
  sink = Sink(str(tmp_path / 'a.fq'), None, 3, chunk_bytes=100)
  sink.unit_size(0, 10)
  sink.put(0, 0, 0, b'0123456789')
  with pytest.raises(IOError):
    sink.close()                            # units 1 and 2 were never written
  # a reader that goes away: the writer's error reaches the producer and close()
  fifo = str(tmp_path / 'gone.fifo')
  os.mkfifo(fifo)
  t = threading.Thread(target=lambda: open(fifo, 'rb').close())
  t.start()
  sink = Sink(fifo, None, 1, slots=2, chunk_bytes=1 << 16)
  t.join()
  sink.unit_size(0, 50 << 16)
  import signal
  old = signal.signal(signal.SIGPIPE, signal.SIG_IGN)
  try:
    with pytest.raises((RuntimeError, IOError)):
      for k in range(50):
        sink.put(0, 0, k << 16, bytes(1 << 16))
      sink.close()
      sink = None
  finally:
    signal.signal(signal.SIGPIPE, old)
    if sink is not None:
      try:
        sink.close()
      except IOError:
        pass


def _shared_worker(rank, world, p1, p2, table, n_units, chunk, seed, ready, go, out_q):
  """One PROCESS of a shared-table run: its own sink on the common files, units pulled dynamically."""
  try:
    rs = np.random.RandomState(seed)
    units = _units(n_units, rs, chunk)            # every process builds the same synthetic units
    if rank != 0:
      go.wait()                                   # the owner has created the table and truncated the outputs
    sink = Sink(p1, p2, n_units, n_producers=1, slots=2, chunk_bytes=chunk, threads=2, table=table, owner=(rank == 0))
    if rank == 0:
      ready.set()
    go.wait()
    mine = []
    while True:
      k = sink.next_unit()
      if k < 0:
        break
      mine.append(k)
      b1, b2 = units[k]
      sink.unit_size(k, b1.size)
      for off in range(0, b1.size, chunk):
        sink.put(0, k, off, b1[off:off + chunk], b2[off:off + chunk])
    sink.close()
    out_q.put((rank, mine, None))
  except BaseException as e:  # noqa: B902
    out_q.put((rank, [], repr(e)))


@pytest.mark.timeout(120)
def test_processes_share_one_pair_of_files(tmp_path):
  """world_size 3, one sink per process, one pair of output files: the shared table hands out the
  units and carries their sizes, every process pwrite()s its units at their final offsets."""
  import multiprocessing as mp
  ctx = mp.get_context('spawn')
  n_units, chunk, seed, world = 30, 500, 11, 3
  p1, p2, table = str(tmp_path / 'a.fq'), str(tmp_path / 'b.fq'), str(tmp_path / 'units.tbl')
  ready, go, q = ctx.Event(), ctx.Event(), ctx.Queue()
  ps = [ctx.Process(target=_shared_worker, args=(r, world, p1, p2, table, n_units, chunk, seed, ready, go, q)) for r in range(world)]
  ps[0].start()
  assert ready.wait(60)
  for p in ps[1:]:
    p.start()
  go.set()
  res = [q.get(timeout=60) for _ in range(world)]
  for p in ps:
    p.join(30)
  assert all(r[2] is None for r in res), res
  taken = sorted(k for r in res for k in r[1])
  assert taken == list(range(n_units))                       # every unit exactly once
  units = _units(n_units, np.random.RandomState(seed), chunk)
  assert open(p1, 'rb').read() == b''.join(u[0].tobytes() for u in units)
  assert open(p2, 'rb').read() == b''.join(u[1].tobytes() for u in units)


@pytest.mark.parametrize('target', ['file', 'gzip', 'fifo'])
@pytest.mark.timeout(120)
def test_streams_of_many_small_units(tmp_path, target):
  """A batch of small units leaves the device as ONE byte stream: slot-sized pieces that run through unit
  boundaries (mg_sink_commit_multi), from two producers whose units interleave in the schedule, mixed
  with ordinary single-unit pieces.  Empty units included."""
  chunk, rs = 1000, np.random.RandomState(3)
  n = 300
  units = []
  for k in range(n):
    size = 0 if k % 11 == 5 else int(rs.randint(1, 400)) if k % 50 else int(rs.randint(1500, 4000))
    units.append((rs.randint(32, 127, size=size).astype(np.uint8), rs.randint(32, 127, size=size).astype(np.uint8)))
  p1, p2 = str(tmp_path / 'a'), str(tmp_path / 'b')
  got = {}
  readers = []
  if target == 'fifo':
    os.mkfifo(p1); os.mkfifo(p2)

    def drain(name, path):
      with open(path, 'rb') as fp:
        got[name] = fp.read()
    readers = [threading.Thread(target=drain, args=(nm, p)) for nm, p in (('a', p1), ('b', p2))]
    for t in readers:
      t.start()
  sink = Sink(p1, p2, n, n_producers=2, slots=2, chunk_bytes=chunk, gzip_level=1 if target == 'gzip' else 0, threads=3)
  errs = []

  def produce(producer, mine):
    try:
      # stretches of units as streams; every 4th stretch unit by unit
      for j in range(0, len(mine), 37):
        part = mine[j:j + 37]
        if (j // 37) % 4 == 3:
          _produce(sink, producer, part, units, chunk, errs)
        else:
          sink.put_stream(producer, part, [units[k][0].size for k in part], np.concatenate([units[k][0] for k in part]),
                          np.concatenate([units[k][1] for k in part]))
    except BaseException as e:  # noqa: B902
      errs.append(e)
  # producer 0: runs of consecutive units (these merge into one piece); producer 1: the units in between
  mine0 = [k for k in range(n) if (k // 10) % 2 == 0]
  mine1 = [k for k in range(n) if (k // 10) % 2 == 1]
  ts = [threading.Thread(target=produce, args=(1, mine1)), threading.Thread(target=produce, args=(0, mine0))]
  for t in ts:
    t.start()
  for t in ts:
    t.join()
  assert not errs, errs
  sink.close()
  for t in readers:
    t.join(timeout=60)
  want1, want2 = b''.join(u[0].tobytes() for u in units), b''.join(u[1].tobytes() for u in units)
  if target == 'fifo':
    assert got['a'] == want1 and got['b'] == want2
  else:
    rd = (lambda p: gzip.open(p, 'rb').read()) if target == 'gzip' else (lambda p: open(p, 'rb').read())
    assert rd(p1) == want1 and rd(p2) == want2
