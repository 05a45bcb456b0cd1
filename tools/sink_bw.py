"""Throughput of the native output sink alone: host threads commit pieces of synthetic 'units' (no GPU
involved) into one pair of files.  python tools/sink_bw.py [target dir] [GB per file] [producers] [writer threads]
MG_SINK_PWRITE=1 switches the regular-file path from mapped windows back to pwrite()."""
import ctypes as C, json, os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from mitty_b200 import _lib
from mitty_b200.engine import Sink

target = sys.argv[1] if len(sys.argv) > 1 else '/dev/shm'
gb = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
producers = int(sys.argv[3]) if len(sys.argv) > 3 else 4
writers = int(sys.argv[4]) if len(sys.argv) > 4 else (os.cpu_count() or 8)
chunk = 64 << 20
unit_bytes = 8 * chunk
n_units = max(producers, int(gb * 2**30 / unit_bytes))
p1, p2 = (os.path.join(target, 'sinkbw.%d.fq' % k) if target != '/dev/null' else '/dev/null' for k in (1, 2))
sink = Sink(p1, p2, n_units, n_producers=producers, slots=6, chunk_bytes=chunk, threads=writers)
L = _lib.lib()
src = np.full(chunk, 65, dtype=np.uint8)


def produce(i):
  while True:
    k = sink.next_unit()
    if k < 0:
      return
    sink.unit_size(k, unit_bytes)
    for off in range(0, unit_bytes, chunk):
      a, b, slot = C.c_void_p(), C.c_void_p(), C.c_void_p()
      assert L.mg_sink_acquire(sink._h, i, C.byref(a), C.byref(b), C.byref(slot)) == 0
      # (the D2H copy would land here; the slots are left as they are: only the write side is measured)
      assert L.mg_sink_commit(sink._h, slot, k, off, chunk) == 0


t0 = time.perf_counter()
ts = [threading.Thread(target=produce, args=(i,)) for i in range(producers)]
[t.start() for t in ts]; [t.join() for t in ts]
w = sink.close()
dt = time.perf_counter() - t0
print(json.dumps({'target': target, 'gb_per_file': n_units * unit_bytes / 2**30, 'producers': producers, 'writer_threads': writers,
                  'mode': 'pwrite' if os.environ.get('MG_SINK_PWRITE') else 'mapped windows', 'gbs_total': (w[0] + w[1]) / dt / 1e9, 'cores': os.cpu_count()}))
for p in (p1, p2):
  if p != '/dev/null' and os.path.exists(p):
    os.remove(p)
