"""Re-encode Mitty's shipped empirical read models (pickled dicts, mitty/data/readmodels/*.pkl,
written by mitty/empirical/bam2illumina.py:116-129) as compressed .npz so the engine can offer the
same builtin model names without unpickling at import time.  The .pkl format itself stays
supported for user models (mitty_b200.cli.get_read_model).  Run once in the build container:

    python tools/import_read_models.py /root/reference/mitty/data/readmodels
"""
import glob
import os
import pickle
import sys

import numpy as np

src = sys.argv[1]
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mitty_b200', 'data', 'readmodels')
os.makedirs(dst, exist_ok=True)
for f in sorted(glob.glob(os.path.join(src, '*.pkl'))):
  m = pickle.load(open(f, 'rb'))
  out = os.path.join(dst, os.path.basename(f)[:-4] + '.npz')
  np.savez_compressed(out, **{k: np.asarray(v) for k, v in m.items()})
  print(out, os.path.getsize(out))
