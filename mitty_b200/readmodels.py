"""Read-model lookup: Mitty's ``get_read_model`` (mitty/cli.py:255-278).

A model is a dict with (at least) ``model_class``, ``mean_rlen``, ``cum_tlen`` f64[1000] and
``cum_bq_mat`` f64[2, 300, 94] (writer: mitty/empirical/bam2illumina.py:116-129).  Builtin names
(``hiseq-X-v2.5-Garvan.pkl`` ...) resolve first, exactly like the reference; anything else is
treated as a literal path to a pickled model.  The builtin models ship re-encoded as ``.npz``
(tools/import_read_models.py); the values are identical to the reference's ``.pkl`` files.
"""
import logging
import os
import pickle

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'readmodels')
_SCALARS = ('model_class', 'model_description', 'mean_rlen', 'min_rlen', 'max_rlen', 'r_cnt', 'min_mq')


def builtin_models():
  return sorted(f[:-4] + '.pkl' for f in os.listdir(_DIR) if f.endswith('.npz'))


def _builtin_path(modelfile):
  base = os.path.basename(modelfile)
  if base != modelfile:
    return None
  stem = base[:-4] if base.endswith('.pkl') else base
  p = os.path.join(_DIR, stem + '.npz')
  return p if os.path.exists(p) else None


def load_model(modelfile):
  p = _builtin_path(modelfile)
  if p is not None:
    logging.debug('Found model {} in builtins'.format(modelfile))
    with np.load(p, allow_pickle=False) as z:
      model = {k: z[k] for k in z.files}
    for k in _SCALARS:
      if k in model:
        model[k] = model[k].item()
    return model
  logging.debug('Treating {} as literal path to model file'.format(modelfile))
  with open(modelfile, 'rb') as fp:
    return pickle.load(fp)


def get_read_model(modelfile):
  """-> (read_module, model), the reference's contract (cli.py:255-278): the module is chosen by
  ``model['model_class']``; an unknown class yields ``None`` (and an AttributeError later)."""
  import mitty_b200.simulation.illumina as illumina
  model = load_model(modelfile)
  read_module = {'illumina': illumina}.get(model['model_class'])
  return read_module, model
