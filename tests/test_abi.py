"""The C-ABI library loads and exports every symbol include/mitty_b200.h declares; without a CUDA
device the engine fails loudly instead of falling back to the CPU (no compute is called here)."""
import ctypes
import os
import re

import pytest

from mitty_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'mitty_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(mg_[a-z_0-9]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
  syms = declared_symbols()
  assert len(syms) >= 17 and 'mg_unit_generate' in syms and 'mg_corrupt_fastq' in syms
  assert os.path.exists(_lib.SO_PATH), 'build the library first: python __graft_entry__.py'
  L = ctypes.CDLL(_lib.SO_PATH)
  for s in syms:
    assert hasattr(L, s), s
  assert sorted(_lib.SYMBOLS) == syms     # the Python binding covers the whole header


def test_unit_desc_layout_matches_header():
  # 8-byte fields first, then the 4-byte pair, as in the header's mg_unit_desc
  assert ctypes.sizeof(_lib.UnitDesc) == 104
  assert _lib.UnitDesc.p_max.offset == 96 and _lib.UnitDesc.corrupt.offset == 80


def test_no_cpu_fallback():
  import torch
  if torch.cuda.is_available():
    pytest.skip('a GPU is present')
  from mitty_b200.engine import Engine
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    Engine(0)


def test_product_does_not_import_the_oracle():
  """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
  pkg = os.path.join(ROOT, 'mitty_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        src = open(os.path.join(dirpath, f)).read()
        assert not re.search(r'^\s*(import|from)\s+oracle\b', src, flags=re.M), os.path.join(dirpath, f)
        assert 'liboracle' not in src and 'mitty_oracle' not in src, os.path.join(dirpath, f)
