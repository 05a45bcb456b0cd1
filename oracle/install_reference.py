"""Recipe: install the UNMODIFIED reference (alenzhao/Mitty, pure Python) from /root/reference into
baseline/_ref/ (git-ignored, travels to the GPU box with the snapshot), next to the pysam stand-in it
needs to import (oracle/refshim/pysam.py: I/O classes only -- pysam/htslib is not installed here).
bench.py times this copy as the CPU baseline (`cpu_baseline.kind = "reference"`, `--impl reference`).

    python oracle/install_reference.py        # build container only: needs /root/reference

The reference's source tree is read-only, so pip builds from a copy under /tmp; --no-deps because
its declared dependencies (pysam, matplotlib, ...) are either shimmed or not on the measured path.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..'))
REF = '/root/reference'
DEST = os.path.join(ROOT, 'baseline', '_ref')


def install(force=False):
  if not os.path.isdir(REF):
    return 'no /root/reference here (GPU box): using the prebuilt baseline/_ref' if os.path.isdir(DEST) else 'no reference available'
  if os.path.isdir(os.path.join(DEST, 'mitty')) and not force:
    shutil.copy(os.path.join(HERE, 'refshim', 'pysam.py'), os.path.join(DEST, 'pysam.py'))
    return 'already installed'
  tmp = tempfile.mkdtemp(prefix='mitty_ref_')
  src = os.path.join(tmp, 'reference')
  shutil.copytree(REF, src, symlinks=True, ignore=shutil.ignore_patterns('.git'))
  os.makedirs(DEST, exist_ok=True)
  cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--find-links', '/opt/wheelhouse',
         '--upgrade', '--target', DEST, src]
  r = subprocess.run(cmd, capture_output=True, text=True)
  shutil.rmtree(tmp, ignore_errors=True)
  if r.returncode != 0:
    return 'pip install failed: ' + (r.stderr or r.stdout)[-400:]
  shutil.copy(os.path.join(HERE, 'refshim', 'pysam.py'), os.path.join(DEST, 'pysam.py'))
  return 'installed'


if __name__ == '__main__':
  print(install(force='--force' in sys.argv))
