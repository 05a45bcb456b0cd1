"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference is imported as is
(``sys.path`` -> /root/reference); ``oracle/refshim/pysam.py`` stands in for the three pysam I/O
classes it uses (pysam/htslib is not installed; it carries no arithmetic).  Every file this
script writes is committed; the GPU box never sees /root/reference.

    python tests/golden/make_golden.py [--full]

--full additionally runs corrupt-reads over the full config-1 FASTQ (about 6 minutes).
"""
import gzip
import hashlib
import json
import os
import pickle
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
REF = '/root/reference'
warnings.filterwarnings('ignore')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'refshim'))
sys.path.insert(0, REF)

import mitty.lib.vcfio as vio                     # noqa: E402  (the reference)
import mitty.simulation.illumina as il             # noqa: E402
import mitty.simulation.readgenerate as rg         # noqa: E402
import mitty.simulation.readcorrupt as rc          # noqa: E402
import mitty.simulation.rpc as rpc                 # noqa: E402
from mitty_b200 import synth                       # noqa: E402


def sha(path):
  h = hashlib.sha256()
  with open(path, 'rb') as fp:
    for blk in iter(lambda: fp.read(1 << 20), b''):
      h.update(blk)
  return h.hexdigest()


def model(name):
  return pickle.load(open(os.path.join(REF, 'mitty', 'data', 'readmodels', name), 'rb'))


def run_pair(wl, tmp, tag, model_name, coverage=30.0, seed=7, corrupt=True, gz=False):
  """reference generate-reads (+ corrupt-reads) with threads=1 -> paths + stats."""
  fa, vcf, bed = synth.write_workload(wl, os.path.join(tmp, tag), gz=gz)
  m = model(model_name)
  p = {k: os.path.join(tmp, '{}.{}.fq'.format(tag, k)) for k in ('r1', 'r2', 'c1', 'c2')}
  rg.process_multi_threaded(fa, vcf, wl['sample'], bed, il, m, coverage, p['r1'], p['r2'], threads=1, seed=seed)
  if corrupt:
    rc.multi_process(il, m, p['r1'], p['c1'], p['r2'], p['c2'], processes=1, seed=seed)
  else:
    p.pop('c1'); p.pop('c2')
  info = {k: {'sha256': sha(v), 'bytes': os.path.getsize(v)} for k, v in p.items()}
  info['pairs'] = sum(1 for _ in open(p['r1'])) // 4
  info.update(model=model_name, coverage=coverage, seed=seed)
  return p, info


def corrupt_stats(c_paths, r_paths, rlen):
  """Per (mate, cycle) BQ histogram and substitution counts from a reference corrupt run."""
  hist = np.zeros((2, rlen, 94), dtype=np.uint32)
  err = np.zeros((2, rlen), dtype=np.uint32)
  for m in (0, 1):
    with open(c_paths[m]) as fc, open(r_paths[m]) as fr:
      while True:
        if not fc.readline():
          break
        fr.readline()
        sc, sr = fc.readline().rstrip('\n'), fr.readline().rstrip('\n')
        fc.readline(); fr.readline()
        q = np.frombuffer(fc.readline().rstrip('\n').encode(), dtype=np.uint8) - 33
        fr.readline()
        hist[m, np.arange(q.size), q] += 1
        a, b = np.frombuffer(sc.encode(), dtype=np.uint8), np.frombuffer(sr.encode(), dtype=np.uint8)
        err[m, :a.size] += (a != b)
  return hist, err


def main():
  full = '--full' in sys.argv
  G = {}
  tmp = tempfile.mkdtemp(prefix='golden')

  # ---- a6 / a8: seed schedule KATs (readgenerate.py:129-159, illumina.py:56-58)
  fake_vcf = [{'v': [[], []]}, {'v': [[], []]}]
  units = list(rg.get_data_for_workers({'passes': 2}, fake_vcf, 7))
  G['schedule_seed7_2x2x2'] = [[u['region_idx'], u['region_cpy'], int(u['rng_seed'])] for u in units]
  fake3 = [{'v': [[], [], []]}, {'v': [[]]}, {'v': [[], []]}]
  units = list(rg.get_data_for_workers({'passes': 4}, fake3, 123456789))
  G['schedule_seed123456789_3-1-2x4'] = [[u['region_idx'], u['region_cpy'], int(u['rng_seed'])] for u in units]
  G['unit_seed_split'] = {str(s): [int(x) for x in np.random.RandomState(s).randint(il.SEED_MAX, size=4)]
                          for s in (1882953283, 0, 4294967295, 7)}
  G['corrupt_worker_seeds_seed7'] = [int(np.random.RandomState(7).randint(il.SEED_MAX))]

  # ---- a5: read_model_params
  G['read_model_params'] = {}
  for name in ('1kg-pcr-free.pkl', 'hiseq-X-v2.5-Garvan.pkl', 'hiseq-2500-v1-pcr-free.pkl'):
    for cov in (30.0, 60.0, 5.0, 0.5, 200.0):
      rm = il.read_model_params(model(name), cov)
      G['read_model_params']['{}@{}'.format(name, cov)] = {'p': rm['p'], 'passes': rm['passes'], 'rlen': int(rm['rlen'])}

  # ---- a9 / a10: generate_reads (template sampling) -> npz
  tpl = {}
  for name, p_min, p_max, seed in (('hiseq-X-v2.5-Garvan.pkl', 1, 100001, 1882953283),
                                   ('1kg-pcr-free.pkl', 20001, 1020051, 976413892),
                                   ('hiseq-X-v2.5-Garvan.pkl', 5, 905, 3)):
    rm = il.read_model_params(model(name), 30.0)
    r = il.generate_reads(rm, p_min, p_max, seed)
    key = '{}_{}_{}_{}'.format(name[:-4], p_min, p_max, seed)
    tpl[key + '_pos0'] = r[0]['pos']; tpl[key + '_pos1'] = r[1]['pos']
    tpl[key + '_fo0'] = r[0]['file_order']; tpl[key + '_fo1'] = r[1]['file_order']
    assert (r[0]['len'] == rm['rlen']).all() and r[0]['len'].dtype == np.uint32
  np.savez_compressed(os.path.join(HERE, 'templates.npz'), **tpl)
  try:
    il.generate_reads(il.read_model_params(model('1kg-pcr-free.pkl'), 30.0), 1, 1000, 1 << 32)
    G['seed_out_of_range'] = 'no error'
  except ValueError as e:
    G['seed_out_of_range'] = str(e)

  # ---- a2-a4, a7: variant loading + node lists on the edge workload
  edge = synth.edge_workload()
  fa, vcf, bed = synth.write_workload(edge, os.path.join(tmp, 'edge'))
  vdf = vio.load_variant_file(vcf, edge['sample'], bed)
  G['edge_variants'] = [{'region': list(r['region']), 'v': [[list(v.tuple()) for v in cp] for cp in r['v']]} for r in vdf]
  import pysam
  fasta = pysam.FastaFile(fa)
  nodes = {}
  for ri, r in enumerate(vdf):
    ref_seq = fasta.fetch(reference=r['region'][0], start=r['region'][1], end=r['region'][2])
    for cpy, vl in enumerate(r['v']):
      nl = rpc.create_node_list(ref_seq, r['region'][1] + 1, vl)
      nodes['{}_{}'.format(ri, cpy)] = [[n.ps, n.pr, n.cigarop, n.oplen, hashlib.md5(n.seq.encode()).hexdigest()[:8] if len(n.seq) > 40 else n.seq, n.v] for n in nl]
  G['edge_nodes'] = nodes

  # ---- a12: reads through every node of the edge workload (explicit positions)
  reads = {}
  r = vdf[0]
  ref_seq = fasta.fetch(reference=r['region'][0], start=r['region'][1], end=r['region'][2])
  nl = rpc.create_node_list(ref_seq, r['region'][1] + 1, r['v'][1])
  rs = np.random.RandomState(5)
  p_min, p_max = nl[0].ps, nl[-1].ps + nl[-1].oplen
  pl = np.sort(rs.randint(p_min, p_max - 160, size=400)).astype(np.int64)
  for L in (150, 37):
    n0, n1 = rpc.get_begin_end_nodes(pl, np.full(pl.size, L, dtype=np.uint32), nl)
    reads['edge0_1_L{}'.format(L)] = [[int(p)] + list(rpc.generate_read(int(p), L, int(a), int(b), nl)) + [int(a), int(b)]
                                      for p, a, b in zip(pl, n0, n1)]
  G['edge_reads'] = reads

  # ---- a13-a18: full FASTQ pairs
  files = {}
  p, info = run_pair(edge, tmp, 'edge', 'hiseq-X-v2.5-Garvan.pkl', coverage=30.0, seed=7)
  for k, v in p.items():
    with open(v, 'rb') as fi, gzip.GzipFile(os.path.join(HERE, 'edge.{}.fq.gz'.format(k)), 'wb', mtime=0) as fo:
      fo.write(fi.read())
  files['edge'] = info
  hist, err = corrupt_stats((p['c1'], p['c2']), (p['r1'], p['r2']), 150)

  p, info = run_pair(edge, tmp, 'edge250', '1kg-pcr-free.pkl', coverage=20.0, seed=99)
  files['edge250'] = info

  # soft-masked reference (half of the bases lower case): sha256 of the four files
  p, info = run_pair(synth.softmask_workload(), tmp, 'softmask', 'hiseq-X-v2.5-Garvan.pkl', coverage=30.0, seed=11)
  files['softmask'] = info

  mid = synth.config1(contig_len=100000)
  p, info = run_pair(mid, tmp, 'mid', 'hiseq-X-v2.5-Garvan.pkl')
  files['mid'] = info
  hist, err = corrupt_stats((p['c1'], p['c2']), (p['r1'], p['r2']), 150)
  np.savez_compressed(os.path.join(HERE, 'mid_corrupt_stats.npz'), bq_hist=hist, sub_count=err, pairs=info['pairs'])
  # coverage histogram of template starts per 1 kb bin (region 0) for the statistical tests
  starts = []
  with open(p['r1']) as fp:
    for i, line in enumerate(fp):
      if i % 4 == 0:
        ri = rg.parse_qname(line[1:].strip())
        if ri[0].chrom == '1':
          starts.append(min(ri[0].pos, ri[1].pos))
  np.savez_compressed(os.path.join(HERE, 'mid_coverage.npz'), hist=np.histogram(starts, bins=100, range=(0, 100000))[0])

  c1 = synth.config1()
  p, info = run_pair(c1, tmp, 'config1', 'hiseq-X-v2.5-Garvan.pkl', corrupt=full)
  if not full and os.path.exists(os.path.join(HERE, 'golden.json')):
    # keep the corrupted-file hashes of an earlier --full run while the perfect files are unchanged
    old = json.load(open(os.path.join(HERE, 'golden.json')))['fastq'].get('config1', {})
    if all(old.get(k) == info[k] for k in ('r1', 'r2')):
      info.update({k: old[k] for k in ('c1', 'c2') if k in old})
  files['config1'] = info
  G['fastq'] = files

  # ---- a17/a18 single-template plugin calls (illumina.corrupt_template)
  m = model('hiseq-X-v2.5-Garvan.pkl')
  rs = np.random.RandomState(42)
  tcases = []
  for L in (150, 100, 1):
    s1 = ''.join('ACGTN'[i] for i in rs.randint(0, 5, size=L))
    s2 = ''.join('ACGTacgtNR'[i] for i in rs.randint(0, 10, size=L))
    out = il.corrupt_template(m, ('q{}'.format(L), s1, s2), np.random.RandomState(1000 + L))
    tcases.append({'in': ['q{}'.format(L), s1, s2], 'seed': 1000 + L, 'out': [list(o) for o in out]})
  G['corrupt_template'] = tcases

  with open(os.path.join(HERE, 'golden.json'), 'w') as fp:
    json.dump(G, fp, indent=1, sort_keys=True)
  print(json.dumps(files, indent=1))


if __name__ == '__main__':
  main()
