"""CPU emulation of the kernels' per-thread logic (mitty_b200/csrc/mg_core.cuh compiled with g++)
against the oracle: node lookup, POS/CIGAR/v_list formatting, 2-bit extraction + reverse
complement, the unaligned word-stream writer, exception patching, the N filter and the
serial-number placement arithmetic.  Runs without a GPU; the same header is what the CUDA kernels
execute, so a failure here is a kernel bug found before any GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from mitty_b200 import synth
from tests import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
PAD = 8


@pytest.fixture(scope='module')
def emul(tmp_path_factory):
  so = str(tmp_path_factory.mktemp('emul') / 'libemul.so')
  # MG_EMUL_SANITIZE=1 (run with LD_PRELOAD=$(g++ -print-file-name=libasan.so)): AddressSanitizer + UBSan over the
  # very header the kernels execute -- the bounds check of the per-thread logic this GPU-less box can do
  san = ['-fsanitize=address,undefined', '-fno-omit-frame-pointer', '-g'] if os.environ.get('MG_EMUL_SANITIZE') else []
  subprocess.check_call(['g++', '-O1', '-std=c++17', '-shared', '-fPIC'] + san + ['-x', 'c++', os.path.join(HERE, 'emul', 'emul.cpp'), '-o', so])
  lib = C.CDLL(so)
  lib.emul_unit.restype = C.c_int64
  lib.emul_digit_sum.restype = C.c_uint64
  lib.emul_digit_sum.argtypes = [C.c_uint64]
  return lib


def device_layout(ref, ref_start_pos, cv, blk_shift=8):
  """Build, from the ORACLE's node list, the arrays the kernels read (MgNode, packed hap, blk, exc)."""
  nodes = oracle.create_node_list(ref, ref_start_pos, cv)
  p_min = nodes[0][0]
  p_max = nodes[-1][0] + nodes[-1][3]
  hap = ''.join(n[4] for n in nodes)
  assert len(hap) == p_max - p_min
  nd = np.zeros(len(nodes), dtype=[('key', 'u4'), ('pr', 'i4'), ('oplen', 'i4'), ('op', 'u4')])
  for i, (ps, pr, op, oplen, seq, v) in enumerate(nodes):
    nd[i] = (ps - p_min + (1 if op == 'D' else 0), pr, oplen, ord(op))
  hb = np.frombuffer(hap.encode(), dtype=np.uint8)
  code = np.full(256, 4, dtype=np.uint8); code[[65, 67, 71, 84]] = [0, 1, 2, 3]
  soft = np.isin(hb, np.frombuffer(b'acgt', dtype=np.uint8))      # lower-case a/c/g/t keep their codes (case runs)
  c = np.where(soft, code[hb ^ 0x20], code[hb])
  n_words = (len(hap) + 15) // 16
  cc = np.zeros(n_words * 16, dtype=np.uint64); cc[:len(hap)] = np.where(c > 3, 0, c)
  words = (cc.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint64))).sum(axis=1).astype(np.uint32)
  packed = np.zeros(n_words + 2 * PAD, dtype=np.uint32); packed[PAD:PAD + n_words] = words
  # exception runs as k_pack_ref / k_exc_map make them: maximal runs of one non-ACGT byte, and
  # maximal runs of lower-case a/c/g/t with the MG_EXC_CASE marker (1) as their byte
  cls = np.where(soft, 1, np.where(c > 3, hb.astype(np.int64), 0))
  exc = []
  i = 0
  bad = np.flatnonzero(cls != 0)
  while i < bad.size:
    j = i
    while j + 1 < bad.size and bad[j + 1] == bad[j] + 1 and cls[bad[j + 1]] == cls[bad[i]]:
      j += 1
    exc.append((bad[i], j - i + 1, cls[bad[i]], 0)); i = j + 1
  ex = np.array(exc, dtype=np.uint32).reshape(-1, 4) if exc else np.zeros((1, 4), dtype=np.uint32)
  n_blk = (len(hap) >> blk_shift) + 1
  blk = (np.searchsorted(nd['key'], np.arange(n_blk, dtype=np.uint64) << blk_shift, side='right') - 1).astype(np.uint32)
  return dict(nodes=nd, packed=packed, exc=ex, n_exc=len(exc), blk=blk, n_blk=n_blk, blk_shift=blk_shift,
              p_min=p_min, p_max=p_max, hap=hap)


def run_emul(lib, lay, L, ts, tl, fo, prefix, mid, corrupt=None, maxw=None):
  cap = int(len(ts)) * (2 * L + 400) + 4096
  o1, o2 = np.zeros(cap, dtype=np.uint8), np.zeros(cap, dtype=np.uint8)
  nb = C.c_int64(0)
  ts_rel = np.ascontiguousarray(ts - lay['p_min'], dtype=np.int64)
  tl = np.ascontiguousarray(tl, dtype=np.int64); fo = np.ascontiguousarray(fo, dtype=np.int8)
  hap_ptr = lay['packed'].ctypes.data + 4 * PAD
  n = lib.emul_unit(C.c_void_p(hap_ptr), C.c_uint32(lay['p_max'] - lay['p_min']), C.c_void_p(lay['nodes'].ctypes.data),
                    C.c_int(lay['nodes'].size), C.c_void_p(lay['blk'].ctypes.data), C.c_int(lay['blk_shift']), C.c_int(lay['n_blk']),
                    C.c_void_p(lay['exc'].ctypes.data), C.c_int(lay['n_exc']), C.c_int(L), C.c_int64(len(ts)),
                    C.c_void_p(ts_rel.ctypes.data), C.c_void_p(tl.ctypes.data), C.c_void_p(fo.ctypes.data),
                    prefix.encode(), mid.encode(), C.c_void_p(o1.ctypes.data), C.c_void_p(o2.ctypes.data), C.c_int64(cap), C.byref(nb),
                    *((C.c_int(0), None, C.c_int(6), C.c_int(0), C.c_int(0), C.c_uint32(0), C.c_uint32(0)) if corrupt is None else
                      (C.c_int(1), C.c_void_p(corrupt['alias'].ctypes.data), C.c_int(corrupt['kshift']), C.c_int(corrupt['code9']),
                       C.c_int(corrupt['alias'].shape[1]), C.c_uint32(corrupt['k0']), C.c_uint32(corrupt['k1']))),
                    C.c_int((12 if L <= 161 else 21 if L <= 305 else 0) if maxw is None else maxw))
  assert n >= 0, n
  return o1[:nb.value].tobytes(), o2[:nb.value].tobytes(), n


def test_digit_sum_and_permutation(emul):
  acc = 0
  for m in range(0, 12000):
    assert emul.emul_digit_sum(m) == acc
    acc += len(str(m + 1))
  for m in (99999, 100000, 123456789, 10 ** 9):
    assert emul.emul_digit_sum(m) == sum((min(m, 10 ** d - 1) - 10 ** (d - 1) + 1) * d for d in range(1, len(str(m)) + 1))
  for n in (1, 2, 3, 4, 5, 17, 1000, 2048, 2049, 4097, 70001, 131072):      # odd and even bit counts, exact powers of two
    out = np.zeros(n, dtype=np.uint32)
    emul.emul_permute(C.c_uint32(n), C.c_uint32(12345), C.c_uint32(678), C.c_void_p(out.ctypes.data))
    assert np.array_equal(np.sort(out), np.arange(n, dtype=np.uint32))
    if n > 1000:
      assert np.corrcoef(out.astype(float), np.arange(n))[0, 1] < 0.05


def test_philox_known_answer(emul):
  """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
  out = (C.c_uint32 * 4)()
  emul.emul_philox(0, 0, 0, 0, 0, 0, out)
  assert [hex(x) for x in out] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
  emul.emul_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
  assert [hex(x) for x in out] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
  emul.emul_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
  assert [hex(x) for x in out] == ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_tiny_kats_through_device_logic(emul, tmp_path):
  """The reference's test_rpc.py KATs (incl. the read from inside an insertion) through the
  kernels' formatting code: POS/CIGAR/v_list are parsed back out of the qname."""
  from tests.test_oracle_golden import RPC_KAT_CPY0, RPC_KAT_CPY1
  from mitty_b200.lib import vcfio
  p = H.write_tiny(tmp_path)
  vcf = vcfio.load_variant_file(p['vcf'], 'g0_s0', p['whole_bed'])
  for cpy, kat in ((1, RPC_KAT_CPY1), (0, RPC_KAT_CPY0)):
    lay = device_layout(H.TINY_SEQ, 1, H.oracle_cv(vcf[0]['v'][cpy]), blk_shift=2)
    for pp, l, n0, n1, (pos, cigar, v_list, seq) in kat:
      f1, f2, n = run_emul(emul, lay, l, np.array([pp]), np.array([l]), np.array([0]), '@t:0:0:', '|1|{}'.format(cpy))
      if n == 0:   # template end beyond p_max: the reference's sampler would never produce it
        continue
      q, s = f1.decode().split('\n')[:2]
      d = q.split('|')
      assert (int(d[4]), d[6], [int(x) for x in d[7].split(',') if x], s) == (pos, cigar, v_list, seq)


@pytest.mark.parametrize('L,maxw', [(150, 12), (150, 21), (150, 0), (37, 12), (16, 0), (16, 12), (1, 12), (161, 12), (250, 21), (305, 21), (330, 0)])
def test_edge_units_match_oracle(emul, L, maxw):
  """Every (region, copy) of the edge workload, dense random templates, all four file orders."""
  regs = H.workload_regions(synth.edge_workload())
  rs = np.random.RandomState(L)
  total = 0
  for ri, r in enumerate(regs):
    for cpy, vl in enumerate(r['v']):
      cv = H.oracle_cv(vl)
      lay = device_layout(r['ref'], r['region'][1] + 1, cv)
      n = 1500
      ts = rs.randint(lay['p_min'] - 2, lay['p_max'], size=n).astype(np.int64)
      ts[:5] = lay['p_min'] + np.arange(5)                    # reads starting on the first bases
      tl = rs.randint(0, 3 * L + 40, size=n).astype(np.int64)
      tl[5:10] = L
      ts[5:10] = lay['p_max'] - L - 1 - np.arange(5)          # reads ending on the last bases
      fo = rs.randint(0, 2, size=n).astype(np.int8)
      f1, f2, cnt = run_emul(emul, lay, L, ts, tl, fo, '@EDGE:0:{}:'.format(ri), '|{}|{}'.format(r['region'][0], cpy), maxw=maxw)
      # the oracle takes the te<p_max survivors (what illumina.generate_reads would hand over)
      tlc = np.maximum(tl, L); te = ts + tlc
      keep = (te < lay['p_max']) & (ts >= lay['p_min'])
      o1, o2, ocnt = oracle.generate_unit(r['ref'], r['region'][1] + 1, cv, L, ts[keep], te[keep], fo[:keep.sum()],
                                          'EDGE:0:{}'.format(ri), r['region'][0], cpy)
      assert cnt == ocnt
      assert f1 == o1 and f2 == o2
      total += cnt
  assert total > (5000 if L < 200 else 3000)


def test_corrupt_call_matches_oracle(emul):
  m = H.model('hiseq-X-v2.5-Garvan.pkl')
  rs = np.random.RandomState(3)
  L = 150
  for mate in (0, 1):
    seq = ''.join('ACGTNacgtR'[i] for i in rs.randint(0, 10, size=L))
    r = np.random.RandomState(77 + mate)
    bq_rnd, call_rnd, base_rnd = r.rand(L), r.rand(L), r.randint(0, 3, size=L).astype(np.uint8)
    call_rnd[::7] *= 1e-4                                      # force some substitutions
    s = np.frombuffer(seq.encode(), dtype=np.uint8).copy(); q = np.zeros(L, dtype=np.uint8)
    rows = np.ascontiguousarray(m['cum_bq_mat'][mate])
    emul.emul_corrupt_det(C.c_void_p(s.ctypes.data), C.c_void_p(q.ctypes.data), C.c_int(L), C.c_void_p(rows.ctypes.data), C.c_int(94),
                          C.c_void_p(oracle.PHRED_P.ctypes.data), C.c_void_p(bq_rnd.ctypes.data), C.c_void_p(call_rnd.ctypes.data),
                          C.c_void_p(base_rnd.ctypes.data))
    # reference semantics restated with numpy (illumina.py:155-160)
    rot = {'A': 'CTG', 'C': 'ATG', 'T': 'ACG', 'G': 'ACT'}
    want_s, want_q = list(seq), []
    for n in range(L):
      bq = min(int(np.searchsorted(rows[n], bq_rnd[n])), 93)
      if call_rnd[n] < oracle.PHRED_P[bq]:
        want_s[n] = rot.get(seq[n], 'NNN')[base_rnd[n]]
      want_q.append(chr(bq + 33))
    assert s.tobytes().decode() == ''.join(want_s) and q.tobytes().decode() == ''.join(want_q)
    assert s.tobytes().decode() != seq


def test_token_numbers(emul):
  """mg_put_num (branch-free decimal text, two tokens, any start phase) == str(v), and nothing outside
  the written range is touched."""
  out = (C.c_uint8 * 32)()
  rs = np.random.RandomState(1)
  vals = [0, 1, 9, 10, 99, 100, 999, 1000, 9999, 10000, 99999, 100000, 9999999, 10000000, 99999999, 100000000, 999999999,
          1000000000, 2147483647, 4294967295, 249250621, 150] + [int(x) for x in rs.randint(0, 1 << 31, size=300)] + \
         [int(10 ** rs.uniform(0, 9.6)) for _ in range(300)]
  for k, v in enumerate(vals):
    for pre in (b'', b'|', b'|1|', b',-'):
      for small in (0, 1):
        n = emul.emul_put_num(C.c_uint32(v), C.c_uint32(int.from_bytes(pre, 'little')), C.c_uint32(len(pre)), C.c_int(k % 4), out, C.c_int(small))
        assert n >= 0 and bytes(out[:n]) == pre + str(v).encode(), (v, pre, n, small, bytes(out[:max(n, 0)]))


@pytest.mark.parametrize('L,rows,maxw', [(150, 150, 12), (150, 300, 0), (37, 150, 21), (5, 300, 12), (150, 300, 12), (3, 150, 12), (2, 300, 0),
                                         (4, 150, 12), (16, 150, 12), (17, 150, 0), (33, 150, 12), (161, 300, 12), (250, 300, 21)])
def test_fused_philox_corruption_matches_numpy_spec(emul, L, rows, maxw):
  """Fused production-mode corruption (inline in the emit path: one Philox word and one joint alias
  lookup per base) == the numpy restatement of its draw layout applied to the perfect reads, incl.
  N / exception bases.  rows = 150: 8-bit codes; rows = 300 includes the all-zero rows beyond the
  model's max_rlen (quality 93): 9-bit codes."""
  from tests import philox_ref as PR
  m = H.model('hiseq-X-v2.5-Garvan.pkl')
  kshift, code9 = PR.table_shape(m['cum_bq_mat'], oracle.PHRED_P, rows)
  assert code9 == (1 if rows > 150 else 0) and kshift == 7
  alias = PR.joint_tables(m['cum_bq_mat'], oracle.PHRED_P, kshift, code9, n_rows=rows)
  # a row encodes the model's joint distribution of (quality, substitution): P(q) from the
  # searchsorted-left outcomes, a miscall with probability phred_p[q], each alternative a third of it
  qb = 7 if code9 else 6
  for mate in (0, 1):
    pm = np.diff(np.concatenate([np.zeros((150, 1)), m['cum_bq_mat'][mate, :150, :]], axis=1), axis=1)
    for cyc in (0, 1, 75, 149):
      d = PR.row_distribution(alias[mate, cyc], kshift, code9)
      tol = 2.0 ** -(13 if code9 else 15)
      assert code9 or pm[cyc][64:].sum() == 0.0
      for q in range(94 if code9 else 64):
        assert abs(d.get(q, 0.0) - pm[cyc][q] * (1.0 - oracle.PHRED_P[q])) < tol
        for sub in (1, 2, 3):
          assert abs(d.get((sub << qb) | q, 0.0) - pm[cyc][q] * oracle.PHRED_P[q] / 3.0) < tol
      assert abs(sum(d.values()) - 1.0) < 1e-9
  regs = H.workload_regions(synth.edge_workload())
  r = regs[0]
  cv = H.oracle_cv(r['v'][1])
  lay = device_layout(r['ref'], r['region'][1] + 1, cv)
  rs = np.random.RandomState(4)
  n = 1200
  ts = rs.randint(lay['p_min'], lay['p_max'] - 3 * L, size=n).astype(np.int64)
  tl = rs.randint(L, 3 * L, size=n).astype(np.int64)
  fo = rs.randint(0, 2, size=n).astype(np.int8)
  p1, p2, cnt = run_emul(emul, lay, L, ts, tl, fo, '@E:0:0:', '|e|1', maxw=maxw)
  cor = dict(alias=alias, kshift=kshift, code9=code9, k0=12345, k1=0xdeadbeef)
  c1, c2, ccnt = run_emul(emul, lay, L, ts, tl, fo, '@E:0:0:', '|e|1', corrupt=cor, maxw=maxw)
  assert ccnt == cnt and cnt > 800
  assert c1 == PR.corrupt_file(p1, 0, (alias, kshift, code9), cor['k0'], cor['k1'])
  assert c2 == PR.corrupt_file(p2, 1, (alias, kshift, code9), cor['k0'], cor['k1'])
  assert c1 != p1 and b'N' in c1


@pytest.mark.parametrize('L,maxw', [(150, 12), (37, 0)])
def test_soft_masked_haplotype(emul, L, maxw):
  """Case runs (lower-case a/c/g/t keep their 2-bit codes, the run only records the case): perfect
  reads == the oracle's, fused corruption == the numpy spec (a miscalled lower-case base becomes N,
  illumina.py:160)."""
  from tests import philox_ref as PR
  r = H.workload_regions(synth.softmask_workload())[0]
  cv = H.oracle_cv(r['v'][0])
  lay = device_layout(r['ref'], r['region'][1] + 1, cv)
  assert (lay['exc'][:, 2] == 1).sum() > 50                 # many case runs
  rs = np.random.RandomState(8)
  n = 1500
  ts = rs.randint(lay['p_min'], lay['p_max'] - 3 * L, size=n).astype(np.int64)
  tl = rs.randint(L, 3 * L, size=n).astype(np.int64)
  fo = rs.randint(0, 2, size=n).astype(np.int8)
  p1, p2, cnt = run_emul(emul, lay, L, ts, tl, fo, '@S:0:0:', '|s|0', maxw=maxw)
  # every emitted read re-derives from its qname the way the reference would have written it
  idx = H.HaplotypeIndex(r['ref'], r['region'][1] + 1, cv)
  from mitty_b200.simulation.readgenerate import parse_qname
  low = 0
  for which, buf in enumerate((p1, p2)):
    lines = buf.decode().split('\n')
    for k in range(0, len(lines) - 1, 4):
      info = parse_qname(lines[k][1:])[which]
      assert idx.check(info, lines[k + 1]) is None, (lines[k], lines[k + 1])
      low += sum(1 for ch in lines[k + 1] if ch.islower())
  assert low > 10000
  cbm = H.model('hiseq-X-v2.5-Garvan.pkl')['cum_bq_mat']
  ks, c9 = PR.table_shape(cbm, oracle.PHRED_P, 150)
  alias = PR.joint_tables(cbm, oracle.PHRED_P, ks, c9, n_rows=150)
  cor = dict(alias=alias, kshift=ks, code9=c9, k0=99, k1=0x1234)
  c1, c2, ccnt = run_emul(emul, lay, L, ts, tl, fo, '@S:0:0:', '|s|0', corrupt=cor, maxw=maxw)
  assert ccnt == cnt
  assert c1 == PR.corrupt_file(p1, 0, (alias, ks, c9), 99, 0x1234)
  assert c2 == PR.corrupt_file(p2, 1, (alias, ks, c9), 99, 0x1234)
