"""Executed warp instructions of one kernel of an ncu report grouped by the enclosing function of
each source line.  Usage: python tools/ncu_funcs.py report.ncu-rep kernel_regex"""
import csv, re, subprocess, sys, os
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--kernel-name', 'regex:' + kern, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; per = {}
for r in rows:
  if not r: continue
  if r[0] == 'File Path': cur = r[1]; continue
  if r[0].isdigit() and len(r) > 9:
    try: per[(cur, int(r[0]))] = (int(r[7]), int(r[8]), int(r[6]))
    except ValueError: pass
fn_of = {}
for path in {k[0] for k in per}:
  local = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mitty_b200', 'csrc', os.path.basename(path))
  if not os.path.exists(local): continue
  name = '?'; names = []
  for ln in open(local):
    m = re.match(r'^(?:template.*>\s*)?(?:MG_HD|MG_NI|static|__global__|__device__|inline)[^;(]*?\b(\w+)\s*\(', ln)
    if m and not ln.startswith(' '): name = m.group(1)
    m2 = re.match(r'^\s*(?:MG_HD|MG_NI)\s+[^;(]*?\b(\w+)\s*\(', ln)
    if m2: name = m2.group(1)
    names.append(name)
  fn_of[path] = names
tot = sum(v[0] for v in per.values()); agg = {}
for (path, ln), v in per.items():
  f = fn_of.get(path, ['?'] * 100000)[ln - 1] if path in fn_of else os.path.basename(path)
  a = agg.setdefault(f, [0, 0, 0]); a[0] += v[0]; a[1] += v[1]; a[2] += v[2]
ts = sum(a[2] for a in agg.values())
print('total warp-inst', tot)
for f, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
  print(f'{f:28s} inst={a[0]/tot:6.1%} thr/inst={a[1]/max(1,a[0]):5.1f} samples={a[2]/max(1,ts):6.1%}')
