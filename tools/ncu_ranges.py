"""Executed warp instructions and stall samples of an ncu report, summed over source-line ranges of mg_core.cuh / mg_kernels.cu
(a coarse 'which part of the kernel' breakdown).  Usage: ncu_ranges.py report.ncu-rep"""
import csv, re, subprocess, sys, os
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; per = {}
for r in rows:
  if not r: continue
  if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
  if r[0] in ('Line No', 'Function Name'): continue
  if r[0].isdigit() and len(r) > 9:
    try: per[(cur, int(r[0]))] = (int(r[7]), int(r[8]), int(r[6]))
    except ValueError: pass
# ranges from function starts in the sources
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mitty_b200', 'csrc')
marks = {}
for fn in ('mg_core.cuh', 'mg_kernels.cu'):
  m = []
  for i, line in enumerate(open(os.path.join(root, fn)), 1):
    g = re.match(r'^(?:MG_HD|MG_NI|template|static|__global__|__device__|struct)\b.*?(\w+)\s*(?:\(|\{|$)', line)
    if re.match(r'^(MG_HD|MG_NI|__global__|__device__|static)\b', line) or re.match(r'^struct \w+', line):
      name = re.findall(r'(\w+)\s*\(', line)
      m.append((i, (name[0] if name else line.split()[1]).strip()))
  marks[fn] = m
agg = {}
for (fn, ln), (inst, thr, samp) in per.items():
  name = '?'
  for s, n in marks.get(fn, []):
    if s <= ln: name = n
    else: break
  a = agg.setdefault((fn, name), [0, 0, 0]); a[0] += inst; a[1] += thr; a[2] += samp
tot = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
  print(f'{k[0]:14s} {k[1]:28s} inst={v[0]/tot:6.1%} thr/inst={v[1]/max(1,v[0]):5.1f} samp={v[2]/max(1,ts):6.1%}')
