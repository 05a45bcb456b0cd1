"""Top source lines per warp-stall reason from an ncu report (cuda,sass page)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; data = []
for r in rows:
  if not r: continue
  if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
  if r[0] == 'Line No': hdr = r; continue
  if r[0] == 'Function Name': continue
  if r[0].isdigit() and hdr and len(r) == len(hdr): data.append((cur, r))
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') or 'Stall' in h]
names = [hdr[i] for i in cols]
print(names)
for i in cols:
  name = hdr[i]
  if name in ('Warp Stall Sampling (Not-issued Samples)',): continue
  vals = []
  for cur, r in data:
    try: v = int(r[i])
    except ValueError: v = 0
    if v: vals.append((v, cur, r[0], r[1]))
  tot = sum(v[0] for v in vals)
  if tot < 200: continue
  print('\n== %s total %d' % (name, tot))
  for v, cur, ln, src in sorted(vals, reverse=True)[:top]:
    print('  %5.1f%%  %s:%s  %s' % (100.0 * v / tot, cur, ln, src.strip()[:95]))
