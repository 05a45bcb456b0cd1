"""Rank source lines of a kernel by executed instructions from an ncu report (cuda,sass page)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; per = {}
for r in rows:
  if not r: continue
  if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
  if r[0] in ('Line No', 'Function Name'): continue
  if r[0].isdigit() and len(r) > 9:
    try: per[(cur, int(r[0]))] = (int(r[7]), int(r[8]), int(r[6]), r[1])
    except ValueError: pass
tot = sum(v[0] for v in per.values()); ts = sum(v[2] for v in per.values())
print('total warp-inst', tot, 'samples', ts)
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
  print(f'{k[0]:14s}:{k[1]:4d} inst={v[0]/tot:6.1%} thr/inst={v[1]/max(1,v[0]):5.1f} samp={v[2]/max(1,ts):6.1%}  {v[3].strip()[:100]}')
