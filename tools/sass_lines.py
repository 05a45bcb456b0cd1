"""Static SASS instruction count per source line for one kernel of libmitty_b200.so."""
import collections, glob, os, re, subprocess, sys, tempfile
pat = sys.argv[1] if len(sys.argv) > 1 else 'k_unit_emitILi12ELb0'
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(root, 'mitty_b200', 'libmitty_b200.so')], cwd=d, capture_output=True)
for f in glob.glob(os.path.join(d, '*.cubin')):
  txt = subprocess.run(['nvdisasm', '--print-line-info', f], capture_output=True, text=True).stdout
  if pat not in txt: continue
  for p in re.split(r'\n\s*\.section\s+\.text\.', txt):
    if pat not in p.split('\n')[0]: continue
    cur = None; cnt = collections.Counter()
    for line in p.split('\n'):
      m = re.search(r'//## File "([^"]+)", line (\d+)', line)
      if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
      if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line): cnt[cur] += 1
    tot = sum(cnt.values()); print('total', tot)
    src = {}
    for (fn, ln), c in cnt.most_common(top):
      text = ''
      path = os.path.join(root, 'mitty_b200', 'csrc', fn)
      if os.path.exists(path):
        src.setdefault(path, open(path).read().split('\n')); text = src[path][ln - 1].strip()[:95]
      print(f'{c:5d} {c/tot:5.1%} {fn}:{ln}  {text}')
