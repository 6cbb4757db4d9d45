"""SASS evidence per kernel of the built library (run where cuobjdump is installed; no GPU needed):
  python tools/sass_summary.py > profiles/r02_sass_summary.txt
tcgen05.mma = UTCHMMA, tcgen05.ld / .st = LDTM / STTM, TMA = UTMALDG / UTMASTG / UBLKCP, tcgen05.commit = UTCBAR (B200_PROFILING.md)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'moephoto_b200', 'lib', 'libmoephoto_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
MN = ['UTCHMMA.2CTA', 'UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'HMMA', 'LDGSTS', 'ATOMG', 'MEMBAR', 'BAR.SYNC']
name, counts, total = None, {}, collections.Counter()
for line in sass.splitlines():
  m = re.search(r'Function : (\S+)', line)
  if m:
    name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
    counts[name] = collections.Counter()
    continue
  if name is None:
    continue
  for mn in MN:
    if re.search(r'\b' + re.escape(mn) + r'\b', line) or (mn + '.') in line and mn in ('LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'ATOMG', 'MEMBAR', 'HMMA', 'LDGSTS'):
      if mn == 'UTCHMMA' and 'UTCHMMA.2CTA' in line:
        continue
      counts[name][mn] += 1
      total[mn] += 1
      break
  if re.match(r'\s+/\*[0-9a-f]{4}\*/', line):
    counts[name]['instructions'] += 1
print('cuobjdump -sass %s  (sm_100a; instruction counts per kernel)' % os.path.relpath(lib, ROOT))
print('%-58s %7s %s' % ('kernel', 'instr', '  '.join('%s' % m for m in MN)))
for k, c in sorted(counts.items()):
  print('%-58s %7d %s' % (k[:58], c['instructions'], '  '.join('%*d' % (len(m), c[m]) for m in MN)))
print('%-58s %7s %s' % ('TOTAL', '', '  '.join('%*d' % (len(m), total[m]) for m in MN)))
print('HMMA (legacy mma.sync path) = %d: the tensor work is tcgen05 only' % total['HMMA'])
