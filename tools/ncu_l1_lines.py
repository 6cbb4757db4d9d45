"""Which instructions of a kernel occupy the L1 / shared-memory data pipe (ncu source page):
   python tools/ncu_l1_lines.py report.ncu-rep <kernel name substring> [top N]
The pipe moves one 128-byte wavefront per cycle and is shared by the tensor core's operand fetch, LSU shared and global accesses."""
import csv, io, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
tables, cur = [], None
for row in csv.reader(io.StringIO(raw)):
  if row and row[0] == 'Kernel Name':
    cur = {'name': row[1], 'rows': []}; tables.append(cur)
  elif cur is not None:
    cur['rows'].append(row)
f = lambda x: float(x or 0)
for t in tables:
  if want not in t['name']:
    continue
  h, rows = t['rows'][0], [x for x in t['rows'][1:] if len(x) > 10]
  c = h.index
  iSrc, iE, iW, iWi, iT, iS = c('Source'), c('Instructions Executed'), c('L1 Wavefronts Shared'), c('L1 Wavefronts Shared Ideal'), c('L1 Tag Requests Global'), c('# Samples')
  iG = c('L2 Theoretical Sectors Global')
  print('==', t['name'][:90])
  print('shared wavefronts %.3g (ideal %.3g)   global tag requests %.3g   global sectors %.3g' % (sum(f(x[iW]) for x in rows), sum(f(x[iWi]) for x in rows), sum(f(x[iT]) for x in rows), sum(f(x[iG]) for x in rows)))
  print('-- shared: wavefronts, ideal, executed, instruction')
  for x in sorted(rows, key=lambda x: -f(x[iW]))[:top]:
    print('%10d %10d %9s  %s' % (f(x[iW]), f(x[iWi]), x[iE], x[iSrc][:90]))
  print('-- global: tag requests, sectors, executed, instruction')
  for x in sorted(rows, key=lambda x: -f(x[iT]))[:top // 2]:
    print('%10d %10d %9s  %s' % (f(x[iT]), f(x[iG]), x[iE], x[iSrc][:90]))
  break
