"""Summarise `ncu --set full` reports (run where ncu is installed; no GPU needed):
  python tools/ncu_summary.py out.txt traffic.json rep1.ncu-rep [rep2.ncu-rep ...]
One line per captured launch (duration, cycles, tensor pipe, DRAM bytes and throughput, registers) and, in traffic.json,
dram__bytes_read.sum + dram__bytes_write.sum per launch keyed by the engine profiler's kernel class (bench.py `roofline.traffic`)."""
import csv, io, json, subprocess, sys

CLASS_OF = [('head_stencil9_kernel', 'head9'), ('arsb_pair_kernel', 'arsb'), ('conv3x3_pair_head_kernel', 'conv_up_head'), ('conv3x3_pair_trunk_kernel', 'conv_trunk'),
            ('conv3x3_pair_kernel', 'conv_up'), ('conv_first_kernel', 'conv_input'), ('head_stencil_kernel', 'head'), ('head_tc_kernel', 'head_tc')]
WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'sm__cycles_elapsed.max.per_second',
        # the L1 / shared-memory data pipe (one 128-byte wavefront per cycle): tensor-core operand fetch and every LSU access share it
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']


def unit_scale(u):
  return {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1.0}.get(u, 1.0)


def main():
  out_txt, out_json, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
  lines, traffic = [], {}
  for rep in reps:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    lines.append('== %s' % rep)
    for r in rows[2:]:
      name = r[ki].replace('(bool)', '').replace('(int)', '').replace('(moe::ConvEpilogue)', '').split('(')[0].replace('moe::', '').replace('void ', '')
      v = {}
      for k in WANT:
        cand = [i for i, h in enumerate(hdr) if h == k or h.endswith('.' + k)]
        if cand:
          try:
            v[k] = float(r[cand[0]].replace(',', '')) * (unit_scale(units[cand[0]]) if 'bytes' in k or 'duration' in k else 1.0)
          except ValueError:
            pass
      rd, wr, dur = v.get('dram__bytes_read.sum', 0), v.get('dram__bytes_write.sum', 0), v.get('gpu__time_duration.sum', 0)
      lines.append('%-40s %9.1f us  %8.0f kcycles @ %.2f GHz | tensor pipe %5.1f %% of elapsed, operand path %5.1f %%, L1 data pipe: tensor %4.1f %% + LSU %4.1f %% | DRAM read %7.3f GB write %7.3f GB = %5.2f TB/s (%4.1f %% of peak) | %3d regs, grid %d'
                   % (name, dur * 1e6, v.get('sm__cycles_elapsed.max', 0) / 1e3, v.get('sm__cycles_elapsed.max.per_second', 0),
                      v.get('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 0),
                      v.get('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0),
                      v.get('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 0), v.get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 0), rd / 1e9, wr / 1e9, (rd + wr) / max(dur, 1e-12) / 1e12,
                      v.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0), int(v.get('launch__registers_per_thread', 0)), int(v.get('launch__grid_size', 0))))
      for key, cls in CLASS_OF:
        if rep != reps[0]:
          break                                                     # traffic.json describes the FIRST report (the bench workload's tile)
        if name.startswith(key):
          traffic.setdefault(cls, []).append(rd + wr)
          break
  open(out_txt, 'w').write('\n'.join(lines) + '\n')
  json.dump({'dram_bytes_per_launch': {k: sum(v) / len(v) for k, v in traffic.items()},
             'launches_averaged': {k: len(v) for k, v in traffic.items()},
             'source': 'ncu --set full --clock-control none (dram__bytes_read.sum + dram__bytes_write.sum), ONE session of the build that produced the '
                       'bench line: every launch of one a4 reference tile (3 x 2160 x 968 LR px = what the 4K bench frame runs four times); ' + ', '.join(reps)},
            open(out_json, 'w'), indent=1)
  print('\n'.join(lines))


if __name__ == '__main__':
  main()
