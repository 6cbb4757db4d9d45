#!/bin/bash
# ONE ncu session of the current build on a B200 (run through gpurun): every launch of one a4 reference tile, a3's fused PixelShuffle(3)
# convolution + head dot products and its stencil kernel, NetDN's fused residual block (K = 48).  The .ncu-rep files (60+ MB) stay on the box; what comes
# back in gpurun_out/ is the summary, DRAM bytes per launch, the raw metric tables and the L1 data-pipe view of the two top kernels.
set -u
R=/tmp/ncu; mkdir -p $R gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 700 $NCU -o $R/a4_tile_all_kernels python tools/prof_tile.py a4 4 2160 968 1 > gpurun_out/ncu_a4.log 2>&1
timeout 400 $NCU -k regex:"pair_head|stencil9" -c 3 -o $R/a3_tile_up_head python tools/prof_tile.py a3 3 2160 1290 1 > gpurun_out/ncu_a3.log 2>&1
timeout 300 $NCU -k regex:arsb --launch-skip 2 -c 1 -o $R/dn_tile_arsb python tools/prof_tile.py dn_lite15 1 1080 1920 1 > gpurun_out/ncu_dn.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_kernels_ncu.txt gpurun_out/traffic.json $R/a4_tile_all_kernels.ncu-rep $R/a3_tile_up_head.ncu-rep $R/dn_tile_arsb.ncu-rep > /dev/null
for n in a4_tile_all_kernels a3_tile_up_head dn_tile_arsb; do ncu -i $R/$n.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/r02_${n}_ncu_raw.csv.gz; done
{ python tools/ncu_l1_lines.py $R/a4_tile_all_kernels.ncu-rep pair_head 24; python tools/ncu_l1_lines.py $R/a4_tile_all_kernels.ncu-rep arsb_pair 24; python tools/ncu_l1_lines.py $R/a4_tile_all_kernels.ncu-rep conv3x3_pair_kernel 16; } > gpurun_out/r02_l1_data_pipe_lines.txt 2>&1
ls -la $R gpurun_out | tail -20
