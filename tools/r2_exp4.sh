#!/bin/bash
O=gpurun_out/r04d; mkdir -p $O
MSDA_SCATTER_CFG=5 MSDA_SORTED_NOSLAB=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'slab32|scatter32|sorted32' -s 4 -c 4 -o $O/prof python tools/profile_split.py 40 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls -la $O
