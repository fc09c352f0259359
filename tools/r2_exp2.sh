#!/bin/bash
# round-2 experiment 2: split backward pieces (gather alone, scatter sort-only / no-flush / full per config)
O=gpurun_out/r04b; mkdir -p $O
timeout 120 python tools/check_bwd.py split32 sorted32 1 2 5 40 > $O/check_bwd_split.txt 2>&1; echo "bwd split rc=$?" >> $O/summary.txt
MSDA_SPLIT_PART=1 timeout 60 python tools/time_variants.py 40 slab32 split32 >> $O/time_parts.txt 2>&1
for cfg in 0 1 2 3 4 5; do
  for probe in 1 2 0; do
    MSDA_SCATTER_CFG=$cfg MSDA_SCATTER_PROBE=$probe MSDA_SPLIT_PART=2 timeout 60 python tools/time_variants.py 40 slab32 split32 >> $O/time_parts.txt 2>&1
  done
done
timeout 60 python tools/time_variants.py 40 slab32 split32 >> $O/time_parts.txt 2>&1
grep -c MISMATCH $O/check_bwd_split.txt; tail -3 $O/check_bwd_split.txt; cat $O/time_parts.txt | sed 's/fwd.*bwd/bwd/'
