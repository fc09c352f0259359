#!/bin/bash
# round-2 experiment 3: slab-less two-CTA sorted32, gather with prefetched grad_out rows
O=gpurun_out/r04c; mkdir -p $O
timeout 120 python tools/check_bwd.py sorted32 split32 1 2 5 40 > $O/check_bwd_noslab.txt 2>&1; echo "noslab vs split rc=$?" >> $O/summary.txt
MSDA_SORTED_NOSLAB=0 timeout 60 python tools/time_variants.py 40 slab32 sorted32 >> $O/time.txt 2>&1
MSDA_SORTED_NOSLAB=1 timeout 60 python tools/time_variants.py 40 slab32 sorted32 >> $O/time.txt 2>&1
MSDA_SPLIT_PART=1 timeout 60 python tools/time_variants.py 40 slab32 split32 >> $O/time.txt 2>&1
for n in 5 10 20 80; do MSDA_SORTED_NOSLAB=1 timeout 60 python tools/time_variants.py $n slab32 sorted32 >> $O/time.txt 2>&1; done
for n in 5 10; do MSDA_SORTED_NOSLAB=0 timeout 60 python tools/time_variants.py $n slab32 sorted32 >> $O/time.txt 2>&1; done
grep -c MISMATCH $O/check_bwd_noslab.txt; tail -3 $O/check_bwd_noslab.txt; cat $O/time.txt
