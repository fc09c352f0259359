#!/bin/bash
# round-2 experiment 5: sorted32 with in-kernel zeroing of grad_value; forward pps sweep; full gpu test-suite
O=gpurun_out/r04e; mkdir -p $O
timeout 100 python tools/check_bwd.py sorted32 warp32 1 2 5 7 40 > $O/check_bwd.txt 2>&1; echo "sorted32 vs warp32 rc=$?" >> $O/summary.txt
timeout 60 python tools/time_variants.py 40 slab32 sorted32 >> $O/time.txt 2>&1
for pps in 5 7 8; do MSDA_SLAB_PPS=$pps timeout 60 python tools/time_variants.py 40 slab32 sorted32 >> $O/time.txt 2>&1; done
for n in 5 10 20 80; do timeout 60 python tools/time_variants.py $n slab32 sorted32 >> $O/time.txt 2>&1; done
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/summary.txt
grep -c MISMATCH $O/check_bwd.txt; cat $O/summary.txt; tail -4 $O/pytest_gpu.txt; cat $O/time.txt
