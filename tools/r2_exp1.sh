#!/bin/bash
# round-2 experiment 1: TMA forward + split backward -- correctness against the round-1 kernels, then timings
mkdir -p gpurun_out/r04a
O=gpurun_out/r04a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
timeout 300 python tools/check_fwd.py slab32 warp32 > $O/check_fwd_tma.txt 2>&1; echo "fwd tma rc=$?" >> $O/summary.txt
MSDA_SLAB_TMA=0 timeout 300 python tools/check_fwd.py slab32 warp32 40 > $O/check_fwd_cpasync.txt 2>&1; echo "fwd cp.async rc=$?" >> $O/summary.txt
timeout 600 python tools/check_bwd.py split32 sorted32 1 2 5 7 40 > $O/check_bwd_split.txt 2>&1; echo "bwd split rc=$?" >> $O/summary.txt
for cfg in 0 1 2 3 4 5; do
  MSDA_SCATTER_CFG=$cfg MSDA_SPLIT_PART=2 timeout 120 python tools/time_variants.py 40 slab32 split32 >> $O/time_scatter_cfg.txt 2>&1
done
MSDA_SPLIT_PART=1 timeout 120 python tools/time_variants.py 40 slab32 split32 >> $O/time_gather.txt 2>&1
for cfg in 0 2 4; do
  MSDA_SCATTER_CFG=$cfg timeout 120 python tools/time_variants.py 40 slab32 split32 >> $O/time_split.txt 2>&1
done
timeout 120 python tools/time_variants.py 40 slab32 sorted32 >> $O/time_split.txt 2>&1
for n in 5 10 20 80; do MSDA_SCATTER_CFG=4 timeout 120 python tools/time_variants.py $n slab32 split32 >> $O/time_split.txt 2>&1; done
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/summary.txt
tail -n 5 $O/*.txt
