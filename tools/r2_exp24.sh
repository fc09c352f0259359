#!/bin/bash
# round 2: TMA L2 prefetch of the next unit's inputs during the walk (MSDA_PF bits: 1 locations + weights, 2 value level, 4 grad_out rows)
O=gpurun_out/r08d; mkdir -p $O
cp combo-avs_b200/libmsda_b200.so /tmp/lib_orig.so
for lib in lib_il lib_pf0 lib_pf1 lib_pf2 lib_pf3 lib_pf4 lib_pf7 lib_pf0 lib_pf7; do
  cp tools/bin/$lib.so combo-avs_b200/libmsda_b200.so
  echo "== $lib" >> $O/bwd.txt
  timeout 200 python tools/check_bwd.py sorted32 warp32 5 7 40 2>&1 | grep -E "sorted32\] |MISMATCH|rror" | tail -3 >> $O/bwd.txt
  for n in 20 80; do timeout 100 python tools/time_variants.py $n 2>&1 | tail -1 | cut -c1-120 >> $O/bwd.txt; done
done
cp /tmp/lib_orig.so combo-avs_b200/libmsda_b200.so
cat $O/bwd.txt
