#!/bin/bash
# round 2: location loads + grad_value zeroing moved ahead of the unit barrier (lib_ro), plus the TMA L2 prefetch of the next unit's
# locations / weights (lib_ro_pf); small-N timings with the host out of the loop (CUDA-graph replay)
O=gpurun_out/r08e; mkdir -p $O
cp combo-avs_b200/libmsda_b200.so /tmp/lib_orig.so
for lib in lib_ro lib_ro_pf lib_ro lib_ro_pf; do
  cp tools/bin/$lib.so combo-avs_b200/libmsda_b200.so
  echo "== $lib" >> $O/bwd.txt
  timeout 200 python tools/check_bwd.py sorted32 warp32 1 5 7 40 2>&1 | grep -E "sorted32\] |MISMATCH|rror" | tail -3 >> $O/bwd.txt
  for n in 20 80; do timeout 100 python tools/time_variants.py $n 2>&1 | tail -1 | cut -c1-120 >> $O/bwd.txt; done
done
echo "== graph replay (lib_ro_pf)" >> $O/bwd.txt
for n in 5 10 20 40; do MSDA_GRAPH=1 timeout 100 python tools/time_variants.py $n 2>&1 | tail -1 | cut -c1-140 >> $O/bwd.txt; done
cp /tmp/lib_orig.so combo-avs_b200/libmsda_b200.so
cat $O/bwd.txt
