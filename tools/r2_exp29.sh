#!/bin/bash
# round 2: cost-balanced warp ranges in the walk (kSwitchCost 3 / 5 / 7 half samples per cell switch) against equal counts (lib_final)
O=gpurun_out/r08q; mkdir -p $O
cp combo-avs_b200/libmsda_b200.so /tmp/lib_orig.so
for lib in lib_final lib_bal3 lib_bal5 lib_bal7 lib_final lib_bal5; do
  cp tools/bin/$lib.so combo-avs_b200/libmsda_b200.so
  echo "== $lib" >> $O/bwd.txt
  timeout 200 python tools/check_bwd.py sorted32 warp32 1 5 7 40 2>&1 | grep -E "sorted32\] |MISMATCH|rror" | tail -3 >> $O/bwd.txt
  for n in 5 20 80; do MSDA_GRAPH=1 timeout 100 python tools/time_variants.py $n 2>&1 | tail -1 | cut -c1-120 >> $O/bwd.txt; done
done
cp tools/bin/lib_stamp.so combo-avs_b200/libmsda_b200.so
MSDA_SORTED_TIMES=$O/walk_stamps.txt timeout 200 python tools/check_bwd.py sorted32 warp32 40 2>&1 | grep -E "sorted32\] |MISMATCH" | tail -2
cp /tmp/lib_orig.so combo-avs_b200/libmsda_b200.so
cat $O/bwd.txt
python - <<PY
import numpy as np
a=np.loadtxt("$O/walk_stamps.txt",dtype=np.int64)
t=a[:,2:]; ok=t[:,31]>0; t=t[ok]
start=t[:,31:32]; ends=t[:,:24]-start
rel=ends/ends.max(1,keepdims=True)
print("stamps (balanced, cost 5): mean walk us %.2f; per-warp mean end / slowest:"%(ends.max(1).mean()/1e3), np.round(rel.mean(0),3))
print("mean (slowest - mean end) / slowest: %.3f" % (1-rel.mean())); print("slowest warp histogram:", np.bincount(ends.argmax(1),minlength=24))
PY
