#!/bin/bash
# round 2: counting pass + per-cell table ahead of the unit barrier (two counter / table sets): lib_db against lib_ro; then the GPU test-suite
O=gpurun_out/r08f; mkdir -p $O
cp combo-avs_b200/libmsda_b200.so /tmp/lib_orig.so
for lib in lib_ro lib_db lib_ro lib_db; do
  cp tools/bin/$lib.so combo-avs_b200/libmsda_b200.so
  echo "== $lib" >> $O/bwd.txt
  timeout 200 python tools/check_bwd.py sorted32 warp32 1 5 7 40 2>&1 | grep -E "sorted32\] |MISMATCH|rror" | tail -3 >> $O/bwd.txt
  for n in 5 10 20 80; do timeout 100 python tools/time_variants.py $n 2>&1 | tail -1 | cut -c1-120 >> $O/bwd.txt; done
done
cp /tmp/lib_orig.so combo-avs_b200/libmsda_b200.so
cat $O/bwd.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_gpu.log
