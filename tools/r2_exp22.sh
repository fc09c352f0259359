#!/bin/bash
# round 2, session 3: clean TMA backward vs HEAD, L2 eviction-policy hints (backward: grad_out evict_last = 1, slab evict_first = 2; forward
# slab evict_first = 4), forward passes-per-slab sweep, DRAM bytes of each under ncu (two metrics only)
O=gpurun_out/r08a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $O/gpu.txt
cp combo-avs_b200/libmsda_b200.so /tmp/lib_orig.so
for lib in lib_head lib_tma lib_h1 lib_h2 lib_h3 lib_tma lib_head; do
  cp tools/bin/$lib.so combo-avs_b200/libmsda_b200.so
  echo "== $lib" >> $O/bwd.txt
  timeout 200 python tools/check_bwd.py sorted32 warp32 5 40 2>&1 | grep -E "sorted32\] |MISMATCH|rror" | tail -3 >> $O/bwd.txt
done
for lib in lib_tma lib_h4 lib_tma lib_h4; do
  cp tools/bin/$lib.so combo-avs_b200/libmsda_b200.so
  echo "== $lib" >> $O/fwd.txt
  for n in 40 80; do timeout 100 python tools/time_variants.py $n 2>&1 | tail -1 >> $O/fwd.txt; done
done
cp tools/bin/lib_tma.so combo-avs_b200/libmsda_b200.so
for pps in 4 5 7 8 9 12; do
  echo "== pps $pps" >> $O/fwd.txt
  MSDA_SLAB_PPS=$pps timeout 100 python tools/time_variants.py 40 2>&1 | tail -1 >> $O/fwd.txt
done
for lib in lib_head lib_tma lib_h3 lib_h4; do
  cp tools/bin/$lib.so combo-avs_b200/libmsda_b200.so
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'slab32|sorted32' -s 6 -c 4 --csv --log-file $O/dram_$lib.csv python tools/time_variants.py 40 > /dev/null 2>&1
done
cp /tmp/lib_orig.so combo-avs_b200/libmsda_b200.so
cat $O/bwd.txt $O/fwd.txt
python - <<PY
import csv,glob
for f in sorted(glob.glob("$O/dram_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    print(f)
    for r in rows[1:]:
        print("  ", r[4][:50], r[-3], r[-2], r[-1])
PY
