#!/bin/bash
# round 2: interleaved unit schedule of the backward (CTA b takes units b, b+grid, ...: the three levels of a (frame, head) run
# concurrently on three SMs, so the L2-resident set shrinks 3x) against the contiguous schedule; DRAM bytes of both
O=gpurun_out/r08b; mkdir -p $O
cp combo-avs_b200/libmsda_b200.so /tmp/lib_orig.so
for lib in lib_new lib_il lib_head lib_new lib_il; do
  cp tools/bin/$lib.so combo-avs_b200/libmsda_b200.so
  echo "== $lib" >> $O/bwd.txt
  timeout 200 python tools/check_bwd.py sorted32 warp32 5 7 40 2>&1 | grep -E "sorted32\] |MISMATCH|rror" | tail -3 >> $O/bwd.txt
  for n in 5 10 20 80; do timeout 100 python tools/time_variants.py $n 2>&1 | tail -1 >> $O/bwd.txt; done
done
for lib in lib_head lib_new lib_il; do
  cp tools/bin/$lib.so combo-avs_b200/libmsda_b200.so
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'sorted32' -s 12 -c 3 --csv --log-file $O/dram_$lib.csv python tools/time_variants.py 40 > /dev/null 2>&1
done
cp /tmp/lib_orig.so combo-avs_b200/libmsda_b200.so
cat $O/bwd.txt
python - <<PY
import csv,glob
for f in sorted(glob.glob("$O/dram_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    print(f)
    for r in rows[1:]:
        print("  ", r[4][:40], r[-3], r[-2], r[-1])
PY
