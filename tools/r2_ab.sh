#!/bin/bash
# same-box A/B of prebuilt libraries under tools/bin/lib_*.so: usage r2_ab.sh <outdir> <lib> [ENV=VAL ...] -- <lib> [ENV=VAL ...] -- ...
O=$1; shift; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $O/gpu.txt
cp combo-avs_b200/libmsda_b200.so /tmp/lib_orig.so
args=("$@"); cur=()
run_one() {
  lib=$1; shift
  cp tools/bin/$lib combo-avs_b200/libmsda_b200.so
  echo "== $lib $*" >> $O/log.txt
  env "$@" MSDA_AB=1 timeout 200 python tools/check_bwd.py sorted32 warp32 40 2>&1 | grep -E "sorted32\] |MISMATCH|rror" | tail -2 >> $O/log.txt
}
for a in "${args[@]}" --; do
  if [ "$a" == "--" ]; then [ ${#cur[@]} -gt 0 ] && run_one "${cur[@]}"; cur=(); else cur+=("$a"); fi
done
cp /tmp/lib_orig.so combo-avs_b200/libmsda_b200.so
cat $O/gpu.txt $O/log.txt
