#!/bin/bash
# One gpurun call: tests, smoke, bench (fp32 / reference / bf16), incumbent comparison, microbenchmarks, ncu launch list + full captures.
# Usage (from the repo root on the GPU box): bash tools/gpu_session.sh [tag]
TAG=${1:-r01z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log; tail -4 $OUT/smoke.log
echo "== microbench"; timeout 300 tools/bin/microbench > $OUT/microbench.txt 2>&1
echo "== incumbent"; for n in 40 20 10 5 80; do timeout 600 python tools/compare_incumbent.py $n > $OUT/incumbent_N$n.json 2>> $OUT/incumbent.err; done
python - <<PY
import json
for n in (5, 10, 20, 40, 80):
    d = json.load(open("$OUT/incumbent_N%d.json" % n))
    print(n, {k: round(v["us"], 1) for k, v in d["ours"].items()}, "ref", round(d["reference_cuda"]["fwd_us"], 1), round(d["reference_cuda"]["bwd_us"], 1))
PY
echo "== bench"; timeout 600 python bench.py > $OUT/bench_fp32.json 2> $OUT/bench.err; cut -c1-300 $OUT/bench_fp32.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
timeout 300 python bench.py --dtype bf16 --no-cpu > $OUT/bench_bf16.json 2>> $OUT/bench.err
timeout 300 python bench.py --dtype bf16 --loc-dtype bf16 --no-cpu > $OUT/bench_bf16_bf16loc.json 2>> $OUT/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_launches_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'slab32|sorted32' -s 4 -c 2 -o $OUT/prof_main python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
echo "== ncu full, fused prologue kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'slab32|sorted32|softmax_backward|zero_masked' -s 40 -c 4 -o $OUT/prof_fused python tools/module_bench.py 40 > $OUT/ncu_fused.log 2>&1
# gpurun brings back at most 64 MiB: summarise the reports here and keep only the text
for r in prof_main prof_fused; do
  if [ -f $OUT/$r.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/$r.ncu-rep > $OUT/${r}_summary.txt 2>&1
    ncu -i $OUT/$r.ncu-rep --page source --csv > $OUT/${r}_source.csv 2>/dev/null
    rm -f $OUT/$r.ncu-rep
  fi
done
ls -la $OUT
