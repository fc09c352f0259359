#!/bin/bash
# One gpurun call: tests, smoke, bench, incumbent comparison, microbenchmarks, ncu launch list + full captures.
# Usage (from the repo root on the GPU box): bash tools/gpu_session.sh [tag]
TAG=${1:-r01a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log; tail -12 $OUT/smoke.log
echo "== microbench"; timeout 300 tools/bin/microbench > $OUT/microbench.txt 2>&1; cat $OUT/microbench.txt
echo "== incumbent"; timeout 600 python tools/compare_incumbent.py 40 > $OUT/incumbent_N40.json 2> $OUT/incumbent.err; cat $OUT/incumbent_N40.json; tail -3 $OUT/incumbent.err
timeout 300 python tools/compare_incumbent.py 5 > $OUT/incumbent_N5.json 2>> $OUT/incumbent.err
echo "== bench"; timeout 600 python bench.py > $OUT/bench_fp32.json 2> $OUT/bench.err; cat $OUT/bench_fp32.json; tail -3 $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; cat $OUT/bench_reference.json
timeout 300 python bench.py --dtype bf16 --no-cpu > $OUT/bench_bf16.json 2>> $OUT/bench.err; cat $OUT/bench_bf16.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_launches_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'warp32|slab|sorted' -s 4 -c 4 -o $OUT/prof_main python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
ls -la $OUT
