#!/bin/bash
# round-2 evidence session (one gpurun call): full GPU test-suite, smoke, bench (fp32 / reference / bf16), launch list, ncu captures, N sweep
TAG=${1:-r06a}; O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu.txt 2>&1
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/smoke.log; tail -2 $O/smoke.log
echo "== bench"; timeout 600 python bench.py > $O/bench_fp32.json 2> $O/bench.err; echo "bench exit $?"; cut -c1-200 $O/bench_fp32.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
timeout 300 python bench.py --dtype bf16 --no-cpu --no-model-step > $O/bench_bf16.json 2>> $O/bench.err
echo "== sweep"; for n in 5 10 20 40 80; do timeout 60 python tools/time_variants.py $n >> $O/sweep_N.txt 2>&1; done
echo "# the same from a CUDA graph (host out of the loop: at small N the eager loop is bound by the Python call, not by the kernels)" >> $O/sweep_N.txt
for n in 5 10 20 40 80; do MSDA_GRAPH=1 timeout 60 python tools/time_variants.py $n >> $O/sweep_N.txt 2>&1; done; cat $O/sweep_N.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-incumbent --no-model-step > $O/ncu_launches_bench.log 2>&1
echo "== ncu full (forward + backward defaults)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'slab32|sorted32' -s 4 -c 2 -o $O/prof_main python bench.py --steps 2 --warmup 3 --no-cpu --no-incumbent --no-model-step > $O/ncu_full.log 2>&1
echo "== ncu full (linear)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'linear_tf32' -s 60 -c 3 -o $O/prof_linear python tools/check_linear.py > $O/ncu_linear.log 2>&1
ls -la $O
