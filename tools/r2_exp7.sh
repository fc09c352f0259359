#!/bin/bash
O=gpurun_out/r04g; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu -k "harness or ops_test_runs" > $O/pytest_new.txt 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 300 python tools/combo_bench.py train 8 6 > $O/combo_train_1gpu.json 2> $O/combo.err; echo "combo rc=$?" >> $O/summary.txt
timeout 300 python tools/combo_bench.py cpu-forward 1 > $O/combo_cpu_forward.json 2>> $O/combo.err
timeout 600 python bench.py --no-cpu > $O/bench_fp32.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -5 $O/pytest_new.txt; cat $O/combo_train_1gpu.json $O/combo_cpu_forward.json; tail -3 $O/combo.err
python - <<PY
import json
d=json.loads(open("$O/bench_fp32.json").read().strip().splitlines()[-1])
print("model_step", json.dumps(d["model_step"])[:1500])
PY
tail -3 $O/bench.err
