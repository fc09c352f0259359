#!/bin/bash
O=gpurun_out/r04f; mkdir -p $O
timeout 1700 python -m pytest tests -x -q -m gpu -k "reference_cuda_extension or avss_batch or ops_test_runs or bench_prints" > $O/pytest_new.txt 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 600 python bench.py > $O/bench_fp32.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -5 $O/pytest_new.txt; python - <<PY
import json
d=json.loads(open("$O/bench_fp32.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["fwd"]["us"], d["roofline"]["bwd"]["us"], d["roofline"]["frac"])
print("e2e", d["e2e"]["value"], "incumbent", d["incumbent"])
print("model_step", d["model_step"])
print("clocks", d["clocks"])
PY
tail -5 $O/bench.err
