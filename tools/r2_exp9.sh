#!/bin/bash
O=gpurun_out/r05b; mkdir -p $O
timeout 120 python tools/check_linear.py > $O/linear_3xtf32.txt 2>&1; echo "check_linear rc=$?" >> $O/summary.txt
timeout 600 python -m pytest tests -x -q -m gpu -k "module or encoder or pixel_decoder or fused or graph or harness" > $O/pytest_sel.txt 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 300 python bench.py --no-cpu --no-incumbent --steps 50 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; cat $O/linear_3xtf32.txt; tail -6 $O/pytest_sel.txt
python - <<PY
import json
d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
m=d["model_step"]; print("pixel decoder", m["weak"], "full", m["full_model"].get("weak"), m["full_model"].get("unavailable"))
PY
