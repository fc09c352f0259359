#!/bin/bash
O=gpurun_out/r04h; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?" >> $O/summary.txt
cat $O/summary.txt
python - <<PY
import json
d=json.loads(open("$O/bench_2gpu.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, "e2e", d["e2e"]["value"])
print("model_step", json.dumps(d["model_step"])[:2500])
PY
tail -5 $O/bench_2gpu.err
