#!/bin/bash
# round 2, f2 epilogue: fused add + LayerNorm and the column-sum kernels: tests, then the pixel-decoder step under torch.profiler, then model_step
O=gpurun_out/r08j; mkdir -p $O
timeout 600 python -m pytest tests/test_layernorm.py tests/test_linear.py -m gpu -x -q > $O/pytest_ln.log 2>&1; echo "pytest ln exit $?"; tail -3 $O/pytest_ln.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_harness.py -m gpu -x -q -k "encoder or decoder or module or harness or graph" > $O/pytest_callers.log 2>&1; echo "pytest callers exit $?"; tail -3 $O/pytest_callers.log
timeout 300 python tools/decoder_step_torchprof.py > $O/torchprof.txt 2>&1; grep -E "Self CUDA time total" $O/torchprof.txt
MSDA_FUSED_ADD_NORM=0 timeout 300 python tools/decoder_step_torchprof.py > $O/torchprof_unfused.txt 2>&1; grep -E "Self CUDA time total" $O/torchprof_unfused.txt
timeout 600 python bench.py --no-cpu --no-incumbent > $O/bench_fp32.json 2> $O/bench.err; python - <<PY
import json
d=json.load(open("$O/bench_fp32.json")); m=d["model_step"]; print(d["value"], m["weak"], m["full_model"]["weak"])
PY
