#!/bin/bash
# round 2: per-warp end-of-walk time stamps (lib_stamp = -DMSDA_WALK_STAMPS): is the skew that the unit barrier waits for systematic per warp?
O=gpurun_out/r08o; mkdir -p $O
cp combo-avs_b200/libmsda_b200.so /tmp/lib_orig.so
cp tools/bin/lib_stamp.so combo-avs_b200/libmsda_b200.so
MSDA_SORTED_TIMES=$O/walk_stamps.txt timeout 200 python tools/check_bwd.py sorted32 warp32 40 2>&1 | grep -E "sorted32\] |MISMATCH" | tail -2
cp /tmp/lib_orig.so combo-avs_b200/libmsda_b200.so
python - <<PY
import numpy as np
a=np.loadtxt("$O/walk_stamps.txt",dtype=np.int64)
t=a[:,2:]
ok=t[:,31]>0
t=t[ok]; u=a[ok,1]
start=t[:,31:32]; ends=t[:,:24]-start
print("units",len(t),"mean walk us",ends.max(1).mean()/1e3)
rel=ends/ends.max(1,keepdims=True)
print("per-warp mean end / slowest:", np.round(rel.mean(0),3))
print("mean (slowest - mean end) / slowest: %.3f" % (1-rel.mean()))
print("which warp is slowest (histogram):", np.bincount(ends.argmax(1),minlength=24))
PY
