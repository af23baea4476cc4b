#!/bin/bash
# diagnostic: where does the 8-GPU sharded step go?  (gpurun --gpus 8 -- 'bash tools/r02_multi_trace.sh')
O=gpurun_out; mkdir -p $O
run() {  # name, extra env...
  name=$1; shift
  env MATINV_MULTI_TRACE=1 "$@" python - > $O/r02_trace_$name.txt 2>&1 <<PY
import sys, time
sys.path.insert(0, ".")
import gpu_matrix_inversion_b200 as m
from oracle.gj_oracle import SEED_UNIFORM
n = 65536
for i in range(2):
    rc, piv, ms = m.sharded_synthetic(n, SEED_UNIFORM + n, "uniform", ngpu=8)
    print("run", i, "rc", rc, "compute_ms (factor + exchange)", ms, flush=True)
PY
  echo "== $name"; grep -E "compute_ms|trace" $O/r02_trace_$name.txt | tail -10
}
run default
run nch2 NCCL_MAX_NCHANNELS=2
run nch16 NCCL_MIN_NCHANNELS=16
