#!/bin/bash
# 8-GPU tuning run of the broadcast communicator (gpurun --gpus 8 -- 'bash tools/r02_multi_trace.sh')
O=gpurun_out; mkdir -p $O
run() {  # name, extra env...
  name=$1; shift
  env "$@" python - > $O/r02_trace_$name.txt 2>&1 <<PY
import sys, time
sys.path.insert(0, ".")
import gpu_matrix_inversion_b200 as m
from oracle.gj_oracle import SEED_UNIFORM
n = 65536
for i in range(3):
    rc, piv, ms = m.sharded_synthetic(n, SEED_UNIFORM + n, "uniform", ngpu=8)
    print("run", i, "rc", rc, "compute_ms (factor + exchange)", ms, flush=True)
PY
  echo "== $name"; grep -E "compute_ms|trace\] rank 0" $O/r02_trace_$name.txt | tail -4
}
run ctas3 MATINV_MULTI_BCAST_CTAS=3
run ctas6 MATINV_MULTI_BCAST_CTAS=6
run ctas8 MATINV_MULTI_BCAST_CTAS=8 MATINV_MULTI_TRACE=1
run ctas4 MATINV_MULTI_BCAST_CTAS=4 MATINV_MULTI_TRACE=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r02_bench_default_g8.json 2> $O/r02_bench_default_g8.err; tail -c 1500 $O/r02_bench_default_g8.json; tail -2 $O/r02_bench_default_g8.err
