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
  echo "== $name"; grep -E "compute_ms|trace\] rank 0" $O/r02_trace_$name.txt | tail -6
}
run ctas1 MATINV_MULTI_BCAST_CTAS=1 MATINV_MULTI_TRACE=1
run ctas2 MATINV_MULTI_BCAST_CTAS=2
run ctas4 MATINV_MULTI_BCAST_CTAS=4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload n65536 --steps 2 --warmup 1 > $O/r02_bench_sh65536_g8.json 2> $O/r02_bench_sh65536_g8.err; tail -c 1800 $O/r02_bench_sh65536_g8.json; tail -2 $O/r02_bench_sh65536_g8.err
