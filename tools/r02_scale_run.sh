#!/bin/bash
# default bench line at N GPUs, as the driver's scaling sweep launches it
N=$1; O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_default_g$N.json 2> $O/r02_bench_default_g$N.err
tail -c 1200 $O/r02_bench_default_g$N.json; tail -3 $O/r02_bench_default_g$N.err
