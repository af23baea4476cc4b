#!/bin/bash
# Round-2 multi-GPU evidence run (gpurun --gpus N -- 'bash tools/r02_multi_gpu_run.sh N'): everything lands in gpurun_out/.
N=${1:-8}
O=gpurun_out
mkdir -p $O
{ echo "== OpenCL ICD probe (BASELINE.md s.3: reference kernels under an OpenCL ICD if one is installed)"; ls -la /etc/OpenCL/vendors 2>&1; which clinfo 2>&1; ls /usr/lib/x86_64-linux-gnu | grep -i -E "opencl|pocl" 2>&1; python -c "import pyopencl" 2>&1 | tail -1; nvidia-smi -L; } > $O/r02_opencl_probe.txt 2>&1
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $O/r02_multi_g$N.txt 2>&1; tail -3 $O/r02_multi_g$N.txt
# single-process C-ABI driver (one host thread per GPU, NCCL inside libmatinv32.so), N=65536 synthetic, 1 warm + 2 timed
python - > $O/r02_cabi_sharded_g$N.json 2> $O/r02_cabi_sharded_g$N.err <<PY
import json, sys, time
sys.path.insert(0, ".")
import gpu_matrix_inversion_b200 as m
from oracle.gj_oracle import SEED_UNIFORM
n = 65536
res = []
for i in range(3):
    t0 = time.perf_counter()
    rc, piv, ms = m.sharded_synthetic(n, SEED_UNIFORM + n, "uniform", ngpu=$N)
    res.append({"rc": rc, "compute_ms": ms, "wall_s": time.perf_counter() - t0})
best = min(r["compute_ms"] for r in res[1:])
print(json.dumps({"entry": "matinv_sharded_synthetic_f32", "n": n, "ngpu": $N, "runs": res, "best_ms": best, "tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12,
                  "nccl_version": m.lib.matinv_nccl_version(), "note": "factorisation + column exchange, CUDA events on rank 0; first run includes ncclCommInitAll"}))
PY
tail -c 600 $O/r02_cabi_sharded_g$N.json; tail -3 $O/r02_cabi_sharded_g$N.err
# torchrun path: BASELINE config 5 as its own workload, then the default line (what the driver's scaling sweep runs)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload n65536 --steps 2 --warmup 1 > $O/r02_bench_sh65536_g$N.json 2> $O/r02_bench_sh65536_g$N.err; tail -c 1500 $O/r02_bench_sh65536_g$N.json; tail -2 $O/r02_bench_sh65536_g$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 > $O/r02_bench_default_g$N.json 2> $O/r02_bench_default_g$N.err; tail -c 1800 $O/r02_bench_default_g$N.json; tail -2 $O/r02_bench_default_g$N.err
