#!/bin/bash
# Round-2 final evidence on one B200 (gpurun -- 'bash tools/r02_final_1gpu.sh')
O=gpurun_out; mkdir -p $O
python -m pytest tests -q -m gpu > $O/r02_pytest_gpu_final.txt 2>&1; tail -4 $O/r02_pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; tail -2 $O/r02_smoke.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_default.json 2> $O/r02_bench_default.err; tail -c 600 $O/r02_bench_default.json; tail -2 $O/r02_bench_default.err
python bench.py --workload n4096 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_n4096.json 2>/dev/null; tail -c 300 $O/r02_bench_n4096.json
python bench.py --workload batched64 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_b64.json 2>/dev/null; tail -c 300 $O/r02_bench_b64.json
python tools/run_single.py 16384 2 > $O/plain_single.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_n16384.csv python tools/run_single.py 16384 1 > $O/ncu_single.log 2>&1; tail -2 $O/ncu_single.log; wc -l $O/r02_launches_n16384.csv
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; tail -c 700 $O/r02_bench_ref.json
