#!/bin/bash
O=gpurun_out/r02_cpp_matrix_inv_32_ngpu8.txt; : > $O
python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "sharded_entry_equals or singular" >> $O 2>&1
echo "== MATINV_NGPU=8 tools/time_matrix_inv_32 32768 2" >> $O
MATINV_NGPU=8 tools/time_matrix_inv_32 32768 2 2>&1 | grep -v "^NCCL version" >> $O
grep -E "passed|failed|matrix_inv_32|==" $O
