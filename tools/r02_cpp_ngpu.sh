#!/bin/bash
# the C++ surface on several GPUs: matrix_inv_32(std::vector<float>, n) with MATINV_NGPU (verdict item 5's done-criterion)
NG=${1:-8}
O=gpurun_out/r02_cpp_matrix_inv_32_ngpu$NG.txt; : > $O
python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "sharded_entry_equals or cpp_caller" >> $O 2>&1
for ng in $NG 1; do
  echo "== MATINV_NGPU=$ng tools/time_matrix_inv_32 32768 2" >> $O
  MATINV_NGPU=$ng tools/time_matrix_inv_32 32768 2 2>&1 | grep -v "^NCCL version" >> $O
done
echo "== MATINV_NGPU=$NG tools/time_matrix_inv_32 16384 2" >> $O
MATINV_NGPU=$NG tools/time_matrix_inv_32 16384 2 2>&1 | grep -v "^NCCL version" >> $O
grep -E "passed|failed|matrix_inv_32|==" $O
