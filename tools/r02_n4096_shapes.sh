#!/bin/bash
O=gpurun_out/r02_n4096_shapes2.txt; : > $O
for n in 1024 2048 4096; do
  for shape in "" 16x1x256; do
    echo -n "n=$n shape=${shape:-default}: " >> $O
    MATINV_SUBPANEL_SHAPE=$shape python tools/run_single.py $n 12 2>&1 | tail -1 >> $O
  done
done
python - >> $O 2>&1 <<PY
import os, sys
sys.path.insert(0, ".")
os.environ["MATINV_SUBPANEL_SHAPE"] = "16x1x256"
import numpy as np
import gpu_matrix_inversion_b200 as m
from oracle import gj_oracle as o
for n in (300, 1100, 2100):
    A = o.uniform(n); X, piv = m.invert(A, want_piv=True); Xo, po, io = o.invert_inplace(A)
    print("parity 16x1x256 n=%d:" % n, np.array_equal(piv, po) and np.array_equal(X.view(np.uint32), Xo.view(np.uint32)))
PY
cat $O
