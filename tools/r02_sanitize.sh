#!/bin/bash
# compute-sanitizer memcheck over the kernels added / re-shaped in round 2, at the smallest sizes that reach them
O=gpurun_out/r02_sanitize_memcheck.txt
cat > /tmp/san_case.py <<PY
import os, sys
sys.path.insert(0, ".")
import numpy as np
import gpu_matrix_inversion_b200 as m
from oracle import gj_oracle as o
A = o.uniform(700)
X, piv = m.invert(A, want_piv=True)                         # subpanel_kernel<16,1,256>, look-ahead off (n < 1024)
Xo, po, io = o.invert_inplace(A)
assert np.array_equal(piv, po) and np.array_equal(X.view(np.uint32), Xo.view(np.uint32))
A2 = o.uniform(1300)
X2 = m.invert(A2)                                           # look-ahead schedule
Xs = m.invert_sharded(A2, ngpu=1)                           # shard primitives + pack / unpack kernels of gj_multi.cu
assert np.array_equal(Xs.view(np.uint32), X2.view(np.uint32))
B = o.batched(64, 0, 40)
Xb, ib = m.invert_batched(B)                                # MATINV_BATCHED from the environment
B32 = o.batched(32, 0, 40)
Xc, ic = m.invert_batched(B32)
assert int((ib != 0).sum()) == 0 and int((ic != 0).sum()) == 0
print("sanitize case ok, batched mode", os.environ.get("MATINV_BATCHED", "default"))
PY
python /tmp/san_case.py > $O 2>&1 && for mode in 3 4; do MATINV_BATCHED=$mode compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_case.py >> $O 2>&1; echo "memcheck MATINV_BATCHED=$mode exit code $?" >> $O; done
grep -E "ERROR SUMMARY|exit code|sanitize case" $O
