#!/bin/bash
# 1-GPU experiment: sub-panel kernel shape / cluster size at small orders (the panel chain is the critical path there)
O=gpurun_out/r02_n4096_shapes.txt; : > $O
for n in 2048 4096 8192; do
  for shape in "" 16x2x512 16x4x512 16x4x256; do
    for ctas in "" 16; do
      [ -z "$shape" ] && [ -n "$ctas" ] && continue
      echo -n "n=$n shape=${shape:-default} ctas=${ctas:-min}: " >> $O
      MATINV_SUBPANEL_SHAPE=$shape MATINV_SUBPANEL_CTAS=$ctas python tools/run_single.py $n 12 2>&1 | tail -1 >> $O
    done
  done
done
cat $O
