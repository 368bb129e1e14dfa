#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
for d in 0 1 2 3; do
  GNNFD_L1_DEPTH=$d timeout 200 python scripts/abl_edge.py depth$d 2>&1 | grep -v "^$" >> $O/depth.log
done
cat $O/depth.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
