#!/bin/bash
O=gpurun_out/r02_final; mkdir -p $O
timeout 200 python scripts/prof_train_kernels.py > $O/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad|mlp_tc_kernel" -s 6 -c 6 -o $O/train_kernels -f python scripts/prof_train_kernels.py > $O/ncu_train.log 2>&1; tail -1 $O/ncu_train.log
