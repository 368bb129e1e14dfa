#!/bin/bash
O=gpurun_out/r02i; mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/train_launches.csv python bench.py --steps 1 --warmup 3 --strong-4m off --no-cpu-baseline > $O/ncu_list.log 2>&1
timeout 200 python scripts/prof_train_kernels.py > $O/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc|mlp_tc_kernel" -s 7 -c 7 -o $O/train_kernels -f python scripts/prof_train_kernels.py > $O/ncu_train.log 2>&1
tail -2 $O/ncu_train.log
ls -la $O
