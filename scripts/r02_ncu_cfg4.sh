#!/bin/bash
# launch lists of the BASELINE configs[4] training steps (VertPotA / StreamFuncA, 8 x 20k-cell meshes)
O=gpurun_out/r02_ncu_cfg4; mkdir -p $O
for w in vertpot_train_8x20k streamfunc_train_8x20k; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/$w.csv python bench.py --workload $w --steps 2 --warmup 3 --strong-4m off --no-cpu-baseline > $O/$w.log 2>&1; tail -1 $O/$w.log | cut -c1-150
done
