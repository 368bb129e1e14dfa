#!/bin/bash
# Last measurement pass of round 2 (after the LEAN instantiations): tests, default bench + reference arm, launch list, ncu --set full
O=gpurun_out/r02_final3; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 200 python scripts/prof_train_kernels.py > $O/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad|mlp_tc_kernel" -s 6 -c 6 -o $O/train_kernels -f python scripts/prof_train_kernels.py > $O/ncu_train.log 2>&1; tail -1 $O/ncu_train.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3800 --csv --log-file $O/train_launches_all.csv python bench.py --steps 2 --warmup 3 --strong-4m off --no-cpu-baseline > $O/ncu_list.log 2>&1; tail -1 $O/ncu_list.log | cut -c1-120
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time; tail -3 $O/bench_default.time; python scripts/print_bench.py $O/bench_default.json
( time timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err ) 2> $O/bench_reference.time; cut -c1-200 $O/bench_reference_arm.json
for w in vertpot_train_8x20k streamfunc_train_8x20k; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w $(python scripts/print_bench.py $O/bench_$w.json 2>/dev/null | head -1)"; done
timeout 300 python scripts/bench_kernels.py > $O/kernel_microbench.log 2>&1; head -8 $O/kernel_microbench.log
