#!/bin/bash
# Round-end refresh of the measurements kept under profiles/ (run on the B200 box through gpurun):
#   gpurun --timeout 2400 -- 'bash scripts/refresh_profiles.sh r02'
TAG=${1:-r02}
O=gpurun_out/refresh_$TAG; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err; python scripts/print_bench.py $O/bench_default.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; head -c 300 $O/bench_reference_arm.json; echo
timeout 200 python bench.py --workload fvgn_fwd_8x20k --steps 20 --warmup 5 > $O/bench_fwd.json 2> $O/bench_fwd.err; python scripts/print_bench.py $O/bench_fwd.json
timeout 200 python scripts/bench_kernels.py > $O/kernel_microbench.log 2>&1
timeout 300 python scripts/abl_edge.py final > $O/fast_path_timing.log 2>&1; cat $O/fast_path_timing.log
# ncu passes (never a bench value): launch list of ONE steady-state training step (skip the setup + 3 warm-up steps), then
# --set full of the training kernels and of the inference edge / node blocks
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 1300 --csv --log-file $O/train_launches.csv python bench.py --steps 2 --warmup 3 --strong-4m off --no-cpu-baseline > $O/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad|mlp_tc_kernel" -s 6 -c 6 -o $O/train_kernels -f python scripts/prof_train_kernels.py > $O/ncu_train.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 -o $O/fwd_edge_fast -f python scripts/prof_fwd_edge.py fast > $O/ncu_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 -o $O/fwd_node_fast -f python scripts/prof_node.py > $O/ncu_node.log 2>&1
ls -la $O
