#!/bin/bash
# Round-2 measurement pass on the B200 box (through gpurun): tests, bench lines, kernel timings, ncu passes.
O=gpurun_out/r02a; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python scripts/abl_edge.py fast > $O/fast_timing.log 2>&1; cat $O/fast_timing.log
timeout 200 python bench.py --steps 10 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err; cat $O/bench_train.json
timeout 120 python bench.py --workload fvgn_fwd_8x20k --steps 20 --warmup 5 > $O/bench_fwd.json 2> $O/bench_fwd.err; cat $O/bench_fwd.json
timeout 120 python bench.py --workload mgn_rollout_2k --steps 50 --warmup 5 > $O/bench_mgn_rollout_2k.json 2> $O/mgn2k.err; cat $O/bench_mgn_rollout_2k.json
timeout 200 python bench.py --workload flux_rollout_200k --steps 10 --warmup 3 > $O/bench_flux_rollout_200k.json 2> $O/flux.err; cat $O/bench_flux_rollout_200k.json
timeout 200 python bench.py --workload cons_rollout_200k --steps 10 --warmup 3 > $O/bench_cons_rollout_200k.json 2> $O/cons.err; cat $O/bench_cons_rollout_200k.json
timeout 200 python scripts/bench_kernels.py > $O/kernel_microbench.log 2>&1; cat $O/kernel_microbench.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/train_launches.csv python bench.py --steps 1 --warmup 1 > $O/ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/fwd_launches.csv python bench.py --workload fvgn_fwd_8x20k --steps 1 --warmup 1 > $O/ncu_list_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 -o $O/fwd_edge_fast -f python scripts/prof_fwd_edge.py fast > $O/ncu_fwd.log 2>&1
ls -la $O
