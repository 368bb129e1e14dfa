set -x
O=gpurun_out/final; mkdir -p $O
timeout 200 python bench.py --steps 10 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
timeout 120 python bench.py --workload fvgn_fwd_8x20k --steps 20 --warmup 5 > $O/bench_fwd.json 2> $O/bench_fwd.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 200 python scripts/bench_kernels.py > $O/kernel_microbench.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/train_launches.csv python bench.py --steps 1 --warmup 1 > $O/ncu_list.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc|mlp_tc_kernel" -s 7 -c 7 -o $O/train_kernels -f python scripts/prof_train_kernels.py > $O/ncu_train.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 -o $O/fwd_edge -f python scripts/prof_fwd_edge.py > $O/ncu_fwd.log 2>&1
ls -la $O
