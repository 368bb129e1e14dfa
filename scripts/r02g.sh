#!/bin/bash
O=gpurun_out/r02g; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rollout.py tests/test_gpu_dist.py -q -m gpu -x 2>&1 | tail -4 > $O/pytest.log; cat $O/pytest.log
timeout 300 python scripts/abl_edge.py base > $O/fast_timing.log 2>&1; cat $O/fast_timing.log
timeout 200 python bench.py --workload fvgn_fwd_8x20k --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_fwd.json 2> $O/bench_fwd.err; tail -3 $O/bench_fwd.err; python scripts/print_bench.py $O/bench_fwd.json
timeout 120 python bench.py --workload mgn_rollout_2k --steps 50 --warmup 5 > $O/bench_mgn_rollout_2k.json 2> $O/mgn2k.err; python scripts/print_bench.py $O/bench_mgn_rollout_2k.json
timeout 200 python bench.py --workload flux_rollout_200k --steps 10 --warmup 3 > $O/bench_flux_rollout_200k.json 2> $O/flux.err; python scripts/print_bench.py $O/bench_flux_rollout_200k.json
