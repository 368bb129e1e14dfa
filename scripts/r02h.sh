#!/bin/bash
O=gpurun_out/r02h; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_dist.py tests/test_gpu_glue.py tests/test_graph_cache.py -q -m gpu -x 2>&1 | tail -4 > $O/pytest.log; cat $O/pytest.log
timeout 120 python bench.py --workload mgn_rollout_2k --steps 50 --warmup 5 > $O/bench_mgn_rollout_2k.json 2> $O/mgn2k.err; tail -2 $O/mgn2k.err; python scripts/print_bench.py $O/bench_mgn_rollout_2k.json
timeout 200 python bench.py --workload flux_rollout_200k --steps 10 --warmup 3 > $O/bench_flux_rollout_200k.json 2> $O/flux.err; python scripts/print_bench.py $O/bench_flux_rollout_200k.json
timeout 300 python bench.py --workload mgn_rollout_4m --steps 5 --warmup 3 > $O/bench_4m_1gpu.json 2> $O/4m.err; python scripts/print_bench.py $O/bench_4m_1gpu.json
