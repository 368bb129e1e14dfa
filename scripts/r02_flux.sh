#!/bin/bash
O=gpurun_out/r02_flux; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_glue.py tests/test_gpu_rollout.py -q -m gpu -x > $O/pytest.log 2>&1; tail -15 $O/pytest.log
timeout 200 python bench.py --workload flux_rollout_200k --steps 20 --warmup 5 --no-cpu-baseline > $O/flux.json 2> $O/flux.err; python scripts/print_bench.py $O/flux.json | head -1
