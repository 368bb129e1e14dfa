#!/bin/bash
O=gpurun_out/r02k; mkdir -p $O
timeout 600 python -m pytest tests/test_graph_cache.py -q -m gpu -x 2>&1 | tail -4 > $O/pytest.log; cat $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err; tail -3 $O/bench_train.err; python scripts/print_bench.py $O/bench_train.json
