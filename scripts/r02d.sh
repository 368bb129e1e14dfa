#!/bin/bash
O=gpurun_out/r02d; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err; tail -3 $O/bench_train.err; cat $O/bench_train.json
timeout 200 python bench.py --workload fvgn_fwd_8x20k --steps 20 --warmup 5 > $O/bench_fwd.json 2> $O/bench_fwd.err; tail -3 $O/bench_fwd.err; cat $O/bench_fwd.json
