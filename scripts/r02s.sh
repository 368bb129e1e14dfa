#!/bin/bash
O=gpurun_out/r02s; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity.py tests/test_graph_cache.py -q -m gpu -x 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err; tail -3 $O/bench_train.err; python scripts/print_bench.py $O/bench_train.json
timeout 600 python bench.py --workload vertpot_train_8x20k --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_vp.json 2> $O/bench_vp.err; python scripts/print_bench.py $O/bench_vp.json | head -1
timeout 600 python bench.py --workload streamfunc_train_8x20k --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_sf.json 2> $O/bench_sf.err; python scripts/print_bench.py $O/bench_sf.json | head -1
