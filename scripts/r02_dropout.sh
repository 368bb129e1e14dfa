#!/bin/bash
# Dropout + L2-hint tests on one B200, then a default training bench line (the default kernels' code is unchanged)
O=gpurun_out/r02_dropout; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_dropout_layout.py tests/test_gpu_l2_hints.py tests/test_abi.py -q -m gpu > $O/pytest.log 2>&1; tail -25 $O/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --strong-4m off --no-cpu-baseline > $O/train.json 2> $O/train.err; python scripts/print_bench.py $O/train.json | head -3
