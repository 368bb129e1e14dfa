#!/bin/bash
O=gpurun_out/r02_gather3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_glue.py tests/test_gpu_training.py tests/test_gpu_rollout.py -q -m gpu -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for r in 1 2; do for v in 0 1; do
GNNFD_GATHER3=$v timeout 300 python bench.py --workload vertpot_train_8x20k --steps 10 --warmup 3 --no-cpu-baseline > $O/vertpot_$v.json 2> $O/vertpot_$v.err; echo "gather3=$v vertpot $(python scripts/print_bench.py $O/vertpot_$v.json 2>/dev/null | head -1)"
done; done
