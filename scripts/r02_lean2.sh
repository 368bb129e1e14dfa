#!/bin/bash
O=gpurun_out/r02_lean2; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest.log 2>&1; tail -2 $O/pytest.log
for v in 0 1; do
  GNNFD_LEAN=$v timeout 300 python bench.py --steps 20 --warmup 5 --strong-4m off --no-cpu-baseline > $O/train_$v.json 2> $O/train_$v.err
  echo "lean=$v train: $(python scripts/print_bench.py $O/train_$v.json 2>/dev/null | head -2 | tr '\n' ' ')"
done
