#!/bin/bash
O=gpurun_out/r02w; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > $O/pytest.log; cat $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python scripts/print_bench.py $O/bench.json | head -2
for w in mgn_rollout_2k flux_rollout_200k; do
timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > $O/$w.json 2> $O/$w.err
echo "$w: $(python scripts/print_bench.py $O/$w.json 2>/dev/null | head -1)"
done
