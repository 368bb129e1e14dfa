#!/bin/bash
O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > $O/pytest.log; cat $O/pytest.log
timeout 300 python scripts/prof_e2e.py > $O/e2e_phases.log 2>&1; cat $O/e2e_phases.log
timeout 200 python bench.py --workload flux_rollout_200k --steps 20 --warmup 3 > $O/bench_flux.json 2> $O/flux.err; python scripts/print_bench.py $O/bench_flux.json
timeout 200 python bench.py --workload cons_rollout_200k --steps 20 --warmup 3 > $O/bench_cons.json 2> $O/cons.err; python scripts/print_bench.py $O/bench_cons.json
