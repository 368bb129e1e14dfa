#!/bin/bash
# pooled epilogue for single-tile launches only: old library (HEAD) vs new (two-group code path restructured; pooled when <= 1 tile per CTA)
O=gpurun_out/r02z; mkdir -p $O
OLD=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_old.so
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rollout.py -q -m gpu -x 2>&1 | tail -3 > $O/pytest_parity.log; cat $O/pytest_parity.log
if ! grep -q "passed" $O/pytest_parity.log || grep -q "failed\|error" $O/pytest_parity.log; then echo "PARITY FAILED - stopping"; exit 0; fi
for i in 1 2; do
echo "== new"; timeout 200 python scripts/abl_edge.py fast 2>&1 | tail -2
echo "== old"; GNNFD_LIB=$OLD timeout 200 python scripts/abl_edge.py fast 2>&1 | tail -2
done
for i in 1 2; do
for v in new nopool old; do
E=""; [ $v = nopool ] && E="GNNFD_POOL=0"; [ $v = old ] && E="GNNFD_LIB=$OLD"
env $E timeout 300 python bench.py --workload mgn_rollout_2k --steps 50 --warmup 5 --no-cpu-baseline > $O/roll2k_${v}_$i.json 2> $O/roll2k_${v}_$i.err; echo "$v 2k rollout: $(python scripts/print_bench.py $O/roll2k_${v}_$i.json 2>/dev/null | head -1)"
done; done
for i in 1 2; do
for v in new old; do
E=""; [ $v = old ] && E="GNNFD_LIB=$OLD"
env $E timeout 300 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_${v}_$i.json 2> $O/bench_${v}_$i.err; echo "$v: $(python scripts/print_bench.py $O/bench_${v}_$i.json 2>/dev/null | head -2 | tr '\n' ' ')"
done; done
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
