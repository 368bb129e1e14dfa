#!/bin/bash
# pooled epilogue (all 16 warps on one phase) vs the two-group schedule
O=gpurun_out/r02y; mkdir -p $O
NP=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_nopool.so
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -4 > $O/pytest_parity.log; cat $O/pytest_parity.log
if ! grep -q "passed" $O/pytest_parity.log || grep -q "failed\|error" $O/pytest_parity.log; then echo "PARITY FAILED - stopping"; exit 0; fi
for i in 1 2; do
echo "== pooled";  timeout 200 python scripts/abl_edge.py fast 2>&1 | tail -2
echo "== two groups"; GNNFD_LIB=$NP timeout 200 python scripts/abl_edge.py fast 2>&1 | tail -2
done
echo "== pooled";  timeout 300 python scripts/bench_kernels.py 2>&1 | sed -n 1,8p
echo "== two groups"; GNNFD_LIB=$NP timeout 300 python scripts/bench_kernels.py 2>&1 | sed -n 1,8p
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_pool_$i.json 2> $O/bench_pool_$i.err; echo "pooled: $(python scripts/print_bench.py $O/bench_pool_$i.json 2>/dev/null | head -2 | tr '\n' ' ')"
GNNFD_LIB=$NP timeout 300 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_grp_$i.json 2> $O/bench_grp_$i.err; echo "two groups: $(python scripts/print_bench.py $O/bench_grp_$i.json 2>/dev/null | head -2 | tr '\n' ' ')"
done
