#!/bin/bash
# A/B: rolled MMA-issuer loops (in-tree library) against the previous build (lib_abl/libgnnfd_A.so)
O=gpurun_out/r02_roll; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest.log 2>&1; tail -2 $O/pytest.log
A=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_A.so
for v in A B A B; do
  if [ $v = A ]; then export GNNFD_LIB=$A; else unset GNNFD_LIB; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --strong-4m off --no-cpu-baseline > $O/train_$v.json 2> $O/train_$v.err
  echo "lib=$v train: $(python scripts/print_bench.py $O/train_$v.json 2>/dev/null | head -3 | tr '\n' ' ' | cut -c1-330)"
done
for v in A B; do
  if [ $v = A ]; then export GNNFD_LIB=$A; else unset GNNFD_LIB; fi
  timeout 200 python bench.py --workload mgn_rollout_2k --steps 200 --warmup 20 --no-cpu-baseline > $O/mgn2k_$v.json 2>/dev/null; echo "lib=$v 2k: $(python scripts/print_bench.py $O/mgn2k_$v.json 2>/dev/null | head -1)"
done
