#!/bin/bash
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity.py tests/test_gpu_dist.py -q -m gpu -x 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
for rep in 1 2; do
echo "== packed"; timeout 200 python scripts/bench_kernels.py 2>&1 | sed -n 1,6p; timeout 200 python scripts/abl_edge.py packed 2>&1 | tail -1
echo "== previous"; GNNFD_LIB=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_prev.so timeout 200 python scripts/bench_kernels.py 2>&1 | sed -n 1,6p; GNNFD_LIB=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_prev.so timeout 200 python scripts/abl_edge.py prev 2>&1 | tail -1
done > $O/ab.log 2>&1; cat $O/ab.log
