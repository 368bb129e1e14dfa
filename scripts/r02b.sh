#!/bin/bash
O=gpurun_out/r02b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -5 > $O/pytest_dist.log; cat $O/pytest_dist.log
GNNFD_LIB=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_prof.so timeout 300 python scripts/prof_roles.py > $O/roles.log 2>&1; cat $O/roles.log
timeout 200 python scripts/prof_node.py > $O/plain.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 -o $O/fwd_node_fast -f python scripts/prof_node.py > $O/ncu.log 2>&1
tail -3 $O/ncu.log
