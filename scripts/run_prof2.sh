#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
GNNFD_LIB=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_prof.so timeout 300 python scripts/prof_roles.py > $O/roles.log 2>&1
cat $O/roles.log
timeout 600 python -m pytest tests/test_gpu_glue.py tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -15 > $O/pytest_glue.log
cat $O/pytest_glue.log
