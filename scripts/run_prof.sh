#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
GNNFD_LIB=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_prof.so timeout 300 python scripts/prof_roles.py > $O/roles.log 2>&1
cat $O/roles.log
timeout 200 python scripts/prof_fwd_edge.py fast > $O/plain.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 -o $O/fwd_edge_fast -f python scripts/prof_fwd_edge.py fast > $O/ncu.log 2>&1
tail -3 $O/ncu.log
