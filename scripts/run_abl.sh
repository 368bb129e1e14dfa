#!/bin/bash
# GPU box: time the ablation builds of the MLP kernel (scripts/abl_edge.py)
O=gpurun_out/abl; mkdir -p $O
timeout 300 python scripts/abl_edge.py base > $O/abl.log 2> $O/abl.err
for n in 1 2 3 4 5 6 7 8 9; do
  GNNFD_LIB=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_abl$n.so timeout 120 python scripts/abl_edge.py abl$n >> $O/abl.log 2>> $O/abl.err
done
cat $O/abl.log
