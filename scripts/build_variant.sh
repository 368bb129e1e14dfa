#!/bin/bash
# Build a variant of the library (profiling counters / ablations) next to the product one:
#   scripts/build_variant.sh prof -DGNNFD_TC_PROF   ->  gnn_fluid_dynamics_b200/lib_abl/libgnnfd_prof.so  (use with GNNFD_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../gnn_fluid_dynamics_b200/csrc"
B=build_$name; mkdir -p $B ../lib_abl
for f in csr segsum mlp_f32 mlp_tc mlp_api wgrad_tc bwd mlp_bwd halo glue; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c $f.cu -o $B/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../lib_abl/libgnnfd_$name.so $B/*.o -lcudart_static -lpthread -ldl -lrt
ls -la ../lib_abl/libgnnfd_$name.so
