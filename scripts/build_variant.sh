#!/bin/bash
# Build a probe variant of the library: scripts/build_variant.sh NAME "-DMCX_OPT_X ..."  ->  lib/libmcx_b200_NAME.so
# (tuning only; select it at run time with MCX_B200_LIB=...)
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../montecarlox.jl_b200/csrc"
out=build_trace/$name; mkdir -p $out
for f in mcx_api k_generic k_ising2d k_resident k_slab k_bc2d k_ising3d k_rows8 k_pt k_flat k_queue k_persist k_bits k_graph k_series k_bc3d; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC $flags -Xptxas -v -c $f.cu -o $out/$f.o 2> $out/$f.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libmcx_b200_$name.so $out/*.o
grep -A2 "k_ising2dILi0ELb0ELb0ELb1ELb1E" $out/k_ising2d.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name"
