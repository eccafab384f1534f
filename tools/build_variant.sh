#!/bin/bash
# Build an experimental variant of libpihm_b200.so into build_exp/<name>/ (git-ignored; travels with gpurun).
# usage: tools/build_variant.sh <name> [extra nvcc flags...]; run with PIHM_B200_LIB=build_exp/<name>/libpihm_b200.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build_exp/$name; mkdir -p $out
src=$root/mm-pihm_b200/csrc
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-O2 -I$root/include $*"
for f in pihm_b200.cu cvode_b200.cu nvector_b200.cu comm.cu partition.cpp; do
  nvcc $FL -c -o $out/${f%.*}.o $src/$f &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libpihm_b200.so $out/*.o -lcudart -ldl
echo built $out/libpihm_b200.so
