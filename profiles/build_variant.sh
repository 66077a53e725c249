#!/bin/bash
# developer tool: build a kernel variant into gpurun_variants/<name>.so with extra nvcc -D flags
# usage: profiles/build_variant.sh name -DAT3D_MINB_FWD1=4 ...
set -e
name=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
out=$ROOT/gpurun_variants; mkdir -p $out/obj_$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -fopenmp -Xcompiler -ffp-contract=off"
for cu in $ROOT/at3d_b200/csrc/*.cu; do
  b=$(basename $cu .cu)
  nvcc $FLAGS "$@" -c $cu -o $out/obj_$name/$b.o 2>/dev/null &
done
wait
nvcc -shared -o $out/$name.so $out/obj_$name/*.o -lcudart -lgomp
rm -rf $out/obj_$name
echo built $out/$name.so
